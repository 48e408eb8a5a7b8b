"""CPU oracle (test infrastructure).  See oracle/muspin_oracle.py."""
