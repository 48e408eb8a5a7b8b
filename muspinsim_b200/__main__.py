"""`python -m muspinsim_b200 input.in [options]` -- the reference's own command line (muspinsim/__main__.py:37-168:
same arguments, input parser, fitting driver, log and `.dat` output files) with `ExperimentRunner.run` routed to
the CUDA path by `adapter.patch_reference()`.  Needs an installed `muspinsim`; this package does not re-implement
the CLI (out of scope), it only switches the hot loop underneath it.
`python -m muspinsim_b200 --mpi ...` calls the reference's MPI entry point (`muspinsim.mpi`) instead."""
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    use_mpi = False
    if argv and argv[0] == "--mpi":
        use_mpi = True
        argv = argv[1:]
    try:
        import muspinsim.__main__ as ref_cli
    except ImportError as exc:  # pragma: no cover - depends on the user's environment
        raise SystemExit("muspinsim_b200: the reference package `muspinsim` is not importable (%s); "
                         "use muspinsim_b200.ExperimentRunner(spec) directly or install muspinsim" % exc)
    from . import adapter

    adapter.patch_reference()
    old = sys.argv
    sys.argv = ["muspinsim"] + argv
    try:
        return ref_cli.main(use_mpi=use_mpi)
    finally:
        sys.argv = old
        adapter.unpatch_reference()


if __name__ == "__main__":
    sys.exit(main())
