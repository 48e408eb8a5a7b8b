# round 2, GPU call I: Celio's method + full GPU test suite
set -x
python -m pytest tests/test_gpu_celio.py -m gpu -q -x > gpurun_out/r2i_tests_celio.log 2>&1; tail -25 gpurun_out/r2i_tests_celio.log
python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; tail -8 gpurun_out/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -8 gpurun_out/r2i_smoke.log
