# round 2, call y (1 GPU, the round's last GPU seconds): exact-reciprocal divisions in the implicit kernels: tests, smoke, the implicit leg's timing
set -x
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_implicit.py -x -q -m gpu > gpurun_out/r2y_pytest.log 2>&1; tail -2 gpurun_out/r2y_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; tail -1 gpurun_out/r2y_smoke.log
timeout 90 python bench.py --no-gs --no-ttt --no-e2e --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 200 gpurun_out/r2y_bench.err
