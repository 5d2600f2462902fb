# round 2, call x (1 GPU): flattened thread mapping of the projection / implicit kernels: their tests, smoke, reduced bench
set -x
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_implicit.py tests/test_gpu_projection.py -x -q -m gpu > gpurun_out/r2x_pytest.log 2>&1; tail -2 gpurun_out/r2x_pytest.log
timeout 100 python bench.py --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; tail -c 200 gpurun_out/r2x_bench.err
