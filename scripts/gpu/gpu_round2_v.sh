# round 2, call v (1 GPU, the last GPU minutes of the round): new GPU tests (implicit solves, projection, fused pass), default bench line
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_implicit.py tests/test_gpu_projection.py tests/test_gpu_parity.py -x -q -m gpu -k "implicit or velocity_bcs or predictor or projection or divergence or correction or fused" > gpurun_out/r2v_pytest.log 2>&1
tail -6 gpurun_out/r2v_pytest.log
timeout 330 python bench.py > gpurun_out/r2v_bench_default.json 2> gpurun_out/r2v_bench_default.err
tail -c 1500 gpurun_out/r2v_bench_default.json; tail -3 gpurun_out/r2v_bench_default.err
