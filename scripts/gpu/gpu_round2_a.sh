# round 2, call a (1 GPU): whole GPU suite after the boundary / mean / bench changes, default bench line
set -x
mkdir -p gpurun_out
export SARASMG_REQUIRE_GPU=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r2a_pytest_gpu.log
cat gpurun_out/r2a_pytest_gpu.log
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
tail -c 1500 gpurun_out/r2a_bench_default.json; tail -5 gpurun_out/r2a_bench_default.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
cat gpurun_out/r2a_bench_reference.json; nproc
