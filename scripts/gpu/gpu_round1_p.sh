# final single-GPU pass of the round: whole GPU suite, default bench, launch list of the final kernels, GS bench
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_p.log
cat gpurun_out/pytest_gpu_p.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_default_p.json 2> gpurun_out/bench_default_p.err
tail -c 400 gpurun_out/bench_default_p.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --cycles 2 --no-e2e --no-cpu-baseline --no-ttt > gpurun_out/bench_under_ncu_p.log 2>&1
timeout 200 python bench.py --gpus 1 --steps 2 --warmup 3 --smoother gs --no-cpu-baseline > gpurun_out/bench_gs_p.json 2> gpurun_out/bench_gs_p.err
tail -c 300 gpurun_out/bench_gs_p.json
