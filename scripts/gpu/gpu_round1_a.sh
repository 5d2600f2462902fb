# first GPU pass of a session: parity tests, default bench, launch list, full ncu capture of the level-0 operators
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 1 --smoother gs --no-cpu-baseline > gpurun_out/bench_gs.json 2> gpurun_out/bench_gs.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --cycles 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stencil|k_error|k_prolong|k_coarsen|k_gs_skew' -c 12 \
    -o gpurun_out/prof_ops python scripts/prof_ops.py 9 1 > gpurun_out/prof_ops.log 2>&1
ls -la gpurun_out
