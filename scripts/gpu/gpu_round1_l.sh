# single GPU: whole GPU suite incl. the full-size tests, default bench (recorded), GS bench, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 > gpurun_out/pytest_gpu_l.log
cat gpurun_out/pytest_gpu_l.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_default_l.json 2> gpurun_out/bench_default_l.err
tail -c 2500 gpurun_out/bench_default_l.json
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --smoother gs --no-cpu-baseline > gpurun_out/bench_gs_l.json 2> gpurun_out/bench_gs_l.err
tail -c 600 gpurun_out/bench_gs_l.json
python -c "import __graft_entry__ as g; g.smoke()"
