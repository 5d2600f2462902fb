# round 2, call w (1 GPU, last GPU minutes): smoke(), GPU suite without the full-size file (run green earlier in the round), default bench
set -x
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -2 gpurun_out/r2w_smoke.log
timeout 200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
timeout 200 python bench.py > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err; tail -c 400 gpurun_out/r2w_bench_default.json; tail -2 gpurun_out/r2w_bench_default.err
