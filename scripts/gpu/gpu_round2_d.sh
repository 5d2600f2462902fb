# round 2, call d (1 GPU): strip GS kernel v2 (short step body): parity, sweep time, ncu capture
set -x
mkdir -p gpurun_out
export SARASMG_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gs_strip" 2>&1 | tail -5 > gpurun_out/r2d_pytest_strip.log
cat gpurun_out/r2d_pytest_strip.log
for idx in 9 10; do timeout 120 python scripts/gs_time.py $idx 3 2>&1 | grep GS_TIME; done > gpurun_out/r2d_gs_time.txt 2>&1
for nq in 3 4 6 8; do SARASMG_GSS_NQ=$nq timeout 120 python scripts/gs_time.py 10 3 2>&1 | grep GS_TIME | sed "s/^/nq=$nq /"; done >> gpurun_out/r2d_gs_time.txt 2>&1
cat gpurun_out/r2d_gs_time.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_strip -s 1 -c 1 -o gpurun_out/r2d_prof_strip_513 python scripts/gs_time.py 9 1 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
