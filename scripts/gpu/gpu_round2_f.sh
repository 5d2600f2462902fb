# round 2, call f (1 GPU): strip GS kernel v3 (no spills): parity, sweep time, ncu capture
set -x
mkdir -p gpurun_out
export SARASMG_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gs_strip" 2>&1 | tail -5 > gpurun_out/r2f_pytest_strip.log
cat gpurun_out/r2f_pytest_strip.log
V=saras_b200/lib/variants
{ SARASMG_VERBOSE=1 timeout 120 python scripts/gs_time.py 9 3 2>&1 | grep -a "GS_TIME\|strip GS on 513"
SARASMG_VERBOSE=1 timeout 120 python scripts/gs_time.py 10 3 2>&1 | grep -a "GS_TIME\|strip GS on 1025"
SARASMG_LIB=$V/libsarasmg_nowait.so timeout 120 python scripts/gs_time.py 10 3 2>&1 | grep GS_TIME
for nq in 2 3 4 8; do SARASMG_GSS_NQ=$nq SARASMG_VERBOSE=1 timeout 120 python scripts/gs_time.py 10 3 2>&1 | grep -a "GS_TIME\|strip GS on 1025" | sed "s/^/nq=$nq /"; done
} > gpurun_out/r2f_gs_time.txt 2>&1
cat gpurun_out/r2f_gs_time.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_strip -s 1 -c 1 -o gpurun_out/r2f_prof_strip_1025 python scripts/gs_time.py 10 1 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
