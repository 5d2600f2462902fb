# round 2, call j (1 GPU): strip GS: lighter CTA fence; single-warp latency vs warps per SM (no-wait timing variants)
set -x
mkdir -p gpurun_out
export SARASMG_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gs_strip" 2>&1 | tail -3
V=saras_b200/lib/variants
{ SARASMG_GSS_NQ=6 SARASMG_GSS_DP=2 SARASMG_VERBOSE=1 timeout 120 python scripts/gs_time.py 10 3 2>&1 | grep -a "GS_TIME\|strip GS on 1025"
for wc in 1 2 4 8 11; do SARASMG_GSS_NQ=6 SARASMG_GSS_DP=2 SARASMG_GSS_WC=$wc SARASMG_VERBOSE=1 SARASMG_LIB=$V/libsarasmg_nowait.so timeout 200 python scripts/gs_time.py 10 2 2>&1 | grep -a "GS_TIME\|strip GS on 1025" | sed "s/^/nowait wc=$wc /"; done
for wc in 1 11; do SARASMG_GSS_NQ=6 SARASMG_GSS_DP=2 SARASMG_GSS_WC=$wc SARASMG_LIB=$V/libsarasmg_nowait_nofence.so timeout 200 python scripts/gs_time.py 10 2 2>&1 | grep -a "GS_TIME" | sed "s/^/nowait_nofence wc=$wc /"; done
} > gpurun_out/r2j_gs_time.txt 2>&1
cat gpurun_out/r2j_gs_time.txt
