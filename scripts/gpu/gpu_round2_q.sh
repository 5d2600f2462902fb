# round 2, call q (1 GPU): fused residual+restriction with 8-row tiles (2 CTAs/SM) vs 16-row variant vs unfused, one box
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "vcycle or mg_solve or c5 or c4" > gpurun_out/r2q_pytest.log 2>&1
tail -4 gpurun_out/r2q_pytest.log
for v in fy8 fy16 nofuse; do
  unset SARASMG_NO_FUSE SARASMG_LIB
  if [ $v = nofuse ]; then export SARASMG_NO_FUSE=1; fi
  if [ $v = fy16 ]; then export SARASMG_LIB=$PWD/saras_b200/lib/variants/libsarasmg_fy16.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2q_bench_$v.json 2> gpurun_out/r2q_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2q_bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
for r in d["per_level_all"]:
    if r["level"] <= 2 and r["op"] in ("residual", "coarsen", "smooth"): print("   ", r["level"], r["op"], r["launches"], round(r["avg_us"], 1))
P
done
