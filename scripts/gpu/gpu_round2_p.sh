# round 2, call p (1 GPU): fused residual+restriction kernel: ncu capture (level 0 and 1), then bench A/B in one box
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resid_coarsen -c 2 -o gpurun_out/r2p_fused python bench.py --steps 1 --warmup 1 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2p_ncu.log 2>&1
tail -3 gpurun_out/r2p_ncu.log
for v in fused nofuse; do
  if [ $v = nofuse ]; then export SARASMG_NO_FUSE=1; else unset SARASMG_NO_FUSE; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2p_bench_$v.json 2> gpurun_out/r2p_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2p_bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
for r in d["per_level_all"]:
    if r["level"] <= 1 and r["op"] in ("residual", "coarsen", "smooth"): print("   ", r["level"], r["op"], r["launches"], round(r["avg_us"], 1))
P
done
