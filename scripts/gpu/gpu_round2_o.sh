# round 2, call o (1 GPU): fused residual+restriction: whole GPU suite, then bench A/B (fused vs SARASMG_NO_FUSE) in one box
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2o_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2o_pytest_gpu.log
for v in fused nofuse; do
  if [ $v = nofuse ]; then export SARASMG_NO_FUSE=1; else unset SARASMG_NO_FUSE; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2o_bench_$v.json 2> gpurun_out/r2o_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2o_bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
for r in d["per_level_all"]:
    if r["level"] <= 1 and r["op"] in ("residual", "coarsen", "smooth"): print("   ", r["level"], r["op"], r["launches"], round(r["avg_us"], 1))
P
done
