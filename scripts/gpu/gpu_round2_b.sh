# round 2, call b (1 GPU): first run of the warp-strip Gauss-Seidel kernel: parity on forced small geometries, then the 1025^3 sweep time
set -x
mkdir -p gpurun_out
export SARASMG_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gs_strip or smooth_bitexact" 2>&1 | tail -15 > gpurun_out/r2b_pytest_strip.log
cat gpurun_out/r2b_pytest_strip.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "c1 or c3 or c4" 2>&1 | tail -5 > gpurun_out/r2b_pytest_full.log
cat gpurun_out/r2b_pytest_full.log
B="--gpus 1 --steps 2 --warmup 1 --smoother gs --no-e2e --no-ttt --no-cpu-baseline"
SARASMG_VERBOSE=1 timeout 200 python bench.py $B > gpurun_out/r2b_gs_default.json 2> gpurun_out/r2b_gs_default.err
grep -a "strip GS" gpurun_out/r2b_gs_default.err | head -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_gs_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), [(e["level"], e["op"], round(e["avg_ms"],3), round(e["frac"],3)) for e in d["per_level"] if e["op"]=="smooth"])
    except Exception as e: print(f, "ERR", e)
PY
for v in "NQ=4" "NQ=6 SARASMG_GSS_DP=2" "NQ=8 SARASMG_GSS_DP=2" "NQ=3 SARASMG_GSS_DP=3"; do
  tag=$(echo $v | tr ' =' '__')
  env SARASMG_GSS_$v timeout 200 python bench.py $B > gpurun_out/r2b_gs_$tag.json 2> gpurun_out/r2b_gs_$tag.err
done
SARASMG_GS_STRIP=0 timeout 200 python bench.py $B > gpurun_out/r2b_gs_old.json 2> gpurun_out/r2b_gs_old.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_gs_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), [(e["level"], e["op"], round(e["avg_ms"],3), round(e["frac"],3)) for e in d["per_level"] if e["op"]=="smooth"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
