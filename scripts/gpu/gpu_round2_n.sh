# round 2, call n (2 GPUs): A/B of the halo push in one box (clocks differ between boxes)
set -x
mkdir -p gpurun_out
for v in push nopush push nopush; do
  if [ $v = nopush ]; then export SARASMG_NO_PUSH=1; else unset SARASMG_NO_PUSH; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2n_bench_n2_$v.json 2> gpurun_out/r2n_bench_n2_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2n_bench_n2_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
P
done
