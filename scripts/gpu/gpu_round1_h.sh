# fused preamble + prolong prefetch variants: parity, full default bench (recorded), GS bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_h.log
cat gpurun_out/pytest_gpu_h.log
QB_LINES=12 bash scripts/quick_bench.sh 10 jacobi --breakdown > gpurun_out/qb_h.txt 2>&1
for v in ppf2 ppf4; do
  echo "=== variant $v" >> gpurun_out/qb_h_variants.txt
  SARASMG_LIB=$PWD/saras_b200/lib/variants/libsarasmg_$v.so bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_h_variants.txt
done
cat gpurun_out/qb_h.txt gpurun_out/qb_h_variants.txt
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_default_h.json 2> gpurun_out/bench_default_h.err
tail -c 1500 gpurun_out/bench_default_h.json
