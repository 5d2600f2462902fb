# parity + quick bench of the power-of-two divisor path, the marching prolongation and error-kernel occupancy variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_b.log
bash scripts/quick_bench.sh 10 jacobi > gpurun_out/qb_b_base.txt 2>&1
for v in "-DMG_ERROR_MINB=3" "-DMG_ERROR_MINB=4"; do
  echo "=== $v" >> gpurun_out/qb_b_variants.txt
  MG_NVCC_EXTRA="$v" python -m saras_b200.build --force 2>&1 | grep -i error
  bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_b_variants.txt
done
cat gpurun_out/pytest_gpu_b.log gpurun_out/qb_b_base.txt gpurun_out/qb_b_variants.txt
