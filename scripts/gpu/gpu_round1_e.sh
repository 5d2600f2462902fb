# TMA-staged marching kernels: parity, then bench variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_e.log
cat gpurun_out/pytest_gpu_e.log
bash scripts/quick_bench.sh 10 jacobi > gpurun_out/qb_e_tma.txt 2>&1
SARASMG_NO_TMA=1 bash scripts/quick_bench.sh 10 jacobi > gpurun_out/qb_e_notma.txt 2>&1
for v in "-DMG_TMA_STAGES=3" "-DMG_TMA_STAGES=5"; do
  echo "=== $v" >> gpurun_out/qb_e_variants.txt
  MG_NVCC_EXTRA="$v" timeout 600 python -m saras_b200.build --force 2>&1 | grep -i error
  bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_e_variants.txt
done
for v in "-DMG_PF_DIST=2" "-DMG_PF_DIST=4"; do
  echo "=== $v (SARASMG_NO_TMA)" >> gpurun_out/qb_e_variants.txt
  MG_NVCC_EXTRA="$v" timeout 600 python -m saras_b200.build --force 2>&1 | grep -i error
  SARASMG_NO_TMA=1 bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_e_variants.txt
done
cat gpurun_out/qb_e_tma.txt gpurun_out/qb_e_notma.txt gpurun_out/qb_e_variants.txt
