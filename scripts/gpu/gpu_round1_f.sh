# TMA v2 (mbarrier full/empty ring), fused face kernels, shared-memory coarse solve: parity, bench, variants, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_f.log
cat gpurun_out/pytest_gpu_f.log
bash scripts/quick_bench.sh 10 jacobi > gpurun_out/qb_f_tma.txt 2>&1
for v in s4 s6; do
  echo "=== variant $v" >> gpurun_out/qb_f_variants.txt
  SARASMG_LIB=$PWD/saras_b200/lib/variants/libsarasmg_$v.so bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_f_variants.txt
done
bash scripts/quick_bench.sh 10 gs > gpurun_out/qb_f_gs.txt 2>&1
QB_LINES=30 bash scripts/quick_bench.sh 9 jacobi --breakdown > gpurun_out/qb_f_idx9.txt 2>&1
cat gpurun_out/qb_f_tma.txt gpurun_out/qb_f_variants.txt gpurun_out/qb_f_gs.txt gpurun_out/qb_f_idx9.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stencil_tma|k_prolong|k_coarsen' -c 8 \
    -o gpurun_out/prof_tma python scripts/prof_ops.py 9 1 > gpurun_out/prof_tma.log 2>&1
tail -3 gpurun_out/prof_tma.log
