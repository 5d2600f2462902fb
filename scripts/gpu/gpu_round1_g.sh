# TMA v3 (two rows per thread): parity, bench, stage variants, ncu metrics at the bench size
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_g.log
cat gpurun_out/pytest_gpu_g.log
bash scripts/quick_bench.sh 10 jacobi > gpurun_out/qb_g_tma.txt 2>&1
for v in s4 s3; do
  echo "=== variant $v" >> gpurun_out/qb_g_variants.txt
  SARASMG_LIB=$PWD/saras_b200/lib/variants/libsarasmg_$v.so bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0" >> gpurun_out/qb_g_variants.txt
done
cat gpurun_out/qb_g_tma.txt gpurun_out/qb_g_variants.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stencil_tma' -c 4 \
    -o gpurun_out/prof_tma3 python scripts/prof_ops.py 9 1 > gpurun_out/prof_tma3.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_stencil_tma|k_prolong|k_coarsen|k_gs_skew' -c 12 \
    --csv --log-file gpurun_out/traffic_1025.csv python scripts/prof_ops.py 10 1 > gpurun_out/traffic_1025.log 2>&1
tail -2 gpurun_out/traffic_1025.log
