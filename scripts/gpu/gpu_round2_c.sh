# round 2, call c (1 GPU): why is the strip GS sweep slow?  timing experiments + one ncu --set full capture at 513^3
set -x
mkdir -p gpurun_out
V=saras_b200/lib/variants
for idx in 9 10; do
  timeout 120 python scripts/gs_time.py $idx 3 2>&1 | grep GS_TIME
  for v in nofence noywait nosleep; do
    SARASMG_LIB=$V/libsarasmg_$v.so timeout 120 python scripts/gs_time.py $idx 3 2>&1 | grep GS_TIME
  done
done > gpurun_out/r2c_gs_time.txt 2>&1
cat gpurun_out/r2c_gs_time.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_strip -s 1 -c 1 -o gpurun_out/r2c_prof_strip_513 python scripts/gs_time.py 9 1 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
