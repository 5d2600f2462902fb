for v in "-DMG_MARCH_MINB=4 -DMG_ERROR_MINB=2" "-DMG_MARCH_MINB=5 -DMG_ERROR_MINB=3" "-DMG_MARCH_MINB=6 -DMG_ERROR_MINB=4"; do
  echo "=== $v"; MG_NVCC_EXTRA="$v" python -m saras_b200.build --force 2>&1 | grep -i error
  bash scripts/quick_bench.sh 10 jacobi 2>&1 | grep -E "GLUP|L0"
done
