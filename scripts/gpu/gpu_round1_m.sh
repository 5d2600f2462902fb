# CUDA-graph replay of the V-cycle: whole GPU suite (incl. full-size tests), bench with/without graphs at 1025^3 and 129^3, fp64 microbenchmark
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14 > gpurun_out/pytest_gpu_m.log
cat gpurun_out/pytest_gpu_m.log
bash scripts/quick_bench.sh 10 jacobi --no-ttt > gpurun_out/qb_m_graph.txt 2>&1
SARASMG_NO_GRAPH=1 bash scripts/quick_bench.sh 10 jacobi --no-ttt > gpurun_out/qb_m_nograph.txt 2>&1
head -3 gpurun_out/qb_m_graph.txt gpurun_out/qb_m_nograph.txt
for idx in 7 8; do
  bash scripts/quick_bench.sh $idx gs --no-ttt --cycles 25 2>&1 | head -1 > gpurun_out/qb_m_small_${idx}_graph.txt
  SARASMG_NO_GRAPH=1 bash scripts/quick_bench.sh $idx gs --no-ttt --cycles 25 2>&1 | head -1 > gpurun_out/qb_m_small_${idx}_nograph.txt
  cat gpurun_out/qb_m_small_${idx}_graph.txt gpurun_out/qb_m_small_${idx}_nograph.txt
done
./scripts/micro/fp64_lat > gpurun_out/fp64_lat.txt 2>&1; cat gpurun_out/fp64_lat.txt
