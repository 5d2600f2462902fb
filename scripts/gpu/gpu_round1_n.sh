# N GPUs: graph replay + refactored peer-memory exchange + table-driven coarsest solve
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "solve_coarsest or mg_solve or vcycle or graph or dropin or c1_ or c2_" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/mp_worker.py gpu > gpurun_out/mp_gpu_${N}_graph.log 2>&1
echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_${N}_graph.log
grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_${N}_graph.log | tail -18
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ttt > gpurun_out/bench_n${N}_graph.json 2> gpurun_out/bench_n${N}_graph.err
tail -1 gpurun_out/bench_n${N}_graph.json | cut -c1-260; tail -3 gpurun_out/bench_n${N}_graph.err
SARASMG_NO_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ttt > gpurun_out/bench_n${N}_nograph.json 2> gpurun_out/bench_n${N}_nograph.err
tail -1 gpurun_out/bench_n${N}_nograph.json | cut -c1-260
