# multi-GPU pass (N = number of GPUs of the box): NCCL halo path parity against the oracle, then the strong-scaling bench
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mp_worker.py gpu > gpurun_out/mp_gpu_$N.log 2>&1
echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_$N.log
tail -25 gpurun_out/mp_gpu_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 2 --warmup 1 --breakdown > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
SARASMG_COARSE_LEGACY=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_n${N}_coarse_legacy.json 2> gpurun_out/bench_n${N}_coarse_legacy.err
tail -3 gpurun_out/bench_n$N.json | cut -c1-1500
grep breakdown gpurun_out/bench_n$N.err | head -45
cut -c1-400 gpurun_out/bench_n${N}_coarse_legacy.json
