# IPC peer-memory halos on N GPUs: parity (both halo paths), bench with both paths
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mp_worker.py gpu > gpurun_out/mp_gpu_${N}_ipc.log 2>&1
echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_${N}_ipc.log
grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_${N}_ipc.log | tail -16
SARASMG_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 2 --warmup 2 --breakdown --no-cpu-baseline --no-e2e > gpurun_out/bench_n${N}_ipc.json 2> gpurun_out/bench_n${N}_ipc.err
tail -1 gpurun_out/bench_n${N}_ipc.json | cut -c1-300
grep -E "breakdown (halo|bc|coarse|total)|IPC" gpurun_out/bench_n${N}_ipc.err | head -24
SARASMG_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
tail -1 gpurun_out/bench_n${N}_nccl.json | cut -c1-300
