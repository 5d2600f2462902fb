# 2 GPUs, tight timeouts: parity of the default multi-rank path, then the opt-in graph path (create / solve / destroy)
N=2
set -x
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 tests/mp_worker.py gpu > gpurun_out/mp_gpu_${N}_final.log 2>&1
echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_${N}_final.log
grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_${N}_final.log | tail -18
MP_GRAPH_CASE=1 MP_CASES=ppn_tanh_graph timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 tests/mp_worker.py gpu > gpurun_out/mp_gpu_${N}_graphcase.log 2>&1
echo "mp_worker graph case rc=$?" >> gpurun_out/mp_gpu_${N}_graphcase.log
grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_${N}_graphcase.log | tail -6
