N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mp_worker.py gpu > gpurun_out/mp_gpu_$N.log 2>&1
echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_$N.log
grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_$N.log | tail -30
