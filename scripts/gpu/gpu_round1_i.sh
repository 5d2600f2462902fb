# 4-GPU box: 2- and 4-rank parity of the NCCL halo path (2x1 and 2x2 pencils), then the strong-scaling bench at N = 4 and 2
set -x
mkdir -p gpurun_out
for N in 4 2; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N tests/mp_worker.py gpu > gpurun_out/mp_gpu_$N.log 2>&1
  echo "mp_worker rc=$?" >> gpurun_out/mp_gpu_$N.log
  grep -E "mp_worker|MP_WORKER|Error|rank0\]:" gpurun_out/mp_gpu_$N.log | tail -14
done
for N in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 2 --warmup 2 --breakdown --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  tail -1 gpurun_out/bench_n$N.json | cut -c1-700
  grep breakdown gpurun_out/bench_n$N.err | head -32
done
