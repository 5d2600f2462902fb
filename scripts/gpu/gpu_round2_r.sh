# round 2, call r (2 GPUs): multi-rank projection step (Pp.syncData over grouped ncclSend/ncclRecv) in the parity worker
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2r_mp_gpu_2.log 2>&1
tail -25 gpurun_out/r2r_mp_gpu_2.log
