# round 2, call u (4 GPUs): parity worker on 2x2 pencils after the stale-pad fix (replayed graphs must mark level-0 pads stale)
set -x
mkdir -p gpurun_out
export MP_CASES=nnn,ppp_tanhz,ppn_tanh,nnn_push_tma,ppn_tanh_nopush
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2u_mp_gpu_4.log 2>&1
grep -a "mp_worker:\|MP_WORKER\|AssertionError" gpurun_out/r2u_mp_gpu_4.log | head -12
