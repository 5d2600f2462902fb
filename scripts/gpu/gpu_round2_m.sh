# round 2, call m (2 GPUs): halo push of the Jacobi sweeps: parity worker, bench at N=2
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2m_mp_gpu_2.log 2>&1
tail -25 gpurun_out/r2m_mp_gpu_2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err
tail -c 300 gpurun_out/r2m_bench_n2.json; tail -3 gpurun_out/r2m_bench_n2.err
