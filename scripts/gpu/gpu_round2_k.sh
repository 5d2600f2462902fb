# round 2, call k (2 GPUs): multi-rank path with CUDA-graph replay on by default: parity worker, both bench arms under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2k_mp_gpu_2.log 2>&1
tail -25 gpurun_out/r2k_mp_gpu_2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
tail -c 600 gpurun_out/r2k_bench_n2.json; tail -3 gpurun_out/r2k_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2k_bench_ref_n2.json 2> gpurun_out/r2k_bench_ref_n2.err
tail -c 400 gpurun_out/r2k_bench_ref_n2.json; tail -3 gpurun_out/r2k_bench_ref_n2.err
