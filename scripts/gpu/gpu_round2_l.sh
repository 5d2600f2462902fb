# round 2, call l (8 GPUs): 8-GPU parity worker (4x2 pencils), bench at N=8, projection-step tests on one GPU
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_projection.py -x -q -m gpu > gpurun_out/r2l_pytest_proj.log 2>&1
tail -5 gpurun_out/r2l_pytest_proj.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2l_mp_gpu_8.log 2>&1
tail -25 gpurun_out/r2l_mp_gpu_8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
tail -c 600 gpurun_out/r2l_bench_n8.json; tail -3 gpurun_out/r2l_bench_n8.err
