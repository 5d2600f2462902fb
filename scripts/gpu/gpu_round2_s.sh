# round 2, call s (8 GPUs): parity worker (4x2 pencils, incl. projection step with Pp.syncData), bench with the halo push and without it
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2s_mp_gpu_8.log 2>&1
grep -a "mp_worker\|MP_WORKER\|rror" gpurun_out/r2s_mp_gpu_8.log | tail -20
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
tail -c 300 gpurun_out/r2s_bench_n8.json; tail -3 gpurun_out/r2s_bench_n8.err
export SARASMG_NO_PUSH=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-gs --no-ttt --no-e2e --no-cpu-baseline > gpurun_out/r2s_bench_n8_nopush.json 2> gpurun_out/r2s_bench_n8_nopush.err
tail -c 300 gpurun_out/r2s_bench_n8_nopush.json
