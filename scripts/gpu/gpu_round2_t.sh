# round 2, call t (4 GPUs): which of today's changes breaks the natural-mean mgSolve on 2-D rank layouts
set -x
mkdir -p gpurun_out
export MP_CASES=nnn
export SARASMG_NO_PUSH=1
for v in nozeroskip nograph fenceall naturalfirst; do
  unset SARASMG_NO_ZEROSKIP SARASMG_GRAPH_MULTI SARASMG_LIB MP_NATURAL_FIRST
  if [ $v = nozeroskip ]; then export SARASMG_NO_ZEROSKIP=1; fi
  if [ $v = nograph ]; then export SARASMG_GRAPH_MULTI=0; fi
  if [ $v = fenceall ]; then export SARASMG_LIB=$PWD/saras_b200/lib/variants/libsarasmg_fenceall.so; fi
  if [ $v = naturalfirst ]; then export MP_NATURAL_FIRST=1; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/mp_worker.py gpu > gpurun_out/r2t_mp_gpu_4_$v.log 2>&1
  echo "== $v"; grep -a "mp_worker\|MP_WORKER\|AssertionError" gpurun_out/r2t_mp_gpu_4_$v.log | head -3
done
