# usage: bash scripts/quick_bench.sh <index> <smoother> [extra bench args]
idx=$1; sm=$2; shift; shift
timeout 600 python bench.py --index $idx --smoother $sm --steps 2 --warmup 1 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/qb.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$sm idx$idx GLUP/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), d['config']['coarse_iters']); [print('  L%d %-9s %8.3f ms  %5.1f%% of peak' % (e['level'], e['op'], e['avg_ms'], 100*e['frac'])) for e in d['per_level']]"
grep breakdown gpurun_out/qb.err | head -${QB_LINES:-0}
