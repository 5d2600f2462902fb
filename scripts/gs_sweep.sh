for zq in 64 128 300; do
echo "1025 ZQ=$zq"; SARASMG_GS_ZQ=$zq timeout 300 python bench.py --index 10 --smoother gs --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(e['level'], e['op'], round(e['avg_ms'],3), round(e['frac'],3)) for e in d['per_level'] if e['op']=='smooth']"
done
