for m in "tiny.en" "base.en" "small"; do
for f in 1 0; do
echo "model $m fused=$f"
WSP_MEGA_FUSED=$f timeout 200 python bench.py --model $m --batch 8 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('  value %.0f  ms %.2f' % (j['value'], j['ms_per_step']), j['stage_ms_per_step'])"
done; done
