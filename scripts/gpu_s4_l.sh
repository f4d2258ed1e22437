bash scripts/gpu_s4_c.sh ffma2
for P in 16000000 32000000; do
python bench.py --no-cpu --steps 3 --warmup 3 --e2e-portion $P 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('portion $P e2e', round(d['e2e']['ms_per_step'],1), {k:round(v,1) for k,v in d['e2e']['phase_ms'].items()})"
done
