timeout 900 python -m pytest tests -m gpu -q -x -k "tc or hotpath or optim or loss or local" 2>&1 | tail -2
for v in 0 1 0 1; do SPF_WGRAD_SIDE=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('side', $v, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['config']['cuda_graph'], d['config']['cuda_graph_note'], d['e2e']['last_loss'])
"; done
