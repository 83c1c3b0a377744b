python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "--threads 128"; do
python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 2 --pairs 2048 $cfg 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(\"cfg $cfg\", \"value\", round(d[\"value\"]), 'ms_per_step', round(d['ms_per_step'],2), \"kernel_ms\", round(d[\"roofline\"][\"kernel_ms\"],2), \"frac\", round(d[\"roofline\"][\"frac\"],4), 'e2e', round(d['e2e']['value']), d[\"accuracy\"][\"max_abs_twist_error_vs_truth\"])
"
done
