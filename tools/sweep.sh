python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "--threads 128" "--threads 128 --approximate-gradient" "--threads 256"; do
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --pairs 2048 $cfg 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(\"cfg $cfg\", \"value\", round(d[\"value\"]), \"kernel_ms\", round(d[\"roofline\"][\"kernel_ms\"],2), \"frac\", round(d[\"roofline\"][\"frac\"],4), 'e2e', round(d['e2e']['value']), d[\"accuracy\"][\"max_abs_twist_error_vs_truth\"], d['e2e']['matches_resident'])
"
done
