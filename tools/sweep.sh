for cfg in "--prefetch 1 --prefetch-rows 2 --threads 128 --blocks-per-sm 3" "--prefetch 1 --prefetch-rows 2 --threads 384" "--prefetch 1 --prefetch-rows 2 --threads 512" "--prefetch 1 --prefetch-rows 2 --threads 256"; do
python bench.py --steps 2 --warmup 3 --pairs 592 --no-cpu --e2e-steps 1 $cfg 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"cfg $cfg\", \"value\", round(d[\"value\"]), \"kernel_ms\", round(d[\"roofline\"][\"kernel_ms\"],2), \"frac\", round(d[\"roofline\"][\"frac\"],4), d[\"accuracy\"][\"max_abs_twist_error_vs_truth\"])"
done
