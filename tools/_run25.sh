cd $GRAFT_REPO_ROOT
for N in 8 4; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_n$N.json'))
print($N, {k:d[k] for k in ['value','ms_per_step','scaling']}, d['config']['pairs_per_gpu'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['h2d_ceiling'], 'weak', d['weak_scaling'])
PY
done
nvidia-smi topo -m 2>&1 | head -12
