cd $GRAFT_REPO_ROOT
DVO_PARITY_REPORT=gpurun_out/parity_percentiles timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_vs_oracle" 2>&1 | tail -3
for pr in 1 3 4; do timeout 300 python bench.py --no-cpu --no-configs --steps 3 --prefetch-rows $pr 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prefetch', $pr, d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"; done
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import dense_visual_odometry_b200 as dvo
from dense_visual_odometry_b200.synthetic import make_sequence
from bench import camera_for
dev=torch.device('cuda',0)
s=make_sequence(1000, device=dev)
cam=camera_for(dvo, 640)
seq=dvo.SequenceAligner(cam,480,640,4,max_frames=1000,weights='tdist')
host=[torch.empty(x.shape,dtype=x.dtype).pin_memory() for x in (s['bgr'],s['depth'])]
host[0].copy_(s['bgr']); host[1].copy_(s['depth']); torch.cuda.synchronize()
for ch in (128,256,334,500,1000):
    seq.align(host[0],host[1],chunk_frames=ch)
    ts=[]
    for _ in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter(); seq.align(host[0],host[1],chunk_frames=ch); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    print('seq e2e chunk', ch, 999/np.median(ts))
t0=time.perf_counter(); seq.align(s['bgr'], s['depth'].clone()); torch.cuda.synchronize()
ts=[]
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); seq.align(s['bgr'], s['depth'].clone()); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
print('seq resident single launch', 999/np.median(ts))
PY
