python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import dense_visual_odometry_b200 as m
from dense_visual_odometry_b200.synthetic import make_pairs_torch, TUM_FR1, TUM_DEPTH_SCALE
B=2048
dev=torch.device('cuda',0)
Km=np.array([[TUM_FR1[0],0,TUM_FR1[2]],[0,TUM_FR1[1],TUM_FR1[3]],[0,0,1]],dtype=np.float32)
cam=m.RGBDCameraModel(Km,TUM_DEPTH_SCALE)
d=make_pairs_torch(range(B),dev)
hb=[torch.empty(d[k].shape,dtype=d[k].dtype).pin_memory() for k in ("bgr_prev","depth_prev","bgr_cur","depth_cur")]
for h,k in zip(hb,("bgr_prev","depth_prev","bgr_cur","depth_cur")): h.copy_(d[k])
# raw H2D ceiling
dst=[torch.empty_like(d[k]) for k in ("bgr_prev","depth_prev","bgr_cur","depth_cur")]
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(3):
    for a,b in zip(dst,hb): a.copy_(b, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
nb=sum(h.numel()*h.element_size() for h in hb)
print('raw pinned H2D GB/s', round(nb/dt/1e9,1), '-> ceiling pose/s', round(B/dt))
al=m.PairBatchAligner(cam,480,640,4,max_pairs=B)
ref=None
for chunk in (128,256,512):
    q,_=al.align(*hb,chunk_pairs=chunk)
    if ref is None: ref=q
    t0=time.perf_counter()
    for _ in range(3): al.align(*hb,chunk_pairs=chunk)
    dt=(time.perf_counter()-t0)/3
    print('chunk',chunk,'e2e pose/s',round(B/dt),'GB/s',round(nb/dt/1e9,1), 'same', bool(np.array_equal(q,ref)))
PY
