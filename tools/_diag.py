import sys, torch
sys.path.insert(0,'.')
import satmvs_b200
from satmvs_b200 import synth
def t(C, D, H=96, W=192, V=3, per_pixel=True, n=30):
    fe=[f.cuda() for f in synth.make_features(1,V,C,H,W)]
    rp=synth.make_rpc_stack(1,V,H,W); dv=synth.make_depth_planes(1,D,H,W,per_pixel=per_pixel).cuda()
    flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
    for _ in range(5): satmvs_b200.build_cost_volume(fe[0],fe[1:],rp[:,0],rp[:,1:],dv,"rpc")
    ev=[]
    for _ in range(n):
        flush.zero_(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record(); satmvs_b200.build_cost_volume(fe[0],fe[1:],rp[:,0],rp[:,1:],dv,"rpc"); e.record(); ev.append((s,e))
    torch.cuda.synchronize()
    ms=sorted(a.elapsed_time(b) for a,b in ev)[n//2]
    cells=D*H*W
    print(f"C={C:3d} D={D:3d} {H}x{W} V={V} per_pixel={per_pixel}: {ms*1e3:7.1f} us  {ms*1e6/cells:6.3f} ns/cell")
t(4,64); t(8,64); t(16,64); t(32,64); t(32,64,per_pixel=False); t(32,64,V=2); t(32,64,V=5)
t(16,32,192,384); t(8,8,384,768); t(32,192,192,384,V=5,n=10)
