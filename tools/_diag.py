import sys, os, torch, ctypes as C
sys.path.insert(0,'.')
os.environ["SATMVS_RED_DEBUG_TIMERS"]="1"
if len(sys.argv)>1: os.environ["SATMVS_RED_DEBUG_BLOCK"]=sys.argv[1]
import satmvs_b200
from satmvs_b200 import synth, _lib, module
Cc,D,H,W=32,64,96,192
m=satmvs_b200.RED_Regularization(Cc,8); m.load_state_dict(synth.make_red_weights(Cc)); m=m.cuda().eval()
x=torch.rand(1,Cc,D,H,W,device='cuda')
for _ in range(3): y=m(x)
torch.cuda.synchronize()
nbytes=_lib.lib().satmvs_red_workspace_bytes(Cc,D,H,W)
ws=module._workspace(nbytes, x.device)
# debug counters live right after the stats block, which is the last region: 64 doubles before the 256B-rounded end
tail=ws[:nbytes].view(torch.int64)
stats_doubles=D*4*3*2
# find: last region starts at nbytes - roundup((stats_doubles+64)*8,256)
reg=((stats_doubles+64)*8+255)//256*256
base=(nbytes-reg)//8
dbg=tail[base+stats_doubles: base+stats_doubles+8].cpu().tolist()
names=["P1 conv","sync","E1","sync","P2 conv","sync","E2","sync"]
tot=sum(dbg)
for n,v in zip(names,dbg): print(f"{n:8s} {v/1e3/D:8.2f} us/plane  {100*v/tot:5.1f}%")
print("total per plane", tot/1e3/D, "us")
u=tail[base+stats_doubles+8: base+stats_doubles+13].cpu().tolist()
for n,v in zip(["stage+mask","main loop","pre+part sync","epilogue","stats atomics"],u): print(f"  unit {n:14s} {v/1e3/D:8.2f} us/plane")
