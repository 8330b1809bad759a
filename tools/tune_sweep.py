"""Time the fused sweep (C-ABI through the Python mirror) for one kernel configuration and print a
digest of the volume so that configurations can be compared bit for bit across processes.
Usage: SATMVS_SWEEP_CFG=n python tools/tune_sweep.py [C D H W V per_pixel geo]"""
import hashlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import satmvs_b200
from satmvs_b200 import synth


def run(C=32, D=64, H=96, W=192, V=3, per_pixel=1, geo="rpc", n=40):
    fe = [f.cuda() for f in synth.make_features(1, V, C, H, W)]
    if geo == "rpc":
        cams = synth.make_rpc_stack(1, V, H, W)
        dv = synth.make_depth_planes(1, D, H, W, per_pixel=bool(per_pixel)).cuda()
    else:
        cams = synth.make_pinhole_stack(1, V, H, W)
        dv = (torch.linspace(0.9, 1.1, D) * 100.0).view(1, D).contiguous().cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    call = lambda: satmvs_b200.build_cost_volume(fe[0], fe[1:], cams[:, 0], cams[:, 1:], dv, geo)
    for _ in range(5):
        out = call()
    ev = []
    for _ in range(n):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); out = call(); e.record(); ev.append((s, e))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    dig = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:12]
    cells = D * H * W
    print(f"cfg={os.environ.get('SATMVS_SWEEP_CFG', '-')} v3={os.environ.get('SATMVS_SWEEP_V3', '-')} "
          f"C={C} D={D} {H}x{W} V={V} pp={per_pixel} {geo}: median {ts[n // 2] * 1e3:7.1f} us  min {ts[0] * 1e3:7.1f} us  "
          f"{ts[n // 2] * 1e6 / cells:6.3f} ns/cell  sha {dig}", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    if a:
        run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), a[6] if len(a) > 6 else "rpc")
    else:
        run()
