#!/usr/bin/env python
"""bench.py — the hot path's headline metric on B200: depth-hypothesis voxels/s (B·V·D·H·W / t).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one pass of the hot path over one synthetic 3-view 768x384 stack at BASELINE.json
configs[1] (cost-volume resolution 96x192, C=32, 64 planes).  With N>1 (torchrun, one rank per GPU)
every rank sweeps its own stack: independent units, no data-path collective, weak scaling.

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB write) between
iterations outside the timed region, barrier + synchronize on both sides, max over ranks.
`e2e` repeats the measurement through the public operator API with pinned HOST inputs, the
host->device copies and the device->host read of the result inside the timed region.
`--impl reference` times the CPU oracle (a port of the reference's torch-CPU path; the reference is
pure Python and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth-hypothesis-voxels/sec (BxVxDxHxW)"
UNIT = "voxels/s"
WORKLOADS = {
    # BASELINE.json configs[1]: 3-view 768x384 image, casred stage 1 (scale 4 -> 96x192 features, C=32), 64 planes
    "cfg2_build": dict(B=1, V=3, C=32, D=64, H=96, W=192, geo="rpc", stage="build",
                       desc="3-view 768x384, 64 planes, casred stage-1 cost-volume build (fused RPC warp + variance)"),
    "cfg2_stage1": dict(B=1, V=3, C=32, D=64, H=96, W=192, geo="rpc", stage="red_train",
                        desc="3-view 768x384, 64 planes, casred stage-1 only: cost volume + RED regulariser + soft-argmin"),
    "cfg5_build": dict(B=1, V=3, C=32, D=64, H=96, W=192, geo="pinhole", stage="build",
                       desc="pin-hole homography sweep 3-view 768x384, 64 planes, cost-volume build"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_inputs(w, seed=0):
    from satmvs_b200 import synth
    fe = synth.make_features(w["B"], w["V"], w["C"], w["H"], w["W"], seed=seed)
    if w["geo"] == "rpc":
        cams = synth.make_rpc_stack(w["B"], w["V"], w["H"], w["W"])
        dv = synth.make_depth_planes(w["B"], w["D"], w["H"], w["W"])
    else:
        cams = synth.make_pinhole_stack(w["B"], w["V"], w["H"], w["W"])
        dv = synth.make_depth_planes(w["B"], w["D"], w["H"], w["W"], lo=90, hi=110, jitter=0.2)
    return fe, cams, dv


def algorithmic_bytes_per_cell(w):
    # variance write 4C + per-pixel hypothesis read 4 + compulsory feature reads 4·C·V/D
    return 4 * w["C"] + 4 + 4.0 * w["C"] * w["V"] / w["D"]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mx = max(int(float(r[2])) for r in self.rows if len(r) >= 9)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 9 and r[5 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_step(w, dev, fe, cams, dv):
    """Returns (step_fn() -> result tensors, kernel launches per step, name of the dominant kernel)."""
    import satmvs_b200
    ref, srcs = fe[0], fe[1:]
    ref_cam, src_cams = cams[:, 0], [cams[:, v] for v in range(1, w["V"])]
    if w["stage"] == "build":
        def step():
            return (satmvs_b200.build_cost_volume(ref, srcs, ref_cam, src_cams, dv, w["geo"]),)
        return step, w["B"]
    if w["stage"] == "red_train":
        from satmvs_b200 import synth
        reg = satmvs_b200.RED_Regularization(w["C"], 8)
        reg.load_state_dict(synth.make_red_weights(w["C"]))
        reg = reg.to(dev).eval()
        all_cams = cams

        def step():
            with torch.no_grad():
                out = satmvs_b200.stage_train_red(fe, all_cams, dv, reg, w["geo"])
            return out["depth"], out["photometric_confidence"]
        # launches per step: 1 sweep + 3 encoders + 8 x-halves + 4 per plane + 3 decoder + 1 head conv + 1 soft-argmin
        return step, w["B"] * (1 + 3 + 8 + 4 * w["D"] + 3 + 1 + 1)
    raise ValueError(w["stage"])


def run_ours(args, w):
    import satmvs_b200  # noqa: F401  (fails loudly if the CUDA library is missing)
    from satmvs_b200 import _lib
    _lib.lib()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fe_h, cams, dv_h = make_inputs(w, seed=rank)
    fe = [f.to(dev) for f in fe_h]
    dv = dv_h.to(dev)
    step, launches = build_step(w, dev, fe, cams, dv)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    voxels = w["B"] * w["V"] * w["D"] * w["H"] * w["W"]
    cells = w["B"] * w["D"] * w["H"] * w["W"]

    def timed(fn, n):
        evs = []
        for _ in range(n):
            flush.zero_()                       # evict L2 between iterations (outside the timed events)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in evs]

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    with ClockSampler(local) as clk:
        times = timed(step, args.steps)
        barrier()
    total_ms = sum(times)

    # end to end: pinned host inputs -> H2D -> operator API -> D2H of the result, all inside the events
    fe_p = [f.pin_memory() for f in fe_h]
    dv_p = dv_h.pin_memory()
    h2d = sum(f.numel() * 4 for f in fe_p) + dv_p.numel() * 4 + cams.numel() * 8
    outs_host = None

    def e2e_step():
        nonlocal outs_host
        f_d = [f.to(dev, non_blocking=True) for f in fe_p]
        d_d = dv_p.to(dev, non_blocking=True)
        st, _ = build_step(w, dev, f_d, cams, d_d)
        res = st()
        if outs_host is None:
            outs_host = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res]
        for h, r in zip(outs_host, res):
            h.copy_(r, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_times = timed(e2e_step, args.steps)
    barrier()
    e2e_ms = sum(e2e_times)
    d2h = sum(h.numel() * h.element_size() for h in outs_host)

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()

    if rank == 0:
        hbm, peak_src = peaks()
        ms = total_ms / args.steps
        line = {
            "metric": METRIC, "value": voxels * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (geometry f64)", "data": "synthetic",
            "config": {"workload": w["desc"], "name": args.workload, "B": w["B"], "V": w["V"], "C": w["C"], "D": w["D"],
                       "H": w["H"], "W": w["W"], "geo_model": w["geo"], "hypotheses": "per-pixel [B,D,H,W]",
                       "l2": "flushed (256 MiB write) between timed iterations", "sharding": "one stack per rank"},
            "clocks": clk.summary(),
            "e2e": {"value": voxels * world / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": (launches if launches is not None else getattr(sys.modules["satmvs_b200"], "last_launch_count", lambda: 0)()) * args.steps,
        }
        if w["stage"] == "build":
            bpc = algorithmic_bytes_per_cell(w)
            ach = cells * bpc / (ms * 1e-3) / 1e9
            line["roofline"] = {"kernel": "sweep_fwd_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                                "frac": ach / hbm, "traffic": None, "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": cells * bpc, "bytes_per_cell": bpc,
                                "avg_launch_us": ms * 1e3}
        if world == 1 and not args.no_cpu_baseline:
            line["gpu_eager_baseline"] = gpu_eager_baseline(w, dev, fe, cams, dv)
            line["cpu_baseline"] = cpu_baseline(w, budget_s=20.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's torch-CPU path
# ------------------------------------------------------------------------------------------------
def cpu_sample(w, planes):
    """One bounded sample: the same workload restricted to `planes` depth planes."""
    from oracle import volume
    ws = dict(w, D=planes)
    fe, cams, dv = make_inputs(ws)

    def run():
        with torch.no_grad():
            return volume.variance_cost_volume(fe, cams, dv, w["geo"])
    return run, ws["B"] * ws["V"] * ws["D"] * ws["H"] * ws["W"]


def gpu_eager_baseline(w, dev, fe, cams, dv):
    """The reference's GPU eager path for the cost-volume build (north_star's >=10x denominator):
    the oracle restates `networks/casred.py:26-53` op for op, so running it on CUDA tensors launches
    the same ATen kernel chain the reference does.  Reported, never shipped."""
    from oracle import volume
    cams_d = cams.to(dev)
    try:
        with torch.no_grad():
            for _ in range(3):
                volume.variance_cost_volume(fe, cams_d, dv, w["geo"])
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats(dev)
            best = float("inf")
            for _ in range(10):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                volume.variance_cost_volume(fe, cams_d, dv, w["geo"])
                e.record()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e))
        vox = w["B"] * w["V"] * w["D"] * w["H"] * w["W"]
        return {"value": vox / (best * 1e-3), "unit": UNIT, "ms": best, "kind": "port of networks/casred.py:26-53 on CUDA tensors "
                "(ATen eager, best of 10)", "peak_alloc_bytes": torch.cuda.max_memory_allocated(dev)}
    except Exception as ex:  # an OOM in the baseline must not lose the bench line
        return {"unavailable": repr(ex)[:200]}


def cpu_baseline(w, budget_s=20.0):
    planes = min(w["D"], 16)
    run, vox = cpu_sample(w, planes)
    run()
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 10):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    best = min(ts)
    return {"value": vox / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle.volume.variance_cost_volume (restatement of networks/casred.py:26-53 + modules/warping.py:310-365) "
                      f"on the first {planes} of {w['D']} planes of the same stack, best of {len(ts)}, torch-CPU "
                      f"{torch.get_num_threads()} threads of {os.cpu_count()} logical cores",
            "seconds_per_sample": best}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    planes = min(w["D"], 16)
    run, vox = cpu_sample(w, planes)
    for _ in range(min(args.warmup, 2)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    val = vox / dt
    cb = {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"oracle port of the reference torch-CPU path, first {planes} of {w['D']} planes per step"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (geometry f64)", "data": "synthetic",
        "config": {"workload": w["desc"], "name": args.workload, "B": w["B"], "V": w["V"], "C": w["C"], "D": w["D"],
                   "H": w["H"], "W": w["W"], "geo_model": w["geo"]},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_build", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
