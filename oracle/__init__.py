"""oracle/ — TEST INFRASTRUCTURE ONLY.  CPU restatement of the SatMVS plane-sweep path.

This package restates, in plain torch-CPU / numpy arithmetic, the algorithms of the
reference's hot path (`modules/warping.py`, `modules/module.py`, `modules/depth_range.py`,
`networks/casred.py`, `networks/casmvs.py`, `tools/rpc_tensor.py`).  Every function cites
the reference file:line it follows.  It is the *checker* for the CUDA kernels in
`satmvs_b200/`; it is never the thing shipped or measured:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
  `--impl reference` legs may import it;
* nothing under `satmvs_b200/` imports it (tests/test_boundary.py greps for that).

Pinning.  The reference ships no golden vectors or tests for this path, so
the oracle is pinned against *outputs of the reference itself*: `oracle/make_golden.py`
imports the unmodified reference from `/root/reference` in the build container, runs it on
seeded synthetic inputs and commits the input/output vectors under `tests/golden/`.
`tests/test_oracle_golden.py` checks this restatement against those vectors (bit-exact for the
fp64 geometry, <=1e-6 for the fp32 network outputs); `python oracle/make_golden.py --out /tmp/x`
re-generates the vectors from the live reference when `/root/reference` is present.

Third-party arithmetic on the path (`F.grid_sample`, `F.interpolate`, conv/BN/GN) lives in
PyTorch (reference pin: pytorch 1.4.0, `environment.yml:121`; here torch 2.11).  The bilinear
sampler is additionally restated explicitly (`geometry.bilinear_sample_zeros`) following the
published `grid_sampler_2d` algorithm and cross-checked against `F.grid_sample`.
"""
from . import geometry, volume, regnets, regress, hypotheses, stages  # noqa: F401
