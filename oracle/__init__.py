"""CPU oracle for the FLAIR hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy / torch-CPU fp32 restatement of the reference algorithm for the
path named in BASELINE.json (sampler update, data-consistency operators, UNet
forward).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and only as the checker /
CPU baseline; nothing under `guided_diffusion/` or `flair_b200/` imports it.

Pinning.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by `tools/gen_golden.py` (which imports `/root/reference` unmodified)
and committed under `tests/golden/`.  Pieces whose arithmetic lives in the
un-vendored mmedit==0.12.0 / mmcv-full==1.4.8 (SPyNet, flow_warp,
ResidualBlocksWithInputConv, ModulatedDeformConv2d) are restated from their
published behaviour: **parity unpinned** for those (video-mode BasicVSR++ only).

`windowing.py` restates `more_itertools.windowed` (requirements.txt:8, un-vendored,
unpinned) from its published algorithm and documented examples.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
