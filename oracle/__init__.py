"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the humanoid ping-pong task hot path.

Nothing in the product package (`isaacgym_b200/`) may import this package.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` use it, and there only as the checker / CPU baseline.

Contents
--------
* `jit_utils_restated`  -- the six quaternion helpers of
  `isaacgymenvs.utils.torch_jit_utils` that the reference calls but does not
  vendor (parity UNPINNED at that boundary: no version is pinned anywhere in the
  reference tree; formulas restated from the public upstream module).
* `pingpong_oracle`     -- CPU (torch fp32) restatement of every live
  observation / reward / reset / action function of the 7 task variants, each
  citing the reference file:line it follows.
* `ref_extract`         -- loads the reference's own function source from
  `/root/reference` (only possible in the build container) so the restatement
  can be pinned against it; `make_golden.py` freezes the outputs of those
  reference functions into `tests/golden/*.npz`.

Parity status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's own
functions executed in the build container (fixtures + generating script are
committed).  The un-vendored `torch_jit_utils` helpers remain "parity unpinned".
"""
