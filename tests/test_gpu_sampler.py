"""Device-side ball launch sampler against the reference's host sampler
(restated in oracle.pingpong_oracle.sample_ball_velocity: same `random.uniform` draws and formulas
as TILT:307-318 / NES:312-323 / ADOF:357-367 / A3:300-302).  Bit parity with Mersenne-Twister is
impossible for a counter-based device stream; parity is statistical: ranges, moments, KS distance."""
import random

import numpy as np
import pytest
import torch
from scipy import stats as sps

from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import make_state
from isaacgym_b200.tasks import make_task
from oracle import pingpong_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("variant", ["a3", "tilt", "nes", "align", "a4", "adof", "align2"])
def test_sampler_matches_reference_distribution(variant):
    cfg = CONFIGS[variant]
    n = 40000
    task = make_task(variant, make_state(cfg, n, seed=1, adversarial=False), device=DEV)
    task.sample_ball_launch(seed=1234, epoch=0)
    torch.cuda.synchronize()
    got = task.st["reset_ball_vel"].cpu().numpy()
    rng = random.Random(99)
    want = np.array([O.sample_ball_velocity(rng, variant) for _ in range(n)], dtype=np.float32)
    assert np.isfinite(got).all()
    for c in range(3):
        if np.ptp(want[:, c]) == 0:                      # A3: vz == 0
            assert (got[:, c] == want[0, c]).all()
            continue
        assert got[:, c].min() >= want[:, c].min() - 1e-3 * abs(want[:, c].min()) - 1e-3
        assert got[:, c].max() <= want[:, c].max() + 1e-3 * abs(want[:, c].max()) + 1e-3
        ks = sps.ks_2samp(got[:, c], want[:, c])
        assert ks.statistic < 0.02, (variant, c, ks)
    # speed |v| follows the reference's construction exactly for the NES/ADOF/A3 forms (|v| = s)
    if variant in ("nes", "adof", "a3"):
        lo, hi = {"nes": (5.4, 5.9), "adof": (5.0, 5.4), "a3": (6.5, 7.5)}[variant]
        sp = np.linalg.norm(got, axis=1)
        assert sp.min() >= lo - 1e-4 and sp.max() <= hi + 1e-4
    if variant == "adof":
        yz = task.st["reset_ball_pos_yz"].cpu().numpy()
        assert yz[:, 0].min() >= -0.5 and yz[:, 0].max() <= 0.1 and yz[:, 1].min() >= 0.96 and yz[:, 1].max() <= 1.05
        assert abs(yz[:, 0].mean() + 0.2) < 0.01 and abs(yz[:, 1].mean() - 1.005) < 0.002


def test_sampler_is_counter_based():
    """Same (seed, epoch) -> same table; a shard draws exactly the rows of the global table;
    a different epoch gives a different table; refresh touches only envs that reset."""
    cfg = CONFIGS["tilt"]
    n = 4096
    st = make_state(cfg, n, seed=2, adversarial=False)
    a = make_task("tilt", st, device=DEV)
    b = make_task("tilt", st, device=DEV)
    a.sample_ball_launch(seed=7, epoch=3)
    b.sample_ball_launch(seed=7, epoch=3)
    assert torch.equal(a.st["reset_ball_vel"], b.st["reset_ball_vel"])
    half = {k: (v[n // 2:].clone() if (v.dim() > 0 and v.shape[0] == n) else v.clone()) for k, v in st.items()}
    s = make_task("tilt", half, device=DEV)
    s.sample_ball_launch(seed=7, epoch=3, env_offset=n // 2)
    assert torch.equal(s.st["reset_ball_vel"], a.st["reset_ball_vel"][n // 2:])
    b.sample_ball_launch(seed=7, epoch=4)
    assert not torch.equal(a.st["reset_ball_vel"], b.st["reset_ball_vel"])
    # refresh only consumed rows
    before = a.st["reset_ball_vel"].clone()
    a.reset_buf.zero_()
    a.reset_buf[::7] = 1
    a.sample_ball_launch(seed=7, epoch=5, refresh_consumed_only=True)
    torch.cuda.synchronize()
    changed = (a.st["reset_ball_vel"] != before).any(dim=1)
    assert torch.equal(changed, a.reset_buf.bool())


@pytest.mark.parametrize("variant", ["tilt", "a4", "adof"])
def test_unaligned_shard_takes_the_generic_path(variant):
    """A shard that starts at an odd env of a larger allocation is only 8-byte aligned: bulk async
    staging is illegal there (KArgs.bulk_ok = 0) and every tile goes through the LDG path.  Results
    must be identical to the aligned (bulk) run on the same envs."""
    from isaacgym_b200 import _native as N
    cfg = CONFIGS[variant]
    n = 1024
    st = make_state(cfg, n + 1, seed=8, adversarial=False)
    lib = N.load()

    def step(state):
        g = {k: v.to(DEV) for k, v in state.items()}
        g["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=DEV)
        g["scratch"] = torch.zeros(16, dtype=torch.int32, device=DEV)
        return g

    big = step(st)
    # shard A: envs 1..n as views into the big tensors (misaligned start)
    view = {k: (v[1:] if (v.dim() > 0 and v.shape[0] == n + 1) else v) for k, v in big.items()}
    assert view["rigid_body_states"].data_ptr() % 16 != 0
    # shard B: the same envs, freshly allocated (aligned)
    copy = {k: (v[1:].clone() if (v.dim() > 0 and v.shape[0] == n + 1) else v.clone()) for k, v in big.items()}
    for g in (view, copy):
        g["progress_buf"] = g["progress_buf"].clone()      # int64 views at odd offsets are still 8-byte aligned
        N.check(lib.ppk_post_physics_step(N.make_task(cfg), N.make_buffers(cfg, g), N.PHASE_ALL, N.current_stream_ptr()), "step")
    torch.cuda.synchronize()
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names + cfg.state_names:
        assert torch.equal(view[name], copy[name]), name
