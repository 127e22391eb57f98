"""CPU tests of the oracle: against the golden vectors frozen from the reference's own
functions, against the reference itself when /root/reference is present (build container),
and property tests of the restated torch_jit_utils helpers (parity unpinned there)."""
import math

import pytest
import torch

from helpers import STEP_VARIANTS, VARIANTS, load_golden
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import clone_state, make_state
from oracle import jit_utils_restated as J
from oracle import ref_extract, task_oracle


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_matches_golden(variant):
    cfg = CONFIGS[variant]
    ins, outs = load_golden(variant)
    st = clone_state(ins)
    if variant != "base":
        st["progress_buf"] += 1
    task_oracle.compute_reward(cfg, st)
    obs = task_oracle.compute_observations(cfg, st)
    # same machine class + same torch build -> the restatement reproduces the reference bit for bit;
    # allow 2 ulp on floats so a different host CPU's libm/SIMD dispatch cannot fail the suite
    torch.testing.assert_close(obs, outs["obs_buf"], rtol=3e-7, atol=1e-7)
    torch.testing.assert_close(st["rew_buf"], outs["rew_buf"], rtol=3e-7, atol=1e-4)
    assert torch.equal(st["reset_buf"], outs["reset_buf"])
    for name in cfg.flag_names + cfg.counter_names + cfg.state_names:
        assert torch.equal(st[name], outs[name]), name


@pytest.mark.skipif(not ref_extract.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_matches_reference_source(variant):
    cfg = CONFIGS[variant]
    ref = task_oracle.ReferenceImpl()
    for seed in (11, 12):
        st0 = make_state(cfg, 2048, seed=seed)
        a, b = clone_state(st0), clone_state(st0)
        if variant != "base":
            a["progress_buf"] += 1
            b["progress_buf"] += 1
        task_oracle.compute_reward(cfg, a)
        oa = task_oracle.compute_observations(cfg, a)
        with ref_extract.quiet():
            task_oracle.compute_reward(cfg, b, ref)
            ob = task_oracle.compute_observations(cfg, b, ref)
        assert torch.equal(oa, ob)
        for k in a:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.skipif(not ref_extract.available(), reason="/root/reference only exists in the build container")
def test_scripted_flag_updates_are_not_in_place():
    """Defect D16: `flag |= x` inside @torch.jit.script (ALIGN:1097, A4:1113/1280) does not
    mutate the caller's tensor, whereas the eager TILT function does."""
    ref = task_oracle.ReferenceImpl()
    for variant, mutates in (("tilt", True), ("align", False), ("a4", False)):
        cfg = CONFIGS[variant]
        st0 = make_state(cfg, 1024, seed=5)
        st = clone_state(st0)
        with ref_extract.quiet():
            task_oracle.compute_reward(cfg, st, ref)
        changed = any(not torch.equal(st[n], st0[n]) for n in cfg.flag_names)
        assert changed == mutates


@pytest.mark.parametrize("variant", [v for v in VARIANTS if v != "base"])
def test_post_physics_step_semantics(variant):
    """reset envs: progress 0, flags at reset values, root/dof rows from the initial tensors,
    ball velocity from the pre-sampled launch; obs of reset envs see the new DOF state."""
    cfg = CONFIGS[variant]
    st0 = make_state(cfg, 512, seed=3)
    st = clone_state(st0)
    env_ids, idx, stats = task_oracle.post_physics_step(cfg, st)
    assert len(env_ids) > 0
    rst = st["reset_buf"].bool()
    assert torch.equal(st["progress_buf"][rst], torch.zeros(int(rst.sum()), dtype=torch.int64))
    assert torch.equal(st["progress_buf"][~rst], st0["progress_buf"][~rst] + 1)
    for name, val in zip(cfg.flag_names, cfg.flag_reset_values):
        assert bool((st[name][rst] == val).all())
    b = cfg.ball_actor
    assert torch.equal(st["root_states"][rst][:, b, 7:10], st0["reset_ball_vel"][rst])
    want = st0["initial_root_states"][rst][:, :, 0:7].clone()
    if variant == "adof":                                   # ADOF:976-979 also re-draws ball y, z
        want[:, b, 1:3] = st0["reset_ball_pos_yz"][rst]
    assert torch.equal(st["root_states"][rst][:, :, 0:7], want)
    assert torch.equal(st["root_states"][~rst], st0["root_states"][~rst])
    if cfg.reset_dof:
        assert torch.equal(st["dof_states"][rst], st0["initial_dof_states"][rst])
    else:
        assert torch.equal(st["dof_states"], st0["dof_states"])
    actor_idx, dof_idx = idx
    assert actor_idx.dtype == torch.int32 and actor_idx.numel() == len(env_ids) * cfg.num_actors
    assert stats["reset_count"] == float(rst.sum())


# ---- helpers (parity unpinned upstream: pinned by properties) ----------------------------------

def _rand_unit_quat(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


def test_rotation_preserves_norm_and_identity():
    q = _rand_unit_quat(1000)
    v = torch.randn(1000, 3, generator=torch.Generator().manual_seed(1))
    r = J.my_quat_rotate(q, v)
    torch.testing.assert_close(r.norm(dim=-1), v.norm(dim=-1), rtol=1e-5, atol=1e-6)
    ident = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(1000, 1)
    assert torch.equal(J.my_quat_rotate(ident, v), v)


def test_heading_of_yaw_quaternion_is_yaw_and_inverse_undoes_it():
    yaw = torch.linspace(-3.0, 3.0, 61)
    q = torch.stack([torch.zeros_like(yaw), torch.zeros_like(yaw), (yaw / 2).sin(), (yaw / 2).cos()], dim=-1)
    torch.testing.assert_close(J.calc_heading(q), yaw, rtol=1e-5, atol=1e-6)
    hq = J.calc_heading_quat_inv(q)
    x = torch.tensor([[1.0, 0.0, 0.0]]).repeat(61, 1)
    fwd = J.my_quat_rotate(q, x)
    back = J.my_quat_rotate(hq, fwd)
    torch.testing.assert_close(back, x, rtol=1e-5, atol=1e-6)


def test_heading_rotation_leaves_z_unchanged_and_degenerate_heading_is_zero():
    q = _rand_unit_quat(500, seed=2)
    hq = J.calc_heading_quat_inv(q)
    v = torch.randn(500, 3, generator=torch.Generator().manual_seed(3))
    r = J.my_quat_rotate(hq, v)
    torch.testing.assert_close(r[:, 2], v[:, 2], rtol=1e-5, atol=1e-6)
    s = math.sqrt(0.5)
    deg = torch.tensor([[0.0, s, 0.0, s], [0.0, 0.0, 0.0, 0.0]])   # x-axis maps to +-z / zero quaternion
    h = J.calc_heading(deg)
    assert torch.isfinite(h).all()
    # zero quaternion: 2w^2-1 = -1 flips the x-axis -> heading pi -> inverse heading quat (0,0,-1,~0)
    hq = J.calc_heading_quat_inv(torch.tensor([[0.0, 0.0, 0.0, 0.0]]))
    torch.testing.assert_close(hq, torch.tensor([[0.0, 0.0, -1.0, 0.0]]), rtol=0, atol=1e-6)
    # identity quaternion: heading 0 -> exact identity
    hq = J.calc_heading_quat_inv(torch.tensor([[0.0, 0.0, 0.0, 1.0]]))
    assert torch.equal(hq.abs(), torch.tensor([[0.0, 0.0, 0.0, 1.0]]))


def test_unnormalised_quaternion_is_not_normalised():
    """my_quat_rotate uses 2w^2-1 directly: scaling q by 2 changes the result."""
    q = _rand_unit_quat(10, seed=4)
    v = torch.randn(10, 3, generator=torch.Generator().manual_seed(5))
    assert not torch.allclose(J.my_quat_rotate(2 * q, v), J.my_quat_rotate(q, v))


# ---- learner side ----------------------------------------------------------------
def test_running_mean_std_restatement_properties():
    """rl_games is absent (parity unpinned): check the restated update against first principles -- merging
    batches one by one equals the statistics of their concatenation plus the one prior pseudo-sample
    (mean 0, var 1, count 1) the module starts from."""
    from oracle import policy_oracle as P
    g = torch.Generator().manual_seed(3)
    rms = P.RunningMeanStd(7)
    batches = [torch.randn(n, 7, generator=g, dtype=torch.float32) * 3 + 1.5 for n in (64, 257, 1000)]
    for b in batches:
        rms.forward(b)
    allx = torch.cat(batches).double()
    n = allx.shape[0]
    mean = allx.mean(0)
    m2 = ((allx - mean) ** 2).sum(0)                       # n-weighted second moment about the data mean
    tot = n + 1.0
    want_mean = mean * n / tot                             # merged with the prior sample at 0
    # chained parallel-variance merges with unbiased batch variances are not the plain pooled variance; replay
    # them in float64 with the same formula on exact batch moments instead
    m, v, c = torch.zeros(7, dtype=torch.float64), torch.ones(7, dtype=torch.float64), 1.0
    for b in batches:
        bd = b.double()
        m, v, c = P.RunningMeanStd._update_mean_var_count_from_moments(m, v, c, bd.mean(0), bd.var(0), bd.shape[0])
    torch.testing.assert_close(rms.running_mean, want_mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rms.running_mean, m, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rms.running_var, v, rtol=1e-5, atol=1e-6)
    assert float(rms.count) == tot and m2.min() > 0
    y = rms.normalize(batches[0])
    assert float(y.abs().max()) <= 5.0 and y.dtype == torch.float32


def test_first_layer_restatement_equals_torch_autocast():
    """The fp16 first layer restated with explicit roundings is what torch.autocast(float16) computes."""
    from oracle import policy_oracle as P
    g = torch.Generator().manual_seed(5)
    for width, units in ((80, 2048), (94, 512), (24, 256)):
        x = (torch.randn(300, width, generator=g) * 2).clamp(-5, 5)
        w = torch.randn(units, width, generator=g) / width ** 0.5
        b = torch.randn(units, generator=g) * 0.1
        for act in ("elu", "None"):
            a = P.first_layer(x, w, b, act).float()
            t = P.first_layer_torch_autocast(x, w, b, act).float()
            # identical up to fp32 summation order: at most one fp16 ulp, on a small fraction of entries
            assert float(((a - t).abs() > 2.0 ** -10 * t.abs() + 1e-7).float().mean()) == 0.0
            assert float((a != t).float().mean()) < 0.01


@pytest.mark.parametrize("variant", STEP_VARIANTS)
def test_oracle_step_matches_reference_method_fixture(variant):
    """tests/golden/<variant>_step.npz: one whole post_physics_step executed from the reference's own class-method
    text (reward -> nonzero -> _reset_idx -> observations).  The restated step reproduces every buffer bit for bit."""
    from isaacgym_b200.config import CONFIGS
    from oracle import task_oracle
    cfg = CONFIGS[variant]
    ins, outs = load_golden(variant, step=True)
    st = {k: v.clone() for k, v in ins.items()}
    env_ids, idx, _ = task_oracle.post_physics_step(cfg, st)
    assert len(env_ids) > 0
    for key, want in outs.items():
        if key == "reset_actor_indices":
            assert torch.equal(idx[0], want)
        else:
            assert torch.equal(st[key], want), (variant, key)
