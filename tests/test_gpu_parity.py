"""GPU parity tests: the CUDA path, called through the C ABI (libppk.so via ctypes), against
 (1) the golden vectors frozen from the reference's own functions,
 (2) the CPU oracle on the same seeded inputs (sizes the oracle finishes in seconds),
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: reset_buf, progress_buf, flags, counters, reset rows -- bit exact;
obs / reward fp32 -- rtol 1e-5 (+ small atol), per field, see helpers.assert_close_fields."""
import pytest
import torch

from helpers import STEP_VARIANTS, VARIANTS, assert_close_fields as _assert_close_fields, load_golden, oracle_step_with_term_scale
from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import clone_state, make_state
from isaacgym_b200.tasks import make_task
from oracle import task_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_state(st):
    g = clone_state(st, DEV)
    g["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=DEV)
    g["scratch"] = torch.zeros(16, dtype=torch.int32, device=DEV)
    return g


def run(cfg, g, phases):
    lib = N.load()
    task = N.make_task(cfg)
    buf = N.make_buffers(cfg, g)
    N.check(lib.ppk_post_physics_step(task, buf, phases, N.current_stream_ptr()), "step")
    torch.cuda.synchronize()


def assert_close_fields(cfg, got_obs, want_obs, got_rew, want_rew, context="", rew_scale=None):
    _assert_close_fields(cfg, got_obs, want_obs, got_rew, want_rew, context, rew_scale=rew_scale)


def check_fields(cfg, g, want, context):
    """obs / reward of the CUDA path against an oracle state produced by `oracle_full_step` (which also recorded,
    per env, the magnitude of the terms that entered the reward: the per-env absolute tolerance)."""
    _assert_close_fields(cfg, g["obs_buf"], want["obs_buf"], g["rew_buf"], want["rew_buf"], context,
                         rew_scale=want.get("_rew_scale"))


def assert_exact(cfg, g, want, names, context):
    for name in names:
        a, b = g[name].cpu(), want[name]
        assert torch.equal(a, b), f"{context}: {name} differs in {int((a != b).sum())} of {a.numel()} entries"


@pytest.mark.parametrize("variant", VARIANTS)
def test_golden_vectors(variant):
    """Outputs of the reference's own functions (tests/golden, made by oracle/make_golden.py)."""
    cfg = CONFIGS[variant]
    ins, outs = load_golden(variant)
    g = gpu_state(ins)
    if variant == "base":
        run(cfg, g, N.PHASE_REWARD | N.PHASE_OBS)
    else:
        run(cfg, g, N.PHASE_PROGRESS | N.PHASE_REWARD | N.PHASE_OBS)
    # the fixtures carry no term breakdown: the oracle on the same inputs says which terms entered each env's reward
    o = clone_state(ins)
    if variant != "base":
        o["progress_buf"] += 1
    _, scale = oracle_step_with_term_scale(cfg, o, lambda c, s_: task_oracle.compute_reward(c, s_))
    assert_close_fields(cfg, g["obs_buf"], outs["obs_buf"], g["rew_buf"], outs["rew_buf"], f"golden[{variant}]", rew_scale=scale)
    assert torch.equal(g["reset_buf"].cpu(), outs["reset_buf"])
    assert_exact(cfg, g, outs, cfg.flag_names + cfg.counter_names + cfg.state_names, f"golden[{variant}]")


def oracle_full_step(cfg, st):
    o = clone_state(st)

    def step(cfg, o):
        if cfg.variant == "base":
            from oracle import pingpong_oracle as O
            o["progress_buf"] += 1
            ids = o["reset_buf"].nonzero(as_tuple=False).flatten()
            if len(ids) > 0:
                O.base_reset_idx(o, ids, o["reset_ball_vel"][0], o["reset_ball_vel"][1])
            o["obs_buf"][:] = task_oracle.compute_observations(cfg, o)
            task_oracle.compute_reward(cfg, o)
            return task_oracle.step_stats(cfg, o)
        return task_oracle.post_physics_step(cfg, o)[2]

    stats, scale = oracle_step_with_term_scale(cfg, o, step)
    o["_rew_scale"] = scale
    return o, stats


STATE_EXACT = ("reset_buf", "progress_buf", "root_states", "dof_states")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n,seed", [(4096, 101), (1000, 102), (33, 103)])
def test_fused_step_matches_oracle(variant, n, seed):
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=seed)
    if variant == "base":
        st["reset_buf"] = (torch.rand(n, generator=torch.Generator().manual_seed(seed)) < 0.1).to(torch.int64)
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    want, stats = oracle_full_step(cfg, st)
    g = gpu_state(st)
    run(cfg, g, N.PHASE_ALL)
    ctx = f"{variant} n={n}"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, ctx)
    check_fields(cfg, g, want, ctx)
    # statistics: fp64 partial sums on the device vs fp64 oracle sums
    got = g["stats"].sum(dim=0).cpu()
    rs = stats["reward_sum"]
    rs = float(rs[0]) if rs.dim() > 0 else float(rs)
    assert abs(float(got[0]) - rs) <= 1e-6 * max(1.0, abs(rs)) + 1e-3
    assert float(got[1]) == float(stats["progress_sum"]) and float(got[2]) == float(stats["reset_count"])
    if variant == "adof":
        order = ("fall_down_count", "closer_to_paddle_count", "hit_paddle_count", "cross_net_count", "hit_table_count")
        for i, name in enumerate(order):
            assert float(got[3 + i]) == float(stats[name]), name


@pytest.mark.parametrize("variant", VARIANTS)
def test_unfused_reference_call_sequence_equals_fused(variant):
    """compute_reward -> reset_idx(nonzero(reset_buf)) -> compute_observations as separate calls
    (the reference's own sequence) gives bit-identical buffers to the single fused kernel."""
    cfg = CONFIGS[variant]
    st = make_state(cfg, 2048, seed=7)
    if variant == "base":
        st["reset_buf"] = (torch.rand(2048, generator=torch.Generator().manual_seed(1)) < 0.1).to(torch.int64)
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    a = make_task(variant, st, device=DEV, fused=True)
    b = make_task(variant, st, device=DEV, fused=False)
    a.post_physics_step()
    b.post_physics_step()
    torch.cuda.synchronize()
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf") + cfg.flag_names + cfg.counter_names + cfg.state_names:
        if variant == "adof" and name == "rew_buf":
            # ADOF's fused step and its single-phase calls are two kernels (ppk_adof2.cuh / ppk_adof.cuh) that add the
            # 23 / 27 terms of the imitation sums in different orders: the reward agrees to fp32 rounding, not bitwise
            torch.testing.assert_close(a.rew_buf, b.rew_buf, rtol=2e-6, atol=2e-6 * 3000.0)
            continue
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(a.root_states, b.root_states) and torch.equal(a.vec_dof_states, b.vec_dof_states)


@pytest.mark.parametrize("variant", [v for v in VARIANTS if v != "base"])
def test_multi_step_trajectory(variant):
    """Flags are stateful across steps: run 12 task steps with a synthetic 'physics' that moves the
    ball, mirrored on the oracle, and compare every step."""
    cfg = CONFIGS[variant]
    n = 512
    st = make_state(cfg, n, seed=21)
    o = clone_state(st)
    task = make_task(variant, st, device=DEV, full_pre_ball_clone=(variant == "align"), launch_seed=None)
    gen = torch.Generator().manual_seed(5)
    b = cfg.ball_actor
    for step in range(12):
        actions = torch.rand(n, cfg.num_dofs, generator=gen) * 2 - 1
        # pre_physics_step
        o["actions"] = actions.clone()
        task_oracle.pre_physics_step(cfg, o)
        task.pre_physics_step(actions.to(DEV))
        torch.testing.assert_close(task.pd_tar.cpu(), o["pd_targets"], rtol=1e-6, atol=1e-7)
        # "physics": ballistic ball + occasional velocity flips, same on both sides
        dv = torch.randn(n, 3, generator=gen) * 0.5
        flip = (torch.rand(n, generator=gen) < 0.2)
        for root in (o["root_states"], None):
            if root is None:
                r = task.root_states
                dvd, flipd = dv.to(DEV), flip.to(DEV)
            else:
                r, dvd, flipd = root, dv, flip
            r[:, b, 0:3] += 0.05 * r[:, b, 7:10]
            r[:, b, 7:10] += dvd
            r[:, b, 7] = torch.where(flipd, -r[:, b, 7], r[:, b, 7])
        _, scale = oracle_step_with_term_scale(cfg, o, lambda c, s_: task_oracle.post_physics_step(c, s_))
        task.post_physics_step()
        torch.cuda.synchronize()
        ctx = f"{variant} step {step}"
        for name in ("reset_buf", "progress_buf") + cfg.flag_names + cfg.counter_names + cfg.state_names:
            assert torch.equal(getattr(task, name).cpu(), o[name]), f"{ctx}: {name}"
        assert torch.equal(task.root_states.cpu(), o["root_states"]), ctx
        assert torch.equal(task.vec_dof_states.cpu(), o["dof_states"]), ctx
        assert_close_fields(cfg, task.obs_buf, o["obs_buf"], task.rew_buf, o["rew_buf"], ctx, rew_scale=scale)


@pytest.mark.parametrize("variant", ["tilt", "a4", "adof", "base"])
@pytest.mark.parametrize("n", [1, 7, 31, 32, 65])
def test_ragged_sizes(variant, n):
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=n, adversarial=False)
    if variant == "base":
        st["reset_ball_vel"] = st["reset_ball_vel"][:1].repeat(2, 1).contiguous()
    want, _ = oracle_full_step(cfg, st)
    g = gpu_state(st)
    # guard rows after the tensors: nothing beyond N may be written (compute-sanitizer is closed on
    # this pool, so out-of-bounds stores are caught with sentinels instead)
    guarded = {}
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names + cfg.state_names:
        t = g[name]
        big = torch.empty((n + 8,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if t.dtype == torch.bool:
            big[:] = True
        else:
            big.fill_(-777)
        big[:n] = t
        guarded[name] = big
        g[name] = big[:n]
    run(cfg, g, N.PHASE_ALL)
    for name, big in guarded.items():
        tail = big[n:]
        ok = bool(tail.all()) if big.dtype == torch.bool else bool((tail == -777).all())
        assert ok, f"{variant} n={n}: {name} was written past its end"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, f"{variant} n={n}")
    check_fields(cfg, g, want, f"{variant} n={n}")


def test_empty_batch_is_a_no_op():
    cfg = CONFIGS["tilt"]
    st = make_state(cfg, 0, seed=0, adversarial=False)
    g = gpu_state(st)
    lib = N.load()
    assert lib.ppk_post_physics_step(N.make_task(cfg), N.make_buffers(cfg, g), N.PHASE_ALL, None) == 0


@pytest.mark.parametrize("variant", [v for v in VARIANTS if v != "base"])
def test_reset_idx_entry_point(variant):
    cfg = CONFIGS[variant]
    n = 700
    st = make_state(cfg, n, seed=9)
    o = clone_state(st)
    env_ids = torch.tensor([0, 3, 64, 65, 699, 311], dtype=torch.int64)
    gen = torch.Generator().manual_seed(2)
    vel = torch.randn(len(env_ids), 3, generator=gen)
    yz = torch.randn(len(env_ids), 2, generator=gen)
    from oracle import pingpong_oracle as O
    want_idx = O.reset_idx(variant, o, env_ids, vel, yz if variant == "adof" else None)
    task = make_task(variant, st, device=DEV)
    got_idx = task.reset_idx(env_ids, ball_vel=vel, ball_pos_yz=yz if variant == "adof" else None)
    torch.cuda.synchronize()
    assert torch.equal(task.root_states.cpu(), o["root_states"])
    assert torch.equal(task.vec_dof_states.cpu(), o["dof_states"])
    assert torch.equal(task.progress_buf.cpu(), o["progress_buf"])
    for name in cfg.flag_names:
        assert torch.equal(getattr(task, name).cpu(), o[name]), name
    assert got_idx[0].dtype == torch.int32 and torch.equal(got_idx[0].cpu(), want_idx[0])
    assert torch.equal(got_idx[1].cpu(), want_idx[1])


def test_pre_ball_clone_layouts_agree():
    """The saved pre-step ball state may be the reference's full 13-float clone or the two floats
    the rewards read; both give identical results."""
    cfg = CONFIGS["align"]
    st = make_state(cfg, 1024, seed=4)
    a = make_task("align", st, device=DEV, full_pre_ball_clone=True)
    b = make_task("align", st, device=DEV, full_pre_ball_clone=False)
    act = torch.zeros(1024, cfg.num_dofs, device=DEV)
    for t in (a, b):
        t.pre_physics_step(act)
        t.post_physics_step()
    torch.cuda.synchronize()
    # the full clone holds the ball row as it was when pre_physics_step ran (TILT:1020), the compact one its vx, vz
    assert torch.equal(a.pre_ball2_root_states[:, [7, 9]], b.pre_ball2_root_states)
    assert torch.equal(a.rew_buf, b.rew_buf) and torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.reset_buf, b.reset_buf)


# ---- full BASELINE.json sizes: properties that do not need the (slow) oracle ----------------------

@pytest.mark.parametrize("variant,n", [("a3", 16384), ("tilt", 65536), ("a4", 65536), ("adof", 32768), ("align", 131072)])
def test_full_size_sharding_and_determinism(variant, n):
    """Envs are independent: running the batch as one shard or as two/four contiguous shards
    gives bit-identical per-env outputs; a second run on the same inputs is bit-identical too;
    rotation preserves the length of every body vector (||local|| == ||global||)."""
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=1000 * cfg.variant_id, device=DEV, adversarial=False)
    keys = ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names
    def fresh():
        g = {k: v.clone() for k, v in st.items()}
        g["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=DEV)
        g["scratch"] = torch.zeros(16, dtype=torch.int32, device=DEV)
        return g
    whole = fresh()
    run(cfg, whole, N.PHASE_ALL)
    again = fresh()
    run(cfg, again, N.PHASE_ALL)
    for k in keys:
        assert torch.equal(whole[k], again[k]), f"non-deterministic {k}"
    shards = 4
    per = n // shards
    for s in range(shards):
        lo, hi = s * per, (s + 1) * per
        part = {k: (v[lo:hi].clone() if (v.dim() > 0 and v.shape[0] == n) else v.clone()) for k, v in st.items()}
        part["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=DEV)
        part["scratch"] = torch.zeros(16, dtype=torch.int32, device=DEV)
        run(cfg, part, N.PHASE_ALL)
        for k in keys:
            assert torch.equal(part[k], whole[k][lo:hi]), f"shard {s}: {k}"
    # norm preservation of the heading-frame rotation on the velocity block
    J = len(cfg.body_ids)
    obs = whole["obs_buf"].reshape(-1, cfg.num_obs)
    lv = obs[:, 3 * J:6 * J].reshape(-1, J, 3).norm(dim=-1)
    ids = torch.tensor(cfg.body_ids, device=DEV)
    gv = st["rigid_body_states"][:, ids, 7:10].norm(dim=-1)
    if cfg.obs_rows == 2:
        ids2 = torch.tensor(cfg.body_ids_2, device=DEV)
        gv = torch.stack((gv, st["rigid_body_states"][:, ids2, 7:10].norm(dim=-1)), dim=1).reshape(-1, J)
    torch.testing.assert_close(lv, gv, rtol=1e-5, atol=1e-5)
    # reward/reset checksum is reproducible across the two runs and finite
    assert torch.isfinite(whole["rew_buf"]).all() and torch.isfinite(whole["obs_buf"]).all()


def test_stats_reduce_and_task_logging():
    cfg = CONFIGS["tilt"]
    n = 4096
    st = make_state(cfg, n, seed=77)
    task = make_task("tilt", st, device=DEV, log_stats=True)
    o = clone_state(st)
    _, _, stats = task_oracle.post_physics_step(cfg, o)
    task.post_physics_step()          # num_steps == 0 -> a logging step
    m = task.stats.means(n)
    assert abs(m["reward_sum"] - float(stats["reward_sum"]) / n) < 1e-6 * max(1.0, abs(float(stats["reward_sum"]) / n)) + 1e-6
    assert m["progress_sum"] == float(stats["progress_sum"]) / n
    assert float(task.stats.slots.abs().sum()) == 0.0      # slots are cleared by the reduce


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("pin", [True, False])
def test_host_session_equals_device_path(variant, pin):
    """ppk_host_post_physics_step (host tensors, chunked H2D/kernel/D2H pipeline; eager on the first
    call, CUDA-graph capture on the second, replay afterwards) gives the same buffers as the device
    path, step after step.  Pinned tensors take the zero-copy route for the small per-env buffers
    and reset rows, pageable ones the all-DMA route."""
    from isaacgym_b200.host_session import HostSession
    cfg = CONFIGS[variant]
    n = 3000                               # not a multiple of the chunk size: exercises tail tiles
    st = make_state(cfg, n, seed=31)
    if variant == "base":
        st["reset_buf"] = (torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.1).to(torch.int64)
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    dev = gpu_state(st)
    if variant != "base":
        dev["pre_ball_states"] = dev["pre_ball_states"][:, [7, 9]].contiguous()
    sess = HostSession(cfg, st, num_chunks=3, pin=pin)
    try:
        for step in range(4):
            sess.post_physics_step(N.PHASE_ALL & ~N.PHASE_STATS)
            run(cfg, dev, N.PHASE_ALL & ~N.PHASE_STATS)
            for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names + cfg.state_names:
                assert torch.equal(sess.state[name], dev[name].cpu()), f"{variant} step {step}: {name}"
        h2d, d2h = sess.traffic()
        assert h2d > 0 and d2h > 0
    finally:
        sess.close()


@pytest.mark.parametrize("n", [4096, 1001, 5])
def test_adof_compact_reference_pose(n):
    """PpkBuffers.initial_balance_states (the reference pose repacked once at init, [N,23,6]) gives the
    same bits as reading the 28 reference rows of initial_body_states, and matches the oracle."""
    from isaacgym_b200.tasks import pack_reference_pose
    cfg = CONFIGS["adof"]
    st = make_state(cfg, n, seed=77)
    want, _ = oracle_full_step(cfg, st)
    full = gpu_state(st)
    run(cfg, full, N.PHASE_ALL)
    comp = gpu_state(st)
    comp["initial_balance_states"] = pack_reference_pose(cfg, comp["initial_body_states"])
    del comp["initial_body_states"]                      # the compact tensor alone is enough
    run(cfg, comp, N.PHASE_ALL)
    for name in ("obs_buf", "rew_buf") + STATE_EXACT + cfg.flag_names + cfg.counter_names:
        if name == "rew_buf":
            # the compact pose runs on the second ADOF design for full tiles (ppk_adof2.cuh), the uncompacted one on the
            # first: same terms, different summation order of the imitation sums
            torch.testing.assert_close(comp[name], full[name], rtol=2e-6, atol=2e-6 * 3000.0)
            continue
        assert torch.equal(comp[name], full[name]), name
    assert_exact(cfg, comp, want, STATE_EXACT + cfg.flag_names + cfg.counter_names, f"adof compact n={n}")
    check_fields(cfg, comp, want, f"adof compact n={n}")


@pytest.mark.parametrize("variant,n", [("tilt", 65536), ("a4", 65536), ("adof", 32768)])
def test_full_size_yaw_invariance_of_observations(variant, n):
    """Size-independent property at BASELINE sizes: the observations live in the heading frame of the root body,
    so turning the whole scene of every env by its own random yaw angle (positions, velocities, root orientation,
    ADOF's reference pose) leaves obs_buf unchanged up to fp32 rounding."""
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=77 + cfg.variant_id, device=DEV, adversarial=False)
    g = torch.Generator(device=DEV).manual_seed(5)
    th = (torch.rand(n, generator=g, device=DEV) * 2 - 1) * 3.0
    c, s = torch.cos(th), torch.sin(th)

    def turn(rows):                       # rows [n, R, 13]: pos, quat xyzw, linvel, angvel
        out = rows.clone()
        for a in (0, 7):
            x, y = rows[..., a], rows[..., a + 1]
            out[..., a] = c[:, None] * x - s[:, None] * y
            out[..., a + 1] = s[:, None] * x + c[:, None] * y
        # q' = qz(th) * q, qz = (0, 0, sin(th/2), cos(th/2))
        hz, hw = torch.sin(th / 2)[:, None], torch.cos(th / 2)[:, None]
        qx, qy, qz, qw = rows[..., 3], rows[..., 4], rows[..., 5], rows[..., 6]
        out[..., 3] = hw * qx - hz * qy
        out[..., 4] = hw * qy + hz * qx
        out[..., 5] = hw * qz + hz * qw
        out[..., 6] = hw * qw - hz * qz
        return out

    a = {k: v.clone() for k, v in st.items()}
    b = {k: v.clone() for k, v in st.items()}
    for key in ("rigid_body_states", "root_states", "initial_body_states"):
        if key in b:
            b[key] = turn(st[key])
    for d in (a, b):
        d["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=DEV)
        d["scratch"] = torch.zeros(16, dtype=torch.int32, device=DEV)
        run(cfg, d, N.PHASE_OBS)
    oa, ob = a["obs_buf"].reshape(n * cfg.obs_rows, -1), b["obs_buf"].reshape(n * cfg.obs_rows, -1)
    scale = oa.abs().amax(dim=1, keepdim=True).clamp_min(1.0)
    err = (oa - ob).abs() / scale
    if variant == "adof":
        err[:, 120] = 0.0                 # y_intersect divides by (-lvx + 1e-6): unbounded amplification near lvx = 0
    # budget: the turn itself (fp32 sin / cos and two roundings per coordinate) plus atan2f / sinf / cosf of the
    # heading, a few 1e-6 rad on vectors up to ~10x the row scale
    assert float(err.max()) < 5e-5, f"max scaled deviation {float(err.max())}"


@pytest.mark.parametrize("variant", [v for v in VARIANTS if v != "base"])
@pytest.mark.parametrize("seed", [1, 2])
def test_randomised_coefficients(variant, seed):
    """The reward coefficients and the episode length are run-time parameters of PpkTask (YAML values in the
    reference): random ones must flow through exactly as they do through the reference functions."""
    import random
    rnd = random.Random(100 * seed + CONFIGS[variant].variant_id)
    base = CONFIGS[variant]
    cfg = base.with_(
        max_episode_length=rnd.randint(20, 300), alpha=rnd.uniform(1.0, 4000.0), power_coefficient=rnd.uniform(1e-4, 1e-2),
        penalty=-rnd.uniform(10.0, 900.0), hit_table_reward=rnd.uniform(100.0, 5000.0),
        not_hit_table_penalty=-rnd.uniform(100.0, 3000.0), cross_net_reward=rnd.uniform(10.0, 2000.0),
        die_penalty=-rnd.uniform(100.0, 5000.0), hit_paddle_reward=rnd.uniform(10.0, 500.0),
        miss_paddle_penalty_coefficient=-rnd.uniform(10.0, 300.0))
    n = 3000
    st = make_state(cfg, n, seed=500 + seed)
    want, _ = oracle_full_step(cfg, st)
    g = gpu_state(st)
    run(cfg, g, N.PHASE_ALL)
    ctx = f"{variant} random coefficients #{seed}"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, ctx)
    check_fields(cfg, g, want, ctx)


@pytest.mark.parametrize("variant", ["tilt", "nes", "a4"])
def test_scattered_body_ids_take_the_generic_staging_path(variant):
    """`bodyStatesId` is a YAML list: ids that are not one run of consecutive rows (and a paddle that is not the
    last id) make the library fall back from the bulk-copy windows to plain loads -- same results."""
    base = CONFIGS[variant]
    ids = (3, 7, 12, 13, 20, 26, 27, 33, 38, 39)                 # scattered, paddle row 39 last as in the reference
    kw = dict(body_ids=ids)
    if variant == "a4":
        kw["body_ids_2"] = tuple(i + 40 for i in (1, 5, 9, 14, 15, 22, 30, 31, 36, 39))
    cfg = base.with_(**kw)
    n = 2000
    st = make_state(cfg, n, seed=321)
    want, _ = oracle_full_step(cfg, st)
    g = gpu_state(st)
    run(cfg, g, N.PHASE_ALL)
    ctx = f"{variant} scattered ids"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, ctx)
    check_fields(cfg, g, want, ctx)


@pytest.mark.parametrize("variant", ["tilt", "adof", "a4"])
def test_tensors_that_are_only_4_byte_aligned(variant):
    """State tensors handed over as views that start 4 bytes into an allocation (legal for fp32 data) cannot be
    bulk-copied in 16-byte pieces: the library takes its plain-load path -- same results."""
    cfg = CONFIGS[variant]
    n = 1500
    st = make_state(cfg, n, seed=77)
    want, _ = oracle_full_step(cfg, st)
    g = gpu_state(st)
    for key in ("rigid_body_states", "root_states", "dof_states", "dof_forces", "initial_body_states"):
        if key in g:
            t = g[key]
            flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=DEV)
            flat[1:].copy_(t.reshape(-1))
            g[key] = flat[1:].view(t.shape)
            assert g[key].data_ptr() % 16 != 0 and g[key].is_contiguous()
    run(cfg, g, N.PHASE_ALL)
    ctx = f"{variant} misaligned"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, ctx)
    check_fields(cfg, g, want, ctx)


def test_host_session_full_size_tilt():
    """The e2e path of bench.py at its own size: 65 536 envs, pinned host tensors, 4 pipeline chunks, graph replay
    from the third call on -- bit-identical to the device path for four consecutive steps."""
    from isaacgym_b200.host_session import HostSession
    cfg = CONFIGS["tilt"]
    n = 65536
    st = make_state(cfg, n, seed=4242, adversarial=False)
    dev = gpu_state(st)
    dev["pre_ball_states"] = dev["pre_ball_states"][:, [7, 9]].contiguous()
    sess = HostSession(cfg, st, num_chunks=4, pin=True)
    try:
        for step in range(4):
            sess.post_physics_step(N.PHASE_ALL & ~N.PHASE_STATS)
            run(cfg, dev, N.PHASE_ALL & ~N.PHASE_STATS)
            for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names:
                assert torch.equal(sess.state[name], dev[name].cpu()), f"step {step}: {name}"
    finally:
        sess.close()


# ---- BASELINE.json's own sizes against the oracle (the port needs well under a second per config) -----------------

@pytest.mark.parametrize("variant,n,with_pre", [("a3", 16384, False), ("tilt", 65536, False), ("a4", 65536, False),
                                                ("adof", 32768, False), ("align", 131072, True), ("nes", 65536, False),
                                                ("align2", 65536, False), ("base", 4096, False)])
def test_baseline_sizes_match_oracle(variant, n, with_pre):
    """configs[0..4] of BASELINE.json at their full per-GPU sizes: the fused step (ALIGN: pre_physics_step + step,
    the "full task step" of configs[4]) against the oracle on the same seeded inputs, every env compared."""
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=1000 * cfg.variant_id + 5, adversarial=True)
    if variant == "base":
        st["reset_buf"] = (torch.rand(n, generator=torch.Generator().manual_seed(11)) < 0.1).to(torch.int64)
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    o_in = clone_state(st)
    g = gpu_state(st)
    if with_pre:
        task_oracle.pre_physics_step(cfg, o_in)            # pd targets + the ball row clone the reward reads
        lib = N.load()
        N.check(lib.ppk_pre_physics_step(N.make_task(cfg), N.make_buffers(cfg, g), N.current_stream_ptr()), "pre")
        torch.testing.assert_close(g["pd_targets"].cpu(), o_in["pd_targets"], rtol=1e-6, atol=1e-7)
        assert torch.equal(g["pre_ball_states"].cpu(), o_in["pre_ball_states"])
    want, stats = oracle_full_step(cfg, o_in)
    run(cfg, g, N.PHASE_ALL)
    ctx = f"{variant} n={n} (BASELINE size)"
    assert_exact(cfg, g, want, STATE_EXACT + cfg.flag_names + cfg.counter_names + cfg.state_names, ctx)
    check_fields(cfg, g, want, ctx)
    got = g["stats"].sum(dim=0).cpu()
    assert float(got[1]) == float(stats["progress_sum"]) and float(got[2]) == float(stats["reset_count"])


def test_adof_has_fallen_at_the_threshold():
    """ADOF's `has_fallen` compares a MEAN of 23 fp32 norms with 0.32 (ADOF:1412): the only flag of the path that
    hangs on a reduction.  Every env here is planted so that the mean lands within +-32 ulp of 0.32.  A sum of 23
    norms of ~0.32 passes through partial sums of up to 7.4 (ulp 4.8e-7): each of the 22 additions may err by half
    of that, 0.35 ulp(0.32) once divided by 23, so two fp32 summation orders can disagree by up to ~8 ulp(0.32).
    Where the fp64 value of the same fp32 inputs is more than 8 ulp away from the threshold every order gives the
    same answer: the kernel (shuffle tree) must agree with the oracle (ATen's CPU order) exactly.  Inside +-8 ulp
    the reference's own answer depends on the reduction order of the ATen build it runs on (CPU vector ISA, or
    the CUDA reduce kernel in production), so either answer is the reference's: those envs are only counted."""
    cfg = CONFIGS["adof"]
    n = 8192
    st = make_state(cfg, n, seed=4321, adversarial=False)
    bal = torch.tensor(cfg.balance_ids)
    gen = torch.Generator().manual_seed(99)
    ref = st["initial_body_states"][:, bal, 0:3].double()
    d = torch.randn(n, len(bal), 3, generator=gen, dtype=torch.float64)
    ulp = 2.0 ** -25                                          # ulp of fp32 in [0.25, 0.5)
    aim = 0.32 + (torch.rand(n, generator=gen, dtype=torch.float64) * 64.0 - 32.0) * ulp     # +-32 ulp around 0.32
    d = d * (aim / d.norm(dim=-1).mean(dim=-1))[:, None, None]
    st["rigid_body_states"][:, bal, 0:3] = (ref + d).float()
    for name in cfg.counter_names:
        st[name][:] = False
    st["progress_buf"][:] = 0          # no time-outs: a reset anywhere would clear the counters of the shard (ADOF:1171-1175)
    cur = st["rigid_body_states"][:, bal, 0:3]
    exact = (cur.double() - st["initial_body_states"][:, bal, 0:3].double()).norm(dim=-1).mean(dim=-1)
    margin = (exact - float(torch.tensor(0.32, dtype=torch.float32))) / ulp
    assert int((margin.abs() < 8).sum()) > n // 8 and int((margin.abs() > 8).sum()) > n // 2, "the planted batch does not straddle the threshold"
    want, _ = oracle_full_step(cfg, st)
    g = gpu_state(st)
    run(cfg, g, N.PHASE_ALL)
    got_fall, want_fall = g["fall_down_count"].cpu(), want["fall_down_count"]
    clear = margin.abs() > 8.0
    assert torch.equal(got_fall[clear], want_fall[clear]), "has_fallen differs where no fp32 summation order can flip it"
    assert torch.equal(want_fall[clear], (margin > 0)[clear])
    flips = int((got_fall != want_fall)[~clear].sum())
    print(f"has_fallen: {int((~clear).sum())} envs within 8 ulp of 0.32, {flips} of them decided differently by the two summation orders")
    # everything else of those envs is still compared, with the envs whose flag legitimately differs left out
    same = got_fall == want_fall
    for name in STATE_EXACT + cfg.flag_names:
        assert torch.equal(g[name].cpu()[same], want[name][same]), name
    _assert_close_fields(cfg, g["obs_buf"].cpu()[same], want["obs_buf"][same], g["rew_buf"].cpu()[same], want["rew_buf"][same],
                         "adof threshold batch", rew_scale=want["_rew_scale"][same])


@pytest.mark.parametrize("variant", ["tilt", "a4", "adof", "base"])
def test_fused_observation_clamp(variant):
    """VecTask.step's `clamp(obs_buf, -clipObservations, clipObservations)` happens inside the step kernels
    (PpkBuffers.clip_observations); 0 / inf = upstream's default, no clamp."""
    cfg = CONFIGS[variant]
    n = 1500
    st = make_state(cfg, n, seed=8)
    if variant == "base":
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    plain = gpu_state(st)
    run(cfg, plain, N.PHASE_ALL)
    assert float(plain["obs_buf"].abs().max()) > 1.5          # the clamp has something to do
    clipped = gpu_state(st)
    clipped["clip_observations"] = 1.5
    run(cfg, clipped, N.PHASE_ALL)
    assert torch.equal(clipped["obs_buf"], plain["obs_buf"].clamp(-1.5, 1.5))
    for name in ("rew_buf", "reset_buf", "progress_buf"):
        assert torch.equal(clipped[name], plain[name]), name
    inf = gpu_state(st)
    inf["clip_observations"] = float("inf")
    run(cfg, inf, N.PHASE_ALL)
    assert torch.equal(inf["obs_buf"], plain["obs_buf"])


@pytest.mark.parametrize("variant", STEP_VARIANTS)
def test_golden_step_fixture(variant):
    """One whole post_physics_step frozen from the reference's own METHOD bodies (oracle/ref_methods.py, made by
    oracle/make_golden.py): the fused kernel -- reset rows, progress, flags, counters bit-exact; obs / reward per field."""
    cfg = CONFIGS[variant]
    ins, outs = load_golden(variant, step=True)
    g = gpu_state(ins)
    g["actor_indices"], g["dof_indices"] = ins["actor_indices"].to(DEV), ins["dof_indices"].to(DEV)
    n = ins["progress_buf"].shape[0]
    g["reset_count"] = torch.zeros(1, dtype=torch.int32, device=DEV)
    g["reset_actor_indices"] = torch.zeros(n * cfg.num_actors, dtype=torch.int32, device=DEV)
    g["reset_dof_indices"] = torch.zeros(n, dtype=torch.int32, device=DEV)
    run(cfg, g, N.PHASE_ALL)
    ctx = f"golden step[{variant}]"
    exact = ("reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names
    assert_exact(cfg, g, outs, exact, ctx)
    o = clone_state(ins)
    _, scale = oracle_step_with_term_scale(cfg, o, lambda c, s_: task_oracle.post_physics_step(c, s_))
    assert_close_fields(cfg, g["obs_buf"], outs["obs_buf"], g["rew_buf"], outs["rew_buf"], ctx, rew_scale=scale)
    # the compacted int32 actor index list the gym setter takes (TILT:876-883); env order within it is unspecified
    k = int(g["reset_count"].item())
    got = g["reset_actor_indices"][:k * cfg.num_actors].cpu().view(k, cfg.num_actors)
    want = outs["reset_actor_indices"].view(-1, cfg.num_actors)
    assert k == want.shape[0] and torch.equal(got[got[:, 0].argsort()], want[want[:, 0].argsort()])


@pytest.mark.parametrize("variant", ["tilt", "adof", "base"])
@pytest.mark.parametrize("pin", [True, False])
def test_host_session_launch_table_is_a_per_step_input(variant, pin):
    """The launch table is refilled IN PLACE between steps (same host pointers, graph replay from the third call on):
    every reset must consume the fresh values, as the reference draws a new velocity per reset (TILT:857-862).  The
    session also returns the logged sums of a PHASE_STATS call (PpkStat order) in its host `stats` array."""
    from isaacgym_b200.host_session import HostSession
    cfg = CONFIGS[variant]
    n = 2048
    st = make_state(cfg, n, seed=61)
    if variant == "base":
        st["reset_buf"] = (torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.1).to(torch.int64)
        st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
    dev = gpu_state(st)
    if variant != "base":
        dev["pre_ball_states"] = dev["pre_ball_states"][:, [7, 9]].contiguous()
    sess = HostSession(cfg, st, num_chunks=2, pin=pin)
    gen = torch.Generator().manual_seed(8)
    try:
        for step in range(5):
            vel = torch.randn(sess.state["reset_ball_vel"].shape, generator=gen)
            sess.state["reset_ball_vel"].copy_(vel)               # in place: the pointer the session saw before
            dev["reset_ball_vel"].copy_(vel.to(DEV))
            if variant == "adof":
                yz = torch.randn(n, 2, generator=gen)
                sess.state["reset_ball_pos_yz"].copy_(yz)
                dev["reset_ball_pos_yz"].copy_(yz.to(DEV))
            dev["stats"].zero_()
            sess.post_physics_step(N.PHASE_ALL)
            run(cfg, dev, N.PHASE_ALL)
            for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names:
                assert torch.equal(sess.state[name], dev[name].cpu()), f"{variant} step {step}: {name}"
            want = dev["stats"].sum(dim=0).cpu()
            got = sess.state["stats"]
            assert float(got[1]) == float(want[1]) and float(got[2]) == float(want[2])
            assert abs(float(got[0]) - float(want[0])) <= 1e-9 * max(1.0, abs(float(want[0])))
        assert int(dev["reset_buf"].sum()) > 0
    finally:
        sess.close()


def test_adof_first_design_still_passes():
    """ADOF's full fused step runs on ppk_adof2.cuh; tail envs, other phase subsets and unaligned tensors stay on the
    first design (ppk_adof.cuh).  PPK_ADOF_V1=1 (read once per process) forces every env onto the first design."""
    import os
    import subprocess
    import sys
    if os.environ.get("PPK_ADOF_V1") is not None:
        pytest.skip("already inside the forced run")
    env = dict(os.environ, PPK_ADOF_V1="1")
    here = os.path.abspath(__file__)
    out = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-k",
                          "(test_fused_step_matches_oracle or test_ragged_sizes or test_multi_step_trajectory or "
                          "test_golden_step_fixture or test_golden_vectors or test_adof_compact_reference_pose) and adof"],
                         env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(os.path.dirname(here)))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def test_small_batches_also_pass_on_the_32_env_tiles():
    """Batches below one wave of 32-env tiles run on 16-env tiles (ppk_api.cu small_batch), so most tests of this
    file exercise those; PPK_SMALL_TILES=0 (read once per process) forces the 32-env instantiations onto the same
    small / ragged cases, tail tiles included."""
    import os
    import subprocess
    import sys
    if os.environ.get("PPK_SMALL_TILES") is not None:
        pytest.skip("already inside the forced run")
    env = dict(os.environ, PPK_SMALL_TILES="0")
    here = os.path.abspath(__file__)
    out = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-k",
                          "(test_fused_step_matches_oracle or test_ragged_sizes or test_multi_step_trajectory or "
                          "test_golden_step_fixture) and (tilt or a3 or nes or align)"],
                         env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(os.path.dirname(here)))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
