"""Synthetic PhysX state: the tensors `gym.acquire_*_tensor` + `gymtorch.wrap_tensor`
would hand to the task (TILT:131-134,153-174,208-214), filled from a seeded
generator.  PhysX itself (the articulation solver) is out of scope; the task hot
path only ever sees these tensors.

Layouts (all fp32, contiguous, AoS):
  rigid_body_states [N,B,13]  root_states [N,A,13]  dof_states [N,D,2]  dof_forces [N,D]
  row = pos(3) quat xyzw(4) linvel(3) angvel(3)

Value distributions follow SURVEY.md 8(d): chosen so that every branch of every
reward fires with non-trivial probability.
"""
import math
from typing import Dict, Optional

import torch

from .config import TaskConfig

State = Dict[str, torch.Tensor]


def _unit_quat(n, g, device):
    q = torch.randn(n, 4, generator=g, device=device)
    return q / q.norm(dim=-1, keepdim=True).clamp_min(1e-6)


def _rows(n, rows, g, device, pos_mean=(0.0, 0.0, 1.0), pos_std=0.4):
    """[n, rows, 13] random rigid-body style rows."""
    t = torch.empty(n, rows, 13, device=device)
    t[..., 0:3] = torch.randn(n, rows, 3, generator=g, device=device) * pos_std + torch.tensor(pos_mean, device=device)
    t[..., 3:7] = _unit_quat(n * rows, g, device).view(n, rows, 4)
    t[..., 7:13] = torch.randn(n, rows, 6, generator=g, device=device)
    return t


def _ball_row(n, g, device):
    """ball x~U(-0.4,3.6) y~U(-0.8,0.8) z~U(0,1.4), v~N(0,4^2) per component."""
    u = lambda lo, hi: torch.rand(n, generator=g, device=device) * (hi - lo) + lo
    row = torch.zeros(n, 13, device=device)
    row[:, 0], row[:, 1], row[:, 2] = u(-0.4, 3.6), u(-0.8, 0.8), u(0.0, 1.4)
    row[:, 6] = 1.0
    row[:, 7:10] = torch.randn(n, 3, generator=g, device=device) * 4.0
    row[:, 10:13] = torch.randn(n, 3, generator=g, device=device)
    return row


def sample_ball_launch(cfg: TaskConfig, n: int, g, device) -> torch.Tensor:
    """Vectorised form of `generate_random_speed_for_ball` (TILT:307-318, NES:312-323,
    ADOF:357-367, A3:300-302): per-env launch velocity [n,3] consumed when an env
    resets.  Same ranges and formulas as the reference's host `random.uniform`
    draws; a device stream cannot be bit-identical to Mersenne-Twister, so the
    velocities are an explicit input of the reset path ("RNG in reset")."""
    u = lambda lo, hi: torch.rand(n, generator=g, device=device, dtype=torch.float64) * (hi - lo) + lo
    rad = math.pi / 180.0
    v = cfg.variant
    if v in ("tilt", "a4", "align", "align2"):
        s = -u(8.0, 8.8 if v in ("align", "align2") else 8.6)
        a, z = u(-5.0, 5.0) * rad, u(2.0, 10.0) * rad
        out = torch.stack((s * a.cos() * z.cos(), s * a.sin() * z.sin(), s * a.sin()), dim=-1)
    elif v in ("nes", "adof"):
        s = u(5.4, 5.9) if v == "nes" else u(5.0, 5.4)
        a = (u(-5.0, 5.0) if v == "nes" else u(-8.0, 3.0)) * rad
        z = (u(10.0, 17.0) if v == "nes" else u(14.0, 24.0)) * rad
        out = torch.stack((-s * a.cos() * z.cos(), s * a.sin() * z.cos(), s * z.sin()), dim=-1)
    else:  # a3, base (BASE:250-266)
        s = -u(6.5, 7.5)
        a = u(-5.0, 5.0) * rad
        out = torch.stack((s * a.cos(), s * a.sin(), torch.zeros_like(s)), dim=-1)
    return out.to(torch.float32)


def make_state(cfg: TaskConfig, num_envs: int, seed: int = 0, device: str = "cpu",
               adversarial: bool = True) -> State:
    """Allocate and fill every tensor one task step reads or writes."""
    n, A, B, D = num_envs, cfg.num_actors, cfg.num_bodies, cfg.num_dofs
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dev = device
    st: State = {}

    rb = _rows(n, B, g, dev)
    root = _rows(n, A, g, dev, pos_std=0.02)
    # humanoid roots: (0,0,1) and (3.5,0,1) +- noise
    root[:, cfg.humanoid_actor[0], 0:3] = torch.tensor((0.0, 0.0, 1.0), device=dev) + 0.05 * torch.randn(n, 3, generator=g, device=dev)
    if cfg.variant in ("a4", "align2", "base"):
        root[:, cfg.humanoid_actor[1], 0:3] = torch.tensor((3.5, 0.0, 1.0), device=dev) + 0.05 * torch.randn(n, 3, generator=g, device=dev)
    root[:, cfg.ball_actor] = _ball_row(n, g, dev)
    if cfg.variant == "base":
        root[:, 4] = _ball_row(n, g, dev)
    pre_ball = _ball_row(n, g, dev)

    # paddles: near the ball for half of the envs so distance / hit terms are exercised
    ball_for = {0: cfg.ball_actor}
    if cfg.variant == "base":
        ball_for = {0: 4, 1: 3}           # paddle1<->ball2, paddle2<->ball1 (BASE:634-647)
    elif cfg.variant in ("a4", "align2"):
        ball_for = {0: 3, 1: 3}
    for k, ball_row in ball_for.items():
        near = torch.rand(n, generator=g, device=dev) < 0.5
        jitter = 0.15 * torch.randn(n, 3, generator=g, device=dev)
        p = cfg.paddle_body[k]
        rb[:, p, 0:3] = torch.where(near.unsqueeze(-1), root[:, ball_row, 0:3] + jitter, rb[:, p, 0:3])

    if cfg.variant == "a3":
        # A3 terminates an env as soon as the ball is behind the paddle (A3:1149,1160): keep that a
        # ~5 % event, and launch reset balls beyond every paddle, so the batch is not one perpetual reset
        p = cfg.paddle_body[0]
        keep = torch.rand(n, generator=g, device=dev) < 0.95
        ahead = torch.minimum(rb[:, p, 0], root[:, cfg.ball_actor, 0] - 0.01)
        rb[:, p, 0] = torch.where(keep, ahead, rb[:, p, 0])

    dof_scale = 1.0
    if cfg.variant == "adof":
        dof_scale = 0.05
    dof = torch.randn(n, D, 2, generator=g, device=dev) * dof_scale
    force = torch.randn(n, D, generator=g, device=dev)

    st["rigid_body_states"] = rb
    st["root_states"] = root
    st["dof_states"] = dof
    st["dof_forces"] = force
    st["pre_ball_states"] = pre_ball

    # reset sources
    init_root = _rows(n, A, g, dev, pos_std=0.02)
    init_root[..., 7:13] = 0.0
    if cfg.variant == "a3":
        init_root[:, cfg.ball_actor, 0] += 3.7
    st["initial_root_states"] = init_root
    if cfg.variant == "adof":
        st["initial_dof_states"] = torch.zeros(n, D, 2, device=dev)            # ADOF:249
    else:
        st["initial_dof_states"] = torch.randn(n, D, 2, generator=g, device=dev) * 0.1
    st["reset_ball_vel"] = sample_ball_launch(cfg, n, g, dev)
    if cfg.variant == "adof":
        u = lambda lo, hi: torch.rand(n, generator=g, device=dev) * (hi - lo) + lo
        st["reset_ball_pos_yz"] = torch.stack((u(-0.5, 0.1), u(0.96, 1.05)), dim=-1)   # ADOF:127-128,976-979
        # reference pose: the initial rigid-body states with balance-body velocities zeroed (ADOF:196,200)
        init_rb = rb.clone()
        bal = torch.tensor(cfg.balance_ids, device=dev)
        init_rb[:, bal, 7:10] = 0.0
        st["initial_body_states"] = init_rb
        # current pose = reference + per-env deviation; sigma spread so has_fallen (mean dist > 0.32) splits the batch
        sigma = torch.rand(n, 1, 1, generator=g, device=dev) * 0.4
        rb[:, :40, 0:3] = init_rb[:, :40, 0:3] + sigma * torch.randn(n, 40, 3, generator=g, device=dev)
        rb[:, :40, 7:10] = 0.5 * torch.randn(n, 40, 3, generator=g, device=dev)
        rb[:, cfg.pelvis_body, 2] = 1.0 + 0.05 * torch.randn(n, generator=g, device=dev)   # around the 0.97 threshold
        near = torch.rand(n, generator=g, device=dev) < 0.5
        jitter = 0.15 * torch.randn(n, 3, generator=g, device=dev)
        rb[:, cfg.paddle_body[0], 0:3] = torch.where(near.unsqueeze(-1), root[:, cfg.ball_actor, 0:3] + jitter,
                                                     rb[:, cfg.paddle_body[0], 0:3])
        # ball z concentrated near the table-height window 0.82..0.83 for a share of envs
        tbl = torch.rand(n, generator=g, device=dev) < 0.2
        root[:, cfg.ball_actor, 2] = torch.where(tbl, 0.815 + 0.02 * torch.rand(n, generator=g, device=dev),
                                                 root[:, cfg.ball_actor, 2])
        # and near the net window 1.72..1.78
        net = torch.rand(n, generator=g, device=dev) < 0.2
        root[:, cfg.ball_actor, 0] = torch.where(net, 1.70 + 0.1 * torch.rand(n, generator=g, device=dev),
                                                 root[:, cfg.ball_actor, 0])

    # VecTask buffers (upstream allocate_buffers dtypes)
    L = cfg.max_episode_length
    st["progress_buf"] = torch.randint(0, L, (n,), generator=g, device=dev, dtype=torch.int64)
    st["reset_buf"] = torch.zeros(n, dtype=torch.int64, device=dev)
    st["rew_buf"] = torch.zeros((n, cfg.obs_rows) if cfg.obs_rows > 1 else (n,), device=dev)
    st["obs_buf"] = torch.zeros((n, cfg.obs_rows, cfg.num_obs) if cfg.obs_rows > 1 else (n, cfg.num_obs), device=dev)
    for name, reset_val in zip(cfg.flag_names, cfg.flag_reset_values):
        p = 0.75 if reset_val else 0.25
        st[name] = torch.rand(n, generator=g, device=dev) < p
    for name in cfg.counter_names:
        st[name] = torch.rand(n, generator=g, device=dev) < 0.25
    for name in cfg.state_names:          # last_hitter in {1, 2}
        st[name] = torch.randint(1, 3, (n,), generator=g, device=dev, dtype=torch.int64)

    # actions and PD scaling (TILT:666-671: offset/scale = 0.5*(hi +- lo) of the DOF limits)
    lo = -1.0 - torch.rand(D, generator=g, device=dev)
    hi = 1.0 + torch.rand(D, generator=g, device=dev)
    st["pd_action_offset"] = 0.5 * (hi + lo)
    st["pd_action_scale"] = 0.5 * (hi - lo)
    st["actions"] = torch.rand(n, D, generator=g, device=dev) * 2.0 - 1.0
    st["pd_targets"] = torch.zeros(n, D, device=dev)
    st["actor_indices"] = torch.arange(n * A, dtype=torch.int64, device=dev)               # TILT:645
    dof_per = 2 if cfg.variant in ("a4", "align2") else 1                                               # A4:889
    st["dof_indices"] = torch.arange(n * dof_per, dtype=torch.int64, device=dev)

    if adversarial and n >= 64:
        _plant_adversarial(cfg, st)
    return st


def _plant_adversarial(cfg: TaskConfig, st: State) -> None:
    """Deterministic edge cases in the first envs: every threshold constant exactly and
    +-1 ulp, vx = 0, pre_vx = 0, atan2(0,0) heading."""
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    root, pre = st["root_states"], st["pre_ball_states"]
    b = cfg.ball_actor
    i = 0

    def ulps(v):
        c = f32(v)
        return [float(torch.nextafter(c, f32(-1e30))), float(c), float(torch.nextafter(c, f32(1e30)))]

    x_thr = (-3.1, 0.4, 1.06, 1.3, 1.7, 1.72, 1.78, 1.8, 1.9, 2.2, 2.44, 2.5, 3.1)
    y_thr = (-0.6, -0.4, 0.4, 0.6)
    z_thr = (0.1, 0.78, 0.82, 0.83, 0.96, 0.98, 1.14, 1.25)
    for col, thrs in ((0, x_thr), (1, y_thr), (2, z_thr)):
        for t in thrs:
            for v in ulps(t):
                if i >= root.shape[0]:
                    return
                root[i, b, col] = v
                if col != 0:
                    root[i, b, 0] = 2.7        # inside the far table so y/z tests decide
                if col == 2:
                    root[i, b, 1] = 0.0
                root[i, b, 7] = 3.0 if (i % 2 == 0) else -3.0
                i += 1
    for v in (0.0, -0.0, 1.0, 1.5):
        for pv in (0.0, -0.0, -1.0, 1.0):
            if i >= root.shape[0]:
                return
            root[i, b, 7] = v
            pre[i, 7] = pv
            i += 1
    # degenerate heading: rotated x-axis has zero xy projection -> atan2(0, 0)
    rb = st["rigid_body_states"]
    s = math.sqrt(0.5)
    for q in ((0.0, s, 0.0, s), (0.0, -s, 0.0, s), (0.0, 0.0, 0.0, 1.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 0.0)):
        if i >= root.shape[0]:
            return
        rb[i, 0, 3:7] = f32(q)
        if cfg.variant in ("a4", "align2", "base") and rb.shape[1] > 40:
            rb[i, 40, 3:7] = f32(q)
        i += 1
    # time-out boundary: progress L-3, L-2 (resets after the +1), L-1
    L = cfg.max_episode_length
    for p in (L - 3, L - 2, L - 1, 0):
        if i >= root.shape[0]:
            return
        st["progress_buf"][i] = p
        i += 1


def clone_state(st: State, device: Optional[str] = None) -> State:
    return {k: (v.to(device).clone() if device is not None else v.clone()) for k, v in st.items()}
