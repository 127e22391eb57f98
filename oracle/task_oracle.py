"""TEST INFRASTRUCTURE ONLY -- task-level composition of the oracle functions.

Mirrors what the reference's class wrappers do around the free functions
(`compute_reward` TILT:739-768, `compute_observations` TILT:770-799,
`pre_physics_step` TILT:1002-1020, `post_physics_step` TILT:1022-1052 and their
counterparts in the other six files) on a plain dict of CPU tensors laid out
like `isaacgym_b200.synth.make_state` produces.  The class methods themselves
cannot run without PhysX, so their (short) bodies are restated here; all
arithmetic lives in `pingpong_oracle.py`.

`impl` selects who does the arithmetic: the restated oracle (default) or the
reference's own source via `ref_extract` (build container only).
"""
import torch

from . import pingpong_oracle as O


class OracleImpl:
    """Arithmetic provider backed by the restatement."""
    humanoid_observations = staticmethod(O.humanoid_observations)
    imitation_observations = staticmethod(O.imitation_observations)
    base_observations = staticmethod(lambda obs_buf, r1, r2, p1, p2, b1, b2: O.base_observations(p1, p2, b1, b2))
    base_reward = staticmethod(O.base_reward)
    a3_reward = staticmethod(O.a3_reward)
    tilt_reward = staticmethod(O.tilt_reward)
    a4_h1_reward = staticmethod(lambda *a: O.tilt_reward(*a, scripted=True))
    a4_h2_reward = staticmethod(lambda *a: O.tilt_reward(*a, mirrored=True, scripted=True))
    nes_reward = staticmethod(O.nes_reward)
    align_reward = staticmethod(O.align_reward)
    adof_reward = staticmethod(O.adof_reward)
    align2_reward = staticmethod(O.align2_reward)

    @staticmethod
    def pingpong_observations(body_states, ids, ball, adof=False):
        return O.pingpong_observations(body_states, ids, ball, y_intersect=adof)


class ReferenceImpl:
    """Arithmetic provider backed by the reference's own source (build container only)."""

    def __init__(self):
        from . import ref_extract as R
        self._R = R
        ld = R.load
        self.humanoid_observations = ld("TILT", "compute_humanoid_observations")
        self._pp_obs = ld("TILT", "compute_pingpong_observations")
        self._pp_obs_adof = ld("ADOF", "compute_pingpong_observations")
        self.imitation_observations = ld("ADOF", "compute_imitation_observations")
        self.base_observations = ld("BASE", "compute_pingpong_observations")
        self.base_reward = ld("BASE", "compute_pingpong_reward")
        self.a3_reward = ld("A3", "compute_pingpong_reward")
        self.tilt_reward = ld("TILT", "compute_pingpong_reward_nv")
        self.a4_h1_reward = ld("A4", "compute_humanoid1_pingpong_reward")
        self.a4_h2_reward = ld("A4", "compute_humanoid2_pingpong_reward")
        self.nes_reward = ld("NES", "compute_pingpong_reward_only_paddle")
        self.align_reward = ld("ALIGN", "compute_pingpong_reward", 0)
        self._adof = R.load_adof_reward()
        self.align2_reward = R.load_align_two_humanoid_reward()

    def pingpong_observations(self, body_states, ids, ball, adof=False):
        return (self._pp_obs_adof if adof else self._pp_obs)(body_states, ids, ball)

    def adof_reward(self, *a, is_train=True):
        # ADOF:802-858 passes (..., body_balance_states_id, is_g1, is_train)
        return self._adof(*a, True, is_train)


def _ids(t):
    return torch.tensor(t, dtype=torch.long)


def _f(x):
    # the reference passes YAML numbers through to the scripted functions annotated `float`
    return float(x)


def compute_observations(cfg, st, impl=OracleImpl):
    """`compute_observations` wrappers: TILT:770-799, A4:773-803 (one row per
    humanoid, defect D5), ADOF:867-904, BASE:493-510.  Returns the new obs_buf."""
    rb, root, dof = st["rigid_body_states"], st["root_states"], st["dof_states"]
    dof_pos, dof_vel = dof[..., 0], dof[..., 1]
    v = cfg.variant
    if v == "base":
        p1, p2 = rb[:, cfg.paddle_body[0], :], rb[:, cfg.paddle_body[1], :]
        return impl.base_observations(st["obs_buf"], p1, p2, p1, p2, root[:, 3, :], root[:, 4, :])
    ball = root[:, cfg.ball_actor, :]
    if v in ("a4", "align2"):
        rows = []
        for ids in (cfg.body_ids, cfg.body_ids_2):
            ids = _ids(ids)
            pp = impl.pingpong_observations(rb, ids, ball)
            hu = impl.humanoid_observations(rb, dof_pos, dof_vel, ids)
            rows.append(torch.cat([hu, pp], dim=-1))
        return torch.stack(rows, dim=1)
    ids = _ids(cfg.body_ids)
    pp = impl.pingpong_observations(rb, ids, ball, adof=(v == "adof"))
    hu = impl.humanoid_observations(rb, dof_pos, dof_vel, ids)
    if v != "adof":
        return torch.cat([hu, pp], dim=-1)
    init_dof = st["initial_dof_states"]
    imi = impl.imitation_observations(rb, st["initial_body_states"], init_dof[..., 0], init_dof[..., 1],
                                      _ids(cfg.balance_ids))
    return torch.cat([hu, pp, imi], dim=-1)


def compute_reward(cfg, st, impl=OracleImpl):
    """`compute_reward` wrappers (TILT:739-759, A3:720-737, NES:745-760, ALIGN:736-753,
    ADOF:802-858, BASE:463-472; A4 per defect D4 = humanoid-1 and humanoid-2 rewards).
    Writes rew_buf / reset_buf and mutates the flag tensors in `st` in place."""
    rb, root, dof = st["rigid_body_states"], st["root_states"], st["dof_states"]
    dof_vel = dof[..., 1]
    force, reset_buf, progress = st["dof_forces"], st["reset_buf"], st["progress_buf"]
    L = _f(cfg.max_episode_length)
    v = cfg.variant
    if v == "base":
        rew, rst = impl.base_reward(rb[:, cfg.paddle_body[0], :], rb[:, cfg.paddle_body[1], :],
                                    root[:, 3, :], root[:, 4, :], reset_buf, progress, L)
        st["rew_buf"][:], st["reset_buf"][:] = rew, rst
        return
    h_root = root[:, cfg.humanoid_actor[0], :]
    paddle = rb[:, cfg.paddle_body[0], :]
    pre_ball, ball = st["pre_ball_states"], root[:, cfg.ball_actor, :]
    common = (h_root, paddle, pre_ball, ball, force, dof_vel, reset_buf, progress, L,
              _f(cfg.alpha), _f(cfg.power_coefficient))
    if v == "a3":
        rew, rst = impl.a3_reward(*common, _f(cfg.penalty))
    elif v == "tilt":
        rew, rst = impl.tilt_reward(*common, _f(cfg.penalty), st["condition_calculated"], _f(cfg.hit_table_reward),
                                    _f(cfg.not_hit_table_penalty), st["reward_calculated"],
                                    st["no_bounce_before_half_mask"])
    elif v == "nes":
        rew, rst = impl.nes_reward(*common, _f(cfg.penalty), st["paddle_condition_calculated"],
                                   st["missed_ball_calculated"])
    elif v == "align":
        rew, rst = impl.align_reward(*common, _f(cfg.penalty), _f(cfg.hit_table_reward),
                                     _f(cfg.not_hit_table_penalty), st["reward_calculated"])
    elif v == "a4":
        rew1, rst = impl.a4_h1_reward(*common, _f(cfg.penalty), st["condition_calculated"],
                                      _f(cfg.hit_table_reward), _f(cfg.not_hit_table_penalty),
                                      st["reward_calculated"], st["no_bounce_before_half_mask"])
        common2 = (root[:, cfg.humanoid_actor[1], :], rb[:, cfg.paddle_body[1], :]) + common[2:]
        rew2, rst2 = impl.a4_h2_reward(*common2, _f(cfg.penalty), st["condition_calculated_2"],
                                       _f(cfg.hit_table_reward), _f(cfg.not_hit_table_penalty),
                                       st["reward_calculated_2"], st["no_bounce_before_half_mask_2"])
        assert torch.equal(rst, rst2)
        rew = torch.stack((rew1, rew2), dim=-1)
    elif v == "align2":
        r1, r2, rst, lh = impl.align2_reward(
            h_root, paddle, root[:, cfg.humanoid_actor[1], :], rb[:, cfg.paddle_body[1], :], pre_ball, ball, force,
            dof_vel, reset_buf, progress, L, _f(cfg.alpha), _f(cfg.power_coefficient), _f(cfg.penalty),
            _f(cfg.hit_table_reward), _f(cfg.not_hit_table_penalty), st["reward_calculated"], st["last_hitter"])
        rew = torch.stack((r1, r2), dim=-1)
        st["last_hitter"][:] = lh            # the caller keeps the returned tensor
    elif v == "adof":
        init_dof = st["initial_dof_states"]
        out = impl.adof_reward(
            h_root, rb[:, cfg.pelvis_body, :], paddle, pre_ball, ball, force, dof_vel, reset_buf, progress, L,
            _f(cfg.alpha), _f(cfg.power_coefficient), st["paddle_condition_calculated"], _f(cfg.hit_paddle_reward),
            _f(cfg.miss_paddle_penalty_coefficient), _f(cfg.cross_net_reward), _f(cfg.hit_table_reward),
            _f(cfg.not_hit_table_penalty), st["hit_table_calculated"], _f(cfg.die_penalty),
            st["die_penalty_calculated"], st["humanoid_die_calculated"], st["closer_to_paddle_count"],
            st["hit_paddle_count"], st["cross_net_count"], st["hit_table_count"], st["fall_down_count"],
            init_dof[..., 0], dof[..., 0], init_dof[..., 1], rb, st["initial_body_states"], _ids(cfg.balance_ids),
            is_train=cfg.is_train)
        rew, rst = out[0], out[1]
        names = ("paddle_condition_calculated", "hit_table_calculated", "die_penalty_calculated",
                 "humanoid_die_calculated") + O.ADOF_COUNTERS
        for name, val in zip(names, out[2:]):
            st[name][:] = val                    # ADOF:802: `self.x[:] = ...` casts back to bool
    else:
        raise ValueError(v)
    st["rew_buf"][:], st["reset_buf"][:] = rew, rst


def pre_physics_step(cfg, st):
    """TILT:1002-1020 (A4 with the offset/scale tiled over both humanoids, defect D6)."""
    st["pd_targets"] = O.pd_targets(st["pd_action_offset"], st["pd_action_scale"], st["actions"].clone())
    if cfg.variant != "base":
        st["pre_ball_states"] = st["root_states"][:, cfg.ball_actor, :].clone()


def reset_idx(cfg, st, env_ids):
    """Per-env launch velocities come from the pre-sampled `reset_ball_vel[N,3]` rows
    (and `reset_ball_pos_yz[N,2]` for ADOF) of the resetting envs."""
    if cfg.variant == "base":
        raise NotImplementedError("BASE reset takes one velocity pair for all envs: use O.base_reset_idx")
    yz = st["reset_ball_pos_yz"][env_ids] if cfg.variant == "adof" else None
    return O.reset_idx(cfg.variant, st, env_ids, st["reset_ball_vel"][env_ids], yz)


def post_physics_step(cfg, st, impl=OracleImpl):
    """TILT:1022-1052 (identical order in A3/NES/ALIGN/A4; ADOF:1149-1192 also clears the
    five counters when any env resets): progress += 1 -> reward -> reset_idx(nonzero(reset_buf))
    -> observations (which see the reset root/DOF tensors but stale rigid bodies)."""
    assert cfg.variant != "base"
    st["progress_buf"] += 1
    compute_reward(cfg, st, impl)
    stats = step_stats(cfg, st)          # the reference logs here: after reward, before reset (TILT:763-766)
    env_ids = st["reset_buf"].nonzero(as_tuple=False).flatten()
    idx = None
    if len(env_ids) > 0:
        idx = reset_idx(cfg, st, env_ids)
        for name in cfg.counter_names:
            st[name].fill_(0)
    st["obs_buf"][:] = compute_observations(cfg, st, impl)
    return env_ids, idx, stats


def step_stats(cfg, st):
    """The scalar episode statistics the reference prints (TILT:763-766, ADOF:1164-1168), fp64."""
    out = {"reward_sum": st["rew_buf"].double().sum(dim=0), "progress_sum": st["progress_buf"].double().sum(),
           "reset_count": st["reset_buf"].double().sum()}
    for name in cfg.counter_names:
        out[name] = st[name].double().sum()
    return out
