"""Host-side mirror of the reference's task classes (tasks/humanoid_pingpong*.py) over libppk.so.

Same method names, argument meaning and call order as the reference:
    pre_physics_step(actions)   TILT:1002-1020
    post_physics_step()         TILT:1022-1052   (one fused kernel here)
    compute_reward(actions)     TILT:739-768
    compute_observations()      TILT:770-799
    reset_idx(env_ids)          TILT:809/847-906
and the same attribute names for the buffers and flag tensors (`obs_buf`, `rew_buf`, `reset_buf`,
`progress_buf`, `condition_calculated`, ...).  PhysX is replaced by a dict of synthetic state
tensors (`synth.make_state`); everything numerical happens in the CUDA library -- there is no
torch or CPU fallback, a missing library raises.
"""
from typing import Dict, Optional

import torch

from . import _native as N
from .config import CONFIGS, TaskConfig
from .stats import EpisodeStats
from .vec_task import VecTask

_STATE_KEYS = ("rigid_body_states", "root_states", "dof_states", "dof_forces", "initial_root_states",
               "initial_dof_states", "initial_body_states", "reset_ball_vel", "reset_ball_pos_yz",
               "pd_action_offset", "pd_action_scale", "actor_indices", "dof_indices")


def pack_reference_pose(cfg: TaskConfig, initial_body_states: torch.Tensor) -> torch.Tensor:
    """initial_body_states[:, balance_ids][..., (0,1,2,7,8,9)] as one contiguous [N, n_balance, 6] tensor --
    everything compute_imitation_reward / compute_imitation_observations read of the reference pose
    (ADOF:1345-1349, ADOF:1908-1909).  `PpkBuffers.initial_balance_states`."""
    ids = torch.tensor(cfg.balance_ids, dtype=torch.long, device=initial_body_states.device)
    rows = initial_body_states[:, ids]
    return torch.cat([rows[..., 0:3], rows[..., 7:10]], dim=-1).contiguous()


class PingpongTask(VecTask):
    """Generic task over one variant; the named subclasses below only pick the variant."""
    variant = "tilt"

    def __init__(self, sim_state: Dict[str, torch.Tensor], cfg: Optional[TaskConfig] = None,
                 device: str = "cuda:0", fused: bool = True, full_pre_ball_clone: bool = False,
                 log_stats: bool = False, envelope: bool = False, compact_reference_pose: bool = True,
                 launch_seed: Optional[int] = 0, env_offset: int = 0, fused_moments: bool = False, **kw):
        """`launch_seed`: the reference draws a fresh ball launch velocity inside every reset (TILT:857-862).  Here the
        fused reset consumes row n of the per-env launch table `reset_ball_vel`; after every `post_physics_step` the rows
        just consumed are redrawn on the device (`ppk_sample_ball_launch`, Philox keyed by (launch_seed, env, step)), so
        no env ever replays a launch.  `launch_seed=None` leaves the table alone: it is then an explicit input the
        caller refills (parity tests feed the oracle the same table).  `env_offset`: first global env id of this shard
        (data-parallel ranks draw disjoint counter ranges).  `fused_moments`: the fused step also leaves the fp64
        column moments of the obs rows it wrote in `self.obs_moments` (`PPK_PHASE_MOMENTS`), for
        `RunningMeanStd.update_from_step` -- the normaliser's update then never re-reads obs_buf."""
        cfg = cfg or CONFIGS[self.variant]
        n = sim_state["root_states"].shape[0]
        super().__init__(cfg, n, device=device, **kw)
        if self.device.type != "cuda":
            raise RuntimeError("the ping-pong task step is CUDA-only (no CPU fallback)")
        self._lib = N.load()
        self._task = N.make_task(cfg)
        self.fused = fused
        self.launch_seed = launch_seed
        self.env_offset = env_offset
        self.fused_moments = bool(fused_moments) and cfg.variant not in ("base", "adof")
        if self.fused_moments:
            self.obs_moments = torch.zeros(N.PPK_MOMENT_SLOTS, 2 * cfg.num_obs, dtype=torch.float64, device=self.device)
        self.log_stats = log_stats
        self.envelope = envelope          # also emit timeout_buf + compacted reset index lists (VecTask.step envelope)
        self._count_zeroed = False
        dev = self.device
        # "acquired" PhysX tensors (zero-copy views in the reference, TILT:153-174)
        self.st: Dict[str, torch.Tensor] = {}
        for key in _STATE_KEYS:
            if key in sim_state:
                self.st[key] = sim_state[key].to(dev).contiguous()
        if cfg.balance_ids and "initial_body_states" in self.st and compact_reference_pose:
            # the imitation reference pose is a constant of the task (ADOF:196-200): repack the fields
            # the step reads of it once, [N,23,6] = (pos, linvel) of the balance bodies
            self.st["initial_balance_states"] = pack_reference_pose(cfg, self.st["initial_body_states"])
        self.root_states = self.st["root_states"]
        self.vec_root_states = self.root_states
        self.body_states = self.st["rigid_body_states"]
        self.vec_dof_states = self.st["dof_states"]
        self.dof_pos = self.vec_dof_states[..., 0]
        self.dof_vel = self.vec_dof_states[..., 1]
        self.dof_force_tensor = self.st["dof_forces"]
        self.ball2_root_states = self.vec_root_states[:, cfg.ball_actor, :]
        self.initial_vec_root_states = self.st["initial_root_states"]
        self.initial_dof_states = self.st["initial_dof_states"]
        self._pd_action_offset = self.st["pd_action_offset"]
        self._pd_action_scale = self.st["pd_action_scale"]
        # VecTask buffers may be seeded from the synthetic state (tests do; a fresh task starts at zero)
        for key in ("progress_buf", "reset_buf"):
            if key in sim_state:
                getattr(self, key).copy_(sim_state[key].to(dev))
        # flag tensors (TILT:241-243, NES:244-248, ALIGN:239, ADOF:279-293)
        for name, reset_val in zip(cfg.flag_names, cfg.flag_reset_values):
            t = sim_state[name].to(dev).clone() if name in sim_state else torch.full((n,), reset_val, dtype=torch.bool, device=dev)
            setattr(self, name, t)
        for name in cfg.counter_names:
            t = sim_state[name].to(dev).clone() if name in sim_state else torch.zeros(n, dtype=torch.bool, device=dev)
            setattr(self, name, t)
        for name in cfg.state_names:          # ALIGN def #2: last_hitter, initialised to 2 (ALIGN:1253)
            t = sim_state[name].to(dev).clone() if name in sim_state else torch.full((n,), 2, dtype=torch.int64, device=dev)
            setattr(self, name, t)
        # saved pre-step ball state: the reference clones the 13-float row (TILT:1020) although only
        # vx (and vz in ALIGN) is read back; by default only those two floats are kept
        width = 13 if full_pre_ball_clone else 2
        self.pre_ball2_root_states = torch.zeros(n, width, device=dev)
        if "pre_ball_states" in sim_state:
            src = sim_state["pre_ball_states"].to(dev)
            self.pre_ball2_root_states.copy_(src if width == 13 else src[:, [7, 9]])
        self.actions = torch.zeros(n, cfg.num_dofs, device=dev)
        self.pd_tar = torch.zeros(n, cfg.num_dofs, device=dev)
        self.stats = EpisodeStats(dev)
        self._scratch = torch.zeros(16, dtype=torch.int32, device=dev)
        # VecTask.step envelope: compacted int32 index lists of the envs reset by the fused step
        # (what gym.set_actor_root_state_tensor_indexed / set_dof_state_tensor_indexed take, TILT:876-888)
        self.reset_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.reset_actor_indices = torch.zeros(max(n * cfg.num_actors, 1), dtype=torch.int32, device=dev)
        dof_per = self.st["dof_indices"].numel() // max(n, 1) if "dof_indices" in self.st else 0
        self.reset_dof_indices = torch.zeros(max(n * dof_per, 1), dtype=torch.int32, device=dev)
        self._buffers = None

    # -- plumbing -----------------------------------------------------------------------------------
    def _tensor_dict(self):
        d = dict(self.st)
        d.update(obs_buf=self.obs_buf, rew_buf=self.rew_buf, reset_buf=self.reset_buf, progress_buf=self.progress_buf,
                 pre_ball_states=self.pre_ball2_root_states, actions=self.actions, pd_targets=self.pd_tar,
                 stats=self.stats.slots, scratch=self._scratch, clip_actions=self.clip_actions,
                 clip_observations=self.clip_obs)
        if self.fused_moments:
            d.update(obs_moments=self.obs_moments)
        if self.envelope:
            d.update(timeout_buf=self.timeout_buf)
        if self.envelope and "actor_indices" in self.st and "dof_indices" in self.st:
            d.update(reset_count=self.reset_count, reset_actor_indices=self.reset_actor_indices,
                     reset_dof_indices=self.reset_dof_indices)
        for name in self.cfg.flag_names + self.cfg.counter_names + self.cfg.state_names:
            d[name] = getattr(self, name)
        return d

    def buffers(self) -> N.PpkBuffers:
        if self._buffers is None:
            self._buffers = N.make_buffers(self.cfg, self._tensor_dict())
        return self._buffers

    def _stream(self):
        return N.current_stream_ptr(self.device)

    def _step(self, phases: int):
        N.check(self._lib.ppk_post_physics_step(self._task, self.buffers(), phases, self._stream()),
                "ppk_post_physics_step")

    # -- the reference's method set ----------------------------------------------------------------
    def pre_physics_step(self, actions):
        """TILT:1002-1020: keep the actions, pd_tar = offset + scale*actions, save the ball state."""
        self.actions.copy_(actions.to(self.device))
        N.check(self._lib.ppk_pre_physics_step(self._task, self.buffers(), self._stream()), "ppk_pre_physics_step")
        self._count_zeroed = True         # the pre-step kernel zeroes reset_count

    def compute_reward(self, actions=None):
        """TILT:739-768 (the statistics the reference prints every `log_every` steps are accumulated
        on the device instead of `.item()`-ed)."""
        self._step(N.PHASE_REWARD | (N.PHASE_STATS if self._log_now() else 0))

    def compute_observations(self):
        self._step(N.PHASE_OBS)

    def reset_idx(self, env_ids, ball_vel=None, ball_pos_yz=None):
        """TILT:809-906.  `ball_vel` [k,3] are the launch velocities the reference draws with host
        `random.uniform`; by default rows `env_ids` of the pre-sampled `reset_ball_vel` are used.
        Returns the int32 (actor_indices, dof_indices) the gym setters take (TILT:876-888)."""
        env_ids = env_ids.to(self.device, torch.int64).contiguous()
        k = env_ids.numel()
        cfg = self.cfg
        actor_out = torch.empty(k * cfg.num_actors, dtype=torch.int32, device=self.device)
        dof_per = self.st["dof_indices"].numel() // self.num_envs if "dof_indices" in self.st else 0
        dof_out = torch.empty(k * max(dof_per, 1), dtype=torch.int32, device=self.device)
        bv = ball_vel.to(self.device, torch.float32).contiguous() if ball_vel is not None else None
        byz = ball_pos_yz.to(self.device, torch.float32).contiguous() if ball_pos_yz is not None else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        N.check(self._lib.ppk_reset_idx(self._task, self.buffers(), env_ids.data_ptr(), k, ptr(bv), ptr(byz),
                                        ptr(self.st.get("actor_indices")), ptr(self.st.get("dof_indices")) if dof_per else None,
                                        dof_per, actor_out.data_ptr(), dof_out.data_ptr() if dof_per else None,
                                        self._stream()), "ppk_reset_idx")
        return actor_out, dof_out

    def sample_ball_launch(self, seed: int, epoch: int, env_offset: int = 0, refresh_consumed_only: bool = False):
        """Device-side `generate_random_speed_for_ball` (TILT:307-318) for the whole shard: refills the
        per-env launch table the predicated reset consumes (Philox4x32-10, counter = global env id, epoch)."""
        N.check(self._lib.ppk_sample_ball_launch(self._task, self.buffers(), seed, epoch, env_offset,
                                                 1 if refresh_consumed_only else 0, self._stream()),
                "ppk_sample_ball_launch")

    def reset_indices(self):
        """(actor_indices, dof_indices) int32 of the envs the last fused step reset -- one host read of
        the counter (the gym setters need the count on the host anyway, TILT:883,888)."""
        k = int(self.reset_count.item())
        dof_per = self.st["dof_indices"].numel() // max(self.num_envs, 1)
        return self.reset_actor_indices[:k * self.cfg.num_actors], self.reset_dof_indices[:k * dof_per]

    def _log_now(self) -> bool:
        le = self.cfg.log_every
        return self.log_stats and le > 0 and self.num_steps % le == 0

    def post_physics_step(self):
        """TILT:1022-1052.  Fused: one kernel does progress+=1, reward, reset mask, flag updates,
        the per-env reset and the observations.  With `fused=False` the reference's own sequence of
        calls is issued (three launches and a host-visible `nonzero`)."""
        log = self._log_now()
        if self.envelope and not self._count_zeroed:
            self.reset_count.zero_()
        self._count_zeroed = False
        if self.fused:
            self._step((N.PHASE_ALL if log else (N.PHASE_ALL & ~N.PHASE_STATS)) | (N.PHASE_MOMENTS if self.fused_moments else 0))
        elif self.cfg.variant == "base":
            self._step(N.PHASE_PROGRESS | N.PHASE_RESET)
            self.compute_observations()
            self.compute_reward()
        else:
            self._step(N.PHASE_PROGRESS)
            self.compute_reward(self.actions)
            env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
            if len(env_ids) > 0:
                self.reset_idx(env_ids)
                for name in self.cfg.counter_names:          # ADOF:1171-1175
                    getattr(self, name).fill_(0)
            self.compute_observations()
        if log:
            self.stats.reduce(self._lib, self._stream())
        if self.launch_seed is not None and self.cfg.variant != "base":
            # redraw the launch rows of the envs that just reset (reset_buf still holds this step's mask, TILT:900)
            self.sample_ball_launch(self.launch_seed, epoch=self.num_steps + 1, env_offset=self.env_offset,
                                    refresh_consumed_only=True)
        self.num_steps += 1


class HumanoidPingpongBase(PingpongTask):          # tasks/humanoid_pingpong.py:66
    variant = "base"


class HumanoidPingpong(PingpongTask):              # tasks/humanoid_interos_edit_pingpong_only_3_actor.py:58
    variant = "a3"


class HumanoidPingpongTilt(PingpongTask):          # tasks/humanoid_pingpong_3_actor_tilt.py:58
    variant = "tilt"


class HumanoidPingpongTiltNoEarlyStop(PingpongTask):   # tasks/humanoid_pingpong_3_actor_tilt_no_earlystop.py:58
    variant = "nes"


class Humanoid12PingpongTilt(PingpongTask):        # tasks/humanoid_pingpong_4_actor_tilt.py:58
    variant = "a4"


class HumanoidPingpongAlignment(PingpongTask):     # tasks/humanoid_pingpong_alignment.py:58 (class HumanoidPingpongTilt there)
    variant = "align"


class HumanoidPingpongAlignmentTwoHumanoid(PingpongTask):   # tasks/humanoid_pingpong_alignment.py:1233 (reward definition #2)
    variant = "align2"


class HumanoidPingpongTiltNESSparse27DOF(PingpongTask):   # tasks/humanoid_pingpong_3_actor_all_dof.py:65
    variant = "adof"


# name -> class, as tasks/__init__.py:92-123 maps task names (plus the three it forgets to register)
isaacgym_task_map = {
    "HumanoidPingpongBase": HumanoidPingpongBase,
    "HumanoidPingpongG1": HumanoidPingpong,
    "HumanoidPingpongTiltG1": HumanoidPingpongTilt,
    "HumanoidPingpongTiltNoEarlyStopG1": HumanoidPingpongTiltNoEarlyStop,
    "Humanoid12PingpongTiltG1": Humanoid12PingpongTilt,
    "HumanoidPingpongAlignmentG1": HumanoidPingpongAlignment,
    "HumanoidPingpongTiltNESSparse27DOFG1": HumanoidPingpongTiltNESSparse27DOF,
    "HumanoidPingpongAlignmentTwoHumanoidG1": HumanoidPingpongAlignmentTwoHumanoid,
}
VARIANT_CLASS = {c.variant: c for c in isaacgym_task_map.values()}


def make_task(variant: str, sim_state, **kw) -> PingpongTask:
    return VARIANT_CLASS[variant](sim_state, **kw)
