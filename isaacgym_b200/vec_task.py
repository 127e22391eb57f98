"""Minimal stand-in for `isaacgymenvs.tasks.base.vec_task.VecTask` (not in the reference tree; the
task classes subclass it, TILT:40,118).  It keeps what the hot path touches -- the buffer names,
dtypes and the `step()` call order -- and nothing else (no viewer, no randomisation, no hydra).

step(actions) (upstream VecTask.step):
    clamp actions to +-clipActions -> pre_physics_step(actions) -> simulate x control_freq_inv
    -> post_physics_step() -> clamp obs to +-clipObservations -> (obs, rew_buf, reset_buf, extras)
Both clamps happen inside the CUDA kernels (`PpkBuffers.clip_actions`, `PpkBuffers.clip_observations`): no ATen
kernel runs between `pre_physics_step` and the returned observations.  `clip_observations` defaults to upstream's
`inf` (no clamp; none of the reference's YAMLs sets clipObservations).
PhysX is replaced by a `physics` callable that refreshes the synthetic state tensors in place.
"""
from typing import Callable, Dict, Optional

import torch

from .config import TaskConfig


class VecTask:
    def __init__(self, cfg: TaskConfig, num_envs: int, device: str = "cuda:0", clip_actions: float = 1.0,
                 clip_observations: float = float("inf"), control_freq_inv: int = 1,
                 physics: Optional[Callable[["VecTask"], None]] = None):
        self.cfg = cfg
        self.num_envs = num_envs
        self.device = torch.device(device)
        self.num_obs = cfg.num_obs
        self.num_actions = cfg.num_dofs
        self.clip_actions = clip_actions           # cfg/task/HumanoidPingpongTiltG1.yaml:26
        self.clip_obs = clip_observations
        self.control_freq_inv = control_freq_inv
        self.max_episode_length = cfg.max_episode_length
        self.physics = physics
        self.extras: Dict[str, torch.Tensor] = {}
        self.num_steps = 0
        self.allocate_buffers()

    def allocate_buffers(self):
        """upstream allocate_buffers: obs/rew fp32, reset/progress/timeout/randomize int64."""
        n, dev, cfg = self.num_envs, self.device, self.cfg
        obs_shape = (n, cfg.obs_rows, cfg.num_obs) if cfg.obs_rows > 1 else (n, cfg.num_obs)
        self.obs_buf = torch.zeros(obs_shape, device=dev, dtype=torch.float32)
        self.rew_buf = torch.zeros((n, cfg.obs_rows) if cfg.obs_rows > 1 else (n,), device=dev, dtype=torch.float32)
        self.reset_buf = torch.ones(n, device=dev, dtype=torch.int64)
        self.timeout_buf = torch.zeros(n, device=dev, dtype=torch.int64)
        self.progress_buf = torch.zeros(n, device=dev, dtype=torch.int64)
        self.randomize_buf = torch.zeros(n, device=dev, dtype=torch.int64)
        self.reset_buf_force = torch.zeros(n, device=dev, dtype=torch.int64)

    # -- hooks the task implements -------------------------------------------------------------
    def pre_physics_step(self, actions):
        raise NotImplementedError

    def post_physics_step(self):
        raise NotImplementedError

    def step(self, actions: torch.Tensor):
        # upstream clamps here (torch.clamp(actions, -clip, clip)); the pre-step kernel does it in place
        self.pre_physics_step(actions)
        for _ in range(self.control_freq_inv):
            if self.physics is not None:
                self.physics(self)
        self.post_physics_step()
        # with the envelope enabled timeout_buf is written by the fused step (progress >= L-1 before
        # the reset cleared it); otherwise it stays zero as in the reference, whose reset has already
        # cleared progress_buf by the time upstream VecTask.step evaluates it
        self.extras["time_outs"] = self.timeout_buf
        # obs_buf is already clamped to +-clip_observations where the step kernel produced it
        return {"obs": self.obs_buf}, self.rew_buf, self.reset_buf, self.extras
