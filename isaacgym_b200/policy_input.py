"""Host-side mirror of what consumes `obs_buf` on the learner side.

`RunningMeanStd` keeps the interface of rl_games' module of the same name
(rl_games/algos_torch/running_mean_std.py, selected by `normalize_input: True`,
cfg/train/HumanoidPingpongTiltG1PPO.yaml:51): buffers `running_mean`, `running_var` (float64 [insize]),
`count` (float64 scalar), `epsilon`, `train()/eval()`, and `forward(input)` which in training mode first
merges the batch moments and then returns the normalised, +-5-clamped input.

`FirstLayer` is the first `nn.Linear(num_obs, units[0])` + activation of the MLP
(`units: [2048, ...]`, `activation: elu`, yaml:29-30) as `torch.autocast(float16)` runs it under
`mixed_precision: True` (yaml:50), fused with the normalisation: one tcgen05 kernel, fp16 output.
`FirstLayer(..., precision="fp32")` is the same layer as the rollout forward runs it (no autocast): fp32 products from
TF32 hi/lo splits on the tensor cores, fp32 output.

CUDA only; every call goes through the C ABI of libppk.so.
"""
import ctypes as C
from typing import Optional

import torch

from . import _native as N


class RunningMeanStd:
    def __init__(self, insize: int, epsilon: float = 1e-05, clip_obs: float = 0.0, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RunningMeanStd is CUDA-only (no CPU fallback)")
        self.insize = int(insize)
        self.epsilon = float(epsilon)
        self.clip_obs = float(clip_obs)           # VecTask.step's clamp of obs_buf; 0 = none
        self.running_mean = torch.zeros(self.insize, dtype=torch.float64, device=self.device)
        self.running_var = torch.ones(self.insize, dtype=torch.float64, device=self.device)
        self.count = torch.ones((), dtype=torch.float64, device=self.device)
        self.training = True
        self._lib = N.load()
        self._moments = torch.zeros(self._lib.ppk_rms_scratch_doubles(self.insize), dtype=torch.float64, device=self.device)
        self._s = self._make_struct()           # the buffers never move: build the descriptor once

    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def _struct(self) -> N.PpkRunningMeanStd:
        s = self._s                              # refreshed each call: callers may rebind the buffers
        s.clip_obs, s.epsilon = self.clip_obs, self.epsilon
        s.running_mean, s.running_var = self.running_mean.data_ptr(), self.running_var.data_ptr()
        s.count = self.count.data_ptr()
        return s

    def _make_struct(self) -> N.PpkRunningMeanStd:
        s = N.PpkRunningMeanStd()
        s.struct_size = C.sizeof(N.PpkRunningMeanStd)
        s.width, s.epsilon, s.clip_obs = self.insize, self.epsilon, self.clip_obs
        s.running_mean, s.running_var = self.running_mean.data_ptr(), self.running_var.data_ptr()
        s.count, s.moments = self.count.data_ptr(), self._moments.data_ptr()
        return s

    def _check(self, x: torch.Tensor):
        if x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device or x.dim() != 2 \
                or x.shape[1] != self.insize:
            raise ValueError(f"expected a contiguous float32 [rows,{self.insize}] tensor on {self.device}")

    def update(self, x: torch.Tensor, group=None, global_rows: Optional[int] = None):
        """Merge the batch moments.  With a process group the moments are all-reduced first, so every rank ends
        with the statistics of the global batch; `global_rows` (the batch size summed over ranks, known to the
        caller when the shards are equal) avoids all-reducing the row count and reading it back."""
        self._check(x)
        s = self._struct()
        if group is None:
            N.check(self._lib.ppk_rms_update(s, x.data_ptr(), x.shape[0], N.current_stream_ptr()), "rms_update")
            return
        import torch.distributed as dist
        N.check(self._lib.ppk_rms_accumulate(s, x.data_ptr(), x.shape[0], N.current_stream_ptr()), "rms_accumulate")
        dist.all_reduce(self._moments[:2 * self.insize], group=group)
        if global_rows is None:
            rows = torch.tensor([float(x.shape[0])], dtype=torch.float64, device=self.device)
            dist.all_reduce(rows, group=group)
            global_rows = rows.item()              # host read: pass global_rows to stay asynchronous
        N.check(self._lib.ppk_rms_merge(s, float(global_rows), N.current_stream_ptr()), "rms_merge")

    def update_from_step(self, obs_moments: torch.Tensor, rows: int, group=None, global_rows: Optional[int] = None):
        """The same merge from the column moments the fused task step left in `obs_moments`
        ([PPK_MOMENT_SLOTS, 2*insize] float64, `PPK_PHASE_MOMENTS` / `PingpongTask(fused_moments=True)`) -- obs_buf is
        not read again.  The moments are those of obs_buf as stored (clamp it in the step: `clip_observations`).
        `rows`: rows of obs_buf the step wrote (envs, or 2 x envs for the two-humanoid variants)."""
        if obs_moments.dtype != torch.float64 or not obs_moments.is_contiguous() or obs_moments.device != self.device \
                or obs_moments.numel() != N.PPK_MOMENT_SLOTS * 2 * self.insize:
            raise ValueError(f"expected a contiguous float64 [{N.PPK_MOMENT_SLOTS}, {2 * self.insize}] tensor on {self.device}")
        s = self._struct()
        if group is None:
            N.check(self._lib.ppk_rms_fold_step_moments(s, obs_moments.data_ptr(), float(rows), 1, N.current_stream_ptr()),
                    "rms_fold_step_moments")
            return
        import torch.distributed as dist
        N.check(self._lib.ppk_rms_fold_step_moments(s, obs_moments.data_ptr(), float(rows), 0, N.current_stream_ptr()),
                "rms_fold_step_moments")
        dist.all_reduce(self._moments[:2 * self.insize], group=group)
        if global_rows is None:
            r = torch.tensor([float(rows)], dtype=torch.float64, device=self.device)
            dist.all_reduce(r, group=group)
            global_rows = r.item()
        N.check(self._lib.ppk_rms_merge(s, float(global_rows), N.current_stream_ptr()), "rms_merge")

    def normalize(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check(x)
        if out is None:
            out = torch.empty_like(x)
        N.check(self._lib.ppk_rms_normalize(self._struct(), x.data_ptr(), x.shape[0], out.data_ptr(),
                                            N.current_stream_ptr()), "rms_normalize")
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            self.update(x)
        return self.normalize(x)

    __call__ = forward


class FirstLayer:
    ACTIVATIONS = {"None": N.PPK_ACT_NONE, None: N.PPK_ACT_NONE, "elu": N.PPK_ACT_ELU}

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], activation="elu",
                 running_mean_std: Optional[RunningMeanStd] = None, precision: str = "autocast_fp16"):
        if weight.device.type != "cuda":
            raise RuntimeError("FirstLayer is CUDA-only (no CPU fallback)")
        if precision not in ("autocast_fp16", "fp32"):
            raise ValueError("precision is 'autocast_fp16' (learner forward) or 'fp32' (rollout forward)")
        self._lib = N.load()
        self.units, self.width = int(weight.shape[0]), int(weight.shape[1])
        self.activation = self.ACTIVATIONS[activation]
        self.rms = running_mean_std
        self.fp32 = precision == "fp32"
        lib = self._lib
        self._packed_bytes, self._pack, self._forward = (
            (lib.ppk_linear_packed_bytes_f32, lib.ppk_linear_pack_f32, lib.ppk_policy_first_layer_f32) if self.fp32 else
            (lib.ppk_linear_packed_bytes, lib.ppk_linear_pack, lib.ppk_policy_first_layer))
        self.out_dtype = torch.float32 if self.fp32 else torch.float16
        nbytes = self._packed_bytes(self.units, self.width)
        if nbytes == 0:
            raise ValueError("units must be a positive multiple of 256")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
        w = weight.detach().to(torch.float32).contiguous()
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        N.check(self._pack(w.data_ptr(), None if b is None else b.data_ptr(), self.units, self.width,
                           self.packed.data_ptr(), nbytes, N.current_stream_ptr()), "linear_pack")

    def forward(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if obs.dtype != torch.float32 or not obs.is_contiguous() or obs.dim() != 2 or obs.shape[1] != self.width:
            raise ValueError(f"expected a contiguous float32 [rows,{self.width}] tensor")
        if out is None:
            out = torch.empty(obs.shape[0], self.units, dtype=self.out_dtype, device=obs.device)
        elif out.dtype != self.out_dtype or not out.is_contiguous() or tuple(out.shape) != (obs.shape[0], self.units):
            raise ValueError(f"out must be a contiguous {self.out_dtype} [rows,{self.units}] tensor")
        rms = None if self.rms is None else self.rms._struct()
        N.check(self._forward(rms, obs.data_ptr(), obs.shape[0], self.width, self.packed.data_ptr(),
                              self.units, self.activation, out.data_ptr(), N.current_stream_ptr()),
                "policy_first_layer")
        return out

    __call__ = forward
