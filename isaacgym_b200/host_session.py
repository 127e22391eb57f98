"""Host-buffer form of the task step (isaacgym CPU pipeline: state tensors live in HOST memory).

Wraps `ppk_host_session_*`: each call moves the rows the step consumes to the device, runs the fused
step there and brings obs / reward / reset / progress / flags (and reset rows) back -- H2D and D2H are
inside the call.  Pinned tensors give full PCIe speed; the per-step pipeline is recorded into a CUDA
graph by the library after the first call with the same buffers."""
import ctypes as C
from typing import Dict

import torch

from . import _native as N
from .config import TaskConfig


class HostSession:
    def __init__(self, cfg: TaskConfig, state: Dict[str, torch.Tensor], num_chunks: int = 4, pin: bool = True):
        self.cfg = cfg
        self._lib = N.load()
        st = dict(state)
        pre = st.get("pre_ball_states")
        if pre is not None and pre.shape[-1] == 13:
            st["pre_ball_states"] = pre[:, [7, 9]].contiguous()
        if cfg.variant == "base":
            st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
        self.state = {}
        for k, v in st.items():
            v = v.contiguous()
            if v.device.type != "cpu":
                v = v.cpu()
            self.state[k] = v.pin_memory() if (pin and v.numel() > 0 and torch.cuda.is_available()) else v
        self.num_envs = self.state["progress_buf"].shape[0]
        # with PPK_PHASE_STATS the session folds its device slots and returns the sums of the call here (PpkStat order)
        stats = torch.zeros(N.PPK_NUM_STATS, dtype=torch.float64)
        self.state["stats"] = stats.pin_memory() if (pin and torch.cuda.is_available()) else stats
        self._buf = N.make_buffers(cfg, self.state, host=True)
        self._sess = C.c_void_p()
        N.check(self._lib.ppk_host_session_create(N.make_task(cfg), self.num_envs, num_chunks, C.byref(self._sess)),
                "ppk_host_session_create")

    def post_physics_step(self, phases: int = N.PHASE_ALL & ~N.PHASE_STATS):
        N.check(self._lib.ppk_host_post_physics_step(self._sess, self._buf, phases), "ppk_host_post_physics_step")

    def stats_means(self):
        """Per-env means of the sums returned by the last call that included PHASE_STATS."""
        return {name: float(x) / self.num_envs for name, x in zip(N.STAT_NAMES, self.state["stats"].tolist())}

    def traffic(self):
        h2d, d2h = C.c_int64(), C.c_int64()
        self._lib.ppk_host_session_traffic(self._sess, C.byref(h2d), C.byref(d2h))
        return int(h2d.value), int(d2h.value)

    def close(self):
        if self._sess:
            self._lib.ppk_host_session_destroy(self._sess)
            self._sess = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
