"""What launches when the reference-facing call is made: `task.step(actions)` of the host mirror (VecTask.step envelope:
action clamp, pre_physics_step, post_physics_step with the observation clamp, time-outs and the compacted reset lists).
Run it under `ncu --metrics gpu__time_duration.sum --clock-control none --csv` WITHOUT a kernel filter: every kernel of
the steps after the warm-up must be one of libppk.so's (no ATen kernel between pre_physics_step and the returned obs).
    python tools/vectask_step_launches.py [variant] [envs] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200.config import CONFIGS           # noqa: E402
from isaacgym_b200.synth import make_state         # noqa: E402
from isaacgym_b200.tasks import make_task          # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "tilt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = CONFIGS[variant]
st = make_state(cfg, n, seed=7, device="cuda:0", adversarial=False)
task = make_task(variant, st, device="cuda:0", clip_actions=1.0, clip_observations=5.0, envelope=True)
actions = torch.rand(n, cfg.num_dofs, device="cuda:0") * 2 - 1
torch.cuda.synchronize()
for i in range(steps):
    torch.cuda.nvtx.range_push(f"task.step {i}")
    obs, rew, reset, extras = task.step(actions)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("steps", steps, "obs", tuple(obs["obs"].shape), "resets last step", int(task.reset_count.item()) if hasattr(task, "reset_count") else "n/a")
