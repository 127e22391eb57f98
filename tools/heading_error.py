"""Largest deviation of the rotated observation fields from the CPU oracle for the loaded library (PPK_LIB selects a
build): err / max(1, row scale), the quantity the parity tolerance (rtol 1e-5 + 1e-6 * row scale) is stated on."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import clone_state, make_state
from oracle import task_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for variant in ("tilt", "a4", "adof"):
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=77)
    want = clone_state(st)
    task_oracle.post_physics_step(cfg, want)
    g = clone_state(st, "cuda:0")
    g["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device="cuda:0")
    g["scratch"] = torch.zeros(16, dtype=torch.int32, device="cuda:0")
    lib = N.load()
    N.check(lib.ppk_post_physics_step(N.make_task(cfg), N.make_buffers(cfg, g), N.PHASE_ALL, N.current_stream_ptr()), "step")
    torch.cuda.synchronize()
    a = g["obs_buf"].double().cpu().reshape(-1, cfg.num_obs)
    b = want["obs_buf"].double().reshape(-1, cfg.num_obs)
    J = len(cfg.body_ids)
    scale = b.abs().amax(dim=-1, keepdim=True).clamp_min(1.0)
    err = ((a - b).abs() / scale)[:, :6 * J]
    rel = ((a - b).abs() / b.abs().clamp_min(1e-30))[:, :6 * J]
    big = b.abs()[:, :6 * J] > 0.1
    print(f"{variant}: rotated fields max |err|/row scale {float(err.max()):.3e}  (tolerance 1e-6 + 1e-5 relative), "
          f"mean {float(err.mean()):.3e}, max relative error where |x| > 0.1: {float(rel[big].max()):.3e}")
