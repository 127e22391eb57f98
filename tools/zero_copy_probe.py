"""How fast is the fused step when EVERY tensor stays in pinned host memory and the kernel reads / writes it
over PCIe itself (no DMA staging)?  Compared with the host session (DMA pipeline) on the same data.
usage (GPU box): python tools/zero_copy_probe.py [variant] [envs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200 import _native as N  # noqa: E402
from isaacgym_b200.config import CONFIGS  # noqa: E402
from isaacgym_b200.host_session import HostSession  # noqa: E402
from isaacgym_b200.synth import make_state  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "tilt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
cfg = CONFIGS[variant]
lib = N.load()
st = make_state(cfg, n, seed=5, device="cpu", adversarial=False)
st["pre_ball_states"] = st["pre_ball_states"][:, [7, 9]].contiguous() if st["pre_ball_states"].shape[-1] == 13 else st["pre_ball_states"]
pin = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
pin["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device="cuda")
pin["scratch"] = torch.zeros(16, dtype=torch.int32, device="cuda")
task = N.make_task(cfg)
buf = N.make_buffers(cfg, {k: v for k, v in pin.items() if k not in ("stats", "scratch")}, host=True)
buf.stats, buf.scratch = pin["stats"].data_ptr(), pin["scratch"].data_ptr()
phases = N.PHASE_ALL & ~N.PHASE_STATS
stream = N.current_stream_ptr()
for _ in range(3):
    N.check(lib.ppk_post_physics_step(task, buf, phases, stream))
torch.cuda.synchronize()
t0 = time.perf_counter()
iters = 20
for _ in range(iters):
    N.check(lib.ppk_post_physics_step(task, buf, phases, stream))
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
print(f"zero-copy kernel ({variant}, {n} envs): {dt * 1e3:.3f} ms/step -> {n / dt / 1e6:.1f} M env-steps/s")

sess = HostSession(cfg, st, num_chunks=4, pin=True)
for _ in range(3):
    sess.post_physics_step(phases)
t0 = time.perf_counter()
for _ in range(iters):
    sess.post_physics_step(phases)
dt2 = (time.perf_counter() - t0) / iters
print(f"host session (DMA pipeline):          {dt2 * 1e3:.3f} ms/step -> {n / dt2 / 1e6:.1f} M env-steps/s")
sess.close()
