"""Per-CTA timeline of one launch of the family step kernel (needs a -DPPK_TRACE build):
    PPK_NVCC_EXTRA=-DPPK_TRACE python -m isaacgym_b200.build --out scratch/libs/libppk_trace.so
    PPK_LIB=scratch/libs/libppk_trace.so python tools/trace_timeline.py [envs] [variant]
Stamps per warp (globaltimer ns): 0 CTA start, 1 copies issued, 2 data arrived, 3 warp done."""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import make_state
from isaacgym_b200.tasks import make_task
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
variant = sys.argv[2] if len(sys.argv) > 2 else "tilt"
cfg = CONFIGS[variant]
lib = N.load()
tasks = [make_task(variant, make_state(cfg, n, seed=s, device="cuda", adversarial=False), device="cuda") for s in range(4)]
tile = {"a4": 16, "align2": 16, "adof": 8}.get(variant, 16 if n <= 71040 else 32)
blocks = (n + tile - 1) // tile
buf = torch.zeros(blocks * 8 * 4, dtype=torch.int64, device="cuda")
lib.ppk_debug_set_trace.argtypes = [C.c_void_p]
ph = N.PHASE_ALL & ~N.PHASE_STATS
for t in tasks: t._step(ph)
torch.cuda.synchronize()
assert lib.ppk_debug_set_trace(buf.data_ptr()) == 0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tasks[0]._step(ph)          # a preceding kernel so the traced launch has a predecessor (as in the graph)
buf.zero_(); torch.cuda.synchronize()
ev0.record(); tasks[1]._step(ph); ev1.record(); torch.cuda.synchronize()
print("event time us", ev0.elapsed_time(ev1) * 1e3)
nw = 6 if variant == "adof" else 4
tr = buf.cpu().numpy().reshape(blocks, 8, 4).astype(np.int64)[:, :nw, :]
base = tr[:, 0, 0].min()
rel = lambda x: (x - base) / 1e3
start = rel(tr[:, 0, 0]); issued = rel(tr[:, 0, 1]); arrived = rel(tr[:, :, 2]); done = rel(tr[:, :, 3])
arrive_w = 0 if variant == "adof" else 1
q = lambda a: np.percentile(a, [0, 10, 50, 90, 100]).round(2)
print("CTA start (us since first)        p0/10/50/90/100:", q(start))
print("copies issued - start (warp 0)                   :", q(issued - start))
print("data arrived - issued                            :", q(arrived[:, arrive_w] - issued))
names = ("warp0 frames+reward", "warp1 balance", "warp2 balance", "warp3 pp bodies", "warp4 dofs", "warp5 reset+ball") if variant == "adof" \
    else ("warp0 frames+rot", "warp1 reward+tail", "warp2 rot", "warp3 rot")
for w, name in enumerate(names):
    print(f"{name:18s} done - arrived              :", q(done[:, w] - arrived[:, w]))
print("CTA lifetime (last warp done - start)            :", q(done.max(axis=1) - start))
print("kernel span (last done - first start) us         :", done.max() - start.min())
ends = done.max(axis=1)
arr = arrived[:, arrive_w]
edges = np.arange(0, ends.max() + 0.5, 0.5)
for i in range(len(edges) - 1):
    a, b = edges[i], edges[i + 1]
    print(f"{a:5.1f}-{b:5.1f} us  started {((start >= a) & (start < b)).sum():5d}  arrived {((arr >= a) & (arr < b)).sum():5d}"
          f"  finished {((ends >= a) & (ends < b)).sum():5d}  alive {((start <= a) & (ends > a)).sum():5d}")
