"""Per-unit timeline of one launch of first_layer_f32_kernel (needs a -DPPK_TRACE build):
    PPK_NVCC_EXTRA=-DPPK_TRACE python -m isaacgym_b200.build --out scratch/libs/libppk_trace.so
    PPK_LIB=scratch/libs/libppk_trace.so python tools/f32_timeline.py [rows] [width] [units]
Stamps (globaltimer ns, see f32_stamp in csrc/ppk_policy_f32.cuh): MMA thread, epilogue warp 0, weight producer, row-tile
preparation, per unit of the CTA's sequence."""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200 import _native as N
from isaacgym_b200.policy_input import FirstLayer
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
width = int(sys.argv[2]) if len(sys.argv) > 2 else 80
units = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
lib = N.load()
g = torch.Generator().manual_seed(1)
x = torch.randn(rows, width, generator=g).cuda(); w = (torch.randn(units, width, generator=g) / 9).cuda(); b = torch.randn(units, generator=g).cuda()
layer = FirstLayer(w, b, "elu", None, precision="fp32")
out = torch.empty(rows, units, device="cuda")
for _ in range(3): layer(x, out)
blocks = 148
buf = torch.zeros(blocks * 4 * 32 * 4, dtype=torch.int64, device="cuda")
lib.ppk_debug_set_trace.argtypes = [C.c_void_p]
assert lib.ppk_debug_set_trace(buf.data_ptr()) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer(x, out); e1.record(); torch.cuda.synchronize()
print("event time us", round(e0.elapsed_time(e1) * 1e3, 1))
tr = buf.cpu().numpy().reshape(blocks, 4, 32, 4).astype(np.float64)
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, (tr - t0) / 1e3, np.nan)
for blk in (0, 1, 76):
    print(f"--- CTA {blk}: unit | mma: waits done, issued | epi w0: acc full, tmem loaded, store0, store1 | producer first stage | prep: start, ready")
    for u in range(14):
        m, e, p, r = rel[blk, 0, u], rel[blk, 1, u], rel[blk, 2, u], rel[blk, 3, u]
        f = lambda v: "   --  " if np.isnan(v) else f"{v:7.2f}"
        print(f"{u:3d} | {f(m[0])} {f(m[1])} | {f(e[0])} {f(e[1])} {f(e[2])} {f(e[3])} | {f(p[0])} | {f(r[0])} {f(r[2])}")
for blk in (0, 76):
    print(f"--- CTA {blk}, second row tile, per K step: loads issued, released, stored, arrived")
    for kb in range(12):
        r = rel[blk, 3, 16 + kb]
        print(f"{kb:3d} | " + " ".join("   --  " if np.isnan(v) else f"{v:7.2f}" for v in r))
m = rel[:, 0, :, :]
per_unit = np.nanmedian(np.diff(m[:, 2:26, 0], axis=1))
print("median time between consecutive units' MMA starts (us):", round(float(per_unit), 2))
print("median MMA issue span per unit (us):", round(float(np.nanmedian(m[:, 2:26, 1] - m[:, 2:26, 0])), 2))
e = rel[:, 1, :, :]
print("median epilogue: acc_full->loaded, loaded->store0, store0->store1 (us):",
      [round(float(np.nanmedian(e[:, 2:26, i + 1] - e[:, 2:26, i])), 2) for i in range(3)])
print("median acc_full(u) - mma issued(u) (us):", round(float(np.nanmedian(e[:, 2:26, 0] - m[:, 2:26, 1])), 2))
