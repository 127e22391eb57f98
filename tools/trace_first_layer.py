"""Timeline of first_layer_kernel from globaltimer stamps (debug build with -DPPK_TRACE, loaded via PPK_LIB).
usage: PPK_LIB=<trace build> python tools/trace_first_layer.py [rows] [units]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200 import _native as N  # noqa: E402
from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
units = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
lib = N.load()
lib.ppk_debug_set_trace.argtypes = [C.c_void_p]
dev = "cuda:0"
x = [torch.randn(rows, 80, device=dev) for _ in range(3)]
rms = RunningMeanStd(80, device=dev).eval()
layer = FirstLayer(torch.randn(units, 80, device=dev) * 0.1, torch.zeros(units, device=dev), "elu", rms)
out = torch.empty(rows, units, dtype=torch.float16, device=dev)
for i in range(3):
    layer(x[i], out)
torch.cuda.synchronize()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
assert lib.ppk_debug_set_trace(buf.data_ptr()) == 0
layer(x[0], out)
buf.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer(x[1], out); e1.record(); torch.cuda.synchronize()
print("event us", e0.elapsed_time(e1) * 1e3)
tr = buf.cpu().numpy().reshape(148, 16).astype(np.int64)
base = tr[:, 0].min()
names = {0: "setup done", 1: "prep tile0 done", 2: "prep tile1 done", 3: "mma unit0 issued", 4: "epi sees unit0",
         5: "epi unit0 stored", 9: "epi unit8 stored", 10: "epi unit16 stored", 6: "mma last issued", 7: "epi last stored", 8: "kernel end"}
for slot in (0, 1, 2, 3, 4, 5, 9, 10, 6, 7, 8):
    v = (tr[:, slot] - base) / 1e3
    v = v[tr[:, slot] > 0]
    print(f"{names[slot]:20s} p0/50/100 us: {np.percentile(v, [0, 50, 100]).round(2)}")
