"""Pieces of the host-buffer (e2e) step, timed one by one on a B200 box, to decide how the PCIe link is best shared:
DMA of the strided rigid-body rows, DMA of the linear tensors, the step kernel with its linear inputs left in pinned
HOST memory (the TMA engine pulls them over PCIe itself), the D2H of the observations.  Not part of the product path."""
import ctypes as C
import glob
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import make_state

rt = C.CDLL(sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[-1])
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = CONFIGS["tilt"]
lib = N.load()
st = make_state(cfg, n, seed=3, adversarial=False)
host = {k: v.contiguous().pin_memory() for k, v in st.items()}
host["pre_ball_states"] = st["pre_ball_states"][:, [7, 9]].contiguous().pin_memory()
dev = {k: v.cuda() for k, v in host.items()}
BD = 12
rb_c = torch.zeros(n, BD, 13, device="cuda")
full = torch.zeros(n, 42, 13, device="cuda")
dev["stats"] = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device="cuda")
dev["scratch"] = torch.zeros(16, dtype=torch.int32, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def timeit(f, reps=10):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def rb_dma(stream, lo=0, m=n):
    src = host["rigid_body_states"].data_ptr() + lo * 42 * 52
    dst = rb_c.data_ptr() + lo * BD * 52
    rt.cudaMemcpy2DAsync(dst, BD * 52, src, 42 * 52, 52, m, 1, stream.cuda_stream)                     # row 0
    rt.cudaMemcpy2DAsync(dst + 52, BD * 52, src + 31 * 52, 42 * 52, 9 * 52, m, 1, stream.cuda_stream)  # rows 31..39


def lin_dma(stream):
    with torch.cuda.stream(stream):
        for k in ("root_states", "dof_states", "dof_forces"):
            dev[k].copy_(host[k], non_blocking=True)


# the compact device task: ids remapped onto the 12 staged rows
import dataclasses
cfg_c = cfg.with_(num_bodies=BD, body_ids=tuple(range(10)), paddle_body=(9, 9))
task = N.make_task(cfg_c)


def buffers(zero_copy_linear):
    d = dict(dev)
    d["rigid_body_states"] = rb_c
    b = N.make_buffers(cfg_c, d)
    if zero_copy_linear:
        b.root_states = host["root_states"].data_ptr()
        b.dof_states = host["dof_states"].data_ptr()
        b.dof_forces = host["dof_forces"].data_ptr()
        b.root_states_out = dev["root_states"].data_ptr()      # resets still land on the device copy
        b.dof_states_out = dev["dof_states"].data_ptr()
    return b


ph = N.PHASE_ALL & ~N.PHASE_STATS
b_dev, b_zc = buffers(False), buffers(True)


def kern(b, stream):
    N.check(lib.ppk_post_physics_step(task, b, ph, stream.cuda_stream), "step")


def obs_d2h(stream):
    with torch.cuda.stream(stream):
        host["obs_buf"].copy_(dev["obs_buf"], non_blocking=True)


rb_dma(s1); lin_dma(s1); torch.cuda.synchronize()
t_rb = timeit(lambda: rb_dma(s1))
t_lin = timeit(lambda: lin_dma(s1))
t_k = timeit(lambda: kern(b_dev, s1))
t_kz = timeit(lambda: kern(b_zc, s1))
t_obs = timeit(lambda: obs_d2h(s1))
rb_bytes, lin_bytes, obs_bytes = n * 520, n * (156 + 56 + 28), n * 320
print(f"n={n}")
print(f"rb rows DMA (52 B + 468 B pieces): {t_rb*1e3:.3f} ms  {rb_bytes/t_rb/1e9:.1f} GB/s")
print(f"linear root/dof/force DMA        : {t_lin*1e3:.3f} ms  {lin_bytes/t_lin/1e9:.1f} GB/s")
print(f"step kernel, all inputs on device: {t_k*1e6:.1f} us")
print(f"step kernel, linear inputs pulled from pinned host by TMA: {t_kz*1e6:.1f} us  {lin_bytes/t_kz/1e9:.1f} GB/s")
print(f"obs D2H                          : {t_obs*1e3:.3f} ms  {obs_bytes/t_obs/1e9:.1f} GB/s")


def serial():
    rb_dma(s1); lin_dma(s1); kern(b_dev, s1); obs_d2h(s1)


def hybrid():          # rb DMA, then the kernel pulls the rest itself
    rb_dma(s1); kern(b_zc, s1); obs_d2h(s1)


def overlap(zc, chunks):
    per = n // chunks
    streams = (s1, s2, s3)
    # chunk views: the same buffers offset by env ranges would need per-chunk PpkBuffers; emulate with full-size calls / chunks
    for c in range(chunks):
        s = streams[c % 3]
        rb_dma(s, c * per, per)
    # (a full pipeline is what ppk_host.cu does; here only the shares are of interest)


print(f"serial  DMA all + kernel + D2H   : {timeit(serial)*1e3:.3f} ms")
print(f"hybrid  DMA rb + zc kernel + D2H : {timeit(hybrid)*1e3:.3f} ms")


def concurrent_rb_and_zc():
    rb_dma(s1)
    kern(b_zc, s2)          # reads the previous rb_c contents: only the link sharing is of interest here


print(f"rb DMA || zero-copy kernel       : {timeit(concurrent_rb_and_zc)*1e3:.3f} ms  (sum of bytes {(rb_bytes+lin_bytes)/1e6:.1f} MB)")


def duplex():
    rb_dma(s1); lin_dma(s1); obs_d2h(s2)


print(f"H2D all || obs D2H               : {timeit(duplex)*1e3:.3f} ms")


# ---- do the three H2D pieces overlap when they sit on different streams (different copy engines)?
def split_rb(sa, sb, lo=0, m=n):
    src = host["rigid_body_states"].data_ptr() + lo * 42 * 52
    dst = rb_c.data_ptr() + lo * BD * 52
    rt.cudaMemcpy2DAsync(dst + 52, BD * 52, src + 31 * 52, 42 * 52, 9 * 52, m, 1, sa.cuda_stream)
    rt.cudaMemcpy2DAsync(dst, BD * 52, src, 42 * 52, 52, m, 1, sb.cuda_stream)


print(f"spans alone                      : {timeit(lambda: rt.cudaMemcpy2DAsync(rb_c.data_ptr() + 52, BD * 52, host['rigid_body_states'].data_ptr() + 31 * 52, 42 * 52, 9 * 52, n, 1, s1.cuda_stream))*1e3:.3f} ms")
print(f"row 0 alone                      : {timeit(lambda: rt.cudaMemcpy2DAsync(rb_c.data_ptr(), BD * 52, host['rigid_body_states'].data_ptr(), 42 * 52, 52, n, 1, s1.cuda_stream))*1e3:.3f} ms")
print(f"spans (s1) || row 0 (s2)         : {timeit(lambda: split_rb(s1, s2))*1e3:.3f} ms")


def three_way():
    split_rb(s1, s2)
    lin_dma(s3)


print(f"spans (s1) || row 0 (s2) || linear (s3): {timeit(three_way)*1e3:.3f} ms")


def three_way_duplex():
    split_rb(s1, s2)
    lin_dma(s3)
    obs_d2h(s2)


print(f"... + obs D2H behind row 0 on s2 : {timeit(three_way_duplex)*1e3:.3f} ms")


def wide():        # one 2-D copy of 624-byte pieces: rows 31..39 of env e, the two non-humanoid rows, row 0 of env e+1
    rt.cudaMemcpy2DAsync(full_c.data_ptr(), 12 * 52, host["rigid_body_states"].data_ptr() + 31 * 52, 42 * 52, 12 * 52, n - 1, 1, s1.cuda_stream)


full_c = torch.zeros(n, 12, 13, device="cuda")
print(f"one 624-byte-piece copy          : {timeit(wide)*1e3:.3f} ms")
for w in (1, 2, 4, 9, 12, 21, 42):
    tt = timeit(lambda: rt.cudaMemcpy2DAsync(full.data_ptr(), 42 * 52, host["rigid_body_states"].data_ptr(), 42 * 52, w * 52, n, 1, s1.cuda_stream))
    print(f"2-D copy, {w*52:5d}-byte pieces      : {tt*1e3:.3f} ms  {w*52*n/tt/1e9:.1f} GB/s")
