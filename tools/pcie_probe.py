import ctypes as C, time, torch
rt = C.CDLL("libcudart.so.12") if False else None
import glob, os
cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(sorted(cands)[-1])
rt.cudaMemcpy2DAsync.argtypes=[C.c_void_p,C.c_size_t,C.c_void_p,C.c_size_t,C.c_size_t,C.c_size_t,C.c_int,C.c_void_p]
n=65536
rb = torch.randn(n,42,13).pin_memory()
dev = torch.empty(n,11,13,device='cuda')
full = torch.empty(n,42,13,device='cuda')
lin = torch.randn(51_000_000//4).pin_memory(); dlin=torch.empty_like(lin,device='cuda')
out = torch.empty(36_000_000//4).pin_memory(); dout=torch.randn(36_000_000//4,device='cuda')
def t(f,reps=10):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
s=torch.cuda.current_stream().cuda_stream
def c2d(width_rows, first):
    rt.cudaMemcpy2DAsync(dev.data_ptr(), 11*52, rb.data_ptr()+first*52, 42*52, width_rows*52, n, 1, s)
print("linear H2D 51MB: %.3f ms -> %.1f GB/s"%((x:=t(lambda: dlin.copy_(lin,non_blocking=True)))*1e3, 51e-3/x))
print("linear D2H 36MB: %.3f ms -> %.1f GB/s"%((x:=t(lambda: out.copy_(dout,non_blocking=True)))*1e3, 36e-3/x))
print("2D rows31-39 (468B x %d): %.3f ms -> %.1f GB/s"%(n,(x:=t(lambda: c2d(9,31)))*1e3, 468*n/1e9/x))
print("2D row0 (52B x %d): %.3f ms -> %.1f GB/s"%(n,(x:=t(lambda: c2d(1,0)))*1e3, 52*n/1e9/x))
print("full rb linear 143MB: %.3f ms -> %.1f GB/s"%((x:=t(lambda: full.copy_(rb,non_blocking=True)))*1e3, rb.numel()*4/1e9/x))
def both():
    dlin.copy_(lin,non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(dout,non_blocking=True)
s2=torch.cuda.Stream()
print("duplex 51MB H2D + 36MB D2H: %.3f ms"%(t(both)*1e3))
# zero-copy kernel read of host memory: gather rows via torch index on mapped memory is not possible; use a copy kernel over a pinned tensor view
import os
print(torch.cuda.get_device_name(), os.cpu_count())
