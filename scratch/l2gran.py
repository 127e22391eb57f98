import ctypes as C, glob, sys, torch
sys.path.insert(0,'.')
rt = C.CDLL(sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[-1])
torch.cuda.init(); torch.zeros(1, device="cuda")
val = C.c_size_t()
rt.cudaDeviceGetLimit(C.byref(val), 5); print("default L2 fetch granularity:", val.value)
import bench
def run(tag):
    for w, n, sets in (("tilt", 65536, 8), ("tilt", 1048576, 2), ("adof", 32768, 8)):
        cfg, tasks = bench.make_tasks(w, n, sets, torch.device("cuda"), 5)
        sec, _ = bench.time_steps(tasks, 1000, 10, False, 0, 1, None)
        print(tag, w, n, round(sec/1000*1e6, 2), "us")
        del tasks; torch.cuda.empty_cache()
run("gran=%d" % val.value)
for g in (32, 128):
    rc = rt.cudaDeviceSetLimit(5, C.c_size_t(g)); rt.cudaDeviceGetLimit(C.byref(val), 5)
    print("set", g, "rc", rc, "now", val.value)
    run("gran=%d" % val.value)
