"""Timeline of one probe launch (tools/tma_probe.cu, argv[6]): bytes landed / tiles finished per 0.5 us bucket."""
import sys
import numpy as np

path, tile = sys.argv[1], int(sys.argv[2])
a = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)
t0 = a[:, 0].min()
start, arrive, end = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
print(f"tiles {len(a)}  first start 0  last start {start.max():.2f}  last end {end.max():.2f} us")
print(f"load latency (start->arrive): median {np.median(arrive - start):.2f}  p10 {np.percentile(arrive - start, 10):.2f} p90 {np.percentile(arrive - start, 90):.2f}")
print(f"store (arrive->end): median {np.median(end - arrive):.2f} p90 {np.percentile(end - arrive, 90):.2f}")
bw = 0.5
edges = np.arange(0, end.max() + bw, bw)
hs, _ = np.histogram(start, edges)
ha, _ = np.histogram(arrive, edges)
he, _ = np.histogram(end, edges)
per_tile_in = tile * 883e-6   # MB
for i in range(len(edges) - 1):
    print(f"{edges[i]:5.1f}-{edges[i+1]:5.1f} us  started {hs[i]:5d}  arrived {ha[i]:5d} ({ha[i]*per_tile_in/bw:6.2f} TB/s in)  finished {he[i]:5d}")
sm = a[:, 3].astype(int)
per_sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
print(f"per-SM last end: min {per_sm_end.min():.2f} median {np.median(per_sm_end):.2f} max {per_sm_end.max():.2f}")
