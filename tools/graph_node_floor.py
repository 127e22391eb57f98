"""What does one dependent kernel node cost in CUDA-graph replay on this box, with nothing to do?
The floor under every per-step time bench.py reports."""
import torch

x = torch.zeros(32, device="cuda")
for _ in range(3):
    x.add_(1.0)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(512):
        x.add_(1.0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    g.replay()
a.record()
for _ in range(20):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"empty dependent kernel node in graph replay: {a.elapsed_time(b) * 1e3 / (20 * 512):.2f} us")
