"""Timing of the learner-side kernels on one GPU, CUDA events, after warm-up.
usage (GPU box): python tools/bench_first_layer.py [rows] [width]
Prints one JSON line per kernel; the torch line (cuBLAS fp16 linear + elementwise kernels under
autocast) is the library baseline the fused kernel is compared with."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd  # noqa: E402

PEAK = 6533.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    sets = [torch.randn(rows, width, device=dev, generator=g) * 2 + 0.5 for _ in range(8)]   # rotated: 8 x 21 MB
    rms = RunningMeanStd(width, device=dev)
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % len(sets)
        return sets[i[0]]

    t = timeit(lambda: rms.update(nxt()))
    print(json.dumps({"kernel": "rms_update (moments + merge)", "rows": rows, "us": t * 1e6,
                      "hbm_frac": rows * width * 4 / t / 1e9 / PEAK}))
    out32 = torch.empty(rows, width, device=dev)
    t = timeit(lambda: rms.normalize(nxt(), out32))
    print(json.dumps({"kernel": "rms_normalize", "rows": rows, "us": t * 1e6, "hbm_frac": rows * width * 8 / t / 1e9 / PEAK}))
    rms.eval()
    unit_list = [int(u) for u in sys.argv[3].split(',')] if len(sys.argv) > 3 else [2048, 4096]
    for units in unit_list:
        w = torch.randn(units, width, device=dev, generator=g) / width ** 0.5
        b = torch.randn(units, device=dev, generator=g) * 0.1
        layer = FirstLayer(w, b, "elu", rms)
        out = torch.empty(rows, units, dtype=torch.float16, device=dev)
        t = timeit(lambda: layer(nxt(), out), iters=30)
        bytes_ = rows * (width * 4 + units * 2)
        print(json.dumps({"kernel": "first_layer_kernel (normalise + linear + elu, fp16 out)", "rows": rows, "units": units,
                          "us": t * 1e6, "hbm_gbs": bytes_ / t / 1e9, "hbm_frac": bytes_ / t / 1e9 / PEAK,
                          "tflops": 2.0 * rows * width * units / t / 1e12}))
        wh, bh = w.half(), b.half()
        mean, den = rms.running_mean.float(), torch.sqrt(rms.running_var.float() + 1e-5)

        def torch_path():
            x = torch.clamp((nxt() - mean) / den, -5.0, 5.0)
            return torch.nn.functional.elu(torch.nn.functional.linear(x.half(), wh, bh))
        t2 = timeit(torch_path, iters=30)
        print(json.dumps({"kernel": "torch: normalise + cuBLAS fp16 linear + elu (library baseline)", "rows": rows,
                          "units": units, "us": t2 * 1e6, "speedup_of_fused": t2 / t}))
        if width <= 95:           # the fp32 (rollout-forward) variant: 3 x TF32 split products, fp32 out
            layer32 = FirstLayer(w, b, "elu", rms, precision="fp32")
            o32 = torch.empty(rows, units, dtype=torch.float32, device=dev)
            t3 = timeit(lambda: layer32(nxt(), o32), iters=30)
            bytes32 = rows * (width * 4 + units * 4)
            print(json.dumps({"kernel": "first_layer_f32_kernel (normalise + 3xTF32 linear + elu, fp32 out)", "rows": rows,
                              "units": units, "us": t3 * 1e6, "hbm_gbs": bytes32 / t3 / 1e9, "hbm_frac": bytes32 / t3 / 1e9 / PEAK,
                              "tflops_tf32_issued": 3 * 2.0 * rows * ((width + 8) // 8 * 8) * units / t3 / 1e12}))

            def torch_fp32():
                x = torch.clamp((nxt() - mean) / den, -5.0, 5.0)
                return torch.nn.functional.elu(torch.nn.functional.linear(x, w, b))
            t4 = timeit(torch_fp32, iters=10)
            print(json.dumps({"kernel": "torch: normalise + cuBLAS fp32 linear (TF32 off) + elu (library baseline)", "rows": rows,
                              "units": units, "us": t4 * 1e6, "speedup_of_fused": t4 / t3}))


if __name__ == "__main__":
    main()
