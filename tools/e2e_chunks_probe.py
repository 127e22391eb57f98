import sys, time, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from bench import run_e2e
for v, n in (("tilt", 65536),):
    for chunks in (1, 2, 3, 4, 6, 8, 12, 16):
        sec, h2d, d2h = run_e2e(v, n, 20, "cuda", chunks=chunks)
        print(v, n, "chunks", chunks, "ms/step %.3f" % (sec / 20 * 1e3), "M env-steps/s %.1f" % (n * 20 / sec / 1e6), h2d, d2h, flush=True)
