#!/usr/bin/env python
"""Benchmark of the fused obs+reward+reset task step (BASELINE.json metric: env-steps/sec and
% of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tilt] [--impl reference]

One "step" = one `post_physics_step` of the task (a single fused kernel: progress+=1, reward,
termination mask, flag updates, per-env reset, observations) over one batch of synthetic PhysX
state.  Default workload: BASELINE.json configs[2], humanoid_pingpong_3_actor_tilt at 65536 envs
per GPU (the size north_star states the roofline target at; configs[1], A3 at 16384 envs, is a
launch-latency-sized batch and is reported under "other_workloads").  Weak scaling: every rank owns
its own 65536-env shard, there is no data-path collective; the only collective is the 8-double
statistics all-reduce.

Timing hygiene: >= 3 warm-up steps; the step rotates over `--sets` (default 8) independent state
sets (~190 MB each for TILT, far beyond the 126 MB L2) so inputs are never L2-resident; timed on
the device with CUDA events around exactly K steps replayed from CUDA graphs; max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# workload name -> (variant, envs per GPU, BASELINE.json config it stands for, pre-step included)
WORKLOADS = {
    "tilt": ("tilt", 65536, "configs[2]: humanoid_pingpong_3_actor_tilt 65536 envs/GPU", False),
    "a3": ("a3", 16384, "configs[1]: HumanoidPingpong single-humanoid (A3) 16384 envs", False),
    "base": ("base", 4096, "configs[0]: humanoid_pingpong.py 4096 envs", False),
    "a4": ("a4", 65536, "configs[3]: humanoid_pingpong_4_actor_tilt shard", False),
    "adof": ("adof", 32768, "configs[3]: humanoid_pingpong_3_actor_all_dof shard", False),
    "align": ("align", 131072, "configs[4]: humanoid_pingpong_alignment full task step, 131072 envs/GPU", True),
    "align2": ("align2", 65536, "humanoid_pingpong_alignment reward definition #2 (two humanoids, last_hitter), 65536 envs/GPU", False),
    "nes": ("nes", 65536, "humanoid_pingpong_3_actor_tilt_no_earlystop 65536 envs/GPU", False),
    "tilt_1m": ("tilt", 1048576, "humanoid_pingpong_3_actor_tilt 1M envs/GPU (launch overhead amortised)", False),
}
# algorithmic bytes per env-step (A_core + 8 for the progress write-back; the
# full ALIGN step adds the pre-step 72 B; reset traffic is ~3 B/env-step and not counted)
ALGO_BYTES = {"base": 212 + 8, "a3": 708 + 8, "tilt": 718 + 8, "nes": 716 + 8, "align": 718 + 8, "a4": 1504 + 8, "align2": 1519 + 8,
              "adof": 3198 + 8}
PRE_STEP_BYTES = {"align": 72}
OTHER_STEPS = 2048          # timed steps of each context workload (other_workloads)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the ncu --set full captures under profiles/
NCU_TRAFFIC = {("tilt", 65536): 68.15e6 + 3.0e6, ("adof", 32768): 113.55e6 + 15.8e6, ("a4", 65536): 116.09e6 + 16.6e6}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        for bit, name in names.items():
            if r & bit and name != "gpu_idle":
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.sample()
            except Exception:
                break
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            try:
                self.sample()
            except Exception:
                pass
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def run_learner_side(tasks, units, peak, iters=40, warm=5):
    """SURVEY 8(f) rank 4 on the obs_buf the task step just wrote: RunningMeanStd update, and the fused
    normalise + first MLP layer (fp16 out).  CUDA events; the obs sets rotate, the fp16 output (rows x units
    x 2 B) is larger than L2 by itself."""
    import torch
    from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd
    obs = [t.obs_buf.view(-1, t.obs_buf.shape[-1]) for t in tasks]        # A4: one row per humanoid
    rows, width = obs[0].shape
    dev = obs[0].device
    g = torch.Generator(device=dev).manual_seed(0)
    rms = RunningMeanStd(width, device=dev)
    w = torch.randn(units, width, device=dev, generator=g) / width ** 0.5
    b = torch.randn(units, device=dev, generator=g) * 0.1
    layer = FirstLayer(w, b, "elu", rms)
    out = torch.empty(rows, units, dtype=torch.float16, device=dev)
    res = {}
    for name, fn in (("rms_update", lambda o: rms.update(o)), ("first_layer", lambda o: layer(o, out))):
        if name == "first_layer":
            rms.eval()
        for i in range(warm):
            fn(obs[i % len(obs)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(obs[i % len(obs)])
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / iters
        nbytes = rows * width * 4 + (rows * units * 2 if name == "first_layer" else 0)
        res[name] = {"us": sec * 1e6, "rows_per_s": rows / sec, "achieved_gbs": nbytes / sec / 1e9,
                     "roofline_frac": nbytes / sec / 1e9 / peak}
    res["first_layer"].update({"units": units, "dtype": "f16 operands, f32 accumulate (tcgen05), f16 out",
                               "tflops": 2.0 * rows * width * units / (res["first_layer"]["us"] * 1e-6) / 1e12,
                               "bound": "hbm (the [rows, units] fp16 write)"})
    return res


def make_tasks(variant, n, sets, device, seed_base):
    from isaacgym_b200.config import CONFIGS
    from isaacgym_b200.synth import make_state
    from isaacgym_b200.tasks import make_task
    cfg = CONFIGS[variant]
    tasks = []
    for s in range(sets):
        st = make_state(cfg, n, seed=seed_base + s, device=str(device), adversarial=False)
        if variant == "base":
            st["reset_ball_vel"] = st["reset_ball_vel"][:2].contiguous()
        tasks.append(make_task(variant, st, device=str(device)))
        del st
    return cfg, tasks


def time_steps(tasks, steps, warmup, with_pre, log_every, world, dist):
    """Exactly `steps` task steps replayed from CUDA graphs, CUDA-event timed; returns seconds
    (max over ranks) and the number of kernels launched in the timed region."""
    from isaacgym_b200 import _native as N
    sets = len(tasks)
    counter = {"i": 0}

    def one_step():
        i = counter["i"]
        t = tasks[i % sets]
        if with_pre:
            N.check(t._lib.ppk_pre_physics_step(t._task, t.buffers(), t._stream()), "pre")
        phases = N.PHASE_ALL if (log_every > 0 and i % log_every == 0) else (N.PHASE_ALL & ~N.PHASE_STATS)
        t._step(phases)
        counter["i"] = i + 1

    for _ in range(max(warmup, 3)):          # eager warm-up (also sets the kernels' smem attributes)
        one_step()
    torch.cuda.synchronize()
    chunk = min(steps, 512)
    full, rem = divmod(steps, chunk)
    graphs = []
    for count in ([chunk] if full else []) + ([rem] if rem else []):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(count):
                one_step()
        graphs.append((g, count))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    if full:
        for _ in range(full):
            graphs[0][0].replay()
    if rem:
        graphs[-1][0].replay()
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sec = start.elapsed_time(end) / 1e3
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    per_step = 1 + (1 if with_pre else 0) + (1 if tasks[0].cfg.variant == "adof" else 0)   # ADOF: + counter-clear kernel
    return sec, steps * per_step


def run_e2e(variant, n, steps, device, chunks=4):
    """The same step through the host-buffer C-ABI session: state tensors in pinned HOST memory,
    H2D of the step's inputs and D2H of its results inside the timed region (wall clock: the call
    returns when the results are on the host)."""
    from isaacgym_b200.config import CONFIGS
    from isaacgym_b200.host_session import HostSession
    from isaacgym_b200.synth import make_state
    cfg = CONFIGS[variant]
    st = make_state(cfg, n, seed=4242, device="cpu", adversarial=False)
    sess = HostSession(cfg, st, num_chunks=chunks)
    try:
        for _ in range(4):                      # eager, graph capture, two replays
            sess.post_physics_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            sess.post_physics_step()
        sec = time.perf_counter() - t0
        h2d, d2h = sess.traffic()
    finally:
        sess.close()
    return sec, h2d, d2h


def run_cpu(variant, n, steps, warmup, threads):
    """The oracle port (same ATen op sequence as the reference functions) on the host CPU."""
    from isaacgym_b200.config import CONFIGS
    from isaacgym_b200.synth import make_state
    from oracle import task_oracle
    cfg = CONFIGS[variant]
    torch.set_num_threads(threads)
    st = make_state(cfg, n, seed=777, device="cpu", adversarial=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if variant == "base":
            st["progress_buf"] += 1
            st["obs_buf"][:] = task_oracle.compute_observations(cfg, st)
            task_oracle.compute_reward(cfg, st)
        else:
            task_oracle.post_physics_step(cfg, st)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tilt", choices=sorted(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-extras", action="store_true", help="skip other_workloads / e2e / cpu_baseline")
    args = ap.parse_args()

    variant, n_default, workload_desc, with_pre = WORKLOADS[args.workload]
    n = args.envs_per_gpu or n_default
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        # the reference's own CPU path: its ATen op chains (oracle port; /root/reference cannot travel
        # to the GPU box) on all host threads; each step is the full batch of the workload
        if rank != 0:
            return
        steps = max(1, min(args.steps, 20))
        warm = max(1, min(args.warmup, 3))
        times = run_cpu(variant, n, steps, warm, cores)
        sec = sum(times)
        value = n * len(times) / sec
        line = {"impl": "reference", "metric": "env-steps/sec of fused obs+reward+reset", "value": value,
                "unit": "env-steps/s", "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
                "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_desc, "variant": variant, "envs_per_gpu": n,
                           "note": "CPU run of the reference's ATen op sequence (oracle port), one process"},
                "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                 "sample": f"{len(times)} steps of the full {n}-env batch"},
                "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from isaacgym_b200 import _native as N
    from isaacgym_b200.config import CONFIGS
    N.load()

    cfg, tasks = make_tasks(variant, n, args.sets, device, seed_base=1000 * CONFIGS[variant].variant_id + 17 * rank)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    sec, launches = time_steps(tasks, args.steps, args.warmup, with_pre, cfg.log_every, world, dist)
    clocks = sampler.stop()
    # the only collective of the path: 8 doubles, off the critical path.  One logging step outside
    # the timed region gives a clean sample (the slots also hold the logging steps of the timed loop).
    tasks[0].stats.slots.zero_()
    tasks[0]._step(N.PHASE_ALL)
    tasks[0].stats.reduce(tasks[0]._lib, tasks[0]._stream())
    stat_means = tasks[0].stats.means(n * world)
    torch.cuda.synchronize()

    total_envs = n * world
    value = total_envs * args.steps / sec
    ms_per_step = 1e3 * sec / args.steps
    peak, peak_src = measured_hbm_peak()
    algo = ALGO_BYTES[variant] + (PRE_STEP_BYTES.get(variant, 0) if with_pre else 0)
    achieved = algo * n / (sec / args.steps) / 1e9          # per GPU: every rank runs the same shard size
    line = {
        "metric": "env-steps/sec of fused obs+reward+reset", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc, "variant": variant, "envs_per_gpu": n, "global_envs": total_envs,
                   "parallelism": f"env-sharded x{world}, stats all-reduce only",
                   "l2_defeat": f"rotating {args.sets} independent state sets per GPU (inputs larger than L2)",
                   "pre_physics_step_in_step": with_pre, "launch": "CUDA graph replay"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC.get((variant, n)), "algorithmic_bytes_per_env": algo, "peak_source": peak_src,
                     "kernel": f"family_step_kernel<{variant}>" if variant not in ("base", "adof") else f"{variant}_step_kernel",
                     "note": "achieved = algorithmic bytes per launch / (CUDA-event time of the timed region / launches). "
                             "traffic = dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture "
                             "(profiles/): reads are ~1.9x the algorithmic read bytes because the AoS rows are 52 B and "
                             "DRAM moves 64-B granules; most of the 343 B/env of outputs are still in L2 when the launch "
                             "ends and reach DRAM later, so the step's real traffic is ~1.36 kB/env vs 726 B algorithmic"},
        "stats_sample": {k: stat_means[k] for k in ("reward_sum", "progress_sum", "reset_count")},
    }
    if rank == 0 and not args.no_extras and variant != "base":
        # the learner side of the path, context only
        try:
            line["learner_side"] = run_learner_side(tasks, 2048, peak)
        except Exception as e:  # noqa: BLE001
            line["learner_side"] = {"error": str(e)[:200]}
    del tasks
    torch.cuda.empty_cache()

    if rank == 0 and not args.no_extras:
        # e2e through the host-buffer C-ABI session, this rank's shard (N=1: the whole job)
        try:
            e_sec, h2d, d2h = run_e2e(variant, n, args.e2e_steps, device)
            line["e2e"] = {"value": n * args.e2e_steps / e_sec * world, "unit": "env-steps/s",
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                           "note": "ppk_host_post_physics_step: pinned host state tensors, chunked H2D/kernel/D2H "
                                   "pipeline; measured on rank 0's shard and scaled by the rank count"}
        except Exception as e:  # noqa: BLE001
            line["e2e"] = {"value": None, "unit": "env-steps/s", "error": str(e)[:200]}
    if rank == 0 and world == 1 and not args.no_extras:
        # CPU baseline: the oracle port on the host cores, bounded sample of the same workload
        times = run_cpu(variant, n, 10, 2, cores)
        line["cpu_baseline"] = {"value": n / statistics.median(times), "unit": "env-steps/s", "cores": cores,
                                "kind": "port", "sample": f"median of {len(times)} steps of the full {n}-env batch, "
                                f"torch.set_num_threads({cores})"}
        times1 = run_cpu(variant, min(n, 16384), 5, 1, 1)
        line["cpu_baseline"]["value_1_thread"] = min(n, 16384) / statistics.median(times1)
        # the other BASELINE.json configs, short runs (parity-tested elsewhere; context only)
        others = {}
        for name in ("base", "a3", "align", "a4", "adof", "tilt_1m"):
            if name == args.workload:
                continue
            try:
                v2, n2, desc2, pre2 = WORKLOADS[name]
                sets2 = 8 if n2 <= 131072 else 2
                cfg2, tasks2 = make_tasks(v2, n2, sets2, device, seed_base=99)
                s2, _ = time_steps(tasks2, OTHER_STEPS, 16, pre2, 0, 1, None)
                algo2 = ALGO_BYTES[v2] + (PRE_STEP_BYTES.get(v2, 0) if pre2 else 0)
                ach2 = algo2 * n2 / (s2 / OTHER_STEPS) / 1e9
                others[name] = {"workload": desc2, "envs": n2, "value": n2 * OTHER_STEPS / s2, "unit": "env-steps/s",
                                "ms_per_step": 1e3 * s2 / OTHER_STEPS, "roofline_frac": ach2 / peak, "achieved_gbs": ach2,
                                "state_sets": sets2}
                del tasks2
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": str(e)[:200]}
        line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
