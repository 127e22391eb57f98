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
the device with CUDA events around exactly K steps replayed from CUDA graphs (every graph is replayed
once, untimed, before the events: the first replay of a freshly instantiated graph is slower); max
over ranks.  Every `log_every` steps (TILT:763, A3:741, ADOF:860) the step accumulates the logged
sums and -- inside the timed region, on a side stream -- they are folded and all-reduced over the ranks
(8 doubles, NCCL): the only collective of the path, shown to stay off the step's critical path.
At N > 1 rank 0's line also carries BASELINE.json configs[3] (A4 and ADOF, 262144 envs sharded over the
ranks) and configs[4] (ALIGN pre + post step, 131072 envs per GPU), timed on all ranks at once.
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


def ncu_traffic(variant, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel, parsed by profiles/summarize.py
    from the committed `ncu --set full` captures into profiles/ncu_traffic.json (None when no capture matches)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = d.get(f"{variant}@{n}")
        return None if e is None else float(e["dram_read_bytes"]) + float(e["dram_write_bytes"])
    except Exception:
        return None


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
    kinds = [("rms_update", lambda o: rms.update(o)), ("first_layer", lambda o: layer(o, out))]
    if width <= 95:          # the rollout forward's fp32 variant (3 x TF32 split products, fp32 out)
        layer32 = FirstLayer(w, b, "elu", rms, precision="fp32")
        out32 = torch.empty(rows, units, dtype=torch.float32, device=dev)
        kinds.append(("first_layer_fp32", lambda o: layer32(o, out32)))
    res = {}
    for name, fn in kinds:
        if name != "rms_update":
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
        nbytes = rows * width * 4 + rows * units * {"rms_update": 0, "first_layer": 2, "first_layer_fp32": 4}[name]
        res[name] = {"us": sec * 1e6, "rows_per_s": rows / sec, "achieved_gbs": nbytes / sec / 1e9,
                     "roofline_frac": nbytes / sec / 1e9 / peak}
    # the same update fed by the step kernel itself (PPK_PHASE_MOMENTS): step + fold vs step, then + separate update
    from isaacgym_b200 import _native as N
    if tasks[0].cfg.variant not in ("base", "adof"):
        for t in tasks:
            t.fused_moments = True
            t.obs_moments = torch.zeros(N.PPK_MOMENT_SLOTS, 2 * width, dtype=torch.float64, device=dev)
            t._buffers = None
        rms2 = RunningMeanStd(width, device=dev)
        ph = N.PHASE_ALL & ~N.PHASE_STATS

        def timed(fn):
            for i in range(warm):
                fn(tasks[i % len(tasks)])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(iters):
                    fn(tasks[i % len(tasks)])
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / iters

        def plain(t):
            t._step(ph)

        def with_update(t):
            t._step(ph)
            rms2.update(t.obs_buf.view(-1, width))

        def fused(t):
            t._step(ph | N.PHASE_MOMENTS)
            rms2.update_from_step(t.obs_moments, rows)

        res["step_plus_rms_update"] = {"step_us": timed(plain), "step_then_update_us": timed(with_update),
                                       "step_with_fused_moments_us": timed(fused),
                                       "note": "CUDA-graph replay of (task step [+ RunningMeanStd update]) over the rotating state sets"}
        for t in tasks:
            t.fused_moments = False
            t._buffers = None
    res["first_layer"].update({"units": units, "dtype": "f16 operands, f32 accumulate (tcgen05), f16 out",
                               "tflops": 2.0 * rows * width * units / (res["first_layer"]["us"] * 1e-6) / 1e12,
                               "bound": "hbm (the [rows, units] fp16 write)"})
    if "first_layer_fp32" in res:
        res["first_layer_fp32"].update({"units": units, "dtype": "3 x tf32 split products, f32 accumulate (tcgen05), f32 out",
                                        "bound": "hbm (the [rows, units] fp32 write)"})
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
    """Exactly `steps` task steps replayed from CUDA graphs, CUDA-event timed; returns seconds (max over ranks), the
    number of our kernels launched in the timed region and the last all-reduced statistics vector.

    Logging steps (i % log_every == 0, as the reference logs: TILT:763, A3:741, ADOF:860) accumulate the statistics;
    right after each of them the graph FORKS: a side branch folds the slots (`ppk_stats_reduce`) and all-reduces the 8
    doubles over the ranks (NCCL, captured into the graph) while the main branch carries on with the next steps; the
    branches join before the next logging step / at the end of the graph.  The collective is inside the timed region
    and off the step's critical path.  If this torch / NCCL build cannot capture the collective, it is issued eagerly
    on the side stream after each graph instead (same work, one exposed latency at the end of the region)."""
    from isaacgym_b200 import _native as N
    sets = len(tasks)
    dev = tasks[0].device
    lib = tasks[0]._lib
    adof = tasks[0].cfg.variant == "adof"
    side = torch.cuda.Stream(dev)
    local = torch.zeros(N.PPK_NUM_STATS, dtype=torch.float64, device=dev)
    total = torch.zeros(N.PPK_NUM_STATS, dtype=torch.float64, device=dev)
    align = torch.zeros(1, dtype=torch.float32, device=dev)

    def one_step(i, log):
        t = tasks[i % sets]
        if with_pre:
            N.check(t._lib.ppk_pre_physics_step(t._task, t.buffers(), t._stream()), "pre")
        t._step(N.PHASE_ALL if log else (N.PHASE_ALL & ~N.PHASE_STATS))

    def fold(i, collective):
        """on the current (side) stream: slots of the logging step's task -> local -> total (-> all-reduce)"""
        slots = tasks[i % sets].stats.slots
        N.check(lib.ppk_stats_reduce(slots.data_ptr(), local.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "ppk_stats_reduce")
        total.copy_(local, non_blocking=True)
        if collective and world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)      # 8 doubles over NVLink

    for i in range(max(warmup, 3)):          # eager warm-up (also sets the kernels' smem attributes)
        one_step(i, False)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)          # the communicator exists before anything is captured
    torch.cuda.synchronize()

    def is_log(i):
        return log_every > 0 and i % log_every == 0

    def capture(first, cnt, collective):
        g = torch.cuda.CUDAGraph()
        main = torch.cuda.current_stream(dev)
        with torch.cuda.graph(g):
            cap = torch.cuda.current_stream(dev)
            forked = False
            for j in range(cnt):
                i = first + j
                if is_log(i):
                    if forked:
                        cap.wait_stream(side)                 # the previous fold has left the slots
                    one_step(i, True)
                    side.wait_stream(cap)                     # fork
                    with torch.cuda.stream(side):
                        fold(i, collective)
                    forked = True
                else:
                    one_step(i, False)
            if forked:
                cap.wait_stream(side)                         # join
        del main
        return g

    # segments of <= 512 steps; segments that look the same (start set, length, phase of the logging cadence) share a graph
    segs, i = [], 0
    while i < steps:
        cnt = min(steps - i, 512)
        segs.append((i, cnt))
        i += cnt
    period = log_every if log_every > 0 else 1
    in_graph_collective = True
    graphs = {}
    for first, cnt in segs:
        key = (first % sets, cnt, first % period)
        if key in graphs:
            continue
        if in_graph_collective:
            try:
                graphs[key] = capture(first, cnt, True)
            except Exception:  # noqa: BLE001 -- this build cannot capture the collective: fold in the graph, all-reduce eagerly
                torch.cuda.synchronize()
                in_graph_collective = False
        if not in_graph_collective:
            graphs[key] = capture(first, cnt, False)
    torch.cuda.synchronize()
    for g in graphs.values():                # untimed first replay of every graph
        g.replay()
    for t in tasks:
        t.stats.slots.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    main = torch.cuda.current_stream(dev)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    folds = sum(1 for i in range(steps) if is_log(i))
    eager = world > 1 and not in_graph_collective and folds > 0
    # keep the GPU busy (~1 ms) while the host enqueues the start event and the first graph launch: the timed region
    # then holds device time of the K steps, not the host's cudaGraphLaunch latency (8 us against 300 us at --steps 20;
    # occasionally far more when the launch waits for the driver lock behind an NVML query of the clock sampler, which
    # showed as a +3..10 % outlier of one rank in one of three runs with a 100-us sleep)
    torch.cuda._sleep(2000000)
    if world > 1:
        # device-side start line: the ranks' streams leave this tiny all-reduce together, so the timed region of a
        # rank whose host came out of the barrier early does not include waiting (inside the captured collective)
        # for a rank whose host came out late -- measured at N=8: up to 1.5 ms of such skew on the first region
        dist.all_reduce(align, op=dist.ReduceOp.SUM)
    start.record()
    for first, cnt in segs:
        graphs[(first % sets, cnt, first % period)].replay()
        if eager and any(is_log(i) for i in range(first, first + cnt)):
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dist.all_reduce(total, op=dist.ReduceOp.SUM)
    if eager:
        main.wait_stream(side)               # the last collective ends inside the timed region
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sec = start.elapsed_time(end) / 1e3
    time_steps.rank_sec = sec
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    per_step = 1 + (1 if with_pre else 0) + (1 if adof else 0)   # ADOF: + counter-clear kernel
    time_steps.collective_mode = ("captured in the step graph" if in_graph_collective else "eager on the side stream") if world > 1 else "single rank: fold only"
    return sec, steps * per_step + folds, total.cpu().tolist()


time_steps.collective_mode = ""


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


def config_dict(workload_desc, variant, n, world, sets, with_pre):
    """`config` of the JSON line -- the same dict for our arm and the reference arm (same workload, same N)."""
    return {"workload": workload_desc, "variant": variant, "envs_per_gpu": n, "global_envs": n * world,
            "parallelism": f"env-sharded x{world}, stats all-reduce only",
            "l2_defeat": f"rotating {sets} independent state sets per GPU (inputs larger than L2)",
            "pre_physics_step_in_step": with_pre, "launch": "CUDA graph replay"}


def bench_workload(name, n, sets, steps, warmup, device, rank, world, dist, peak):
    """One context workload (BASELINE.json configs other than the headline one), all ranks at once."""
    from isaacgym_b200.config import CONFIGS
    v, _, desc, pre = WORKLOADS[name]
    cfg, tasks = make_tasks(v, n, sets, device, seed_base=99 + 17 * rank)
    sec, _, _ = time_steps(tasks, steps, warmup, pre, cfg.log_every, world, dist)
    algo = ALGO_BYTES[v] + (PRE_STEP_BYTES.get(v, 0) if pre else 0)
    ach = algo * n / (sec / steps) / 1e9
    del tasks
    torch.cuda.empty_cache()
    return {"workload": desc, "envs_per_gpu": n, "global_envs": n * world, "value": n * world * steps / sec,
            "unit": "env-steps/s", "ms_per_step": 1e3 * sec / steps, "roofline_frac": ach / peak, "achieved_gbs": ach,
            "state_sets": sets, "steps": steps}


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
    warm = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own CPU path: its ATen op chains (oracle port; /root/reference cannot travel
        # to the GPU box) on all host threads; each step is the full batch of the workload
        if rank != 0:
            return
        steps = max(1, min(args.steps, 20))
        times = run_cpu(variant, n, steps, min(warm, 3), cores)
        sec = sum(times)
        value = n * len(times) / sec
        line = {"impl": "reference", "metric": "env-steps/sec of fused obs+reward+reset", "value": value,
                "unit": "env-steps/s", "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
                "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(workload_desc, variant, n, args.gpus, args.sets, with_pre),
                "reference_note": "CPU run of the reference's ATen op sequence (oracle port), one process on rank 0's host "
                                  f"cores; at most 20 timed steps of the full {n}-env batch",
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
    sec, launches, totals = time_steps(tasks, args.steps, args.warmup, with_pre, cfg.log_every, world, dist)
    clocks = sampler.stop()
    stat_means = {name: x / (n * world) for name, x in zip(N.STAT_NAMES, totals)}   # the last logging step's all-reduced sums
    rank_ms = None
    if world > 1:        # every rank's own device time of the timed region (the line reports their maximum)
        mine = torch.tensor([time_steps.rank_sec], dtype=torch.float64, device="cuda")
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        rank_ms = [round(1e3 * float(x.item()) / args.steps, 5) for x in everyone]

    total_envs = n * world
    value = total_envs * args.steps / sec
    ms_per_step = 1e3 * sec / args.steps
    peak, peak_src = measured_hbm_peak()
    algo = ALGO_BYTES[variant] + (PRE_STEP_BYTES.get(variant, 0) if with_pre else 0)
    achieved = algo * n / (sec / args.steps) / 1e9          # per GPU: every rank runs the same shard size
    traffic = ncu_traffic(variant, n)
    line = {
        "metric": "env-steps/sec of fused obs+reward+reset", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(workload_desc, variant, n, world, args.sets, with_pre),
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_env": algo, "peak_source": peak_src,
                     "kernel": {"base": "base_step_kernel", "adof": "adof2_step_kernel"}.get(variant, f"family_step_kernel<{variant}>"),
                     "note": "achieved = algorithmic bytes per launch / (CUDA-event time of the timed region / launches). "
                             "traffic = dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture "
                             "(profiles/ncu_traffic.json): the AoS rows are 52 B, the TMA engine fetches 64-B granules, and most "
                             "of the 343 B/env of outputs are still in L2 when the launch ends and reach DRAM later"},
        "stats_collective": {"every_steps": cfg.log_every, "in_timed_region": True, "stream": "side branch of the step graph",
                             "payload": "8 doubles, all-reduce(SUM)", "mode": time_steps.collective_mode},
        "stats_sample": {k: stat_means[k] for k in ("reward_sum", "progress_sum", "reset_count")},
    }
    if rank_ms is not None:
        line["rank_ms_per_step"] = rank_ms
    if rank == 0 and not args.no_extras and variant != "base":
        # the learner side of the path, context only
        try:
            line["learner_side"] = run_learner_side(tasks, 2048, peak)
        except Exception as e:  # noqa: BLE001
            line["learner_side"] = {"error": str(e)[:200]}
    del tasks
    torch.cuda.empty_cache()

    if not args.no_extras:
        # e2e through the host-buffer C-ABI session: EVERY rank runs its own shard at the same time (they share the
        # host's memory and PCIe complex), the job's throughput is all envs over the slowest rank's wall time
        try:
            if world > 1:
                dist.barrier()
            e_sec, h2d, d2h = run_e2e(variant, n, args.e2e_steps, device)
            if world > 1:
                t = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e_sec = float(t.item())
            line["e2e"] = {"value": n * world * args.e2e_steps / e_sec, "unit": "env-steps/s",
                           "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
                           "note": "ppk_host_post_physics_step: pinned host state tensors, chunked H2D/kernel/D2H "
                                   "pipeline; measured on all ranks at once, max wall time over the ranks"}
        except Exception as e:  # noqa: BLE001
            line["e2e"] = {"value": None, "unit": "env-steps/s", "error": str(e)[:200]}
    if not args.no_extras and world > 1:
        # BASELINE.json configs[3] (A4 and ADOF, 262144 envs sharded over the ranks: strong scaling) and configs[4]
        # (ALIGN pre + post step at 131072 envs per GPU), all ranks at once, the stats all-reduce in the loop
        others = {}
        for name, n2, sets2 in (("a4", 262144 // world, 4), ("adof", 262144 // world, 4), ("align", 131072, 4)):
            try:
                r = bench_workload(name, n2, sets2, 1024, 16, device, rank, world, dist, peak)
                r["scaling"] = "strong (262144 envs over the ranks)" if name != "align" else "weak (131072 envs per GPU)"
                others[name] = r
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": str(e)[:200]}
        line["other_workloads"] = others
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
                n2 = WORKLOADS[name][1]
                others[name] = bench_workload(name, n2, 8 if n2 <= 131072 else 2, OTHER_STEPS, 16, device, 0, 1, None, peak)
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
