"""Two-GPU NCCL tests (skipped on a single-GPU box): the collectives of the path are scalar-sized --
the episode-statistics all-reduce of the task step and the moments all-reduce of RunningMeanStd."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from isaacgym_b200 import _native as N
        from isaacgym_b200.config import CONFIGS
        from isaacgym_b200.policy_input import RunningMeanStd
        from isaacgym_b200.stats import shard_range
        from isaacgym_b200.synth import make_state
        from isaacgym_b200.tasks import make_task
        cfg = CONFIGS["tilt"]
        n = 6000
        st = make_state(cfg, n, seed=9, adversarial=False)        # same global batch on every rank
        lo, hi = shard_range(n, rank, world)
        shard = {k: (v[lo:hi].clone() if (isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n) else v)
                 for k, v in st.items()}
        task = make_task("tilt", shard, device=str(dev), log_stats=True)
        task._step(N.PHASE_ALL)
        task.stats.reduce(task._lib, task._stream())        # slots -> local -> all_reduce(SUM) on a side stream
        means = task.stats.means(n)
        rms = RunningMeanStd(cfg.num_obs, device=dev)
        rms.update(task.obs_buf, group=dist.group.WORLD)
        torch.cuda.synchronize()
        torch.save({"means": means, "obs": task.obs_buf.cpu(), "rew": task.rew_buf.cpu(), "progress": task.progress_buf.cpu(),
                    "reset": task.reset_buf.cpu(), "mean": rms.running_mean.cpu(), "var": rms.running_var.cpu(),
                    "count": float(rms.count)}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_stats_and_moments_allreduce(tmp_path):
    from oracle import policy_oracle as P
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(world)]
    obs = torch.cat([x["obs"] for x in r])
    rew = torch.cat([x["rew"] for x in r])
    # both ranks hold the statistics of the GLOBAL batch
    ref = P.RunningMeanStd(obs.shape[1])
    ref.update(obs)
    for x in r:
        torch.testing.assert_close(x["mean"], ref.running_mean, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(x["var"], ref.running_var, rtol=1e-6, atol=1e-7)
        assert x["count"] == float(ref.count)
        assert abs(x["means"]["reward_sum"] - float(rew.double().mean())) <= 1e-6 * max(1.0, abs(float(rew.double().mean())))
        assert x["means"] == r[0]["means"]
    torch.testing.assert_close(r[0]["mean"], r[1]["mean"], rtol=0, atol=0)
