"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous env sharding and the one
collective of the path, the SUM all-reduce of the 8-double statistics vector."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.stats import EpisodeStats, shard_range
from isaacgym_b200.synth import clone_state, make_state
from oracle import task_oracle


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 64, 65536, 262144 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a and d >= c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = CONFIGS["tilt"]
        st = make_state(cfg, n, seed=5)                  # every rank builds the same global batch ...
        lo, hi = shard_range(n, rank, world)             # ... and steps only its own env block
        shard = {k: (v[lo:hi].clone() if (v.dim() > 0 and v.shape[0] == n) else v.clone()) for k, v in st.items()}
        shard["actor_indices"] = torch.arange((hi - lo) * cfg.num_actors, dtype=torch.int64)   # per-process sim
        shard["dof_indices"] = torch.arange(hi - lo, dtype=torch.int64)
        _, _, s = task_oracle.post_physics_step(cfg, shard)
        stats = EpisodeStats("cpu")
        stats.local[0] = float(s["reward_sum"])
        stats.local[1] = float(s["progress_sum"])
        stats.local[2] = float(s["reset_count"])
        stats.all_reduce()
        m = stats.means(n)
        torch.save({"means": m, "obs": shard["obs_buf"], "reset": shard["reset_buf"], "lo": lo, "hi": hi},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_step_and_stats_allreduce(tmp_path):
    n, world = 1000, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    cfg = CONFIGS["tilt"]
    whole = clone_state(make_state(cfg, n, seed=5))
    _, _, s = task_oracle.post_physics_step(cfg, whole)
    parts = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    # per-env outputs of the shards concatenate to the unsharded result, bit for bit
    assert torch.equal(torch.cat([p["obs"] for p in parts]), whole["obs_buf"])
    assert torch.equal(torch.cat([p["reset"] for p in parts]), whole["reset_buf"])
    # every rank holds the same global means after the all-reduce
    for p in parts:
        m = p["means"]
        assert abs(m["reward_sum"] - float(s["reward_sum"]) / n) < 1e-9 * max(1.0, abs(float(s["reward_sum"])))
        assert m["progress_sum"] == float(s["progress_sum"]) / n
        assert m["reset_count"] == float(s["reset_count"]) / n
    assert len(N.STAT_NAMES) == N.PPK_NUM_STATS
