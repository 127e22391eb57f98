"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the REFERENCE's own functions.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden [variant ...]   # writes tests/golden/<variant>.npz (default: all)

For each task variant a small seeded synthetic state (inputs included in the
fixture, so it is self-contained) is pushed through the reference's
`compute_*_reward` and `compute_*_observations` functions, cut unmodified out of
/root/reference by `oracle/ref_extract.py`, composed exactly as the class
wrappers do (`oracle/task_oracle.py`).  The GPU box has no /root/reference;
there the fixtures are what pins both the oracle and the CUDA path.

`<variant>_step.npz` (a3, tilt, nes, align, adof): the same for one whole `post_physics_step` executed from the
reference's own METHOD bodies (`oracle/ref_methods.py`: progress += 1 -> compute_reward -> nonzero(reset_buf) ->
reset_idx / _reset_idx with its host `random.uniform` draws -> compute_observations, ADOF's counter clear included).
The launch values the reference drew are stored in the rows of `in__reset_ball_vel` (/ `in__reset_ball_pos_yz`) of
the envs it reset, which is where the fused CUDA step reads them.
"""
import random
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from isaacgym_b200.config import CONFIGS  # noqa: E402
from isaacgym_b200.synth import clone_state, make_state  # noqa: E402
from oracle import ref_extract, task_oracle  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NUM_ENVS = 96
SEED = 20261018


STEP_SEED = 20261019


def make_step_fixtures(only):
    from oracle import ref_methods as M
    for variant in M.STEP_VARIANTS:
        if only and variant not in only:
            continue
        cfg = CONFIGS[variant]
        st_in = make_state(cfg, NUM_ENVS, seed=STEP_SEED + cfg.variant_id)
        st_in["progress_buf"][[3, 40, 77]] = cfg.max_episode_length - 2          # time-outs: every variant resets some envs
        st = clone_state(st_in)
        M.bind_step_functions(variant)
        task = M.make_ref_task(cfg, st)
        random.seed(STEP_SEED)
        with ref_extract.quiet():
            task.post_physics_step()
        env_ids = st["reset_buf"].nonzero(as_tuple=False).flatten()
        b = cfg.ball_actor
        st_in["reset_ball_vel"][env_ids] = st["root_states"][env_ids, b, 7:10]      # what the reference drew
        if variant == "adof":
            st_in["reset_ball_pos_yz"][env_ids] = st["root_states"][env_ids, b, 1:3]
        out = {"in__" + k: v.numpy() for k, v in st_in.items()}
        for key in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "root_states", "dof_states") + cfg.flag_names + cfg.counter_names:
            out["out__" + key] = st[key].numpy()
        out["out__reset_actor_indices"] = task.gym.last("set_actor_root_state_tensor_indexed")[2].numpy()
        path = os.path.join(GOLDEN_DIR, f"{variant}_step.npz")
        np.savez_compressed(path, **out)
        print(f"{variant}_step: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), resets={len(env_ids)}")


def main():
    assert ref_extract.available(), "needs /root/reference"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    steps_only = "--steps-only" in sys.argv
    if steps_only:
        sys.argv.remove("--steps-only")
        make_step_fixtures(set(sys.argv[1:]))
        return
    ref = task_oracle.ReferenceImpl()
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    for variant, cfg in CONFIGS.items():
        if only and variant not in only:
            continue
        st_in = make_state(cfg, NUM_ENVS, seed=SEED + cfg.variant_id)
        st = clone_state(st_in)
        if variant != "base":
            st["progress_buf"] += 1                      # post_physics_step does this before compute_reward
        with ref_extract.quiet():
            task_oracle.compute_reward(cfg, st, ref)
            obs = task_oracle.compute_observations(cfg, st, ref)
        out = {"in__" + k: v.numpy() for k, v in st_in.items()}
        out["out__obs_buf"] = obs.numpy()
        out["out__rew_buf"] = st["rew_buf"].numpy()
        out["out__reset_buf"] = st["reset_buf"].numpy()
        for name in cfg.flag_names + cfg.counter_names + cfg.state_names:
            out["out__" + name] = st[name].numpy()
        path = os.path.join(GOLDEN_DIR, f"{variant}.npz")
        np.savez_compressed(path, **out)
        print(f"{variant}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), "
              f"resets={int(st['reset_buf'].sum())}, mean reward={float(st['rew_buf'].mean()):.4f}")
    make_step_fixtures(only)


if __name__ == "__main__":
    main()
