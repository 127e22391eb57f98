"""VecTask.step envelope: action clamp inside the pre-step kernel,
timeout_buf, and the compacted int32 actor / dof index lists of the envs the fused step resets."""
import pytest
import torch

from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import clone_state, make_state
from isaacgym_b200.tasks import make_task
from oracle import task_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("variant", ["a3", "tilt", "nes", "align", "a4", "adof", "align2"])
def test_fused_step_emits_reset_index_lists_and_timeouts(variant):
    cfg = CONFIGS[variant]
    n = 3000
    st = make_state(cfg, n, seed=17)
    o = clone_state(st)
    p_before = o["progress_buf"].clone()
    env_ids, (want_actor, want_dof), _ = task_oracle.post_physics_step(cfg, o)
    task = make_task(variant, st, device=DEV, envelope=True)
    task.post_physics_step()
    got_actor, got_dof = task.reset_indices()
    k = int(task.reset_count.item())
    assert k == len(env_ids)
    A = cfg.num_actors
    # env order inside the lists is unspecified: compare as sets of per-env rows
    ga = got_actor.cpu().view(k, A)
    wa = want_actor.view(k, A)
    assert torch.equal(ga[ga[:, 0].argsort()], wa[wa[:, 0].argsort()])
    dof_per = want_dof.numel() // max(k, 1)
    gd, wd = got_dof.cpu().view(k, dof_per), want_dof.view(k, dof_per)
    assert torch.equal(gd[gd[:, 0].argsort()], wd[wd[:, 0].argsort()])
    assert got_actor.dtype == torch.int32
    want_timeout = (p_before + 1 >= cfg.max_episode_length - 1).to(torch.int64)
    assert torch.equal(task.timeout_buf.cpu(), want_timeout)


def test_step_clamps_actions_in_the_pre_step_kernel():
    cfg = CONFIGS["tilt"]
    n = 1024
    st = make_state(cfg, n, seed=3)
    task = make_task("tilt", st, device=DEV, clip_actions=0.5, envelope=True)
    actions = (torch.rand(n, cfg.num_dofs, generator=torch.Generator().manual_seed(1)) * 4 - 2).to(DEV)
    obs, rew, reset, extras = task.step(actions)
    clamped = torch.clamp(actions, -0.5, 0.5)
    assert torch.equal(task.actions, clamped)
    want = task._pd_action_offset + task._pd_action_scale * clamped
    torch.testing.assert_close(task.pd_tar, want, rtol=1e-6, atol=1e-7)
    assert obs["obs"].abs().max() <= task.clip_obs and "time_outs" in extras
    assert rew.shape == (n,) and reset.dtype == torch.int64
