"""Closed-loop use of the whole API the way a trainer would drive it: VecTask.step() with the
envelope on, a synthetic physics callback that moves the ball, and the device sampler refreshing
the launch table after every step.  Checks the invariants that must hold after any number of steps."""
import pytest
import torch

from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import make_state
from isaacgym_b200.tasks import make_task

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("variant", ["tilt", "nes", "a4", "adof", "align2"])
def test_sixty_steps_keep_every_invariant(variant):
    cfg = CONFIGS[variant]
    n = 4096
    st = make_state(cfg, n, seed=55, adversarial=False)
    b = cfg.ball_actor
    gen = torch.Generator(device=DEV).manual_seed(9)

    def physics(task):
        r = task.root_states
        r[:, b, 0:3] += 0.0166 * r[:, b, 7:10]
        r[:, b, 9] -= 9.81 * 0.0166
        hit = torch.rand(n, generator=gen, device=DEV) < 0.05
        r[:, b, 7] = torch.where(hit, r[:, b, 7].abs(), r[:, b, 7])

    task = make_task(variant, st, device=DEV, envelope=True, physics=physics, log_stats=True, launch_seed=42,
                     clip_observations=5.0)
    # envs at random phases of their episode, so time-outs happen within the 60 steps for every variant
    task.progress_buf.copy_(torch.randint(0, cfg.max_episode_length - 2, (n,), generator=gen, device=DEV))
    task.sample_ball_launch(seed=42, epoch=0)
    init_root = task.initial_vec_root_states.clone()
    total_resets = 0
    for step in range(60):
        prog_before = task.progress_buf.clone()
        table_before = task.st["reset_ball_vel"].clone()
        actions = torch.rand(n, cfg.num_dofs, generator=gen, device=DEV) * 3 - 1.5
        obs, rew, reset, extras = task.step(actions)      # redraws the consumed launch rows itself (launch_seed)
        rst = reset.bool()
        total_resets += int(rst.sum())
        # progress: +1, or 0 where the env reset; time-outs are resets
        assert torch.equal(task.progress_buf[~rst], prog_before[~rst] + 1)
        assert int(task.progress_buf[rst].abs().sum()) == 0
        assert bool((extras["time_outs"].bool() <= rst).all())
        assert bool((task.progress_buf < cfg.max_episode_length).all())
        # reset envs: root rows from the initial tensor, ball launched with the table row it consumed
        if rst.any():
            assert torch.equal(task.root_states[rst][:, :, 3:7], init_root[rst][:, :, 3:7])
            assert torch.equal(task.root_states[rst][:, b, 7:10], table_before[rst])
            for name, val in zip(cfg.flag_names, cfg.flag_reset_values):
                assert bool((getattr(task, name)[rst] == val).all()), name
        # consumed table rows were redrawn, the others kept
        changed = (task.st["reset_ball_vel"] != table_before).any(dim=1)
        assert torch.equal(changed, rst)
        # compacted index lists agree with the mask
        k = int(task.reset_count.item())
        assert k == int(rst.sum())
        envs = torch.div(task.reset_actor_indices[:k * cfg.num_actors].view(k, cfg.num_actors)[:, 0], cfg.num_actors,
                         rounding_mode="floor")
        assert torch.equal(envs.sort().values, rst.nonzero().flatten().to(envs.dtype))
        # outputs are finite, clamped obs respect the clip, actions were clamped in place
        assert torch.isfinite(task.obs_buf).all() and torch.isfinite(rew).all()
        assert float(obs["obs"].abs().max()) <= task.clip_obs
        assert float(task.actions.abs().max()) <= task.clip_actions
    assert total_resets > 0
    m = task.stats.means(n)
    assert m["reset_count"] >= 0.0 and abs(m["progress_sum"]) < cfg.max_episode_length
