"""CPU tests that EXECUTE the reference's own class-method bodies (`_reset_idx`, `pre_physics_step`,
`post_physics_step`, `compute_reward`, `compute_observations`: oracle/ref_methods.py cuts them out of
/root/reference unmodified and binds them to a stub `self`) and pin the restated wrappers of
oracle/task_oracle.py / oracle/pingpong_oracle.py against them, bit for bit.  Build container only
(needs /root/reference); on the GPU box the golden step fixtures (tests/golden/*_step.npz, frozen from
these same method bodies by oracle/make_golden.py) carry the result."""
import random

import pytest
import torch

from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import clone_state, make_state
from oracle import pingpong_oracle as O
from oracle import ref_extract as R
from oracle import ref_methods as M
from oracle import task_oracle

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference is only present in the build container")

STEP_VARIANTS = M.STEP_VARIANTS                                   # classes whose post_physics_step runs as shipped
RESET_VARIANTS = STEP_VARIANTS + ("a4",)


def _state_keys(cfg):
    return ("root_states", "dof_states", "progress_buf", "reset_buf", "rew_buf", "obs_buf") + cfg.flag_names + cfg.counter_names


@pytest.mark.parametrize("variant", [v for v in RESET_VARIANTS])
def test_sampler_ranges_are_the_constructor_s(variant):
    got = M.read_ranges_from_reference(variant)
    for key, val in M._RANGES[variant].items():
        assert got[key] == val, (variant, key, got.get(key), val)


@pytest.mark.parametrize("variant", RESET_VARIANTS)
def test_reference_reset_idx_body(variant):
    """`_reset_idx` executed from the reference text == `pingpong_oracle.reset_idx` fed with the launch values the
    restated sampler draws from the same `random` stream (pins the sampler formula, the draw order, the row
    rewrites, the int64 -> int32 index gather and the flag resets)."""
    cfg = CONFIGS[variant]
    n = 96
    st = make_state(cfg, n, seed=300 + cfg.variant_id)
    ref_st, our_st = clone_state(st), clone_state(st)
    env_ids = torch.tensor([0, 5, 17, 18, 64, 95], dtype=torch.int64)
    task = M.make_ref_task(cfg, ref_st)
    random.seed(777)
    with R.quiet():
        task._reset_idx(env_ids)
    vel, yz = M.replay_launch_draws(variant, 777, len(env_ids))
    a_idx, d_idx = O.reset_idx(variant, our_st, env_ids, vel, yz)
    # the A4 class owns one flag set (A4:908-911); the second humanoid's set exists only because its reward, which the
    # class never calls (defect D4), is run here: the restatement resets it like the first, the reference cannot
    ref_flags = cfg.flag_names[:3] if variant == "a4" else cfg.flag_names
    for key in ("root_states", "dof_states", "progress_buf") + ref_flags:
        assert torch.equal(ref_st[key], our_st[key]), (variant, key)
    if variant == "a4":
        for key in cfg.flag_names[3:]:
            assert torch.equal(ref_st[key], st[key]), key       # untouched by the reference
            assert bool((our_st[key][env_ids] == cfg.flag_reset_values[cfg.flag_names.index(key)]).all())
    got = task.gym.last("set_actor_root_state_tensor_indexed")
    assert got[2].dtype == torch.int32 and torch.equal(got[2], a_idx) and got[3] == len(env_ids) * cfg.num_actors
    dof_call = task.gym.last("set_dof_state_tensor_indexed")
    if variant == "nes":
        assert dof_call is None                                  # NES:871-918: the DOF reset is commented out
    else:
        assert dof_call[2].dtype == torch.int32 and torch.equal(dof_call[2], d_idx)


def test_a4_pre_physics_step_cannot_broadcast_as_shipped():
    """Defect D6: offset[7] + scale[7] * actions[N,14] (A4:1014).  The restated wrapper tiles offset / scale."""
    cfg = CONFIGS["a4"]
    st = make_state(cfg, 8, seed=1)
    task = M.make_ref_task(cfg, clone_state(st), overrides={"_pd_action_offset": st["pd_action_offset"][:7],
                                                            "_pd_action_scale": st["pd_action_scale"][:7]})
    with pytest.raises(RuntimeError):
        task.pre_physics_step(st["actions"])


@pytest.mark.parametrize("variant", STEP_VARIANTS)
def test_reference_pre_physics_step_body(variant):
    cfg = CONFIGS[variant]
    st = make_state(cfg, 64, seed=11)
    task = M.make_ref_task(cfg, clone_state(st))
    task.pre_physics_step(st["actions"])
    pd = task.gym.last("set_dof_position_target_tensor")[1]
    ours = clone_state(st)
    task_oracle.pre_physics_step(cfg, ours)
    assert torch.equal(pd, ours["pd_targets"])
    assert torch.equal(task.pre_ball2_root_states, ours["pre_ball_states"])
    assert task.actions.data_ptr() != st["actions"].data_ptr() and torch.equal(task.actions, st["actions"])


_bind = M.bind_step_functions


@pytest.mark.parametrize("variant", STEP_VARIANTS)
@pytest.mark.parametrize("seed", [1, 2])
def test_reference_post_physics_step_body(variant, seed):
    """The whole `post_physics_step` of the reference class (progress += 1 -> compute_reward -> nonzero ->
    reset_idx -> compute_observations, ADOF's counter clear included), its own text calling its own free
    functions, against `task_oracle.post_physics_step`: every buffer the step writes, bit for bit."""
    cfg = CONFIGS[variant]
    n = 128
    st = make_state(cfg, n, seed=40 * seed + cfg.variant_id)
    ref_st, our_st = clone_state(st), clone_state(st)
    _bind(variant)
    task = M.make_ref_task(cfg, ref_st)
    random.seed(1234 + seed)
    with R.quiet():
        task.post_physics_step()
    env_ids = ref_st["reset_buf"].nonzero(as_tuple=False).flatten()
    assert len(env_ids) > 0, "the batch must exercise the reset"
    vel, yz = M.replay_launch_draws(variant, 1234 + seed, len(env_ids))
    our_st["reset_ball_vel"][env_ids] = vel
    if yz is not None:
        our_st["reset_ball_pos_yz"][env_ids] = yz
    task_oracle.post_physics_step(cfg, our_st)
    for key in _state_keys(cfg):
        assert torch.equal(ref_st[key], our_st[key]), (variant, key, int((ref_st[key] != our_st[key]).sum()))
    assert task.num_steps == 1


def test_a4_compute_reward_calls_an_undefined_function():
    """Defect D4: A4:743 calls `compute_pingpong_reward_nv`, which the file never defines."""
    cfg = CONFIGS["a4"]
    task = M.make_ref_task(cfg, clone_state(make_state(cfg, 8, seed=1)))
    with pytest.raises(NameError):
        task.compute_reward(task.actions)
