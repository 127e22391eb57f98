"""TEST INFRASTRUCTURE ONLY -- CPU oracle (torch fp32 restatement) of the humanoid
ping-pong task hot path of mjmj531/isaacgym.

Every function restates, operation by operation and in the reference's own
floating-point evaluation order, one live function of the reference; the
docstrings cite the reference file:line (aliases as in SURVEY.md: BASE, A3,
TILT, NES, A4, ALIGN, ADOF).  `tests/test_oracle_vs_reference.py` pins each of
them bit-for-bit against the reference's own source executed in the build
container, and `tests/golden/` holds outputs of the reference functions.

The product (`isaacgym_b200/`) never imports this file.
"""
import math

import torch

from .jit_utils_restated import calc_heading_quat_inv, my_quat_rotate

# --------------------------------------------------------------------------
# observations
# --------------------------------------------------------------------------


def _heading_frame(body_states, body_ids):
    """root position and inverse heading quaternion of body `body_ids[0]`
    (A3:1409-1415 / A3:1433-1442)."""
    root = body_states[:, body_ids[0], :]
    return root[:, 0:3], calc_heading_quat_inv(root[:, 3:7])


def humanoid_observations(body_states, dof_pos, dof_vel, body_ids):
    """A3:1429-1468 (byte-identical at TILT:1669, NES:1777, A4:1589, ALIGN:1607,
    ADOF:1849): [R(hq)(pos_j - pos_root), R(hq) vel_j, dof_pos, 0.1*dof_vel]."""
    n, j = body_states.shape[0], len(body_ids)
    pos = body_states[:, body_ids, 0:3]
    vel = body_states[:, body_ids, 7:10]
    root_pos, hq = _heading_frame(body_states, body_ids)
    hq_flat = hq.unsqueeze(-2).repeat((1, j, 1)).reshape(n * j, 4)
    local_pos = my_quat_rotate(hq_flat, (pos - root_pos.unsqueeze(1)).view(n * j, -1)).reshape(n, -1)
    local_vel = my_quat_rotate(hq_flat, vel.view(n * j, -1)).reshape(n, -1)
    return torch.cat([local_pos, local_vel, dof_pos.clone(), dof_vel.clone() * 0.1], dim=-1)


def pingpong_observations(body_states, body_ids, ball_root_states, y_intersect=False):
    """A3:1400-1426 (identical in TILT/NES/A4/ALIGN); with `y_intersect` the ADOF
    form ADOF:1811-1846, which appends ly + (lvy / (-lvx + 1e-6)) * lx."""
    root_pos, hq = _heading_frame(body_states, body_ids)
    lpos = my_quat_rotate(hq, ball_root_states[..., 0:3] - root_pos)
    lvel = my_quat_rotate(hq, ball_root_states[..., 7:10])
    if not y_intersect:
        return torch.cat((lpos, lvel), dim=-1)
    yi = lpos[..., 1] + (lvel[..., 1] / (-lvel[..., 0] + 1e-6)) * (lpos[..., 0])
    return torch.cat((lpos, lvel, yi.unsqueeze(-1)), dim=-1)


def imitation_observations(body_states, ref_body_states, ref_dof_pos, ref_dof_vel, balance_ids):
    """ADOF:1891-1927 (time_steps == 1): [10*R(hq)(ref_pos - pos), R(hq)(ref_vel - vel),
    ref_dof_pos, ref_dof_vel]; heading from body `balance_ids[0]`."""
    n, j = body_states.shape[0], len(balance_ids)
    pos = body_states[:, balance_ids, 0:3]
    vel = body_states[:, balance_ids, 7:10]
    rpos = ref_body_states[:, balance_ids, 0:3]
    rvel = ref_body_states[:, balance_ids, 7:10]
    hq = calc_heading_quat_inv(body_states[:, balance_ids, 3:7][:, 0])
    hq_flat = hq.unsqueeze(-2).repeat((1, j, 1)).view(-1, 4)
    dpos = rpos.view(n, 1, j, 3) - pos.view(n, 1, j, 3)
    dvel = rvel.view(n, 1, j, 3) - vel.view(n, 1, j, 3)
    lpos = my_quat_rotate(hq_flat, dpos.view(-1, 3))
    lvel = my_quat_rotate(hq_flat, dvel.view(-1, 3))
    return torch.cat([lpos.view(n, -1) * 10., lvel.view(n, -1), ref_dof_pos, ref_dof_vel], dim=-1).view(n, -1)


def base_observations(paddle1, paddle2, ball1, ball2):
    """BASE:776-813: plain concat of pos/vel of paddle1, paddle2, ball1, ball2."""
    pv = lambda s: torch.cat((s[..., 0:3], s[..., 7:10]), dim=-1)
    return torch.cat((pv(paddle1), pv(paddle2), pv(ball1), pv(ball2)), dim=-1)


# --------------------------------------------------------------------------
# rewards
# --------------------------------------------------------------------------


# Tests set TERM_SCALE to a list: every reward function then appends, per env, the sum of the magnitudes of the
# terms that entered that env's reward -- the scale the fp32 rounding of the sum lives on (per-env tolerance).
TERM_SCALE = None


def _note_terms(*terms):
    if TERM_SCALE is not None:
        TERM_SCALE.append(sum(torch.as_tensor(t).abs().double() for t in terms))


def _dist3(a, b):
    return torch.sqrt((a[..., 0] - b[..., 0]) ** 2 + (a[..., 1] - b[..., 1]) ** 2 + (a[..., 2] - b[..., 2]) ** 2)


def _power_reward(dof_force, dof_vel, power_coefficient):
    """TILT:1246-1247."""
    return -power_coefficient * torch.abs(torch.multiply(dof_force, dof_vel)).sum(dim=-1)


def _reset_mask(ball_z, threshold, progress_buf, reset_buf, max_episode_length, early_stop=True, extra_die=None):
    """TILT:1253-1265: die on ball height, time-out at L-1 (int64 output)."""
    ones = torch.ones_like(reset_buf)
    die = torch.zeros_like(reset_buf)
    if extra_die is not None:
        die = torch.where(extra_die, ones, die)
    if early_stop:
        die = torch.where(ball_z < threshold, ones, die)
    return torch.where(progress_buf >= max_episode_length - 1, ones, die)


def base_reward(paddle1, paddle2, ball1, ball2, reset_buf, progress_buf, max_episode_length):
    """BASE:622-667.  `die` tests ball1 twice (BASE:662, kept: defect D3)."""
    d1 = _dist3(paddle1[..., 0:3], ball2[..., 0:3])
    d2 = _dist3(paddle2[..., 0:3], ball1[..., 0:3])
    reward = 1.0 / (1.0 + d1 * d1) + 1.0 / (1.0 + d2 * d2)
    _note_terms(reward)
    z1 = ball1[..., 2]
    ones = torch.ones_like(reset_buf)
    die = torch.where((z1 < 0.1) & (z1 < 0.1), ones, torch.zeros_like(reset_buf))
    return reward, torch.where(progress_buf >= max_episode_length - 1, ones, die)


def a3_reward(humanoid_root, paddle, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
              max_episode_length, alpha, power_coefficient, penalty):
    """A3:1080-1173.  The per-env Python loop A3:1126-1128 is the masked
    assignment alpha*|vx| where pre_vx < 0 and vx > 0."""
    ppos, bpos = paddle[..., 0:3], ball[..., 0:3]
    dist = _dist3(ppos, bpos)
    pos_reward = 1.0 / (1.0 + 1.5 * dist * dist)
    vx = ball[..., 7]
    hit = (pre_ball[..., 7] < 0) & (vx > 0)
    vel_reward = torch.where(hit, alpha * torch.abs(vx), torch.zeros_like(vx))
    reward = pos_reward + _power_reward(dof_force, dof_vel, power_coefficient) + vel_reward
    missed = bpos[..., 0] < ppos[..., 0] - 1e-3
    _note_terms(pos_reward, _power_reward(dof_force, dof_vel, power_coefficient), vel_reward, missed * float(penalty))
    reward = torch.where(missed, reward + penalty, reward)
    return reward, _reset_mask(bpos[..., 2], 0.1, progress_buf, reset_buf, max_episode_length, extra_die=missed)


def tilt_reward(humanoid_root, paddle, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
                max_episode_length, alpha, power_coefficient, penalty, condition_calculated,
                hit_table_reward, not_hit_table_penalty, reward_calculated, no_bounce_before_half_mask,
                mirrored=False, scripted=False):
    """TILT:1105-1270 (`compute_pingpong_reward_nv`; A4:1113-1278 is the same
    text).  `mirrored=True` gives the far-side player A4:1280-1439.
    In the eager TILT function `flag |= x` mutates the caller's three bool flag
    tensors IN PLACE.  Under `@torch.jit.script` (A4:1113,1280; `scripted=True`)
    the same statement compiles to the out-of-place `aten::__or__`/`__and__`:
    the updates are local to the call and the caller's tensors never change
    (observed by running the reference; defect D16 in DESIGN.md)."""
    if scripted:
        condition_calculated = condition_calculated.clone()
        reward_calculated = reward_calculated.clone()
        no_bounce_before_half_mask = no_bounce_before_half_mask.clone()
    bpos = ball[..., 0:3]
    bx, by, bz = bpos[:, 0], bpos[:, 1], bpos[:, 2]
    dist = _dist3(paddle[..., 0:3], bpos)
    pos_reward = 1.0 / (1.0 + 1.5 * dist * dist)
    pre_vx, vx = pre_ball[..., 7], ball[..., 7]
    zero = torch.zeros_like(vx)
    if not mirrored:
        outgoing = vx > 0                                   # ball flying away from this player
        condition = (pre_vx < 0) & outgoing                 # TILT:1153
        missed = bx < humanoid_root[..., 0] - 0.05          # TILT:1169
        near_half, far_table, beyond = bx < 2.44, (bx > 2.44) & (bx < 3.1), bx >= 3.1
    else:
        outgoing = vx < 0
        condition = (pre_vx > 0) & outgoing                 # A4:1328
        missed = bx > humanoid_root[..., 0] + 0.05          # A4:1344
        near_half, far_table, beyond = bx > 1.06, (bx < 1.06) & (bx > 0.4), bx <= 0.4
    vel_reward = torch.where(condition & ~condition_calculated, alpha * torch.abs(vx), zero)
    condition_calculated |= condition
    reward = torch.where(missed, zero + penalty, zero)
    bounce_up = (bz < 0.83) & outgoing & (by < 0.6) & (by > -0.6)      # TILT:1184
    # stage A: bounced on own half -> penalty, once (TILT:1187-1196)
    stage_a = near_half & bounce_up
    hit_reward = torch.where(stage_a & ~reward_calculated, not_hit_table_penalty, zero)
    reward_calculated |= stage_a
    no_bounce_before_half_mask &= ~stage_a
    # stage B: first bounce on the far table (TILT:1199-1206)
    stage_b = far_table & bounce_up & no_bounce_before_half_mask
    hit_reward = torch.where(stage_b & ~reward_calculated, hit_table_reward, hit_reward)
    reward_calculated |= stage_b
    # stage C: flew past the table (TILT:1209-1214)
    hit_reward = torch.where(beyond & outgoing & ~reward_calculated, not_hit_table_penalty, hit_reward)
    reward_calculated |= beyond
    over_net = (bx > 1.7) & (bx < 1.8) & outgoing & (by < 0.4) & (by > -0.4) & (bz > 0.98) & (bz < 1.14)
    net_reward = torch.where(over_net, 400, zero)                        # TILT:1226-1244
    _note_terms(reward, pos_reward, _power_reward(dof_force, dof_vel, power_coefficient), vel_reward, hit_reward, net_reward)
    reward += pos_reward + _power_reward(dof_force, dof_vel, power_coefficient) + vel_reward + hit_reward + net_reward
    return reward, _reset_mask(bz, 0.1, progress_buf, reset_buf, max_episode_length)


def nes_reward(humanoid_root, paddle, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
               max_episode_length, alpha, power_coefficient, penalty, paddle_condition_calculated,
               missed_ball_calculated):
    """NES:1115-1322 (`compute_pingpong_reward_only_paddle`); no early stop."""
    ppos, bpos = paddle[..., 0:3], ball[..., 0:3]
    pre_vx, vx = pre_ball[..., 7], ball[..., 7]
    zero = torch.zeros_like(vx)
    hit = (pre_vx < 0) & (vx > 1.0)                                      # NES:1154
    bx, px, hx = bpos[..., 0], ppos[..., 0], humanoid_root[..., 0]
    missed = (bx < hx - 0.05) | (bx < px - 0.1)                          # NES:1163
    reward = torch.where((~missed_ball_calculated) & missed, zero + penalty, zero)
    missed_ball_calculated |= missed
    dist = torch.sqrt((ppos[..., 1] - bpos[..., 1]) ** 2 + (ppos[..., 2] - bpos[..., 2]) ** 2)
    pos_reward = torch.where((~paddle_condition_calculated) | (bx < hx - 0.05),
                             1.0 * torch.exp(-20.0 * dist * dist), zero)  # NES:1188-1195
    vel_reward = torch.where(hit & ~paddle_condition_calculated, alpha * torch.abs(vx), zero)
    paddle_condition_calculated |= hit
    _note_terms(reward, pos_reward, _power_reward(dof_force, dof_vel, power_coefficient), vel_reward, (bpos[..., 2] < 0.1) * 800.0)
    reward += pos_reward + _power_reward(dof_force, dof_vel, power_coefficient) + vel_reward
    reward = torch.where(bpos[..., 2] < 0.1, -800 + reward, reward)       # NES:1313-1315
    return reward, _reset_mask(bpos[..., 2], 0.1, progress_buf, reset_buf, max_episode_length, early_stop=False)


def align_reward(humanoid_root, paddle, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
                 max_episode_length, alpha, power_coefficient, penalty, hit_table_reward,
                 not_hit_table_penalty, reward_calculated):
    """ALIGN:1097-1230 (definition #1, the one that compiles).  The hit-table award
    condition is unsatisfiable as written (ALIGN:1164,1170,1174; defect D8).
    The function is TorchScript: `reward_calculated |= ...` is out-of-place there,
    so the caller's flag tensor is read but never written (defect D16)."""
    reward_calculated = reward_calculated.clone()
    bpos = ball[..., 0:3]
    bx = bpos[:, 0]
    dist = _dist3(paddle[..., 0:3], bpos)
    pos_reward = 1.0 / (1.0 + 1.5 * dist * dist)
    pre_vx, pre_vz = pre_ball[..., 7], pre_ball[..., 9]
    vx, vz = ball[..., 7], ball[..., 9]
    zero = torch.zeros_like(vx)
    condition = (pre_vx < 0) & (vx > 0)
    vel_reward = torch.where(condition, alpha * torch.abs(vx), zero)     # masked assignment ALIGN:1153
    in_table = (bx > 2.2) & (bx < 3.1)
    bounce_up = (pre_vz < 0) & (vz > 0)                                  # ALIGN:1167
    no_bounce_before_half = (bx < 2.2) & ~bounce_up                      # ALIGN:1170
    award = in_table & bounce_up & no_bounce_before_half
    hit_reward = torch.where(award & ~reward_calculated, hit_table_reward, zero)
    reward_calculated |= award
    hit_reward = torch.where((bx >= 3.1) & (vx > 0) & ~reward_calculated, not_hit_table_penalty, hit_reward)
    reward_calculated |= bx >= 3.1
    reward = pos_reward + _power_reward(dof_force, dof_vel, power_coefficient) + vel_reward + hit_reward
    missed = bx < humanoid_root[..., 0] - 0.05
    _note_terms(pos_reward, _power_reward(dof_force, dof_vel, power_coefficient), vel_reward, hit_reward, missed * float(penalty))
    reward = torch.where(missed, reward + penalty, reward)
    return reward, _reset_mask(bpos[..., 2], 0.1, progress_buf, reset_buf, max_episode_length)


def align2_reward(h1_root, paddle1, h2_root, paddle2, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
                  max_episode_length, alpha, power_coefficient, penalty, hit_table_reward, not_hit_table_penalty,
                  reward_calculated, last_hitter):
    """ALIGN:1233-1351, the two-humanoid definition (never live in the reference: defect D7; it loads
    once its return annotation is repaired, `ref_extract.load_align_two_humanoid_reward`).  TorchScript:
    `reward_calculated |= ...` stays local to the call (D16); `last_hitter` is returned."""
    reward_calculated = reward_calculated.clone()
    bpos = ball[..., 0:3]
    bx = bpos[:, 0]
    pre_vx, pre_vz, vx, vz = pre_ball[..., 7], pre_ball[..., 9], ball[..., 7], ball[..., 9]
    zero = torch.zeros_like(vx)
    d1, d2 = _dist3(paddle1[..., 0:3], bpos), _dist3(paddle2[..., 0:3], bpos)
    pos1, pos2 = 1.0 / (1.0 + 1.5 * d1 * d1), 1.0 / (1.0 + 1.5 * d2 * d2)
    c1 = (pre_vx < 0) & (vx > 0)                              # humanoid 1 hit the ball
    c2 = (pre_vx > 0) & (vx < 0)                              # humanoid 2 hit the ball
    vel1 = torch.where(c1, alpha * torch.abs(vx), zero)
    vel2 = torch.where(c2, alpha * torch.abs(vx), zero)
    range1 = (bx > 2.2) & (bx < 3.1)                          # humanoid 2's half of the table
    range2 = (bx < 1.3) & (bx > 0.4)                          # humanoid 1's half
    bounce_up = (pre_vz < 0) & (vz > 0)
    hit1 = torch.where(range1 & bounce_up & (last_hitter == 1) & ~reward_calculated, hit_table_reward, zero)
    hit2 = torch.where(range2 & bounce_up & (last_hitter == 2) & ~reward_calculated, hit_table_reward, zero)
    reward_calculated |= (range1 & bounce_up) | (range2 & bounce_up)
    hit1 = torch.where((bx >= 3.1) & (last_hitter == 1) & ~reward_calculated, not_hit_table_penalty, hit1)
    hit2 = torch.where((bx <= -3.1) & (last_hitter == 2) & ~reward_calculated, not_hit_table_penalty, hit2)
    power_reward = _power_reward(dof_force, dof_vel, power_coefficient)
    r1 = pos1 + power_reward + vel1 + hit1
    r2 = pos2 + power_reward + vel2 + hit2
    _note_terms(pos1, power_reward, vel1, hit1, (bx < h1_root[..., 0] - 0.05) * float(penalty))
    _note_terms(pos2, power_reward, vel2, hit2, (bx > h2_root[..., 0] + 0.05) * float(penalty))
    r1 = torch.where(bx < h1_root[..., 0] - 0.05, r1 + penalty, r1)
    r2 = torch.where(bx > h2_root[..., 0] + 0.05, r2 + penalty, r2)
    reset = _reset_mask(bpos[..., 2], 0.1, progress_buf, reset_buf, max_episode_length)
    last_hitter = torch.where(c1, torch.ones_like(last_hitter), last_hitter)
    last_hitter = torch.where(c2, torch.full_like(last_hitter, 2), last_hitter)
    return r1, r2, reset, last_hitter


def adof_gradient_penalty(ball_pos, vx, hit_table_reward, not_hit_table_penalty, hit_table_calculated,
                          hit_table_count, humanoid_die_calculated):
    """ADOF:1245-1301."""
    x, y, z = ball_pos[..., 0], ball_pos[..., 1], ball_pos[..., 2]
    z_in_range = (z >= 0.82) & (z <= 0.83) & (vx > 0)
    distance = torch.sqrt((x - 2.5) ** 2 + (y - 0.0) ** 2)
    in_range = (x >= 1.9) & (x <= 3.1) & (y >= -0.6) & (y <= 0.6)
    hit_table_count = torch.where(z_in_range & in_range, torch.ones_like(hit_table_count), hit_table_count)
    out = torch.where(z_in_range & (~hit_table_calculated) & (~humanoid_die_calculated),
                      torch.where(in_range, hit_table_reward, not_hit_table_penalty * distance),
                      torch.zeros_like(distance))
    hit_table_calculated |= z_in_range
    return out, hit_table_calculated, hit_table_count


def adof_imitation_reward(dof_pos, rest_dof_pos, dof_vel, rest_dof_vel, body_states, initial_body_states,
                          balance_ids, is_train=True):
    """ADOF:1313-1418, `is_g1=True` branch (every shipped config sets is_g1: true)."""
    k_pos, k_vel, k_dof_pos, k_dof_vel = 50., 4.0, 5.0, 0.05
    w_pos, w_vel, w_dof_pos, w_dof_vel = 0.4, 0.2, 0.2, 0.2
    pos = body_states[:, balance_ids, 0:3]
    vel = body_states[:, balance_ids, 7:10]
    rpos = initial_body_states[:, balance_ids, 0:3]
    rvel = initial_body_states[:, balance_ids, 7:10]
    r_body_pos = torch.exp(-k_pos * ((rpos - pos) ** 2).mean(dim=-1).mean(dim=-1))
    r_body_vel = torch.exp(-k_vel * ((rvel - vel) ** 2).mean(dim=-1).mean(dim=-1))
    dq2 = (rest_dof_pos - dof_pos) ** 2
    first22 = (w_dof_pos * 50.0) * torch.exp(-(k_dof_pos * 500.0) * dq2[..., :22].mean(dim=-1))
    last5 = w_dof_pos * torch.exp(-k_dof_pos * dq2[..., 22:].mean(dim=-1))
    dqd = ((rest_dof_vel[..., :22] - dof_vel[..., :22]) ** 2).mean(dim=-1)
    r_dof_vel = torch.exp(-k_dof_vel * dqd)
    ref_reward = first22 + last5 + w_dof_vel * r_dof_vel + w_pos * r_body_pos + w_vel * r_body_vel
    termination_distance = 0.32 if is_train else 1e6
    has_fallen = torch.any(torch.norm(pos - rpos, dim=-1).mean(dim=-1, keepdim=True) > termination_distance, dim=-1)
    ref_reward = torch.where(has_fallen, torch.ones_like(ref_reward) * -50.0, ref_reward)
    return ref_reward, has_fallen


def adof_reward(humanoid_root, pelvis, paddle, pre_ball, ball, dof_force, dof_vel, reset_buf, progress_buf,
                max_episode_length, alpha, power_coefficient, paddle_condition_calculated, hit_paddle_reward,
                miss_paddle_penalty_coefficient, cross_net_reward_float, hit_table_reward, not_hit_table_penalty,
                hit_table_calculated, die_penalty_float, die_penalty_calculated, humanoid_die_calculated,
                closer_to_paddle_count, hit_paddle_count, cross_net_count, hit_table_count, fall_down_count,
                init_dof_pos, dof_pos, init_dof_vel, body_states, initial_body_states, balance_ids, is_train=True):
    """ADOF:1440-1690.  Returns the reference's 11-tuple with every flag/counter as
    bool (the wrapper ADOF:802 assigns them into bool tensors; defect D10)."""
    ppos, bpos = paddle[..., 0:3], ball[..., 0:3]
    pre_vx, vx = pre_ball[..., 7], ball[..., 7]
    zero = torch.zeros_like(vx)
    ref_reward, has_fallen = adof_imitation_reward(dof_pos, init_dof_pos, dof_vel, init_dof_vel, body_states,
                                                   initial_body_states, balance_ids, is_train)
    pelvis_height = pelvis[..., 2]
    fall_down_count = torch.where(has_fallen > 0, torch.ones_like(pelvis_height), fall_down_count)
    bx, by, bz = bpos[..., 0], bpos[..., 1], bpos[..., 2]
    px, py, pz = ppos[..., 0], ppos[..., 1], ppos[..., 2]
    x_close = torch.abs(bx - px) < 0.2
    first_time_close = x_close & ~paddle_condition_calculated
    yz = torch.sqrt((by - py) ** 2 + (bz - pz) ** 2)
    in_circle = yz < 0.15
    pos_reward = torch.where(first_time_close & ~humanoid_die_calculated,
                             torch.where(in_circle, hit_paddle_reward, miss_paddle_penalty_coefficient * yz), zero)
    closer_to_paddle_count = torch.where(first_time_close & in_circle, torch.ones_like(closer_to_paddle_count),
                                         closer_to_paddle_count)
    hit = (pre_vx < 0) & (vx > 1.5)
    hit_paddle_count = torch.where(hit, torch.ones_like(hit_paddle_count), hit_paddle_count)
    vel_reward = torch.where(hit & ~paddle_condition_calculated & ~humanoid_die_calculated,
                             alpha * torch.abs(vx), zero)
    paddle_condition_calculated |= x_close
    time_penalty = torch.where((bx > humanoid_root[..., 0]) & (vx < 0), -0.01 * progress_buf.float(),
                               torch.zeros_like(progress_buf))
    table, hit_table_calculated, hit_table_count = adof_gradient_penalty(
        bpos, vx, hit_table_reward, not_hit_table_penalty, hit_table_calculated, hit_table_count,
        humanoid_die_calculated)
    when_over_net = (bx > 1.72) & (bx < 1.78) & (vx > 0)
    suitable = (bz > 0.96) & (bz < 1.25)
    over_height = torch.where(~suitable, torch.where(bz > 1.25, bz - 1.25, 0.96 - bz), zero)
    net = torch.where(when_over_net & ~humanoid_die_calculated,
                      torch.where(suitable, cross_net_reward_float, -400 * over_height), zero)
    cross_net_count = torch.where(net > 0, torch.ones_like(cross_net_count), cross_net_count)
    power_reward = _power_reward(dof_force, dof_vel, power_coefficient)
    low = bz < 0.78
    die_penalty = torch.where(low & ~die_penalty_calculated & ~humanoid_die_calculated, die_penalty_float, zero)
    die_penalty_calculated |= low
    humanoid_die_calculated |= (pelvis_height < 0.97)
    _note_terms(pos_reward, power_reward, vel_reward, table, net, die_penalty, time_penalty, ref_reward)
    reward = zero + (pos_reward + power_reward + vel_reward + table + net + die_penalty + time_penalty + ref_reward)
    reset = _reset_mask(bz, 0.78, progress_buf, reset_buf, max_episode_length, early_stop=False)
    return (reward, reset, paddle_condition_calculated, hit_table_calculated, die_penalty_calculated,
            humanoid_die_calculated, closer_to_paddle_count.bool(), hit_paddle_count.bool(), cross_net_count.bool(),
            hit_table_count.bool(), fall_down_count.bool())


# --------------------------------------------------------------------------
# pre-step, reset, ball sampling
# --------------------------------------------------------------------------


def pd_targets(pd_action_offset, pd_action_scale, actions):
    """TILT:1006 (hash-identical in A3/NES/ALIGN/ADOF; A4 with tiled offset/scale, D6)."""
    return pd_action_offset + pd_action_scale * actions


def sample_ball_velocity(rng, variant):
    """Host-side ball launch velocity, one `random.Random` stream like the reference
    (TILT:307-318, NES:312-323, ADOF:357-367, A3:300-302, BASE:250-266)."""
    u = rng.uniform
    rad = math.radians
    if variant in ("tilt", "a4", "align", "align2"):
        s = -u(8.0, 8.8 if variant in ("align", "align2") else 8.6)
        a = u(-5.0, 5.0)
        z = u(2.0, 10.0)
        return (s * math.cos(rad(a)) * math.cos(rad(z)), s * math.sin(rad(a)) * math.sin(rad(z)), s * math.sin(rad(a)))
    if variant in ("nes", "adof"):
        s = u(5.4, 5.9) if variant == "nes" else u(5.0, 5.4)
        a = u(-5.0, 5.0) if variant == "nes" else u(-8.0, 3.0)
        z = u(10.0, 17.0) if variant == "nes" else u(14.0, 24.0)
        return (-s * math.cos(rad(a)) * math.cos(rad(z)), s * math.sin(rad(a)) * math.cos(rad(z)), s * math.sin(rad(z)))
    if variant == "a3":
        s = -u(6.5, 7.5)
        a = u(-5.0, 5.0)
        return (s * math.cos(rad(a)), s * math.sin(rad(a)), 0.0)
    raise ValueError(variant)


# per-variant static description of the task step (what the class wrappers hard-code)
VARIANTS = {
    # name: actors, ball actor row, flags (name, value after reset), resets dof, stats period
    "a3":    dict(actors=3, ball=2, flags=(), reset_dof=True, log_every=20),
    "tilt":  dict(actors=3, ball=2, flags=(("reward_calculated", False), ("condition_calculated", False),
                                           ("no_bounce_before_half_mask", True)), reset_dof=True, log_every=40),
    "nes":   dict(actors=3, ball=2, flags=(("paddle_condition_calculated", False), ("missed_ball_calculated", False)),
                  reset_dof=False, log_every=40),
    "align": dict(actors=3, ball=2, flags=(("reward_calculated", False),), reset_dof=True, log_every=40),
    "a4":    dict(actors=4, ball=3, flags=(("reward_calculated", False), ("condition_calculated", False),
                                           ("no_bounce_before_half_mask", True),
                                           ("reward_calculated_2", False), ("condition_calculated_2", False),
                                           ("no_bounce_before_half_mask_2", True)), reset_dof=True, log_every=40),
    # ALIGN def #2 has no class code of its own: reward_calculated is cleared like ALIGN:897, and a fresh
    # rally starts with last_hitter = 2, its documented initial value (ALIGN:1253)
    "align2": dict(actors=4, ball=3, flags=(("reward_calculated", False),), reset_dof=True, log_every=40,
                   states=(("last_hitter", 2),)),
    "adof":  dict(actors=3, ball=2, flags=(("paddle_condition_calculated", False), ("die_penalty_calculated", False),
                                           ("humanoid_die_calculated", False), ("hit_table_calculated", False)),
                  reset_dof=True, log_every=32),
}
ADOF_COUNTERS = ("closer_to_paddle_count", "hit_paddle_count", "cross_net_count", "hit_table_count", "fall_down_count")


def reset_idx(variant, st, env_ids, ball_vel, ball_pos_yz=None):
    """`_reset_idx` of the 3-/4-actor classes: TILT:847-906, A3:826-872, NES:871-918
    (no DOF reset), A4:853-912, ALIGN:845-898, ADOF:965-1028 (also ball y,z).
    `st` is a dict of tensors; `ball_vel[k,3]` are the host-sampled launch
    velocities for `env_ids` (the reference draws them with `random.uniform`).
    Returns the int32 actor / dof index lists handed to the gym setters."""
    v = VARIANTS[variant]
    root, init_root = st["root_states"], st["initial_root_states"]
    root[env_ids, :, 0:3] = init_root[env_ids, :, 0:3]
    root[env_ids, :, 3:7] = init_root[env_ids, :, 3:7]
    root[env_ids, :, 7:13] = torch.zeros_like(root[env_ids, :, 7:13])
    for k, e in enumerate(env_ids):
        if ball_pos_yz is not None:
            root[e, v["ball"], 1] = ball_pos_yz[k, 0]
            root[e, v["ball"], 2] = ball_pos_yz[k, 1]
        root[e, v["ball"], 7:10] = ball_vel[k]
    if v["reset_dof"]:
        st["dof_states"][env_ids, :, :] = st["initial_dof_states"][env_ids, :, :]
    n, a = root.shape[0], root.shape[1]
    actor_indices = st["actor_indices"].view(n, a)[env_ids].flatten().to(torch.int32)
    dof_per = st["dof_indices"].numel() // n
    dof_indices = st["dof_indices"].view(n, dof_per)[env_ids].flatten().to(torch.int32)
    st["progress_buf"][env_ids] = 0
    for name, val in v["flags"]:
        st[name][env_ids] = val
    for name, val in v.get("states", ()):
        st[name][env_ids] = val
    return actor_indices, dof_indices


def base_reset_idx(st, env_ids, ball1_vel, ball2_vel):
    """BASE:530-579: one velocity pair for all `env_ids`, velocities 10:13 and other
    actors' 7:13 are NOT zeroed, `reset_buf[env_ids] = 0`."""
    root, init_root = st["root_states"], st["initial_root_states"]
    root[env_ids, :, 0:3] = init_root[env_ids, :, 0:3]
    root[env_ids, :, 3:7] = init_root[env_ids, :, 3:7]
    root[env_ids, 3, 7:10] = ball1_vel
    root[env_ids, 4, 7:10] = ball2_vel
    st["dof_states"][env_ids, :, :] = st["initial_dof_states"][env_ids, :, :]
    st["progress_buf"][env_ids] = 0
    st["reset_buf"][env_ids] = 0
    return st["actor_indices"].view(root.shape[0], 5)[env_ids].flatten().to(torch.int32)
