"""TEST INFRASTRUCTURE ONLY.

Restatement of the quaternion helpers the reference imports with
`from isaacgymenvs.utils.torch_jit_utils import *`
(/root/reference/tasks/humanoid_interos_edit_pingpong_only_3_actor.py:38) and
calls at A3:1415,1419-1420,1442,1449-1450 and ADOF:1826,1830-1831,1862,
1869-1870,1905,1912-1913.

PARITY UNPINNED: `isaacgymenvs` is neither vendored in /root/reference nor
pinned by any requirements/lock file there.  The formulas
below restate the published algorithm of NVIDIA-Omniverse/IsaacGymEnvs
`isaacgymenvs/utils/torch_jit_utils.py` (1.x line): quaternions are xyzw, the
input quaternion is NOT normalised, the vector rotation is the
`v(2w^2-1) + 2w(q x v) + 2q(q.v)` form with the dot product taken by `bmm`.
Their pinning here is by property tests (tests/test_oracle_helpers.py).
"""
import torch


def normalize(x, eps: float = 1e-9):
    # x / max(||x||_2, eps) along the last axis
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def quat_unit(a):
    return normalize(a)


def quat_from_angle_axis(angle, axis):
    theta = (angle / 2).unsqueeze(-1)
    xyz = normalize(axis) * theta.sin()
    w = theta.cos()
    return quat_unit(torch.cat([xyz, w], dim=-1))


def my_quat_rotate(q, v):
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3),
                          v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c


def calc_heading(q):
    ref_dir = torch.zeros_like(q[..., 0:3])
    ref_dir[..., 0] = 1
    rot_dir = my_quat_rotate(q, ref_dir)
    return torch.atan2(rot_dir[..., 1], rot_dir[..., 0])


def calc_heading_quat_inv(q):
    heading = calc_heading(q)
    axis = torch.zeros_like(q[..., 0:3])
    axis[..., 2] = 1
    return quat_from_angle_axis(-heading, axis)


HELPER_NAMES = ("normalize", "quat_unit", "quat_from_angle_axis",
                "my_quat_rotate", "calc_heading", "calc_heading_quat_inv")
