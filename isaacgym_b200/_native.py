"""ctypes binding of libppk.so (include/ppk.h).  There is NO fallback: if the CUDA library is
missing or a call fails, the product path raises."""
import ctypes as C
import os

import torch

from .config import TaskConfig

PPK_MAX_BODY_IDS = 32
PPK_MAX_FLAGS = 12
PPK_STATS_SLOTS = 64
PPK_NUM_STATS = 8
ABI_VERSION = 2

PHASE_PROGRESS, PHASE_REWARD, PHASE_RESET, PHASE_OBS, PHASE_STATS = 1, 2, 4, 8, 16
PHASE_ALL = 31
PHASE_MOMENTS = 32          # opt-in: column moments of obs_buf for RunningMeanStd (learner side)
PPK_MOMENT_SLOTS = 64

STAT_NAMES = ("reward_sum", "progress_sum", "reset_count", "fall_down_count", "closer_to_paddle_count",
              "hit_paddle_count", "cross_net_count", "hit_table_count")

_f32p = C.POINTER(C.c_float)


class PpkTask(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("variant", C.c_int32), ("num_actors", C.c_int32), ("num_bodies", C.c_int32),
        ("num_dofs", C.c_int32), ("humanoid_actor", C.c_int32 * 2), ("ball_actor", C.c_int32),
        ("paddle_body", C.c_int32 * 2), ("pelvis_body", C.c_int32), ("num_body_ids", C.c_int32),
        ("body_ids", (C.c_int32 * PPK_MAX_BODY_IDS) * 2), ("num_balance_ids", C.c_int32),
        ("balance_ids", C.c_int32 * PPK_MAX_BODY_IDS), ("max_episode_length", C.c_int64),
        ("alpha", C.c_float), ("power_coefficient", C.c_float), ("penalty", C.c_float),
        ("hit_table_reward", C.c_float), ("not_hit_table_penalty", C.c_float), ("cross_net_reward", C.c_float),
        ("die_penalty", C.c_float), ("hit_paddle_reward", C.c_float), ("miss_paddle_penalty_coefficient", C.c_float),
        ("is_train", C.c_int32), ("reset_dof", C.c_int32), ("write_flags", C.c_int32),
    ]


class PpkBuffers(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_envs", C.c_int64),
        ("rigid_body_states", C.c_void_p), ("root_states", C.c_void_p), ("dof_states", C.c_void_p),
        ("dof_forces", C.c_void_p), ("pre_ball_states", C.c_void_p),
        ("pre_ball_stride", C.c_int32), ("pre_vx_offset", C.c_int32), ("pre_vz_offset", C.c_int32),
        ("initial_root_states", C.c_void_p), ("initial_dof_states", C.c_void_p), ("initial_body_states", C.c_void_p),
        ("reset_ball_vel", C.c_void_p), ("reset_ball_pos_yz", C.c_void_p),
        ("obs_buf", C.c_void_p), ("rew_buf", C.c_void_p), ("reset_buf", C.c_void_p), ("progress_buf", C.c_void_p),
        ("flags", C.c_void_p * PPK_MAX_FLAGS),
        ("actions", C.c_void_p), ("pd_action_offset", C.c_void_p), ("pd_action_scale", C.c_void_p),
        ("pd_targets", C.c_void_p), ("stats", C.c_void_p), ("scratch", C.c_void_p),
        ("root_states_out", C.c_void_p), ("dof_states_out", C.c_void_p),
        ("clip_actions", C.c_float), ("dof_indices_per_env", C.c_int32), ("timeout_buf", C.c_void_p),
        ("actor_indices", C.c_void_p), ("dof_indices", C.c_void_p), ("reset_count", C.c_void_p),
        ("reset_actor_indices", C.c_void_p), ("reset_dof_indices", C.c_void_p),
        ("last_hitter", C.c_void_p), ("initial_balance_states", C.c_void_p),
        ("clip_observations", C.c_float), ("obs_moments", C.c_void_p),
    ]


class PpkRunningMeanStd(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_int32), ("epsilon", C.c_float), ("clip_obs", C.c_float),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("count", C.c_void_p),
                ("moments", C.c_void_p)]


PPK_ACT_NONE, PPK_ACT_ELU = 0, 1

_LIB = None


def lib_path() -> str:
    # PPK_LIB: load another build of the same ABI (A/B measurements of kernel variants)
    return os.environ.get("PPK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libppk.so")


def load():
    """Load libppk.so.  Where the sources and nvcc are present (the build container, the GPU box) the library is
    rebuilt first whenever its recorded source hash differs from the sources -- a stale binary is never loaded
    silently; `build()` returns at once when the hash matches.  PPK_LIB loads another build as is.  Raises if
    unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.environ.get("PPK_LIB"):
        from . import build as _build
        if os.path.isdir(_build.CSRC) and (os.path.exists(_build.nvcc_path()) or not os.path.exists(path)):
            _build.build()
        elif os.path.isdir(_build.CSRC) and not _build.stamp_matches():
            raise RuntimeError(f"{path} was built from other sources than isaacgym_b200/csrc (hash mismatch) and nvcc "
                               "is not available to rebuild it: run `python -m isaacgym_b200.build` where nvcc is")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: the CUDA extension is the only implementation of the "
                           "ping-pong task step (no CPU fallback). Run `python -m isaacgym_b200.build`.")
    lib = C.CDLL(path)
    lib.ppk_abi_version.restype = C.c_int
    lib.ppk_strerror.restype = C.c_char_p
    lib.ppk_strerror.argtypes = [C.c_int]
    for name in ("ppk_compute_reward", "ppk_compute_observations", "ppk_pre_physics_step"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(PpkTask), C.POINTER(PpkBuffers), C.c_void_p]
    lib.ppk_post_physics_step.restype = C.c_int
    lib.ppk_post_physics_step.argtypes = [C.POINTER(PpkTask), C.POINTER(PpkBuffers), C.c_uint32, C.c_void_p]
    lib.ppk_reset_idx.restype = C.c_int
    lib.ppk_reset_idx.argtypes = [C.POINTER(PpkTask), C.POINTER(PpkBuffers), C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ppk_sample_ball_launch.restype = C.c_int
    lib.ppk_sample_ball_launch.argtypes = [C.POINTER(PpkTask), C.POINTER(PpkBuffers), C.c_uint64, C.c_uint64,
                                           C.c_int64, C.c_int32, C.c_void_p]
    lib.ppk_stats_reduce.restype = C.c_int
    lib.ppk_stats_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    if hasattr(lib, "ppk_host_session_create"):
        lib.ppk_host_session_create.restype = C.c_int
        lib.ppk_host_session_create.argtypes = [C.POINTER(PpkTask), C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
        lib.ppk_host_session_destroy.restype = C.c_int
        lib.ppk_host_session_destroy.argtypes = [C.c_void_p]
        lib.ppk_host_post_physics_step.restype = C.c_int
        lib.ppk_host_post_physics_step.argtypes = [C.c_void_p, C.POINTER(PpkBuffers), C.c_uint32]
        lib.ppk_host_session_traffic.restype = C.c_int
        lib.ppk_host_session_traffic.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    rp = C.POINTER(PpkRunningMeanStd)
    lib.ppk_rms_accumulate.restype = C.c_int
    lib.ppk_rms_accumulate.argtypes = [rp, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ppk_rms_merge.restype = C.c_int
    lib.ppk_rms_merge.argtypes = [rp, C.c_double, C.c_void_p]
    lib.ppk_rms_fold_step_moments.restype = C.c_int
    lib.ppk_rms_fold_step_moments.argtypes = [rp, C.c_void_p, C.c_double, C.c_int32, C.c_void_p]
    lib.ppk_rms_update.restype = C.c_int
    lib.ppk_rms_update.argtypes = [rp, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ppk_rms_normalize.restype = C.c_int
    lib.ppk_rms_normalize.argtypes = [rp, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.ppk_rms_scratch_doubles.restype = C.c_size_t
    lib.ppk_rms_scratch_doubles.argtypes = [C.c_int32]
    lib.ppk_linear_packed_bytes.restype = C.c_size_t
    lib.ppk_linear_packed_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.ppk_linear_pack.restype = C.c_int
    lib.ppk_linear_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ppk_policy_first_layer.restype = C.c_int
    lib.ppk_policy_first_layer.argtypes = [rp, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p]
    lib.ppk_linear_packed_bytes_f32.restype = C.c_size_t
    lib.ppk_linear_packed_bytes_f32.argtypes = [C.c_int32, C.c_int32]
    lib.ppk_linear_pack_f32.restype = C.c_int
    lib.ppk_linear_pack_f32.argtypes = lib.ppk_linear_pack.argtypes
    lib.ppk_policy_first_layer_f32.restype = C.c_int
    lib.ppk_policy_first_layer_f32.argtypes = lib.ppk_policy_first_layer.argtypes
    if lib.ppk_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libppk.so ABI {lib.ppk_abi_version()} != binding ABI {ABI_VERSION}")
    _LIB = lib
    return lib


def check(rc: int, what: str = "ppk"):
    if rc != 0:
        msg = load().ppk_strerror(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def make_task(cfg: TaskConfig) -> PpkTask:
    t = PpkTask()
    t.struct_size = C.sizeof(PpkTask)
    t.variant = cfg.variant_id
    t.num_actors, t.num_bodies, t.num_dofs = cfg.num_actors, cfg.num_bodies, cfg.num_dofs
    t.humanoid_actor[0], t.humanoid_actor[1] = cfg.humanoid_actor
    t.ball_actor = cfg.ball_actor
    t.paddle_body[0], t.paddle_body[1] = cfg.paddle_body
    t.pelvis_body = cfg.pelvis_body
    t.num_body_ids = len(cfg.body_ids)
    for j, b in enumerate(cfg.body_ids):
        t.body_ids[0][j] = b
    for j, b in enumerate(cfg.body_ids_2):
        t.body_ids[1][j] = b
    t.num_balance_ids = len(cfg.balance_ids)
    for j, b in enumerate(cfg.balance_ids):
        t.balance_ids[j] = b
    t.max_episode_length = int(cfg.max_episode_length)
    t.alpha, t.power_coefficient, t.penalty = cfg.alpha, cfg.power_coefficient, cfg.penalty
    t.hit_table_reward, t.not_hit_table_penalty = cfg.hit_table_reward, cfg.not_hit_table_penalty
    t.cross_net_reward, t.die_penalty = cfg.cross_net_reward, cfg.die_penalty
    t.hit_paddle_reward = cfg.hit_paddle_reward
    t.miss_paddle_penalty_coefficient = cfg.miss_paddle_penalty_coefficient
    t.is_train = 1 if cfg.is_train else 0
    t.reset_dof = 1 if cfg.reset_dof else 0
    # eager reward functions mutate the caller's flags; the TorchScript ones (ALIGN, A4) do not (D16)
    t.write_flags = 0 if cfg.variant in ("align", "a4", "align2") else 1
    return t


def _ptr(t, dtype, what, host=False):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a tensor")
    if t.dtype != dtype:
        raise TypeError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{what}: must be contiguous")
    if host != (t.device.type == "cpu"):
        raise ValueError(f"{what}: expected a {'host' if host else 'CUDA'} tensor, got {t.device}")
    return t.data_ptr()


def make_buffers(cfg: TaskConfig, st: dict, host: bool = False) -> PpkBuffers:
    """Fill a PpkBuffers from a dict of tensors named as in `synth.make_state` / the task shim."""
    b = PpkBuffers()
    b.struct_size = C.sizeof(PpkBuffers)
    n = st["progress_buf"].shape[0]
    b.num_envs = n
    f32, i64, u8 = torch.float32, torch.int64, torch.bool

    def get(name, dtype, shape=None):
        t = st.get(name)
        if t is not None and shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return _ptr(t, dtype, name, host)

    A, B, D = cfg.num_actors, cfg.num_bodies, cfg.num_dofs
    b.rigid_body_states = get("rigid_body_states", f32, (n, B, 13))
    b.root_states = get("root_states", f32, (n, A, 13))
    b.dof_states = get("dof_states", f32, (n, D, 2))
    b.dof_forces = get("dof_forces", f32, (n, D))
    pre = st.get("pre_ball_states")
    if pre is not None:
        b.pre_ball_states = _ptr(pre, f32, "pre_ball_states", host)
        if pre.shape[-1] == 13:
            b.pre_ball_stride, b.pre_vx_offset, b.pre_vz_offset = 13, 7, 9      # full clone, TILT:1020
        elif pre.shape[-1] == 2:
            b.pre_ball_stride, b.pre_vx_offset, b.pre_vz_offset = 2, 0, 1       # (vx, vz) only
        else:
            raise ValueError("pre_ball_states must be [N,13] or [N,2]")
    b.initial_root_states = get("initial_root_states", f32, (n, A, 13))
    b.initial_dof_states = get("initial_dof_states", f32, (n, D, 2))
    b.initial_body_states = get("initial_body_states", f32, (n, B, 13))
    b.initial_balance_states = get("initial_balance_states", f32, (n, len(cfg.balance_ids), 6))
    rbv = st.get("reset_ball_vel")
    b.reset_ball_vel = _ptr(rbv, f32, "reset_ball_vel", host)
    b.reset_ball_pos_yz = get("reset_ball_pos_yz", f32, (n, 2))
    b.obs_buf = get("obs_buf", f32)
    b.rew_buf = get("rew_buf", f32)
    b.reset_buf = get("reset_buf", i64, (n,))
    b.progress_buf = get("progress_buf", i64, (n,))
    for i, name in enumerate(cfg.flag_names + cfg.counter_names):
        b.flags[i] = get(name, u8, (n,))
    b.actions = get("actions", f32, (n, D))
    b.pd_action_offset = get("pd_action_offset", f32, (D,))
    b.pd_action_scale = get("pd_action_scale", f32, (D,))
    b.pd_targets = get("pd_targets", f32, (n, D))
    b.stats = _ptr(st.get("stats"), torch.float64, "stats", host)
    b.scratch = _ptr(st.get("scratch"), torch.int32, "scratch", host)
    b.last_hitter = _ptr(st.get("last_hitter"), i64, "last_hitter", host)
    b.obs_moments = _ptr(st.get("obs_moments"), torch.float64, "obs_moments", host)
    # optional VecTask.step envelope outputs
    b.clip_actions = float(st.get("clip_actions", 0.0) or 0.0)
    clip_obs = float(st.get("clip_observations", 0.0) or 0.0)
    b.clip_observations = clip_obs if (clip_obs > 0.0 and clip_obs != float("inf")) else 0.0
    b.timeout_buf = _ptr(st.get("timeout_buf"), i64, "timeout_buf", host)
    if st.get("reset_count") is not None:
        ai, di = st["actor_indices"], st["dof_indices"]
        b.actor_indices = _ptr(ai, i64, "actor_indices", host)
        b.dof_indices = _ptr(di, i64, "dof_indices", host)
        b.dof_indices_per_env = di.numel() // max(n, 1)
        b.reset_count = _ptr(st["reset_count"], torch.int32, "reset_count", host)
        b.reset_actor_indices = _ptr(st["reset_actor_indices"], torch.int32, "reset_actor_indices", host)
        b.reset_dof_indices = _ptr(st["reset_dof_indices"], torch.int32, "reset_dof_indices", host)
    return b


def current_stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream
