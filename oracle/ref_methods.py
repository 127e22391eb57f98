"""TEST INFRASTRUCTURE ONLY -- runs the reference's own CLASS METHOD bodies on a stub `self`.

`ref_extract.py` pins the free functions (reward / observation arithmetic).  The step logic around
them lives in methods of the task classes -- `_reset_idx` (TILT:847-906, A3:825, NES:871, A4:853,
ALIGN:842, ADOF:965), `pre_physics_step` (TILT:1002-1020), `post_physics_step` (TILT:1022-1052,
ADOF:1149-1192), `compute_reward`, `compute_observations`, `reset_idx`,
`generate_random_speed_for_ball` -- and the classes cannot be instantiated here (they need PhysX).
The methods can still be EXECUTED: each one is cut out of the class body by AST position, dedented,
written unmodified into a file-backed temporary module and bound to a plain object whose attributes are
the synthetic state tensors (`isaacgym_b200.synth.make_state`) under the names the reference's
`__init__` gives them (TILT:153-243), with a recording no-op `gym`, `gymtorch.unwrap_tensor = identity`
and `gymapi.Vec3` stand-ins.  Host randomness (`random.uniform`, TILT:307-318) is left to Python's own
generator: tests seed it and replay the same stream through the restated sampler.

Works only where `/root/reference` exists (the build container).  Nothing is copied into the repository.
"""
import ast
import importlib.util
import os
import random
import sys
import tempfile
import textwrap
import types

import torch

from . import ref_extract as R

_HEADER = (
    "import math\n"
    "import random\n"
    "import torch\n"
    "from torch import Tensor\n"
    "from typing import Tuple, Dict, List, Optional\n"
    "from oracle.jit_utils_restated import *\n"
    "from oracle.ref_methods import gymtorch, gymapi\n"
)

# the task class of each file (the first ClassDef that defines post_physics_step)
_METHODS = ("generate_random_speed_for_ball", "compute_reward", "compute_observations", "refresh_sim_tensors", "reset_idx",
            "_reset_idx", "pre_physics_step", "post_physics_step")


class _GymTorch:
    """gymtorch.unwrap_tensor hands a torch tensor to the simulator: identity here."""
    @staticmethod
    def unwrap_tensor(t):
        return t


class _Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class _GymApi:
    Vec3 = _Vec3


gymtorch = _GymTorch()
gymapi = _GymApi()


class RecordingGym:
    """Stand-in for the `gym` handle: every call is recorded and returns True (the reference ignores the
    results, TILT:881-888).  The indexed setters' index tensors are what the tests compare."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def call(*args, **kw):
            self.calls.append((name, args))
            return True
        return call

    def last(self, name):
        for n, a in reversed(self.calls):
            if n == name:
                return a
        return None


_tmpdir = None
_modules = {}


def _method_sources(alias):
    path = os.path.join(R.REFERENCE_ROOT, R.FILES[alias])
    text = open(path, encoding="utf-8").read()
    lines = text.splitlines(keepends=True)
    for node in ast.parse(text).body:
        if isinstance(node, ast.ClassDef) and any(isinstance(c, ast.FunctionDef) and c.name == "post_physics_step" for c in node.body):
            out = {}
            for c in node.body:
                if isinstance(c, ast.FunctionDef) and c.name in _METHODS:
                    out[c.name] = (c.lineno, textwrap.dedent("".join(lines[c.lineno - 1:c.end_lineno])))
            return out
    raise KeyError(f"no task class found in {R.FILES[alias]}")


def load_methods(alias):
    """Module whose top-level functions are the reference's methods of file `alias` (unmodified text)."""
    global _tmpdir
    if alias in _modules:
        return _modules[alias]
    if not R.available():
        raise RuntimeError(f"{R.REFERENCE_ROOT} is not present; the reference loader only works in the build container")
    if _tmpdir is None:
        _tmpdir = tempfile.mkdtemp(prefix="ppk_refm_")
    modname = f"_ppkrefm_{alias}"
    fpath = os.path.join(_tmpdir, modname + ".py")
    with open(fpath, "w", encoding="utf-8") as f:
        f.write(_HEADER)
        for name, (lineno, src) in _method_sources(alias).items():
            f.write(f"\n# cut from {R.FILES[alias]}:{lineno}\n{src}\n")
    spec = importlib.util.spec_from_file_location(modname, fpath)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _modules[alias] = mod
    return mod


ALIAS_OF = {"a3": "A3", "tilt": "TILT", "nes": "NES", "align": "ALIGN", "a4": "A4", "adof": "ADOF", "base": "BASE"}

# launch-sampler ranges set in the constructors: TILT:114-116, A3:111-112, NES:118-120, ALIGN:112-114 (8.8), A4 = TILT,
# ADOF:127-131
_RANGES = {
    "tilt": dict(initial_speed_range=(8.0, 8.6), tilt_angle_range=(-5.0, 5.0), tilt_z_angle_range=(2.0, 10.0)),
    "a4": dict(initial_speed_range=(8.0, 8.6), tilt_angle_range=(-5.0, 5.0), tilt_z_angle_range=(2.0, 10.0)),
    "align": dict(initial_speed_range=(8.0, 8.8), tilt_angle_range=(-5.0, 5.0), tilt_z_angle_range=(2.0, 10.0)),
    "a3": dict(initial_speed_range=(6.5, 7.5), tilt_angle_range=(-5.0, 5.0)),
    "nes": dict(initial_speed_range=(5.4, 5.9), tilt_angle_range=(-5.0, 5.0), tilt_z_angle_range=(10.0, 17.0)),
    "adof": dict(initial_speed_range=(5.0, 5.4), tilt_angle_range=(-8.0, 3.0), tilt_z_angle_range=(14.0, 24.0),
                 initial_pos_y_range=(-0.5, 0.1), initial_pos_z_range=(0.96, 1.05)),
}


def read_ranges_from_reference(variant):
    """The same ranges parsed out of the constructor text, to pin `_RANGES` (and with it the restated sampler)."""
    path = os.path.join(R.REFERENCE_ROOT, R.FILES[ALIAS_OF[variant]])
    out = {}
    for node in ast.walk(ast.parse(open(path, encoding="utf-8").read())):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Attribute):
            name = node.targets[0].attr
            if name in ("initial_speed_range", "tilt_angle_range", "tilt_z_angle_range", "initial_pos_y_range", "initial_pos_z_range"):
                try:
                    out[name] = tuple(float(x) for x in ast.literal_eval(node.value))
                except Exception:
                    pass
    return out


def make_ref_task(cfg, st, overrides=None):
    """A stub `self` over the tensors of `st` (CPU), attribute names as in the reference constructor
    (TILT:125-243, A4:125-190, ADOF:157-293).  Views alias `st`, so what the methods write is visible in `st`."""
    v = cfg.variant
    mod = load_methods(ALIAS_OF[v])
    task = types.SimpleNamespace()
    for name in _METHODS:
        fn = getattr(mod, name, None)
        if fn is not None:
            setattr(task, name, types.MethodType(fn, task))
    n = st["root_states"].shape[0]
    A, D = cfg.num_actors, cfg.num_dofs
    task.device = "cpu"
    task.num_envs, task.actors_per_env = n, A
    task.gym, task.sim = RecordingGym(), object()
    task.randomize, task.headless = False, True
    task.num_steps = 0
    task.max_episode_length = cfg.max_episode_length
    # root / rigid-body / DOF tensors and their views (TILT:156-214)
    task.vec_root_states = st["root_states"]
    task.root_states = st["root_states"].view(n * A, 13)
    task.vec_rb_states = st["rigid_body_states"]
    task.body_states = st["rigid_body_states"]
    task.vec_dof_states = st["dof_states"]
    task.dof_states = st["dof_states"].view(n * D, 2)
    task.dof_pos, task.dof_vel = st["dof_states"][..., 0], st["dof_states"][..., 1]
    task.dof_force_tensor = st["dof_forces"]
    task.initial_vec_root_states = st["initial_root_states"]
    task.initial_pos = st["initial_root_states"][:, :, 0:3]
    task.initial_rot = st["initial_root_states"][:, :, 3:7]
    task.initial_dof_states = st["initial_dof_states"]
    task.humanoid1_root_states = task.vec_root_states[:, cfg.humanoid_actor[0], :]
    task.humanoid1_paddle_rb_states = task.vec_rb_states[:, cfg.paddle_body[0], :]
    task.ball2_root_states = task.vec_root_states[:, cfg.ball_actor, :]
    task.pre_ball2_root_states = st["pre_ball_states"]
    task.body_states_id = torch.tensor(cfg.body_ids, dtype=torch.long)
    task.actor_indices, task.dof_indices = st["actor_indices"], st["dof_indices"]
    task._pd_action_offset, task._pd_action_scale = st["pd_action_offset"], st["pd_action_scale"]
    task.actions = st["actions"]
    # VecTask buffers
    task.obs_buf, task.rew_buf = st["obs_buf"], st["rew_buf"]
    task.reset_buf, task.progress_buf = st["reset_buf"], st["progress_buf"]
    task.randomize_buf = torch.zeros(n, dtype=torch.int64)
    task.reset_buf_force = torch.zeros(n, dtype=torch.int64)
    # reward constants (TILT:99-107)
    task.alpha, task.power_coefficient, task.penalty = cfg.alpha, cfg.power_coefficient, cfg.penalty
    task.hit_table_reward, task.not_hit_table_penalty = cfg.hit_table_reward, cfg.not_hit_table_penalty
    for name in cfg.flag_names + cfg.counter_names + cfg.state_names:
        setattr(task, name, st[name])
    if v == "nes":      # NES:244-248 allocates five flags, two are live; _reset_idx writes all five (NES:913-917)
        for name in ("reward_calculated", "no_bounce_before_half_mask", "net_condition_calculated"):
            setattr(task, name, torch.zeros(n, dtype=torch.bool))
    if v == "adof":     # ADOF:98-116,196-206,249-251
        task.humanoid1_pelvis_rb_states = task.vec_rb_states[:, cfg.pelvis_body, :]
        task.hit_paddle_reward = cfg.hit_paddle_reward
        task.miss_paddle_penalty_coefficient = cfg.miss_paddle_penalty_coefficient
        task.cross_net_reward_float = cfg.cross_net_reward
        task.die_penalty_float = cfg.die_penalty
        task.initial_body_states = st["initial_body_states"]
        task.initial_dof_pos = st["initial_dof_states"][..., 0]        # ADOF:250-251
        task.initial_dof_vel = st["initial_dof_states"][..., 1]
        task.body_balance_states_id = torch.tensor(cfg.balance_ids, dtype=torch.long)
        task.is_g1, task.is_train = True, cfg.is_train
    for key, val in _RANGES.get(v, {}).items():
        setattr(task, key, val)
    for key, val in (overrides or {}).items():
        setattr(task, key, val)
    return task


STEP_VARIANTS = ("a3", "tilt", "nes", "align", "adof")          # classes whose post_physics_step runs as shipped


def bind_step_functions(variant):
    """Bind, for the methods of `variant`'s class, the free functions they call by module-level name to the reference's
    own text (ref_extract).  ALIGN's second `compute_pingpong_reward` shadows the live one and does not compile
    (defect D7): the first definition (ALIGN:1097) is bound, as in the file up to that line."""
    ld = R.load
    alias = ALIAS_OF[variant]
    names = {"compute_humanoid_observations": ld(alias, "compute_humanoid_observations"),
             "compute_pingpong_observations": ld(alias, "compute_pingpong_observations")}
    if variant == "a3":
        names["compute_pingpong_reward"] = ld("A3", "compute_pingpong_reward")
    elif variant == "tilt":
        names["compute_pingpong_reward_nv"] = ld("TILT", "compute_pingpong_reward_nv")
    elif variant == "nes":
        names["compute_pingpong_reward_only_paddle"] = ld("NES", "compute_pingpong_reward_only_paddle")
    elif variant == "align":
        names["compute_pingpong_reward"] = ld("ALIGN", "compute_pingpong_reward", 0)
    elif variant == "adof":
        names["compute_pingpong_reward_nv"] = R.load_adof_reward()
        names["compute_imitation_observations"] = ld("ADOF", "compute_imitation_observations")
    return bind_free_functions(variant, names)


def bind_free_functions(variant, names):
    """The methods call the file's free functions by module-level name (TILT:740 compute_pingpong_reward_nv, ...):
    give the methods module the reference's own functions (ref_extract), under those names."""
    mod = load_methods(ALIAS_OF[variant])
    for name, fn in names.items():
        setattr(mod, name, fn)
    return mod


def replay_launch_draws(variant, seed, count):
    """What `_reset_idx` draws for `count` resetting envs after `random.seed(seed)`, through the RESTATED sampler
    (oracle.pingpong_oracle.sample_ball_velocity): (ball_vel [count,3] fp32, ball_pos_yz [count,2] fp32 or None).
    The reference builds `torch.tensor([vx, vy, vz])` from Python floats: fp64 -> fp32 rounding, as here."""
    from . import pingpong_oracle as O
    rng = random.Random(seed)
    vel, yz = [], []
    for _ in range(count):
        if variant == "adof":                 # ADOF:976-977 draws the position first
            yz.append((rng.uniform(*_RANGES["adof"]["initial_pos_y_range"]), rng.uniform(*_RANGES["adof"]["initial_pos_z_range"])))
        vel.append(O.sample_ball_velocity(rng, variant))
    vel_t = torch.tensor(vel, dtype=torch.float64).to(torch.float32).reshape(count, 3)
    yz_t = torch.tensor(yz, dtype=torch.float64).to(torch.float32).reshape(count, 2) if yz else None
    return vel_t, yz_t
