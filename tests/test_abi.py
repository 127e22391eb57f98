"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, exports every
symbol include/ppk.h declares, the ctypes mirrors match the C struct sizes, and argument
validation fails loudly (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from isaacgym_b200 import _native as N
from isaacgym_b200 import build as B
from isaacgym_b200.config import CONFIGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppk.h")


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.load()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppk_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert "ppk_post_physics_step" in names and "ppk_reset_idx" in names and len(names) >= 9
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ppk.h but not exported"


def test_abi_version_and_strerror(lib):
    assert lib.ppk_abi_version() == N.ABI_VERSION
    assert lib.ppk_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert len(lib.ppk_strerror(code)) > 3


def test_struct_sizes_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ppk.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(PpkTask), sizeof(PpkBuffers), sizeof(PpkRunningMeanStd));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(N.PpkTask) and b == C.sizeof(N.PpkBuffers) and c == C.sizeof(N.PpkRunningMeanStd)


def test_argument_validation_returns_error_codes(lib):
    task = N.make_task(CONFIGS["tilt"])
    buf = N.PpkBuffers()
    buf.struct_size = C.sizeof(N.PpkBuffers)
    buf.num_envs = 128
    assert lib.ppk_post_physics_step(None, None, N.PHASE_ALL, None) == -1           # PPK_ERR_NULL
    assert lib.ppk_post_physics_step(task, buf, N.PHASE_ALL, None) == -1            # missing tensors
    bad = N.PpkBuffers()
    bad.struct_size = 4
    assert lib.ppk_post_physics_step(task, bad, N.PHASE_ALL, None) == -6            # PPK_ERR_ABI
    t2 = N.make_task(CONFIGS["tilt"])
    t2.variant = 99
    assert lib.ppk_post_physics_step(t2, buf, N.PHASE_ALL, None) == -4              # PPK_ERR_VARIANT
    t3 = N.make_task(CONFIGS["tilt"])
    t3.num_dofs = 9
    buf2 = N.PpkBuffers()
    buf2.struct_size = C.sizeof(N.PpkBuffers)
    buf2.num_envs = 4
    for f in ("rigid_body_states", "root_states", "dof_states", "dof_forces", "pre_ball_states", "obs_buf", "rew_buf",
              "reset_buf", "progress_buf", "initial_root_states", "initial_dof_states", "reset_ball_vel"):
        setattr(buf2, f, 0x1000)
    for i in range(3):
        buf2.flags[i] = 0x1000
    buf2.pre_ball_stride, buf2.pre_vx_offset, buf2.pre_vz_offset = 2, 0, 1
    assert lib.ppk_post_physics_step(t3, buf2, N.PHASE_ALL & ~N.PHASE_STATS, None) == -2   # PPK_ERR_SHAPE
    buf2.progress_buf = 0x1004
    assert lib.ppk_post_physics_step(task, buf2, N.PHASE_ALL & ~N.PHASE_STATS, None) == -3  # PPK_ERR_ALIGN
    with pytest.raises(RuntimeError):
        N.check(-2, "x")
    # PPK_PHASE_MOMENTS: needs PPK_PHASE_OBS, a moments buffer, and a variant of the family kernel
    buf2.progress_buf = 0x1000
    assert lib.ppk_post_physics_step(task, buf2, N.PHASE_ALL | N.PHASE_MOMENTS, None) == -1          # obs_moments NULL
    buf2.obs_moments = 0x1004
    assert lib.ppk_post_physics_step(task, buf2, N.PHASE_ALL | N.PHASE_MOMENTS, None) == -3          # not 8-byte aligned
    buf2.obs_moments = 0x1000
    assert lib.ppk_post_physics_step(task, buf2, N.PHASE_REWARD | N.PHASE_MOMENTS, None) == -4       # without PHASE_OBS
    assert lib.ppk_post_physics_step(N.make_task(CONFIGS["adof"]), buf2, N.PHASE_ALL | N.PHASE_MOMENTS, None) == -4
    assert lib.ppk_post_physics_step(task, buf2, 64, None) == -4                                      # unknown phase bit


def test_task_descriptor_follows_the_reference_configs():
    t = N.make_task(CONFIGS["adof"])
    assert (t.num_actors, t.num_bodies, t.num_dofs) == (3, 42, 27)
    assert list(t.body_ids[0])[:10] == [0, 31, 32, 33, 34, 35, 36, 37, 38, 39]
    assert t.num_balance_ids == 23 and t.max_episode_length == 160
    assert N.make_task(CONFIGS["a4"]).write_flags == 0 and N.make_task(CONFIGS["tilt"]).write_flags == 1
    assert N.make_task(CONFIGS["nes"]).reset_dof == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "isaacgym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_learner_side_argument_validation(lib):
    assert lib.ppk_linear_packed_bytes(2048, 80) == 2048 * 96 * 2      # K padded to a multiple of 16 with room for the bias column
    assert lib.ppk_linear_packed_bytes(2048, 94) == 2048 * 96 * 2
    assert lib.ppk_linear_packed_bytes(1000, 80) == 0                             # units % 256 != 0
    rms = N.PpkRunningMeanStd()
    rms.struct_size = C.sizeof(N.PpkRunningMeanStd)
    rms.width = 80
    assert lib.ppk_rms_update(rms, 0x1000, 16, None) == -1                        # NULL statistics
    rms.struct_size = 8
    assert lib.ppk_rms_normalize(rms, 0x1000, 16, 0x1000, None) == -6             # PPK_ERR_ABI
    assert lib.ppk_policy_first_layer(None, 0x1000, 16, 80, 0x1000, 1000, 1, 0x1000, None) == -2   # units
    assert lib.ppk_policy_first_layer(None, 0x1000, 16, 200, 0x1000, 2048, 1, 0x1000, None) == -2  # width without a tiling
    assert lib.ppk_policy_first_layer(None, 0x1000, 16, 80, 0x1000, 2048, 7, 0x1000, None) == -4   # activation
    assert lib.ppk_policy_first_layer(None, None, 16, 80, 0x1000, 2048, 1, 0x1000, None) == -1
    assert lib.ppk_policy_first_layer(None, 0x1000, 0, 80, 0x1000, 2048, 1, 0x1000, None) == 0     # empty batch
    # fp32 (rollout-forward) variant: 8 bytes per packed element (TF32 hi + lo), K padded to a multiple of 8 with the bias column
    assert lib.ppk_linear_packed_bytes_f32(2048, 80) == 2048 * 88 * 8
    assert lib.ppk_linear_packed_bytes_f32(2048, 94) == 2048 * 96 * 8
    assert lib.ppk_linear_packed_bytes_f32(256, 24) == 256 * 32 * 8
    assert lib.ppk_linear_packed_bytes_f32(1000, 80) == 0
    assert lib.ppk_linear_pack_f32(None, None, 256, 80, 0x1000, 1 << 20, None) == -1
    assert lib.ppk_linear_pack_f32(0x1000, None, 256, 80, 0x1000, 16, None) == -2                  # packed buffer too small
    assert lib.ppk_policy_first_layer_f32(None, 0x1000, 16, 80, 0x1000, 1000, 1, 0x1000, None) == -2   # units
    assert lib.ppk_policy_first_layer_f32(None, 0x1000, 16, 80, 0x1000, 2048, 7, 0x1000, None) == -4   # activation
    assert lib.ppk_policy_first_layer_f32(None, None, 16, 80, 0x1000, 2048, 1, 0x1000, None) == -1
    assert lib.ppk_policy_first_layer_f32(None, 0x1000, 16, 80, 0x1004, 2048, 1, 0x1000, None) == -3   # packed weights not 16-byte aligned
    assert lib.ppk_policy_first_layer_f32(None, 0x1000, 0, 80, 0x1000, 2048, 1, 0x1000, None) == 0     # empty batch


def test_learner_side_host_mirror_is_cuda_only():
    """No CPU or torch fallback behind the host mirrors: they refuse to exist on the CPU."""
    import torch
    from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd
    with pytest.raises(RuntimeError):
        RunningMeanStd(80, device="cpu")
    with pytest.raises(RuntimeError):
        FirstLayer(torch.zeros(256, 80), torch.zeros(256))
