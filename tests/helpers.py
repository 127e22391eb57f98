"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ("base", "a3", "tilt", "nes", "align", "a4", "adof", "align2")


STEP_VARIANTS = ("a3", "tilt", "nes", "align", "adof")     # fixtures of one whole post_physics_step (<variant>_step.npz)


def load_golden(variant, step=False):
    """-> (inputs: dict of CPU tensors, outputs: dict of CPU tensors) frozen from the reference: the free functions
    (`<variant>.npz`) or, `step=True`, one post_physics_step run from the reference's own method bodies."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{variant}_step.npz" if step else f"{variant}.npz"))
    ins = {k[4:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("in__")}
    outs = {k[5:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("out__")}
    return ins, outs


def oracle_step_with_term_scale(cfg, st, fn):
    """Run oracle call `fn(cfg, st)` and return (its result, per-env sum of |terms| that entered each reward),
    shaped like rew_buf.  The scale is what the fp32 rounding of the reward sum lives on."""
    from oracle import pingpong_oracle as O
    O.TERM_SCALE = []
    try:
        res = fn(cfg, st)
        scales = list(O.TERM_SCALE)
    finally:
        O.TERM_SCALE = None
    assert scales, "the oracle reward did not report its terms"
    scale = scales[0] if len(scales) == 1 else torch.stack(scales, dim=-1)
    return res, scale


def assert_close_fields(cfg, got_obs, want_obs, got_rew, want_rew, context="", rew_scale=None):
    """Per-field tolerances of SURVEY.md 8(d): rotated fields rtol 1e-5 + atol 1e-6*scale,
    copied fields rtol 1e-6.  Reward: rtol 1e-5 plus an absolute term PER ENV: 2e-6 * (1 + sum of the
    magnitudes of the terms that entered that env's reward) when the oracle reported them (`rew_scale`,
    see oracle_step_with_term_scale) -- an env whose reward is only 1/(1+1.5 d^2) gets ~1e-5, an env that was
    paid 2000 gets 4e-3.  Without `rew_scale` (golden fixtures, which carry no term breakdown) the absolute
    term is 1e-5 * max(1, |reward|) per env."""
    got_obs, want_obs = got_obs.double().cpu(), want_obs.double().cpu()
    if cfg.variant == "base":
        torch.testing.assert_close(got_obs, want_obs, rtol=0, atol=0, msg=lambda m: f"{context} base obs: {m}")
    else:
        n = got_obs.shape[0]
        g = got_obs.reshape(n * cfg.obs_rows, cfg.num_obs)
        w = want_obs.reshape(n * cfg.obs_rows, cfg.num_obs)
        J, D = len(cfg.body_ids), cfg.num_dofs
        rot_end = 6 * J
        dof_end = rot_end + 2 * D
        scale = w.abs().amax(dim=-1, keepdim=True).clamp_min(1.0)
        def chk(lo, hi, rtol, atol_scale, what):
            err = (g[:, lo:hi] - w[:, lo:hi]).abs()
            tol = rtol * w[:, lo:hi].abs() + atol_scale * scale
            bad = err > tol
            assert not bad.any(), (f"{context} {what}: {int(bad.sum())} of {bad.numel()} outside tolerance; "
                                   f"max err {float(err.max()):.3e}")
        chk(0, rot_end, 1e-5, 1e-6, "rotated body pos/vel")
        chk(rot_end, dof_end, 1e-6, 0.0, "dof_pos / 0.1*dof_vel")
        chk(dof_end, dof_end + 6, 1e-5, 1e-6, "ball local pos/vel")
        if cfg.variant == "adof":
            yi = dof_end + 6
            # y_intersect divides by (-lvx + 1e-6): rtol 1e-4 and the amplification of lvx's own error
            lvx = w[:, dof_end + 3]
            amp = (w[:, yi].abs() * 4e-6 * scale[:, 0] / (lvx - 1e-6).abs().clamp_min(1e-30))
            err = (g[:, yi] - w[:, yi]).abs()
            tol = 1e-4 * w[:, yi].abs() + 1e-5 + amp
            assert not (err > tol).any(), f"{context} y_intersect: max err {float(err.max()):.3e}"
            nb = len(cfg.balance_ids)
            chk(yi + 1, yi + 1 + 6 * nb, 1e-5, 1e-5, "imitation pos/vel diffs")
            chk(yi + 1 + 6 * nb, cfg.num_obs, 1e-6, 0.0, "reference dof pos/vel")
    gr, wr = got_rew.double().cpu(), want_rew.double().cpu()
    err = (gr - wr).abs()
    if rew_scale is not None:
        tol = 1e-5 * wr.abs() + 2e-6 * (1.0 + rew_scale.double().cpu().reshape(wr.shape))
    else:
        tol = 1e-5 * wr.abs() + 1e-5 * wr.abs().clamp_min(1.0)
    bad = err > tol
    assert not bad.any(), (f"{context} reward: {int(bad.sum())} envs outside tolerance; worst err "
                           f"{float((err - tol).max() + tol[(err - tol).argmax()]):.3e} vs tol {float(tol[(err - tol).argmax()]):.3e}")
