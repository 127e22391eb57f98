"""GPU parity of the fp32 (rollout-forward) first layer: `FirstLayer(precision="fp32")` -> ppk_policy_first_layer_f32.

Oracle: plain torch fp32 on the CPU, elu(linear(running_mean_std(obs))) (`oracle/policy_oracle.first_layer_fp32`).
Tolerance (north_star: fp32 within 1e-5 relative): |got - want| <= 1e-5 |want| + 2e-6 (sum_k |x_k w_k| + |b|).  The second
term is the rounding noise of the dot product itself -- two fp32 summation orders of the same 80..96 products differ by
that much where the terms cancel -- and is what the TF32 hi/lo split products add (<= 3 * 2^-22 per product)."""
import os
import subprocess
import sys

import pytest
import torch

from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd
from oracle import policy_oracle as P

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def batch(rows, width, seed, scale=2.0, shift=0.7):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, width, generator=g) * scale + shift


def assert_fp32_close(got, want, scale, ctx):
    g = got.float().cpu()
    err = (g - want).abs()
    tol = 1e-5 * want.abs() + 2e-6 * scale
    bad = err > tol
    assert int(bad.sum()) == 0, f"{ctx}: {int(bad.sum())} of {bad.numel()} entries outside tolerance; worst " \
                                f"err/tol {float((err / tol).max()):.2f}, max abs err {float(err.max()):.3e}"
    # and the typical entry is far inside it: the split products are not a 1e-5 method
    assert float((err / tol).mean()) < 0.1, f"{ctx}: mean err/tol {float((err / tol).mean()):.3f}"


def reference_state(width):
    rms_ref = P.RunningMeanStd(width)
    rms_ref.update(batch(5000, width, 1))
    rms = RunningMeanStd(width, device=DEV).eval()
    rms.running_mean.copy_(rms_ref.running_mean)
    rms.running_var.copy_(rms_ref.running_var)
    return rms_ref, rms


@pytest.mark.parametrize("rows,width,units", [(4096, 80, 2048), (1000, 80, 512), (129, 94, 256), (5, 24, 256),
                                              (20000, 80, 4096), (128 * 149, 80, 256), (128 * 148 + 1, 94, 512)])
@pytest.mark.parametrize("act", ["elu", "None"])
def test_fp32_first_layer_matches_torch_fp32(rows, width, units, act):
    g = torch.Generator().manual_seed(rows + width)
    x = batch(rows, width, rows)
    w = torch.randn(units, width, generator=g) / width ** 0.5
    b = torch.randn(units, generator=g) * 0.1
    rms_ref, rms = reference_state(width)
    layer = FirstLayer(w.to(DEV), b.to(DEV), activation=act, running_mean_std=rms, precision="fp32")
    got = layer(x.to(DEV))
    torch.cuda.synchronize()
    want, scale = P.first_layer_fp32(rms_ref.normalize(x), w, b, act)
    assert got.dtype == torch.float32 and tuple(got.shape) == (rows, units)
    assert_fp32_close(got, want, scale, f"rows={rows} width={width} units={units} act={act}")


def test_fp32_first_layer_without_normalisation_and_without_bias():
    g = torch.Generator().manual_seed(8)
    x = batch(777, 80, 8, scale=1.0, shift=0.0)
    w = torch.randn(256, 80, generator=g) / 80 ** 0.5
    got = FirstLayer(w.to(DEV), None, activation="elu", running_mean_std=None, precision="fp32")(x.to(DEV))
    want, scale = P.first_layer_fp32(x, w, None, "elu")
    assert_fp32_close(got, want, scale, "no rms, no bias")


def test_fp32_elu_is_accurate_near_zero_and_in_the_tail():
    """Pre-activations planted across the two branches of the device ELU (polynomial for x in (-0.25, 0], exp2 - 1 below):
    identity weights pick single inputs, so want = elu(x) exactly and the relative error of the activation is visible."""
    width, units = 80, 256
    xs = torch.cat([-torch.logspace(-6, 1.2, 4000), torch.logspace(-6, 1, 96), torch.tensor([0.0, -0.25, -0.2500001, -0.2499999])])
    rows = xs.numel()
    x = torch.zeros(rows, width)
    x[:, 3] = xs
    w = torch.zeros(units, width)
    w[:, 3] = 1.0
    got = FirstLayer(w.to(DEV), None, activation="elu", running_mean_std=None, precision="fp32")(x.to(DEV)).cpu()
    want = torch.nn.functional.elu(xs.double()).float()
    rel = ((got[:, 0] - want).abs() / want.abs().clamp_min(1e-30))
    assert float(rel.max()) < 2e-6, f"ELU relative error {float(rel.max()):.3e} at x={float(xs[rel.argmax()])}"
    assert torch.equal(got[:, 0], got[:, 255])


def test_fp32_first_layer_full_size_properties():
    """65 536 x 80 -> 2048: deterministic, rows independent of their position in the batch (and of which CTA of a cluster
    pair gets them), a random sample of rows against the oracle."""
    rows, width, units = 65536, 80, 2048
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(rows, width, device=DEV, generator=g) * 2 + 0.3
    w = torch.randn(units, width, device=DEV, generator=g) / width ** 0.5
    b = torch.randn(units, device=DEV, generator=g) * 0.1
    rms = RunningMeanStd(width, device=DEV)
    rms.update(x)
    rms.eval()
    layer = FirstLayer(w, b, "elu", rms, precision="fp32")
    full = layer(x)
    assert torch.equal(full, layer(x))
    for lo, hi in ((0, 1000), (12345, 13345), (65536 - 77, 65536), (128, 256)):
        part = layer(x[lo:hi].contiguous())
        assert torch.equal(part, full[lo:hi]), f"rows {lo}:{hi} depend on their position in the batch"
    idx = torch.randint(0, rows, (1500,), generator=torch.Generator().manual_seed(1))
    ref = P.RunningMeanStd(width)
    ref.running_mean, ref.running_var = rms.running_mean.cpu(), rms.running_var.cpu()
    want, scale = P.first_layer_fp32(ref.normalize(x[idx.to(DEV)].cpu()), w.cpu(), b.cpu(), "elu")
    assert_fp32_close(full[idx.to(DEV)], want, scale, "full-size sample")


def test_fp32_first_layer_against_torch_cuda_fp32():
    """What the rollout runs on this GPU: torch fp32 linear (cuBLAS SGEMM, TF32 off = torch's default) + elu."""
    rows, width, units = 8192, 80, 2048
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(rows, width, device=DEV, generator=g) * 1.5 + 0.2
    lin = torch.nn.Linear(width, units).to(DEV)
    rms = RunningMeanStd(width, device=DEV)
    rms.update(x)
    rms.eval()
    got = FirstLayer(lin.weight, lin.bias, "elu", rms, precision="fp32")(x)
    xn = rms.normalize(x)
    assert not torch.backends.cuda.matmul.allow_tf32
    with torch.no_grad():
        want = torch.nn.functional.elu(lin(xn))
        scale = xn.abs() @ lin.weight.abs().t() + lin.bias.abs()
    assert_fp32_close(got, want.cpu(), scale.cpu(), "torch cuda fp32")


def test_fp32_first_layer_unsupported_width_is_an_error():
    w = torch.zeros(256, 313, device=DEV)
    layer = FirstLayer(w, None, "elu", None, precision="fp32")
    with pytest.raises(RuntimeError):
        layer(torch.zeros(10, 313, device=DEV))


def test_fp32_first_layer_cta_pair_variant_also_passes():
    """PPK_FL32_CLUSTER=2 (CTA pairs, cta_group::2: one MMA of M = 256 over the two CTAs of a cluster, each holding half
    of the weight stage) is the A/B build of the same kernel; rows must not depend on which CTA of the pair gets them."""
    code = (
        "import torch\n"
        "from isaacgym_b200.policy_input import FirstLayer\n"
        "g = torch.Generator().manual_seed(5)\n"
        "x = torch.randn(3000, 80, generator=g); w = torch.randn(512, 80, generator=g) / 9; b = torch.randn(512, generator=g)\n"
        "layer = FirstLayer(w.cuda(), b.cuda(), 'elu', None, precision='fp32')\n"
        "got = layer(x.cuda()).cpu()\n"
        "want = torch.nn.functional.elu(torch.nn.functional.linear(x, w, b))\n"
        "scale = x.abs() @ w.abs().t() + b.abs()\n"
        "assert bool(((got - want).abs() <= 1e-5 * want.abs() + 2e-6 * scale).all())\n"
        "assert torch.equal(layer(x[128:900].contiguous().cuda()).cpu(), got[128:900])\n"
        "print('ok')\n")
    env = dict(os.environ, PPK_FL32_CLUSTER="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
