"""GPU parity of the learner-side consumers of obs_buf through the C ABI:
RunningMeanStd (fp64 statistics, fp32 output) and the fused normalise + first MLP layer (tcgen05, fp16).
Tolerances: running_mean / running_var rtol 1e-6 (fp64 sums on the device vs torch's fp32 batch moments
promoted to fp64); normalised obs bit-exact given the same statistics; first layer: identical up to the
fp32 summation order inside the dot product -> at most ONE fp16 ulp of the linear output (+2e-5 absolute
where the terms cancel; two ulps after ELU, see assert_fp16_close), on a small fraction of the entries."""
import pytest
import torch

from isaacgym_b200.policy_input import FirstLayer, RunningMeanStd
from oracle import policy_oracle as P

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def batch(rows, width, seed, scale=2.0, shift=0.7):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, width, generator=g) * scale + shift


@pytest.mark.parametrize("width", [80, 94, 313, 24, 1, 512])      # 1: rl_games' value normalisation (normalize_value)
def test_running_mean_std_matches_restatement(width):
    ours = RunningMeanStd(width, device=DEV)
    ref = P.RunningMeanStd(width)
    for i, rows in enumerate((4096, 1000, 33)):
        x = batch(rows, width, 10 + i)
        y = ours.forward(x.to(DEV))
        want = ref.forward(x)
        torch.testing.assert_close(ours.running_mean.cpu(), ref.running_mean, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(ours.running_var.cpu(), ref.running_var, rtol=1e-6, atol=1e-7)
        assert float(ours.count) == float(ref.count)
        torch.testing.assert_close(y.cpu(), want, rtol=1e-5, atol=1e-6)
    # frozen statistics (rollout / eval mode): the output is bit-exact given the same fp64 state
    ours.eval()
    ref.training = False
    ref.running_mean, ref.running_var = ours.running_mean.cpu(), ours.running_var.cpu()
    x = batch(2048, width, 99)
    assert torch.equal(ours(x.to(DEV)).cpu(), ref.forward(x))
    assert float(ours.count) == float(ref.count)          # eval mode does not update


def test_observation_clamp_feeds_the_statistics():
    ours = RunningMeanStd(80, clip_obs=1.5, device=DEV)
    ref = P.RunningMeanStd(80)
    x = batch(3000, 80, 4, scale=3.0)
    y = ours.forward(x.to(DEV))
    want = ref.forward(P.clip_observations(x, 1.5))
    torch.testing.assert_close(ours.running_var.cpu(), ref.running_var, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(y.cpu(), want, rtol=1e-5, atol=1e-6)


def assert_fp16_close(got, want, ctx, ulps=1):
    g, w = got.float().cpu(), want.float()
    # one fp16 ulp, plus the fp32 summation-order noise of an 80..96-term dot product whose terms are O(1)
    # (it dominates when the terms cancel to a result near zero, where an fp16 ulp is tiny)
    # ELU: a one-ulp flip of the fp16 pre-activation x in (-0.69, -0.5) moves exp(x)-1, which has a finer
    # ulp than x, by up to two of ITS ulps -> ulps=2 for activated outputs
    bad = (g - w).abs() > ulps * 2.0 ** -10 * w.abs() + 2e-5
    assert int(bad.sum()) == 0, f"{ctx}: {int(bad.sum())} of {bad.numel()} entries off by more than one fp16 ulp; " \
                                f"max abs err {float((g - w).abs().max())}"
    assert float((g != w).float().mean()) < 0.02, f"{ctx}: too many entries differ in the last bit"


@pytest.mark.parametrize("rows,width,units", [(4096, 80, 2048), (1000, 80, 512), (129, 94, 256), (5, 24, 256),
                                              (20000, 80, 4096), (3000, 313, 1024), (700, 313, 256)])
@pytest.mark.parametrize("act", ["elu", "None"])
def test_first_layer_matches_autocast_restatement(rows, width, units, act):
    g = torch.Generator().manual_seed(rows + width)
    x = batch(rows, width, rows)
    w = torch.randn(units, width, generator=g) / width ** 0.5
    b = torch.randn(units, generator=g) * 0.1
    rms_ref = P.RunningMeanStd(width)
    rms_ref.update(batch(5000, width, 1))
    rms = RunningMeanStd(width, device=DEV).eval()
    rms.running_mean.copy_(rms_ref.running_mean)
    rms.running_var.copy_(rms_ref.running_var)
    layer = FirstLayer(w.to(DEV), b.to(DEV), activation=act, running_mean_std=rms)
    got = layer(x.to(DEV))
    torch.cuda.synchronize()
    want = P.first_layer(rms_ref.normalize(x), w, b, act)
    assert got.dtype == torch.float16 and tuple(got.shape) == (rows, units)
    assert_fp16_close(got, want, f"rows={rows} width={width} units={units} act={act}", ulps=2 if act == "elu" else 1)


def test_first_layer_without_normalisation_and_without_bias():
    g = torch.Generator().manual_seed(8)
    x = batch(777, 80, 8, scale=1.0, shift=0.0)
    w = torch.randn(256, 80, generator=g) / 80 ** 0.5
    layer = FirstLayer(w.to(DEV), None, activation="elu", running_mean_std=None)
    got = layer(x.to(DEV))
    assert_fp16_close(got, P.first_layer(x, w, None, "elu"), "no rms, no bias", ulps=2)


def test_first_layer_consumes_the_task_step_output():
    """obs_buf straight out of the fused task step -> normalise -> first layer, both sides of the path."""
    from isaacgym_b200.config import CONFIGS
    from isaacgym_b200.synth import make_state
    from isaacgym_b200.tasks import make_task
    cfg = CONFIGS["tilt"]
    st = make_state(cfg, 4096, seed=12, adversarial=False)
    task = make_task("tilt", st, device=DEV)
    task.post_physics_step()
    g = torch.Generator().manual_seed(1)
    w = torch.randn(2048, 80, generator=g) / 80 ** 0.5
    b = torch.zeros(2048)
    rms = RunningMeanStd(80, device=DEV)
    rms.update(task.obs_buf)
    rms.eval()
    got = FirstLayer(w.to(DEV), b.to(DEV), "elu", rms)(task.obs_buf)
    ref = P.RunningMeanStd(80)
    ref.running_mean, ref.running_var = rms.running_mean.cpu(), rms.running_var.cpu()
    want = P.first_layer(ref.normalize(task.obs_buf.cpu()), w, b, "elu")
    assert_fp16_close(got, want, "task step -> first layer", ulps=2)


def test_sharded_moments_merge_equals_single_update():
    """Data-parallel form: every rank accumulates the moments of its shard, the moments and the row count are
    summed across ranks (what all_reduce(SUM) does), every rank merges the global batch.  Emulated on one GPU
    with two shards of unequal size; equals the single-rank update on the concatenated batch."""
    import ctypes as C
    from isaacgym_b200 import _native as N
    lib = N.load()
    x = batch(5000, 80, 21).to(DEV)
    shards = [x[:1800].contiguous(), x[1800:].contiguous()]
    single = RunningMeanStd(80, device=DEV)
    single.update(batch(700, 80, 20).to(DEV))             # non-trivial prior state
    ranks = [RunningMeanStd(80, device=DEV) for _ in shards]
    for r in ranks:
        r.running_mean.copy_(single.running_mean); r.running_var.copy_(single.running_var); r.count.copy_(single.count)
    single.update(x)
    for r, sh in zip(ranks, shards):
        N.check(lib.ppk_rms_accumulate(r._struct(), sh.data_ptr(), sh.shape[0], N.current_stream_ptr()))
    total = ranks[0]._moments[:160] + ranks[1]._moments[:160]          # the all-reduce
    for r in ranks:
        r._moments[:160].copy_(total)
        N.check(lib.ppk_rms_merge(r._struct(), float(x.shape[0]), N.current_stream_ptr()))
        torch.testing.assert_close(r.running_mean, single.running_mean, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(r.running_var, single.running_var, rtol=1e-11, atol=1e-12)
        assert float(r.count) == float(single.count)
        assert float(r._moments[:160].abs().sum()) == 0.0             # cleared for the next batch


def test_first_layer_full_size_properties():
    """65 536 x 80 -> 2048 (the TILT batch): deterministic, every output row depends on its input row only (a slice
    of the batch run on its own gives bit-identical rows, whatever tile it lands in), and a random sample of rows
    matches the oracle."""
    rows, width, units = 65536, 80, 2048
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(rows, width, device=DEV, generator=g) * 2 + 0.3
    w = torch.randn(units, width, device=DEV, generator=g) / width ** 0.5
    b = torch.randn(units, device=DEV, generator=g) * 0.1
    rms = RunningMeanStd(width, device=DEV)
    rms.update(x)
    rms.eval()
    layer = FirstLayer(w, b, "elu", rms)
    full = layer(x)
    again = layer(x)
    assert torch.equal(full, again)
    for lo, hi in ((0, 1000), (12345, 13345), (65536 - 77, 65536)):
        part = layer(x[lo:hi].contiguous())
        assert torch.equal(part, full[lo:hi]), f"rows {lo}:{hi} depend on their position in the batch"
    idx = torch.randint(0, rows, (1500,), generator=torch.Generator().manual_seed(1))
    ref = P.RunningMeanStd(width)
    ref.running_mean, ref.running_var = rms.running_mean.cpu(), rms.running_var.cpu()
    want = P.first_layer(ref.normalize(x[idx.to(DEV)].cpu()), w.cpu(), b.cpu(), "elu")
    assert_fp16_close(full[idx.to(DEV)], want, "full-size sample", ulps=2)


@pytest.mark.parametrize("width,units", [(80, 2048), (313, 512)])
def test_first_layer_against_torch_cuda_autocast(width, units):
    """The thing the reference's learner actually runs: torch.autocast("cuda", float16) over
    elu(linear(normalised obs)) -- cuBLAS fp16 GEMM (fp32 accumulate) + torch's elementwise kernels, on this GPU.
    Same tolerance as against the CPU restatement (summation order inside the dot product)."""
    rows = 8192
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(rows, width, device=DEV, generator=g) * 1.5 + 0.2
    lin = torch.nn.Linear(width, units).to(DEV)
    rms = RunningMeanStd(width, device=DEV)
    rms.update(x)
    rms.eval()
    got = FirstLayer(lin.weight, lin.bias, "elu", rms)(x)
    xn = rms.normalize(x)                                   # bit-exact fp32 normalisation (tested above)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        want = torch.nn.functional.elu(lin(xn))
    assert want.dtype == torch.float16
    assert_fp16_close(got, want.cpu(), f"torch cuda autocast width={width}", ulps=2)


def test_first_layer_with_a_4_byte_aligned_observation_view():
    """obs handed over as a view that starts 4 bytes into an allocation: the row-tile preparation takes scalar
    loads instead of 16-byte ones, the moments / normalise kernels their one-column layout -- same results."""
    rows, width, units = 3000, 80, 512
    g = torch.Generator(device=DEV).manual_seed(9)
    flat = torch.randn(rows * width + 1, device=DEV, generator=g)
    x = flat[1:].view(rows, width)
    assert x.data_ptr() % 16 != 0 and x.is_contiguous()
    xa = x.clone()                                            # the same values, 16-byte aligned
    w = torch.randn(units, width, device=DEV, generator=g) / width ** 0.5
    b = torch.randn(units, device=DEV, generator=g) * 0.1
    r1, r2 = RunningMeanStd(width, device=DEV), RunningMeanStd(width, device=DEV)
    r1.update(x)
    r2.update(xa)
    torch.testing.assert_close(r1.running_mean, r2.running_mean, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(r1.running_var, r2.running_var, rtol=1e-12, atol=1e-12)
    r1.eval(); r2.eval()
    r2.running_mean.copy_(r1.running_mean); r2.running_var.copy_(r1.running_var)
    assert torch.equal(r1.normalize(x), r2.normalize(xa))
    assert torch.equal(FirstLayer(w, b, "elu", r1)(x), FirstLayer(w, b, "elu", r2)(xa))


@pytest.mark.parametrize("variant,n", [("tilt", 4096), ("tilt", 1001), ("a4", 2000), ("nes", 65536)])
def test_step_kernel_moments_feed_running_mean_std(variant, n):
    """PPK_PHASE_MOMENTS: the fused task step leaves the fp64 column sums / sums of squares of the obs rows it wrote;
    RunningMeanStd.update_from_step merges them without reading obs_buf -- same statistics as update(obs_buf) and as the
    oracle's RunningMeanStd, over several steps (the slots are cleared by every fold)."""
    from isaacgym_b200.config import CONFIGS
    from isaacgym_b200.synth import make_state
    from isaacgym_b200.tasks import make_task
    cfg = CONFIGS[variant]
    task = make_task(variant, make_state(cfg, n, seed=5), device=DEV, fused_moments=True, clip_observations=4.0)
    fused, plain = RunningMeanStd(cfg.num_obs, device=DEV), RunningMeanStd(cfg.num_obs, device=DEV)
    ref = P.RunningMeanStd(cfg.num_obs)
    for step in range(3):
        task.post_physics_step()
        rows = task.obs_buf.view(-1, cfg.num_obs)
        fused.update_from_step(task.obs_moments, rows.shape[0])
        plain.update(rows)
        ref.update(rows.cpu())
        assert float(task.obs_moments.abs().sum()) == 0.0
        for a, b in ((fused.running_mean, plain.running_mean), (fused.running_var, plain.running_var)):
            torch.testing.assert_close(a, b, rtol=1e-11, atol=1e-12)
        torch.testing.assert_close(fused.running_mean.cpu(), ref.running_mean, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(fused.running_var.cpu(), ref.running_var, rtol=1e-6, atol=1e-7)
        assert float(fused.count) == float(ref.count)
        task.root_states[:, cfg.ball_actor, 0:3] += 0.01          # the next step sees other observations
    assert float(task.obs_buf.abs().max()) <= 4.0
