"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the learner-side consumers of obs_buf.

PARITY UNPINNED for `RunningMeanStd`: rl_games is an un-vendored, un-pinned dependency of the reference
(absent from /root/reference and from this image).  What follows restates the published algorithm of
rl_games/algos_torch/running_mean_std.py (class RunningMeanStd: `_update_mean_var_count_from_moments`,
`forward`), which the reference selects with `normalize_input: True`
(cfg/train/HumanoidPingpongTiltG1PPO.yaml:51).  The first layer is pinned against torch itself (present
here): `torch.autocast("cpu", float16)` over `F.elu(F.linear(x, W, b))`, see tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F


class RunningMeanStd:
    def __init__(self, insize, epsilon=1e-05):
        self.epsilon = epsilon
        self.running_mean = torch.zeros(insize, dtype=torch.float64)
        self.running_var = torch.ones(insize, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.float64)
        self.training = True

    @staticmethod
    def _update_mean_var_count_from_moments(mean, var, count, batch_mean, batch_var, batch_count):
        delta = batch_mean - mean
        tot_count = count + batch_count
        new_mean = mean + delta * batch_count / tot_count
        m_a = var * count
        m_b = batch_var * batch_count
        M2 = m_a + m_b + delta ** 2 * count * batch_count / tot_count
        new_var = M2 / tot_count
        return new_mean, new_var, tot_count

    def update(self, x):
        mean = x.mean(0)
        var = x.var(0)                  # unbiased, as torch's default
        self.running_mean, self.running_var, self.count = self._update_mean_var_count_from_moments(
            self.running_mean, self.running_var, self.count, mean, var, x.size()[0])

    def normalize(self, x):
        # the module's text is torch.sqrt(var.float() + eps); torch's CPU float32 sqrt (SLEEF, 0.5+ ulp) is not
        # correctly rounded in ~0.7 % of inputs while the CUDA sqrt the reference runs on is, so the square
        # root is taken in float64 and rounded once (= the correctly rounded float32 result)
        den = torch.sqrt((self.running_var.float() + self.epsilon).double()).float()
        y = (x - self.running_mean.float()) / den
        return torch.clamp(y, min=-5.0, max=5.0)

    def forward(self, x):
        if self.training:
            self.update(x)
        return self.normalize(x)


def clip_observations(obs, clip_obs):
    """VecTask.step: obs_dict["obs"] = clamp(obs_buf, -clip_obs, clip_obs) (upstream; clip_obs defaults to inf)."""
    return obs if clip_obs <= 0 else torch.clamp(obs, -clip_obs, clip_obs)


def first_layer(x_norm, weight, bias, activation="elu"):
    """What autocast(float16) computes for act(linear(x)): fp16 operands, fp32 accumulation, fp16 result of
    the linear layer, activation evaluated in fp32 on that fp16 tensor and rounded to fp16."""
    xh, wh = x_norm.half().float(), weight.half().float()
    y = xh @ wh.t()
    if bias is not None:
        y = y + bias.half().float()
    y = y.half()
    if activation == "elu":
        y = F.elu(y.float()).half()
    return y


def first_layer_fp32(x_norm, weight, bias, activation="elu"):
    """The rollout forward (no autocast): plain fp32 act(linear(x)) as torch computes it on the CPU.  Also returns the
    scale of the rounding noise of the dot product, sum_k |x_k w_k| + |b|, that tolerances are stated against."""
    y = F.linear(x_norm, weight, bias)
    scale = x_norm.abs() @ weight.abs().t()
    if bias is not None:
        scale = scale + bias.abs()
    if activation == "elu":
        y = F.elu(y)
    return y, scale


def first_layer_torch_autocast(x_norm, weight, bias, activation="elu"):
    """The same through torch's own autocast machinery (CPU)."""
    with torch.autocast("cpu", dtype=torch.float16):
        y = F.linear(x_norm, weight, bias)
        if activation == "elu":
            y = F.elu(y)
    return y
