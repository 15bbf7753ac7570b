"""Helpers shared by make_golden_big.py (runs the unmodified reference at BASELINE.json's full sizes) and
the tests that replay those fixtures: seeded regeneration of the large inputs, strided sampling and fp64
checksums of large outputs."""
import numpy as np
import torch


def gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def regen_data(case):
    """The input volume: U[0,1) like the min-max normalised MR inputs."""
    return torch.rand(*case["size"], generator=gen(case["seed"]))


def regen_delta(case):
    """The injected start value of the AdvNoise parameter: unit-L2 per sample Gaussian noise
    (adv_noise.py:41-49 draws the same distribution from the global RNG)."""
    d = torch.randn(*case["size"], generator=gen(case["seed"] + 1))
    n = d.reshape(d.shape[0], -1).norm(dim=1).view(-1, *([1] * (d.dim() - 1)))
    return d / (n + 1e-20)


def _sign_pattern(numel, seed=12345):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, 2, size=numel, dtype=np.int8).astype(np.float64) * 2.0 - 1.0)


def checksum(t):
    """[sum, sum of squares, dot with a fixed +-1 pattern, max |x|] in fp64."""
    x = t.detach().double().cpu().reshape(-1)
    return torch.stack([x.sum(), (x * x).sum(), (x * _sign_pattern(x.numel())).sum(), x.abs().max()])


def strided(t, case):
    """Sample every `stride`-th voxel along each spatial axis, starting at `offset`."""
    s, o = case["stride"], case["offset"]
    idx = (slice(None), slice(None)) + tuple(slice(o, None, s) for _ in t.shape[2:])
    return t.detach()[idx].contiguous()


def checksum_err(t, ck):
    """Largest deviation of the checksums of `t` from the recorded ones, each scaled by what an
    element-wise error of 1 x max|x| would do to it (N for the sums, sqrt(N) for the +-1 dot)."""
    mine = checksum(t)
    n = float(t.numel())
    mx = max(float(ck[3]), 1e-30)
    e_sum = abs(float(mine[0] - ck[0])) / (n * mx)
    e_sq = abs(float(mine[1] - ck[1])) / (2.0 * n * mx * mx)
    e_dot = abs(float(mine[2] - ck[2])) / (n ** 0.5 * mx)
    return max(e_sum, e_sq, e_dot)
