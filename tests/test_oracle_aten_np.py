"""Pins the numpy restatement of the ATen primitives (oracle/aten_np.py) to torch's own CPU kernels:
the oracle's arithmetic is a third-party dependency of the reference (torch), so its published
algorithms are restated independently and checked here, op by op, in float64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import aten_np as A

TOL = 1e-10


def _t(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("pad", ["zeros", "border", "reflection"])
def test_grid_sample(d, mode, pad):
    rng = np.random.default_rng(d * 10 + len(mode) + len(pad))
    sp = (6, 7) if d == 2 else (4, 5, 6)
    out_sp = (5, 4) if d == 2 else (3, 4, 5)
    inp = rng.standard_normal((2, 3) + sp)
    grid = rng.uniform(-1.4, 1.4, (2,) + out_sp + (d,))
    if mode == "nearest":       # keep clear of the half-integer ties (rint vs nearbyint agree anyway)
        grid = np.round(grid * 7.3) / 7.3 + 1e-3
    ref = F.grid_sample(_t(inp), _t(grid), mode=mode, padding_mode=pad, align_corners=True).numpy()
    got = A.grid_sample(inp, grid, mode=mode, padding_mode=pad)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("pad", ["zeros", "border", "reflection"])
def test_grid_sample_bicubic(pad):
    rng = np.random.default_rng(5)
    inp = rng.standard_normal((2, 2, 6, 7))
    grid = rng.uniform(-1.3, 1.3, (2, 5, 4, 2))
    ref = F.grid_sample(_t(inp), _t(grid), mode="bicubic", padding_mode=pad, align_corners=True).numpy()
    got = A.grid_sample(inp, grid, mode="bicubic", padding_mode=pad)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("size", [[2, 1, 5, 7], [2, 1, 4, 5, 6], [1, 1, 1, 9]])
def test_affine_grid(size):
    rng = np.random.default_rng(len(size))
    d = len(size) - 2
    theta = rng.standard_normal((size[0], d, d + 1))
    ref = F.affine_grid(_t(theta), size, align_corners=True).numpy()
    got = A.affine_grid(theta, size)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("in_sp,out_sp", [((4, 4), (16, 24)), ((8, 8, 4), (32, 32, 16)), ((3, 5), (7, 11))])
def test_interpolate_linear_by_size(in_sp, out_sp):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3) + in_sp)
    mode = "bilinear" if len(in_sp) == 2 else "trilinear"
    ref = F.interpolate(_t(x), size=out_sp, mode=mode, align_corners=False).numpy()
    got = A.interpolate_linear(x, out_sp)
    assert np.abs(got - ref).max() < TOL


def test_interpolate_linear_by_scale_factor():
    """adv_bias.py:325-327 passes scale_factor in 3-D: ATen then uses 1/scale_factor as the ratio."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((1, 1, 5, 6, 3))
    sf = (4.0, 4.0, 4.0)
    ref = F.interpolate(_t(x), scale_factor=sf, mode="trilinear", align_corners=False).numpy()
    got = A.interpolate_linear(x, ref.shape[2:], scales=sf)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("sp", [(9, 11), (6, 7, 8)])
def test_depthwise_gaussian(sp):
    """The reference's 9^d Gaussian (sigma 1) is the outer product of its 1-D factor up to ~1e-8."""
    from oracle import advchain_oracle as orc
    rng = np.random.default_rng(6)
    d = len(sp)
    x = rng.standard_normal((2, d) + sp).astype(np.float32)
    ref = orc.gaussian_smooth(torch.from_numpy(x)).numpy()
    k1 = orc.gaussian_kernel_1d() if hasattr(orc, "gaussian_kernel_1d") else None
    if k1 is None:
        ks = 9
        ax = np.arange(ks, dtype=np.float64) - (ks - 1) / 2.0
        k1 = np.exp(-ax ** 2 / 2.0)
        k1 = k1 / k1.sum()
    got = A.depthwise_separable_conv(x.astype(np.float64), list(np.asarray(k1, dtype=np.float64)))
    assert np.abs(got - ref).max() < 5e-6


def _gauss1d():
    ax = np.arange(9, dtype=np.float64) - 4.0
    k = np.exp(-ax ** 2 / 2.0)
    return list(k / k.sum())


def _base_grid_np(n, spatial):
    mesh = np.meshgrid(*[np.linspace(-1.0, 1.0, s) for s in spatial], indexing="ij")
    d = len(spatial)
    return np.broadcast_to(np.stack([mesh[d - 1 - k] for k in range(d)], 0), (n, d) + tuple(spatial)).copy()


def _cl(x):           # channel-first N x d x spatial -> channel-last grid
    return np.moveaxis(x, 1, -1)


@pytest.mark.parametrize("spatial,vsize", [((24, 32), (3, 4)), ((12, 16, 20), (2, 2, 3))])
def test_field_build_restated_without_torch(spatial, vsize):
    """The whole deformation-field build (adv_morph.py:454-491: low-res Gaussian, upsample, scaling and
    squaring with quirk Q1, compose-with-base, full-res Gaussian, clamp) restated on the numpy
    primitives in float64 agrees with the oracle's fp32 torch evaluation to fp32 rounding."""
    from oracle import advchain_oracle as orc
    rng = np.random.default_rng(11)
    d = len(spatial)
    v = rng.uniform(-1, 1, (2, d) + vsize)
    v = v / np.sqrt((v.reshape(2, -1) ** 2).sum(1)).reshape(2, *([1] * (d + 1)))
    k1 = _gauss1d()
    u = A.depthwise_separable_conv(1.5 * v, k1)
    u = A.interpolate_linear(u, spatial)
    n = 8
    if d == 3:
        while np.sqrt(((u / 2.0 ** n) ** 2).sum()) > 0.5:
            n += 1
    base = _base_grid_np(2, spatial)
    phi0 = base + u / 2.0 ** n
    phi = phi0
    for _ in range(n):
        phi = A.grid_sample(phi, _cl(phi), padding_mode="border")
    off = phi - phi0                                       # quirk Q1: minus phi_0, not minus base
    comp = A.grid_sample(base, _cl(off + base), padding_mode="border")
    comp = A.depthwise_separable_conv(comp - base, k1) + base
    want = np.clip(comp, -1, 1)
    got = orc.morph_field(torch.from_numpy(v.astype(np.float32)), 1.5, spatial).numpy()
    assert np.abs(got - want).max() < 2e-5
