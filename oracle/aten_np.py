"""numpy restatement of the ATen primitives the reference's hot path is made of.  TEST INFRASTRUCTURE ONLY.

The reference (`advchain/augmentor/*.py`) does its arithmetic through PyTorch ATen, a third-party
dependency (requirements.txt:4 `torch>=1.6.0`).  `oracle/advchain_oracle.py` calls the same ATen ops;
this module restates the published algorithms of those ops in plain numpy (float64 by default) so that
the oracle does not rest on torch alone: `tests/test_oracle_aten_np.py` checks every function here
against torch's CPU kernels, and the CUDA kernels follow the same formulas.

Sources (PyTorch 2.x, aten/src/ATen/native):
  * GridSampler.h : grid_sampler_unnormalize, clip_coordinates, reflect_coordinates,
                    compute_coordinates, cubic convolution coefficients (A = -0.75)
  * GridSampler.cpp / cuda/GridSampler.cu : grid_sampler_{2,3}d (bilinear / nearest / bicubic)
  * AffineGridGenerator.cpp : affine_grid (align_corners=True: linspace(-1, 1, size))
  * UpSample.h : area_pixel_compute_source_index (align_corners=False), linear upsampling
All functions here use align_corners=True for grid_sample / affine_grid and align_corners=False for
interpolate, which is what the reference passes (adv_affine.py:297-313, adv_morph.py:133-135, 462-464,
546-557, adv_bias.py:318-327).
"""
import numpy as np


# ----------------------------------------------------------------------------- coordinates

def unnormalize(coord, size):
    """grid_sampler_unnormalize, align_corners=True: [-1, 1] -> [0, size-1]."""
    return (coord + 1.0) / 2.0 * (size - 1)


def reflect(x, twice_low, twice_high):
    """reflect_coordinates (bounds given doubled, like ATen)."""
    if twice_low == twice_high:
        return np.zeros_like(x)
    mn = twice_low / 2.0
    span = (twice_high - twice_low) / 2.0
    x = np.abs(x - mn)
    extra = np.fmod(x, span)
    flips = np.floor(x / span)
    return np.where(flips % 2 == 0, extra + mn, span - extra + mn)


def compute_coordinates(x, size, padding_mode):
    """compute_coordinates on pixel coordinates (align_corners=True)."""
    if padding_mode == "border":
        x = np.clip(x, 0, size - 1)
    elif padding_mode == "reflection":
        x = reflect(x, 0, 2 * (size - 1))
        x = np.clip(x, 0, size - 1)
    return x


def source_index(coord, size, padding_mode):
    return compute_coordinates(unnormalize(coord, size), size, padding_mode)


# ----------------------------------------------------------------------------- grid_sample

def _gather(inp, idx, valid):
    """inp: (N, C, *S); idx: tuple of integer index arrays (N, *out) per spatial axis (slow -> fast);
    valid: bool (N, *out).  Out-of-bounds taps read 0."""
    n, c = inp.shape[:2]
    sp = inp.shape[2:]
    safe = [np.clip(i, 0, s - 1) for i, s in zip(idx, sp)]
    nn = np.arange(n).reshape((n,) + (1,) * (idx[0].ndim - 1))
    out = inp[(nn, slice(None)) + tuple(safe)]            # (N, *out, C): advanced indices first
    out = np.moveaxis(out, -1, 1)
    return out * valid[:, None]


def grid_sample(inp, grid, mode="bilinear", padding_mode="zeros"):
    """F.grid_sample(inp, grid, mode, padding_mode, align_corners=True).
    inp (N, C, H, W) with grid (N, Ho, Wo, 2), or inp (N, C, D, H, W) with grid (N, Do, Ho, Wo, 3);
    grid[..., 0] addresses the LAST spatial axis (x)."""
    inp = np.asarray(inp)
    grid = np.asarray(grid)
    d = inp.ndim - 2
    sizes = inp.shape[2:]                                   # slow -> fast
    # coordinate per tensor axis, slow -> fast: grid channel d-1-a addresses axis a
    if mode == "bicubic":
        assert d == 2, "bicubic is 2-D only"
        return _bicubic(inp, grid, padding_mode)
    pix = [source_index(grid[..., d - 1 - a], sizes[a], padding_mode) for a in range(d)]
    if mode == "nearest":
        idx = [np.rint(p).astype(np.int64) for p in pix]   # nearbyint: round half to even
        valid = np.ones(idx[0].shape, bool)
        for i, s in zip(idx, sizes):
            valid &= (i >= 0) & (i < s)
        return _gather(inp, idx, valid)
    lo = [np.floor(p) for p in pix]
    out = 0.0
    for corner in range(1 << d):
        w = 1.0
        idx = []
        valid = np.ones(pix[0].shape, bool)
        for a in range(d):
            up = (corner >> a) & 1
            i = lo[a] + up
            # ATen: weight of the low corner is (hi - x), of the high corner (x - lo)
            w = w * ((pix[a] - lo[a]) if up else ((lo[a] + 1.0) - pix[a]))
            ii = i.astype(np.int64)
            valid &= (ii >= 0) & (ii < sizes[a])
            idx.append(ii)
        out = out + _gather(inp, idx, valid) * w[:, None]
    return out


_A = -0.75


def _cc1(x):
    return ((_A + 2) * x - (_A + 3)) * x * x + 1


def _cc2(x):
    return ((_A * x - 5 * _A) * x + 8 * _A) * x - 4 * _A


def cubic_coefficients(t):
    """get_cubic_upsampling_coefficients."""
    return [_cc2(t + 1.0), _cc1(t), _cc1(1.0 - t), _cc2(2.0 - t)]


def _bicubic(inp, grid, padding_mode):
    h, w = inp.shape[2:]
    ix = unnormalize(grid[..., 0], w)
    iy = unnormalize(grid[..., 1], h)
    x0, y0 = np.floor(ix), np.floor(iy)
    cx, cy = cubic_coefficients(ix - x0), cubic_coefficients(iy - y0)
    out = 0.0
    for j in range(4):
        # get_value_bounded: the INTEGER tap coordinate goes through compute_coordinates, then a bounds test
        ty = compute_coordinates(y0 - 1 + j, h, padding_mode).astype(np.int64)
        row = 0.0
        for i in range(4):
            tx = compute_coordinates(x0 - 1 + i, w, padding_mode).astype(np.int64)
            valid = (tx >= 0) & (tx < w) & (ty >= 0) & (ty < h)
            row = row + _gather(inp, [ty, tx], valid) * cx[i][:, None]
        out = out + row * cy[j][:, None]
    return out


# ----------------------------------------------------------------------------- affine_grid

def linspace_from_neg_one(size):
    """AffineGridGenerator.cpp, align_corners=True; a single point sits at 0."""
    if size <= 1:
        return np.zeros(size)
    return np.linspace(-1.0, 1.0, size)


def affine_grid(theta, size):
    """F.affine_grid(theta, size, align_corners=True).  theta (N, d, d+1); size (N, C, *spatial).
    Returns (N, *spatial, d) with channel 0 = x."""
    theta = np.asarray(theta)
    sp = size[2:]
    d = len(sp)
    axes = [linspace_from_neg_one(s) for s in sp]
    mesh = np.meshgrid(*axes, indexing="ij")                # slow -> fast
    base = np.stack([mesh[d - 1 - k] for k in range(d)] + [np.ones(sp)], -1)      # (..., d+1): x, y[, z], 1
    return np.einsum("...k,nck->n...c", base, theta)


# ----------------------------------------------------------------------------- interpolate

def upsample_axis_weights(in_size, out_size, scale=None):
    """area_pixel_compute_source_index(align_corners=False) + linear weights for one axis.
    `scale` = the user-given scale_factor (ATen then uses 1/scale_factor), else in/out."""
    ratio = (1.0 / scale) if scale is not None else float(in_size) / float(out_size)
    dst = np.arange(out_size)
    src = np.maximum(ratio * (dst + 0.5) - 0.5, 0.0)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = src - i0
    return i0, i1, 1.0 - l1, l1


def interpolate_linear(x, out_sizes, scales=None):
    """F.interpolate(x, size=..., mode='bilinear'|'trilinear', align_corners=False) (scales: the
    scale_factor per axis when the caller passed scale_factor instead of size, adv_bias.py:325-327)."""
    x = np.asarray(x)
    for a, o in enumerate(out_sizes):
        ax = 2 + a
        i0, i1, l0, l1 = upsample_axis_weights(x.shape[ax], o, None if scales is None else scales[a])
        shape = [1] * x.ndim
        shape[ax] = o
        x = np.take(x, i0, axis=ax) * l0.reshape(shape) + np.take(x, i1, axis=ax) * l1.reshape(shape)
    return x


# ----------------------------------------------------------------------------- depthwise conv

def depthwise_separable_conv(x, k1d):
    """nn.ConvNd(C, C, ks, groups=C, padding=ks//2, bias=False) with the outer-product kernel
    k1d x k1d [x k1d] applied to every channel (adv_morph.py:377-452; zero padding)."""
    x = np.asarray(x)
    r = len(k1d) // 2
    for ax in range(2, x.ndim):
        pad = [(0, 0)] * x.ndim
        pad[ax] = (r, r)
        xp = np.pad(x, pad)
        acc = 0.0
        for t, wt in enumerate(k1d):
            sl = [slice(None)] * x.ndim
            sl[ax] = slice(t, t + x.shape[ax])
            acc = acc + wt * xp[tuple(sl)]
        x = acc
    return x
