"""Lane-level numpy models of `ss_step_bwd_lean_kernel` (the default squaring-step adjoint) and
`ss_step_bwd_box_kernel` (advchain_b200/csrc/advk_morph.cu).

The lean kernel hands the x1 corner column of a lane to its neighbour by WEIGHT (the receiver forms
g*w0 + g_prev*w1_prev), redirects corners outside the volume to inside ones (their weight is exactly 0
under border padding) and builds the Jacobian term by separable differences.  The warp-box variant
hands contributions from lane to lane along x, y and z before it issues a RED, and lets the Jacobian
term ride on the corner that lands on the voxel itself.  Whether those hand-offs add up to the plain
scatter for ANY field (ragged boxes at the
volume faces, displacements of many voxels, coordinates clipped by the border padding) is index
logic that can be checked without a GPU: this test executes the kernel's statements for 32 lanes at a
time, with CUDA's shuffle semantics, and compares with the direct adjoint
(grid_sample(phi, phi, border, align_corners=True) backward, adv_morph.py:133-135 / 166-168).
The GPU parity of the real kernels is in tests/test_gpu_kernels.py::test_morph_field (tile masks >= 16).
"""
import numpy as np
import pytest


def _axis(coord, size):
    """make_axis_border of advk_common.cuh, float64."""
    mx = float(size - 1)
    x = (coord + 1.0) / 2.0 * mx
    inside = (x > 0) & (x < mx)
    x = np.minimum(mx, np.maximum(x, 0.0))
    mult = np.where(inside, mx / 2.0, 0.0)
    f = np.floor(x)
    i0 = f.astype(np.int64)
    return i0, (f + 1.0) - x, x - f, i0 + 1 < size, mult


def _direct(phi, up):
    """out[y] = sum_x up[x] * w(phi[x], y)  +  Jacobian term at x.  phi, up: [D, H, W, 3]."""
    D, H, W, _ = phi.shape
    out = np.zeros_like(up)
    for z in range(D):
        for y in range(H):
            for x in range(W):
                f, g = phi[z, y, x], up[z, y, x]
                ix, wx0, wx1, vx1, mx = _axis(f[0], W)
                iy, wy0, wy1, vy1, my = _axis(f[1], H)
                iz, wz0, wz1, vz1, mz = _axis(f[2], D)
                j = np.zeros(3)
                for dz in range(2):
                    for dy in range(2):
                        for dx in range(2):
                            if (dz and not vz1) or (dy and not vy1) or (dx and not vx1):
                                continue
                            wz, wy, wx = (wz1 if dz else wz0), (wy1 if dy else wy0), (wx1 if dx else wx0)
                            out[iz + dz, iy + dy, ix + dx] += g * (wx * wy * wz)
                            dot = float(phi[iz + dz, iy + dy, ix + dx] @ g)
                            j[0] += (dot if dx else -dot) * wy * wz
                            j[1] += (dot if dy else -dot) * wx * wz
                            j[2] += (dot if dz else -dot) * wx * wy
                out[z, y, x] += j * np.array([mx, my, mz])
    return out


def _shfl_down(v, d):
    r = v.copy()
    r[:32 - d] = v[d:]
    return r


def _shfl_up(v, d):
    r = v.copy()
    r[d:] = v[:32 - d]
    return r


def _box_kernel(phi, up, lbx, lby, lbz):
    """The kernel, one warp (= one box) at a time; returns (out, number of RED lane-operations)."""
    D, H, W, _ = phi.shape
    BX, BY, BZ = 1 << lbx, 1 << lby, 1 << lbz
    HW = H * W
    src = phi.reshape(-1, 3)
    out = np.zeros((D * H * W, 3))
    reds = 0
    lane = np.arange(32)
    lx, ly, lz = lane & (BX - 1), (lane >> lbx) & (BY - 1), lane >> (lbx + lby)

    def red(mask, addr, val):
        nonlocal reds
        for i in np.nonzero(mask)[0]:
            out[addr[i]] += val[i]
            reds += 1

    for bz in range(-(-D // BZ)):
        for by in range(-(-H // BY)):
            for bx in range(-(-W // BX)):
                x, y, z = bx * BX + lx, by * BY + ly, bz * BZ + lz
                live = (x < W) & (y < H) & (z < D)
                p = (z * H + y) * W + x
                ps = np.where(live, p, 0)
                f = np.where(live[:, None], src[ps], 0.0)
                g = np.where(live[:, None], up.reshape(-1, 3)[ps], 0.0)
                ix, wx0, wx1, vx1, mx = _axis(f[:, 0], W)
                iy, wy0, wy1, vy1, my = _axis(f[:, 1], H)
                iz, wz0, wz1, vz1, mz = _axis(f[:, 2], D)
                a000 = iz * HW + iy * W + ix
                # pass 1
                j = np.zeros((32, 3))
                for dz in range(2):
                    for dy in range(2):
                        row = live & (vy1 if dy else True) & (vz1 if dz else True)
                        for dx in range(2):
                            m = row & (vx1 if dx else True)
                            a = np.where(m, a000 + dz * HW + dy * W + dx, 0)
                            dot = np.where(m, (src[a] * g).sum(1), 0.0)
                            wz, wy, wx = (wz1 if dz else wz0), (wy1 if dy else wy0), (wx1 if dx else wx0)
                            j[:, 0] += (dot if dx else -dot) * wy * wz
                            j[:, 1] += (dot if dy else -dot) * wx * wz
                            j[:, 2] += (dot if dz else -dot) * wx * wy
                j *= np.stack([mx, my, mz], 1)
                # pass 2
                key = np.where(live, a000, -1)
                hand_x = live & vx1 & (lx < BX - 1) & (_shfl_down(key, 1) == a000 + 1)
                gp = _shfl_up(g, 1)
                ddy, ddz = y - iy, z - iz
                frow = np.where(live & (ix == x) & (ddy >= 0) & (ddy < 2) & (ddz >= 0) & (ddz < 2), ddz * 2 + ddy, -1)
                R = np.zeros((2, 2, 32, 3))
                for dz in range(2):
                    for dy in range(2):
                        row = live & (vy1 if dy else True) & (vz1 if dz else True)
                        wr = (wy1 if dy else wy0) * (wz1 if dz else wz0)
                        w0, w1 = wx0 * wr, wx1 * wr
                        ws = _shfl_up(np.where(hand_x & row, w1, 0.0), 1)
                        ws[0] = 0.0
                        c0 = g * w0[:, None] + gp * ws[:, None]
                        c0 = c0 + np.where((frow == dz * 2 + dy)[:, None], j, 0.0)
                        red(row & vx1 & ~hand_x, a000 + dz * HW + dy * W + 1, g * w1[:, None])
                        R[dz, dy] = c0
                hand_y = np.zeros(32, bool)
                if lby > 0:
                    hand_y = live & vy1 & (ly < BY - 1) & (_shfl_down(key, BX) == a000 + W)
                for dz in range(2):
                    rv1 = live & vy1 & (vz1 if dz else True)
                    if lby > 0:
                        h = hand_y & rv1
                        r = _shfl_up(np.where(h[:, None], R[dz, 1], 0.0), BX)
                        R[dz, 0] += np.where((ly > 0)[:, None], r, 0.0)
                    red(rv1 & ~hand_y, a000 + dz * HW + W, R[dz, 1])
                rv1 = live & vz1
                hand_z = np.zeros(32, bool)
                if lbz > 0:
                    hand_z = rv1 & (lz < BZ - 1) & (_shfl_down(key, BX * BY) == a000 + HW)
                    r = _shfl_up(np.where(hand_z[:, None], R[1, 0], 0.0), BX * BY)
                    R[0, 0] += np.where((lz > 0)[:, None], r, 0.0)
                red(rv1 & ~hand_z, a000 + HW, R[1, 0])
                red(live, a000, R[0, 0])
                red(live & (frow < 0), p, j)
    return out.reshape(D, H, W, 3), reds


def _lean_kernel(phi, up):
    """`ss_step_bwd_lean_kernel`: 32 consecutive voxels per warp, x hand-off by weight, outside corners
    redirected to corner 0 of their axis (zero weight / zero mult), Jacobian by separable differences."""
    D, H, W, _ = phi.shape
    HW, S = H * W, D * H * W
    src, upf = phi.reshape(-1, 3), up.reshape(-1, 3)
    out = np.zeros((S, 3))
    lane = np.arange(32)
    for w in range(-(-S // 32)):
        p = w * 32 + lane
        live = p < S
        ps = np.where(live, p, 0)
        f = np.where(live[:, None], src[ps], 0.0)
        g = np.where(live[:, None], upf[ps], 0.0)
        ix, wx0, wx1, vx1, mx = _axis(f[:, 0], W)
        iy, wy0, wy1, vy1, my = _axis(f[:, 1], H)
        iz, wz0, wz1, vz1, mz = _axis(f[:, 2], D)
        dxo, dyo, dzo = np.where(vx1, 1, 0), np.where(vy1, W, 0), np.where(vz1, HW, 0)
        a000 = iz * HW + iy * W + ix
        d = np.zeros((2, 2, 2, 32))
        for dz in range(2):
            for dy in range(2):
                r = a000 + dz * dzo + dy * dyo
                d[dz, dy, 0] = (src[r] * g).sum(1)
                d[dz, dy, 1] = (src[r + dxo] * g).sum(1)
        jxz, eyd, ey = [], [], []
        for dz in range(2):
            e0 = wx0 * d[dz, 0, 0] + wx1 * d[dz, 0, 1]
            e1 = wx0 * d[dz, 1, 0] + wx1 * d[dz, 1, 1]
            jxz.append(wy0 * (d[dz, 0, 1] - d[dz, 0, 0]) + wy1 * (d[dz, 1, 1] - d[dz, 1, 0]))
            eyd.append(e1 - e0)
            ey.append(wy0 * e0 + wy1 * e1)
        j = np.stack([(wz0 * jxz[0] + wz1 * jxz[1]) * mx, (wz0 * eyd[0] + wz1 * eyd[1]) * my, (ey[1] - ey[0]) * mz], 1)
        nxt = _shfl_down(np.where(live, a000, -1), 1)
        hand = live & vx1 & (lane < 31) & (nxt == a000 + 1)
        gp = _shfl_up(g, 1)
        for dz in range(2):
            for dy in range(2):
                wr = (wy1 if dy else wy0) * (wz1 if dz else wz0)
                w0, w1 = wx0 * wr, wx1 * wr
                ws = _shfl_up(np.where(hand, w1, 0.0), 1)
                ws[0] = 0.0
                c0 = g * w0[:, None] + gp * ws[:, None]
                r = a000 + dz * dzo + dy * dyo
                for i in np.nonzero(live)[0]:
                    out[r[i]] += c0[i]
                for i in np.nonzero(live & vx1 & ~hand)[0]:
                    out[r[i] + 1] += g[i] * w1[i]
        for i in np.nonzero(live)[0]:
            out[p[i]] += j[i]
    return out.reshape(D, H, W, 3)


def _field(rng, D, H, W, amp_vox):
    """Absolute sampling coordinates base + displacement (normalised), smooth-ish + a few outliers."""
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    base = np.stack([xx, yy, zz], -1)
    disp = np.stack([np.sin(3 * xx + 2 * yy + zz), np.cos(2 * xx - yy + 3 * zz), np.sin(xx + 4 * yy - 2 * zz)], -1)
    scale = amp_vox * 2.0 / np.array([max(W - 1, 1), max(H - 1, 1), max(D - 1, 1)])
    phi = base + disp * scale
    k = rng.integers(0, D * H * W, size=6)
    phi.reshape(-1, 3)[k] += rng.uniform(-1.5, 1.5, size=(6, 3))      # far jumps and out-of-range coordinates
    return phi


SHAPES = [(5, 0, 0), (4, 1, 0), (3, 2, 0), (3, 1, 1)]


@pytest.mark.parametrize("lbx,lby,lbz", SHAPES)
@pytest.mark.parametrize("size,amp", [((5, 7, 19), 0.4), ((4, 6, 9), 2.7), ((1, 9, 21), 0.6)])
def test_box_handoff_equals_plain_scatter(lbx, lby, lbz, size, amp):
    rng = np.random.default_rng(11)
    D, H, W = size
    phi = _field(rng, D, H, W, amp)
    up = rng.standard_normal((D, H, W, 3))
    ref = _direct(phi, up)
    out, reds = _box_kernel(phi, up, lbx, lby, lbz)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("size,amp", [((5, 7, 19), 0.4), ((4, 6, 9), 2.7), ((1, 9, 21), 0.6), ((3, 5, 1), 0.8)])
def test_lean_kernel_equals_plain_scatter(size, amp):
    """Coordinates ON and beyond the last voxel of every axis are in the field (clipped by the border
    padding): the redirected corners must contribute nothing."""
    rng = np.random.default_rng(12)
    D, H, W = size
    phi = _field(rng, D, H, W, amp)
    phi[..., -1, 0] = 1.0            # x exactly on the last voxel
    phi[:, -1, :, 1] = 1.3           # y clipped
    phi[0, :, :, 2] = -1.0           # z exactly on the first voxel
    up = rng.standard_normal((D, H, W, 3))
    ref = _direct(phi, up)
    out = _lean_kernel(phi, up)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_box_red_count():
    """Sub-voxel smooth field on a volume the boxes tile exactly: the RED lane-operations per voxel the
    kernel's header quotes (5.1 for the lane-combined kernel = 4 rows + 1 Jacobian + row ends)."""
    rng = np.random.default_rng(5)
    D, H, W = 8, 16, 32
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    phi = np.stack([xx + 0.3 * 2 / (W - 1), yy + 0.4 * 2 / (H - 1), zz + 0.2 * 2 / (D - 1)], -1)   # all displacements in [0, 1) voxel
    up = rng.standard_normal((D, H, W, 3))
    per_voxel = {}
    for s in SHAPES:
        _, reds = _box_kernel(phi, up, *s)
        per_voxel[s] = reds / float(D * H * W)
    assert per_voxel[(5, 0, 0)] < 4.2
    assert per_voxel[(4, 1, 0)] < 3.4
    assert per_voxel[(3, 2, 0)] < 3.1
    assert per_voxel[(3, 1, 1)] < 3.1
