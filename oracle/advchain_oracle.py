"""CPU oracle for the advchain chained-augmentation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (plain PyTorch fp32 ops + autograd) of the
algorithm the reference implements in `advchain/augmentor/*.py` and the part of
`advchain/common/loss.py` the inner loop calls.  It exists so that the CUDA product path in
`advchain_b200/` can be checked against something that runs on the GPU box, where
`/root/reference` does not exist.

Rules (see DESIGN.md "Oracle"):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    leg may import this module; nothing under `advchain_b200/` does.
  * parity pinning: the reference ships no tests or golden vectors, so the
    oracle is pinned against outputs of the *unmodified reference itself*, generated in the
    build container by `tests/golden/make_golden.py` and committed under `tests/golden/*.npz`
    (`tests/test_oracle_golden.py` replays them).
  * the arithmetic underneath (grid_sample, affine_grid, interpolate, conv*) is PyTorch ATen,
    a third-party dependency of the reference (requirements.txt:4 `torch>=1.6.0`; run here
    with torch 2.11.0).  `oracle/aten_np.py` restates those primitives in numpy.

Each function cites the reference lines it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- helpers


def unit_l2(d):
    """Per-sample L2 normalisation. adv_transformation_base.py:151-156."""
    nrm = d.reshape(d.shape[0], -1).norm(dim=1).view(-1, *([1] * (d.dim() - 1)))
    return d / (nrm + 1e-20)


def base_grid(n, spatial):
    """Identity sampling grid, channel-first, channel 0 = x = LAST tensor axis.
    adv_morph.py:14-55 (meshgrid of linspace(-1,1,size), 'ij', concatenated x,y[,z])."""
    mesh = torch.meshgrid([torch.linspace(-1, 1, s) for s in spatial], indexing="ij")
    g = torch.stack(mesh[::-1], 0)
    return g.unsqueeze(0).repeat(n, *([1] * (len(spatial) + 1)))


def _cl(field):
    """channel-first N x d x spatial -> channel-last grid for grid_sample."""
    d = field.dim() - 2
    return field.permute(0, *range(2, 2 + d), 1)


def gaussian_kernel(d, sigma=1.0, ksize=5):
    """adv_morph.py:391-421: size is raised to 2*int(4*sigma+0.5)+1 (=9 for sigma 1),
    kernel = exp(-|r|^2 / (2 sigma^2)) normalised by its total sum."""
    ksize = max(ksize, 2 * int(4 * sigma + 0.5) + 1)
    ax = torch.arange(ksize).to(torch.get_default_dtype()) - (ksize - 1) / 2.0
    sq = sum(m ** 2.0 for m in torch.meshgrid(*([ax] * d), indexing="ij"))
    k = torch.exp(-sq / (2 * sigma ** 2.0))
    return k / k.sum()


def gaussian_smooth(x, sigma=1.0):
    """Depthwise Gaussian, zero padding. adv_morph.py:377-389, 423-445."""
    d = x.dim() - 2
    k = gaussian_kernel(d, sigma)
    w = k.view(1, 1, *k.shape).repeat(x.shape[1], 1, *([1] * d))
    conv = F.conv2d if d == 2 else F.conv3d
    return conv(x, w, padding=k.shape[0] // 2, groups=x.shape[1])


def ss_steps_for(u, n=8):
    """3-D only: smallest n>=8 with ||u / 2^n||_F <= 0.5 over the whole batch tensor.
    adv_morph.py:159-162."""
    if u.dim() == 5:
        while torch.norm(u / (2.0 ** n)) > 0.5:
            n += 1
    return n


def integrate_velocity(u, n=8):
    """Scaling and squaring. adv_morph.py:116-177, 179-202.
    Quirk Q1: `integrate_by_add` mutates the base grid in place, so the value
    subtracted at the end is phi_0 = base + u/2^n, not base."""
    n = ss_steps_for(u.detach(), n)
    phi0 = base_grid(u.shape[0], u.shape[2:]) + u / (2.0 ** n)
    phi = phi0
    for _ in range(n):
        phi = F.grid_sample(phi, _cl(phi), padding_mode="border", align_corners=True)
    return phi - phi0, n


def morph_field(v, scale, spatial):
    """`DemonsCompose(duv=scale*v, base, smooth=True)`. adv_morph.py:454-491."""
    d = len(spatial)
    u = gaussian_smooth(scale * v)
    u = F.interpolate(u, size=tuple(spatial), mode="bilinear" if d == 2 else "trilinear",
                      align_corners=False)
    off, _ = integrate_velocity(u)
    base = base_grid(v.shape[0], spatial)
    comp = F.grid_sample(base, _cl(off + base), padding_mode="border", align_corners=True)
    comp = gaussian_smooth(comp - base) + base
    return torch.clamp(comp, -1, 1)


def warp(data, grid_cf=None, theta=None, interp="bilinear", padding="zeros"):
    """grid_sample wrapper with the reference's padding options.
    adv_morph.py:524-558 (field) / adv_affine.py:289-314 (theta).
    Deviation Q13: 'lowest' uses the per-sample minimum broadcast correctly for any N."""
    if theta is not None:
        grid = F.affine_grid(theta, list(data.shape), align_corners=True)
    else:
        grid = _cl(grid_cf)
    if padding == "lowest":
        pv = data.reshape(data.shape[0], -1).min(dim=1).values.detach()
        pv = pv.view(-1, *([1] * (data.dim() - 1)))
    elif isinstance(padding, (int, float)) and not isinstance(padding, bool):
        pv = padding
    else:
        return F.grid_sample(data, grid, mode=interp, padding_mode=padding, align_corners=True)
    return F.grid_sample(data - pv, grid, mode=interp, padding_mode="zeros",
                         align_corners=True) + pv


# ----------------------------------------------------------------------------- bias geometry


def bspline_kernel(stride, order):
    """adv_bias.py:12-49. 2-D: padding grows as i*stride (zero-padded result, quirk Q6);
    3-D: fixed padding stride-1 (tight support)."""
    d = len(stride)
    ones = torch.ones(1, 1, *stride)
    k = ones
    conv = F.conv2d if d == 2 else F.conv3d
    for i in range(1, order + 1):
        pad = [i * s for s in stride] if d == 2 else [s - 1 for s in stride]
        k = conv(k, ones, padding=pad) / float(np.prod(stride))
    return k[0, 0]


class BiasGeometry:
    """Control-point lattice + crop window. adv_bias.py:84-102, 202-277, 358-374."""

    def __init__(self, data_size, spacing, downscale, order):
        img = np.array(data_size[2:])
        self.image_size = img
        self.stride = [s // downscale for s in spacing]
        st = np.array(self.stride)
        lowres = img / (1.0 * downscale)
        cells = np.ceil(lowres / st).astype(int)
        inner = st * cells - (st - 1)
        self.cp_shape = (cells + 2).tolist()
        diff = inner - lowres
        fl = np.floor(np.abs(diff) / 2) * np.sign(diff)
        self.crop_start = (fl + np.remainder(diff, 2) * np.sign(diff)).astype(int)
        self.crop_end = fl.astype(int)
        self.kernel = bspline_kernel(self.stride, order)
        self.pad = ((np.array(self.kernel.shape) - 1) / 2).astype(int).tolist()


def bias_field(cp, g, eps, use_log=True):
    """cp -> conv_transpose -> crop -> linear upsample -> exp -> clip.
    adv_bias.py:279-335 and 337-356."""
    d = len(g.stride)
    convt = F.conv_transpose2d if d == 2 else F.conv_transpose3d
    f = convt(cp, g.kernel[None, None], padding=g.pad, stride=g.stride)
    sl = [slice(None), slice(None)] + [slice(s + a, -s - b) for s, a, b in
                                      zip(g.stride, g.crop_start, g.crop_end)]
    low = f[tuple(sl)]
    sf = [g.image_size[i] / low.shape[2 + i] for i in range(d)]
    up = low
    if any(s > 1 for s in sf):
        if d == 2:
            up = F.interpolate(low, size=(int(g.image_size[0]), int(g.image_size[1])),
                               mode="bilinear", align_corners=False)
        else:
            up = F.interpolate(low, scale_factor=tuple(float(s) for s in sf), mode="trilinear",
                               align_corners=False)
    b = torch.exp(up) if use_log else 1 + up
    return 1 + torch.clamp(b - 1, -eps, eps)


# ----------------------------------------------------------------------------- affine


def affine_theta(param, cfg, d):
    """5 / 9 params -> 2x3 / 3x4 matrix. adv_affine.py:210-273."""
    p = F.hardtanh(param)
    pi = math.pi
    if d == 2:
        ang = p[:, 0] * cfg["rot"] * pi
        a = 1 + p[:, 1] * cfg["scale_x"]
        b = 1 + p[:, 2] * cfg["scale_y"]
        r0 = torch.stack([a * torch.cos(ang), b * (-torch.sin(ang)), p[:, 3] * cfg["shift_x"]], -1)
        r1 = torch.stack([a * torch.sin(ang), b * torch.cos(ang), p[:, 4] * cfg["shift_y"]], -1)
        return torch.stack([r0, r1], 1)
    n = param.shape[0]
    O, I = torch.zeros(n), torch.ones(n)

    def mat(rows):
        return torch.stack([torch.stack(r, -1) for r in rows], 1)

    T = mat([[I, O, O, p[:, 6] * cfg["shift_x"]], [O, I, O, p[:, 7] * cfg["shift_y"]],
             [O, O, I, p[:, 8] * cfg["shift_z"]], [O, O, O, I]])
    S = mat([[1 + p[:, 3] * cfg["scale_x"], O, O, O], [O, 1 + p[:, 4] * cfg["scale_y"], O, O],
             [O, O, 1 + p[:, 5] * cfg["scale_z"], O], [O, O, O, I]])
    ph, th, ps = p[:, 0] * cfg["rot_x"] * pi, p[:, 1] * cfg["rot_y"] * pi, p[:, 2] * cfg["rot_z"] * pi
    c, s = torch.cos, torch.sin
    R = mat([[c(th) * c(ps), -c(ph) * s(ps) + s(ph) * s(th) * c(ps), s(ph) * s(ps) + c(ph) * s(th) * c(ps), O],
             [c(th) * s(ps), c(ph) * c(ps) + s(ph) * s(th) * s(ps), -s(ph) * c(ps) + c(ph) * s(th) * s(ps), O],
             [-s(th), s(ph) * c(th), c(ph) * c(th), O],
             [O, O, O, I]])
    return torch.matmul(T, torch.matmul(R, S))[:, :3, :4]


def affine_inverse(theta):
    """First d rows of inverse([theta; 0..0 1]). adv_affine.py:316-324."""
    n, d = theta.shape[0], theta.shape[1]
    homo = torch.eye(d + 1).unsqueeze(0).repeat(n, 1, 1)
    homo = torch.cat([theta, homo[:, d:, :]], 1)
    return homo.inverse()[:, :d, :]


# ----------------------------------------------------------------------------- loss


def _sobel_kernels(d):
    """common/loss.py:143-203, including quirk Q10 (3-D: gy := gx, gz from the last assignment)."""
    if d == 2:
        kx = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]])
        ky = torch.tensor([[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]])
        return [kx, ky]
    h = torch.tensor([1., 2., 1.])
    hp = torch.tensor([1., 0., -1.])
    gx = h[:, None, None] * hp[None, :, None] * h[None, None, :]
    gz = h[:, None, None] * h[None, :, None] * hp[None, None, :]
    return [gx, gx, gz]


def contour_term(inp, tgt, mask):
    """Single-class Sobel-edge MSE. common/loss.py:102-220 (ignore_background=False,
    one_hot_target=False, one channel at a time)."""
    d = inp.dim() - 2
    conv = F.conv2d if d == 2 else F.conv3d
    m = mask[:, :1]
    tot = 0.
    for k in _sobel_kernels(d):
        w = k[None, None]
        tot = tot + F.mse_loss(conv(inp, w, padding=1) * m, conv(tgt, w, padding=1) * m)
    return tot / len(_sobel_kernels(d))


def consistency_loss(output, reference, types=("mse", "contour"), weights=(1.0, 0.5),
                     mask=None, is_gt=False):
    """common/loss.py:8-87 with scales=[0]; keeps quirk Q9 (mse divided again by N*S)."""
    k = reference.shape[1]
    if mask is None:
        mask = torch.ones_like(output)
    tgt = reference if is_gt else torch.softmax(reference, 1)
    dist = 0.
    for t, w in zip(types, weights):
        if t == "kl":
            p = tgt if not is_gt else torch.where(reference == 0, 1e-8, 1 - 1e-8)
            logp = F.log_softmax(reference, 1) if not is_gt else torch.log(p)
            plogp = (mask * (p * logp)).sum(1)
            plogq = (mask * (p * F.log_softmax(output, 1))).sum(1)
            loss = torch.mean(plogp - plogq)
        elif t == "mse":
            pred = torch.softmax(output, 1)
            loss = F.mse_loss(pred * mask, tgt * mask) / (mask.numel() / k)
        elif t == "contour":
            pred = torch.softmax(output, 1)
            loss = 0.
            for c in range(1, k):
                loss = loss + contour_term(pred[:, [c]], tgt[:, [c]], mask)
            if k > 1:
                loss = loss / (k - 1)
        else:
            raise NotImplementedError(t)
        dist = dist + w * loss
    return dist


# ----------------------------------------------------------------------------- transforms


class _Stage:
    geometric = False
    power_iteration = False

    def __init__(self):
        self.param = None
        self.training = False

    def train(self):
        self.training = True
        p = self.param.detach()
        if self.power_iteration:
            p = self._power_start(p)
        self.param = p.clone().requires_grad_(True)

    def eval(self):
        self.training = False
        self.param = self.param.detach()

    def _power_start(self, p):
        return unit_l2(p)

    def pfwd(self, x):
        return x

    def pbwd(self, x):
        return x

    def update(self, step):
        g = unit_l2(self.param.grad)
        self.param = (g if self.power_iteration else self.param + step * g).detach()


class Noise(_Stage):
    """adv_noise.py:33-114."""
    name = "noise"

    def __init__(self, cfg, ignore_values=None):
        super().__init__()
        self.eps, self.xi, self.size = cfg["epsilon"], cfg["xi"], cfg["data_size"]
        self.ignore = ignore_values

    def init(self):
        self.param = unit_l2(torch.randn(*self.size))

    def fwd(self, x):
        s = self.xi if (self.power_iteration and self.training) else self.eps
        out = x + s * self.param
        if self.ignore is not None:
            out = torch.where((x - self.ignore).abs() < 1e-8, torch.full_like(out, self.ignore), out)
        return out

    def rescale(self):
        self.param = unit_l2(self.param)


class Bias(_Stage):
    """adv_bias.py:84-188."""
    name = "bias"

    def __init__(self, cfg, ignore_values=None):
        super().__init__()
        self.cfg = cfg
        self.eps, self.xi = cfg["epsilon"], 1e-6
        self.use_log = cfg["space"] == "log"
        self.geom = BiasGeometry(cfg["data_size"], cfg["control_point_spacing"], cfg["downscale"],
                                 cfg["interpolation_order"])
        self.ignore = ignore_values
        if self.use_log:
            self.low, self.high = math.log(1 - self.eps), math.log(1 + self.eps)
        else:
            self.low, self.high = -self.eps, self.eps

    def init(self):
        shape = [self.cfg["data_size"][0], 1] + self.geom.cp_shape
        mode = self.cfg["init_mode"]
        if mode == "random":
            self.param = torch.rand(*shape) * (self.high - self.low) + self.low
        elif mode == "gaussian":
            self.param = torch.ones(*shape).normal_(0, 0.5)
            self.low, self.high = -np.inf, np.inf
        else:
            self.param = torch.zeros(*shape)
            self.low, self.high = -np.inf, np.inf

    def field(self):
        cp = self.xi * self.param if (self.power_iteration and self.training) else self.param
        return bias_field(cp, self.geom, self.eps, self.use_log)

    def fwd(self, x):
        b = self.field()
        out = x * b
        if self.ignore is not None:
            out = torch.where((x - self.ignore).abs() < 1e-8, torch.full_like(out, self.ignore), out)
        return out

    def rescale(self):
        self.param = torch.clamp(self.param, self.low, self.high)


class Morph(_Stage):
    """adv_morph.py:247-347, 493-522."""
    name = "morph"
    geometric = True

    def __init__(self, cfg, padding="zeros"):
        super().__init__()
        self.eps, self.xi = cfg["epsilon"], 0.5
        self.size, self.vsize = cfg["data_size"], cfg["vector_size"]
        self.fi = cfg.get("forward_interp", "bilinear")
        self.bi = cfg.get("backward_interp", "bilinear")
        self.padding = padding

    def init(self):
        v = torch.rand(self.size[0], len(self.vsize), *self.vsize) * 2 - 1
        self.param = unit_l2(v)

    def _scale(self):
        return self.xi if (self.power_iteration and self.training) else self.eps

    def field(self, sign):
        return morph_field(self.param, sign * self._scale(), self.size[2:])

    def fwd(self, x):
        return warp(x, torch.clamp(self.field(+1), -1, 1), interp=self.fi, padding=self.padding)

    def bwd(self, x):
        return warp(x, self.field(-1), interp=self.bi, padding=self.padding)

    pfwd, pbwd = fwd, bwd

    def rescale(self):
        self.param = unit_l2(self.param)


class Affine(_Stage):
    """adv_affine.py:73-208."""
    name = "affine"
    geometric = True

    def __init__(self, cfg, padding="zeros"):
        super().__init__()
        self.cfg, self.xi = cfg, 1e-6
        self.d = len(cfg["data_size"]) - 2
        self.fi = cfg.get("forward_interp", "bilinear")
        self.bi = cfg.get("backward_interp", "bilinear")
        self.padding = padding

    def init(self):
        n = self.cfg["data_size"][0]
        self.param = F.hardtanh(2 * torch.rand(n, 5 if self.d == 2 else 9) - 1)

    def _power_start(self, p):
        return p.sign()

    def theta(self):
        p = self.xi * self.param if (self.power_iteration and self.training) else self.param
        return affine_theta(p, self.cfg, self.d)

    def fwd(self, x):
        return warp(x, theta=self.theta(), interp=self.fi, padding=self.padding)

    def bwd(self, x):
        return warp(x, theta=affine_inverse(self.theta()), interp=self.bi, padding=self.padding)

    pfwd, pbwd = fwd, bwd

    def update(self, step):
        g = self.param.grad.sign()
        self.param = (g if self.power_iteration else self.param + step * g).detach()

    def rescale(self):
        pass


def make_stage(name, cfg, **kw):
    return {"noise": Noise, "bias": Bias, "morph": Morph, "affine": Affine}[name](cfg, **kw)


# ----------------------------------------------------------------------------- solver


class Solver:
    """The PGD inner loop and the final loss. adv_compose_solver.py:43-405."""

    def __init__(self, stages, types=("mse", "contour"), weights=(1.0, 0.5), if_norm_image=False,
                 min_intensity=None, max_intensity=None, is_gt=False):
        self.stages = stages
        self.types, self.weights, self.is_gt = types, weights, is_gt
        self.norm, self.lo, self.hi = if_norm_image, min_intensity, max_intensity

    def geometric(self):
        return any(s.geometric for s in self.stages)

    def forward(self, data):
        """:148-176."""
        t = data.detach().clone()
        for s in self.stages:
            t = s.fwd(t)
        if self.norm:
            lo = data.min() if self.lo is None else self.lo
            hi = data.max() if self.hi is None else self.hi
            t = torch.clamp(t, lo, hi)
        return t

    def predict_forward(self, x):
        for s in self.stages:
            x = s.pfwd(x)
        return x

    def predict_backward(self, x):
        for s in reversed(self.stages):
            x = s.pbwd(x)
        return x

    def valid_mask(self, like):
        """:321-325."""
        m = self.predict_backward(self.predict_forward(torch.ones_like(like)))
        return (m != 0).to(like.dtype)

    def loss(self, pred, ref, mask=None):
        return consistency_loss(pred, ref, self.types, self.weights, mask, self.is_gt)

    def step_loss(self, model, data, init_output):
        """One pass of :312-341 up to the scalar `dist` (transforms already in train mode)."""
        out = model(self.forward(data))
        if self.geometric():
            warped = self.predict_backward(out)
            with torch.no_grad():
                mask = self.valid_mask(init_output)
            return self.loss(warped, init_output, mask), warped, mask
        return self.loss(out, init_output.detach()), out, None

    def inner_loop(self, model, data, init_output, n_iter, step_sizes=None, flags=None,
                   record=None):
        """:289-405 without the anatomy-preservation branch."""
        k = len(self.stages)
        flags = [True] * k if flags is None else flags
        steps = [1] * k if step_sizes is None else step_sizes
        for it in range(n_iter):
            model.zero_grad()
            for f, s in zip(flags, self.stages):
                if f:
                    s.train()
            dist, _, _ = self.step_loss(model, data, init_output)
            if not (torch.isnan(dist) or torch.isinf(dist)):
                dist.backward()
                if record is not None:
                    record(it, dist.detach(), [s.param.grad.clone() if f else None
                                               for f, s in zip(flags, self.stages)])
                for f, s in zip(flags, self.stages):
                    if f:
                        # quirk Q15: the reference indexes step_sizes with a counter it never
                        # increments (adv_compose_solver.py:349-357) -> every transform is
                        # updated with step_sizes[0].
                        s.update(steps[0])
            model.zero_grad()
        for f, s in zip(flags, self.stages):
            if f:
                s.rescale()
                s.eval()

    def final_loss(self, model, data, init_output):
        """:236-279 (model.train()/_fix_dropout handling is the caller's business here)."""
        for s in self.stages:
            s.eval()
        adv = self.forward(data)
        out = model(adv.detach().clone())
        if self.geometric():
            mask = self.valid_mask(init_output)
            warped = self.predict_backward(out)
            return self.loss(warped, init_output.detach(), mask), adv, out, warped
        return self.loss(out, init_output.detach()), adv, out, out
