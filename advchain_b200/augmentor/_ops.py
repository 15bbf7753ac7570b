"""torch.autograd.Function wrappers around the advk_* C ABI (include/advk.h).

PyTorch supplies tensors (device memory), the current stream and the autograd tape; all device
arithmetic is in libadvchain_b200.so.  Fields are carried as fp32 tensors of shape
[N, *spatial, 2] (2-D) or [N, *spatial, 4] (3-D, last lane unused) -- the interleaved layout
the kernels gather from with one vector load per stencil corner.
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib
from .._lib import call, ptr, stream

_PAD = {"zeros": _lib.PAD_ZEROS, "border": _lib.PAD_BORDER, "reflection": _lib.PAD_REFLECTION}
_INTERP = {"bilinear": _lib.INTERP_LINEAR, "trilinear": _lib.INTERP_LINEAR,
           "linear": _lib.INTERP_LINEAR, "nearest": _lib.INTERP_NEAREST, "bicubic": _lib.INTERP_BICUBIC}


def parse_interp(interp):
    if interp not in _INTERP:
        raise NotImplementedError("interpolation mode %r is not supported by the CUDA path "
                                  "(bilinear/trilinear/nearest/bicubic are)" % (interp,))
    return _INTERP[interp]


def parse_padding(padding_mode, data):
    """-> (pad_mode, per-sample pad-value tensor or None).
    'lowest' / numeric padding = sample (x - v) with zero padding, add v back
    (adv_affine.py:299-311, adv_morph.py:542-554).  Deviation from the reference (quirk Q13):
    'lowest' uses the per-sample minimum for any batch size; the reference's broadcast only
    works for N == 1."""
    if isinstance(padding_mode, str):
        if padding_mode == "lowest":
            pv = data.detach().reshape(data.shape[0], -1).min(dim=1).values.float().contiguous()
            return _lib.PAD_ZEROS, pv
        if padding_mode not in _PAD:
            raise ValueError("unknown padding mode %r" % (padding_mode,))
        return _PAD[padding_mode], None
    if isinstance(padding_mode, (int, float)) and not isinstance(padding_mode, bool):
        pv = torch.full((data.shape[0],), float(padding_mode), dtype=torch.float32, device=data.device)
        return _lib.PAD_ZEROS, pv
    raise ValueError("unknown padding mode %r" % (padding_mode,))


def fused_padding(padding_mode, data):
    """Padding for a stage of the fused chain, or None when the mode needs the generic path
    ('lowest' takes the minimum of the stage's own input, which only exists mid-kernel)."""
    if isinstance(padding_mode, str) and padding_mode == "lowest":
        return None
    return parse_padding(padding_mode, data)


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def field_lanes(d):
    return 2 if d == 2 else 4


# --------------------------------------------------------------------------------------- affine


class AffineTheta(Function):
    """param (N x 5|9) -> theta, theta_inv (N x d x (d+1)). adv_affine.py:210-273, 316-324."""

    @staticmethod
    def forward(ctx, param, cfg, pscale):
        param = _f32c(param)
        n, d = param.shape[0], cfg.d
        theta = torch.empty(n, d, d + 1, dtype=torch.float32, device=param.device)
        theta_inv = torch.empty_like(theta)
        call("advk_affine_theta_fwd", C.byref(cfg), ptr(param), pscale, n, ptr(theta), ptr(theta_inv),
             stream())
        ctx.save_for_backward(param)
        ctx.cfg, ctx.pscale = cfg, pscale
        return theta, theta_inv

    @staticmethod
    @once_differentiable
    def backward(ctx, g_theta, g_theta_inv):
        (param,) = ctx.saved_tensors
        g_param = torch.empty_like(param)
        call("advk_affine_theta_bwd", C.byref(ctx.cfg), ptr(param), ctx.pscale, param.shape[0],
             ptr(_f32c(g_theta)), ptr(_f32c(g_theta_inv)), ptr(g_param), stream())
        return g_param, None, None


# --------------------------------------------------------------------------------------- warps


class WarpAffine(Function):
    """F.affine_grid + F.grid_sample with the grid generated on the fly. adv_affine.py:297-313."""

    @staticmethod
    def forward(ctx, src, theta, pad, interp, padv):
        src, theta = _f32c(src), _f32c(theta)
        g = _lib.geom(src.shape)
        out = torch.empty_like(src)
        call("advk_warp_affine_fwd", C.byref(g), src.shape[1], ptr(src), ptr(theta), pad, interp,
             ptr(padv), ptr(out), stream())
        ctx.save_for_backward(src, theta, padv)
        ctx.meta = (g, pad, interp)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        src, theta, padv = ctx.saved_tensors
        g, pad, interp = ctx.meta
        g_src = torch.zeros_like(src) if ctx.needs_input_grad[0] else None
        g_theta = torch.zeros_like(theta) if ctx.needs_input_grad[1] else None
        if g_src is not None or g_theta is not None:
            call("advk_warp_affine_bwd", C.byref(g), src.shape[1], ptr(_f32c(g_out)), ptr(src),
                 ptr(theta), pad, interp, ptr(padv), ptr(g_src), ptr(g_theta), stream())
        return g_src, g_theta, None, None, None


class WarpField(Function):
    """F.grid_sample with a dense interleaved field. adv_morph.py:546-557."""

    @staticmethod
    def forward(ctx, src, field, pad, interp, padv):
        src = _f32c(src)
        g = _lib.geom(src.shape)
        out = torch.empty_like(src)
        call("advk_warp_field_fwd", C.byref(g), src.shape[1], ptr(src), ptr(field), pad, interp,
             ptr(padv), ptr(out), stream())
        ctx.save_for_backward(src, field, padv)
        ctx.meta = (g, pad, interp)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        src, field, padv = ctx.saved_tensors
        g, pad, interp = ctx.meta
        g_src = torch.zeros_like(src) if ctx.needs_input_grad[0] else None
        g_field = torch.empty_like(field) if ctx.needs_input_grad[1] else None
        if g_src is not None or g_field is not None:
            call("advk_warp_field_bwd", C.byref(g), src.shape[1], ptr(_f32c(g_out)), ptr(src),
                 ptr(field), pad, interp, ptr(padv), ptr(g_src), ptr(g_field), stream())
        return g_src, g_field, None, None, None


# --------------------------------------------------------------------------------------- morph


def morph_unorm2(v, size, cfg, scale):
    """||upsample(G * (scale v))||_F^2 over the batch -- a device scalar (adv_morph.py:159-162)."""
    v = _f32c(v)
    g = _lib.geom(size)
    u_lr = torch.empty_like(v)
    out = torch.empty(1, dtype=torch.float32, device=v.device)
    call("advk_morph_unorm2", C.byref(g), C.byref(cfg), ptr(v), float(scale), ptr(u_lr), ptr(out),
         stream())
    return out


class MorphField(Function):
    """velocity -> (unclamped) deformation field; DemonsCompose, adv_morph.py:454-491."""

    @staticmethod
    def forward(ctx, v, size, cfg, scale, nb_steps, norm_out=None):
        v = _f32c(v)
        g = _lib.geom(size)
        lanes = field_lanes(g.d)
        spatial = tuple(size[2:])
        nvox = g.N * g.D * g.H * g.W
        u_lr = torch.empty_like(v)
        levels = torch.empty((nb_steps + 2) * nvox * lanes, dtype=torch.float32, device=v.device)
        field = torch.empty((g.N,) + spatial + (lanes,), dtype=torch.float32, device=v.device)
        call("advk_morph_field_fwd", C.byref(g), C.byref(cfg), ptr(v), float(scale), nb_steps,
             ptr(u_lr), ptr(levels), ptr(field), ptr(norm_out), stream())
        ctx.save_for_backward(levels, field)
        ctx.meta = (g, cfg, float(scale), nb_steps, v.shape)
        return field

    @staticmethod
    @once_differentiable
    def backward(ctx, g_field):
        levels, field = ctx.saved_tensors
        g, cfg, scale, nb, vshape = ctx.meta
        g_field = _f32c(g_field)
        scratch = torch.empty(5 * field.numel(), dtype=torch.float32, device=field.device)
        nlr = _lib.load().advk_morph_lr_scratch_floats(C.byref(g), C.byref(cfg))
        lr_scratch = torch.empty(nlr, dtype=torch.float32, device=field.device)
        g_v = torch.empty(vshape, dtype=torch.float32, device=field.device)
        call("advk_morph_field_bwd", C.byref(g), C.byref(cfg), scale, nb, ptr(levels), ptr(field),
             ptr(g_field), ptr(scratch), ptr(lr_scratch), ptr(g_v), stream())
        return g_v, None, None, None, None, None


# --------------------------------------------------------------------------------------- intensity

ORDER_NOISE, ORDER_BIAS, ORDER_NOISE_BIAS, ORDER_BIAS_NOISE = 0, 1, 2, 3


class Intensity(Function):
    """AdvNoise / AdvBias stage(s): adv_noise.py:79-90, adv_bias.py:152-188, 279-356.
    `plan` is a BiasPlan (geometry + device matrices) or None for noise only."""

    @staticmethod
    def forward(ctx, x, delta, cp, order, noise_scale, plan, cp_scale, ignore):
        x = _f32c(x)
        g = _lib.geom(x.shape)
        low = None
        cfg_ref = None
        if order != ORDER_NOISE:
            cp = _f32c(cp)
            cfg_ref = C.byref(plan.cfg(x.device))
            low = torch.empty((g.N,) + tuple(plan.low_size), dtype=torch.float32, device=x.device)
            call("advk_bias_lowfield_fwd", cfg_ref, g.N, ptr(cp), float(cp_scale), ptr(low), stream())
        if order != ORDER_BIAS:
            delta = _f32c(delta)
        out = torch.empty_like(x)
        use_ig = 0 if ignore is None else 1
        call("advk_intensity_fwd", C.byref(g), x.shape[1], order, ptr(x),
             ptr(delta) if order != ORDER_BIAS else None, float(noise_scale), ptr(low), cfg_ref,
             use_ig, float(ignore or 0.0), ptr(out), None, stream())
        ctx.save_for_backward(x, delta if order != ORDER_BIAS else None, low)
        ctx.meta = (g, order, float(noise_scale), plan, float(cp_scale), use_ig, float(ignore or 0.0),
                    None if cp is None else cp.shape)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        x, delta, low = ctx.saved_tensors
        g, order, ns, plan, cp_scale, use_ig, ig, cp_shape = ctx.meta
        need_x, need_d, need_cp = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_d = need_d and order != ORDER_BIAS
        need_cp = need_cp and order != ORDER_NOISE
        g_x = torch.empty_like(x) if need_x else None
        g_delta = torch.empty_like(x) if need_d else None
        nvox = g.N * g.D * g.H * g.W
        g_up = torch.empty(nvox, dtype=torch.float32, device=x.device) if need_cp else None
        cfg_ref = C.byref(plan.cfg(x.device)) if order != ORDER_NOISE else None
        call("advk_intensity_bwd", C.byref(g), x.shape[1], order, ptr(_f32c(g_out)), ptr(x), ptr(delta),
             ns, ptr(low), cfg_ref, use_ig, ig, ptr(g_x), ptr(g_delta), ptr(g_up), stream())
        g_cp = None
        if need_cp:
            lib = _lib.load()
            ns_ = lib.advk_bias_scratch_floats(C.byref(g), cfg_ref)
            scratch = torch.empty(max(int(ns_), 1), dtype=torch.float32, device=x.device)
            g_low = torch.empty_like(low)
            call("advk_bias_upsample_adjoint", C.byref(g), cfg_ref, ptr(g_up), ptr(scratch), ptr(g_low),
                 stream())
            g_cp = torch.empty(cp_shape, dtype=torch.float32, device=x.device)
            call("advk_bias_lowfield_bwd", cfg_ref, g.N, ptr(g_low), cp_scale, ptr(g_cp), stream())
        return g_x, g_delta, g_cp, None, None, None, None, None


def bias_field_only(cp, plan, size, cp_scale=1.0):
    """The clipped bias field N x 1 x spatial (the reference's `bias_field` attribute)."""
    cp = _f32c(cp.detach())
    g = _lib.geom(size)
    dev = cp.device
    cfg_ref = C.byref(plan.cfg(dev))
    low = torch.empty((g.N,) + tuple(plan.low_size), dtype=torch.float32, device=dev)
    call("advk_bias_lowfield_fwd", cfg_ref, g.N, ptr(cp), float(cp_scale), ptr(low), stream())
    ones = torch.ones((g.N, 1) + tuple(size[2:]), dtype=torch.float32, device=dev)
    out = torch.empty_like(ones)
    call("advk_intensity_fwd", C.byref(g), 1, ORDER_BIAS, ptr(ones), None, 0.0, ptr(low), cfg_ref, 0, 0.0,
         ptr(out), None, stream())
    return out


# --------------------------------------------------------------------------------------- fused chain


class BiasLow(Function):
    """control points -> low-res bias field (conv_transposeNd + crop, adv_bias.py:293-307)."""

    @staticmethod
    def forward(ctx, cp, plan, cp_scale):
        cp = _f32c(cp)
        n = cp.shape[0]
        low = torch.empty((n,) + tuple(plan.low_size), dtype=torch.float32, device=cp.device)
        call("advk_bias_lowfield_fwd", C.byref(plan.cfg(cp.device)), n, ptr(cp), float(cp_scale), ptr(low),
             stream())
        ctx.meta = (plan, float(cp_scale), cp.shape)
        return low

    @staticmethod
    @once_differentiable
    def backward(ctx, g_low):
        plan, cp_scale, shape = ctx.meta
        g_low = _f32c(g_low)
        g_cp = torch.empty(shape, dtype=torch.float32, device=g_low.device)
        call("advk_bias_lowfield_bwd", C.byref(plan.cfg(g_low.device)), shape[0], ptr(g_low), cp_scale,
             ptr(g_cp), stream())
        return g_cp, None, None


class ChainSpec(object):
    """Host description of a fused chain: `stages` is a list of dicts
         {"kind": "intensity", "order", "noise_scale", "ignore", "plan", "delta": slot|None, "low": slot|None}
         {"kind": "field" | "affine", "pad", "interp", "padv": tensor|None, "param": slot}
    where a slot indexes the flat tensor list handed to ChainApply.apply."""

    def __init__(self, stages, clamp=None, want_mask=False, binarize=False):
        self.stages = stages
        self.clamp = clamp
        self.want_mask = want_mask
        self.binarize = binarize


def _fill_desc(spec, g, channels, params, grads=None):
    d = _lib.ChainDesc()
    d.g, d.C, d.n_stages = g, channels, len(spec.stages)
    keep = []
    for k, st in enumerate(spec.stages):
        e = d.stages[k]
        if st["kind"] == "intensity":
            e.kind = _lib.STAGE_INTENSITY
            e.intensity_order = st["order"]
            e.noise_scale = float(st["noise_scale"])
            e.use_ignore = 0 if st["ignore"] is None else 1
            e.ignore_value = float(st["ignore"] or 0.0)
            if st["delta"] is not None:
                e.delta = ptr(params[st["delta"]])
            if st["low"] is not None:
                cfg = st["plan"].cfg(params[st["low"]].device)
                keep.append(cfg)
                e.bias = C.pointer(cfg)
                e.low = ptr(params[st["low"]])
            if grads is not None:
                if st["delta"] is not None and grads[st["delta"]] is not None:
                    e.g_delta = ptr(grads[st["delta"]])
                if st["low"] is not None and grads[st["low"]] is not None:
                    e.g_up = ptr(grads[st["low"]])
        else:
            field = st["kind"] == "field"
            e.kind = _lib.STAGE_WARP_FIELD if field else _lib.STAGE_WARP_AFFINE
            e.pad_mode, e.interp = st["pad"], st["interp"]
            e.pad_values = ptr(st["padv"])
            if field:
                e.field = ptr(params[st["param"]])
            else:
                e.theta = ptr(params[st["param"]])
            if grads is not None and grads[st["param"]] is not None:
                if field:
                    e.g_field = ptr(grads[st["param"]])
                else:
                    e.g_theta = ptr(grads[st["param"]])
    if spec.clamp is not None:
        d.do_clamp, d.clamp_lo, d.clamp_hi = 1, float(spec.clamp[0]), float(spec.clamp[1])
    d.want_mask = 1 if spec.want_mask else 0
    d.binarize_mask = 1 if spec.binarize else 0
    return d, keep


class ChainApply(Function):
    """The fused chain-apply launch (advk_chain_apply_fwd/bwd, include/advk.h): solver.forward,
    predict_forward, backward / predict_backward and the valid-region mask of
    adv_compose_solver.py:148-219, 321-325."""

    @staticmethod
    def forward(ctx, src, mask_src, spec, *params):
        src = _f32c(src)
        g = _lib.geom(src.shape)
        params = [None if p is None else _f32c(p) for p in params]
        d, keep = _fill_desc(spec, g, src.shape[1], params)
        n_stash, n_scr = _lib._Z(), _lib._Z()
        call("advk_chain_workspace_floats", C.byref(d), C.byref(n_stash), C.byref(n_scr))
        stash = torch.empty(max(int(n_stash.value), 1), dtype=torch.float32, device=src.device)
        out = torch.empty_like(src)
        mask_out = None
        if spec.want_mask:
            mask_out = torch.empty((g.N, 1) + tuple(src.shape[2:]), dtype=torch.float32, device=src.device)
            if mask_src is not None:
                mask_src = _f32c(mask_src)
        call("advk_chain_apply_fwd", C.byref(d), ptr(src), ptr(mask_src) if spec.want_mask else None,
             ptr(stash), ptr(out), ptr(mask_out), stream())
        ctx.save_for_backward(src, stash, *params)
        ctx.meta = (spec, g, int(n_scr.value))
        # the mask and the stash never carry a gradient: without this autograd hands backward() freshly
        # zero-filled tensors of their sizes (19 M + 2 M floats for the prediction chain at 128^3: 16 us of fills
        # per chain adjoint, profiles/r02u_launch_list.md)
        ctx.set_materialize_grads(False)
        if mask_out is not None:
            ctx.mark_non_differentiable(mask_out, stash)
        else:
            ctx.mark_non_differentiable(stash)
        return out, mask_out, stash

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out, _g_mask, _g_stash):
        saved = ctx.saved_tensors
        src, stash, params = saved[0], saved[1], list(saved[2:])
        spec, g, n_scr = ctx.meta
        need = ctx.needs_input_grad
        if g_out is None:                                   # the chain output itself was not used
            return (None, None, None) + (None,) * len(params)
        grads = [None] * len(params)
        ups = []
        for st in spec.stages:
            if st["kind"] == "intensity":
                if st["delta"] is not None and need[3 + st["delta"]]:
                    grads[st["delta"]] = torch.empty_like(params[st["delta"]])
                if st["low"] is not None and need[3 + st["low"]]:
                    grads[st["low"]] = torch.empty(g.N * g.D * g.H * g.W, dtype=torch.float32, device=src.device)
                    ups.append(st)
            elif need[3 + st["param"]]:
                prm = params[st["param"]]
                grads[st["param"]] = torch.empty_like(prm) if st["kind"] == "field" else torch.zeros_like(prm)
        g_src = torch.empty_like(src) if need[0] else None
        d, keep = _fill_desc(spec, g, src.shape[1], params, grads)
        scratch = torch.empty(max(n_scr, 1), dtype=torch.float32, device=src.device)
        call("advk_chain_apply_bwd", C.byref(d), ptr(_f32c(g_out)), ptr(src), ptr(stash), ptr(scratch),
             ptr(g_src), stream())
        for st in ups:           # g_up (N x S) -> g_low through the adjoint of the linear upsample
            low = params[st["low"]]
            cfg_ref = C.byref(st["plan"].cfg(src.device))
            ns_ = _lib.load().advk_bias_scratch_floats(C.byref(g), cfg_ref)
            tmp = torch.empty(max(int(ns_), 1), dtype=torch.float32, device=src.device)
            g_low = torch.empty_like(low)
            call("advk_bias_upsample_adjoint", C.byref(g), cfg_ref, ptr(grads[st["low"]]), ptr(tmp), ptr(g_low),
                 stream())
            grads[st["low"]] = g_low
        return (g_src, None, None) + tuple(grads)


def run_chain(src, stages, clamp=None, mask_src=None, want_mask=False, binarize=False):
    """stages: list of host stage dicts holding TENSORS ("delta", "cp", "field", "theta"); adjacent
    noise/bias stages with the same ignore value are merged into one INTENSITY stage.
    Returns (out, mask_out or None, per-stage-boundary views of the stash)."""
    merged = []
    for st in stages:
        if (st["kind"] == "intensity" and merged and merged[-1]["kind"] == "intensity"
                and merged[-1]["ignore"] == st["ignore"] and merged[-1]["order"] in (ORDER_NOISE, ORDER_BIAS)
                and st["order"] in (ORDER_NOISE, ORDER_BIAS) and merged[-1]["order"] != st["order"]):
            a = merged[-1]
            first_noise = a["order"] == ORDER_NOISE
            n_, b_ = (a, st) if first_noise else (st, a)
            merged[-1] = dict(kind="intensity", order=ORDER_NOISE_BIAS if first_noise else ORDER_BIAS_NOISE,
                              noise_scale=n_["noise_scale"], ignore=a["ignore"], plan=b_["plan"],
                              cp_scale=b_["cp_scale"], delta=n_["delta"], cp=b_["cp"], span=a.get("span", 1) + 1)
        else:
            merged.append(dict(st))
    if len(merged) > _lib.CHAIN_MAX_STAGES:
        return None
    params, spec_stages = [], []
    for st in merged:
        if st["kind"] == "intensity":
            e = dict(kind="intensity", order=st["order"], noise_scale=st.get("noise_scale", 0.0),
                     ignore=st["ignore"], plan=st.get("plan"), delta=None, low=None)
            if st["order"] != ORDER_BIAS:
                e["delta"] = len(params)
                params.append(st["delta"])
            if st["order"] != ORDER_NOISE:
                e["low"] = len(params)
                params.append(BiasLow.apply(st["cp"], st["plan"], st["cp_scale"]))
        else:
            e = dict(kind=st["kind"], pad=st["pad"], interp=st["interp"], padv=st.get("padv"), param=len(params))
            params.append(st["field"] if st["kind"] == "field" else st["theta"])
        spec_stages.append(e)
    spec = ChainSpec(spec_stages, clamp=clamp, want_mask=want_mask, binarize=binarize)
    out, mask_out, stash = ChainApply.apply(src, mask_src, spec, *params)
    return out, mask_out, (stash, merged)


# --------------------------------------------------------------------------------------- loss


class ConsistencyLoss(Function):
    """calc_segmentation_consistency (common/loss.py:8-87) for scales=[0], types mse/contour/kl."""

    @staticmethod
    def forward(ctx, output, reference, mask, w_mse, w_contour, is_gt, w_kl=0.0):
        output, reference = _f32c(output), _f32c(reference.detach())
        g = _lib.geom(output.shape)
        k = output.shape[1]
        if mask is not None:
            mask = _f32c(mask.detach()[:, :1])
        n = _lib.load().advk_loss_scratch_floats(C.byref(g), k)
        scratch = torch.empty(int(n), dtype=torch.float32, device=output.device)
        loss = torch.empty((), dtype=torch.float32, device=output.device)
        call("advk_consistency_loss_fwd", C.byref(g), k, ptr(output), ptr(reference), ptr(mask),
             float(w_mse), float(w_contour), float(w_kl), 1 if is_gt else 0, ptr(scratch), ptr(loss), stream())
        ctx.save_for_backward(scratch, mask)
        ctx.meta = (g, k, float(w_mse), float(w_contour), float(w_kl), 1 if is_gt else 0, output.shape)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loss):
        scratch, mask = ctx.saved_tensors
        g, k, w_mse, w_contour, w_kl, is_gt, shape = ctx.meta
        g_out = torch.empty(shape, dtype=torch.float32, device=scratch.device)
        call("advk_consistency_loss_bwd", C.byref(g), k, ptr(mask), w_mse, w_contour, w_kl, is_gt, ptr(scratch),
             ptr(_f32c(g_loss).reshape(1)), ptr(g_out), stream())
        return g_out, None, None, None, None, None, None


# --------------------------------------------------------------------------------------- glue


class Clamp(Function):
    """torch.clamp(x, lo, hi) of solver.forward (adv_compose_solver.py:167-175)."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        x = _f32c(x)
        out = torch.empty_like(x)
        call("advk_clamp", ptr(x), float(lo), float(hi), ptr(out), x.numel(), stream())
        ctx.save_for_backward(x)
        ctx.lohi = (float(lo), float(hi))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        (x,) = ctx.saved_tensors
        g_x = torch.empty_like(x)
        call("advk_clamp_bwd", ptr(_f32c(g_out)), ptr(x), ctx.lohi[0], ctx.lohi[1], ptr(g_x), x.numel(),
             stream())
        return g_x, None, None


def nonzero_mask_(t):
    """t = (t != 0) in place (adv_compose_solver.py:325)."""
    call("advk_nonzero_mask", ptr(t), t.numel(), stream())
    return t


def pgd_update_(param, grad, step, mode, guard=None):
    """In-place PGD update of a parameter tensor; see ADVK_UPD_* in include/advk.h.  `guard`: device
    scalar (the step's loss); a non-finite value leaves `param` unchanged (device-side NaN guard)."""
    n = param.shape[0]
    per = param.numel() // n
    ss = torch.empty(n, dtype=torch.float64, device=param.device)
    call("advk_pgd_update_guarded", ptr(param), ptr(_f32c(grad)), float(step), mode, n, per,
         ptr(ss, torch.float64),
         None if guard is None else ptr(guard), stream())
    return param
