"""ctypes binding of libadvchain_b200.so (the C ABI declared in include/advk.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
product path raises.  PyTorch is used only for device memory, streams and autograd plumbing;
every device computation of the hot path goes through the `advk_*` entry points.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadvchain_b200.so")
ABI_VERSION = 3

PAD_ZEROS, PAD_BORDER, PAD_REFLECTION = 0, 1, 2
INTERP_LINEAR, INTERP_NEAREST, INTERP_BICUBIC = 0, 1, 2
UPD_L2_ASCENT, UPD_SIGN_ASCENT, UPD_L2_POWER, UPD_SIGN_POWER = 0, 1, 2, 3


class Geom(C.Structure):
    _fields_ = [("d", C.c_int), ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int)]


class AffineCfg(C.Structure):
    _fields_ = [("d", C.c_int), ("rot", C.c_float * 3), ("scale", C.c_float * 3),
                ("shift", C.c_float * 3)]


class MorphCfg(C.Structure):
    _fields_ = [("lr", C.c_int * 3), ("ktaps", C.c_int), ("gauss", C.c_float * 15)]


class BiasCfg(C.Structure):
    _fields_ = [("n_cp", C.c_int * 3), ("low", C.c_int * 3), ("A", C.c_void_p * 3),
                ("up_scale", C.c_float * 3), ("upsample", C.c_int), ("use_log", C.c_int),
                ("magnitude", C.c_float)]


class ChainStage(C.Structure):
    _fields_ = [("kind", C.c_int), ("intensity_order", C.c_int), ("noise_scale", C.c_float),
                ("use_ignore", C.c_int), ("ignore_value", C.c_float), ("bias", C.POINTER(BiasCfg)),
                ("delta", C.c_void_p), ("low", C.c_void_p), ("field", C.c_void_p), ("theta", C.c_void_p),
                ("pad_mode", C.c_int), ("interp", C.c_int), ("pad_values", C.c_void_p),
                ("g_delta", C.c_void_p), ("g_up", C.c_void_p), ("g_field", C.c_void_p),
                ("g_theta", C.c_void_p)]


CHAIN_MAX_STAGES = 6
STAGE_INTENSITY, STAGE_WARP_FIELD, STAGE_WARP_AFFINE = 0, 1, 2


class ChainDesc(C.Structure):
    _fields_ = [("g", Geom), ("C", C.c_int), ("n_stages", C.c_int), ("stages", ChainStage * CHAIN_MAX_STAGES),
                ("do_clamp", C.c_int), ("clamp_lo", C.c_float), ("clamp_hi", C.c_float),
                ("want_mask", C.c_int), ("binarize_mask", C.c_int)]


_P = C.c_void_p
_F = C.c_float
_I = C.c_int
_Z = C.c_size_t
_G = C.POINTER(Geom)

# name -> (restype, argtypes); mirrors include/advk.h one to one
SIGNATURES = {
    "advk_abi_version": (_I, []),
    "advk_last_error": (C.c_char_p, []),
    "advk_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_Z)]),
    "advk_kernel_count": (_I, []),
    "advk_kernel_name": (C.c_char_p, [_I]),
    "advk_launch_count": (C.c_ulonglong, [_I, _I]),
    "advk_prof_configure": (_I, [_I, _I]),
    "advk_prof_collect": (_I, [C.POINTER(_I), C.POINTER(_F), _I]),
    "advk_prof_group_runs": (_I, [_I]),
    "advk_set_pdl": (_I, [_I]),
    "advk_prof_collect_runs": (_I, [C.POINTER(_I), C.POINTER(_F), C.POINTER(_I), _I]),
    "advk_affine_theta_fwd": (_I, [C.POINTER(AffineCfg), _P, _F, _I, _P, _P, _P]),
    "advk_affine_theta_bwd": (_I, [C.POINTER(AffineCfg), _P, _F, _I, _P, _P, _P, _P]),
    "advk_warp_affine_fwd": (_I, [_G, _I, _P, _P, _I, _I, _P, _P, _P]),
    "advk_warp_affine_bwd": (_I, [_G, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "advk_warp_field_fwd": (_I, [_G, _I, _P, _P, _I, _I, _P, _P, _P]),
    "advk_warp_field_bwd": (_I, [_G, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "advk_morph_unorm2": (_I, [_G, C.POINTER(MorphCfg), _P, _F, _P, _P, _P]),
    "advk_morph_field_fwd": (_I, [_G, C.POINTER(MorphCfg), _P, _F, _I, _P, _P, _P, _P, _P]),
    "advk_morph_lr_scratch_floats": (_Z, [_G, C.POINTER(MorphCfg)]),
    "advk_morph_field_bwd": (_I, [_G, C.POINTER(MorphCfg), _F, _I, _P, _P, _P, _P, _P, _P, _P]),
    "advk_bias_lowfield_fwd": (_I, [C.POINTER(BiasCfg), _I, _P, _F, _P, _P]),
    "advk_bias_lowfield_bwd": (_I, [C.POINTER(BiasCfg), _I, _P, _F, _P, _P]),
    "advk_intensity_fwd": (_I, [_G, _I, _I, _P, _P, _F, _P, C.POINTER(BiasCfg), _I, _F, _P, _P, _P]),
    "advk_intensity_bwd": (_I, [_G, _I, _I, _P, _P, _P, _F, _P, C.POINTER(BiasCfg), _I, _F, _P, _P,
                                _P, _P]),
    "advk_bias_scratch_floats": (_Z, [_G, C.POINTER(BiasCfg)]),
    "advk_bias_upsample_adjoint": (_I, [_G, C.POINTER(BiasCfg), _P, _P, _P, _P]),
    "advk_chain_set_cooperative": (_I, [_I]),
    "advk_chain_set_lean": (_I, [_I]),
    "advk_chain_set_packed": (_I, [_I]),
    "advk_chain_workspace_floats": (_I, [C.POINTER(ChainDesc), C.POINTER(_Z), C.POINTER(_Z)]),
    "advk_chain_apply_fwd": (_I, [C.POINTER(ChainDesc), _P, _P, _P, _P, _P, _P]),
    "advk_chain_apply_bwd": (_I, [C.POINTER(ChainDesc), _P, _P, _P, _P, _P, _P]),
    "advk_loss_scratch_floats": (_Z, [_G, _I]),
    "advk_loss_tune": (_I, [_I]),
    "advk_consistency_loss_fwd": (_I, [_G, _I, _P, _P, _P, _F, _F, _F, _I, _P, _P, _P]),
    "advk_consistency_loss_bwd": (_I, [_G, _I, _P, _F, _F, _F, _I, _P, _P, _P, _P]),
    "advk_pgd_update": (_I, [_P, _P, _F, _I, _I, _Z, _P, _P]),
    "advk_pgd_update_guarded": (_I, [_P, _P, _F, _I, _I, _Z, _P, _P, _P]),
    "advk_morph_tune": (_I, [_I]),
    "advk_morph_steps_check": (_I, [_P, _I, _I, _P, _P]),
    "advk_publish_verdict": (_I, [_P, _P, _P, _P, _P, _P]),
    "advk_clamp": (_I, [_P, _F, _F, _P, _Z, _P]),
    "advk_clamp_bwd": (_I, [_P, _P, _F, _F, _P, _Z, _P]),
    "advk_nonzero_mask": (_I, [_P, _Z, _P]),
}

_lib = None


def load():
    """Loads the shared library (once). Raises if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "advchain_b200: %s not found. Build it with `python -m advchain_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.advk_abi_version() != ABI_VERSION:
        raise RuntimeError("advchain_b200: ABI mismatch (library %d, binding %d)"
                           % (lib.advk_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def call(name, *args):
    """Calls an int-returning launch function; raises RuntimeError with the library's message."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.advk_last_error().decode()))


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of exactly `dtype` (None -> NULL).  The kernels launch
    on the calling thread's current device and stream (`stream()`): a tensor living on another device is
    refused here instead of being dereferenced on the wrong one."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("advchain_b200 runs on CUDA tensors only (got %s); there is no CPU "
                           "fallback" % t.device)
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError("advchain_b200: expected a contiguous %s tensor, got %s contiguous=%s"
                           % (dtype, t.dtype, t.is_contiguous()))
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError("advchain_b200: tensor on %s but the current CUDA device is %d; run under "
                           "`torch.cuda.device(tensor.device)` (one process per GPU sets it once)"
                           % (t.device, torch.cuda.current_device()))
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def geom(size):
    """size: [N, C, (D,) H, W] -> Geom (C is not part of it)."""
    sp = list(size[2:])
    if len(sp) == 2:
        return Geom(2, int(size[0]), 1, int(sp[0]), int(sp[1]))
    if len(sp) == 3:
        return Geom(3, int(size[0]), int(sp[0]), int(sp[1]), int(sp[2]))
    raise ValueError("only 2-D / 3-D data is supported, got size %s" % (list(size),))


def device_info():
    lib = load()
    sm, ma, mi, l2 = _I(), _I(), _I(), _Z()
    rc = lib.advk_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2))
    if rc != 0:
        raise RuntimeError(lib.advk_last_error().decode())
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), l2_bytes=l2.value)


def kernel_names():
    lib = load()
    return [lib.advk_kernel_name(i).decode() for i in range(lib.advk_kernel_count())]


def launch_count(kernel=None, reset=False):
    """Kernels launched by the library since the last reset (all, or one kernel by name)."""
    kid = -1 if kernel is None else kernel_names().index(kernel)
    return int(load().advk_launch_count(kid, 1 if reset else 0))


def prof_configure(kernel="all", capacity=8192):
    """Bracket launches of `kernel` ("all", None = off, or a kernel name) with CUDA events."""
    kid = -2 if kernel is None else (-1 if kernel == "all" else kernel_names().index(kernel))
    call("advk_prof_configure", kid, capacity)


def prof_group_runs(enable):
    """Bracket a run of back-to-back launches of the profiled kernel once (see include/advk.h). -> previous setting"""
    return bool(load().advk_prof_group_runs(1 if enable else 0))


def prof_collect_runs(capacity=8192):
    """-> {kernel name: [(ms, launches covered), ...]} for the records since the last collect."""
    lib = load()
    ids, ms, cnt = (_I * capacity)(), (_F * capacity)(), (_I * capacity)()
    n = lib.advk_prof_collect_runs(ids, ms, cnt, capacity)
    if n < 0:
        raise RuntimeError("advk_prof_collect_runs failed: %s" % lib.advk_last_error().decode())
    names = kernel_names()
    out = {}
    for i in range(n):
        out.setdefault(names[ids[i]], []).append((float(ms[i]), int(cnt[i])))
    return out


def prof_collect(capacity=8192):
    """-> {kernel name: [ms, ...]} for the records since the last collect."""
    lib = load()
    ids = (_I * capacity)()
    ms = (_F * capacity)()
    n = lib.advk_prof_collect(ids, ms, capacity)
    if n < 0:
        raise RuntimeError("advk_prof_collect failed: %s" % lib.advk_last_error().decode())
    names = kernel_names()
    out = {}
    for i in range(n):
        out.setdefault(names[ids[i]], []).append(float(ms[i]))
    return out
