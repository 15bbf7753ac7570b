"""Consistency loss the solver calls every inner step (reference: advchain/common/loss.py:8-249).

On CUDA tensors the single-scale configurations of the solver -- scales=[0], divergence types 'mse',
'contour' and/or 'kl', a channel-uniform mask -- run in the fused advk_consistency_loss_* kernels
(SURVEY.md section 8f rank f1).  Everything else (multi-scale, per-channel masks, CPU tensors)
takes the plain PyTorch formulation below, written to reproduce the reference's numbers
including its quirks -- Q9 (the 'mse' term is divided a second time by N*S) and Q10 (3-D contour
loss uses the x-kernel for y and the last-assigned kernel for z).
"""
import torch
import torch.nn.functional as F

_SOBEL_CACHE = {}


def _sobel(d, device, dtype):
    key = (d, str(device), dtype)
    if key not in _SOBEL_CACHE:
        if d == 2:
            kx = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]])
            ky = torch.tensor([[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]])
            ks = [kx, ky]
        else:
            h = torch.tensor([1., 2., 1.])
            hp = torch.tensor([1., 0., -1.])
            gx = h[:, None, None] * hp[None, :, None] * h[None, None, :]
            gz = h[:, None, None] * h[None, :, None] * hp[None, None, :]
            ks = [gx, gx, gz]
        _SOBEL_CACHE[key] = [k[None, None].to(device=device, dtype=dtype) for k in ks]
    return _SOBEL_CACHE[key]


def contour_loss(input, target, mask=None, **unused):
    """Sobel-edge MSE for one class channel (loss.py:102-220 with ignore_background=False,
    one_hot_target=False, as the consistency loss calls it)."""
    d = input.dim() - 2
    conv = F.conv2d if d == 2 else F.conv3d
    ks = _sobel(d, input.device, input.dtype)
    m = 1.0 if mask is None else mask[:, :input.shape[1]]
    total = 0.
    for k in ks:
        w = k.expand(input.shape[1], input.shape[1], *k.shape[2:])
        total = total + F.mse_loss(conv(input, w, padding=1) * m, conv(target, w, padding=1) * m)
    return total / len(ks)


def kl_divergence(reference, pred, mask=None, is_gt=False):
    """loss.py:223-249."""
    if mask is None:
        mask = torch.ones_like(pred)
    if not is_gt:
        p = F.softmax(reference, dim=1)
        log_p = F.log_softmax(reference, dim=1)
    else:
        p = torch.where(reference == 0, 1e-8, 1 - 1e-8)
        log_p = torch.log(p)
    plogp = torch.sum(mask * (p * log_p), dim=1)
    plogq = torch.sum(mask * (p * F.log_softmax(pred, dim=1)), dim=1)
    return torch.mean(plogp - plogq)


def _fusable(output, reference, divergence_types, divergence_weights, scales, mask):
    if not (output.is_cuda and reference.is_cuda and output.dtype == torch.float32):
        return False
    if list(scales) != [0] or output.shape != reference.shape or output.dim() not in (4, 5):
        return False
    if len(divergence_types) != len(divergence_weights) or not divergence_types:
        return False
    if any(t not in ('mse', 'contour', 'kl') for t in divergence_types):
        return False
    if reference.requires_grad and torch.is_grad_enabled():
        return False                      # the fused backward only produces dL/d(output)
    if mask is not None:
        if not mask.is_cuda or mask.shape[0] != output.shape[0] or mask.shape[2:] != output.shape[2:]:
            return False
        # a K-channel view of ONE channel (the solver's valid-region mask).  An N x 1 x spatial mask changes
        # the reference's Q9 divisor numel(mask)/K (loss.py:62-64) to N*S/K, which the kernel does not
        # implement: it takes the PyTorch formulation below.
        if not (mask.shape[1] == output.shape[1] and (mask.stride(1) == 0 or output.shape[1] == 1)):
            return False
    return True


def calc_segmentation_consistency(output, reference, divergence_types=['kl', 'contour'],
                                  divergence_weights=[1.0, 0.5], class_weights=None, scales=[0],
                                  mask=None, is_gt=False):
    """loss.py:8-87."""
    if class_weights is not None:
        raise NotImplementedError
    if _fusable(output, reference, divergence_types, divergence_weights, scales, mask):
        from ..augmentor import _ops
        w = {'mse': 0.0, 'contour': 0.0, 'kl': 0.0}
        for name, weight in zip(divergence_types, divergence_weights):
            w[name] += float(weight)
        return _ops.ConsistencyLoss.apply(output, reference, mask, w['mse'], w['contour'], bool(is_gt), w['kl'])
    num_classes = reference.size(1)
    spatial_dims = output.dim() - 2
    assert spatial_dims in (2, 3), 'only support 2d or 3d segmentation'
    assert output.dim() == reference.dim()
    if mask is None:
        mask = torch.ones_like(output)
    dist = 0.
    for scale in scales:
        if scale > 0:
            pool = F.avg_pool2d if spatial_dims == 2 else F.avg_pool3d
            ref_s, out_s = pool(reference, 2 ** scale), pool(output, 2 ** scale)
        else:
            ref_s, out_s = reference, output
        for name, weight in zip(divergence_types, divergence_weights):
            if name == 'kl':
                loss = kl_divergence(pred=out_s, reference=ref_s, mask=mask, is_gt=is_gt)
            elif name == 'mse':
                target = ref_s if is_gt else torch.softmax(ref_s, dim=1)
                pred = torch.softmax(out_s, dim=1)
                loss = F.mse_loss(pred * mask, target * mask)
                loss = loss / (torch.numel(mask) / num_classes)          # quirk Q9
            elif name == 'contour':
                target = ref_s if is_gt else torch.softmax(ref_s, dim=1)
                pred = torch.softmax(out_s, dim=1)
                loss, cnt = 0., 0
                for i in range(1, num_classes):
                    cnt += 1
                    loss = loss + contour_loss(pred[:, [i]], target[:, [i]], mask=mask)
                if cnt > 0:
                    loss = loss / cnt
            else:
                raise NotImplementedError(name)
            dist = dist + 2 ** scale * (weight * loss)
    return dist / (1.0 * len(scales))
