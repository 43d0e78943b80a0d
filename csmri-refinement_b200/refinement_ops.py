"""Pointwise ops of the refinement path on the GPU.

Mirrors, with the reference's names and argument meaning:

* ``scale`` / ``unscale``            models/refinement_wrapper.py:51-92 (``_scale``, ``_unscale``)
* ``magnitude_image``                utils/tensor_transforms.py:78-99
* ``complex_abs``                    utils/tensor_transforms.py:62-75, differentiable
* ``refinement_real_penalty_add``    RefinementWrapper._refinement_real_penalty_add,
                                     models/refinement_wrapper.py:173-197

Each is one reduction pass plus one map pass through ``libcsmri_dc.so``
(``csmri_plane_minmax`` / ``csmri_plane_scale`` / ``csmri_refine_real_penalty_add``)
instead of the reference's 6-10 elementwise launches, and evaluates the same
float32 expression op by op, so results are bit-identical to the torch
evaluation.  CUDA tensors only; no fallback.
"""
import torch
from torch.autograd.function import once_differentiable

from . import _lib, rec_transforms


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_bchw(name, t, channels=None):
    if t.dim() != 4 or (channels is not None and t.size(1) != channels):
        raise ValueError('%s must be (B,%s,H,W), got %s' % (name, channels or 'C', tuple(t.shape)))
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: the refinement ops have no CPU fallback' % name)


def _minmax(t):
    """t (B,C,H,W) contiguous -> minimum, maximum of shape (B,C,1)."""
    b, c, h, w = t.shape
    minimum = torch.empty((b, c, 1), dtype=torch.float32, device=t.device)
    maximum = torch.empty_like(minimum)
    _lib.check(_lib.lib().csmri_plane_minmax(t.data_ptr(), minimum.data_ptr(), maximum.data_ptr(),
                                             b * c, h * w, h * w, _stream()))
    return minimum, maximum


def _map(t, minimum, maximum, mode):
    b, c, h, w = t.shape
    if minimum.numel() != b * c or maximum.numel() != b * c:
        raise ValueError('minimum / maximum must hold one value per example and channel')
    out = torch.empty_like(t)
    _lib.check(_lib.lib().csmri_plane_scale(
        t.data_ptr(), minimum.contiguous().data_ptr(), maximum.contiguous().data_ptr(),
        out.data_ptr(), b * c, h * w, h * w, h * w, mode, _stream()))
    return out


def scale(tensor):
    """``_scale``: per example and channel to the range (-1, 1).
    -> (scaled_tensor, minimum, maximum), minimum / maximum of shape (B,C,1)."""
    _check_bchw('tensor', tensor)
    tensor = tensor.contiguous()
    with torch.cuda.device(tensor.device):
        minimum, maximum = _minmax(tensor)
        return _map(tensor, minimum, maximum, 1), minimum, maximum


def unscale(tensor, minimum, maximum):
    """``_unscale``: back from (-1, 1) with the minimum / maximum of :func:`scale`."""
    _check_bchw('tensor', tensor)
    tensor = tensor.contiguous()
    with torch.cuda.device(tensor.device):
        return _map(tensor, minimum, maximum, 2)


def magnitude_image(tensor):
    """``magnitude_image``: (B,2,H,W) -> magnitude (B,1,H,W) scaled to (0, 1)
    with the minimum and maximum of each image."""
    mag = rec_transforms.magnitude(tensor, lo=float('-inf'), hi=float('inf'))
    with torch.cuda.device(mag.device):
        minimum, maximum = _minmax(mag)
        return _map(mag, minimum, maximum, 0)


class _RefineRealPenaltyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out_pretrained, out_learnable, scale_param):
        b, _, h, w = out_pretrained.shape
        dev = out_pretrained.device
        with torch.cuda.device(dev):
            pred = torch.empty_like(out_pretrained)
            minimum = torch.empty((b,), dtype=torch.float32, device=dev)
            maximum = torch.empty_like(minimum)
            _lib.check(_lib.lib().csmri_refine_real_penalty_add(
                out_pretrained.data_ptr(), out_learnable.data_ptr(), scale_param.data_ptr(),
                pred.data_ptr(), minimum.data_ptr(), maximum.data_ptr(), b, h, w, _stream()))
        ctx.save_for_backward(out_learnable, scale_param, maximum)
        return pred

    @staticmethod
    @once_differentiable     # raw kernel on data pointers
    def backward(ctx, grad_pred):
        out_learnable, scale_param, maximum = ctx.saved_tensors
        grad_pred = grad_pred.contiguous()
        b, _, h, w = grad_pred.shape
        lib = _lib.lib()
        with torch.cuda.device(grad_pred.device):
            grad_learn = torch.empty_like(out_learnable)
            partial = torch.empty((b, lib.csmri_refine_partials()), dtype=torch.float32,
                                  device=grad_pred.device)
            _lib.check(lib.csmri_refine_real_penalty_add_backward(
                grad_pred.data_ptr(), out_learnable.data_ptr(), scale_param.data_ptr(),
                maximum.data_ptr(), grad_learn.data_ptr(), partial.data_ptr(), b, h, w, _stream()))
        grad_scale = partial.sum().reshape(scale_param.shape) if ctx.needs_input_grad[2] else None
        # the pretrained output is detached in the reference (refinement_wrapper.py:211-212)
        return None, (grad_learn if ctx.needs_input_grad[1] else None), grad_scale


def refinement_real_penalty_add(out_pretrained, out_learnable, scale_param):
    """``RefinementWrapper._refinement_real_penalty_add`` for a given output of
    the learnable model.  ``out_pretrained`` (B,2,H,W) is treated as detached
    (the reference freezes and detaches the pretrained path); gradients flow to
    ``out_learnable`` (B,1,H,W) and ``scale_param`` (1,).  Returns the
    reference's dict."""
    _check_bchw('out_pretrained', out_pretrained, 2)
    _check_bchw('out_learnable', out_learnable, 1)
    if out_pretrained.requires_grad:
        raise RuntimeError('out_pretrained must be detached (frozen pretrained path, '
                           'models/refinement_wrapper.py:211-212)')
    if out_learnable.shape[0] != out_pretrained.shape[0] or \
            out_learnable.shape[2:] != out_pretrained.shape[2:]:
        raise ValueError('out_learnable %s does not match out_pretrained %s' % (
            tuple(out_learnable.shape), tuple(out_pretrained.shape)))
    if scale_param.numel() != 1 or not scale_param.is_cuda or scale_param.dtype != torch.float32:
        raise ValueError('scale_param must be a float32 CUDA tensor with one element')
    pred = _RefineRealPenaltyAdd.apply(out_pretrained.contiguous(), out_learnable.contiguous(),
                                       scale_param)
    return {'pred': pred, 'pretrained': out_pretrained, 'prescaled_refinement': out_learnable,
            'scaled_refinement': scale_param * out_learnable}


class _ComplexAbs(torch.autograd.Function):
    """|z| of a planar complex batch through ``csmri_magnitude_clamp`` (bit-identical
    to ``(re**2 + im**2) ** 0.5``); backward d|z| = (re, im) / |z| (0 where |z| = 0)."""

    @staticmethod
    def forward(ctx, x):
        mag = rec_transforms.magnitude(x, lo=float('-inf'), hi=float('inf'))
        ctx.save_for_backward(x, mag)
        return mag

    @staticmethod
    def backward(ctx, grad):
        x, mag = ctx.saved_tensors
        return x * (grad / mag.clamp_min(torch.finfo(torch.float32).tiny)) * (mag > 0)


def complex_abs(tensor):
    """``complex_abs`` (utils/tensor_transforms.py:62-75): (B,2,H,W) -> (B,1,H,W),
    differentiable (the generator losses of config 5 back-propagate through it)."""
    _check_bchw('tensor', tensor, 2)
    return _ComplexAbs.apply(tensor.contiguous())
