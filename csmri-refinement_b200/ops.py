"""PyTorch custom ops on top of the C ABI (include/csmri_dc.h).

PyTorch is plumbing here: it owns device memory and streams, and
``torch.library`` gives the kernels an autograd formula.  The arithmetic lives
in ``csrc/csmri_dc.cu``.  Ops exist for CUDA tensors only; calling them on CPU
tensors raises (no fallback).

Reference interfaces replaced (paths relative to the reference root):
  data/reconstruction/deep_med_lib/my_pytorch/myfft.py:78-128   Fft2d / Ifft2d
  data/reconstruction/deep_med_lib/my_pytorch/myfft.py:131-163  blend + perform
"""
from typing import Optional, Tuple

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_planar(name, t, like=None):
    if t.dim() != 4 or t.size(1) != 2:
        raise ValueError('%s must be (B,2,H,W), got %s' % (name, tuple(t.shape)))
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: the DC path has no CPU fallback' % name)
    if like is not None and (t.shape != like.shape or t.device != like.device):
        raise ValueError('%s: shape/device %s/%s do not match %s/%s' % (
            name, tuple(t.shape), t.device, tuple(like.shape), like.device))


# ---------------------------------------------------------------------------
# csmri::dc_prepare - once per batch (k0 and mask are constant over the cascade)
# ---------------------------------------------------------------------------
@torch.library.custom_op('csmri::dc_prepare', mutates_args=(), device_types='cuda')
def dc_prepare(k0: torch.Tensor, mask: torch.Tensor, noise_lvl: float,
               with_addend: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dtab (B,H), addend (B,2,H,W), row_constant int32[1]).  ``with_addend``
    False skips the row transform of k0 (addend comes back empty): the mask
    analysis alone, for callers that first want to know which path applies."""
    _check_planar('k0', k0)
    _check_planar('mask', mask, k0)
    k0 = k0.contiguous()
    mask = mask.contiguous()
    B, _, H, W = k0.shape
    with torch.cuda.device(k0.device):
        dtab = torch.empty((B, H), dtype=torch.float32, device=k0.device)
        addend = torch.empty_like(k0) if with_addend else k0.new_empty((0,))
        flag = torch.empty((1,), dtype=torch.int32, device=k0.device)
        _lib.check(_lib.lib().csmri_dc_prepare(
            _ptr(k0), _ptr(mask), B, H, W, float(noise_lvl), _ptr(dtab),
            _ptr(addend) if with_addend else None, _ptr(flag), None, _stream()))
    return dtab, addend, flag


@dc_prepare.register_fake
def _(k0, mask, noise_lvl, with_addend=True):
    B, _, H, W = k0.shape
    return (k0.new_empty((B, H)), torch.empty_like(k0) if with_addend else k0.new_empty((0,)),
            k0.new_empty((1,), dtype=torch.int32))


# ---------------------------------------------------------------------------
# csmri::dc_prepare_lines - the same plan from a compact Cartesian description
# ---------------------------------------------------------------------------
@torch.library.custom_op('csmri::dc_prepare_lines', mutates_args=(), device_types='cuda')
def dc_prepare_lines(k0_lines: torch.Tensor, rows: torch.Tensor, noise_lvl: float,
                     width: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """k0_lines (B,2,L,W) = kspace[b][:, rows[b] != 0, :], rows (B,H) uint8
    -> (dtab (B,H), addend (B,2,H,W), lines_ok int32[1])."""
    if k0_lines.dim() != 4 or k0_lines.size(1) != 2 or k0_lines.dtype != torch.float32:
        raise ValueError('k0_lines must be a float32 (B,2,L,W) tensor, got %s %s'
                         % (tuple(k0_lines.shape), k0_lines.dtype))
    if not k0_lines.is_cuda:
        raise RuntimeError('k0_lines must be a CUDA tensor: the DC path has no CPU fallback')
    B, _, L, W = k0_lines.shape
    if W != width:
        raise ValueError('k0_lines has width %d, expected %d' % (W, width))
    if rows.dim() != 2 or rows.size(0) != B or rows.dtype != torch.uint8 or \
            rows.device != k0_lines.device:
        raise ValueError('rows must be a uint8 (B,H) tensor on the same device')
    H = rows.size(1)
    k0_lines, rows = k0_lines.contiguous(), rows.contiguous()
    with torch.cuda.device(k0_lines.device):
        dtab = torch.empty((B, H), dtype=torch.float32, device=k0_lines.device)
        addend = torch.empty((B, 2, H, W), dtype=torch.float32, device=k0_lines.device)
        ok = torch.empty((1,), dtype=torch.int32, device=k0_lines.device)
        _lib.check(_lib.lib().csmri_dc_prepare_lines(
            _ptr(k0_lines), _ptr(rows), B, H, W, L, float(noise_lvl), _ptr(dtab), _ptr(addend),
            _ptr(ok), _stream()))
    return dtab, addend, ok


@dc_prepare_lines.register_fake
def _(k0_lines, rows, noise_lvl, width):
    B, H = rows.shape
    return (k0_lines.new_empty((B, H)), k0_lines.new_empty((B, 2, H, width)),
            k0_lines.new_empty((1,), dtype=torch.int32))


# ---------------------------------------------------------------------------
# csmri::dc_cartesian - out = iFFT_H(dtab * FFT_H(x [+ residual]) [+ addend])
# ---------------------------------------------------------------------------
@torch.library.custom_op('csmri::dc_cartesian', mutates_args=(), device_types='cuda')
def dc_cartesian(x: torch.Tensor, residual: Optional[torch.Tensor], dtab: torch.Tensor,
                 addend: Optional[torch.Tensor]) -> torch.Tensor:
    _check_planar('x', x)
    x = x.contiguous()
    if residual is not None:
        _check_planar('residual', residual, x)
        residual = residual.contiguous()
    if addend is not None:
        _check_planar('addend', addend, x)
        addend = addend.contiguous()
    B, _, H, W = x.shape
    if dtab.shape != (B, H) or dtab.dtype != torch.float32 or not dtab.is_contiguous():
        raise ValueError('dtab must be a contiguous float32 (B,H) tensor')
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        _lib.check(_lib.lib().csmri_dc_forward_cartesian(
            _ptr(x), _ptr(residual), _ptr(dtab), _ptr(addend), _ptr(out), B, H, W, _stream()))
    return out


@dc_cartesian.register_fake
def _(x, residual, dtab, addend):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _dc_cartesian_setup(ctx, inputs, output):
    x, residual, dtab, addend = inputs
    ctx.save_for_backward(dtab)
    ctx.has_residual = residual is not None
    ctx.has_addend = addend is not None


def _dc_cartesian_backward(ctx, grad):
    # F^-1 D F is self-adjoint (real diagonal D, unitary F): the backward of
    # myfft.py:92-102,119-128 chained with the blend is the same kernel.
    (dtab,) = ctx.saved_tensors
    gx = dc_cartesian(grad, None, dtab, None) if ctx.needs_input_grad[0] or (
        ctx.has_residual and ctx.needs_input_grad[1]) else None
    g_res = gx if ctx.has_residual and ctx.needs_input_grad[1] else None
    if ctx.has_addend and ctx.needs_input_grad[3]:
        # the reference never differentiates w.r.t. k0 (loader output,
        # requires_grad=False: utils/__init__.py:75-83)
        raise RuntimeError('gradient w.r.t. the prepared k0 term is not supported')
    return (gx if ctx.needs_input_grad[0] else None), g_res, None, None


dc_cartesian.register_autograd(_dc_cartesian_backward, setup_context=_dc_cartesian_setup)


# ---------------------------------------------------------------------------
# csmri::dc_general / csmri::dc_general_adjoint - arbitrary dense masks
# ---------------------------------------------------------------------------
@torch.library.custom_op('csmri::dc_general', mutates_args=(), device_types='cuda')
def dc_general(x: torch.Tensor, residual: Optional[torch.Tensor], k0: torch.Tensor,
               mask: torch.Tensor, noise_lvl: float) -> torch.Tensor:
    _check_planar('x', x)
    _check_planar('k0', k0, x)
    _check_planar('mask', mask, x)
    x, k0, mask = x.contiguous(), k0.contiguous(), mask.contiguous()
    if residual is not None:
        _check_planar('residual', residual, x)
        residual = residual.contiguous()
    B, _, H, W = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        scratch = torch.empty_like(x)
        _lib.check(_lib.lib().csmri_dc_forward_general(
            _ptr(x), _ptr(residual), _ptr(k0), _ptr(mask), _ptr(out), B, H, W,
            float(noise_lvl), _ptr(scratch), _stream()))
    return out


@dc_general.register_fake
def _(x, residual, k0, mask, noise_lvl):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.custom_op('csmri::dc_general_adjoint', mutates_args=(), device_types='cuda')
def dc_general_adjoint(grad: torch.Tensor, mask: torch.Tensor, noise_lvl: float) -> torch.Tensor:
    _check_planar('grad', grad)
    _check_planar('mask', mask, grad)
    grad, mask = grad.contiguous(), mask.contiguous()
    B, _, H, W = grad.shape
    with torch.cuda.device(grad.device):
        out = torch.empty_like(grad)
        scratch = torch.empty_like(grad)
        _lib.check(_lib.lib().csmri_dc_adjoint_general(
            _ptr(grad), _ptr(mask), _ptr(out), B, H, W, float(noise_lvl), _ptr(scratch),
            _stream()))
    return out


@dc_general_adjoint.register_fake
def _(grad, mask, noise_lvl):
    return torch.empty_like(grad, memory_format=torch.contiguous_format)


def _dc_general_setup(ctx, inputs, output):
    x, residual, k0, mask, noise_lvl = inputs
    ctx.save_for_backward(mask)
    ctx.noise_lvl = noise_lvl
    ctx.has_residual = residual is not None


def _dc_general_backward(ctx, grad):
    (mask,) = ctx.saved_tensors
    need = ctx.needs_input_grad[0] or (ctx.has_residual and ctx.needs_input_grad[1])
    gx = dc_general_adjoint(grad, mask, ctx.noise_lvl) if need else None
    return ((gx if ctx.needs_input_grad[0] else None),
            (gx if ctx.has_residual and ctx.needs_input_grad[1] else None), None, None, None)


dc_general.register_autograd(_dc_general_backward, setup_context=_dc_general_setup)


def _dc_general_adjoint_setup(ctx, inputs, output):
    grad, mask, noise_lvl = inputs
    ctx.save_for_backward(mask)
    ctx.noise_lvl = noise_lvl


def _dc_general_adjoint_backward(ctx, g):
    (mask,) = ctx.saved_tensors
    return dc_general_adjoint(g, mask, ctx.noise_lvl), None, None


dc_general_adjoint.register_autograd(_dc_general_adjoint_backward,
                                     setup_context=_dc_general_adjoint_setup)


# ---------------------------------------------------------------------------
# plain ortho FFT2 / iFFT2 (Fft2d / Ifft2d forward, myfft.py:78-128)
# ---------------------------------------------------------------------------
def fft2_planar(x: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    _check_planar('x', x)
    x = x.contiguous()
    B, _, H, W = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        scratch = torch.empty_like(x)
        _lib.check(_lib.lib().csmri_fft2(_ptr(x), _ptr(out), B, H, W, int(bool(inverse)),
                                         _ptr(scratch), _stream()))
    return out


# ---------------------------------------------------------------------------
# undersampling (cs.undersample + to_tensor_format, compressed_sensing.py:460-512)
# ---------------------------------------------------------------------------
def undersample(img: torch.Tensor, rows: torch.Tensor, with_plan: bool = False):
    """img (B,H,W) float32 CUDA, rows (B,H) uint8 CUDA (1 = sampled line)
    -> dict(inp, kspace, mask, target), each (B,2,H,W) float32
    (the batch dict of scar_segmentation.py:212-218).  With ``with_plan`` also
    returns (dtab, addend): the noiseless DC plan of (kspace, mask), a by-product
    of the same kernels."""
    if img.dim() != 3 or img.dtype != torch.float32 or not img.is_cuda:
        raise ValueError('img must be a float32 CUDA tensor of shape (B,H,W)')
    B, H, W = img.shape
    if rows.shape != (B, H) or rows.dtype != torch.uint8 or rows.device != img.device:
        raise ValueError('rows must be a uint8 (B,H) tensor on the same device')
    img, rows = img.contiguous(), rows.contiguous()
    with torch.cuda.device(img.device):
        outs = [torch.empty((B, 2, H, W), dtype=torch.float32, device=img.device)
                for _ in range(5)]
        inp, kspace, mask, target, scratch = outs
        dtab = torch.empty((B, H), dtype=torch.float32, device=img.device) if with_plan else None
        addend = scratch if with_plan else None      # doubles as the intermediate
        _lib.check(_lib.lib().csmri_undersample(
            _ptr(img), _ptr(rows), _ptr(inp), _ptr(kspace), _ptr(mask), _ptr(target),
            _ptr(dtab), _ptr(addend), B, H, W, _ptr(scratch), _stream()))
    batch = {'inp': inp, 'kspace': kspace, 'mask': mask, 'target': target}
    if with_plan:
        return batch, (dtab, addend)
    return batch
