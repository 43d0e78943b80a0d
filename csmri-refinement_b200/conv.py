"""nn.Conv2d whose kernels come from libcsmri_dc.

RecNet's convolutions (models/recnet.py:37-48) stay what they are in the
reference - ``torch.nn.Conv2d`` modules with the same parameters, state-dict
keys and forward values.  What runs underneath on a B200:

* 32 -> 32 layers (zero padding 1, H % 8 == 0, W % 128 == 0): forward and data
  gradient through ``csmri_conv3x3_tc``, weight gradient through
  ``csmri_conv3x3_wgrad`` - the tcgen05 tensor cores with an error-compensated
  TF32 split (fp32-level accuracy: rel-L2 against float64 2.6e-7 / 4.7e-7,
  cuDNN's fp32 kernels 2.5e-7), bias + LeakyReLU fused into the forward kernel;
* RecNet's thin first / last layers (2 -> 32, 32 -> 2): everything through
  ``csmri_conv3x3_thin`` and the thin weight-gradient kernels;
* other channel counts that are multiples of 32: cuDNN forward / data gradient,
  the SIMT weight-gradient kernel (cuDNN's fp32 weight gradient was 64 % of the
  D5C5 step, profiles/r1_recnet_step_kernels.txt) and the fused epilogues.

Anything else (other kernel sizes, strides, dilations, channels_last, autocast,
CPU tensors) keeps torch's own kernels.
"""
import torch
from torch.autograd.function import once_differentiable
from torch import nn

from . import _lib

_ENABLED = True
_TC_ENABLED = True
_THIN_IN_BIAS = True      # 32 -> 2 layer: bias gradient from the weight-gradient kernel
# experimental: weight gradient on a side stream, concurrently with the data gradient
import os as _os
_WGRAD_STREAM = _os.environ.get('CSMRI_WGRAD_STREAM', '0') == '1'
_SIDE_STREAMS = {}


def _side_stream(device):
    s = _SIDE_STREAMS.get(device.index)
    if s is None:
        s = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device)
    return s


def set_tensor_core_conv(flag):
    """Switch the tcgen05 forward / data-gradient kernel of the 32 -> 32 layers on / off
    (off = cuDNN's fp32 kernels; A/B timing, tests)."""
    global _TC_ENABLED
    _TC_ENABLED = bool(flag)


def set_fast_wgrad(flag):
    """Switch the hand-written weight gradient on / off (A/B timing, tests)."""
    global _ENABLED
    _ENABLED = bool(flag)


def _require_cuda_f32(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError('the convolution kernels of libcsmri_dc need float32 CUDA tensors '
                               '(got %s on %s); there is no CPU implementation' % (t.dtype, t.device))


def conv3x3_wgrad(x, grad_out, pad):
    """dW (CO,CI,3,3) of a stride-1 3x3 convolution; x (N,CI,H+2-2*pad,W+2-2*pad),
    grad_out (N,CO,H,W), float32 CUDA."""
    _require_cuda_f32(x, grad_out)
    x, grad_out = x.contiguous(), grad_out.contiguous()
    n, ci = x.shape[0], x.shape[1]
    co, h, w = grad_out.shape[1], grad_out.shape[2], grad_out.shape[3]
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        dw = torch.empty((co, ci, 3, 3), dtype=torch.float32, device=x.device)
        ws = torch.empty((lib.csmri_conv3x3_wgrad_workspace_bytes(ci, co) // 4,),
                         dtype=torch.float32, device=x.device)
        _lib.check(lib.csmri_conv3x3_wgrad(
            x.data_ptr(), grad_out.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, ci, co, h, w,
            int(pad), torch.cuda.current_stream().cuda_stream))
    return dw


def _eligible(x, weight, stride, padding, dilation, groups):
    if not (_ENABLED and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32):
        return False
    if torch.is_autocast_enabled():
        return False          # fp16 / bf16 autocast: the kernels are fp32-only, torch handles it
    if x.dim() != 4 or not x.is_contiguous() or not weight.is_contiguous():
        return False          # e.g. channels_last: the kernels address plain NCHW
    if tuple(weight.shape[2:]) != (3, 3) or groups != 1:
        return False
    if tuple(stride) != (1, 1) or tuple(dilation) != (1, 1) or tuple(padding) not in ((0, 0), (1, 1)):
        return False
    co, ci = weight.shape[0], weight.shape[1]
    h = x.shape[2] - 2 + 2 * padding[0]
    w = x.shape[3] - 2 + 2 * padding[1]
    if h <= 0 or w <= 0 or w % 32 != 0:
        return False
    if x.shape[0] * max(ci, co) > 65535:        # grid.y of the epilogue kernels
        return False
    if (ci, co) in ((2, 32), (32, 2)):          # RecNet's first / last layer of a block
        return h % 16 == 0
    return ci % 32 == 0 and co % 32 == 0 and h % 4 == 0


def _aligned16(t):
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def bias_lrelu_(z, bias, slope):
    """z <- leaky_relu(z + bias[c], slope) in one pass, in place (z contiguous NCHW)."""
    _require_cuda_f32(z, bias)
    n, c, h, w = z.shape
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().csmri_bias_lrelu(z.data_ptr(), bias.data_ptr(), n, c, h, w,
                                               float(slope), torch.cuda.current_stream().cuda_stream))
    return z


def bias_lrelu_backward(grad_y, y, slope):
    """-> (grad_z, grad_bias) from the forward OUTPUT y, one pass over grad_y and y."""
    _require_cuda_f32(grad_y, y)
    grad_y = _aligned16(grad_y.contiguous())
    n, c, h, w = y.shape
    with torch.cuda.device(y.device):
        grad_z = torch.empty_like(y)
        grad_b = torch.empty((c,), dtype=torch.float32, device=y.device)
        partial = torch.empty((n * c * 8,), dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib().csmri_bias_lrelu_backward(
            grad_y.data_ptr(), y.data_ptr(), grad_z.data_ptr(), grad_b.data_ptr(),
            partial.data_ptr(), n, c, h, w, float(slope), torch.cuda.current_stream().cuda_stream))
    return grad_z, grad_b


def conv3x3_thin(x, weight, bias, slope=0.0):
    """RecNet's thin layers (2 -> 32 or 32 -> 2 channels, zero padding 1) through
    ``csmri_conv3x3_thin``: act(conv(x, weight) + bias), slope 0 = no activation."""
    _require_cuda_f32(x, weight, bias)
    x, weight = x.contiguous(), weight.contiguous()
    n, a, h, w = x.shape
    b = weight.shape[0]
    with torch.cuda.device(x.device):
        y = torch.empty((n, b, h, w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csmri_conv3x3_thin(
            x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
            y.data_ptr(), n, a, b, h, w, float(slope), torch.cuda.current_stream().cuda_stream))
    return y


def conv3x3_tc(x, weight, bias, slope=0.0, transpose_flip=False):
    """RecNet's 32 -> 32 layers (zero padding 1) through ``csmri_conv3x3_tc``
    (tcgen05 tensor cores, error-compensated TF32 split, fp32 accumulation):
    act(conv(x, weight) + bias); ``transpose_flip`` gives the data gradient of
    the same layer from the same weight tensor."""
    _require_cuda_f32(x, weight, bias)
    x, weight = x.contiguous(), weight.contiguous()
    n, c, h, w = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        _lib.check(_lib.lib().csmri_conv3x3_tc(
            x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
            y.data_ptr(), n, c, h, w, float(slope), int(bool(transpose_flip)),
            torch.cuda.current_stream().cuda_stream))
    return y


def conv3x3_tc_signs(x, weight, bias, slope, want_in_signs=False):
    """``conv3x3_tc`` forward that also returns the sign word of every output pixel
    ((N,H,W) int32, bit c = output channel c > 0) for :func:`conv3x3_tc_masked` and, with
    ``want_in_signs``, the same word for the input ``x`` (else ``None``)."""
    _require_cuda_f32(x, weight, bias)
    x, weight = x.contiguous(), weight.contiguous()
    n, c, h, w = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        signs = torch.empty((n, h, w), dtype=torch.int32, device=x.device)
        in_signs = torch.empty((n, h, w), dtype=torch.int32, device=x.device) if want_in_signs else None
        _lib.check(_lib.lib().csmri_conv3x3_tc_signs(
            x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
            y.data_ptr(), signs.data_ptr(), None if in_signs is None else in_signs.data_ptr(),
            n, c, h, w, float(slope), torch.cuda.current_stream().cuda_stream))
    return y, signs, in_signs


def conv3x3_tc_masked(x, weight, signs, act_slope, transpose_flip=True):
    """``conv(x, weight)[:, c] * (bit c of signs ? 1 : act_slope)`` through
    ``csmri_conv3x3_tc_masked``: with ``transpose_flip`` the data gradient of a 32 -> 32
    layer followed by the backward of the LeakyReLU that produced that layer's input, whose
    output signs :func:`conv3x3_tc_signs` recorded - in one pass."""
    _require_cuda_f32(x, weight)
    x, weight = x.contiguous(), weight.contiguous()
    n, c, h, w = x.shape
    if signs.dtype != torch.int32 or tuple(signs.shape) != (n, h, w) or not signs.is_cuda \
            or not signs.is_contiguous():
        raise RuntimeError('signs must be a contiguous CUDA int32 tensor of shape (N, H, W)')
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        _lib.check(_lib.lib().csmri_conv3x3_tc_masked(
            x.data_ptr(), weight.data_ptr(), signs.data_ptr(), y.data_ptr(), n, c, h, w,
            float(act_slope), int(bool(transpose_flip)), torch.cuda.current_stream().cuda_stream))
    return y


def conv3x3_wgrad_bias(x, grad_out):
    """(dW, db) of a 32 -> 32 layer with zero padding 1 on the tensor-core path
    (H % 16 == 0, W % 64 == 0): the bias gradient is a by-product of the kernel."""
    _require_cuda_f32(x, grad_out)
    x, grad_out = _aligned16(x.contiguous()), _aligned16(grad_out.contiguous())
    n, _, h, w = grad_out.shape
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        dw = torch.empty((32, 32, 3, 3), dtype=torch.float32, device=x.device)
        db = torch.empty((32,), dtype=torch.float32, device=x.device)
        ws = torch.empty((lib.csmri_conv3x3_wgrad_workspace_bytes(32, 32) // 4,),
                         dtype=torch.float32, device=x.device)
        _lib.check(lib.csmri_conv3x3_wgrad_bias(
            x.data_ptr(), grad_out.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, h, w,
            torch.cuda.current_stream().cuda_stream))
    return dw, db


def conv3x3_wgrad_thin_bias(x, grad_out):
    """(dW, db) of RecNet's 2 -> 32 layer (zero padding 1): the bias gradient is a
    by-product of the weight-gradient kernel."""
    _require_cuda_f32(x, grad_out)
    x, grad_out = _aligned16(x.contiguous()), _aligned16(grad_out.contiguous())
    n, _, h, w = grad_out.shape
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        dw = torch.empty((32, 2, 3, 3), dtype=torch.float32, device=x.device)
        db = torch.empty((32,), dtype=torch.float32, device=x.device)
        ws = torch.empty((lib.csmri_conv3x3_wgrad_workspace_bytes(2, 32) // 4,),
                         dtype=torch.float32, device=x.device)
        _lib.check(lib.csmri_conv3x3_wgrad_thin_bias(
            x.data_ptr(), grad_out.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, h, w,
            torch.cuda.current_stream().cuda_stream))
    return dw, db


def conv3x3_wgrad_thin_in_bias(x, grad_out):
    """(dW, db) of RecNet's 32 -> 2 layer (zero padding 1), bias gradient as a by-product."""
    _require_cuda_f32(x, grad_out)
    x, grad_out = _aligned16(x.contiguous()), _aligned16(grad_out.contiguous())
    n, _, h, w = grad_out.shape
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        dw = torch.empty((2, 32, 3, 3), dtype=torch.float32, device=x.device)
        db = torch.empty((2,), dtype=torch.float32, device=x.device)
        ws = torch.empty((lib.csmri_conv3x3_wgrad_workspace_bytes(32, 2) // 4,),
                         dtype=torch.float32, device=x.device)
        _lib.check(lib.csmri_conv3x3_wgrad_thin_in_bias(
            x.data_ptr(), grad_out.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, h, w,
            torch.cuda.current_stream().cuda_stream))
    return dw, db


def _is_tc(x, weight, pad):
    """Shapes csmri_conv3x3_tc covers: 32 -> 32 channels, padding 1, H % 8 == 0, W % 128 == 0."""
    return (_TC_ENABLED and pad == 1 and tuple(weight.shape[:2]) == (32, 32) and
            x.shape[2] % 8 == 0 and x.shape[3] % 128 == 0)


def conv3x3_thin_masked(x, weight, signs, act_slope):
    """2 -> 32 thin convolution (no bias) times the LeakyReLU derivative selected by
    ``signs`` ((N,H,W) int32 from :func:`conv3x3_tc_signs`): the data gradient of RecNet's
    32 -> 2 layer (``weight`` = its weights flipped and transposed) and the backward of the
    activation in front of it, in one pass."""
    _require_cuda_f32(x, weight)
    x, weight = x.contiguous(), weight.contiguous()
    n, a, h, w = x.shape
    if a != 2 or tuple(weight.shape) != (32, 2, 3, 3):
        raise RuntimeError('conv3x3_thin_masked is the 2 -> 32 form (got %s, %s)'
                           % (tuple(x.shape), tuple(weight.shape)))
    if signs.dtype != torch.int32 or tuple(signs.shape) != (n, h, w) or not signs.is_cuda \
            or not signs.is_contiguous():
        raise RuntimeError('signs must be a contiguous CUDA int32 tensor of shape (N, H, W)')
    with torch.cuda.device(x.device):
        y = torch.empty((n, 32, h, w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csmri_conv3x3_thin_masked(
            x.data_ptr(), weight.data_ptr(), signs.data_ptr(), y.data_ptr(), n, h, w,
            float(act_slope), torch.cuda.current_stream().cuda_stream))
    return y


def conv3x3_thin_dgrad(grad_out, weight, signs=None, act_slope=0.0):
    """Data gradient of a thin layer (2 -> 32 or 32 -> 2, padding 1) from the layer's own
    weights - no flipped / transposed copy.  ``signs`` (32 -> 2 layer only): also apply the
    derivative of the LeakyReLU(act_slope) that produced the layer's input."""
    _require_cuda_f32(grad_out, weight)
    grad_out, weight = _aligned16(grad_out.contiguous()), weight.contiguous()
    n, co, h, w = grad_out.shape
    ci = weight.shape[1]
    if tuple(weight.shape) != (co, ci, 3, 3) or (ci, co) not in ((2, 32), (32, 2)):
        raise RuntimeError('conv3x3_thin_dgrad covers the 2 -> 32 and 32 -> 2 layers (got %s, %s)'
                           % (tuple(grad_out.shape), tuple(weight.shape)))
    if signs is not None and (signs.dtype != torch.int32 or tuple(signs.shape) != (n, h, w)
                              or not signs.is_cuda or not signs.is_contiguous()):
        raise RuntimeError('signs must be a contiguous CUDA int32 tensor of shape (N, H, W)')
    with torch.cuda.device(grad_out.device):
        dx = torch.empty((n, ci, h, w), dtype=torch.float32, device=grad_out.device)
        _lib.check(_lib.lib().csmri_conv3x3_thin_dgrad(
            grad_out.data_ptr(), weight.data_ptr(), signs.data_ptr() if signs is not None else None,
            dx.data_ptr(), n, ci, co, h, w, float(act_slope), torch.cuda.current_stream().cuda_stream))
    return dx


def _is_thin(weight, pad):
    return pad == 1 and (weight.shape[1], weight.shape[0]) in ((2, 32), (32, 2))


class _Conv3x3(torch.autograd.Function):
    """conv (+ bias) [+ LeakyReLU when ``slope`` is given].  32 -> 32 layers and the
    thin layers (2 -> 32, 32 -> 2): everything runs through libcsmri_dc.  Other
    32k -> 32k layers: forward and data gradient stay cuDNN, epilogues and the
    weight gradient run through libcsmri_dc."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad, slope, in_signs=None, in_slope=0.0, out_premasked=False):
        # out_premasked (thin 2 -> 32 layer with fused LeakyReLU only): the consumer (_TcChain)
        # sends the gradient back already multiplied by this layer's activation derivative
        ctx.out_premasked = bool(out_premasked)
        # in_signs (thin 32 -> 2 layer only): sign words of x, the output of a LeakyReLU whose
        # producer (_TcChain) expects its incoming gradient already multiplied by the derivative
        ctx.in_signs, ctx.in_slope = in_signs, in_slope
        ctx.pad, ctx.slope = pad, slope
        ctx.has_bias = bias is not None
        ctx.thin = _is_thin(weight, pad) and (slope is None or weight.shape[1] == 2)
        ctx.tc = _is_tc(x, weight, pad)
        if ctx.thin:
            y = conv3x3_thin(x, weight, bias, slope or 0.0)
        elif ctx.tc:
            y = conv3x3_tc(x, weight, bias, slope or 0.0)
        elif slope is None:
            y = torch.ops.aten.convolution(x, weight, bias, [1, 1], [pad, pad], [1, 1], False,
                                           [0, 0], 1)
        else:
            y = torch.ops.aten.convolution(x, weight, None, [1, 1], [pad, pad], [1, 1], False,
                                           [0, 0], 1)
            bias_lrelu_(y, bias, slope)
        if slope is None:
            ctx.save_for_backward(x, weight)
        else:
            ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    @once_differentiable     # raw kernels on data pointers: a second derivative would be silently wrong
    def backward(ctx, grad_out):
        pad = ctx.pad
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        grad_out = grad_out.contiguous()
        gx = gb = gw = None
        if ctx.slope is None:
            x, weight = ctx.saved_tensors
        elif ctx.out_premasked:
            x, weight, y = ctx.saved_tensors
            gw, gb = conv3x3_wgrad_thin_bias(x, grad_out) if need_w or need_b else (None, None)
            if need_x:
                gx = conv3x3_thin_dgrad(grad_out, weight)
            return gx, gw if need_w else None, gb if need_b else None, None, None, None, None, None
        else:
            x, weight, y = ctx.saved_tensors
            grad_out, gb = bias_lrelu_backward(grad_out, y, ctx.slope)
            if not need_b:
                gb = None
        fork = None
        if need_w and _WGRAD_STREAM and not ctx.thin:
            # fork here, before the data gradient is launched on the current stream
            cur = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gw = conv3x3_wgrad(x, grad_out, pad)
            fork = (cur, side)
        if ctx.thin:
            if need_b and gb is None and need_w and weight.shape[0] == 2 and _THIN_IN_BIAS \
                    and x.shape[2] % 8 == 0 and x.shape[3] % 32 == 0:
                gw, gb = conv3x3_wgrad_thin_in_bias(x, grad_out)   # db rides on the weight gradient
                need_w = False
            elif need_b and gb is None:
                gb = grad_out.sum(dim=(0, 2, 3))
            if need_x and ctx.in_signs is not None:
                gx = conv3x3_thin_dgrad(grad_out, weight, ctx.in_signs, ctx.in_slope)
            elif need_x:   # the same kernel family on flipped, transposed weights
                gx = conv3x3_thin_dgrad(grad_out, weight)
        elif ctx.tc:
            if need_b and gb is None:
                gb = grad_out.sum(dim=(0, 2, 3))
            if need_x:     # the same kernel with the operator transposed and mirrored
                gx = conv3x3_tc(grad_out, weight, None, 0.0, transpose_flip=True)
        elif ctx.slope is None:
            if need_x or need_b:
                gx, _, gb = torch.ops.aten.convolution_backward(
                    grad_out, x, weight, [weight.shape[0]], [1, 1], [pad, pad], [1, 1], False,
                    [0, 0], 1, [need_x, False, need_b])
        elif need_x:
            gx = torch.ops.aten.convolution_backward(
                grad_out, x, weight, None, [1, 1], [pad, pad], [1, 1], False, [0, 0], 1,
                [True, False, False])[0]
        if fork is not None:
            cur, side = fork
            cur.wait_stream(side)
            gw.record_stream(cur)
            grad_out.record_stream(side)
        elif need_w and gw is None:
            gw = conv3x3_wgrad(x, grad_out, pad)
        return gx, gw, gb, None, None, None, None, None


class _TcChain(torch.autograd.Function):
    """A run of consecutive 32 -> 32 layers, each followed by LeakyReLU(slope), as ONE
    autograd node (models/recnet.py:37-47: the inner layers of a ConvBlock).  Forward is
    the fused conv + bias + LeakyReLU kernel per layer.  Backward applies the activation
    derivative of layer k inside the data-gradient kernel of layer k + 1 (its saved input IS
    that activation's output) and takes the bias gradients from the weight-gradient kernel.
    The activation signs travel as one 32-bit word per pixel written by the forward kernel.
    The derivative of the run's LAST activation is a pass of its own unless the consumer
    applies it (``out_premasked``: the block's 32 -> 2 layer does, in its data gradient), and
    the gradient returned for ``x`` carries the derivative of the activation that produced
    ``x`` when ``in_premask`` is set (the block's 2 -> 32 layer then skips its own pass): see
    :func:`tc_chain`."""

    @staticmethod
    def forward(ctx, x, slope, out_premasked, in_premask, *params):
        # out_premasked: the consumer of the output promises to multiply the gradient it sends
        # back by the last activation's derivative (it gets the sign words for that);
        # in_premask: x is itself the output of a LeakyReLU(slope) whose producer expects the
        # gradient returned for x to carry that derivative (the first kernel records x's signs)
        ws, bs = params[0::2], params[1::2]
        acts, signs, in_signs = [x], [], None
        for k, (w, b) in enumerate(zip(ws, bs)):
            y, sg, isg = conv3x3_tc_signs(acts[-1], w, b, slope, want_in_signs=bool(in_premask) and k == 0)
            if k == 0:
                in_signs = isg
            signs.append(sg)
            acts.append(y)
        ctx.slope, ctx.n, ctx.out_premasked = slope, len(ws), bool(out_premasked)
        ctx.in_signs = in_signs
        ctx.save_for_backward(*acts, *ws, *signs[:-1])
        ctx.mark_non_differentiable(signs[-1])
        return acts[-1], signs[-1]

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_signs):
        n, slope = ctx.n, ctx.slope
        saved = ctx.saved_tensors
        acts, ws, signs = saved[:n + 1], saved[n + 1:2 * n + 1], saved[2 * n + 1:]
        grads = [None] * (2 * n)
        if ctx.out_premasked:                   # the consumer applied the last activation's derivative
            gz, gb, have_gb = grad_out.contiguous(), None, False
        else:                                   # a pass of its own (+ bias gradient)
            gz, gb = bias_lrelu_backward(grad_out.contiguous(), acts[n], slope)
            have_gb = True
        for k in range(n - 1, -1, -1):          # layer k: input acts[k], output acts[k + 1]
            need_w, need_b = ctx.needs_input_grad[4 + 2 * k], ctx.needs_input_grad[5 + 2 * k]
            fresh = k < n - 1 or not have_gb    # gb of this layer not known yet
            if need_w and fresh:
                grads[2 * k], gb = conv3x3_wgrad_bias(acts[k], gz)
            elif need_w:
                grads[2 * k] = conv3x3_wgrad(acts[k], gz, 1)
            elif need_b and fresh:
                gb = gz.sum(dim=(0, 2, 3))
            if need_b:
                grads[2 * k + 1] = gb
            if k > 0:                           # data gradient + derivative of layer k - 1's activation
                gz = conv3x3_tc_masked(gz, ws[k], signs[k - 1], slope)
            elif ctx.needs_input_grad[0] and ctx.in_signs is not None:
                gz = conv3x3_tc_masked(gz, ws[0], ctx.in_signs, slope)
            elif ctx.needs_input_grad[0]:
                gz = conv3x3_tc(gz, ws[0], None, 0.0, transpose_flip=True)
            else:
                gz = None
        return (gz, None, None, None) + tuple(grads)


def tc_chain_eligible(x, convs):
    """True when every module of ``convs`` is a 32 -> 32 ``Conv2d`` with bias, padding 1
    and the same fused LeakyReLU slope, and ``x`` has a shape both tensor-core kernels
    cover (H % 16 == 0, W % 128 == 0)."""
    if len(convs) < 1 or not (_ENABLED and _TC_ENABLED) or torch.is_autocast_enabled():
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()):
        return False
    if x.shape[1] != 32 or x.shape[2] % 16 != 0 or x.shape[3] % 128 != 0:
        return False
    slope = convs[0].fused_slope
    for m in convs:
        if not isinstance(m, Conv2d) or m.bias is None or m.fused_slope is None or \
                m.fused_slope != slope or not slope > 0:
            return False
        if tuple(m.weight.shape) != (32, 32, 3, 3) or m.padding_mode != 'zeros' or \
                isinstance(m.padding, str) or tuple(m.padding) != (1, 1) or \
                tuple(m.stride) != (1, 1) or tuple(m.dilation) != (1, 1) or m.groups != 1:
            return False
        if not (m.weight.is_contiguous() and m.weight.dtype == torch.float32):
            return False
    return True


def _plain_thin(m, cin, cout, with_act):
    return isinstance(m, Conv2d) and m.bias is not None and tuple(m.weight.shape) == (cout, cin, 3, 3) and \
        m.padding_mode == 'zeros' and not isinstance(m.padding, str) and tuple(m.padding) == (1, 1) and \
        tuple(m.stride) == (1, 1) and tuple(m.dilation) == (1, 1) and m.groups == 1 and \
        m.weight.is_contiguous() and m.weight.dtype == torch.float32 and \
        ((m.fused_slope is not None and m.fused_slope > 0) if with_act else m.fused_slope is None)


def tc_chain(x, convs, last=None, first=None):
    """Run ``convs`` (see :func:`tc_chain_eligible`) as one fused autograd node.
    ``last`` (optional): the 32 -> 2 ``Conv2d`` that consumes the result - if it is the plain
    thin layer (padding 1, bias, no activation) it is applied too, and its data gradient carries
    the derivative of the run's last LeakyReLU.  ``first`` (optional): the 2 -> 32 ``Conv2d``
    with the same fused LeakyReLU that PRODUCES the run's input - then ``x`` is that layer's
    input, it is applied first, and the run's data gradient carries its activation derivative.
    With both, no LeakyReLU-backward pass is left in the block."""
    params = []
    for m in convs:
        params += [m.weight, m.bias]
    slope = float(convs[0].fused_slope)
    fuse_first = first is not None and _plain_thin(first, 2, 32, True) and float(first.fused_slope) == slope \
        and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    if first is not None:
        x = _Conv3x3.apply(x, first.weight, first.bias, 1, slope, None, 0.0, True) if fuse_first else first(x)
    fuse_last = last is not None and _plain_thin(last, 32, 2, False)
    y, signs = _TcChain.apply(x, slope, fuse_last, fuse_first, *params)
    if last is None:
        return y
    if fuse_last:
        return _Conv3x3.apply(y, last.weight, last.bias, 1, None, signs, slope)
    return last(y)


class Conv2d(nn.Conv2d):
    """Drop-in ``nn.Conv2d``: same constructor, parameters and forward values.
    ``fused_slope`` (set by the owner, not a constructor argument) makes the
    module apply the LeakyReLU that follows it in RecNet's blocks, so that bias
    add and activation are one pass in both directions."""

    fused_slope = None

    def forward(self, x):
        slope = self.fused_slope
        if self.padding_mode == 'zeros' and not isinstance(self.padding, str) and \
                _eligible(x, self.weight, self.stride, self.padding, self.dilation, self.groups) \
                and (slope is None or (self.bias is not None and slope > 0)):
            pad = int(self.padding[0])
            if torch.is_grad_enabled() and (self.weight.requires_grad or x.requires_grad):
                return _Conv3x3.apply(x, self.weight, self.bias, pad, slope)
            # inference: same kernels, nothing saved
            if _is_thin(self.weight, pad) and (slope is None or self.weight.shape[1] == 2):
                return conv3x3_thin(x, self.weight, self.bias, slope or 0.0)
            if _is_tc(x, self.weight, pad):
                return conv3x3_tc(x, self.weight, self.bias, slope or 0.0)
            if slope is not None:
                y = torch.ops.aten.convolution(x, self.weight, None, [1, 1], [pad, pad], [1, 1],
                                               False, [0, 0], 1)
                return bias_lrelu_(y, self.bias, slope)
        out = super(Conv2d, self).forward(x)
        if slope is not None:
            out = nn.functional.leaky_relu(out, slope, inplace=True)
        return out
