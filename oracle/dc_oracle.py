"""CPU oracle for the k-space data-consistency (DC) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product path
(``csmri_refinement_b200``) never imports this module and fails loudly when the
CUDA library is missing.

It restates, in numpy / ``torch.fft``, the arithmetic of these reference files
(all paths relative to ``/root/reference``):

* ``data/reconstruction/deep_med_lib/my_pytorch/myfft.py:131-142``  blend
* ``data/reconstruction/deep_med_lib/my_pytorch/myfft.py:145-163``  DC perform
* ``data/reconstruction/deep_med_lib/my_pytorch/myfft.py:78-128``   ortho FFT2/iFFT2
  (third-party ``pytorch-fft==0.14`` -> cuFFT C2C; absent from the tree, its
  convention is the one ``myfft.py:225,241-242`` pins: numpy's ``norm='ortho'``)
* ``data/reconstruction/deep_med_lib/utils/compressed_sensing.py:82-123``  cartesian_mask
* ``data/reconstruction/deep_med_lib/utils/compressed_sensing.py:460-512`` undersample
* ``data/reconstruction/deep_med_lib/utils/compressed_sensing.py:515-529`` numpy DC
* ``data/reconstruction/deep_med_lib/utils/dnn_io.py:4-22,47-61``          2-channel packing
* ``data/reconstruction/deep_med_lib/my_pytorch/myImageTransformations.py:1215-1238`` Undersample group
* ``data/reconstruction/deep_med_lib/my_pytorch/myImageTransformations.py:105-117,935-954`` CenterCropInKspace
* ``utils/tensor_transforms.py:62-75``, ``data/reconstruction/rec_transforms.py:79-85``,
  ``metrics/image_metrics.py:7-19``  complex_abs / output_transform / PSNR

Parity pinning: the reference ships no golden vectors and its torch DC cannot
run (CUDA-only ``pytorch_fft``, legacy autograd Functions).  The oracle is
pinned instead against outputs of the reference's own importable numpy
functions and its pure ``myfft.data_consistency`` blend, generated in the build
container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# numpy side
# --------------------------------------------------------------------------


def blend_np(k, k0, mask, noise_lvl=None):
    """k-space blend, myfft.py:131-142 (same formula, same truthiness test).

    ``noise_lvl`` of ``None`` *or* ``0`` selects the noiseless branch, exactly
    like the reference's ``if v:``.
    """
    v = noise_lvl
    if v:
        return (1 - mask) * k + mask * (k + v * k0) / (1 + v)
    return (1 - mask) * k + k0


def planar_to_complex(x):
    """(B,2,H,W) planar real -> (B,H,W) complex; channel 0 = Re, 1 = Im
    (myfft.py:159: ``x[:,0:1], x[:,1:2]``; dnn_io.py:16-22)."""
    return x[:, 0] + 1j * x[:, 1]


def complex_to_planar(z, dtype=np.float32):
    """(B,H,W) complex -> (B,2,H,W) planar real of ``dtype``
    (dnn_io.complex2real, dnn_io.py:4-22)."""
    return np.stack([z.real, z.imag], axis=1).astype(dtype)


def dc_perform_np(x, k0, mask, noise_lvl=None, dtype=np.float64):
    """DataConsistencyInKspace.perform restated with numpy (myfft.py:153-163).

    ``x, k0, mask``: (B,2,H,W) planar.  The blend is applied per channel with
    the per-channel mask, as the reference does on the concatenated planes.
    Computation is carried out in ``dtype`` (float64 by default: the
    high-precision arm of the oracle).
    """
    x = np.asarray(x, dtype=dtype)
    k0 = np.asarray(k0, dtype=dtype)
    mask = np.asarray(mask, dtype=dtype)
    kc = np.fft.fft2(planar_to_complex(x), norm='ortho')
    k = np.stack([kc.real, kc.imag], axis=1)
    out = blend_np(k, k0, mask, noise_lvl)
    xc = np.fft.ifft2(planar_to_complex(out), norm='ortho')
    return np.stack([xc.real, xc.imag], axis=1).astype(dtype)


def dc_diag_np(mask, noise_lvl=None):
    """Diagonal D of the DC Jacobian F^-1 D F: autograd of
    myfft.py:139/141 w.r.t. k."""
    v = noise_lvl
    if v:
        return (1 - mask) + mask / (1 + v)
    return 1 - mask


def dc_adjoint_np(g, mask, noise_lvl=None, dtype=np.float64):
    """Vector-Jacobian product of dc_perform w.r.t. x (myfft.py:92-102,119-128
    chained with the blend's autograd): gx = iFFT2o(D * FFT2o(g))."""
    g = np.asarray(g, dtype=dtype)
    d = dc_diag_np(np.asarray(mask, dtype=dtype), noise_lvl)
    gc = np.fft.fft2(planar_to_complex(g), norm='ortho')
    gk = np.stack([gc.real, gc.imag], axis=1) * d
    xc = np.fft.ifft2(planar_to_complex(gk), norm='ortho')
    return np.stack([xc.real, xc.imag], axis=1).astype(dtype)


def cs_data_consistency_np(x, y, mask):
    """numpy DC on complex arrays, compressed_sensing.py:515-529 with
    ``centered=False, norm='ortho'`` (the only in-tree CPU DC)."""
    xf = np.fft.fft2(x, norm='ortho')
    return np.fft.ifft2((1 - mask) * xf + y, norm='ortho')


def normal_pdf(length, sensitivity):
    """compressed_sensing.py:13-14."""
    return np.exp(-sensitivity * (np.arange(length) - length / 2) ** 2)


def cartesian_lines(n_img, nx, acc, sample_n=8, rng=None):
    """Sampled phase-encode lines BEFORE ifftshift: (n_img, nx) float64 in {0,1}.

    Restates the index selection of compressed_sensing.py:93-113: Gaussian pdf
    plus a uniform floor, centre ``sample_n`` lines forced, and one
    ``rng.choice(nx, n_lines, replace=False, p=pdf)`` draw per image.  The RNG
    call sequence is identical, so with the same ``RandomState`` the indices
    are bit-identical.
    """
    if rng is None:
        rng = np.random
    pdf = normal_pdf(nx, 0.5 / (nx / 10.) ** 2)
    lmda = nx / (2. * acc)
    n_lines = nx // acc
    pdf += lmda * 1. / nx
    lo, hi = nx // 2 - sample_n // 2, nx // 2 + sample_n // 2
    if sample_n:
        pdf[lo:hi] = 0
        pdf /= np.sum(pdf)
        n_lines -= sample_n
    rows = np.zeros((n_img, nx))
    for i in range(n_img):
        idx = rng.choice(nx, int(n_lines), False, pdf)
        rows[i, idx] = 1
    if sample_n:
        rows[:, lo:hi] = 1
    return rows


def cartesian_mask(shape, acc, sample_n=10, centred=False, rng=None):
    """compressed_sensing.py:82-123: lines broadcast along the last axis, then
    ``ifftshift`` over the last two axes unless ``centred``."""
    n_img, nx, ny = int(np.prod(shape[:-2])), shape[-2], shape[-1]
    rows = cartesian_lines(n_img, nx, acc, sample_n, rng)
    mask = np.broadcast_to(rows[:, :, None], (n_img, nx, ny)).reshape(shape)
    if not centred:
        mask = np.fft.ifftshift(mask, axes=(-1, -2))
    return np.ascontiguousarray(mask)


def undersample(x, mask, norm='ortho', noise=0, rng=None):
    """compressed_sensing.py:460-512, ``centred=False`` branch.

    The two ``rng.normal`` draws happen even when ``noise == 0`` (:494-495);
    they are kept so that the RNG stream stays in step with the reference.
    Returns ``(x_u, x_fu)`` complex128.
    """
    if rng is None:
        rng = np.random
    assert x.shape == mask.shape
    nz = np.sqrt(.5) * (rng.normal(0, 1, x.shape) + 1j * rng.normal(0, 1, x.shape))
    nz = nz * np.sqrt(noise)
    if norm == 'ortho':
        nz = nz * np.sqrt(np.prod(mask.shape[-2:]))
    else:
        nz = nz * np.prod(mask.shape[-2:])
    x_f = np.fft.fft2(x, norm=norm)
    x_fu = mask * (x_f + nz)
    x_u = np.fft.ifft2(x_fu, norm=norm)
    return x_u, x_fu


def to_tensor_format(x, mask=False):
    """dnn_io.py:47-61 for (n,nx,ny) inputs: complex -> (n,2,nx,ny) float32;
    a mask is first multiplied by (1+1j) so both channels carry it."""
    if mask:
        x = x * (1 + 1j)
    return complex_to_planar(np.asarray(x), np.float32)


def undersample_group(image, mask, rng=None):
    """myImageTransformations.py:1215-1238 ``Undersample.__call__`` for one
    (nx,ny,1) real image and a (1,nx,ny) mask: returns (nx,ny,8) float32 with
    channels [inp(2), kspace(2), mask(2), target(2)]
    (scar_segmentation.py:212-218 splits them in that order)."""
    image = image.transpose((2, 0, 1))
    im_und, k_und = undersample(image, mask, norm='ortho', rng=rng)
    grp = np.concatenate([to_tensor_format(im_und), to_tensor_format(k_und),
                          to_tensor_format(mask, mask=True),
                          to_tensor_format(image)], axis=1)
    return grp.squeeze().transpose((1, 2, 0))


def center_crop_in_kspace(img, size):
    """myImageTransformations.CenterCropInKspace (:935-954) for one (nx,ny[,c])
    array: centred FFT2 over axes (0,1), crop_image_at (:105-117) around
    (nx//2, ny//2), centred inverse, magnitude."""
    sx, sy = (size, size) if np.isscalar(size) else size
    nx, ny = img.shape[:2]
    ax = (0, 1)
    k = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(img, axes=ax), norm='ortho', axes=ax), axes=ax)
    cx, cy, r1, r2 = nx // 2, ny // 2, sx // 2, sy // 2
    x1, x2, y1, y2 = cx - r1, cx + r1, cy - r2, cy + r2
    x1_, x2_, y1_, y2_ = max(x1, 0), min(x2, nx), max(y1, 0), min(y2, ny)
    crop = k[x1_:x2_, y1_:y2_]
    crop = np.pad(crop, ((x1_ - x1, x2 - x2_), (y1_ - y1, y2 - y2_)) + ((0, 0),) * (crop.ndim - 2),
                  'constant')
    out = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(crop, axes=ax), norm='ortho', axes=ax),
                          axes=ax)
    return abs(out)


def complex_abs_np(x):
    """utils/tensor_transforms.py:62-75 on a (B,2,H,W) float32 array -> (B,1,H,W),
    with torch's operation order (separately rounded squares, sum, sqrt)."""
    x = np.asarray(x, dtype=np.float32)
    return np.sqrt(x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1], dtype=np.float32)[:, None]


def output_transform_np(pred, target):
    """rec_transforms.py:79-85."""
    return (np.clip(complex_abs_np(pred), 0.0, 1.0), np.clip(complex_abs_np(target), 0.0, 1.0))


def psnr_np(pred, target):
    """metrics/image_metrics.py:7-19 on output_transform(pred, target)."""
    p, t = output_transform_np(pred, target)
    mse = np.mean((p.astype(np.float64) - t.astype(np.float64)) ** 2)
    return 10.0 * np.log10(1.0 / mse)


# --------------------------------------------------------------------------
# torch side (fp32/fp64, differentiable; also the timed CPU baseline)
# --------------------------------------------------------------------------


def dc_perform_torch(x, k0, mask, noise_lvl=None):
    """torch.fft restatement of myfft.py:153-163 with ``norm='ortho'``
    (myfft.py:86-89,113-116).  Differentiable w.r.t. ``x`` through torch's own
    autograd, which reproduces myfft.py:92-102,119-128."""
    import torch
    xc = torch.complex(x[:, 0], x[:, 1])
    kc = torch.fft.fft2(xc, norm='ortho')
    k = torch.stack([kc.real, kc.imag], dim=1)
    v = noise_lvl
    if v:
        out = (1 - mask) * k + mask * (k + v * k0) / (1 + v)
    else:
        out = (1 - mask) * k + k0
    oc = torch.fft.ifft2(torch.complex(out[:, 0], out[:, 1]), norm='ortho')
    return torch.stack([oc.real, oc.imag], dim=1)


class OracleDataConsistencyInKspace(object):
    """Interface twin of myfft.DataConsistencyInKspace (myfft.py:145-163) on
    top of :func:`dc_perform_torch`; used by tests to drive RecNet on CPU."""

    def __init__(self, noise_lvl=None, norm='ortho'):
        assert norm == 'ortho'
        self.noise_lvl = noise_lvl

    def perform(self, x, k0, mask):
        return dc_perform_torch(x, k0, mask, self.noise_lvl)


def rel_l2(a, b):
    """||a-b|| / ||b|| in float64."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


# ---------------------------------------------------------------------------
# refinement-path pointwise ops
# ---------------------------------------------------------------------------
def scale_np(x):
    """models/refinement_wrapper.py:51-73 ``_scale``: per example and channel,
    (x - min) / max(x - min) * 2 - 1, evaluated op by op in float32.
    -> (scaled (B,C,H,W), minimum (B,C,1), maximum (B,C,1))"""
    x = np.asarray(x, dtype=np.float32)
    b, c, h, w = x.shape
    out = x.reshape(b, c, h * w)
    minimum = out.min(axis=2, keepdims=True)
    out = out - minimum
    maximum = out.max(axis=2, keepdims=True)
    out = out / maximum
    out = out * np.float32(2) - np.float32(1)
    return out.reshape(b, c, h, w), minimum, maximum


def unscale_np(t, minimum, maximum):
    """models/refinement_wrapper.py:76-92 ``_unscale``: ((t + 1) / 2) * max + min."""
    t = np.asarray(t, dtype=np.float32)
    b, c, h, w = t.shape
    out = t.reshape(b, c, h * w)
    out = (out + np.float32(1)) / np.float32(2)
    out = out * maximum + minimum
    return out.reshape(b, c, h, w)


def magnitude_image_np(x):
    """utils/tensor_transforms.py:78-99: complex_abs, then min-max to (0, 1)."""
    m = complex_abs_np(np.asarray(x, dtype=np.float32))
    b, c, h, w = m.shape
    out = m.reshape(b, c, h * w)
    minimum = out.min(axis=2, keepdims=True)
    out = out - minimum
    maximum = out.max(axis=2, keepdims=True)
    return (out / maximum).reshape(b, c, h, w)


def refinement_real_penalty_add_torch(out_pretrained, out_learnable, scale):
    """models/refinement_wrapper.py:173-197 with the learnable model's output
    given: the real channel of the (detached) pretrained output is scaled to
    (-1, 1), ``scale * out_learnable`` is added, the sum is mapped back with the
    same min / max; the imaginary channel passes through.  torch, so that
    autograd supplies the reference gradients."""
    import torch
    real = out_pretrained[:, 0].unsqueeze(1).contiguous()
    imag = out_pretrained[:, 1].unsqueeze(1).contiguous()
    b, c, h, w = real.shape
    out = real.view(b, c, h * w)
    minimum, _ = out.min(dim=2, keepdim=True)
    out = out - minimum
    maximum, _ = out.max(dim=2, keepdim=True)
    out = out / maximum
    scaled = (out * 2 - 1).view(b, c, h, w)
    refined = scaled + scale * out_learnable
    o = refined.view(b, c, h * w)
    o = (o + 1) / 2
    o = o * maximum + minimum
    return torch.cat((o.view(b, c, h, w), imag), dim=1)
