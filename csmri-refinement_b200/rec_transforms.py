"""GPU mirror of data/reconstruction/rec_transforms.py (the loader tail and the
reporting transform around the DC path).

* ``center_crop_in_kspace``  myImageTransformations.CenterCropInKspace (:935-954):
  fft2c -> crop at the centre (``crop_image_at`` :105-117) -> ifft2c -> abs.
  The FFTs are the library's ``csmri_fft2``; the shifts, the crop and its zero
  padding are composed into one gather pass per stage (``csmri_shift_crop``,
  csrc/loader_tail.cuh), bit-identical to the roll / slice / pad chain.
* ``normalize_by_max``        ``x / np.max(np.abs(x))`` (rec_transforms.py:47,65)
* ``TrainTransform`` / ``TestTransform``  rec_transforms.py:18-76 without the
  augmentations (``augmentation`` is None in the shipped configs,
  data/transform_wrappers.py:31-36): crop -> normalise -> Undersample -> batch dict
* ``output_transform`` / ``psnr``  rec_transforms.py:79-85 + metrics/image_metrics.py:7-19
  through the fused ``csmri_magnitude_clamp`` / ``csmri_psnr_sum`` kernels.
"""
import math

import torch

from . import _lib, ops, undersampling


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_images(img):
    if img.dim() != 3 or img.dtype != torch.float32 or not img.is_cuda:
        raise ValueError('images must be a float32 CUDA tensor of shape (B,H,W)')


def _shift_crop(x, out_ch, out_hw, in_roll, off, out_roll, absmax=None):
    """One gather pass through the composed roll / crop / pad / roll index map
    (csmri_shift_crop, csrc/loader_tail.cuh)."""
    if x.dim() == 3:
        B, in_ch, (ih, iw) = x.shape[0], 1, x.shape[1:]
    else:
        B, in_ch, ih, iw = x.shape
    oh, ow = out_hw
    with torch.cuda.device(x.device):
        out = torch.empty((B, out_ch, oh, ow), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csmri_shift_crop(
            x.data_ptr(), out.data_ptr(), B, in_ch, ih, iw, out_ch, oh, ow,
            in_roll[0] % ih, in_roll[1] % iw, off[0], off[1], out_roll[0] % oh, out_roll[1] % ow,
            absmax.data_ptr() if absmax is not None else None, _stream()))
    return out


def _center_crop(images, size, want_max):
    _check_images(images)
    if isinstance(size, (tuple, list)):
        sx, sy = int(size[0]), int(size[1])
    else:
        sx = sy = int(size)
    images = images.contiguous()
    B, nx, ny = images.shape
    # fft2c = fftshift(fft2(ifftshift(x)))   (deep_med_lib/utils/mymath.py:18-29)
    # ifftshift = roll by -(n // 2): x'[i] = x[(i + n // 2) mod n]; the real image gets its zero
    # imaginary plane in the same pass
    x = _shift_crop(images, 2, (nx, ny), (nx // 2, ny // 2), (0, 0), (0, 0))
    k = ops.fft2_planar(x)
    # fftshift (roll by +n // 2), crop_image_at(im_k, nx//2, ny//2, sx, sy) = box
    # [c - s//2, c + s//2) zero padded, then the ifftshift of ifft2c on the cropped size
    r1, r2 = sx // 2, sy // 2
    cnx, cny = 2 * r1, 2 * r2
    if cnx < 1 or cny < 1:
        raise ValueError('crop size %r is too small' % (size,))
    crop = _shift_crop(k, 2, (cnx, cny), (-(nx // 2), -(ny // 2)), (nx // 2 - r1, ny // 2 - r2),
                       (cnx // 2, cny // 2))
    y = ops.fft2_planar(crop, inverse=True)
    # the fftshift of ifft2c and np.abs, with max |.| per image as a by-product
    amax = torch.empty((B,), dtype=torch.float32, device=images.device) if want_max else None
    out = _shift_crop(y, 1, (cnx, cny), (-(cnx // 2), -(cny // 2)), (0, 0), (0, 0), absmax=amax)
    return out[:, 0], amax


def center_crop_in_kspace(images, size):
    """(B,H,W) real -> (B,size,size): magnitude of the image whose centred
    k-space has been cropped (or zero-padded) to size x size
    (myImageTransformations.CenterCropInKspace :935-954).  Two library FFTs and
    three gather passes (csmri_shift_crop); no torch kernels."""
    return _center_crop(images, size, False)[0]


def _divide(images, denom):
    B, h, w = images.shape
    with torch.cuda.device(images.device):
        out = torch.empty_like(images)
        _lib.check(_lib.lib().csmri_plane_divide(images.data_ptr(), denom.data_ptr(),
                                                 out.data_ptr(), B, h * w, _stream()))
    return out


def normalize_by_max(images):
    """x / max|x| per image (rec_transforms.py:47), bit-identical to the torch
    expression ``x / x.abs().amax((1, 2), keepdim=True)``."""
    _check_images(images)
    images = images.contiguous()
    B, h, w = images.shape
    with torch.cuda.device(images.device):
        amax = torch.empty((B,), dtype=torch.float32, device=images.device)
        _lib.check(_lib.lib().csmri_plane_absmax(images.data_ptr(), amax.data_ptr(), B, h * w,
                                                 _stream()))
    return _divide(images, amax)


def crop_and_normalize(images, size):
    """``normalize_by_max(center_crop_in_kspace(images, size))`` with the maximum
    taken inside the last gather pass (train_transform :40-47)."""
    x, amax = _center_crop(images, size, True)
    return _divide(x.contiguous(), amax)


def magnitude(x, lo=0.0, hi=float('inf')):
    """complex_abs (utils/tensor_transforms.py:62-75) fused with a clamp:
    (B,2,H,W) -> (B,1,H,W)."""
    ops._check_planar('x', x)
    x = x.contiguous()
    B, _, H, W = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csmri_magnitude_clamp(x.data_ptr(), out.data_ptr(), B, H, W,
                                                    float(lo), float(hi), _stream()))
    return out


def output_transform():
    """rec_transforms.py:79-85: (pred, target) -> clamp(|.|, 0, 1) of both."""
    def transform(pred, target):
        return magnitude(pred, 0.0, 1.0), magnitude(target, 0.0, 1.0)
    return transform


def psnr(pred, target):
    """compute_psnr (metrics/image_metrics.py:7-19) of output_transform(pred,
    target) in one fused pass; returns a Python float (one 8-byte read)."""
    ops._check_planar('pred', pred)
    ops._check_planar('target', target, pred)
    pred, target = pred.contiguous(), target.contiguous()
    B, _, H, W = pred.shape
    with torch.cuda.device(pred.device):
        acc = torch.zeros((1,), dtype=torch.float64, device=pred.device)
        _lib.check(_lib.lib().csmri_psnr_sum(pred.data_ptr(), target.data_ptr(), acc.data_ptr(),
                                             B, H, W, 0.0, 1.0, _stream()))
    mse = float(acc.item()) / (B * H * W)
    return 10.0 * math.log10(1.0 / mse)


class _Transform(object):
    def __init__(self, cs_params, image_size, downscale, fixed_mask, num_images):
        self.size = image_size // downscale
        self.undersample = undersampling.Undersample(
            cs_params['sampling_scheme'], (1, self.size, self.size),
            cs_params['acceleration_factor'],
            variable=(False if fixed_mask else cs_params.get('variable_acceleration', False)),
            fixed_mask=fixed_mask, num_fixed_masks=num_images)

    def __call__(self, images):
        return self.undersample(crop_and_normalize(images, self.size))


class TrainTransform(_Transform):
    """rec_transforms.train_transform (:18-57), batch/GPU, no augmentation."""

    def __init__(self, cs_params, image_size, downscale=1):
        super(TrainTransform, self).__init__(cs_params, image_size, downscale, False, 1)


class TestTransform(_Transform):
    """rec_transforms.test_transform (:60-76): fixed masks from RandomState(0)."""

    def __init__(self, cs_params, image_size, downscale=1, num_images=1):
        super(TestTransform, self).__init__(cs_params, image_size, downscale, True, num_images)
