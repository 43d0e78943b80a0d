"""GPU mirror of data/reconstruction/rec_transforms.py (the loader tail and the
reporting transform around the DC path).

* ``center_crop_in_kspace``  myImageTransformations.CenterCropInKspace (:935-954):
  fft2c -> crop at the centre (``crop_image_at`` :105-117) -> ifft2c -> abs.
  The FFTs are the library's ``csmri_fft2``; shifts / crop are index plumbing.
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


def center_crop_in_kspace(images, size):
    """(B,H,W) real -> (B,size,size): magnitude of the image whose centred
    k-space has been cropped (or zero-padded) to size x size."""
    _check_images(images)
    if isinstance(size, (tuple, list)):
        sx, sy = int(size[0]), int(size[1])
    else:
        sx = sy = int(size)
    B, nx, ny = images.shape
    x = torch.stack([images, torch.zeros_like(images)], dim=1)
    # fft2c = fftshift(fft2(ifftshift(x)))   (deep_med_lib/utils/mymath.py:18-29)
    x = torch.roll(x, shifts=(-(nx // 2), -(ny // 2)), dims=(2, 3))
    k = ops.fft2_planar(x.contiguous())
    k = torch.roll(k, shifts=(nx // 2, ny // 2), dims=(2, 3))
    # crop_image_at(im_k, nx//2, ny//2, sx, sy): box [c - s//2, c + s//2), zero padded
    cx, cy, r1, r2 = nx // 2, ny // 2, sx // 2, sy // 2
    x1, x2, y1, y2 = cx - r1, cx + r1, cy - r2, cy + r2
    crop = k[:, :, max(x1, 0):min(x2, nx), max(y1, 0):min(y2, ny)]
    pad = (max(0, -y1), max(0, y2 - ny), max(0, -x1), max(0, x2 - nx))
    if any(pad):
        crop = torch.nn.functional.pad(crop, pad)
    cnx, cny = crop.shape[2], crop.shape[3]
    crop = torch.roll(crop, shifts=(-(cnx // 2), -(cny // 2)), dims=(2, 3))
    y = ops.fft2_planar(crop.contiguous(), inverse=True)
    y = torch.roll(y, shifts=(cnx // 2, cny // 2), dims=(2, 3))
    return magnitude(y.contiguous())[:, 0]


def normalize_by_max(images):
    """x / max|x| per image (rec_transforms.py:47)."""
    _check_images(images)
    return images / images.abs().amax(dim=(1, 2), keepdim=True)


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
        x = center_crop_in_kspace(images, self.size)
        x = normalize_by_max(x)
        return self.undersample(x)


class TrainTransform(_Transform):
    """rec_transforms.train_transform (:18-57), batch/GPU, no augmentation."""

    def __init__(self, cs_params, image_size, downscale=1):
        super(TrainTransform, self).__init__(cs_params, image_size, downscale, False, 1)


class TestTransform(_Transform):
    """rec_transforms.test_transform (:60-76): fixed masks from RandomState(0)."""

    def __init__(self, cs_params, image_size, downscale=1, num_images=1):
        super(TestTransform, self).__init__(cs_params, image_size, downscale, True, num_images)
