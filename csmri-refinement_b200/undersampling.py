"""Undersampling path: host-side line selection + GPU mask application / k0.

Mirrors, for the Cartesian scheme every shipped config uses
(myImageTransformations.py:71-81: "varden" falls through to the Cartesian
branch):

* ``cs.cartesian_mask``  data/reconstruction/deep_med_lib/utils/compressed_sensing.py:82-123
  - stays on the HOST: it draws from numpy's legacy ``RandomState``
  (``rng.choice(..., replace=False, p=pdf)``), which cannot be reproduced
  bit-exactly on a GPU.  Only the sampled *line indices* are produced here
  (N bytes per slice instead of an N x N float64 mask).
* ``cs.undersample``     compressed_sensing.py:460-512 (centred=False, ortho)
  + ``dnn_io.to_tensor_format`` (dnn_io.py:47-61) + the channel split of
  scar_segmentation.py:212-218 - run on the GPU (``csmri_undersample``).
* ``Undersample``        myImageTransformations.py:1196-1238 (the transform).
"""
import numpy as np
import torch

from . import ops


def _normal_pdf(length, sensitivity):
    return np.exp(-sensitivity * (np.arange(length) - length / 2) ** 2)


def cartesian_rows(shape, acc, sample_n=10, centred=False, rng=None):
    """Sampled phase-encode lines of ``cs.cartesian_mask(shape, acc, sample_n,
    centred, rng)`` as a uint8 array (N, Nx); ``mask[n, x, y] == rows[n, x]``.

    Consumes ``rng`` exactly like the reference (one ``choice`` per image), so
    the selected indices are bit-identical for the same ``RandomState``.
    """
    if rng is None:
        rng = np.random
    n_img, nx = int(np.prod(shape[:-2])), shape[-2]
    pdf_x = _normal_pdf(nx, 0.5 / (nx / 10.) ** 2)
    lmda = nx / (2. * acc)
    n_lines = nx // acc
    pdf_x += lmda * 1. / nx
    lo, hi = nx // 2 - sample_n // 2, nx // 2 + sample_n // 2
    if sample_n:
        pdf_x[lo:hi] = 0
        pdf_x /= np.sum(pdf_x)
        n_lines -= sample_n
    rows = np.zeros((n_img, nx), dtype=np.uint8)
    for i in range(n_img):
        idx = rng.choice(nx, int(n_lines), False, pdf_x)
        rows[i, idx] = 1
    if sample_n:
        rows[:, lo:hi] = 1
    if not centred:
        rows = np.fft.ifftshift(rows, axes=-1)
    return np.ascontiguousarray(rows)


def cartesian_mask(shape, acc, sample_n=10, centred=False, rng=None):
    """Dense float64 mask with the reference's shape and values (for callers
    that want the array ``cs.cartesian_mask`` returns)."""
    rows = cartesian_rows(shape, acc, sample_n, centred, rng)
    ny = shape[-1]
    mask = np.repeat(rows[:, :, None], ny, axis=2).astype(np.float64)
    return mask.reshape(shape)


def consume_noise_draws(rng, shape):
    """``cs.undersample`` draws two normal arrays even for noise == 0
    (compressed_sensing.py:494-495).  Call this to keep a shared RNG stream in
    step with the reference when masks and (zero) noise use the same rng
    (myImageTransformations.py:1199-1207,1227)."""
    rng.normal(0, 1, shape)
    rng.normal(0, 1, shape)


def undersample(img, rows, register_dc_plan=True):
    """GPU ``cs.undersample`` + tensor formatting.  ``img`` (B,H,W) float32
    CUDA; ``rows`` (B,H) uint8 (numpy or tensor).  Returns the batch dict
    ``{inp, kspace, mask, target}`` of scar_segmentation.py:212-218.

    With ``register_dc_plan`` the per-batch constants of the (noiseless) DC
    layers - which fall out of the same kernels - are registered with
    :mod:`myfft`, so ``RecNet.forward`` on this batch runs without a prepare
    pass and without any host synchronisation."""
    if isinstance(rows, np.ndarray):
        rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.uint8))
    rows = rows.to(device=img.device, dtype=torch.uint8, non_blocking=True)
    if not register_dc_plan:
        return ops.undersample(img, rows)
    from . import myfft
    batch, (dtab, addend) = ops.undersample(img, rows, with_plan=True)
    myfft.register_plan(batch['kspace'], batch['mask'], dtab, addend)
    return batch


def compact_lines(kspace, rows):
    """Sampled lines of a Cartesian k-space batch, ``kspace[b][:, rows[b] != 0, :]``
    stacked to (B,2,L,W) in ascending row order - the compact form
    ``myfft.plan_from_lines`` / ``HostDCPipeline.forward_backward_lines`` take
    (k0 is zero everywhere else, compressed_sensing.py:510).  Works on host and
    device tensors; every slice must have the same number of sampled rows
    (``cs.cartesian_mask`` always samples ``Nx // acc`` lines)."""
    if isinstance(rows, np.ndarray):
        rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.uint8))
    rows = rows.to(kspace.device)
    counts = (rows != 0).sum(dim=1)
    n_lines = int(counts[0].item())
    if n_lines == 0 or not bool((counts == n_lines).all().item()):
        raise ValueError('every slice needs the same, non-zero number of sampled rows')
    b, _, h, w = kspace.shape
    idx = (rows != 0).nonzero()[:, 1].reshape(b, n_lines)          # ascending per slice
    gather = idx[:, None, :, None].expand(b, 2, n_lines, w)
    return torch.gather(kspace, 2, gather).contiguous()


class Undersample(object):
    """Batch/GPU version of myImageTransformations.Undersample (:1196-1238).

    ``__call__(images)`` takes a float32 CUDA tensor (B,H,W) (already scaled
    to [0,1], rec_transforms.py:47) and returns the batch dict.  Mask lines
    are drawn on the host in the reference's order: with ``fixed_mask`` from
    ``RandomState(0)`` at construction and then cycled, else from the global
    ``np.random`` per image.
    """

    def __init__(self, mask_type, im_shape, acceleration_rate=4, variable=False,
                 fixed_mask=False, num_fixed_masks=1, keep_rng_in_step=True):
        if mask_type == 'radial':
            raise NotImplementedError('radial sampling is not reachable from the shipped '
                                      'configs and is not implemented')
        self.im_shape = tuple(im_shape)
        self.acc = acceleration_rate
        self.variable = variable
        self.keep_rng_in_step = keep_rng_in_step
        if fixed_mask:
            self.rng = np.random.RandomState(seed=0)
            self.current_mask = 0
            self.fixed_rows = [self._draw() for _ in range(num_fixed_masks)]
        else:
            self.rng = np.random
            self.fixed_rows = None

    def _draw(self):
        central_lines = 8
        n = self.im_shape[0]
        if self.variable:
            rows = np.zeros((n, self.im_shape[-2]), dtype=np.uint8)
            for i in range(n):
                acc_r = float(self.rng.uniform(1, self.acc * 1.5))
                rows[i] = cartesian_rows(self.im_shape[1:], acc_r, central_lines,
                                         centred=False, rng=self.rng)[0]
            return rows
        return cartesian_rows(self.im_shape, self.acc, central_lines, centred=False,
                              rng=self.rng)

    def next_rows(self, batch):
        """Rows for ``batch`` images, drawn in the reference's per-sample order."""
        out = []
        per = self.im_shape[0]
        while sum(r.shape[0] for r in out) < batch:
            if self.fixed_rows is None:
                rows = self._draw()
            else:
                rows = self.fixed_rows[self.current_mask]
                self.current_mask = (self.current_mask + 1) % len(self.fixed_rows)
            out.append(rows)
            if self.keep_rng_in_step:
                consume_noise_draws(self.rng, (per,) + self.im_shape[1:])
        return np.concatenate(out, 0)[:batch]

    def __call__(self, images):
        rows = self.next_rows(images.shape[0])
        return undersample(images, rows)
