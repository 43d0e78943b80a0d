"""Drop-in for data/reconstruction/deep_med_lib/my_pytorch/myfft.py.

Public names of the reference module that sit on the DC path keep their name and
argument meaning: ``make_contiguous``, ``contiguous_clone`` (:10-18), ``Fft2d`` /
``Ifft2d`` (:78-128, two-tensor real / imaginary calling convention, autograd
included), ``data_consistency(k, k0, mask, noise_lvl)`` (:131-142, the k-space
blend alone) and ``DataConsistencyInKspace`` (:145-163).  The 1-D ``Fft`` / ``Ifft``
(:21-75) are not on the path and are not provided.

``DataConsistencyInKspace(noise_lvl=None, norm='ortho').perform(x, k0, mask)``
keeps the reference's name, arguments and semantics; the work is done by the
sm_100a kernels of ``csrc/csmri_dc.cu`` through the C ABI.  Like the reference
object it is not an ``nn.Module`` and owns no parameters or buffers, so
``RecNet.state_dict()`` keeps its keys (models/recnet.py:128-134: ``dc_layers``
is a plain list).

``k0`` and ``mask`` are the same tensors for every DC layer of a cascade
(models/recnet.py:139-151), so what depends only on them - the proof that the
mask is row-constant (compressed_sensing.py:115-116), the per-row diagonal and
the k0 term in hybrid space ``iFFT_W(k0)`` - is computed once per batch and
cached here.
"""
import collections
import contextlib
import threading

import torch

from . import ops


class DCPlan(object):
    """Per-batch constants of the DC operator (see csmri_dc_prepare).

    ``flag`` is the device int32 the prepare pass wrote (1 = every mask row is
    constant / the line table is consistent); it is only *read* here when the
    caller did not state ``assume_row_constant``.
    """

    __slots__ = ('k0', 'mask', 'noise_lvl', 'dtab', 'addend', 'row_constant', 'key', 'flag')

    def __init__(self, k0, mask, noise_lvl, assume_row_constant=None, prepared=None):
        v = float(noise_lvl) if noise_lvl else 0.0   # `if v:` of myfft.py:137
        self.k0, self.mask, self.noise_lvl = k0, mask, v
        self.flag = None
        self.key = None
        if prepared is not None:
            # handed over by the loader (undersampling.undersample) or built from
            # a line table (plan_from_lines): row-constant by construction and
            # (dtab, addend) already exist
            self.dtab, self.addend = prepared
            self.row_constant = True
            return
        if assume_row_constant is not None and not assume_row_constant:
            # general path stated by the caller: the k0 row transform is not needed
            dtab, _, flag = ops.dc_prepare(k0.detach(), mask.detach(), v, False)
            self.dtab, self.addend, self.flag, self.row_constant = dtab, None, flag, False
            return
        dtab, addend, flag = ops.dc_prepare(k0.detach(), mask.detach(), v, True)
        self.dtab, self.addend, self.flag = dtab, addend, flag
        if assume_row_constant is None:
            # one 4-byte device->host read per batch
            self.row_constant = bool(flag.item())
            if not self.row_constant:
                self.addend = None
        else:
            self.row_constant = True
            rec = _ASSUMPTION_RECORDER
            if rec is not None:
                rec.flags.append(flag)     # checked later by whoever made the assumption


def _tensor_key(t):
    return (t.data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version,
            t.device.index)


_PLAN_CACHE = collections.OrderedDict()
_PLAN_CACHE_SIZE = 4     # keys; an entry pins k0, mask and the prepared k0 term of one batch
_PLAN_LOCK = threading.RLock()   # DataParallel-style callers reach the cache from several threads
_ASSUME_ROW_CONSTANT = None     # process-wide default, see assume_row_constant()
_ASSUMPTION_RECORDER = None


class _Assumption(object):
    """What ``assume_row_constant`` yields: the device flags of every plan that
    was built on the assumption inside the block.  ``violations()`` returns a
    0-dim int64 tensor (number of batches whose mask was NOT row-constant)
    without synchronising, so it can be evaluated inside a CUDA graph."""

    def __init__(self, value):
        self.value = value
        self.flags = []

    def violations(self):
        if not self.flags:
            return None
        seen, uniq = set(), []
        for f in self.flags:            # every DC layer of a cascade reports the same plan
            if id(f) not in seen:
                seen.add(id(f))
                uniq.append(f.reshape(()))
        return (1 - torch.stack(uniq)).sum()


@contextlib.contextmanager
def assume_row_constant(value):
    """Within the block, skip the per-batch device->host read that proves the
    mask is row-constant and take ``value`` (True: Cartesian strip kernel,
    False: general path, None: check) instead.  Needed under CUDA-graph
    capture, where a host read is illegal.  With ``True`` the proof is still
    computed on the device; the yielded object collects those flags so the
    caller can verify the assumption after the fact (ShardedTrainer does)."""
    global _ASSUME_ROW_CONSTANT, _ASSUMPTION_RECORDER
    prev, prev_rec = _ASSUME_ROW_CONSTANT, _ASSUMPTION_RECORDER
    rec = _Assumption(value)
    _ASSUME_ROW_CONSTANT = value
    _ASSUMPTION_RECORDER = rec if value else None
    try:
        yield rec
    finally:
        _ASSUME_ROW_CONSTANT, _ASSUMPTION_RECORDER = prev, prev_rec


def get_plan(k0, mask, noise_lvl=None, assume_row_constant=None):
    """Cached :class:`DCPlan` for this (k0, mask, noise_lvl).

    The cache keeps the two tensors alive, so their storage cannot be handed
    to another batch while an entry exists; in-place writes through torch bump
    ``_version`` and miss.  Writes torch cannot see (a CUDA-graph replay filling
    static input buffers, ``.data`` writes, raw C-ABI writes) do NOT miss: such
    callers must bypass the cache - pass ``plan=`` to :func:`dc_perform` or call
    :func:`clear_plan_cache` (ShardedTrainer rebuilds the plan inside its graph).
    Entries are evicted LRU (4 keys = 2-4 batches).
    """
    v = float(noise_lvl) if noise_lvl else 0.0
    if assume_row_constant is None:
        assume_row_constant = _ASSUME_ROW_CONSTANT
    key = (_tensor_key(k0), _tensor_key(mask), v, assume_row_constant)
    with _PLAN_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
            if plan.flag is not None and assume_row_constant and _ASSUMPTION_RECORDER is not None:
                _ASSUMPTION_RECORDER.flags.append(plan.flag)
            return plan
    plan = DCPlan(k0, mask, v, assume_row_constant)
    plan.key = key
    with _PLAN_LOCK:
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _PLAN_CACHE.popitem(last=False)
    return plan


def register_plan(k0, mask, dtab, addend):
    """Install a noiseless plan the loader produced as a by-product
    (``ops.undersample(..., with_plan=True)``): the first DC layer then needs no
    prepare pass and no device->host read for this batch."""
    plan = DCPlan(k0, mask, 0.0, prepared=(dtab, addend))
    with _PLAN_LOCK:
        for arc in (None, True):
            key = (_tensor_key(k0), _tensor_key(mask), 0.0, arc)
            _PLAN_CACHE[key] = plan
            _PLAN_CACHE.move_to_end(key)
        plan.key = key
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _PLAN_CACHE.popitem(last=False)
    return plan


def plan_from_lines(k0_lines, rows, noise_lvl=None, check=True):
    """:class:`DCPlan` from the compact description of a Cartesian acquisition:
    ``rows`` (B,H) uint8 sampled-line table and ``k0_lines`` (B,2,L,W) =
    ``kspace[b][:, rows[b] != 0, :]`` (k0 is zero off the sampled lines by
    construction, compressed_sensing.py:510).  Nothing dense has to exist on
    the host or cross PCIe.  ``check`` reads the 4-byte consistency flag (every
    slice has exactly L sampled rows); pass False under graph capture / in a
    pipeline and look at ``plan.flag`` later.  Not cached: the caller owns it."""
    v = float(noise_lvl) if noise_lvl else 0.0
    dtab, addend, ok = ops.dc_prepare_lines(k0_lines, rows, v, k0_lines.shape[-1])
    if check and not bool(ok.item()):
        raise ValueError('rows / k0_lines are inconsistent: every slice must have exactly '
                         '%d sampled rows' % k0_lines.shape[2])
    plan = DCPlan(None, None, v, prepared=(dtab, addend))
    plan.flag = ok
    return plan


def clear_plan_cache():
    with _PLAN_LOCK:
        _PLAN_CACHE.clear()


def make_contiguous(*Xs):
    """myfft.py:10-11."""
    return tuple(X if X.is_contiguous() else X.contiguous() for X in Xs)


def contiguous_clone(X):
    """myfft.py:14-18."""
    return X.clone() if X.is_contiguous() else X.contiguous()


class _Fft2Planar(torch.autograd.Function):
    """scale * (i)FFT2_ortho of a planar (B,2,H,W) tensor.  The adjoint of a unitary
    transform is its inverse, which is what the reference's swapped-plane
    backward (myfft.py:92-102, 119-128) evaluates."""

    @staticmethod
    def forward(ctx, x, inverse, scale):
        ctx.inverse, ctx.scale = inverse, scale
        out = ops.fft2_planar(x, inverse=inverse)
        return out if scale == 1.0 else out * scale

    @staticmethod
    def backward(ctx, grad):
        g = ops.fft2_planar(grad.contiguous(), inverse=not ctx.inverse)
        return (g if ctx.scale == 1.0 else g * ctx.scale), None, None


class _Fft2dBase(object):
    _inverse = False

    def __init__(self, norm=None):
        if norm not in (None, 'ortho'):
            raise ValueError("norm must be None or 'ortho', got %r" % (norm,))
        self.norm = norm

    def __call__(self, X_re, X_im):
        if X_re.shape != X_im.shape or X_re.dim() < 2:
            raise ValueError('real and imaginary parts must have the same (..., H, W) shape')
        h, w = X_re.shape[-2], X_re.shape[-1]
        lead = X_re.shape[:-2]
        x = torch.stack((X_re.reshape(-1, h, w), X_im.reshape(-1, h, w)), dim=1)
        # pytorch_fft: fft2 un-normalised, ifft2 divides by H*W (myfft.py:85,112)
        n = float(h * w) ** 0.5
        scale = 1.0 if self.norm == 'ortho' else (1.0 / n if self._inverse else n)
        out = _Fft2Planar.apply(x, self._inverse, scale)
        return out[:, 0].reshape(*lead, h, w), out[:, 1].reshape(*lead, h, w)

    forward = __call__


class Fft2d(_Fft2dBase):
    """myfft.py:78-102: ``k_re, k_im = Fft2d(norm)(X_re, X_im)`` over the last two axes."""
    _inverse = False


class Ifft2d(_Fft2dBase):
    """myfft.py:105-128: ``x_re, x_im = Ifft2d(norm)(k_re, k_im)``."""
    _inverse = True


def data_consistency(k, k0, mask, noise_lvl=None):
    """myfft.py:131-142, the blend alone, for callers that hold k-space already
    (same tensor expression as the reference, hence bit-identical; inside
    :meth:`DataConsistencyInKspace.perform` it is fused into the strip kernel).
    k    - input in k-space
    k0   - initially sampled elements in k-space
    mask - corresponding nonzero location
    """
    v = noise_lvl
    if v:  # noisy case
        out = (1 - mask) * k + mask * (k + v * k0) / (1 + v)
    else:  # noiseless case
        out = (1 - mask) * k + k0
    return out


def dc_perform(x, k0, mask, noise_lvl=None, residual=None, plan=None):
    """iFFT2o(blend(FFT2o(x [+ residual]), k0, mask)) - myfft.py:153-163 in one op."""
    if not x.is_cuda:
        raise RuntimeError('DataConsistencyInKspace needs CUDA tensors: the B200 DC path has '
                           'no CPU fallback (use oracle/ for CPU checks)')
    if plan is None:
        plan = get_plan(k0, mask, noise_lvl)
    if plan.row_constant:
        return ops.dc_cartesian(x, residual, plan.dtab, plan.addend)
    return ops.dc_general(x, residual, plan.k0, plan.mask, plan.noise_lvl)


class DataConsistencyInKspace(object):
    """Create data consistency operator (interface of myfft.py:145-163)."""

    def __init__(self, noise_lvl=None, norm='ortho'):
        if norm != 'ortho':
            # models/recnet.py:131 only ever builds norm='ortho'
            raise ValueError("only norm='ortho' is supported, got %r" % (norm,))
        self.noise_lvl = noise_lvl
        self.norm = norm

    def perform(self, x, k0, mask, residual=None):
        """
        x    - input in image space, (B,2,H,W)
        k0   - initially sampled elements in k-space
        mask - corresponding nonzero location
        residual - optional extension: added to x before the transform
                   (the `x + block_input` of models/recnet.py:147-148)
        """
        return dc_perform(x, k0, mask, self.noise_lvl, residual)

    def perform_lines(self, x, k0_lines, rows, residual=None, plan=None):
        """Same operator for a Cartesian acquisition given compactly (extension,
        see :func:`plan_from_lines`): ``rows`` (B,H) uint8 line table and
        ``k0_lines`` (B,2,L,W), the sampled lines of k0 only."""
        if plan is None:
            plan = plan_from_lines(k0_lines, rows, self.noise_lvl)
        return dc_perform(x, None, None, self.noise_lvl, residual, plan=plan)
