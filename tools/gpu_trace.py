"""Per-CTA timeline of the two-column persistent strip kernel (tuning probe)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, undersampling  # noqa: E402

lib = _lib.lib()
dev = torch.device('cuda:0')
stream = torch.cuda.current_stream().cuda_stream
n, B = 256, int(os.environ.get('TRACE_B', '296'))
variant = int(os.environ.get('TRACE_VARIANT', '31'))
img = torch.rand(B, n, n, device=dev)
rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
batch = undersampling.undersample(img, rows)
plan = myfft.get_plan(batch['kspace'], batch['mask'])
xs = [torch.randn(B, 2, n, n, device=dev) for _ in range(2)]
out = torch.empty_like(xs[0])
lib.csmri_set_tuning(0, variant)
trace = torch.zeros(2 * 1024 * 40, dtype=torch.int64, device=dev)
for adj in (0, 1):
    for i in range(6):
        if i == 4:
            trace.zero_()
            lib.csmri_set_trace(trace.data_ptr())
        if adj:
            lib.csmri_dc_adjoint_cartesian(xs[i % 2].data_ptr(), plan.dtab.data_ptr(), out.data_ptr(),
                                           B, n, n, stream)
        else:
            lib.csmri_dc_forward_cartesian(xs[i % 2].data_ptr(), None, plan.dtab.data_ptr(),
                                           plan.addend.data_ptr(), out.data_ptr(), B, n, n, stream)
    torch.cuda.synchronize()
    lib.csmri_set_trace(None)
    tt = trace.cpu().numpy().reshape(2, 1024, 40)
    prev, t = tt[0], tt[1]
    prev = prev[prev[:, 0] > 0]
    t = t[t[:, 0] > 0]
    pn = (prev[:, 1:39] > 0).sum(1)
    prev_end = max(row[k] for row, k in zip(prev, pn))
    print('  gap: prev kernel last tile end -> this kernel first entry %.2f us; entry spread %.2f us; '
          'entry->prologue done (median) %.2f us' % (
              (t[:, 39].min() - prev_end) / 1e3, (t[:, 39].max() - t[:, 39].min()) / 1e3,
              np.median(t[:, 0] - t[:, 39]) / 1e3))
    t0 = t[:, 0].min()
    ntile = (t[:, 1:39] > 0).sum(1)
    ends = np.array([row[k] for row, k in zip(t, ntile)])
    print('adj' if adj else 'fwd', 'CTAs', len(t), 'tiles/CTA min/max', ntile.min(), ntile.max())
    print('  start spread us %.2f   first-tile-done us: min %.2f med %.2f max %.2f' % (
        (t[:, 0].max() - t0) / 1e3, (t[:, 1].min() - t0) / 1e3, (np.median(t[:, 1]) - t0) / 1e3,
        (t[:, 1].max() - t0) / 1e3))
    print('  end us: min %.2f med %.2f max %.2f' % ((ends.min() - t0) / 1e3,
                                                     (np.median(ends) - t0) / 1e3, (ends.max() - t0) / 1e3))
    d = np.diff(t[:, :ntile.min() + 1], axis=1) / 1e3
    print('  per-tile us (median over CTAs):', ' '.join('%.2f' % v for v in np.median(d, 0)))
    print('  per-tile us (max over CTAs):   ', ' '.join('%.2f' % v for v in d.max(0)))
