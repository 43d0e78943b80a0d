"""HostDCPipeline chunk-size sweep (context probe for the e2e number)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import hostpipe, undersampling  # noqa: E402

dev = torch.device('cuda:0')
B, n = 256, 256
img = torch.rand(B, n, n, device=dev)
rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
batch = undersampling.undersample(img, rows)
hx = torch.randn(B, 2, n, n).pin_memory()
hg = torch.randn(B, 2, n, n).pin_memory()
hk0, hm = batch['kspace'].cpu().pin_memory(), batch['mask'].cpu().pin_memory()
ho, hgx = torch.empty_like(hx).pin_memory(), torch.empty_like(hx).pin_memory()
# raw PCIe reference points
d = torch.empty_like(hx, device=dev)
for name, fn in (('H2D 128MiB', lambda: d.copy_(hx, non_blocking=True)),
                 ('D2H 128MiB', lambda: ho.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record(); torch.cuda.synchronize()
    print(name, '%.1f GB/s' % (5 * hx.numel() * 4 / a.elapsed_time(b) / 1e6))
for chunk, depth in ((16, 3), (32, 3), (64, 3), (128, 2), (64, 4)):
    pipe = hostpipe.HostDCPipeline(dev, chunk=chunk, depth=depth)
    for _ in range(2):
        pipe.forward_backward(hx, hk0, hm, hg, ho, hgx)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        pipe.forward_backward(hx, hk0, hm, hg, ho, hgx)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print('chunk %d depth %d: %.2f ms/step  %.0f slices/s  H2D %.1f GB/s' % (
        chunk, depth, ms, B / ms * 1e3, 4 * hx.numel() * 4 / ms / 1e6))
