"""HostDCPipeline sweep (context probe for the e2e number): chunk size, buffers in
flight, copy streams per direction, with and without CUDA-graph replay, for the
compact (line table + sampled k0 lines) and the dense input form."""
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
hrows = torch.from_numpy(rows).pin_memory()
hk0l = undersampling.compact_lines(batch['kspace'], hrows.to(dev)).cpu().pin_memory()
ho, hgx = torch.empty_like(hx).pin_memory(), torch.empty_like(hx).pin_memory()
prop = torch.cuda.get_device_properties(dev)
print('async engines:', getattr(prop, 'async_engine_count', '?'))
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
for chunk, depth, cs, graph in ((64, 3, 1, True), (32, 3, 1, True), (16, 4, 1, True), (64, 3, 2, True),
                                (32, 4, 2, True), (16, 4, 2, True), (32, 4, 2, False), (8, 6, 2, True)):
    pipe = hostpipe.HostDCPipeline(dev, chunk=chunk, depth=depth, copy_streams=cs, use_graph=graph)
    res = []
    for form in ('lines', 'dense'):
        if form == 'lines':
            call = lambda: pipe.forward_backward_lines(hx, hk0l, hrows, hg, ho, hgx)   # noqa: E731
        else:
            call = lambda: pipe.forward_backward(hx, hk0, hm, hg, ho, hgx)             # noqa: E731
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            call()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 5)
    print('chunk %3d depth %d copy_streams %d graph %d: lines %.2f ms (%.0f slices/s)  dense %.2f ms' % (
        chunk, depth, cs, graph, res[0], B / res[0] * 1e3, res[1]))
