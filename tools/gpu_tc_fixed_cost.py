"""Fixed cost per launch of the tensor-core kernels: one work item (N=1, 16x128) vs the training shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from csmri_refinement_b200 import conv  # noqa: E402

dev = torch.device('cuda:0')
for n, h, w in ((1, 16, 128), (1, 64, 128), (4, 128, 128), (32, 256, 256)):
    x = torch.randn(n, 32, h, w, device=dev)
    gy = torch.randn(n, 32, h, w, device=dev)
    wt = torch.randn(32, 32, 3, 3, device=dev) * 0.1
    b = torch.randn(32, device=dev)
    for name, fn in (('forward', lambda: conv.conv3x3_tc(x, wt, b, 0.01)),
                     ('weight gradient', lambda: conv.conv3x3_wgrad(x, gy, 1))):
        for _ in range(5):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        c.record()
        torch.cuda.synchronize()
        print('N=%2d %3dx%3d %-16s %.1f us per launch (20 launches in a CUDA graph)' % (n, h, w, name, a.elapsed_time(c) * 50))
