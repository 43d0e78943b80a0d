"""The three tensor-core kernels of a 32 -> 32 RecNet layer at the D5C5 training shape
(batch 32, 256 x 256), a few launches each - the target of `ncu -k regex:conv3x3.*tc`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from csmri_refinement_b200 import conv  # noqa: E402

dev = torch.device('cuda:0')
n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 256, 256)))
x = torch.randn(n, 32, h, w, device=dev)
gy = torch.randn(n, 32, h, w, device=dev) * 0.05
wt = torch.randn(32, 32, 3, 3, device=dev) * 0.08
b = torch.randn(32, device=dev)
for _ in range(3):
    y = conv.conv3x3_tc(x, wt, b, 0.01)
    gx = conv.conv3x3_tc(gy, wt, None, 0.0, transpose_flip=True)
    dw = conv.conv3x3_wgrad(x, gy, 1)
torch.cuda.synchronize()
act, sg, _ = conv.conv3x3_tc_signs(x, wt, b, 0.01)
for _ in range(3):
    gm = conv.conv3x3_tc_masked(gy, wt, sg, 0.01)
    dwb = conv.conv3x3_wgrad_bias(x, gy)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
for _ in range(10):
    y = conv.conv3x3_tc(x, wt, b, 0.01)
ev[1].record()
for _ in range(10):
    gx = conv.conv3x3_tc(gy, wt, None, 0.0, transpose_flip=True)
ev[2].record()
for _ in range(10):
    dw = conv.conv3x3_wgrad(x, gy, 1)
ev[3].record()
for _ in range(10):
    gm = conv.conv3x3_tc_masked(gy, wt, sg, 0.01)
ev[4].record()
for _ in range(10):
    dwb = conv.conv3x3_wgrad_bias(x, gy)
ev[5].record()
torch.cuda.synchronize()
flop = 2.0 * 9 * 32 * 32 * n * h * w
byt = 2.0 * x.numel() * 4
for name, a, c in (('forward (bias + LeakyReLU fused)', 0, 1), ('data gradient', 1, 2), ('weight gradient', 2, 3),
                   ('data gradient x LeakyReLU derivative', 3, 4), ('weight + bias gradient', 4, 5)):
    ms = ev[a].elapsed_time(ev[c]) / 10
    print('%-38s %.3f ms  %.0f TFLOP/s fp32-equivalent  %.0f GB/s of operand traffic' % (
        name, ms, flop / ms / 1e9, byt / ms / 1e6))

# the thin 32 -> 2 layer (forward of a block's last layer / data gradient of its first), TMA vs cp.async staging
from csmri_refinement_b200 import _lib  # noqa: E402
w2 = torch.randn(2, 32, 3, 3, device=dev) * 0.1
b2 = torch.randn(2, device=dev)
for key, name in ((1, 'thin 32 -> 2, TMA-staged'), (0, 'thin 32 -> 2, cp.async-staged')):
    _lib.lib().csmri_set_tuning(8, key)
    for _ in range(3):
        t2 = conv.conv3x3_thin(x, w2, b2, 0.0)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        t2 = conv.conv3x3_thin(x, w2, b2, 0.0)
    c.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(c) / 10
    print('%-38s %.3f ms  %.0f GB/s of operand traffic' % (name, ms, (x.numel() + t2.numel()) * 4 / ms / 1e6))
_lib.lib().csmri_set_tuning(8, 1)

# the thin 2 -> 32 layer (forward of a block's first layer / data gradient of its last), plain and with the sign-word mask
w3 = torch.randn(32, 2, 3, 3, device=dev) * 0.1
x2 = torch.randn(n, 2, h, w, device=dev)
for name, fn in (('thin 2 -> 32', lambda: conv.conv3x3_thin(x2, w3, None, 0.0)),
                 ('thin 2 -> 32 x LeakyReLU derivative', lambda: conv.conv3x3_thin_masked(x2, w3, sg, 0.01))):
    for _ in range(3):
        fn()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        t3 = fn()
    c.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(c) / 10
    print('%-38s %.3f ms  %.0f GB/s of operand traffic' % (name, ms, (x2.numel() + t3.numel()) * 4 / ms / 1e6))
