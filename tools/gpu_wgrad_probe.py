"""Time csmri_conv3x3_wgrad at the RecNet D5C5 shapes against torch's own backward-weight."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, conv  # noqa: E402

_lib.lib().csmri_set_tuning(5, int(os.environ.get('WGRAD_COT', '8')))

torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
N, H, W = 32, 256, 256
reps = int(os.environ.get('WGRAD_REPS', '10'))
for ci, co in ((32, 32), (2, 32), (32, 2), (64, 64)):
    x = torch.randn(N, ci, H, W, device='cuda')
    gy = torch.randn(N, co, H, W, device='cuda')
    w = torch.randn(co, ci, 3, 3, device='cuda')

    def ours():
        return conv.conv3x3_wgrad(x, gy, 1)

    def theirs():
        return torch.ops.aten.convolution_backward(gy, x, w, [co], [1, 1], [1, 1], [1, 1], False,
                                                   [0, 0], 1, [False, True, False])[1]
    res = {}
    for name, fn in (('ours', ours), ('cudnn', theirs)):
        if name == 'cudnn' and os.environ.get('WGRAD_ONLY_OURS', '0') == '1':
            continue
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = (a.elapsed_time(b) / reps, out)
    flop = 2.0 * 9 * ci * co * N * H * W
    line = '%2d->%2d ' % (ci, co)
    for name, (ms, _) in res.items():
        line += ' %s %.3f ms (%.1f TFLOP/s)' % (name, ms, flop / ms / 1e9)
    if len(res) == 2:
        line += '  rel diff %.2e' % ((res['ours'][1] - res['cudnn'][1]).norm() /
                                     res['cudnn'][1].norm()).item()
    print(line, flush=True)
