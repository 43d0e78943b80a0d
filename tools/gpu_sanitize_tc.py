"""Small workload of the round-2 kernels for compute-sanitizer racecheck / synccheck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from csmri_refinement_b200 import conv  # noqa: E402

dev = torch.device('cuda:0')
xt = torch.randn(1, 32, 16, 128, device=dev)
gy = torch.randn(1, 32, 16, 128, device=dev)
wt = torch.randn(32, 32, 3, 3, device=dev) * 0.1
y, sg, isg = conv.conv3x3_tc_signs(xt, wt, torch.randn(32, device=dev), 0.01, want_in_signs=True)
conv.conv3x3_tc_masked(gy, wt, sg, 0.01)
conv.conv3x3_wgrad_bias(xt, gy)
x2 = torch.randn(1, 2, 16, 128, device=dev)
conv.conv3x3_thin(xt, torch.randn(2, 32, 3, 3, device=dev), torch.randn(2, device=dev), 0.0)      # TMA-staged 32 -> 2
conv.conv3x3_wgrad(xt, torch.randn(1, 2, 16, 128, device=dev), 1)                                  # TMA-staged 32 -> 2 weight gradient
conv.conv3x3_thin_masked(x2, torch.randn(32, 2, 3, 3, device=dev), sg, 0.01)
conv.conv3x3_wgrad_thin_bias(x2, gy)
torch.cuda.synchronize()
print('ok round-2 kernels')
