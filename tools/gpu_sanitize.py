"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import (conv, myfft, ops, rec_transforms, recnet, refinement_ops,  # noqa: E402
                                   undersampling)

dev = torch.device('cuda:0')
for n, B in ((64, 5), (128, 3), (256, 3), (320, 2), (512, 2), (32, 4)):
    img = torch.rand(B, n, n, device=dev)
    rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
    batch = undersampling.undersample(img, rows)
    x = torch.randn(B, 2, n, n, device=dev, requires_grad=True)
    for noise in (None, 0.1):
        out = myfft.DataConsistencyInKspace(noise_lvl=noise).perform(x, batch['kspace'], batch['mask'])
        out.sum().backward()
    gm = (torch.rand(B, 1, n, n, device=dev) < 0.3).float().expand(B, 2, n, n).contiguous()
    o2 = ops.dc_general(x.detach(), None, batch['kspace'], gm, 0.1)
    ops.dc_general_adjoint(o2, gm, 0.1)
    rec_transforms.psnr(out.detach(), batch['target'])
    torch.cuda.synchronize()
    print('ok', n)

# refinement-path pointwise ops
t = torch.randn(3, 2, 40, 52, device=dev)
s_, mn, mx = refinement_ops.scale(t)
refinement_ops.unscale(s_, mn, mx)
refinement_ops.magnitude_image(t)
learn = torch.randn(3, 1, 40, 52, device=dev, requires_grad=True)
sc = torch.nn.Parameter(torch.full((1,), 0.3, device=dev))
refinement_ops.refinement_real_penalty_add(t, learn, sc)['pred'].sum().backward()
# training-step kernels: weight gradients (thick / thin), thin convolutions, epilogues,
# through one small RecNet step (image edges, first / last tiles included)
for ci, co, h, w, pad in ((32, 32, 8, 32, 1), (64, 32, 4, 64, 0), (2, 32, 16, 32, 1), (32, 2, 32, 64, 0)):
    conv.conv3x3_wgrad(torch.randn(2, ci, h + 2 - 2 * pad, w + 2 - 2 * pad, device=dev),
                       torch.randn(2, co, h, w, device=dev), pad)
# tensor-core kernels (tcgen05): forward / data gradient and weight gradient of the 32 -> 32 layers
xt = torch.randn(2, 32, 16, 128, device=dev)
wt = torch.randn(32, 32, 3, 3, device=dev) * 0.1
conv.conv3x3_tc(xt, wt, torch.randn(32, device=dev), 0.01)
conv.conv3x3_tc(xt, wt, None, 0.0, transpose_flip=True)
conv.conv3x3_wgrad(xt, torch.randn(2, 32, 16, 128, device=dev), 1)
_, sg, _ = conv.conv3x3_tc_signs(xt, wt, torch.randn(32, device=dev), 0.01, want_in_signs=True)
conv.conv3x3_wgrad_thin_bias(torch.randn(2, 2, 16, 128, device=dev), xt)
conv.conv3x3_thin_masked(torch.randn(2, 2, 16, 128, device=dev), torch.randn(32, 2, 3, 3, device=dev), sg, 0.01)
conv.conv3x3_tc_masked(xt, wt, sg, 0.01)
conv.conv3x3_wgrad_bias(xt, torch.randn(2, 32, 16, 128, device=dev))
torch.cuda.synchronize()
print('ok tensor-core convolutions')
img = torch.rand(2, 32, 32, device=dev)
rows = undersampling.cartesian_rows((2, 32, 32), 4, 8, False, np.random.RandomState(1))
batch = undersampling.undersample(img, rows)
net = recnet.construct_model({'num_blocks': 2, 'num_convs': 3, 'num_filters': 32}).to(dev)
torch.nn.functional.mse_loss(net(batch['inp'], batch['kspace'], batch['mask']),
                             batch['target']).backward()
torch.cuda.synchronize()
print('ok refinement + training-step kernels')
