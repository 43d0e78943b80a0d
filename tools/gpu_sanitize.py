"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import myfft, ops, rec_transforms, undersampling  # noqa: E402

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
