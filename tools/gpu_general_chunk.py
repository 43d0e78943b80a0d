"""General-mask DC path: sweep of the chunk size that keeps the hybrid scratch in L2
(csmri_set_tuning key 9, MiB of scratch per chunk; 0 = the whole batch per pass)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, ops, undersampling  # noqa: E402
from tools.gpu_time_aux import time_fn  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    lib = _lib.lib()
    out = []
    for n, B in ((256, 256), (512, 64), (128, 1024), (320, 164)):
        img = torch.rand(B, n, n, device=dev)
        rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
        batch = undersampling.undersample(img, torch.from_numpy(rows).to(dev))
        k0, mask = batch['kspace'], batch['mask']
        x = torch.randn(B, 2, n, n, device=dev)
        nb = 8.0 * n * n * B
        ref = None
        for mib in (0, 4, 8, 12, 16, 24, 32, 48, 64):
            lib.csmri_set_tuning(9, mib)
            y = ops.dc_general(x, None, k0, mask, 0.1)
            if ref is None:
                ref = y.clone()
            same = bool(torch.equal(y, ref))
            tf = time_fn(lambda: ops.dc_general(x, None, k0, mask, 0.0), 30)
            ta = time_fn(lambda: ops.dc_general_adjoint(x, mask, 0.0), 30)
            r = {'N': n, 'B': B, 'chunk_mib': mib, 'fwd_us': round(tf, 1), 'adj_us': round(ta, 1),
                 'fwd_GBps_24N2': round(3 * nb / tf / 1e3), 'bit_identical_to_unchunked': same}
            print(json.dumps(r), flush=True)
            out.append(r)
        lib.csmri_set_tuning(9, 0)
        myfft.clear_plan_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_general_chunk_sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
