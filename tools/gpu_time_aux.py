"""Time the once-per-batch / loader-side / general-mask kernels (context numbers)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import myfft, ops, undersampling  # noqa: E402


def time_fn(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    dev = torch.device('cuda:0')
    out = []
    for n, B in ((256, 256), (512, 64), (128, 1024), (320, 164)):
        img = torch.rand(B, n, n, device=dev)
        rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
        rows_d = torch.from_numpy(rows).to(dev)
        batch = undersampling.undersample(img, rows_d)
        k0, mask = batch['kspace'], batch['mask']
        x = torch.randn(B, 2, n, n, device=dev)
        nb = 8.0 * n * n * B   # bytes of one (B,2,n,n) tensor
        r = {'N': n, 'B': B}
        t = time_fn(lambda: ops.dc_prepare(k0, mask, 0.0))
        r['prepare_us'] = round(t, 1)
        r['prepare_GBps'] = round(4 * nb / t / 1e3)       # read k0, mask; write addend (+mask ch1)
        t = time_fn(lambda: ops.undersample(img, rows_d))
        r['undersample_us'] = round(t, 1)
        r['undersample_GBps_alg'] = round(4.5 * nb / t / 1e3)   # read img (0.5), write 4 tensors
        t = time_fn(lambda: ops.fft2_planar(x))
        r['fft2_us'] = round(t, 1)
        r['fft2_GBps_alg'] = round(2 * nb / t / 1e3)
        t = time_fn(lambda: ops.dc_general(x, None, k0, mask, 0.0))
        r['general_fwd_us'] = round(t, 1)
        r['general_fwd_GBps_alg'] = round(3 * nb / t / 1e3)
        print(json.dumps(r), flush=True)
        out.append(r)
        myfft.clear_plan_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'aux_timing.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
