"""A/B timing of the strip-kernel variants (csmri_set_variant) per slice size.

  python tools/gpu_variant_probe.py [N:B:variant,variant ...]

Forward and adjoint launches timed separately through the raw C ABI (CUDA
events, 2 rotating input sets, 128 MiB per tensor), each variant checked
against the fp64 oracle on 2 slices.  Writes gpurun_out/variant_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, undersampling  # noqa: E402
from oracle import dc_oracle as orc  # noqa: E402  (checker only)

PEAK = 6459.0


def main():
    cases = sys.argv[1:] or ['512:64:0,3', '1024:16:0,3', '256:256:0']
    lib = _lib.lib()
    dev = torch.device('cuda:0')
    res = []
    for case in cases:
        n, B, variants = case.split(':')
        n, B = int(n), int(B)
        g = torch.Generator(device=dev).manual_seed(n)
        img = torch.rand(B, n, n, device=dev, generator=g)
        rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(n))
        batch = undersampling.undersample(img, rows)
        plan = myfft.get_plan(batch['kspace'], batch['mask'])
        xs = [torch.randn(B, 2, n, n, device=dev, generator=g) for _ in range(2)]
        out, gx = torch.empty_like(xs[0]), torch.empty_like(xs[0])
        stream = torch.cuda.current_stream().cuda_stream
        ref = orc.dc_perform_np(xs[0][:2].cpu().numpy(), batch['kspace'][:2].cpu().numpy(),
                                batch['mask'][:2].cpu().numpy())
        gref = orc.dc_adjoint_np(xs[1][:2].cpu().numpy(), batch['mask'][:2].cpu().numpy())
        for v in (int(t) for t in variants.split(',')):
            lib.csmri_set_variant(v)

            def fwd(i):
                _lib.check(lib.csmri_dc_forward_cartesian(
                    xs[i % 2].data_ptr(), None, plan.dtab.data_ptr(), plan.addend.data_ptr(),
                    out.data_ptr(), B, n, n, stream))

            def adj(i):
                _lib.check(lib.csmri_dc_adjoint_cartesian(
                    xs[i % 2].data_ptr(), plan.dtab.data_ptr(), gx.data_ptr(), B, n, n, stream))

            fwd(0)
            adj(1)
            torch.cuda.synchronize()
            err = orc.rel_l2(out[:2].cpu().numpy(), ref)
            gerr = orc.rel_l2(gx[:2].cpu().numpy(), gref)
            t = {}
            for name, fn in (('fwd', fwd), ('adj', adj)):
                for i in range(5):
                    fn(i)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(100):
                    fn(i)
                b.record()
                torch.cuda.synchronize()
                t[name] = a.elapsed_time(b) / 100 * 1e3
            pair = 40.0 * n * n * B / ((t['fwd'] + t['adj']) * 1e-6) / 1e9
            r = {'N': n, 'B': B, 'variant': v, 'fwd_us': round(t['fwd'], 2),
                 'adj_us': round(t['adj'], 2), 'pair_frac': round(pair / PEAK, 3),
                 'fwd_frac': round(24.0 * n * n * B / (t['fwd'] * 1e-6) / 1e9 / PEAK, 3),
                 'adj_frac': round(16.0 * n * n * B / (t['adj'] * 1e-6) / 1e9 / PEAK, 3),
                 'rel_l2_fwd': float('%.2e' % err), 'rel_l2_adj': float('%.2e' % gerr)}
            print(json.dumps(r), flush=True)
            res.append(r)
        lib.csmri_set_variant(0)
        del batch, plan, xs, out, gx, img
        myfft.clear_plan_cache()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'variant_probe.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
