"""Time the strip-kernel tuning variants on the GPU box (writes gpurun_out/probe.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, undersampling  # noqa: E402


def time_fn(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    lib = _lib.lib()
    dev = torch.device('cuda:0')
    stream = torch.cuda.current_stream().cuda_stream
    res = []
    cases = [(256, 256, [0, 1, 10, 11]), (512, 64, [0, 1, 10]), (128, 1024, [0, 10, 11]),
             (64, 4096, [0]), (1024, 16, [0])]
    extra = [int(v) for v in os.environ.get('PROBE_VARIANTS', '').split(',') if v]
    for n, B, variants in cases:
        img = torch.rand(B, n, n, device=dev)
        rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
        batch = undersampling.undersample(img, rows)
        plan = myfft.get_plan(batch['kspace'], batch['mask'])
        xs = [torch.randn(B, 2, n, n, device=dev) for _ in range(2)]
        out = torch.empty_like(xs[0])
        for v in variants + (extra if n == 256 else []):
            lib.csmri_set_variant(v)
            it = [0]

            def fwd():
                it[0] += 1
                _lib.check(lib.csmri_dc_forward_cartesian(
                    xs[it[0] % 2].data_ptr(), None, plan.dtab.data_ptr(), plan.addend.data_ptr(),
                    out.data_ptr(), B, n, n, stream))

            def adj():
                it[0] += 1
                _lib.check(lib.csmri_dc_adjoint_cartesian(
                    xs[it[0] % 2].data_ptr(), plan.dtab.data_ptr(), out.data_ptr(), B, n, n,
                    stream))

            # correctness of the variant against variant 0 on identical inputs
            it[0] = 0
            fwd()
            got_f = out.clone()
            it[0] = 0
            adj()
            got_a = out.clone()
            if v == variants[0]:
                ref_f, ref_a = got_f, got_a
            err_f = ((got_f - ref_f).norm() / ref_f.norm()).item()
            err_a = ((got_a - ref_a).norm() / ref_a.norm()).item()
            tf, ta = time_fn(fwd), time_fn(adj)
            r = {'N': n, 'B': B, 'variant': v, 'relerr_vs_v0': [err_f, err_a], 'fwd_ms': tf, 'adj_ms': ta,
                 'fwd_GBps': 24 * n * n * B / tf / 1e6, 'adj_GBps': 16 * n * n * B / ta / 1e6,
                 'pair_GBps': 40 * n * n * B / (tf + ta) / 1e6,
                 'pair_slices_per_s': B / (tf + ta) * 1e3}
            print(json.dumps(r), flush=True)
            res.append(r)
        lib.csmri_set_variant(0)
        # general path and torch.fft on the same GPU for context
        k0, mask = batch['kspace'], batch['mask']
        from csmri_refinement_b200 import ops
        tg = time_fn(lambda: ops.dc_general(xs[0], None, k0, mask, 0.0), 10)

        def torch_dc():
            xc = torch.complex(xs[0][:, 0], xs[0][:, 1])
            kc = torch.fft.fft2(xc, norm='ortho')
            k = torch.stack([kc.real, kc.imag], 1)
            o = (1 - mask) * k + k0
            oc = torch.fft.ifft2(torch.complex(o[:, 0], o[:, 1]), norm='ortho')
            return torch.stack([oc.real, oc.imag], 1)

        tt = time_fn(torch_dc, 10)
        r = {'N': n, 'B': B, 'general_fwd_ms': tg, 'general_fwd_GBps': 24 * n * n * B / tg / 1e6,
             'torch_cufft_fwd_ms': tt, 'torch_cufft_fwd_GBps': 24 * n * n * B / tt / 1e6}
        print(json.dumps(r), flush=True)
        res.append(r)
        del batch, plan, xs, out, img
        myfft.clear_plan_cache()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
