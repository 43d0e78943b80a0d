"""Time the strip-kernel tuning variants on the GPU box (writes gpurun_out/probe.json).

PROBE_CASES="N:B,N:B"       sizes (default 256:256,512:64,128:1024)
PROBE_LIST="v,v"            kernel variants (0 default, 1 one-column pipeline, 2 direct)
PROBE_CONTEXT=0             skip the general-path / torch.fft context timings
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, ops, undersampling  # noqa: E402


def time_fn(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if os.environ.get('PROBE_GRAPH', '0') == '1':
        # replay the launches from a CUDA graph: no host launch path in the timing
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    lib = _lib.lib()
    lib.csmri_set_tuning(4, int(os.environ.get('PROBE_PDL', '1')))
    dev = torch.device('cuda:0')
    stream = torch.cuda.current_stream().cuda_stream
    cases = [tuple(int(v) for v in c.split(':')) for c in
             os.environ.get('PROBE_CASES', '256:256,512:64,128:1024').split(',')]
    plist = [tuple(int(v) for v in c.split(':')) for c in
             os.environ.get('PROBE_LIST', '0:0').split(',')]
    plist = [p if len(p) == 3 else p + (0,) for p in plist]   # (variant, prefetch, dephase)
    context = os.environ.get('PROBE_CONTEXT', '1') != '0'
    res = []
    for n, B in cases:
        img = torch.rand(B, n, n, device=dev)
        rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
        batch = undersampling.undersample(img, rows)
        plan = myfft.get_plan(batch['kspace'], batch['mask'])
        xs = [torch.randn(B, 2, n, n, device=dev) for _ in range(2)]
        out = torch.empty_like(xs[0])
        ref_f = ref_a = None
        for v, pf, dph in [(0, 0, 0)] + plist:
            lib.csmri_set_tuning(0, v)
            it = [0]

            def fwd():
                it[0] += 1
                _lib.check(lib.csmri_dc_forward_cartesian(
                    xs[it[0] % 2].data_ptr(), None, plan.dtab.data_ptr(), plan.addend.data_ptr(),
                    out.data_ptr(), B, n, n, torch.cuda.current_stream().cuda_stream))

            def adj():
                it[0] += 1
                _lib.check(lib.csmri_dc_adjoint_cartesian(
                    xs[it[0] % 2].data_ptr(), plan.dtab.data_ptr(), out.data_ptr(), B, n, n,
                    torch.cuda.current_stream().cuda_stream))

            it[0] = 0
            fwd()
            got_f = out.clone()
            it[0] = 0
            adj()
            got_a = out.clone()
            if ref_f is None:
                ref_f, ref_a = got_f, got_a
            err_f = ((got_f - ref_f).norm() / ref_f.norm()).item()
            err_a = ((got_a - ref_a).norm() / ref_a.norm()).item()
            tf, ta = time_fn(fwd), time_fn(adj)
            r = {'N': n, 'B': B, 'variant': v, 'pf': pf, 'dephase': dph, 'relerr_vs_v0': [err_f, err_a],
                 'fwd_us': round(tf * 1e3, 2), 'adj_us': round(ta * 1e3, 2),
                 'fwd_GBps': round(24 * n * n * B / tf / 1e6), 'adj_GBps': round(16 * n * n * B / ta / 1e6),
                 'pair_GBps': round(40 * n * n * B / (tf + ta) / 1e6),
                 'pair_frac_6459': round(40 * n * n * B / (tf + ta) / 1e6 / 6459, 3),
                 'pair_slices_per_s': round(B / (tf + ta) * 1e3)}
            print(json.dumps(r), flush=True)
            res.append(r)
        lib.csmri_set_tuning(0, 0)
        if context:
            k0, mask = batch['kspace'], batch['mask']
            tg = time_fn(lambda: ops.dc_general(xs[0], None, k0, mask, 0.0), 10)

            def torch_dc():
                xc = torch.complex(xs[0][:, 0], xs[0][:, 1])
                kc = torch.fft.fft2(xc, norm='ortho')
                k = torch.stack([kc.real, kc.imag], 1)
                o = (1 - mask) * k + k0
                oc = torch.fft.ifft2(torch.complex(o[:, 0], o[:, 1]), norm='ortho')
                return torch.stack([oc.real, oc.imag], 1)

            tt = time_fn(torch_dc, 10)
            r = {'N': n, 'B': B, 'general_fwd_us': round(tg * 1e3, 1),
                 'general_fwd_GBps': round(24 * n * n * B / tg / 1e6),
                 'torch_cufft_fwd_us': round(tt * 1e3, 1),
                 'torch_cufft_fwd_GBps': round(24 * n * n * B / tt / 1e6)}
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
