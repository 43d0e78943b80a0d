"""the CPU baselines B1-B4, measured with the oracle on this host,
next to the same work on the GPU (if one is present).  Reported numbers, not targets."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import dc_oracle as orc  # noqa: E402


def timeit(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cpu = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                cpu = line.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    res = {'cpu': cpu, 'logical_cores': threads, 'torch_threads': torch.get_num_threads()}
    n, B = 256, 64
    rs = np.random.RandomState(0)
    m1 = orc.cartesian_mask((B, n, n), 4, 8, False, np.random.RandomState(0))
    img = rs.uniform(0, 1, (B, n, n))
    x_u, x_fu = orc.undersample(img, m1, rng=np.random.RandomState(0))
    xin = rs.normal(size=(B, n, n)) + 1j * rs.normal(size=(B, n, n))
    # B1: numpy cs.data_consistency, complex128, 1 thread
    t = timeit(lambda: orc.cs_data_consistency_np(xin, x_fu, m1), 3)
    res['B1_numpy_dc_fwd'] = {'ms': t * 1e3, 'slices_per_s': B / t, 'threads': 1,
                              'what': 'cs.data_consistency restatement, complex128, B=64, 256^2'}
    # B2: torch.fft restatement fp32, forward and forward+backward, noiseless and lambda=0.1
    x = torch.from_numpy(orc.complex_to_planar(xin)).requires_grad_(True)
    k0 = torch.from_numpy(orc.complex_to_planar(x_fu))
    mask = torch.from_numpy(np.stack([m1, m1], 1).astype(np.float32))
    w = torch.randn(B, 2, n, n)
    for v in (None, 0.1):
        with torch.no_grad():
            t = timeit(lambda: orc.dc_perform_torch(x, k0, mask, v), 5)
        res['B2_torch_dc_fwd_noise_%s' % v] = {'ms': t * 1e3, 'slices_per_s': B / t, 'threads': threads}

        def fb():
            out = orc.dc_perform_torch(x, k0, mask, v)
            torch.autograd.grad(out, x, w)
        t = timeit(fb, 5)
        res['B2_torch_dc_fwd_bwd_noise_%s' % v] = {'ms': t * 1e3, 'slices_per_s': B / t,
                                                    'threads': threads}
    # B3: config 1 - RecNet D5C5 forward, B=4, 256^2, 4x mask, DC = oracle (nf 32 and 64)
    from csmri_refinement_b200 import recnet
    b4 = {'inp': torch.from_numpy(orc.to_tensor_format(x_u[:4])),
          'kspace': torch.from_numpy(orc.to_tensor_format(x_fu[:4])),
          'mask': torch.from_numpy(orc.to_tensor_format(m1[:4], mask=True))}
    for nf in (32, 64):
        torch.manual_seed(0)
        net = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': nf},
                                     dc_factory=orc.OracleDataConsistencyInKspace)
        with torch.no_grad():
            t = timeit(lambda: net(b4['inp'], b4['kspace'], b4['mask']), 2)
        res['B3_recnet_D5C5_nf%d_fwd_cpu' % nf] = {'ms': t * 1e3, 'slices_per_s': 4 / t,
                                                   'threads': threads}
        if torch.cuda.is_available():
            torch.backends.cudnn.allow_tf32 = False
            torch.manual_seed(0)
            gnet = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': nf}).cuda()
            gb = {k: v.cuda() for k, v in b4.items()}
            with torch.no_grad():
                out = gnet(gb['inp'], gb['kspace'], gb['mask'])
                ref = net(b4['inp'], b4['kspace'], b4['mask'])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    gnet(gb['inp'], gb['kspace'], gb['mask'])
                torch.cuda.synchronize()
                tg = (time.perf_counter() - t0) / 10
            res['B3_recnet_D5C5_nf%d_fwd_b200' % nf] = {
                'ms': tg * 1e3, 'slices_per_s': 4 / tg,
                'rel_l2_vs_cpu_oracle_net': orc.rel_l2(out.cpu().numpy(), ref.numpy())}
    # B4: host data path per sample (what csmri_undersample replaces)
    for size in (256, 512):
        im1 = rs.uniform(0, 1, (1, size, size))

        def host():
            m = orc.cartesian_mask((1, size, size), 4, 8, False, np.random)
            orc.undersample(im1, m, rng=np.random)
        t = timeit(host, 10)
        res['B4_host_mask_undersample_%d' % size] = {'ms_per_sample': t * 1e3, 'threads': 1}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'cpu_baselines.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
