"""BASELINE configs[3]: DC slice-size / acceleration sweep (fwd+adjoint), one rank per GPU.

  python tools/sweep.py                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py

Sizes 128/256/320/512 at 4x/8x/12x Cartesian masks, batch chosen so one tensor is 128 MiB per
GPU (B = 1024/256/164/64).  Every case is also checked against the CPU oracle on 2 slices.
Rank 0 prints one JSON line per case and writes gpurun_out/sweep_n<N>.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from csmri_refinement_b200 import _lib, myfft, parallel, undersampling  # noqa: E402
from oracle import dc_oracle as orc  # noqa: E402  (checker only)

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def main():
    rank, world, dev = parallel.init_distributed()
    lib = _lib.lib()
    res = []
    for n, B in ((128, 1024), (256, 256), (320, 164), (512, 64)):
        for acc in (4, 8, 12):
            g = torch.Generator(device=dev).manual_seed(rank * 100 + n + acc)
            img = torch.rand(B, n, n, device=dev, generator=g)
            rows = undersampling.cartesian_rows((B, n, n), acc, 8, False,
                                                np.random.RandomState(rank * 100 + acc))
            assert int(rows[0].sum()) == n // acc
            batch = undersampling.undersample(img, rows)
            plan = myfft.get_plan(batch['kspace'], batch['mask'])
            assert plan.row_constant
            xs = [torch.randn(B, 2, n, n, device=dev, generator=g) for _ in range(2)]
            out = torch.empty_like(xs[0])
            gx = torch.empty_like(xs[0])
            stream = torch.cuda.current_stream().cuda_stream

            def step(i):
                lib.csmri_dc_forward_cartesian(xs[i % 2].data_ptr(), None, plan.dtab.data_ptr(),
                                               plan.addend.data_ptr(), out.data_ptr(), B, n, n, stream)
                lib.csmri_dc_adjoint_cartesian(xs[(i + 1) % 2].data_ptr(), plan.dtab.data_ptr(),
                                               gx.data_ptr(), B, n, n, stream)

            step(0)
            torch.cuda.synchronize()
            ref = orc.dc_perform_np(xs[0][:2].cpu().numpy(), batch['kspace'][:2].cpu().numpy(),
                                    batch['mask'][:2].cpu().numpy())
            err = orc.rel_l2(out[:2].cpu().numpy(), ref)
            gref = orc.dc_adjoint_np(xs[1][:2].cpu().numpy(), batch['mask'][:2].cpu().numpy())
            gerr = orc.rel_l2(gx[:2].cpu().numpy(), gref)
            assert err < 1e-5 and gerr < 1e-5, (n, acc, err, gerr)
            for i in range(5):
                step(i)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50
            a.record()
            for i in range(reps):
                step(i)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            gbs = 40.0 * n * n * B / (ms * 1e-3) / 1e9
            r = {'N': n, 'acc': acc, 'B_per_gpu': B, 'n_gpus': world, 'ms_fwd_adj': round(ms, 4),
                 'slices_per_s': round(world * B / (ms * 1e-3)), 'GBps_per_gpu': round(gbs),
                 'frac_of_measured_peak': round(gbs / PEAK, 3), 'sampled_lines': n // acc,
                 'rel_l2_fwd': float('%.2e' % err), 'rel_l2_adj': float('%.2e' % gerr)}
            if rank == 0:
                print(json.dumps(r), flush=True)
            res.append(r)
            del batch, plan, xs, out, gx, img
            myfft.clear_plan_cache()
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'sweep_n%d.json' % world), 'w') as f:
            json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
