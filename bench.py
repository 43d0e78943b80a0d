#!/usr/bin/env python
"""Benchmark of the DC hot path (BASELINE.json metric: DC-op slices/s + HBM GB/s
at 256^2 fp32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input:
DC forward (noiseless blend, myfft.py:141) + DC adjoint (the backward of
myfft.py:92-128) on B=256 slices of 256x256 fp32 per GPU (BASELINE configs[1]).
Batches are independent, so N GPUs = N ranks each with its own batch (weak
scaling, no collective on the data path).

Prints ONE JSON line (rank 0).  `value` = slices/s with inputs resident in HBM;
`e2e` = the same through DataConsistencyInKspace.perform + autograd backward
with HOST (pinned) buffers, copies inside the timed region; `roofline` = the
strip kernel's algorithmic bytes / its CUDA-event time against the measured
HBM peak; `cpu_baseline` = the oracle port on the host cores (bounded sample).
Extra keys: `noisy_lambda` (the same microbench with noise_lvl = 0.1) and
`recnet_train` (BASELINE configs[2]: RecNet D5C5 fp32 training step, batch 32 per
GPU, with `tf32_convs` as a context number).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'dc_fwd_adjoint_slices_per_s'
UNIT = 'slices/s'
B_PER_GPU = 256
N = 256
ACC = 4


# Libraries print to stdout on their own (NCCL's version banner under NCCL_DEBUG, cuDNN
# notices): keep fd 1 for the ONE JSON line and send everything else to stderr.
_JSON_FD = None


def capture_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------
class ClockSampler(object):
    def __init__(self, index=0, period=0.05):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# --------------------------------------------------------------------------
# synthetic workload: host-chosen Cartesian lines, k0 = mask*FFT2(img)
# --------------------------------------------------------------------------
def make_batch(dev, B, n, seed):
    from csmri_refinement_b200 import undersampling
    g = torch.Generator(device=dev).manual_seed(seed)
    img = torch.rand(B, n, n, device=dev, generator=g)
    rows = undersampling.cartesian_rows((B, n, n), ACC, 8, False, np.random.RandomState(seed))
    batch = undersampling.undersample(img, rows)
    x = torch.randn(B, 2, n, n, device=dev, generator=g)
    w = torch.randn(B, 2, n, n, device=dev, generator=g)
    return x, w, batch['kspace'], batch['mask']


def recnet_train_bench(dev, rank, world, steps=8, warmup=3, batch=32, n=256, blocks=5, convs=5,
                       filters=32, tf32=False):
    """BASELINE configs[2]: RecNet D5C5 MSE training step, batch 32 per GPU,
    batch-sharded, one flat-bucket NCCL allreduce per step, fp32 (no TF32),
    whole step captured in a CUDA graph.  Returns slices/s over all ranks."""
    import torch.distributed as dist
    from csmri_refinement_b200 import parallel, recnet, undersampling
    # tf32=True is torch's stock conv setting (TF32 tensor cores): a context
    # number only, the parity gate (1e-5) is stated for fp32 arithmetic
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True              # let cuDNN pick its fastest fp32 algorithms
    torch.manual_seed(0)                               # identical replicas on every rank
    model = recnet.construct_model({'num_blocks': blocks, 'num_convs': convs,
                                    'num_filters': filters}).to(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    batches = []
    for i in range(2):
        img = torch.rand(batch, n, n, device=dev, generator=g)
        rows = undersampling.cartesian_rows((batch, n, n), ACC, 8, False,
                                            np.random.RandomState(100 + rank * 7 + i))
        batches.append(undersampling.undersample(img, rows))
    trainer = parallel.ShardedTrainer(model, lr=2e-4, cuda_graph=True, assume_row_constant=True)
    for i in range(warmup):
        trainer.step(batches[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = trainer.step(batches[i % 2])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {'metric': 'recnet_train_slices_per_s', 'value': world * batch * steps / (ms * 1e-3),
            'unit': UNIT, 'ms_per_step': ms / steps, 'steps': steps,
            'config': 'RecNet D%dC%d nf=%d, MSE + Adam(2e-4), batch %d/GPU, %dx%d, %s, '
                      'DC = fused strip kernel, CUDA-graph step, '
                      'flat-bucket NCCL allreduce (%d bytes)' % (
                          blocks, convs, filters, batch, n, n,
                          'cuDNN convs with TF32 tensor cores (torch default; not parity-gated)'
                          if tf32 else 'fp32 convs (cuDNN forward / data gradient, TF32 off; weight '
                          'gradient = csmri_conv3x3_wgrad)', trainer.bucket.nbytes()),
            'loss': float(loss.item())}


def cpu_baseline(sample_b=64, reps=6, noise=None):
    """Oracle port (torch.fft restatement of myfft.py:131-163) forward+backward
    on the host cores: the reference's DC has no CPU path of its own."""
    from oracle import dc_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rs = np.random.RandomState(0)
    rows = np.fft.ifftshift(orc.cartesian_lines(sample_b, N, ACC, 8, np.random.RandomState(0)), -1)
    m1 = np.broadcast_to(rows[:, :, None], (sample_b, N, N))
    k0c = m1 * np.fft.fft2(rs.uniform(0, 1, (sample_b, N, N)), norm='ortho')
    k0 = torch.from_numpy(orc.complex_to_planar(k0c))
    mask = torch.from_numpy(np.stack([m1, m1], 1).astype(np.float32))
    x = torch.randn(sample_b, 2, N, N, requires_grad=True)
    w = torch.randn(sample_b, 2, N, N)

    def step():
        out = orc.dc_perform_torch(x, k0, mask, noise)
        (g,) = torch.autograd.grad(out, x, w)
        return g

    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {'value': sample_b / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': '%d reps of fwd+backward on %d slices of %dx%d (torch.fft oracle port, '
                      '%d threads)' % (reps, sample_b, N, N, threads),
            'ms_per_sample_step': dt * 1e3}


def run_reference(args, rank, world):
    """--impl reference: the reference's DC is CUDA-only (pytorch_fft) and pure
    Python, so its CPU stand-in is the oracle port on all host threads."""
    if rank != 0:
        return
    sample_b = 64
    for _ in range(args.warmup):
        cpu_baseline(sample_b, 1)
    t0 = time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = cpu_baseline(sample_b, 1)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    # cpu_baseline() runs one untimed + one timed rep; use its own timed figure
    vals = res['value']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': vals, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['ms_per_sample_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'DC fwd+adjoint, %dx%d fp32, 4x Cartesian mask; CPU sample of %d '
                               'slices per step (B=256/GPU on the GPU arm)' % (N, N, sample_b)},
        'cpu_baseline': {'value': vals, 'unit': UNIT, 'cores': res['cores'], 'kind': 'port',
                         'sample': res['sample']},
        'e2e': {'value': vals, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_ms_per_step_incl_setup': dt * 1e3,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--variant', type=int, default=None, help='kernel tuning variant (debug)')
    ap.add_argument('--no-recnet', action='store_true', help='skip the RecNet training leg')
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the DC path)')
    import torch.distributed as dist
    from csmri_refinement_b200 import _lib, myfft, ops

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    if args.variant is not None:
        _lib.lib().csmri_set_variant(args.variant)

    B = B_PER_GPU
    nbuf = 2   # rotating buffer sets; one step already streams 640 MiB >> 126 MB L2
    sets = [make_batch(dev, B, N, seed=rank * 16 + i) for i in range(nbuf)]
    plans = [myfft.get_plan(k0, mask) for (_, _, k0, mask) in sets]
    assert all(p.row_constant for p in plans)
    outs = [torch.empty_like(sets[0][0]) for _ in range(2)]

    def step(i):
        x, w, _, _ = sets[i % nbuf]
        p = plans[i % nbuf]
        out = ops.dc_cartesian(x, None, p.dtab, p.addend)   # forward
        gx = ops.dc_cartesian(w, None, p.dtab, None)        # adjoint
        return out, gx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank, period=0.002)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # ---- per-kernel roofline: forward and adjoint launches timed separately ----
    def time_kernel(fn, reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def raw_fwd(i):
        x, _, _, _ = sets[i % nbuf]
        p = plans[i % nbuf]
        lib.csmri_dc_forward_cartesian(x.data_ptr(), None, p.dtab.data_ptr(), p.addend.data_ptr(),
                                       outs[0].data_ptr(), B, N, N, stream)

    def raw_adj(i):
        _, w, _, _ = sets[i % nbuf]
        p = plans[i % nbuf]
        lib.csmri_dc_adjoint_cartesian(w.data_ptr(), p.dtab.data_ptr(), outs[1].data_ptr(),
                                       B, N, N, stream)

    reps = max(args.steps, 20)
    ms_fwd = time_kernel(raw_fwd, reps)
    ms_adj = time_kernel(raw_adj, reps)
    clocks = sampler.stop()   # sampled over the timed region and the per-kernel timing loops
    peak, peak_src = measured_peak()
    bytes_fwd, bytes_adj = 24 * N * N * B, 16 * N * N * B
    gbs_fwd = bytes_fwd / (ms_fwd * 1e-3) / 1e9
    gbs_adj = bytes_adj / (ms_adj * 1e-3) / 1e9
    gbs_pair = (bytes_fwd + bytes_adj) / ((ms_fwd + ms_adj) * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_dram_traffic.json')) as f:
            traffic = json.load(f).get('bytes_per_fwd_adj_pair')
    except Exception:
        pass
    roofline = {
        'bound': 'hbm', 'kernel': 'dc_strip_pipev_kernel<256,16,16,...> (forward + adjoint launches)',
        'achieved': gbs_pair, 'peak': peak, 'unit': 'GB/s', 'frac': gbs_pair / peak,
        'traffic': traffic, 'peak_source': peak_src,
        'bytes_per_launch': {'forward': bytes_fwd, 'adjoint': bytes_adj},
        'forward': {'ms': ms_fwd, 'GBps': gbs_fwd, 'frac': gbs_fwd / peak},
        'adjoint': {'ms': ms_adj, 'GBps': gbs_adj, 'frac': gbs_adj / peak},
        'frac_of_nominal_8000': gbs_pair / 8000.0,
    }

    # ---- noisy-lambda variant of configs[1] (myfft.py:139): same kernels, other plan ---
    nplans = [myfft.DCPlan(k0, mask, 0.1) for (_, _, k0, mask) in sets]

    def noisy_step(i):
        x, w, _, _ = sets[i % nbuf]
        p = nplans[i % nbuf]
        lib.csmri_dc_forward_cartesian(x.data_ptr(), None, p.dtab.data_ptr(), p.addend.data_ptr(),
                                       outs[0].data_ptr(), B, N, N, stream)
        lib.csmri_dc_adjoint_cartesian(w.data_ptr(), p.dtab.data_ptr(), outs[1].data_ptr(),
                                       B, N, N, stream)

    ms_noisy = time_kernel(noisy_step, reps)
    if world > 1:
        t = torch.tensor([ms_noisy], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_noisy = float(t.item())
    noisy = {'noise_lvl': 0.1, 'value': world * B / (ms_noisy * 1e-3), 'unit': UNIT,
             'ms_per_step': ms_noisy,
             'frac': (bytes_fwd + bytes_adj) / (ms_noisy * 1e-3) / 1e9 / peak}
    del nplans

    # ---- e2e: public API, host buffers, copies inside the timed region ---------
    x0, w0, k00, m0 = sets[0]
    hx, hk0, hm, hw = (t.cpu().pin_memory() for t in (x0, k00, m0, w0))
    h_out = torch.empty_like(hx).pin_memory()
    h_gx = torch.empty_like(hx).pin_memory()
    dc = myfft.DataConsistencyInKspace()

    def e2e_step():
        x = hx.to(dev, non_blocking=True).requires_grad_(True)
        k0 = hk0.to(dev, non_blocking=True)
        m = hm.to(dev, non_blocking=True)
        w = hw.to(dev, non_blocking=True)
        out = dc.perform(x, k0, m)
        (gx,) = torch.autograd.grad(out, x, w)
        h_out.copy_(out.detach(), non_blocking=True)
        h_gx.copy_(gx, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        e2e_step()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    tensor_bytes = hx.numel() * 4
    e2e_simple = {'value': world * B * e2e_steps / (e2e_ms * 1e-3), 'unit': UNIT,
                  'h2d_bytes_per_step': 4 * tensor_bytes, 'd2h_bytes_per_step': 2 * tensor_bytes,
                  'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps,
                  'api': 'DataConsistencyInKspace.perform + autograd backward, pinned host '
                         'x/k0/mask/grad-seed in, out/grad_x back, copies and compute in sequence; '
                         'includes the per-batch prepare'}

    # same work through the streamed public API: chunks of 64 slices on three
    # streams (H2D / DC forward+adjoint incl. prepare / D2H), full-duplex PCIe
    from csmri_refinement_b200 import hostpipe
    pipe = hostpipe.HostDCPipeline(dev, chunk=64, depth=3)
    for _ in range(2):
        pipe.forward_backward(hx, hk0, hm, hw, h_out, h_gx)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        pipe.forward_backward(hx, hk0, hm, hw, h_out, h_gx)
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {'value': world * B * e2e_steps / (e2e_ms * 1e-3), 'unit': UNIT,
           'h2d_bytes_per_step': 4 * tensor_bytes, 'd2h_bytes_per_step': 2 * tensor_bytes,
           'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps,
           'api': 'hostpipe.HostDCPipeline.forward_backward: pinned host x/k0/mask/grad-seed in, '
                  'out/grad_x back to pinned host, 64-slice chunks on 3 streams; includes the '
                  'per-chunk prepare and the row-constancy verification read',
           'unpipelined': e2e_simple}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'DC operator microbench (BASELINE configs[1]): batch %d/GPU of '
                               '%dx%d fp32 slices, 4x Cartesian mask, noiseless, fwd+adjoint'
                               % (B, N, N),
                   'l2': 'inputs larger than L2: one step streams 640 MiB over 2 rotating '
                         'buffer sets (L2 = 126 MB)',
                   'sharding': 'batch-sharded, one rank per GPU, no data-path collective'},
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': 2 * args.steps, 'clocks': clocks,
        'hbm_GBps_fwd_adj': value / world * 40 * N * N / 1e9,
        'noisy_lambda': noisy,
    }
    if not args.no_recnet:
        try:
            line['recnet_train'] = recnet_train_bench(dev, rank, world)
            line['recnet_train']['tf32_convs'] = recnet_train_bench(dev, rank, world, tf32=True)
        except Exception as e:  # secondary leg: never lose the headline line
            line.setdefault('recnet_train', {})['error'] = repr(e)[:300]
        torch.backends.cudnn.allow_tf32 = False
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline()
    elif rank == 0:
        line['cpu_baseline'] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == '__main__':
    main()
