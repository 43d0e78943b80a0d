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
The timed region of `value` is ONE CUDA graph holding the K steps (2K launches
of the csmri::dc_cartesian custom op), so it measures the GPU and not the Python
dispatch of a 50-60 us kernel; `roofline.frac` is derived from that same region
(`roofline.kernel_frac` / `forward` / `adjoint` are the per-launch figures from
raw C-ABI loops).
Extra keys: `gpu_baseline_torch_fft` (cuFFT + pointwise, the reference's
structure, on the same GPU and tensors), `noisy_lambda` (the same microbench
with noise_lvl = 0.1), `recnet_train` (BASELINE configs[2]: RecNet D5C5
training step, batch 32 per GPU, 32->32 layers on the tcgen05 tensor cores at
fp32-level accuracy; `cudnn_fp32` = round 1's cuDNN fp32 path and `tf32_convs`
= cuDNN's plain-TF32 path as context numbers),
`recnet_1json` (configs/1-recnet.json unchanged: D3C3-nf32, GLOBAL batch 20,
512^2, 8x; strong scaling) and `refinement_train` (BASELINE configs[4]).
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
                       filters=32, mode='tc'):
    """BASELINE configs[2]: RecNet D5C5 MSE training step, batch 32 per GPU,
    batch-sharded, one flat-bucket NCCL allreduce per step, whole step captured
    in a CUDA graph.  Returns slices/s over all ranks.  ``mode``:
    'tc'         the product path: 32 -> 32 layers (forward, data and weight gradient) on
                 the tcgen05 tensor cores with the error-compensated TF32 split
                 (fp32-level accuracy, parity-gated), thin layers / epilogues / DC = library kernels;
    'cudnn_fp32' context: round 1's path - cuDNN fp32 forward / data gradient (TF32 off) and the
                 SIMT weight-gradient kernel;
    'cudnn_tf32' context: torch's stock setting, cuDNN with plain TF32 tensor cores (10-bit
                 mantissa products, NOT parity-gated) + the SIMT weight gradient."""
    import torch.distributed as dist
    from csmri_refinement_b200 import _lib, conv, parallel, recnet, undersampling
    tf32 = mode == 'cudnn_tf32'
    conv.set_tensor_core_conv(mode == 'tc')
    _lib.lib().csmri_set_tuning(7, 1 if mode == 'tc' else 0)
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
    conv.set_tensor_core_conv(True)
    _lib.lib().csmri_set_tuning(7, 1)
    how = {'tc': '32->32 convolutions (forward, data gradient, weight gradient) = tcgen05 kernels, '
                 '3-term TF32 split with fp32 accumulation (fp32-level accuracy, parity-gated); '
                 'thin layers and bias/LeakyReLU epilogues = library kernels',
           'cudnn_fp32': 'context: cuDNN fp32 forward / data gradient (TF32 off), SIMT weight '
                         'gradient (round 1 path)',
           'cudnn_tf32': 'context: cuDNN convs with plain TF32 tensor cores (torch default; not '
                         'parity-gated), SIMT weight gradient'}[mode]
    return {'metric': 'recnet_train_slices_per_s', 'value': world * batch * steps / (ms * 1e-3),
            'unit': UNIT, 'ms_per_step': ms / steps, 'steps': steps,
            'config': 'RecNet D%dC%d nf=%d, MSE + Adam(2e-4), batch %d/GPU, %dx%d, %s, '
                      'DC = fused strip kernel, CUDA-graph step, '
                      'flat-bucket NCCL allreduce (%d bytes)' % (
                          blocks, convs, filters, batch, n, n, how, trainer.bucket.nbytes()),
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


def pin_rank_to_cores(local_rank, local_world):
    """Give every rank its own contiguous slice of the host cores this process
    may use (8 ranks + their NVML / autograd threads otherwise migrate over the
    same 32 cores).  Returns the core list, or None if affinity is unavailable."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(cores) < local_world:
            return cores
        per = len(cores) // local_world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return None


def torch_fft_dc(x, k0, mask):
    """The reference's structure on this GPU: library FFT (cuFFT through
    torch.fft) + separate pointwise kernels, myfft.py:153-163 noiseless branch.
    Only a same-box comparison number ('beat the cuFFT+pointwise
    restatement'); never on the product path."""
    kc = torch.fft.fft2(torch.complex(x[:, 0], x[:, 1]), norm='ortho')
    k = torch.stack((kc.real, kc.imag), 1)
    out = (1 - mask) * k + k0
    oc = torch.fft.ifft2(torch.complex(out[:, 0], out[:, 1]), norm='ortho')
    return torch.stack((oc.real, oc.imag), 1)


def event_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def rank_stats(ms, dev, world):
    """(max over ranks, {'min','median','max'} over ranks) of a per-rank time."""
    if world <= 1:
        return ms, {'min': ms, 'median': ms, 'max': ms}
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    v = sorted(float(x.item()) for x in all_t)
    return v[-1], {'min': v[0], 'median': float(np.median(v)), 'max': v[-1]}


def recnet_1json_bench(dev, rank, world, steps=6, warmup=3):
    """configs/1-recnet.json UNCHANGED: D3C3-nf32, global batch 20, 512x512,
    8x Cartesian undersampling, MSE + Adam(2e-4).  The JSON's
    batch_size is the GLOBAL batch (DataParallel splits it), so this leg is
    strong scaling: 20 slices per step over however many GPUs there are."""
    import torch.distributed as dist
    from csmri_refinement_b200 import harness
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    conf = harness.load_config(harness.config_path('1-recnet.json'))
    trainer, local_b = harness.recnet_trainer(conf, dev, rank, world, cuda_graph=True)
    batches = [harness.synthetic_batch(conf, local_b, dev, seed=1000 + 16 * rank + i)
               for i in range(2)]
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
    ms, stats = rank_stats(a.elapsed_time(b) / steps, dev, world)
    n_img = batches[0]['inp'].shape[-1]
    return {'metric': 'recnet_train_slices_per_s', 'value': int(conf.batch_size) / (ms * 1e-3),
            'unit': UNIT, 'ms_per_step': ms, 'ms_per_step_ranks': stats, 'steps': steps,
            'scaling': 'strong', 'global_batch': int(conf.batch_size), 'local_batch': local_b,
            'config': 'configs/1-recnet.json unchanged: RecNet D%dC%d nf=%d, global batch %d '
                      '(this rank: %d), %dx%d, %dx Cartesian, MSE + Adam(%g), fp32 (TF32 off), '
                      'CUDA-graph step, gradient weighted by shard size, flat-bucket allreduce '
                      '(%d bytes)' % (conf.model['num_blocks'], conf.model['num_convs'],
                                      conf.model['num_filters'], conf.batch_size, local_b, n_img,
                                      n_img, conf.undersampling['acceleration_factor'],
                                      conf.optimizer['learning_rate'], trainer.bucket.nbytes()),
            'loss': float(loss.item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--variant', type=int, default=None, help='kernel tuning variant (debug)')
    ap.add_argument('--no-recnet', action='store_true', help='skip the RecNet training legs')
    ap.add_argument('--no-refinement', action='store_true', help='skip the config-5 leg')
    ap.add_argument('--eager', action='store_true',
                    help='time the headline through per-step Python dispatch instead of one '
                         'CUDA graph (debug)')
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', str(world)))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the DC path)')
    cores = pin_rank_to_cores(local_rank, local_world)
    import torch.distributed as dist
    from csmri_refinement_b200 import _lib, hostpipe, myfft, ops, undersampling

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

    # ---- headline: W warm-up steps, then EXACTLY K steps in the timed region ----
    # The K steps (2K launches of the product's custom op) are captured into ONE
    # CUDA graph, so the timed region is a single host call: a 50-60 us kernel
    # issued step by step from Python measures the host (8 ranks share 32 cores)
    # as much as the GPU.
    for i in range(args.warmup):
        step(i)
    barrier()
    graph = None
    if not args.eager:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(args.steps):
                step(i)
        graph.replay()            # untimed: first replay uploads the graph
    sampler = ClockSampler(local_rank, period=0.05)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(args.steps):
            step(i)
    ev1.record()
    barrier()
    ms_local = ev0.elapsed_time(ev1)
    ms, ms_ranks = rank_stats(ms_local, dev, world)
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # ---- per-kernel figures: forward and adjoint launches timed separately -------
    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    it = [0]

    def raw_fwd():
        i = it[0] = it[0] + 1
        x, _, _, _ = sets[i % nbuf]
        p = plans[i % nbuf]
        lib.csmri_dc_forward_cartesian(x.data_ptr(), None, p.dtab.data_ptr(), p.addend.data_ptr(),
                                       outs[0].data_ptr(), B, N, N, stream)

    def raw_adj():
        i = it[0] = it[0] + 1
        _, w, _, _ = sets[i % nbuf]
        p = plans[i % nbuf]
        lib.csmri_dc_adjoint_cartesian(w.data_ptr(), p.dtab.data_ptr(), outs[1].data_ptr(),
                                       B, N, N, stream)

    reps = max(args.steps, 50)
    ms_fwd = event_ms(raw_fwd, reps, warm=3)
    ms_adj = event_ms(raw_adj, reps, warm=3)
    peak, peak_src = measured_peak()
    bytes_fwd, bytes_adj = 24 * N * N * B, 16 * N * N * B
    bytes_step = bytes_fwd + bytes_adj
    gbs_fwd = bytes_fwd / (ms_fwd * 1e-3) / 1e9
    gbs_adj = bytes_adj / (ms_adj * 1e-3) / 1e9
    gbs_step = bytes_step / (ms_per_step * 1e-3) / 1e9       # same region as `value`
    gbs_kernels = bytes_step / ((ms_fwd + ms_adj) * 1e-3) / 1e9
    traffic = None
    for name in ('r2_dram_traffic.json', 'r1_dram_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                traffic = json.load(f).get('bytes_per_fwd_adj_pair')
            break
        except Exception:
            pass
    roofline = {
        'bound': 'hbm', 'kernel': 'dc_strip_pipev_kernel<256,16,16,...> (forward + adjoint launches)',
        'achieved': gbs_step, 'peak': peak, 'unit': 'GB/s', 'frac': gbs_step / peak,
        'traffic': traffic, 'peak_source': peak_src,
        'derived_from': 'the timed region of `value`: (24+16)*N^2*B bytes / ms_per_step',
        'bytes_per_step': bytes_step,
        'bytes_per_launch': {'forward': bytes_fwd, 'adjoint': bytes_adj},
        'kernel_frac': gbs_kernels / peak,
        'forward': {'ms': ms_fwd, 'GBps': gbs_fwd, 'frac': gbs_fwd / peak},
        'adjoint': {'ms': ms_adj, 'GBps': gbs_adj, 'frac': gbs_adj / peak},
        'frac_of_nominal_8000': gbs_step / 8000.0,
    }

    # ---- the reference's structure on the same GPU and tensors (cuFFT + pointwise) ----
    gpu_baseline = None
    if rank == 0:
        try:
            xs = sets[0][0].detach().clone().requires_grad_(True)
            wseed, k0d, md = sets[0][1], sets[0][2], sets[0][3]

            def tf_step():
                out = torch_fft_dc(xs, k0d, md)
                torch.autograd.grad(out, xs, wseed)

            ms_tf = event_ms(tf_step, 5, warm=2)
            gpu_baseline = {'value': B / (ms_tf * 1e-3), 'unit': UNIT, 'ms_per_step': ms_tf,
                            'frac': bytes_step / (ms_tf * 1e-3) / 1e9 / peak,
                            'what': 'torch.fft.fft2 -> blend -> torch.fft.ifft2 + autograd backward '
                                    '(cuFFT + pointwise kernels, the reference\'s structure) on '
                                    'the same GPU, tensors and batch; 1 GPU'}
            del xs
        except Exception as e:
            gpu_baseline = {'error': repr(e)[:200]}
        torch.cuda.empty_cache()

    # ---- noisy-lambda variant of configs[1] (myfft.py:139): same kernels, other plan ---
    nplans = [myfft.DCPlan(k0, mask, 0.1) for (_, _, k0, mask) in sets]

    def noisy_step():
        i = it[0] = it[0] + 1
        x, w, _, _ = sets[i % nbuf]
        p = nplans[i % nbuf]
        lib.csmri_dc_forward_cartesian(x.data_ptr(), None, p.dtab.data_ptr(), p.addend.data_ptr(),
                                       outs[0].data_ptr(), B, N, N, stream)
        lib.csmri_dc_adjoint_cartesian(w.data_ptr(), p.dtab.data_ptr(), outs[1].data_ptr(),
                                       B, N, N, stream)

    ms_noisy, _ = rank_stats(event_ms(noisy_step, reps, warm=3), dev, world)
    noisy = {'noise_lvl': 0.1, 'value': world * B / (ms_noisy * 1e-3), 'unit': UNIT,
             'ms_per_step': ms_noisy, 'frac': bytes_step / (ms_noisy * 1e-3) / 1e9 / peak}
    del nplans

    # ---- e2e: public API, HOST buffers, copies inside the timed region -----------
    x0, w0, k00, m0 = sets[0]
    rows0 = (m0[:, 0, :, 0] != 0).to(torch.uint8)
    k0l = undersampling.compact_lines(k00, rows0)
    hx, hk0, hm, hw, hk0l, hrows = (t.cpu().pin_memory() for t in (x0, k00, m0, w0, k0l, rows0))
    h_out = torch.empty_like(hx).pin_memory()
    h_gx = torch.empty_like(hx).pin_memory()
    tensor_bytes = hx.numel() * 4
    e2e_steps = max(3, min(args.steps, 10))

    def timed_e2e(fn):
        for _ in range(2):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            fn()
        b.record()
        barrier()
        return rank_stats(a.elapsed_time(b) / e2e_steps, dev, world)

    pipe = hostpipe.HostDCPipeline(dev, chunk=32, depth=3)   # best of tools/gpu_e2e_probe.py
    h2d_lines = 2 * tensor_bytes + hk0l.numel() * 4 + hrows.numel()
    d2h = 2 * tensor_bytes
    ms_lines, st_lines = timed_e2e(
        lambda: pipe.forward_backward_lines(hx, hk0l, hrows, hw, h_out, h_gx))
    ms_dense, st_dense = timed_e2e(
        lambda: pipe.forward_backward(hx, hk0, hm, hw, h_out, h_gx))

    # the PCIe floor of the same step: the copies alone, both directions at once
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    d_in = [torch.empty_like(t, device=dev) for t in (hx, hk0l, hrows, hw)]

    def copies_only():
        cur = torch.cuda.current_stream()
        s_up.wait_stream(cur)
        s_dn.wait_stream(cur)
        with torch.cuda.stream(s_up):
            for d, h in zip(d_in, (hx, hk0l, hrows, hw)):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_out.copy_(outs[0], non_blocking=True)
            h_gx.copy_(outs[1], non_blocking=True)
        cur.wait_stream(s_up)
        cur.wait_stream(s_dn)

    ms_copy, _ = timed_e2e(copies_only)
    del d_in
    clocks = sampler.stop()   # sampled from the timed region to the end of the e2e legs

    e2e = {'value': world * B / (ms_lines * 1e-3), 'unit': UNIT,
           'h2d_bytes_per_step': h2d_lines, 'd2h_bytes_per_step': d2h,
           'steps': e2e_steps, 'ms_per_step': ms_lines, 'ms_per_step_ranks': st_lines,
           'h2d_GBps_per_gpu': h2d_lines / (ms_lines * 1e-3) / 1e9,
           'd2h_GBps_per_gpu': d2h / (ms_lines * 1e-3) / 1e9,
           'copies_only_ms_per_step': ms_copy,
           'frac_of_copy_floor': ms_copy / ms_lines,
           'api': 'hostpipe.HostDCPipeline.forward_backward_lines: pinned host x / grad-seed / '
                  'sampled k0 lines (B,2,L,W) / line table (B,H) uint8 in, out / grad_x back to '
                  'pinned host, 32-slice chunks on 3 streams, the whole schedule replayed as one CUDA '
                  'graph; includes the per-chunk '
                  'csmri_dc_prepare_lines and the consistency read. copies_only = the same '
                  'bytes moved with no kernel in between (the PCIe floor of this step)',
           'dense_interface': {
               'value': world * B / (ms_dense * 1e-3), 'unit': UNIT, 'ms_per_step': ms_dense,
               'ms_per_step_ranks': st_dense,
               'h2d_bytes_per_step': 4 * tensor_bytes, 'd2h_bytes_per_step': d2h,
               'api': 'HostDCPipeline.forward_backward: the reference loader\'s dense k0 and dense '
                      '2-channel mask cross PCIe too (4 tensors in); includes csmri_dc_prepare and '
                      'the row-constancy verification read'}}
    del hx, hk0, hm, hw, hk0l, hrows, h_out, h_gx, pipe
    torch.cuda.empty_cache()

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'ms_per_step_ranks': {k: v / args.steps for k, v in ms_ranks.items()},
        'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'DC operator microbench (BASELINE configs[1]): batch %d/GPU of '
                               '%dx%d fp32 slices, 4x Cartesian mask, noiseless, fwd+adjoint'
                               % (B, N, N),
                   'l2': 'inputs larger than L2: one step streams 640 MiB over 2 rotating '
                         'buffer sets (L2 = 126 MB)',
                   'sharding': 'batch-sharded, one rank per GPU, no data-path collective',
                   'timed_region': ('the K steps are 2K launches of the csmri::dc_cartesian custom '
                                    'op captured in ONE CUDA graph (replayed once untimed, then '
                                    'once timed)') if graph is not None else
                                   'K steps issued one by one from Python (--eager)',
                   'host_cores_this_rank': len(cores) if cores else None},
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': 2 * args.steps, 'clocks': clocks,
        'hbm_GBps_fwd_adj': gbs_step,
        'gpu_baseline_torch_fft': gpu_baseline,
        'noisy_lambda': noisy,
    }
    del sets, plans, outs, graph
    torch.cuda.empty_cache()
    if not args.no_recnet:
        try:
            line['recnet_train'] = recnet_train_bench(dev, rank, world)
            line['recnet_train']['cudnn_fp32'] = recnet_train_bench(dev, rank, world, mode='cudnn_fp32')
            line['recnet_train']['tf32_convs'] = recnet_train_bench(dev, rank, world, mode='cudnn_tf32')
        except Exception as e:  # secondary leg: never lose the headline line
            line.setdefault('recnet_train', {})['error'] = repr(e)[:300]
        torch.backends.cudnn.allow_tf32 = False
        torch.cuda.empty_cache()
        try:
            line['recnet_1json'] = recnet_1json_bench(dev, rank, world)
        except Exception as e:
            line['recnet_1json'] = {'error': repr(e)[:300]}
        torch.cuda.empty_cache()
    if not args.no_refinement:
        try:
            from csmri_refinement_b200 import refinement_harness
            line['refinement_train'] = refinement_harness.bench_leg(dev, rank, world)
        except Exception as e:
            line['refinement_train'] = {'error': repr(e)[:300]}
        torch.cuda.empty_cache()
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
