"""One process per GPU, batch sharded by rank, one flat-bucket gradient allreduce.

Replaces ``utils/custom_data_parallel.py`` (a dict-aware ``nn.DataParallel``,
:6-35) and the multi-GPU branch of ``utils/__init__.py:cudaify`` (:52-72): the
reference replicates the model, scatters the batch and reduces gradients to
GPU 0 inside ONE process every step.  Here every rank owns a contiguous,
rank-major shard of the batch from the loader on (no scatter, no gather of
outputs), the DC operator needs no communication at all, and the only
collective of a training step is one ``all_reduce(SUM)`` over a flat fp32
gradient buffer (0.13-2.3 MB for RecNet), followed by ``1/world``.

:class:`ShardedTrainer` restates ``Runner._train_step``
(training/runner.py:154-178): zero_grad, ``model(inp, kspace, mask)``, MSE on
the raw 2-channel tensors (models/criteria.py:69-83), backward, optimizer step
with ``Adam(lr, betas=(0.9, 0.999))`` (training/optimizers.py:18-21).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT (torchrun).  Returns (rank, world, device)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local_rank)
        device = torch.device('cuda', local_rank)
    else:
        device = torch.device('cpu')
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if use_cuda else 'gloo'
        kwargs = {'device_id': device} if backend == 'nccl' else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, device


def shard_range(n_total, rank, world):
    """Contiguous rank-major shard [lo, hi) of ``n_total`` items; sizes differ
    by at most one (the first ``n_total % world`` ranks get the extra item)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """Slice every tensor of a batch dict (scar_segmentation.py:212-218) to
    this rank's shard along dim 0."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


class FlatGradBucket(object):
    """All gradients of ``params`` as views into ONE flat fp32 buffer, so the
    gradient exchange is a single collective with no packing copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def zero_(self):
        self.flat.zero_()

    def check_views(self):
        """Autograd accumulates in place, so the views must survive backward."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                return False
            off += p.numel()
        return True

    def allreduce_mean(self):
        """SUM over ranks then * 1/world: the gradient of the mean loss over the
        global batch when every rank holds the same number of slices."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


class ShardedTrainer(object):
    """RecNet MSE training step on this rank's shard (training/runner.py:154-178).

    ``cuda_graph=True`` captures zero_grad + forward + loss + backward +
    allreduce + Adam into one CUDA graph replayed per step on static input
    buffers (the ~100 small kernels of a D5C5 step are launch-bound otherwise).

    Uneven shards (``shard_range`` when the global batch is not a multiple of
    the world size): every rank scales its loss by ``local_n * world / global_n``
    before backward, so SUM-allreduce / world is the gradient of the mean loss
    over the GLOBAL batch - what the reference's single-process DataParallel
    computes (utils/__init__.py:59-68 scatters one batch, the loss is taken on
    the gathered output).

    ``assume_row_constant=True`` (required under a CUDA graph, where the
    per-batch device->host read of the mask proof is illegal) is *verified*: the
    proof is still computed on the device inside the step, accumulated into a
    sticky counter and copied to pinned host memory asynchronously; ``step()``
    raises as soon as it sees the counter non-zero - at the latest in the call
    after the offending batch.
    """

    def __init__(self, model, lr=2e-4, betas=(0.9, 0.999), cuda_graph=False,
                 assume_row_constant=None):
        self.model = model
        self.bucket = FlatGradBucket(model.parameters())
        self.cuda_graph = bool(cuda_graph)
        self.assume_row_constant = assume_row_constant
        dev = self.bucket.flat.device
        self.device = dev
        self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, betas=betas,
                                          capturable=self.cuda_graph and dev.type == 'cuda')
        self.criterion = torch.nn.MSELoss()
        self._graph = None
        self._static = None
        self._loss = None
        self._loss_weight = None     # local_n * world / global_n, fixed at the first step
        self._local_n = None
        # verification of assume_row_constant=True
        self._bad = None             # device int64 counter (sticky)
        self._bad_host = None        # pinned mirror
        self._prev_event = None
        if assume_row_constant and dev.type == 'cuda':
            self._bad = torch.zeros((), dtype=torch.int64, device=dev)
            self._bad_host = torch.zeros((), dtype=torch.int64).pin_memory()

    def _weight_for(self, batch):
        n = int(batch['inp'].shape[0])
        if self._loss_weight is None or n != self._local_n:
            if self._graph is not None:
                raise ValueError('the CUDA-graph step was captured for %d slices per rank, got %d'
                                 % (self._local_n, n))
            world = self.bucket.world
            if world > 1:
                t = torch.tensor([n], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                total = int(t.item())
            else:
                total = n
            self._local_n = n
            self._loss_weight = float(n) * world / float(total)
        return self._loss_weight

    def _step_eager(self, batch):
        from . import myfft
        weight = self._loss_weight if self._loss_weight is not None else 1.0
        self.bucket.zero_()
        with myfft.assume_row_constant(self.assume_row_constant) as assumed:
            out = self.model(batch['inp'], batch['kspace'], batch['mask'])
        if self._bad is not None:
            viol = assumed.violations()
            if viol is not None:
                self._bad.add_(viol)
                self._bad_host.copy_(self._bad, non_blocking=True)
        pred = out['pred'] if isinstance(out, dict) else out
        loss = self.criterion(pred, batch['target'])
        (loss if weight == 1.0 else loss * weight).backward()
        self.bucket.allreduce_mean()
        self.optimizer.step()
        return loss.detach()

    def _capture(self, batch):
        from . import myfft
        self._static = {k: v.clone() for k, v in batch.items()}
        params = self.bucket.params
        # warm-up and capture must not advance training: snapshot weights and
        # optimizer state, restore them (in place - the graph holds the
        # addresses) once the graph exists
        saved_p = [p.detach().clone() for p in params]
        saved_s = {i: {k: v.clone() for k, v in self.optimizer.state[p].items()
                       if torch.is_tensor(v)}
                   for i, p in enumerate(params) if p in self.optimizer.state}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):           # warm-up off the capture stream
            for _ in range(3):
                self._step_eager(self._static)
        torch.cuda.current_stream().wait_stream(s)
        # the per-batch DC plan (D table, addend) must be recomputed INSIDE the
        # graph on every replay: drop the entry the warm-up left in the cache
        myfft.clear_plan_cache()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss = self._step_eager(self._static)
        myfft.clear_plan_cache()
        with torch.no_grad():
            for i, p in enumerate(params):
                p.copy_(saved_p[i])
                for k, v in self.optimizer.state[p].items():
                    if torch.is_tensor(v):
                        if i in saved_s and k in saved_s[i]:
                            v.copy_(saved_s[i][k])
                        else:
                            v.zero_()

    def _check_assumption(self):
        """Raise if an EARLIER step saw a mask that is not row-constant (the
        flag of the step just launched is looked at by the next call)."""
        if self._bad_host is None:
            return
        prev, self._prev_event = self._prev_event, torch.cuda.Event()
        self._prev_event.record()
        if prev is not None:
            prev.synchronize()         # the step before the one just launched has finished
            if int(self._bad_host.item()) != 0:
                raise RuntimeError(
                    'ShardedTrainer(assume_row_constant=True): %d batch(es) had a mask that is '
                    'not constant along W (not a Cartesian mask, compressed_sensing.py:115-116); '
                    'the Cartesian strip kernel gave wrong results for them. Use '
                    'assume_row_constant=None (checked, eager) or False (general path).'
                    % int(self._bad_host.item()))

    def step(self, batch):
        """One optimizer step; returns the (local-shard) loss as a 0-dim tensor."""
        self._weight_for(batch)
        if not self.cuda_graph:
            loss = self._step_eager(batch)
            self._check_assumption()
            return loss
        if self.assume_row_constant is None:
            raise ValueError('cuda_graph=True needs assume_row_constant=True/False: the mask '
                             'check is a device->host read, which a graph cannot contain')
        if self._graph is None:
            self._capture(batch)
            if self._bad is not None:      # the warm-up steps ran on this batch too
                self._bad_host.copy_(self._bad, non_blocking=True)
        for k, v in batch.items():
            self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        self._check_assumption()
        return self._loss
