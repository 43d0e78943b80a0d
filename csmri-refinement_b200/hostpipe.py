"""Host-resident batches through the DC operator with copies hidden behind compute.

The reference moves a batch to the GPU, runs the model and reads results back
strictly one after the other (``BaseRunner._request_data`` ->
``utils.cudaify``, training/base_runner.py:29-41).  For the DC operator alone
that is PCIe time plus kernel time.  :class:`HostDCPipeline` cuts a pinned host
batch into chunks and runs three CUDA streams - host->device, DC forward +
adjoint, device->host - so that, after the first chunk, the step costs what the
(full-duplex) PCIe link costs and nothing else.

Two input forms:

* :meth:`forward_backward` takes what the reference's loader produces: dense
  ``k0`` and dense 2-channel ``mask`` (4 tensors host->device per step);
* :meth:`forward_backward_lines` takes the compact description of a Cartesian
  acquisition - the sampled-line table ``rows`` (B,H) uint8 and only the sampled
  lines of k0, ``k0_lines`` (B,2,L,W) - and expands it on the device
  (``csmri_dc_prepare_lines``).  At 4x acceleration that is 2.25 instead of 4
  tensors over PCIe, at 8x 2.13.

It is the same public operator underneath
(:func:`csmri_refinement_b200.myfft.dc_perform` semantics,
myfft.py:131-163 forward and :92-128 backward).  With dense inputs masks are
*assumed* row-constant while streaming and the assumption is verified for every
chunk at the end with a single device->host read - if any chunk fails it, the
step is redone through the general path.  With compact inputs the consistency
of the line table is verified the same way and a violation raises.
"""
import torch

from . import ops


class HostDCPipeline(object):
    def __init__(self, device, chunk=32, depth=3, noise_lvl=None):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.depth = int(depth)
        self.noise_lvl = float(noise_lvl) if noise_lvl else 0.0
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = {}

    def _buffers(self, tag, shapes):
        """``depth`` sets of device staging buffers, one per name in ``shapes``
        ({name: (per-slice shape, dtype)})."""
        key = (tag,) + tuple(sorted((k, tuple(v[0]), v[1]) for k, v in shapes.items()))
        bufs = self._bufs.get(tag)
        if bufs is None or bufs[0] != key:
            sets = [{k: torch.empty((self.chunk,) + tuple(shp), dtype=dt, device=self.device)
                     for k, (shp, dt) in shapes.items()} for _ in range(self.depth)]
            bufs = self._bufs[tag] = (key, sets)
        return bufs[1]

    @staticmethod
    def _check_pinned(*tensors):
        for t in tensors:
            if t.device.type != 'cpu' or not t.is_pinned():
                raise ValueError('HostDCPipeline needs pinned host tensors')

    def _run(self, B, bufs, host_in, prepare, h_out, h_gx):
        """Common three-stream loop.  ``host_in``: {name: pinned host tensor};
        ``prepare(buf, n)`` -> (dtab, addend, flag) on the compute stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        flags, done = [], [None] * self.depth
        n_chunks = (B + self.chunk - 1) // self.chunk
        for c in range(n_chunks):
            lo, hi = c * self.chunk, min(B, (c + 1) * self.chunk)
            n = hi - lo
            buf = bufs[c % self.depth]
            with torch.cuda.stream(self.s_in):
                if done[c % self.depth] is not None:      # buffer still feeding an older chunk
                    self.s_in.wait_event(done[c % self.depth])
                for k, h in host_in.items():
                    buf[k][:n].copy_(h[lo:hi], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(e_in)
                dtab, addend, flag = prepare(buf, n)
                out = ops.dc_cartesian(buf['x'][:n], None, dtab, addend)
                gx = ops.dc_cartesian(buf['g'][:n], None, dtab, None)
                flags.append(flag)
                e_run = torch.cuda.Event()
                e_run.record(self.s_run)
                done[c % self.depth] = e_run
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(e_run)
                h_out[lo:hi].copy_(out, non_blocking=True)
                h_gx[lo:hi].copy_(gx, non_blocking=True)
                out.record_stream(self.s_out)
                gx.record_stream(self.s_out)
        with torch.cuda.stream(self.s_run):
            ok = torch.stack(flags).min()
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        self.s_out.synchronize()
        return int(ok.item()) == 1

    def forward_backward(self, hx, hk0, hmask, hgrad, h_out, h_gx):
        """out = DC(x; k0, mask) and gx = (dDC/dx)^T g for pinned host tensors
        (B,2,H,W); results are written into the pinned ``h_out`` / ``h_gx``.
        Returns after all copies have completed."""
        self._check_pinned(hx, hk0, hmask, hgrad, h_out, h_gx)
        B = hx.shape[0]
        per = tuple(hx.shape[1:])
        f32 = torch.float32
        bufs = self._buffers('dense', {'x': (per, f32), 'k0': (per, f32), 'mask': (per, f32),
                                       'g': (per, f32)})

        def prepare(buf, n):
            return ops.dc_prepare(buf['k0'][:n], buf['mask'][:n], self.noise_lvl)

        ok = self._run(B, bufs, {'x': hx, 'k0': hk0, 'mask': hmask, 'g': hgrad}, prepare,
                       h_out, h_gx)
        if not ok:
            # some chunk's mask was not row-constant: redo through the general path
            self._general(hx, hk0, hmask, hgrad, h_out, h_gx)

    def forward_backward_lines(self, hx, hk0_lines, hrows, hgrad, h_out, h_gx):
        """The same step from the compact Cartesian description: ``hk0_lines``
        (B,2,L,W) pinned float32 = the sampled lines of k0 in ascending row order,
        ``hrows`` (B,H) pinned uint8 line table.  Raises if a slice does not have
        exactly L sampled rows."""
        self._check_pinned(hx, hk0_lines, hrows, hgrad, h_out, h_gx)
        B, _, H, W = hx.shape
        if hrows.shape != (B, H) or hrows.dtype != torch.uint8:
            raise ValueError('rows must be a uint8 (B,H) tensor')
        if hk0_lines.dim() != 4 or hk0_lines.shape[0] != B or hk0_lines.shape[1] != 2 or \
                hk0_lines.shape[3] != W:
            raise ValueError('k0_lines must be (B,2,L,W)')
        per = tuple(hx.shape[1:])
        f32 = torch.float32
        bufs = self._buffers('lines', {'x': (per, f32), 'g': (per, f32),
                                       'k0l': (tuple(hk0_lines.shape[1:]), f32),
                                       'rows': ((H,), torch.uint8)})

        def prepare(buf, n):
            return ops.dc_prepare_lines(buf['k0l'][:n], buf['rows'][:n], self.noise_lvl, W)

        ok = self._run(B, bufs, {'x': hx, 'k0l': hk0_lines, 'rows': hrows, 'g': hgrad}, prepare,
                       h_out, h_gx)
        if not ok:
            raise ValueError('rows / k0_lines are inconsistent: every slice must have exactly '
                             '%d sampled rows' % hk0_lines.shape[2])

    def _general(self, hx, hk0, hmask, hgrad, h_out, h_gx):
        B = hx.shape[0]
        for lo in range(0, B, self.chunk):
            hi = min(B, lo + self.chunk)
            x, k0, m, g = (t[lo:hi].to(self.device, non_blocking=True)
                           for t in (hx, hk0, hmask, hgrad))
            h_out[lo:hi].copy_(ops.dc_general(x, None, k0, m, self.noise_lvl))
            h_gx[lo:hi].copy_(ops.dc_general_adjoint(g, m, self.noise_lvl))
        torch.cuda.synchronize(self.device)
