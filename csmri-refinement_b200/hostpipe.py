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

from . import _lib, ops


class HostDCPipeline(object):
    """``chunk`` slices per pipeline step, ``depth`` chunk buffers in flight.
    Everything a chunk needs on the device - staging inputs, the plan (D table,
    addend), both results, the verification flag - is allocated once per
    (shape, input form) and the kernels are enqueued through the raw C ABI: the
    per-chunk host cost is a handful of asynchronous calls, so small chunks
    (short pipeline fill / drain) do not turn the step host-bound."""

    def __init__(self, device, chunk=32, depth=3, noise_lvl=None):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.depth = int(depth)
        self.noise_lvl = float(noise_lvl) if noise_lvl else 0.0
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = {}

    def _buffers(self, tag, shapes, n_chunks):
        """``depth`` sets of device buffers, one tensor per name in ``shapes``
        ({name: (per-slice shape, dtype)}), plus one flag per chunk."""
        key = (tag, n_chunks) + tuple(sorted((k, tuple(v[0]), v[1]) for k, v in shapes.items()))
        bufs = self._bufs.get(tag)
        if bufs is None or bufs[0] != key:
            sets = []
            for _ in range(self.depth):
                d = {k: torch.empty((self.chunk,) + tuple(shp), dtype=dt, device=self.device)
                     for k, (shp, dt) in shapes.items()}
                d['e_in'], d['e_run'], d['e_out'] = (torch.cuda.Event() for _ in range(3))
                sets.append(d)
            flags = torch.ones((n_chunks,), dtype=torch.int32, device=self.device)
            bufs = self._bufs[tag] = (key, sets, flags)
        return bufs[1], bufs[2]

    @staticmethod
    def _check_pinned(*tensors):
        for t in tensors:
            if t.device.type != 'cpu' or not t.is_pinned():
                raise ValueError('HostDCPipeline needs pinned host tensors')

    def _run(self, B, H, W, bufs, flags, host_in, prepare, h_out, h_gx):
        """Common three-stream loop.  ``host_in``: {name: pinned host tensor};
        ``prepare(buf, n, flag_ptr, stream)`` enqueues the plan kernels of a chunk."""
        lib = _lib.lib()
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        run = self.s_run.cuda_stream
        n_chunks = (B + self.chunk - 1) // self.chunk
        with torch.cuda.device(self.device):
            for c in range(n_chunks):
                lo, hi = c * self.chunk, min(B, (c + 1) * self.chunk)
                n = hi - lo
                buf = bufs[c % self.depth]
                with torch.cuda.stream(self.s_in):
                    if c >= self.depth:          # inputs of the chunk that used this set are consumed
                        self.s_in.wait_event(buf['e_run'])
                    for k, h in host_in.items():
                        buf[k][:n].copy_(h[lo:hi], non_blocking=True)
                    buf['e_in'].record(self.s_in)
                self.s_run.wait_event(buf['e_in'])
                if c >= self.depth:              # results of that chunk have left the device
                    self.s_run.wait_event(buf['e_out'])
                prepare(buf, n, flags[c:c + 1].data_ptr(), run)
                _lib.check(lib.csmri_dc_forward_cartesian(
                    buf['x'].data_ptr(), None, buf['dtab'].data_ptr(), buf['addend'].data_ptr(),
                    buf['out'].data_ptr(), n, H, W, run))
                _lib.check(lib.csmri_dc_adjoint_cartesian(
                    buf['g'].data_ptr(), buf['dtab'].data_ptr(), buf['gx'].data_ptr(), n, H, W, run))
                buf['e_run'].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(buf['e_run'])
                    h_out[lo:hi].copy_(buf['out'][:n], non_blocking=True)
                    h_gx[lo:hi].copy_(buf['gx'][:n], non_blocking=True)
                    buf['e_out'].record(self.s_out)
            with torch.cuda.stream(self.s_run):
                ok = flags[:n_chunks].min()
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        self.s_out.synchronize()
        return int(ok.item()) == 1

    def forward_backward(self, hx, hk0, hmask, hgrad, h_out, h_gx):
        """out = DC(x; k0, mask) and gx = (dDC/dx)^T g for pinned host tensors
        (B,2,H,W); results are written into the pinned ``h_out`` / ``h_gx``.
        Returns after all copies have completed."""
        self._check_pinned(hx, hk0, hmask, hgrad, h_out, h_gx)
        B, _, H, W = hx.shape
        per = tuple(hx.shape[1:])
        f32 = torch.float32
        n_chunks = (B + self.chunk - 1) // self.chunk
        bufs, flags = self._buffers('dense', {
            'x': (per, f32), 'k0': (per, f32), 'mask': (per, f32), 'g': (per, f32),
            'addend': (per, f32), 'out': (per, f32), 'gx': (per, f32), 'dtab': ((H,), f32)},
            n_chunks)
        lib = _lib.lib()

        def prepare(buf, n, flag_ptr, stream):
            _lib.check(lib.csmri_dc_prepare(
                buf['k0'].data_ptr(), buf['mask'].data_ptr(), n, H, W, self.noise_lvl,
                buf['dtab'].data_ptr(), buf['addend'].data_ptr(), flag_ptr, None, stream))

        ok = self._run(B, H, W, bufs, flags, {'x': hx, 'k0': hk0, 'mask': hmask, 'g': hgrad},
                       prepare, h_out, h_gx)
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
                hk0_lines.shape[3] != W or hk0_lines.dtype != torch.float32:
            raise ValueError('k0_lines must be a float32 (B,2,L,W) tensor')
        L = int(hk0_lines.shape[2])
        per = tuple(hx.shape[1:])
        f32 = torch.float32
        n_chunks = (B + self.chunk - 1) // self.chunk
        bufs, flags = self._buffers('lines', {
            'x': (per, f32), 'g': (per, f32), 'k0l': ((2, L, W), f32), 'rows': ((H,), torch.uint8),
            'addend': (per, f32), 'out': (per, f32), 'gx': (per, f32), 'dtab': ((H,), f32)},
            n_chunks)
        lib = _lib.lib()

        def prepare(buf, n, flag_ptr, stream):
            _lib.check(lib.csmri_dc_prepare_lines(
                buf['k0l'].data_ptr(), buf['rows'].data_ptr(), n, H, W, L, self.noise_lvl,
                buf['dtab'].data_ptr(), buf['addend'].data_ptr(), flag_ptr, stream))

        ok = self._run(B, H, W, bufs, flags, {'x': hx, 'k0l': hk0_lines, 'rows': hrows, 'g': hgrad},
                       prepare, h_out, h_gx)
        if not ok:
            raise ValueError('rows / k0_lines are inconsistent: every slice must have exactly '
                             '%d sampled rows' % L)

    def _general(self, hx, hk0, hmask, hgrad, h_out, h_gx):
        B = hx.shape[0]
        for lo in range(0, B, self.chunk):
            hi = min(B, lo + self.chunk)
            x, k0, m, g = (t[lo:hi].to(self.device, non_blocking=True)
                           for t in (hx, hk0, hmask, hgrad))
            h_out[lo:hi].copy_(ops.dc_general(x, None, k0, m, self.noise_lvl))
            h_gx[lo:hi].copy_(ops.dc_general_adjoint(g, m, self.noise_lvl))
        torch.cuda.synchronize(self.device)
