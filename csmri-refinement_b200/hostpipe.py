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

    def __init__(self, device, chunk=32, depth=3, noise_lvl=None, use_graph=True, copy_streams=1):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.depth = int(depth)
        self.noise_lvl = float(noise_lvl) if noise_lvl else 0.0
        # copy_streams > 1: the tensors of a chunk are spread over several streams per
        # direction, so the fixed cost of each DMA (~25 us) overlaps with its neighbours'
        self.s_ins = [torch.cuda.Stream(self.device) for _ in range(max(1, int(copy_streams)))]
        self.s_outs = [torch.cuda.Stream(self.device) for _ in range(max(1, int(copy_streams)))]
        self.s_in, self.s_out = self.s_ins[0], self.s_outs[0]
        self.s_run = torch.cuda.Stream(self.device)
        self._bufs = {}
        # The whole three-stream schedule of one call (copies, kernels, events) is a
        # fixed DAG once the pinned host buffers and shapes are fixed, which is how a
        # loader's staging buffers are used: the second call with the same buffers
        # captures it into a CUDA graph, later calls replay it with ONE launch
        # (issuing ~25 asynchronous calls per chunk from Python costs ~0.5 ms per
        # chunk and made 16-slice chunks host-bound).
        self.use_graph = bool(use_graph)
        self._graphs = {}

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
                d['e_run'] = torch.cuda.Event()
                d['e_in'] = [torch.cuda.Event() for _ in self.s_ins]
                d['e_out'] = [torch.cuda.Event() for _ in self.s_outs]
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
        """One call: eager the first time a (buffers, shapes) key is seen, captured
        into a CUDA graph the second time, replayed from then on."""
        if not self.use_graph:
            ok = self._enqueue(B, H, W, bufs, flags, host_in, prepare, h_out, h_gx)
        else:
            key = (id(bufs), B, H, W) + tuple((k, t.data_ptr(), tuple(t.shape))
                                              for k, t in host_in.items()) + \
                (h_out.data_ptr(), h_gx.data_ptr())
            entry = self._graphs.get(key)
            if entry is None:
                ok = self._enqueue(B, H, W, bufs, flags, host_in, prepare, h_out, h_gx)
                self._graphs[key] = 'seen'
            else:
                if entry == 'seen':
                    if len(self._graphs) > 8:          # staging buffers changed a lot: start over
                        self._graphs = {}
                    torch.cuda.synchronize(self.device)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.device(self.device), torch.cuda.graph(graph):
                        ok_t = self._enqueue(B, H, W, bufs, flags, host_in, prepare, h_out, h_gx,
                                             capturing=True)
                    entry = self._graphs[key] = (graph, ok_t)
                graph, ok = entry
                graph.replay()
        torch.cuda.current_stream(self.device).synchronize()
        return int(ok.item()) == 1

    def _enqueue(self, B, H, W, bufs, flags, host_in, prepare, h_out, h_gx, capturing=False):
        """The three-stream schedule.  ``host_in``: {name: pinned host tensor};
        ``prepare(buf, n, flag_ptr, stream)`` enqueues the plan kernels of a chunk.
        Returns the device tensor min(flags) (1 = every chunk's plan was valid)."""
        lib = _lib.lib()
        cur = torch.cuda.current_stream(self.device)
        for s in self.s_ins + self.s_outs + [self.s_run]:
            s.wait_stream(cur)
        run = self.s_run.cuda_stream
        n_chunks = (B + self.chunk - 1) // self.chunk
        # largest tensors first, dealt round-robin over the input streams
        names = sorted(host_in, key=lambda k: -host_in[k][0:1].numel())
        with torch.cuda.device(self.device):
            for c in range(n_chunks):
                lo, hi = c * self.chunk, min(B, (c + 1) * self.chunk)
                n = hi - lo
                buf = bufs[c % self.depth]
                for i, s_in in enumerate(self.s_ins):
                    with torch.cuda.stream(s_in):
                        if c >= self.depth:      # inputs of the chunk that used this set are consumed
                            s_in.wait_event(buf['e_run'])
                        for k in names[i::len(self.s_ins)]:
                            buf[k][:n].copy_(host_in[k][lo:hi], non_blocking=True)
                        buf['e_in'][i].record(s_in)
                for e in buf['e_in']:
                    self.s_run.wait_event(e)
                if c >= self.depth:              # results of that chunk have left the device
                    for e in buf['e_out']:
                        self.s_run.wait_event(e)
                prepare(buf, n, flags[c:c + 1].data_ptr(), run)
                _lib.check(lib.csmri_dc_forward_cartesian(
                    buf['x'].data_ptr(), None, buf['dtab'].data_ptr(), buf['addend'].data_ptr(),
                    buf['out'].data_ptr(), n, H, W, run))
                _lib.check(lib.csmri_dc_adjoint_cartesian(
                    buf['g'].data_ptr(), buf['dtab'].data_ptr(), buf['gx'].data_ptr(), n, H, W, run))
                buf['e_run'].record(self.s_run)
                outs = ((h_out, 'out'), (h_gx, 'gx'))
                for i, s_out in enumerate(self.s_outs):
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(buf['e_run'])
                        for h, k in outs[i::len(self.s_outs)]:
                            h[lo:hi].copy_(buf[k][:n], non_blocking=True)
                        buf['e_out'][i].record(s_out)
            with torch.cuda.stream(self.s_run):
                ok = flags[:n_chunks].min()
        for s in self.s_ins + self.s_outs + [self.s_run]:
            cur.wait_stream(s)
        return ok

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
