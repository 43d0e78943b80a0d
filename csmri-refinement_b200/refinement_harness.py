"""``configs/2-refinement.json`` (BASELINE configs[4]) on the B200 path: frozen
RecNet + learnable U-Net refinement, PatchGAN discriminator, VGG19 perceptual
loss, alternating adversarial step, one process per GPU.

What the reference does: ``adversarial_runner.build_runner``
(training/adversarial_runner.py:22-132) builds ``RefinementWrapper``
(models/refinement_wrapper.py:95-220) and ``CNNDiscriminator``, then every step
``_train_single_step`` (:322-389) runs

    out_gen = gen(inp, kspace, mask)            # frozen RecNet (3 DC layers, no grad)
                                                # -> U-Net -> real-penalty-add
    D(|fake|.detach() via ImagePool), D(|target|) -> discriminator loss (GAN, label smoothing)
    D(|fake|) with grad -> generator losses: 0.5*gan + 1*FeatureMatching + 10*VGG19
                                             + 2*FeaturePenalty(L1 on prescaled_refinement)
    D update, then G update (Adam 2e-4, beta1 0.5 each).

Here the hot path of that step - the frozen RecNet's DC layers, its thin
convolutions, the min-max / real-penalty-add recombination and the magnitude
images - run through ``libcsmri_dc.so``; U-Net, discriminator and VGG19 stay
cuDNN through torch (not on the DC path).  Multi-GPU replaces the reference's
single-process ``nn.DataParallel`` (utils/__init__.py:52-72): every rank owns
``batch_size`` slices, BatchNorm statistics and the image pool stay rank-local
(DataParallel semantics), and the two gradient exchanges are flat buckets -
3.7 MB for the generator, **112 MB for the discriminator**, the one
bandwidth-bound collective of the whole task, which is cut into chunks that are
all-reduced while backward is still producing the earlier layers' gradients.

Deviations forced by the inputs that do not exist, all stated in
the bench line: the JSON's ``discriminator_model`` has no ``name`` (defaults to
CNNDiscriminator), ``pretrained_weights`` is a placeholder path (RecNet keeps its
seeded init), VGG19 cannot be downloaded (seeded random weights), data is
synthetic.  One deliberate difference: the reference back-propagates the
generator loss *after* the in-place discriminator update (legal in torch 0.3,
an error in torch >= 1.5); here both gradients are taken at the pre-update
discriminator weights - the standard simultaneous GAN step.
"""
import random

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import harness, recnet, refinement_models, refinement_ops
from .config import Configuration


# ---------------------------------------------------------------------------
# model builders (training/adversarial_runner.py:22-38, models/refinement_wrapper.py:27-45)
# ---------------------------------------------------------------------------
def build_learnable_model(conf):
    gen = Configuration.from_dict(conf.generator_model, conf)
    sub = Configuration.from_dict(gen.learnable_model, conf)
    if sub.name != 'UNET':
        raise ValueError('expected a UNET learnable model, got %r' % (sub.name,))
    return refinement_models.construct_unet(sub)


def build_discriminator(conf):
    sub = Configuration.from_dict(conf.discriminator_model, conf)
    name = sub.get_attr('name', default='CNNDiscriminator')   # absent in the shipped JSON
    if name != 'CNNDiscriminator':
        raise ValueError('Unknown discriminator %r' % (name,))
    return refinement_models.construct_discriminator(sub)


class RefinementWrapper(nn.Module):
    """models/refinement_wrapper.py:95-220 for the frozen-pretrained case: a
    RecNet whose output is detached, a learnable model on top, and the
    ``real-penalty-add`` / ``add`` recombination.  ``forward(inp, kspace, mask)``
    keeps the argument names the runner binds by (training/base_runner.py:43-63).
    state_dict keys: ``scale``, ``pretrained_model.*``, ``learnable_model.*``."""

    def __init__(self, pretrained_model, learnable_model, mode='add', input_mode='input',
                 freeze_pretrained_model=True):
        super(RefinementWrapper, self).__init__()
        if not freeze_pretrained_model:
            raise NotImplementedError('only the frozen pretrained path of the shipped config')
        if mode not in ('add', 'real-penalty-add'):
            raise ValueError('Unknown mode {}'.format(mode))
        if input_mode not in ('input', 'output', 'concat'):
            raise ValueError('Unknown input mode {}'.format(input_mode))
        self.mode, self.input_mode = mode, input_mode
        self.freeze_pretrained_model = True
        self.pretrained_model = pretrained_model
        self.learnable_model = learnable_model
        if mode == 'real-penalty-add':
            self.scale = nn.Parameter(torch.zeros(1))
        for p in self.pretrained_model.parameters():
            p.requires_grad = False

    def trainable_parameters(self):
        """What the reference's overridden ``parameters()`` yields (:146-162)."""
        return [p for p in self.parameters() if p.requires_grad]

    def _learnable_input(self, inp, out_pretrained):
        if self.input_mode == 'input':
            return inp
        if self.input_mode == 'output':
            return out_pretrained
        return torch.cat((inp, out_pretrained), dim=1)

    def forward(self, inp, kspace, mask):
        with torch.no_grad():                      # _var_without_grad + .detach() (:208-220)
            # the RecNet path is parity-gated in fp32 (1e-5): no TF32 in its convolutions
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False,
                                            benchmark=torch.backends.cudnn.benchmark,
                                            deterministic=torch.backends.cudnn.deterministic):
                out_pretrained = self.pretrained_model(inp, kspace, mask)
        out_learnable = self.learnable_model(self._learnable_input(inp, out_pretrained))
        if self.mode == 'add':
            return out_pretrained + out_learnable
        if out_pretrained.is_cuda:
            # min-max scale, scale * refinement, unscale, concat: one fused kernel
            return refinement_ops.refinement_real_penalty_add(out_pretrained, out_learnable,
                                                              self.scale)
        return real_penalty_add_reference(out_pretrained, out_learnable, self.scale)


def real_penalty_add_reference(out_pretrained, out_learnable, scale):
    """models/refinement_wrapper.py:51-92,173-197 as plain tensor ops - the CPU
    statement the fused kernel is tested against (refinement_ops.npz); used by
    the CPU tests of this module only."""
    re, im = out_pretrained[:, 0:1].contiguous(), out_pretrained[:, 1:2].contiguous()
    b, c, h, w = re.shape
    flat = re.view(b, c, h * w)
    minimum = flat.min(dim=2, keepdim=True)[0]
    flat = flat - minimum
    maximum = flat.max(dim=2, keepdim=True)[0]
    scaled = (flat / maximum * 2 - 1).view(b, c, h, w)
    scaled_ref = scale * out_learnable
    refined = (scaled + scaled_ref).view(b, c, h * w)
    out_real = ((refined + 1) / 2 * maximum + minimum).view(b, c, h, w)
    return {'pred': torch.cat((out_real, im), dim=1), 'pretrained': out_pretrained,
            'prescaled_refinement': out_learnable, 'scaled_refinement': scaled_ref}


def build_generator(conf, dc_factory=None):
    """RefinementWrapper from ``conf.generator_model``; construction order =
    RNG order of the reference (pretrained model, then learnable model).
    ``dc_factory`` is the RecNet mirror's test-only injection point."""
    gen = Configuration.from_dict(conf.generator_model, conf)
    if gen.name != 'RefinementWrapper':
        raise ValueError('expected a RefinementWrapper generator, got %r' % (gen.name,))
    pre_conf = Configuration.from_dict(gen.pretrained_model, conf)
    kwargs = {} if dc_factory is None else {'dc_factory': dc_factory}
    pretrained = recnet.construct_model(pre_conf, pre_conf.name, **kwargs)
    learnable = build_learnable_model(conf)
    return RefinementWrapper(pretrained, learnable, mode=gen.get_attr('mode', default='add'),
                             input_mode=gen.get_attr('input_mode', default='input'),
                             freeze_pretrained_model=gen.get_attr('freeze_pretrained_model',
                                                                  default=True))


# ---------------------------------------------------------------------------
# losses (models/adversarial_loss.py, models/vgg_loss.py, models/criteria.py)
# ---------------------------------------------------------------------------
def _bce_to(prob, label):
    return F.binary_cross_entropy(prob, torch.full_like(prob, label))


def gan_loss_disc(out_fake, out_real, label_smoothing=0.0):
    """GANLoss('disc') (:103-115,73-85): BCE(fake, 0) + BCE(real, 1 - smoothing)."""
    return _bce_to(out_fake['prob'], 0.0) + _bce_to(out_real['prob'], 1.0 - label_smoothing)


def gan_loss_gen(out_fake):
    """GANLoss('gen') (:87-92): BCE(fake, 1)."""
    return _bce_to(out_fake['prob'], 1.0)


def feature_matching_loss(out_fake, out_real, distance=F.l1_loss):
    """FeatureMatchingLoss('gen') (:141-160): mean over the discriminator's
    feature maps of distance(fake, real.detach())."""
    feats = list(zip(out_fake['features'], out_real['features']))
    return sum(distance(f, r.detach()) for f, r in feats) / len(feats)


def _magnitude(x):
    return refinement_ops.complex_abs(x) if x.is_cuda else \
        ((x[:, 0] ** 2 + x[:, 1] ** 2) ** 0.5).unsqueeze(1)


def vgg_loss(vgg, prediction, target, criterion=F.mse_loss):
    """VGGLoss.forward for complex inputs (models/vgg_loss.py:44-65): magnitude,
    replicated to 3 channels, distance between the feature blocks."""
    p = _magnitude(prediction)
    t = _magnitude(target.detach())
    fp = vgg(torch.cat((p, p, p), dim=1))
    with torch.no_grad():
        ft = vgg(torch.cat((t, t, t), dim=1))
    return sum(criterion(a, b) for a, b in zip(fp, ft))


class ImagePool(object):
    """utils/image_pool.py:7-60: history of generated images shown to the
    discriminator; per rank, Python ``random`` (seed it per rank)."""

    def __init__(self, pool_size, p_pool_image=0.5):
        self.pool_size, self.p_pool_image = pool_size, p_pool_image
        self.images = []

    def query(self, image_batch):
        if self.pool_size == 0:
            return image_batch
        result = []
        for image in image_batch.detach():
            image = image.unsqueeze(0)
            if len(self.images) < self.pool_size:
                self.images.append(image)
                result.append(image)
            elif random.uniform(0, 1) < self.p_pool_image:
                idx = random.randint(0, self.pool_size - 1)
                result.append(self.images[idx].clone())
                self.images[idx] = image
            else:
                result.append(image)
        return torch.cat(result, 0)


# ---------------------------------------------------------------------------
# gradient exchange: flat bucket, chunks all-reduced while backward still runs
# ---------------------------------------------------------------------------
class OverlappedGradBucket(object):
    """All gradients of ``params`` as views into one flat buffer laid out in
    REVERSE parameter order - the order backward finishes them - and cut into
    chunks of ~``chunk_bytes``.  A post-accumulate hook on every parameter
    counts its chunk down; a complete chunk is all-reduced at once
    (``async_op=True``: NCCL's stream, concurrent with the rest of backward).
    ``finish()`` reduces what is left, waits, and scales by 1/world."""

    def __init__(self, params, chunk_bytes=16 << 20, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = bool(overlap) and self.world > 1
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self.chunks = []                 # [lo, hi, n_params]
        self._chunk_of = {}
        off, lo, count = 0, 0, 0
        for p in reversed(self.params):
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._chunk_of[p] = len(self.chunks)
            off += n
            count += 1
            if (off - lo) * self.flat.element_size() >= chunk_bytes:
                self.chunks.append([lo, off, count])
                lo, count = off, 0
        if count:
            self.chunks.append([lo, off, count])
        self._pending = [c[2] for c in self.chunks]
        self._launched = [False] * len(self.chunks)
        self._work = []
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def begin(self):
        """Before backward: zero the gradients, re-arm the chunk counters."""
        self.flat.zero_()
        self._pending = [c[2] for c in self.chunks]
        self._launched = [False] * len(self.chunks)
        self._work = []

    def _launch(self, k):
        lo, hi, _ = self.chunks[k]
        self._launched[k] = True
        self._work.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, p):
        k = self._chunk_of[p]
        self._pending[k] -= 1
        if self._pending[k] == 0 and not self._launched[k]:
            self._launch(k)

    def finish(self):
        """After backward: SUM over ranks / world in ``flat`` (every ``p.grad``)."""
        if self.world == 1:
            return
        if self.overlap:
            for k in range(len(self.chunks)):
                if not self._launched[k]:       # parameters that received no gradient
                    self._launch(k)
            for w in self._work:
                w.wait()
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / self.world)


# ---------------------------------------------------------------------------
# the adversarial step
# ---------------------------------------------------------------------------
class AdversarialTrainer(object):
    """``AdversarialRunner._train_single_step`` (training/adversarial_runner.py:
    322-389) for the loss set of the shipped config, on this rank's batch."""

    def __init__(self, conf, device, rank=0, overlap=True, vgg_seed=1234, dc_factory=None,
                 chunk_bytes=16 << 20, channels_last=False):
        self.conf, self.device = conf, device
        harness.set_random_seeds(conf.seed)            # identical replicas on every rank
        self.gen = build_generator(conf, dc_factory).to(device)
        self.disc = build_discriminator(conf).to(device)
        self.gen.pretrained_model.eval()               # frozen path: no dropout / BN in RecNet anyway
        g = torch.Generator().manual_seed(vgg_seed)
        self.vgg = refinement_models.VGG19()
        for p in self.vgg.parameters():                # ImageNet weights cannot be downloaded here
            if p.dim() > 1:
                p.data.normal_(0.0, (2.0 / (p.shape[1] * 9)) ** 0.5, generator=g)
            else:
                p.data.zero_()
        self.vgg = self.vgg.to(device).eval()
        if channels_last:
            # library-side layout choice for the cuDNN models only (the RecNet path and every
            # kernel of libcsmri_dc stay NCHW planar): saves cuDNN's per-layer NCHW<->NHWC
            # transposes around its tensor-core kernels
            for m in (self.gen.learnable_model, self.disc, self.vgg):
                m.to(memory_format=torch.channels_last)
        random.seed(conf.seed * 1000 + rank)           # image pool draws: per rank
        dconf = Configuration.from_dict(conf.discriminator_model, conf)
        pool = dconf.get_attr('image_pool_size', default=5 * conf.batch_size) \
            if dconf.get_attr('use_image_pool', default=False) else 0
        self.pool = ImagePool(pool, dconf.get_attr('image_pool_sample_prob', default=0.5))
        if dconf.get_attr('input_method', default='simple') != 'simple-magnitude':
            raise NotImplementedError('discriminator input_method of the shipped config only')
        self.label_smoothing = conf.get_attr('discriminator_label_smoothing', default=0.0)
        adv, other = conf.generator_adversarial_losses, conf.generator_losses
        if list(adv) != ['gan', 'FeatureMatching'] or list(other) != ['VGG19', 'FeaturePenalty'] \
                or list(conf.discriminator_losses) != ['gan']:
            raise NotImplementedError('loss set of the shipped config only')
        w = conf.get_attr('generator_loss_weights', default={})
        # order matters: adversarial criteria first (training/base_runner.py:19-27)
        self.gen_weights = [w.get(n, 1.0) for n in list(adv) + list(other)]
        fp = conf.feature_penalty
        self.penalty_key = fp['input_key']
        self.penalty = {'L1': F.l1_loss, 'MSE': F.mse_loss}[fp.get('criterion', 'MSE')]
        self.gen_bucket = OverlappedGradBucket(self.gen.trainable_parameters(), chunk_bytes,
                                               overlap)
        self.disc_bucket = OverlappedGradBucket(self.disc.parameters(), chunk_bytes, overlap)
        self.gen_opt = torch.optim.Adam(self.gen_bucket.params,
                                        **harness.adam_args(conf.generator_optimizer, conf))
        self.disc_opt = torch.optim.Adam(self.disc_bucket.params,
                                         **harness.adam_args(conf.discriminator_optimizer, conf))

    def step(self, batch):
        """One D update + one G update; returns {'disc_loss', 'gen_loss', ...} (0-dim tensors)."""
        out_gen = self.gen(batch['inp'], batch['kspace'], batch['mask'])
        fake_mag = _magnitude(out_gen['pred'])
        # discriminator: pooled, detached fakes vs real targets
        out_fake_d = self.disc(self.pool.query(fake_mag.detach()))
        out_real = self.disc(_magnitude(batch['target']).detach())
        disc_loss = gan_loss_disc(out_fake_d, out_real, self.label_smoothing)
        # generator: through the discriminator with gradients
        out_fake_g = self.disc(fake_mag)
        losses = [gan_loss_gen(out_fake_g), feature_matching_loss(out_fake_g, out_real),
                  vgg_loss(self.vgg, out_gen['pred'], batch['target']),
                  self.penalty(out_gen[self.penalty_key],
                               torch.zeros_like(out_gen[self.penalty_key]))]
        gen_loss = sum(wt * l for wt, l in zip(self.gen_weights, losses))
        # both gradients at the pre-update weights, each confined to its own parameters
        self.disc_bucket.begin()
        disc_loss.backward(inputs=self.disc_bucket.params)
        self.gen_bucket.begin()
        gen_loss.backward(inputs=self.gen_bucket.params)
        self.disc_bucket.finish()
        self.gen_bucket.finish()
        self.disc_opt.step()
        self.gen_opt.step()
        return {'disc_loss': disc_loss.detach(), 'gen_loss': gen_loss.detach(),
                'gen_losses': [l.detach() for l in losses]}


def bench_leg(dev, rank, world, steps=4, warmup=2):
    """BASELINE configs[4]: one adversarial training step of 2-refinement.json,
    ``batch_size`` (5) slices of 512x512 per GPU (weak scaling)."""
    conf = harness.load_config(harness.config_path('2-refinement.json'))
    b = int(conf.batch_size)
    # U-Net / discriminator / VGG19 are not parity-gated: torch's stock convolution
    # setting (TF32 tensor cores); the frozen RecNet path stays fp32 (RefinementWrapper.forward)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    batches = [harness.synthetic_batch(conf, b, dev, seed=2000 + 16 * rank + i) for i in range(2)]
    res = {}
    for mode in (('overlapped', True),) + ((('serial', False),) if world > 1 else ()):
        trainer = AdversarialTrainer(conf, dev, rank, overlap=mode[1], channels_last=True)
        for i in range(warmup):
            trainer.step(batches[i % 2])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            out = trainer.step(batches[i % 2])
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res[mode[0]] = ms
        last = (trainer, out)
    trainer, out = last
    ar_ms = None
    if world > 1:                                  # the 112 MB exchange on its own
        buf = trainer.disc_bucket.flat
        for _ in range(2):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.all_reduce(buf)
        e.record()
        torch.cuda.synchronize()
        ar_ms = a.elapsed_time(e) / 5
    ms = res['overlapped']
    return {
        'metric': 'refinement_train_slices_per_s', 'value': world * b / (ms * 1e-3),
        'unit': 'slices/s', 'ms_per_step': ms, 'steps': steps, 'scaling': 'weak',
        'ms_per_step_serial_allreduce': res.get('serial'),
        'disc_allreduce_ms_alone': ar_ms,
        'disc_bucket_bytes': trainer.disc_bucket.nbytes(),
        'disc_bucket_chunks': len(trainer.disc_bucket.chunks),
        'gen_bucket_bytes': trainer.gen_bucket.nbytes(),
        'gen_loss': float(out['gen_loss'].item()), 'disc_loss': float(out['disc_loss'].item()),
        'config': 'configs/2-refinement.json unchanged: RefinementWrapper(real-penalty-add) = frozen '
                  'RecNet D3C3 (3 DC layers, forward only) + U-Net (920,033 params), PatchGAN '
                  'discriminator (27,941,697 params), losses 0.5*gan + FeatureMatching + 10*VGG19 + '
                  '2*FeaturePenalty, Adam(2e-4, beta1 0.5) x2, batch %d/GPU of 512x512, 8x Cartesian; '
                  'synthetic data, discriminator name defaulted, RecNet and VGG19 weights are '
                  'seeded random (checkpoint / ImageNet weights do not exist offline: SURVEY D6); U-Net / '
                  'discriminator / VGG19 convolutions with torch\'s stock TF32 setting, RecNet path fp32; '
                  'discriminator gradients all-reduced in %d chunks overlapped with backward'
                  % (b, len(trainer.disc_bucket.chunks))}
