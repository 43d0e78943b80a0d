"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``

Everything stored here is an output of unmodified reference code imported from
``/root/reference``:

* ``cs.cartesian_mask``      compressed_sensing.py:82-123
* ``cs.undersample``         compressed_sensing.py:460-512
* ``cs.data_consistency``    compressed_sensing.py:515-529
* ``myfft.data_consistency`` myfft.py:131-142 (pure blend; ``pytorch_fft`` is
  stubbed only so that the module imports - the stub is never called)
* ``Undersample.__call__``   myImageTransformations.py:1196-1238
* ``models.recnet.RecNet``   models/recnet.py:65-161 with ``dc_layers`` (a plain
  list, recnet.py:128-134) swapped for an op that chains the reference blend
  with ``torch.fft`` (the reference's own FFT wrapper is CUDA-only).

* ``_scale`` / ``_unscale`` / ``_refinement_real_penalty_add``
  models/refinement_wrapper.py:51-92,173-197 and ``magnitude_image``
  utils/tensor_transforms.py:78-99
* ``configs/1-recnet.json`` and ``configs/2-refinement.json`` copied byte for
  byte (they are the INPUTS north_star says must run unchanged; sha256 recorded)
  and what ``Configuration.from_json`` + ``construct_model`` build from the
  first one under its own seed (keys, shapes, initial weights)

The fixtures are small (< 1 MB in total) and are what pins ``oracle/``.
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    pkg = types.ModuleType('pytorch_fft')
    sub = types.ModuleType('pytorch_fft.fft')
    for name in ('fft,ifft,fft2,ifft2,fft3,ifft3,rfft,irfft,rfft2,irfft2,'
                 'rfft3,irfft3').split(','):
        setattr(sub, name, None)
    pkg.fft = sub
    sys.modules['pytorch_fft'] = pkg
    sys.modules['pytorch_fft.fft'] = sub
    import data.reconstruction.deep_med_lib.utils.compressed_sensing as cs
    import data.reconstruction.deep_med_lib.my_pytorch.myfft as myfft
    import data.reconstruction.deep_med_lib.my_pytorch.myImageTransformations as mit
    import models.recnet as recnet
    return cs, myfft, mit, recnet


def main():
    cs, myfft, mit, recnet = import_reference()

    # ---- 1. masks: sampled rows for every sweep size / acceleration ----------
    masks = {}
    for n in (32, 64, 128, 256, 320, 512):
        for acc in (4, 8, 12):
            if n // acc <= 8:
                continue
            m = cs.cartesian_mask((3, n, n), acc, 8, centred=False,
                                  rng=np.random.RandomState(0))
            assert (m == m[:, :, :1]).all()
            masks['rows_n%d_acc%d' % (n, acc)] = m[:, :, 0].astype(np.uint8)
    masks['full_n32_acc4'] = cs.cartesian_mask(
        (2, 32, 32), 4, 8, centred=False, rng=np.random.RandomState(0))
    masks['full_n32_acc4_centred'] = cs.cartesian_mask(
        (2, 32, 32), 4, 8, centred=True, rng=np.random.RandomState(0))
    # the default sample_n=10 and a non-integer acceleration
    masks['full_n64_acc3p5_s10'] = cs.cartesian_mask(
        (2, 64, 64), 3.5, centred=False, rng=np.random.RandomState(7))
    np.savez_compressed(os.path.join(HERE, 'masks.npz'), **masks)

    # ---- 2. undersample + numpy DC ------------------------------------------
    rs = np.random.RandomState(1)
    n = 32
    img = rs.uniform(0, 1, (2, n, n))
    mask = cs.cartesian_mask((2, n, n), 4, 8, centred=False,
                             rng=np.random.RandomState(0))
    rng = np.random.RandomState(5)
    x_u, x_fu = cs.undersample(img, mask, centred=False, norm='ortho', rng=rng)
    after = rng.normal()                    # pins how much RNG was consumed
    x_u_nz, x_fu_nz = cs.undersample(img, mask, centred=False, norm='ortho',
                                     noise=0.01, rng=np.random.RandomState(5))
    xin = rs.normal(size=(2, n, n)) + 1j * rs.normal(size=(2, n, n))
    xd = cs.data_consistency(xin, x_fu, mask, centered=False, norm='ortho')
    # general (non row-constant) mask too
    gmask = (rs.uniform(size=(2, n, n)) < 0.3).astype(np.float64)
    gk0 = gmask * np.fft.fft2(img, norm='ortho')
    xd_g = cs.data_consistency(xin, gk0, gmask, centered=False, norm='ortho')
    np.savez_compressed(
        os.path.join(HERE, 'undersample_dc.npz'), img=img, mask=mask,
        x_u=x_u, x_fu=x_fu, rng_after=after, x_u_nz=x_u_nz, x_fu_nz=x_fu_nz,
        xin=xin, xd=xd, gmask=gmask, gk0=gk0, xd_g=xd_g)

    # ---- 3. reference torch blend (fp32) -------------------------------------
    g = torch.Generator().manual_seed(3)
    k = torch.randn(2, 2, n, n, generator=g)
    k0 = torch.randn(2, 2, n, n, generator=g)
    m = torch.from_numpy(np.stack([mask, mask], 1).astype(np.float32))
    k0 = k0 * m
    blend = {
        'k': k.numpy(), 'k0': k0.numpy(), 'mask': m.numpy(),
        'out_none': myfft.data_consistency(k, k0, m, None).numpy(),
        'out_zero': myfft.data_consistency(k, k0, m, 0).numpy(),
        'out_0p1': myfft.data_consistency(k, k0, m, 0.1).numpy(),
        'out_2p5': myfft.data_consistency(k, k0, m, 2.5).numpy(),
    }
    np.savez_compressed(os.path.join(HERE, 'blend.npz'), **blend)

    # ---- 4. Undersample transform group (fixed masks, RandomState(0)) --------
    tr = mit.Undersample('varden', (1, n, n), acceleration_rate=4,
                         fixed_mask=True, num_fixed_masks=2)
    im1 = rs.uniform(0, 1, (n, n, 1))
    im2 = rs.uniform(0, 1, (n, n, 1))
    grp1 = tr(im1.copy())
    grp2 = tr(im2.copy())
    np.savez_compressed(os.path.join(HERE, 'undersample_group.npz'),
                        im1=im1, im2=im2, grp1=grp1, grp2=grp2)

    # ---- 5. reference RecNet, DC = reference blend + torch.fft ---------------
    class RefBlendDC(object):
        def __init__(self, noise_lvl=None):
            self.noise_lvl = noise_lvl

        def perform(self, x, k0, mask):
            kc = torch.fft.fft2(torch.complex(x[:, 0], x[:, 1]), norm='ortho')
            kk = torch.stack([kc.real, kc.imag], 1)
            out = myfft.data_consistency(kk, k0, mask, self.noise_lvl)
            oc = torch.fft.ifft2(torch.complex(out[:, 0], out[:, 1]),
                                 norm='ortho')
            return torch.stack([oc.real, oc.imag], 1)

    torch.manual_seed(0)
    net = recnet.RecNet(num_blocks=2, num_convs=3, num_filters=4)
    from models.weight_inits import initialize_weights
    for blk in net.conv_blocks:
        initialize_weights(blk, {})
    net.dc_layers = [RefBlendDC() for _ in net.dc_layers]
    nn_ = 32
    img = torch.from_numpy(rs.uniform(0, 1, (2, nn_, nn_)))
    mk = cs.cartesian_mask((2, nn_, nn_), 4, 4, centred=False,
                           rng=np.random.RandomState(0))
    xu, xfu = cs.undersample(img.numpy(), mk, centred=False, norm='ortho',
                             rng=np.random.RandomState(0))
    inp = torch.from_numpy(np.stack([xu.real, xu.imag], 1).astype(np.float32))
    ksp = torch.from_numpy(np.stack([xfu.real, xfu.imag], 1).astype(np.float32))
    msk = torch.from_numpy(np.stack([mk, mk], 1).astype(np.float32))
    tgt = torch.stack([img.float(), torch.zeros_like(img).float()], 1)
    out = net(inp, ksp, msk)
    loss = torch.nn.functional.mse_loss(out, tgt)
    loss.backward()
    fix = {'inp': inp.numpy(), 'kspace': ksp.numpy(), 'mask': msk.numpy(),
           'target': tgt.numpy(), 'out': out.detach().numpy(),
           'loss': np.float64(loss.item())}
    for name, p in net.state_dict().items():
        fix['w:' + name] = p.numpy()
    for name, p in net.named_parameters():
        fix['g:' + name] = p.grad.numpy()
    # residual ("use_refinement") variant and the dict-returning variant
    torch.manual_seed(0)
    net2 = recnet.RecNet(num_blocks=2, num_convs=3, num_filters=4,
                         use_refinement=True, return_intermediate_recs=True,
                         skip_final_dc=True)
    net2.load_state_dict(net.state_dict())
    net2.dc_layers = [RefBlendDC() for _ in net2.dc_layers]
    o2 = net2(inp, ksp, msk)
    fix['out_refine_pred'] = o2['pred'].detach().numpy()
    fix['out_refine_rec0'] = o2['reconstructions'][0].detach().numpy()
    fix['n_refine_recs'] = np.int64(len(o2['reconstructions']))
    # full DC with the noisy branch
    dcn = RefBlendDC(0.1).perform(inp * 0.5 + 0.1, ksp, msk)
    fix['dc_noisy_0p1'] = dcn.numpy()
    np.savez_compressed(os.path.join(HERE, 'recnet_tiny.npz'), **fix)

    # ---- 6. loader tail and reporting transform ---------------------------------
    from utils.tensor_transforms import complex_abs
    import data.reconstruction.rec_transforms as rt
    im64 = rs.uniform(0, 1, (64, 64, 1))
    tail = {'im64': im64,
            'crop_same': mit.CenterCropInKspace(64)(im64.copy()),
            'crop_32': mit.CenterCropInKspace(32)(im64.copy()),
            'crop_pad96': mit.CenterCropInKspace(96)(im64.copy())}
    pt = torch.from_numpy((rs.normal(size=(2, 2, 32, 32)) * 0.7).astype(np.float32))
    tt = torch.from_numpy((rs.normal(size=(2, 2, 32, 32)) * 0.7).astype(np.float32))
    op, ot = rt.output_transform()(pt, tt)
    tail.update(pred=pt.numpy(), target=tt.numpy(), abs_pred=complex_abs(pt).numpy(),
                out_pred=op.numpy(), out_target=ot.numpy(),
                psnr=np.float64(10. * np.log10(1. / torch.nn.functional.mse_loss(op, ot).item())))
    np.savez_compressed(os.path.join(HERE, 'loader_tail.npz'), **tail)

    # ---- 7. refinement-path pointwise ops (models/refinement_wrapper.py:51-92,173-197,
    #         utils/tensor_transforms.py:78-99) ------------------------------------
    import models.refinement_wrapper as rw
    from utils.tensor_transforms import magnitude_image
    xs = torch.from_numpy((rs.normal(size=(3, 2, 32, 48)) * 1.3 + 0.2).astype(np.float32))
    scaled, mn, mx = rw._scale(xs)
    ref = {'x': xs.numpy(), 'scaled': scaled.numpy(), 'minimum': mn.numpy(),
           'maximum': mx.numpy(), 'unscaled': rw._unscale(scaled * 1.5, mn, mx).numpy(),
           'magnitude_image': magnitude_image(xs).numpy()}
    pre = torch.from_numpy(rs.normal(size=(3, 2, 32, 48)).astype(np.float32))
    learn = torch.from_numpy((rs.normal(size=(3, 1, 32, 48)) * 0.3).astype(np.float32))
    learn.requires_grad_(True)
    scale_p = torch.nn.Parameter(torch.tensor([0.37]))
    fake_self = types.SimpleNamespace(scale=scale_p, learnable_model=lambda inp: learn,
                                      _learnable_model_input_fn=lambda inp, out: inp)
    res = rw.RefinementWrapper._refinement_real_penalty_add(fake_self, None, pre)
    cot = torch.from_numpy(rs.normal(size=(3, 2, 32, 48)).astype(np.float32))
    (res['pred'] * cot).sum().backward()
    ref.update(pre=pre.numpy(), learn=learn.detach().numpy(), scale=scale_p.detach().numpy(),
               pred=res['pred'].detach().numpy(), cot=cot.numpy(),
               grad_learn=learn.grad.numpy(), grad_scale=scale_p.grad.numpy())
    np.savez_compressed(os.path.join(HERE, 'refinement_ops.npz'), **ref)

    # ---- 8. the shipped run configurations, byte for byte, and what the reference
    #         builds from configs/1-recnet.json (utils/config.py:212-250,
    #         training/runner.py:18-21, models/recnet.py:20-26) ---------------------
    import hashlib
    import shutil
    from utils.config import Configuration
    from models import construct_model
    cdir = os.path.join(HERE, 'configs')
    os.makedirs(cdir, exist_ok=True)
    conf_fix = {}
    for name in ('1-recnet.json', '2-refinement.json'):
        shutil.copyfile(os.path.join(REF, 'configs', name), os.path.join(cdir, name))
        with open(os.path.join(cdir, name), 'rb') as f:
            conf_fix['sha256:' + name] = np.array(hashlib.sha256(f.read()).hexdigest())
    conf = Configuration.from_json(os.path.join(REF, 'configs', '1-recnet.json'))
    model_conf = Configuration.from_dict(conf.model, conf)
    torch.manual_seed(conf.seed)
    net1 = construct_model(model_conf, model_conf.name)
    conf_fix['recnet1:seed'] = np.int64(conf.seed)
    conf_fix['recnet1:num_params'] = np.int64(sum(p.numel() for p in net1.parameters()))
    conf_fix['recnet1:num_dc'] = np.int64(len(net1.dc_layers))
    conf_fix['recnet1:keys'] = np.array(list(net1.state_dict().keys()))
    for k, v in net1.state_dict().items():
        conf_fix['recnet1:w:' + k] = v.numpy()
    conf_fix['recnet1:batch_size'] = np.int64(conf.batch_size)
    conf_fix['recnet1:acc'] = np.int64(conf.undersampling['acceleration_factor'])
    conf_fix['recnet1:lr'] = np.float64(conf.optimizer['learning_rate'])
    conf2 = Configuration.from_json(os.path.join(REF, 'configs', '2-refinement.json'))
    conf_fix['refine2:seed'] = np.int64(conf2.seed)
    conf_fix['refine2:top_keys'] = np.array(sorted(k for k in conf2.__dict__
                                                   if not k.startswith('_')))
    gen_conf = Configuration.from_dict(conf2.generator_model, conf2)
    conf_fix['refine2:gen_mode'] = np.array(gen_conf.mode)
    conf_fix['refine2:disc_has_name'] = np.bool_(
        Configuration.from_dict(conf2.discriminator_model, conf2).has_attr('name'))
    np.savez_compressed(os.path.join(HERE, 'configs.npz'), **conf_fix)

    # ---- 9. the models around the frozen RecNet in configs/2-refinement.json:
    #         models/unet.py, models/discriminators.py built by the reference's own
    #         construct_model under a seed (keys, init RNG order, forward values) ----
    rm = {}

    def checksum(sd):
        return np.array([[float(v.double().sum()), float(v.double().pow(2).sum())]
                         for v in sd.values()])

    # (a) the full-size models of the shipped config: keys, shapes, weight checksums
    gconf = Configuration.from_dict(conf2.generator_model, conf2)
    uconf = Configuration.from_dict(gconf.learnable_model, conf2)
    torch.manual_seed(conf2.seed)
    unet_full = construct_model(uconf, uconf.name)
    dconf_d = dict(conf2.discriminator_model)
    dconf_d.setdefault('name', 'CNNDiscriminator')          # the JSON has no name
    dconf = Configuration.from_dict(dconf_d, conf2)
    disc_full = construct_model(dconf, dconf.name)
    for tag, net in (('unet_full', unet_full), ('disc_full', disc_full)):
        sd = net.state_dict()
        rm[tag + ':keys'] = np.array(list(sd.keys()))
        rm[tag + ':shapes'] = np.array([str(tuple(v.shape)) for v in sd.values()])
        rm[tag + ':checksum'] = checksum(sd)
        rm[tag + ':num_params'] = np.int64(sum(p.numel() for p in net.parameters()))
    with torch.no_grad():
        unet_full.eval()
        xin = torch.from_numpy(rs.normal(size=(1, 2, 64, 64)).astype(np.float32))
        rm['unet_full:x'] = xin.numpy()
        rm['unet_full:y_eval'] = unet_full(xin).numpy()
    del unet_full, disc_full

    # (b) small instances of the same architectures, weights and outputs stored
    small_u = dict(uconf.__dict__)
    small_u.update(encode_filters=[4, 8, 16], decode_filters=[8, 4])
    torch.manual_seed(3)
    unet_s = construct_model(Configuration.from_dict(small_u), 'UNET')
    small_d = dict(dconf_d)
    small_d.update(num_filters_per_layer=[4, 8, 8, 16, 16, 16], spatial_shape=[128, 128])
    disc_s = construct_model(Configuration.from_dict(small_d), 'CNNDiscriminator')
    for tag, net in (('unet_s', unet_s), ('disc_s', disc_s)):
        for k, v in net.state_dict().items():
            rm[tag + ':w:' + k] = v.numpy().copy()     # BN buffers change in the forward below
    xu = torch.from_numpy(rs.normal(size=(2, 2, 48, 40)).astype(np.float32))
    rm['unet_s:x'] = xu.numpy()
    unet_s.train()
    yu = unet_s(xu)
    yu.square().mean().backward()
    rm['unet_s:y_train'] = yu.detach().numpy()
    rm['unet_s:g_head'] = unet_s.head[0].weight.grad.numpy()
    rm['unet_s:g_first'] = unet_s.encode_units[0].encode[1].weight.grad.numpy()
    unet_s.eval()
    with torch.no_grad():
        rm['unet_s:y_eval'] = unet_s(xu).numpy()       # uses the running stats updated above
    xd = torch.from_numpy(rs.uniform(0, 1, size=(2, 1, 128, 128)).astype(np.float32))
    rm['disc_s:x'] = xd.numpy()
    disc_s.eval()
    with torch.no_grad():
        od = disc_s(xd)
    rm['disc_s:logits_eval'] = od['logits'].numpy()
    rm['disc_s:prob_eval'] = od['prob'].numpy()
    rm['disc_s:num_features'] = np.int64(len(od['features']))
    for i, f in enumerate(od['features']):
        rm['disc_s:feat%d_eval' % i] = f.numpy()
    # (c) adversarial / feature losses on those outputs (models/adversarial_loss.py)
    from models.adversarial_loss import GANLoss, FeatureMatchingLoss
    with torch.no_grad():
        od2 = disc_s(xd.flip(0) * 0.5)
    rm['loss:gan_disc'] = np.float64(GANLoss('disc', '', 0.1)(od, od2).item())
    rm['loss:gan_gen'] = np.float64(GANLoss('gen', '', 0.1)(od, od2).item())
    rm['loss:fm_gen'] = np.float64(FeatureMatchingLoss('gen', 'L1')(od, od2).item())
    np.savez_compressed(os.path.join(HERE, 'refinement_models.npz'), **rm)

    tot = sum(os.path.getsize(os.path.join(HERE, f))
              for f in os.listdir(HERE) if f.endswith('.npz'))
    print('golden fixtures written, %d bytes' % tot)


if __name__ == '__main__':
    main()
