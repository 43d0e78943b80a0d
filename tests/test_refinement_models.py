"""CPU tests of the config-5 model mirrors (csmri_refinement_b200.refinement_models)
against fixtures produced by the reference's own classes
(tests/golden/make_golden.py section 9: models/unet.py, models/discriminators.py,
models/adversarial_loss.py built / evaluated by the unmodified reference)."""
import os

import numpy as np
import torch

from oracle import dc_oracle as orc


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, 'refinement_models.npz'))


def _checksum(sd):
    return np.array([[float(v.double().sum()), float(v.double().pow(2).sum())]
                     for v in sd.values()])


def test_full_size_unet_and_discriminator_of_2_refinement_json(golden_dir):
    """Built from the UNCHANGED configs/2-refinement.json under its own seed:
    same state_dict keys and shapes, same parameter counts (920,033 and
    27,941,697) and the same initial weights as the reference -
    i.e. the init rules and their RNG order carry over."""
    from csmri_refinement_b200 import harness, refinement_harness as rh
    g = _load(golden_dir)
    conf = harness.load_config(harness.config_path('2-refinement.json'))
    torch.manual_seed(conf.seed)
    unet = rh.build_learnable_model(conf)
    disc = rh.build_discriminator(conf)
    for tag, net in (('unet_full', unet), ('disc_full', disc)):
        sd = net.state_dict()
        assert list(sd.keys()) == [str(k) for k in g[tag + ':keys']], tag
        assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g[tag + ':shapes']]
        assert sum(p.numel() for p in net.parameters()) == int(g[tag + ':num_params'])
        assert np.allclose(_checksum(sd), g[tag + ':checksum'], rtol=1e-12, atol=0), tag
    assert int(g['unet_full:num_params']) == 920033
    assert int(g['disc_full:num_params']) == 27941697
    unet.eval()
    with torch.no_grad():
        y = unet(torch.from_numpy(g['unet_full:x']))
    assert orc.rel_l2(y.numpy(), g['unet_full:y_eval']) < 1e-5
    out = disc.eval()(torch.zeros(1, 1, 512, 512))
    assert out['logits'].shape == (1, 1, 13, 13) and len(out['features']) == 7


def test_small_unet_matches_reference_outputs_and_gradients(golden_dir):
    from csmri_refinement_b200 import harness, refinement_models as rm
    from csmri_refinement_b200.config import Configuration
    g = _load(golden_dir)
    conf = harness.load_config(harness.config_path('2-refinement.json'))
    small = dict(conf.generator_model['learnable_model'])
    small.update(encode_filters=[4, 8, 16], decode_filters=[8, 4])
    torch.manual_seed(3)
    net = rm.construct_unet(Configuration.from_dict(small))
    for k, v in net.state_dict().items():
        assert np.array_equal(v.numpy(), g['unet_s:w:' + k]), k       # same seed, same weights
    x = torch.from_numpy(g['unet_s:x'])
    net.train()
    y = net(x)
    y.square().mean().backward()
    assert orc.rel_l2(y.detach().numpy(), g['unet_s:y_train']) < 1e-5
    assert orc.rel_l2(net.head[0].weight.grad.numpy(), g['unet_s:g_head']) < 1e-4
    assert orc.rel_l2(net.encode_units[0].encode[1].weight.grad.numpy(), g['unet_s:g_first']) < 1e-4
    net.eval()
    with torch.no_grad():
        assert orc.rel_l2(net(x).numpy(), g['unet_s:y_eval']) < 1e-5


def test_small_discriminator_and_losses_match_reference(golden_dir):
    from csmri_refinement_b200 import harness, refinement_harness as rh, refinement_models as rm
    from csmri_refinement_b200.config import Configuration
    g = _load(golden_dir)
    conf = harness.load_config(harness.config_path('2-refinement.json'))
    small = dict(conf.discriminator_model)
    small.update(num_filters_per_layer=[4, 8, 8, 16, 16, 16], spatial_shape=[128, 128])
    # the generator of section 9(b) was seeded once for both small models
    torch.manual_seed(3)
    su = dict(conf.generator_model['learnable_model'])
    su.update(encode_filters=[4, 8, 16], decode_filters=[8, 4])
    rm.construct_unet(Configuration.from_dict(su))
    disc = rm.construct_discriminator(Configuration.from_dict(small))
    for k, v in disc.state_dict().items():
        assert np.array_equal(v.numpy(), g['disc_s:w:' + k]), k
    x = torch.from_numpy(g['disc_s:x'])
    disc.eval()
    with torch.no_grad():
        out = disc(x)
        out2 = disc(x.flip(0) * 0.5)
    assert orc.rel_l2(out['logits'].numpy(), g['disc_s:logits_eval']) < 1e-5
    assert orc.rel_l2(out['prob'].numpy(), g['disc_s:prob_eval']) < 1e-5
    assert len(out['features']) == int(g['disc_s:num_features'])
    for i, f in enumerate(out['features']):
        assert orc.rel_l2(f.numpy(), g['disc_s:feat%d_eval' % i]) < 1e-5, i
    assert abs(rh.gan_loss_disc(out, out2, 0.1).item() - float(g['loss:gan_disc'])) < 1e-6
    assert abs(rh.gan_loss_gen(out).item() - float(g['loss:gan_gen'])) < 1e-6
    assert abs(rh.feature_matching_loss(out, out2).item() - float(g['loss:fm_gen'])) < 1e-6


def test_vgg19_blocks_follow_the_torchvision_layout():
    """models/vgg.py:37-45: block k ends before a max-pool, names are torchvision's
    `features` indices; a torchvision VGG19 state dict loads by key mapping."""
    from csmri_refinement_b200 import refinement_models as rm
    import torchvision
    torch.manual_seed(0)
    tv = torchvision.models.vgg19(weights=None).features
    vgg = rm.VGG19()
    assert len(vgg.blocks) == 5
    assert [n for n, _ in vgg.blocks[0].named_children()] == ['0', '1', '2', '3']
    assert [n for n, _ in vgg.blocks[4].named_children()][0] == '27'
    assert not any(p.requires_grad for p in vgg.parameters())
    vgg.load_torchvision_features({'features.' + k: v for k, v in tv.state_dict().items()})
    x = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        want = tv[:36]((x - vgg.mean) / vgg.std)          # up to relu5_4, before the last pool
        got = vgg(x)
    assert len(got) == 1 and torch.allclose(got[0], want, atol=1e-6)
