"""Models around the frozen RecNet in ``configs/2-refinement.json`` (BASELINE
configs[4]): the learnable U-Net, the PatchGAN discriminator and the VGG19
feature extractor of the perceptual loss.

They are NOT on the DC hot path - their convolutions stay stock cuDNN through
torch ("out of scope as kernels; the harness for config 5").  What
must carry over from the reference is the *contract*: constructor arguments as
the JSON spells them, ``state_dict`` keys (checkpoints), the output dicts, and
the weight-init RNG order (same seed -> same weights).  ``tests/golden/
refinement_models.npz`` pins all of that against the reference's own classes.

Mirrors (paths relative to the reference root):
  models/unet.py:32-290            ConvEncodeUnit / ConvDecodeUnit / UNET
  models/discriminators.py:51-247  CNNDiscriminator
  models/vgg.py:8-80               VGG19 (feature blocks split at the max-pools)
  models/weight_inits.py:5-114     per-class init rules, applied in Module.apply order
  models/utils.py:55-85            "same" padding layers (even kernels pad (l, l+1))
Only the option values the shipped config reaches (plus their obvious
neighbours) are implemented; anything else raises instead of silently
building a different network.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import Configuration


# ---------------------------------------------------------------------------
# padding and init helpers
# ---------------------------------------------------------------------------
_PAD_LAYERS = {'zero': nn.ZeroPad2d, 'reflection': nn.ReflectionPad2d,
               'replication': nn.ReplicationPad2d}


def same_padding_layer(kernel_size, stride, mode='zero', dilation=1):
    """models/utils.py:55-85: total = ceil((k_eff - 1) / stride), split (t//2, t//2 [+1])."""
    if stride not in (1, 2):
        raise ValueError('same padding is defined for stride 1 or 2')
    if mode not in _PAD_LAYERS:
        raise ValueError('Unknown padding mode %r' % (mode,))
    k_eff = kernel_size + (kernel_size - 1) * (dilation - 1)
    total = int(math.ceil((k_eff - 1.0) / stride))
    side = total // 2
    pad = side if total % 2 == 0 else (side, side + 1, side, side + 1)
    return _PAD_LAYERS[mode](pad)


DEFAULT_INITS = {                       # models/weight_inits.py:5-14
    'conv_weight': ('he_normal', 0.0), 'conv_bias': ('constant', 0.0),
    'conv_transposed_weight': ('he_normal', 0.0), 'conv_transposed_bias': ('constant', 0.0),
    'batchnorm_weight': ('constant', 1.0), 'batchnorm_bias': ('constant', 0.0),
    'linear_weight': ('xavier_normal', 'linear'), 'linear_bias': ('constant', 0.0),
}


def _apply_init(spec, tensor):
    """One rule of models/weight_inits.py:17-66 on one tensor."""
    name = spec[0] if isinstance(spec, (tuple, list)) else spec
    args = list(spec[1:]) if isinstance(spec, (tuple, list)) else []
    if name == 'torch_default':
        return
    if name == 'zero':
        nn.init.constant_(tensor, 0.0)
    elif name == 'constant':
        nn.init.constant_(tensor, args[0])
    elif name == 'normal':
        nn.init.normal_(tensor, mean=args[0], std=args[1])
    elif name == 'uniform':
        nn.init.uniform_(tensor, a=args[0], b=args[1])
    elif name in ('xavier_normal', 'xavier', 'xavier_uniform'):
        gain = nn.init.calculate_gain(args[0]) if isinstance(args[0], str) else args[0]
        (nn.init.xavier_normal_ if name == 'xavier_normal' else nn.init.xavier_uniform_)(
            tensor, gain=gain)
    elif name in ('he_normal', 'he_uniform'):
        a = args[0] if args else 0.0
        (nn.init.kaiming_normal_ if name == 'he_normal' else nn.init.kaiming_uniform_)(tensor, a=a)
    elif name == 'orthogonal':
        gain = args[0] if args else 1.0
        if isinstance(gain, str):
            gain = nn.init.calculate_gain(gain, args[1] if len(args) > 1 else None)
        nn.init.orthogonal_(tensor, gain=gain)
    else:
        raise ValueError('Unknown weight init %r' % (name,))


_CLASS_KEYS = (('ConvTranspose2d', 'conv_transposed'), ('Conv2d', 'conv'), ('Linear', 'linear'),
               ('BatchNorm2d', 'batchnorm'))


def initialize_weights(model, model_rules, user_rules=None):
    """models/weight_inits.py:109-114: defaults < the model's own rules < the
    config's ``weight_init`` section, applied to every module in
    ``Module.apply`` order (children first) - which fixes the RNG sequence."""
    rules = dict(DEFAULT_INITS)
    rules.update(model_rules)
    rules.update(user_rules or {})

    def visit(m):
        cls = type(m).__name__
        for needle, prefix in _CLASS_KEYS:
            if needle in cls:
                w = rules.get(prefix + '_weight') if getattr(m, 'weight', None) is not None else None
                b = rules.get(prefix + '_bias') if getattr(m, 'bias', None) is not None else None
                if w is not None:
                    _apply_init(w, m.weight.data)
                if b is not None:
                    _apply_init(b, m.bias.data)
                return

    model.apply(visit)


# ---------------------------------------------------------------------------
# U-Net (learnable_model of the refinement wrapper)
# ---------------------------------------------------------------------------
class ConvEncodeUnit(nn.Module):
    """models/unet.py:42-75: (pad, conv, [bn], [lrelu]) x num_layers, optional 2x2 max-pool;
    with pooling, forward returns (pooled, before_pooling)."""

    def __init__(self, in_channels, num_layers, num_filters, kernel_size, relu_leakiness, use_bn,
                 downsample, use_act=True, padding='zero'):
        super(ConvEncodeUnit, self).__init__()
        self.downsample = downsample
        mods = []
        for _ in range(num_layers):
            mods.append(same_padding_layer(kernel_size, 1, padding))
            mods.append(nn.Conv2d(in_channels, num_filters, kernel_size, stride=1, bias=not use_bn))
            in_channels = num_filters
            if use_bn:
                mods.append(nn.BatchNorm2d(num_filters))
            if use_act:
                mods.append(nn.LeakyReLU(relu_leakiness, inplace=True))
        self.encode = nn.Sequential(*mods)
        if downsample:
            self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    def forward(self, inp):
        x = self.encode(inp)
        if self.downsample:
            return self.pool(x), x
        return x


class ConvDecodeUnit(nn.Module):
    """models/unet.py:78-156 for the resize-convolution and transposed modes."""

    def __init__(self, in_channels, encoder_channels, num_filters, relu_leakiness, use_bn,
                 use_act=True, kernel_size=3, transposed_kernel_size=2, num_layers=0,
                 mode='transposed', padding='zero', act_upsampling_only=False):
        super(ConvDecodeUnit, self).__init__()
        use_bias = not use_bn or encoder_channels == 0
        if mode == 'transposed':
            up = [nn.ConvTranspose2d(in_channels, num_filters, kernel_size=transposed_kernel_size,
                                     stride=2, bias=use_bias)]
            in_channels = num_filters
        elif mode in ('nn-resize-conv', 'nn-biresize-conv'):
            up = [nn.Upsample(scale_factor=2,
                              mode='nearest' if mode == 'nn-resize-conv' else 'bilinear'),
                  same_padding_layer(kernel_size, 1, padding),
                  nn.Conv2d(in_channels, num_filters, kernel_size, stride=1, bias=use_bias)]
            in_channels = num_filters
        elif mode in ('nn', 'bilinear'):
            up = [nn.Upsample(scale_factor=2, mode='nearest' if mode == 'nn' else 'bilinear')]
        else:
            raise NotImplementedError('upsampling_mode %r is not reachable from the shipped '
                                      'configs' % (mode,))
        dec = []
        norm_act = up if act_upsampling_only else dec
        width = in_channels if act_upsampling_only else in_channels + encoder_channels
        if use_bn:
            norm_act.append(nn.BatchNorm2d(width))
        if use_act:
            norm_act.append(nn.LeakyReLU(relu_leakiness, inplace=True))
        if num_layers > 0:
            dec.append(ConvEncodeUnit(in_channels + encoder_channels, num_layers, num_filters,
                                      kernel_size, relu_leakiness, use_bn, downsample=False,
                                      use_act=use_act, padding=padding))
        self.upsample = nn.Sequential(*up)
        self.decode = nn.Sequential(*dec)

    def forward(self, decode_path, encode_path=None):
        x = self.upsample(decode_path)
        if encode_path is not None:
            dh = encode_path.shape[2] - x.shape[2]
            dw = encode_path.shape[3] - x.shape[3]
            if dh or dw:                                   # models/unet.py:31-39
                x = F.pad(x, (0, dw, 0, dh), mode='reflect')
            x = torch.cat((encode_path, x), dim=1)
        return self.decode(x)


UNET_REQUIRED = ['num_inputs', 'num_outputs', 'num_layers_per_scale', 'encode_filters',
                 'decode_filters', 'output_activation']
UNET_OPTIONAL = ['kernel_size', 'transposed_kernel_size', 'relu_leakiness', 'use_bn',
                 'upsampling_mode', 'padding', 'encoder_features', 'use_refinement',
                 'decoder_act_upsampling_only']


class UNET(nn.Module):
    """models/unet.py:159-290."""

    DEFAULT_RELU_LEAKINESS = 0.1

    def __init__(self, num_inputs, num_outputs, num_layers_per_scale, encode_filters,
                 decode_filters, output_activation, kernel_size=3, transposed_kernel_size=2,
                 relu_leakiness=DEFAULT_RELU_LEAKINESS, use_bn=True, upsampling_mode='transposed',
                 padding='zero', encoder_features=None, use_refinement=False,
                 decoder_act_upsampling_only=False):
        super(UNET, self).__init__()
        if output_activation not in ('softmax', 'tanh', 'none'):
            raise ValueError('output_activation must be softmax, tanh or none')
        self.encoder_features = encoder_features
        self.use_refinement = use_refinement
        if isinstance(relu_leakiness, float):
            relu_leakiness = (relu_leakiness, relu_leakiness)
        n_enc = len(encode_filters)
        ch = num_inputs
        enc = []
        for s, nf in enumerate(encode_filters):
            enc.append(ConvEncodeUnit(ch, num_layers_per_scale, nf, kernel_size, relu_leakiness[0],
                                      use_bn, downsample=(s != n_enc - 1), padding=padding))
            ch = nf
        common = dict(kernel_size=kernel_size, transposed_kernel_size=transposed_kernel_size,
                      num_layers=num_layers_per_scale, mode=upsampling_mode, padding=padding,
                      act_upsampling_only=decoder_act_upsampling_only)
        cat_dec = []
        for s, nf in enumerate(decode_filters[:n_enc - 1]):
            cat_dec.append(ConvDecodeUnit(ch, encode_filters[-(s + 2)], nf, relu_leakiness[1],
                                          use_bn, **common))
            ch = nf
        dec = []
        for nf in decode_filters[n_enc - 1:]:
            dec.append(ConvDecodeUnit(ch, 0, nf, relu_leakiness[1], use_bn, **common))
            ch = nf
        head = [nn.Conv2d(ch, num_outputs, kernel_size=1, stride=1, padding=0, bias=True)]
        if output_activation == 'softmax':
            head.append(nn.Softmax(dim=1))
        elif output_activation == 'tanh':
            head.append(nn.Tanh())
        self.encode_units = nn.ModuleList(enc)
        self.concat_decode_units = nn.ModuleList(cat_dec)
        self.decode_units = nn.ModuleList(dec)
        self.head = nn.Sequential(*head)

    @staticmethod
    def weight_init_params():
        return {'conv_weight': ('he_normal', UNET.DEFAULT_RELU_LEAKINESS),
                'conv_transposed_weight': ('he_normal', UNET.DEFAULT_RELU_LEAKINESS),
                'batchnorm_weight': ('uniform', 0.98, 1.02)}

    def forward(self, inp):
        x = inp
        skips = []
        last = None
        for unit in self.encode_units:
            if unit.downsample:
                x, before = unit(x)
                skips.append(before)
            else:
                x = last = unit(x)
        for s, unit in enumerate(self.concat_decode_units):
            x = unit(x, skips[-(s + 1)])
        for unit in self.decode_units:
            x = unit(x)
        pred = self.head(x)
        if self.use_refinement:
            pred = inp + pred
        if self.encoder_features is not None:
            feats = skips + [last]
            return {'pred': pred, 'features': [feats[i] for i in self.encoder_features]}
        return pred


def construct_unet(conf):
    """models/unet.py:22-26."""
    conf = Configuration.from_dict(conf)
    model = UNET(**conf.to_param_dict(UNET_REQUIRED, UNET_OPTIONAL))
    initialize_weights(model, UNET.weight_init_params(), conf.get_attr('weight_init', default={}))
    return model


# ---------------------------------------------------------------------------
# PatchGAN discriminator
# ---------------------------------------------------------------------------
DISC_REQUIRED = ['num_inputs', 'num_filters_per_layer', 'strides']
DISC_OPTIONAL = ['kernel_sizes', 'fc_layers', 'spatial_shape', 'act_fn', 'relu_leakiness',
                 'use_norm_layers', 'norm_layer', 'use_weightnorm', 'padding',
                 'final_conv_kernel_size', 'final_average_pooling', 'use_biases',
                 'compute_features', 'dropout_after', 'dropout_prob']


def _needs_bias(use_norm_layers, norm_layer):
    """models/utils.py:44-52."""
    return (not use_norm_layers) or use_norm_layers == 'not-first' or norm_layer == 'instance'


class CNNDiscriminator(nn.Module):
    """models/discriminators.py:51-247 without fully connected layers (the
    shipped config is a PatchGAN: a final 4x4 convolution to one logit map)."""

    DEFAULT_RELU_LEAKINESS = 0.2

    def __init__(self, num_inputs, num_filters_per_layer, strides, kernel_sizes=None,
                 fc_layers=(), spatial_shape=None, act_fn='lrelu',
                 relu_leakiness=DEFAULT_RELU_LEAKINESS, use_norm_layers=True, norm_layer='batch',
                 use_weightnorm=False, padding='zero', final_conv_kernel_size=1, use_biases=True,
                 final_average_pooling=False, compute_features=False, dropout_after=(),
                 dropout_prob=0.5):
        super(CNNDiscriminator, self).__init__()
        if len(fc_layers) > 0 or use_weightnorm:
            raise NotImplementedError('fully connected heads / weight norm are not reachable '
                                      'from the shipped configs')
        if act_fn not in ('lrelu', 'relu'):
            raise NotImplementedError('act_fn %r' % (act_fn,))
        if norm_layer not in ('batch', 'instance', 'instance-affine'):
            raise ValueError('Unknown normalization layer %r' % (norm_layer,))
        if kernel_sizes is None:
            kernel_sizes = 3
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * len(num_filters_per_layer)
        assert len(num_filters_per_layer) == len(strides) == len(kernel_sizes)
        self.compute_features = compute_features
        self.feature_layers = set()
        ch = num_inputs
        layers = []
        for idx, (nf, k, s) in enumerate(zip(num_filters_per_layer, kernel_sizes, strides)):
            bias = use_biases and _needs_bias(use_norm_layers, norm_layer)
            layers.append(same_padding_layer(k, s, padding))
            layers.append(nn.Conv2d(ch, nf, kernel_size=k, stride=s, bias=bias))
            if use_norm_layers == 'not-first':
                use_norm_layers = True                     # the first layer stays un-normalised
            elif use_norm_layers:
                layers.append(nn.BatchNorm2d(nf, affine=True) if norm_layer == 'batch' else
                              nn.InstanceNorm2d(nf, affine=(norm_layer == 'instance-affine')))
            layers.append(nn.LeakyReLU(relu_leakiness, inplace=True) if act_fn == 'lrelu'
                          else nn.ReLU(inplace=True))
            if compute_features:
                self.feature_layers.add(layers[-1])
            if idx in dropout_after:
                # out of place: torch >= 1.5 refuses to back-propagate through an in-place
                # dropout on the saved output of an in-place LeakyReLU (see forward)
                layers.append(nn.Dropout2d(p=dropout_prob, inplace=False))
            ch = nf
        self.convs = nn.Sequential(*layers)
        self.fcs = None
        final = [nn.Conv2d(ch, 1, kernel_size=final_conv_kernel_size, stride=1, bias=use_biases)]
        if final_average_pooling:
            final.append(nn.AdaptiveAvgPool2d((1, 1)))
        self.final_conv = nn.Sequential(*final)

    @staticmethod
    def weight_init_params():
        return {'conv_weight': ('normal', 0.0, 0.02), 'linear_weight': ('normal', 0.0, 0.02),
                'batchnorm_weight': ('normal', 1.0, 0.02)}

    def forward(self, inp):
        x = inp
        features = []
        fresh = False                      # features[-1] is the activation just computed
        for layer in self.convs:
            x = layer(x)
            if self.compute_features and layer in self.feature_layers:
                features.append(x)
                fresh = True
            elif fresh and isinstance(layer, nn.Dropout2d):
                # the reference's dropout is in place on the tensor it has just put into
                # `features` (discriminators.py:170-176,215-220), so the feature map of such
                # a layer is the dropped-out one
                features[-1] = x
                fresh = False
            else:
                fresh = False
        x = self.final_conv(x)
        out = {'prob': torch.sigmoid(x), 'logits': x}
        if self.compute_features:
            features.append(x)
            out['features'] = features
        return out


def construct_discriminator(conf):
    """models/discriminators.py:27-34; a ``weight_init`` override with the
    reference's per-layer ``final_layer_bias`` hack is not supported."""
    conf = Configuration.from_dict(conf)
    user = conf.get_attr('weight_init', default={})
    if 'final_layer_bias' in user:
        raise NotImplementedError('final_layer_bias init override')
    model = CNNDiscriminator(**conf.to_param_dict(DISC_REQUIRED, DISC_OPTIONAL))
    initialize_weights(model, CNNDiscriminator.weight_init_params(), user)
    return model


# ---------------------------------------------------------------------------
# VGG19 feature blocks (perceptual loss)
# ---------------------------------------------------------------------------
_VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M',
              512, 512, 512, 512, 'M']


class VGG19(nn.Module):
    """models/vgg.py:8-80: torchvision's VGG19 ``features`` cut into blocks that
    each END before a max-pool (block k > 0 starts with one); module names inside a
    block are the torchvision ``features`` indices, so ``load_torchvision_features``
    maps ``features.N.*`` of a stock checkpoint onto ``blocks.B.N.*``.  Weights are
    whatever the caller loads; the reference downloads the ImageNet ones
    (models/vgg.py:35), which is impossible offline - the harness uses a seeded
    random init and says so."""

    LAST_FEATURE_MAP = 4

    def __init__(self, output_blocks=(LAST_FEATURE_MAP,), requires_grad=False):
        super(VGG19, self).__init__()
        self.output_blocks = sorted(output_blocks)
        last = self.output_blocks[-1]
        assert 0 <= last <= 5, 'VGG19 has at most 6 blocks'
        self.blocks = nn.ModuleList([nn.Sequential()])
        ch, idx = 3, 0
        for v in _VGG19_CFG:
            if v == 'M':
                if len(self.blocks) - 1 == last:
                    break
                self.blocks.append(nn.Sequential())
                self.blocks[-1].add_module(str(idx), nn.MaxPool2d(kernel_size=2, stride=2))
                idx += 1
            else:
                self.blocks[-1].add_module(str(idx), nn.Conv2d(ch, v, kernel_size=3, padding=1))
                self.blocks[-1].add_module(str(idx + 1), nn.ReLU(inplace=True))
                ch, idx = v, idx + 2
        for p in self.parameters():
            p.requires_grad = requires_grad
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]).reshape(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]).reshape(1, 3, 1, 1))

    def load_torchvision_features(self, state_dict):
        own = self.state_dict()
        for key in list(own):
            if key.startswith('blocks.'):
                _, _, idx, leaf = key.split('.')
                own[key] = state_dict['features.%s.%s' % (idx, leaf)]
        self.load_state_dict(own)

    def forward(self, inp):
        """inp (B,3,H,W) in (0, 1) -> list of the requested blocks' feature maps."""
        x = inp.sub(self.mean).div(self.std)
        out = []
        for b, block in enumerate(self.blocks):
            x = block(x)
            if b in self.output_blocks:
                out.append(x)
        return out
