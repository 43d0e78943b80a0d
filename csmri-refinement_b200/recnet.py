"""RecNet cascade with the B200 DC layer - mirror of models/recnet.py.

Same constructor arguments, same ``forward(self, inp, kspace, mask)`` parameter
names (the reference runner binds batch keys by name, training/base_runner.py:
43-63, and RefinementWrapper compares signatures, models/refinement_wrapper.py:
131-144), same ``state_dict`` keys (``conv_blocks.{i}.layers.{1,4,..}.weight``)
and the same weight-init RNG sequence (models/recnet.py:20-26,54-59 +
models/weight_inits.py:95-114), so checkpoints and seeds carry over.

What differs is only how the work is issued:
* the DC layers are :class:`csmri_refinement_b200.myfft.DataConsistencyInKspace`
  (one fused kernel instead of ~25 launches, models/recnet.py:151);
* the optional residual add of models/recnet.py:147-148 is folded into that
  kernel's load stage when a DC layer follows;
* zero "same" padding is given to cuDNN as the convolution's own padding
  instead of a separate ZeroPad2d pass over HBM (an ``nn.Identity`` keeps the
  Sequential indices, hence the checkpoint keys, unchanged).
The convolutions stay cuDNN fp32 through torch, except for the
weight gradient of the 3x3 layers between equal multiples of 32 channels, which
:mod:`csmri_refinement_b200.conv` computes with its own kernel (cuDNN's is half
of the fp32 training step on a B200).
"""
import math

import torch
import torch.nn as nn

from . import myfft
from .conv import Conv2d

RECNET_REQUIRED_PARAMS = ['num_blocks', 'num_convs', 'num_filters']
RECNET_OPTIONAL_PARAMS = ['num_final_outputs', 'dilations_per_conv', 'kernel_size',
                          'relu_leakiness', 'padding', 'use_refinement', 'skip_final_dc',
                          'return_intermediate_recs']

DEFAULT_RELU_LEAKINESS = 0.01


def _same_padding(kernel_size, dilation):
    # models/utils.py:76-85 for stride 1
    eff = kernel_size + (kernel_size - 1) * (dilation - 1)
    return int(math.ceil(eff - 1.0))


def _pad_and_conv(in_ch, out_ch, kernel_size, dilation, mode):
    total = _same_padding(kernel_size, dilation)
    side = total // 2
    if mode == 'zero' and total % 2 == 0:
        return nn.Identity(), Conv2d(in_ch, out_ch, kernel_size=kernel_size, stride=1,
                                     bias=True, dilation=dilation, padding=side)
    layers = {'zero': nn.ZeroPad2d, 'reflection': nn.ReflectionPad2d,
              'replication': nn.ReplicationPad2d}
    if mode not in layers:
        raise ValueError('Unknown padding mode %r' % (mode,))
    pad = side if total % 2 == 0 else (side, side + 1, side, side + 1)
    return layers[mode](pad), Conv2d(in_ch, out_ch, kernel_size=kernel_size, stride=1,
                                     bias=True, dilation=dilation)


class _FakeAct(object):
    """Shape / placement stand-in for the first layer's output (32 channels at the input's
    spatial size), so that the run's eligibility can be decided before that layer runs."""

    def __init__(self, x):
        self.is_cuda, self.dtype = x.is_cuda, x.dtype
        self.shape = (x.shape[0], 32, x.shape[2], x.shape[3])

    def dim(self):
        return 4

    def is_contiguous(self):
        return True


class ConvBlock(nn.Module):
    """models/recnet.py:29-62: (pad, conv, lrelu) x (num_convs-1), pad, conv."""

    def __init__(self, num_convs, num_filters, kernel_size, relu_leakiness, dilations,
                 padding='zero', num_inputs=2, num_outputs=2, final_act=False):
        super(ConvBlock, self).__init__()
        in_ch = num_inputs
        modules = []
        for i in range(num_convs - 1):
            pad, conv = _pad_and_conv(in_ch, num_filters, kernel_size, dilations[i], padding)
            # the activation is applied by the convolution module itself (one fused
            # bias + LeakyReLU pass); the Identity keeps the Sequential indices
            conv.fused_slope = relu_leakiness
            modules += [pad, conv, nn.Identity()]
            in_ch = num_filters
        pad, conv = _pad_and_conv(in_ch, num_outputs, kernel_size, dilations[-1], padding)
        modules += [pad, conv]
        if final_act:
            modules.append(nn.LeakyReLU(relu_leakiness, inplace=True))
        self.layers = nn.Sequential(*modules)

    def initialize(self):
        """Reference init (recnet.py:54-59, weight_inits.py:95-114), visiting the
        convolutions in ``Module.apply`` order so the RNG stream is identical:
        first conv xavier-uniform gain 1 with its bias LEFT at torch's default
        init (the per-module override at recnet.py:58 names only 'weight', so
        weight_inits.py:72-79 finds no bias rule for it), the other convs
        He-normal with a=0.01 and bias 0."""
        first = True
        for m in self.layers:
            if isinstance(m, nn.Conv2d):
                if first:
                    nn.init.xavier_uniform_(m.weight.data, gain=1.0)
                    first = False
                else:
                    nn.init.kaiming_normal_(m.weight.data, a=DEFAULT_RELU_LEAKINESS)
                    nn.init.constant_(m.bias.data, 0.0)

    def forward(self, x):
        # the run of 32 -> 32 layers in the middle of the block as one autograd node (the
        # LeakyReLU derivatives ride on the data-gradient kernels); same values, same
        # parameters - anything the fused node does not cover takes the plain Sequential
        mods = list(self.layers)
        convs = [m for m in mods if isinstance(m, nn.Conv2d)]
        inner = convs[1:-1]
        if torch.is_grad_enabled() and len(inner) >= 1 and all(
                isinstance(m, (nn.Conv2d, nn.Identity)) for m in mods) and convs[0].fused_slope is not None:
            from . import conv as _conv
            # the run's eligibility depends on placement and spatial shape only, so it can be
            # decided before the first layer has produced the run's actual input
            if x.dim() == 4 and x.is_contiguous() and _conv.tc_chain_eligible(_FakeAct(x), inner):
                return _conv.tc_chain(x, inner, last=convs[-1], first=convs[0])
            y = convs[0](x)
            for m in inner:
                y = m(y)
            return convs[-1](y)
        return self.layers(x)


class RecNet(nn.Module):
    """Reconstruction network (Schlemper et al. deep cascade), models/recnet.py:65-161."""

    DEFAULT_RELU_LEAKINESS = DEFAULT_RELU_LEAKINESS

    def __init__(self, num_blocks, num_convs, num_filters, num_final_outputs=2,
                 dilations_per_conv=1, kernel_size=3, relu_leakiness=DEFAULT_RELU_LEAKINESS,
                 padding='zero', use_refinement=False, skip_final_dc=False,
                 return_intermediate_recs=False, dc_factory=None):
        super(RecNet, self).__init__()
        if isinstance(num_filters, int):
            num_filters = [num_filters] * num_blocks
        if isinstance(dilations_per_conv, int):
            dilations_per_conv = [dilations_per_conv] * num_convs
        assert len(num_filters) == num_blocks, \
            'Number of given filters must match number of blocks'
        assert len(dilations_per_conv) == num_convs, \
            'Number of dilations must match number of convolutions'

        blocks = []
        for idx, nf in enumerate(num_filters):
            n_out = 2 if idx < num_blocks - 1 else num_final_outputs
            blocks.append(ConvBlock(num_convs, nf, kernel_size, relu_leakiness,
                                    padding=padding, num_outputs=n_out,
                                    dilations=dilations_per_conv))
        if dc_factory is None:
            dc_factory = lambda: myfft.DataConsistencyInKspace(norm='ortho')  # noqa: E731
        n_dc = num_blocks if not skip_final_dc else num_blocks - 1
        self.conv_blocks = nn.ModuleList(blocks)
        # a plain list on purpose (recnet.py:128-134): no parameters, no buffers
        self.dc_layers = [dc_factory() for _ in range(n_dc)]
        self.use_refinement = use_refinement
        self.skip_final_dc = skip_final_dc
        self.return_intermediate_recs = return_intermediate_recs

    def forward(self, inp, kspace, mask):
        x = inp
        reconstructions = []
        for idx in range(len(self.conv_blocks)):
            block_input = x
            x = self.conv_blocks[idx](x)
            has_dc = idx < len(self.dc_layers)
            if has_dc:
                dc = self.dc_layers[idx]
                if self.use_refinement and isinstance(dc, myfft.DataConsistencyInKspace):
                    x = dc.perform(x, kspace, mask, residual=block_input)
                else:
                    if self.use_refinement:
                        x = x + block_input
                    x = dc.perform(x, kspace, mask)
                if self.return_intermediate_recs:
                    reconstructions.append(x)
            elif self.use_refinement:
                x = x + block_input
        if self.return_intermediate_recs:
            return {'pred': x, 'reconstructions': reconstructions}
        return x


def construct_model(conf, model_name=None, **kwargs):
    """models/recnet.py:20-26: build from a reference ``Configuration`` (or any
    object with ``to_param_dict``) or a plain dict of the same keys."""
    user_init = conf.get_attr('weight_init', default={}) if hasattr(conf, 'get_attr') \
        else conf.get('weight_init', {})
    if user_init:
        # the reference merges these rules into the per-block init
        # (models/recnet.py:23-24, weight_inits.py:109-114); neither shipped config
        # sets them for RecNet, and silently ignoring them would change the weights
        raise NotImplementedError('a RecNet `weight_init` override is not supported: %r'
                                  % (user_init,))
    if hasattr(conf, 'to_param_dict'):
        params = conf.to_param_dict(RECNET_REQUIRED_PARAMS, RECNET_OPTIONAL_PARAMS)
    else:
        missing = [k for k in RECNET_REQUIRED_PARAMS if k not in conf]
        if missing:
            raise ValueError('missing RecNet parameters: %s' % missing)
        params = {k: conf[k] for k in RECNET_REQUIRED_PARAMS + RECNET_OPTIONAL_PARAMS
                  if k in conf}
    params.update(kwargs)
    model = RecNet(**params)
    for block in model.conv_blocks:
        block.initialize()
    return model
