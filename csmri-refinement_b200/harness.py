"""Run the reference's JSON configurations on the B200 path.

``train.py configs/1-recnet.json`` in the reference is:
``Configuration.from_json`` -> ``set_random_seeds(conf.seed)`` ->
``build_runner`` (``construct_model(Configuration.from_dict(conf.model, conf))``,
``get_criterion('MSE')``, ``cudaify``, ``get_optimizer``: training/runner.py:18-76,
training/optimizers.py:5-24) -> ``DataLoader(batch_size=conf.batch_size)`` over
ScarSeg 512x512 slices pushed through ``rec_transforms.train_transform``
(CenterCropInKspace, /max, ``Undersample('varden', acc, ...)``) -> per batch
``Runner._train_step``.

The runners, CLI and dataset are out of scope (torch-0.3 idioms, proprietary
data); what is kept is everything the JSON *decides*: architecture, seed and
init, loss, optimizer and its hyper-parameters, the GLOBAL batch size, the
undersampling scheme and acceleration.  The JSON files themselves are read
unchanged.  Data is synthetic (uniform images of the dataset's 512x512 shape,
scar_segmentation.py:22), sharded by rank the way ``shard_range`` splits the
global batch; the step is :class:`csmri_refinement_b200.parallel.ShardedTrainer`.
"""
import os

import numpy as np
import torch

from . import parallel, recnet, undersampling
from .config import Configuration

# data/reconstruction/scar_seg/scar_segmentation.py:22 (no `downscale` key in the
# shipped configs => train_transform's default 1, data/transform_wrappers.py:33-36)
SCARSEG_IMAGE_SIZE = 512
CENTRAL_LINES = 8            # myImageTransformations.py:80-81


def load_config(path):
    return Configuration.from_json(path)


def set_random_seeds(seed):
    """utils/__init__.py:24-30."""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def build_recnet(conf):
    """training/runner.py:19-21 for ``conf.model.name == 'RecNet'``: the model is
    built under the config's seed, so the initial weights are the reference's."""
    model_conf = Configuration.from_dict(conf.model, conf)
    if model_conf.name != 'RecNet':
        raise ValueError('expected a RecNet model section, got %r' % (model_conf.name,))
    return recnet.construct_model(model_conf, model_conf.name)


def adam_args(opt_section, parent=None):
    """training/optimizers.py:18-21."""
    opt = Configuration.from_dict(opt_section, parent)
    if opt.name != 'Adam':
        raise ValueError('only the Adam optimizer of the shipped configs is supported, got %r'
                         % (opt.name,))
    return {'lr': opt.learning_rate,
            'betas': (opt.get_attr('beta1', default=0.9), opt.get_attr('beta2', default=0.999))}


def undersampling_args(conf):
    cs = conf.undersampling
    if cs['sampling_scheme'] == 'radial':
        raise NotImplementedError('radial sampling is not reachable from the shipped configs')
    return {'acc': cs['acceleration_factor'], 'variable': cs.get('variable_acceleration', False)}


def synthetic_batch(conf, n_slices, device, seed, image_size=None):
    """``n_slices`` training samples of the config's data path on the GPU:
    uniform (0,1) images of the dataset's shape, Cartesian lines drawn on the
    host exactly as ``cs.cartesian_mask`` does (acc and the 8 central lines of
    myImageTransformations.py:71-81), ``cs.undersample`` on the device.
    Returns the batch dict {inp, kspace, mask, target}."""
    n = image_size or conf.get_attr('image_size', default=SCARSEG_IMAGE_SIZE)
    us = undersampling_args(conf)
    if us['variable']:
        raise NotImplementedError('variable acceleration: use undersampling.Undersample')
    g = torch.Generator(device=device).manual_seed(seed)
    img = torch.rand(n_slices, n, n, device=device, generator=g)
    rows = undersampling.cartesian_rows((n_slices, n, n), us['acc'], CENTRAL_LINES, False,
                                        np.random.RandomState(seed))
    return undersampling.undersample(img, rows)


def recnet_trainer(conf, device, rank=0, world=1, cuda_graph=True):
    """(trainer, local_batch): the MSE + Adam training step of the config on this
    rank's shard of the GLOBAL ``conf.batch_size`` (the reference's batch is the
    global one, split by DataParallel: SURVEY 8e)."""
    if conf.get_attr('loss_name') != 'MSE':
        raise ValueError('expected loss_name MSE (models/criteria.py:112-128)')
    set_random_seeds(conf.seed)                    # identical replicas on every rank
    model = build_recnet(conf).to(device)
    lo, hi = parallel.shard_range(int(conf.batch_size), rank, world)
    trainer = parallel.ShardedTrainer(model, cuda_graph=cuda_graph and device.type == 'cuda',
                                      assume_row_constant=True if device.type == 'cuda' else None,
                                      **adam_args(conf.optimizer, conf))
    return trainer, hi - lo


def config_path(name):
    """The shipped configs, byte-identical copies of the reference's
    ``configs/<name>`` (tests/golden/configs; sha256 pinned in configs.npz)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, 'tests', 'golden', 'configs', name)
