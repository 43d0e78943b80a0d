"""world_size-2 gloo tests (CPU) of the multi-rank host logic: rank-major batch
sharding, the flat gradient bucket and its single allreduce, and that a
2-rank sharded RecNet step equals the 1-rank step on the whole batch.
The DC layers are the CPU oracle here (the product DC op is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import dc_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_batch(B, n, seed=0):
    rs = np.random.RandomState(seed)
    img = rs.uniform(0, 1, (B, n, n))
    m1 = orc.cartesian_mask((B, n, n), 4, 8, False, np.random.RandomState(seed))
    xu, xfu = orc.undersample(img, m1, rng=np.random.RandomState(seed))
    return {'inp': torch.from_numpy(orc.to_tensor_format(xu)),
            'kspace': torch.from_numpy(orc.to_tensor_format(xfu)),
            'mask': torch.from_numpy(orc.to_tensor_format(m1, mask=True)),
            'target': torch.from_numpy(orc.to_tensor_format(img))}


def _build(seed=0):
    from csmri_refinement_b200 import recnet
    torch.manual_seed(seed)
    return recnet.construct_model({'num_blocks': 2, 'num_convs': 2, 'num_filters': 4},
                                  dc_factory=orc.OracleDataConsistencyInKspace)


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from csmri_refinement_b200 import parallel
    r, w, dev = parallel.init_distributed('gloo')
    assert (r, w) == (rank, world) and dev.type == 'cpu'
    model = _build()
    trainer = parallel.ShardedTrainer(model, lr=1e-3)
    assert trainer.bucket.world == world
    full = _make_batch(4, 32)
    shard = parallel.shard_batch(full, rank, world)
    assert shard['inp'].shape[0] == 2
    losses = [trainer.step(shard).item()]
    grad1 = trainer.bucket.flat.clone()          # allreduced gradient of step 1
    losses.append(trainer.step(shard).item())
    assert trainer.bucket.check_views()
    torch.save({'sd': model.state_dict(), 'losses': losses, 'grad1': grad1,
                'grad': trainer.bucket.flat.clone()}, os.path.join(out_dir, 'r%d.pt' % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_range_is_contiguous_rank_major():
    from csmri_refinement_b200 import parallel
    for n, world in ((256, 8), (10, 4), (3, 8), (7, 2)):
        spans = [parallel.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_flat_bucket_views_survive_backward():
    from csmri_refinement_b200 import parallel
    model = _build()
    bucket = parallel.FlatGradBucket(model.parameters())
    assert bucket.flat.numel() == sum(p.numel() for p in model.parameters())
    b = _make_batch(2, 32)
    out = model(b['inp'], b['kspace'], b['mask'])
    torch.nn.functional.mse_loss(out, b['target']).backward()
    assert bucket.check_views()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.equal(ref, bucket.flat) and bucket.flat.abs().sum() > 0
    bucket.zero_()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())


@pytest.mark.timeout(300)
def test_two_rank_step_equals_single_rank_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), 'r0.pt'))
    r1 = torch.load(os.path.join(str(tmp_path), 'r1.pt'))
    # identical replicas after the allreduce'd updates
    for k in r0['sd']:
        assert torch.equal(r0['sd'][k], r1['sd'][k]), k
    assert torch.equal(r0['grad'], r1['grad'])
    # and equal to one process stepping on the whole batch
    from csmri_refinement_b200 import parallel
    model = _build()
    trainer = parallel.ShardedTrainer(model, lr=1e-3)
    full = _make_batch(4, 32)
    losses = [trainer.step(full).item()]
    # the allreduced mean of the shard gradients IS the full-batch gradient
    assert orc.rel_l2(r0['grad1'].numpy(), trainer.bucket.flat.numpy()) < 1e-5
    losses.append(trainer.step(full).item())
    # Adam divides by sqrt(v): parameters whose true gradient is ~0 (the bias in
    # front of a DC layer) move by O(lr) in a rounding-noise direction, so the
    # weights are compared with an absolute tolerance well below lr = 1e-3
    gmax = max(float(p.grad.abs().max()) for p in trainer.bucket.params)
    for (k, v), p in zip(model.named_parameters(), trainer.bucket.params):
        noise_only = float(p.grad.abs().max()) < 1e-5 * gmax
        tol = 2.5e-3 if noise_only else 2e-5          # 2 steps * lr for noise-driven ones
        assert (r0['sd'][k] - v.detach()).abs().max().item() < tol, k
    # mean of the two shard losses == full-batch loss (equal shard sizes)
    for i in range(2):
        assert abs(0.5 * (r0['losses'][i] + r1['losses'][i]) - losses[i]) < 1e-6


def _uneven_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from csmri_refinement_b200 import parallel
    parallel.init_distributed('gloo')
    model = _build()
    trainer = parallel.ShardedTrainer(model, lr=1e-3)
    full = _make_batch(3, 32)
    shard = parallel.shard_batch(full, rank, world)
    assert shard['inp'].shape[0] == (2 if rank == 0 else 1)
    trainer.step(shard)
    torch.save({'grad1': trainer.bucket.flat.clone(), 'w': trainer._loss_weight},
               os.path.join(out_dir, 'u%d.pt' % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_uneven_shards_give_the_global_batch_gradient(tmp_path):
    """3 slices over 2 ranks (shards of 2 and 1).  Each rank weights
    its mean loss by local_n * world / global_n, so the allreduced mean is the
    gradient of the mean loss over all 3 slices (what the reference's
    single-process DataParallel step computes), not the mean of shard means."""
    world, port = 2, _free_port()
    mp.spawn(_uneven_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), 'u0.pt'))
    r1 = torch.load(os.path.join(str(tmp_path), 'u1.pt'))
    assert abs(r0['w'] - 4.0 / 3.0) < 1e-12 and abs(r1['w'] - 2.0 / 3.0) < 1e-12
    assert torch.equal(r0['grad1'], r1['grad1'])
    from csmri_refinement_b200 import parallel
    model = _build()
    trainer = parallel.ShardedTrainer(model, lr=1e-3)
    trainer.step(_make_batch(3, 32))
    assert orc.rel_l2(r0['grad1'].numpy(), trainer.bucket.flat.numpy()) < 1e-5


def test_harness_builds_the_1recnet_step_on_cpu():
    """configs/1-recnet.json -> model, optimizer arguments and this rank's share of
    the GLOBAL batch 20 (strong scaling: 8 ranks get 3,3,3,3,2,2,2,2)."""
    from csmri_refinement_b200 import harness, parallel
    conf = harness.load_config(harness.config_path('1-recnet.json'))
    sizes = []
    for r in range(8):
        lo, hi = parallel.shard_range(conf.batch_size, r, 8)
        sizes.append(hi - lo)
    assert sizes == [3, 3, 3, 3, 2, 2, 2, 2]
    trainer, local_b = harness.recnet_trainer(conf, torch.device('cpu'), rank=5, world=8)
    assert local_b == 2 and not trainer.cuda_graph
    opt = trainer.optimizer.param_groups[0]
    assert opt['lr'] == 2e-4 and tuple(opt['betas']) == (0.9, 0.999)
    assert trainer.bucket.nbytes() == 31302 * 4


def _small_refinement_conf(tmp_dir):
    """configs/2-refinement.json with the widths cut down so that one adversarial
    step runs on the CPU in seconds (same keys, same code path)."""
    import json
    from csmri_refinement_b200 import harness
    with open(harness.config_path('2-refinement.json')) as f:
        c = json.load(f)
    c['generator_model']['learnable_model'].update(encode_filters=[4, 8, 16], decode_filters=[8, 4])
    c['generator_model']['pretrained_model'].update(num_filters=4)
    c['discriminator_model'].update(num_filters_per_layer=[4, 8, 8, 16, 16, 16],
                                    spatial_shape=[128, 128], image_pool_size=3)
    c['batch_size'] = 2
    path = os.path.join(tmp_dir, 'small-refinement.json')
    with open(path, 'w') as f:
        json.dump(c, f)
    return harness.load_config(path)


def _refine_worker(rank, world, port, out_dir, overlap):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from csmri_refinement_b200 import parallel, refinement_harness as rh
    if world > 1:
        parallel.init_distributed('gloo')
    conf = _small_refinement_conf(out_dir if rank == 0 else os.path.join(out_dir, 'r1'))
    trainer = rh.AdversarialTrainer(conf, torch.device('cpu'), rank, overlap=overlap,
                                    dc_factory=orc.OracleDataConsistencyInKspace,
                                    chunk_bytes=2048)
    assert len(trainer.disc_bucket.chunks) > 3
    res = []
    for i in range(3):                       # third step draws from the (size 3) image pool
        out = trainer.step(_make_batch(2, 128, seed=10 * rank + i))
        res.append({k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v])
                    for k, v in out.items()})
    torch.save({'gen': [p.detach().clone() for p in trainer.gen_bucket.params],
                'disc': [p.detach().clone() for p in trainer.disc_bucket.params],
                'disc_grad': trainer.disc_bucket.flat.clone(), 'res': res,
                'gen_keys': list(trainer.gen.state_dict().keys())},
               os.path.join(out_dir, 'ref_%d_%d_%d.pt' % (world, int(overlap), rank)))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_adversarial_step_two_ranks_overlapped_allreduce(tmp_path):
    """BASELINE configs[4] host logic on 2 gloo ranks: the discriminator / generator
    gradient buckets are all-reduced chunk by chunk from backward hooks; both
    ranks end with identical weights, and the overlapped exchange equals the
    single all-reduce after backward."""
    os.makedirs(os.path.join(str(tmp_path), 'r1'), exist_ok=True)
    for overlap in (True, False):
        mp.spawn(_refine_worker, args=(2, _free_port(), str(tmp_path), overlap), nprocs=2,
                 join=True)
    load = lambda w, o, r: torch.load(os.path.join(str(tmp_path), 'ref_%d_%d_%d.pt' % (w, o, r)))  # noqa: E731
    a0, a1, b0 = load(2, 1, 0), load(2, 1, 1), load(2, 0, 0)
    for key in ('gen', 'disc'):
        for p, q in zip(a0[key], a1[key]):
            assert torch.equal(p, q)                   # identical replicas
        for p, q in zip(a0[key], b0[key]):
            assert torch.allclose(p, q, rtol=0, atol=1e-6)   # overlapped == serial exchange
    assert torch.allclose(a0['disc_grad'], b0['disc_grad'], rtol=1e-5, atol=1e-8)
    assert a0['disc_grad'].abs().sum() > 0
    assert a0['gen_keys'][0] == 'scale' and a0['gen_keys'][1].startswith('pretrained_model.')
    r = a0['res'][-1]
    assert len(r['gen_losses']) == 4 and all(torch.isfinite(t) for t in r['gen_losses'])
    assert torch.isfinite(r['disc_loss']) and torch.isfinite(r['gen_loss'])
    # frozen RecNet: its parameters are not in the generator bucket (3 conv blocks x 3 convs x 2)
    from csmri_refinement_b200 import refinement_harness as rh
    conf = _small_refinement_conf(str(tmp_path))
    gen = rh.build_generator(conf, dc_factory=orc.OracleDataConsistencyInKspace)
    n_frozen = sum(1 for p in gen.pretrained_model.parameters())
    assert n_frozen == 18 and len(a0['gen']) == len(list(gen.parameters())) - n_frozen
