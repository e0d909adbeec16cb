"""Data-parallel path on CPU (`-m "not gpu"`): two processes, gloo backend, world_size 2.
Kernel launchers are the TEST-ONLY torch stand-ins (tests/fake_backend.py); what is under test is
the product's host logic: flat gradient bucket -> all-reduce(sum) -> 1/world folded into Adam,
identical replicas after the update, rank-dependent random streams and data shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _MP:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _worker(rank, world, port, script, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests import fake_backend, parity
    import importlib
    fake_backend.install(_MP())
    mod = importlib.import_module(parity.SCRIPTS[script][0])
    np.random.seed(1234)                                        # same initial weights on every rank
    tr = mod.Trainer(device='cpu', seed=1234 + rank, act_dtype=torch.float32, batch_size=4)
    inputs = parity.make_inputs(script, 4, 50 + rank)           # a different data shard per rank
    # local gradients first (no exchange) ...
    tr.disc_opt.zero_grad()
    tr.critic_forward_backward(*inputs)
    local = tr.disc_opt.flat_g.clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    # ... then the product path: all-reduce + Adam with grad_scale 1/world
    p_before = tr.disc_opt.flat_p.clone()
    w = tr.disc_opt.all_reduce()
    assert w == world
    assert torch.allclose(tr.disc_opt.flat_g, sum(gathered), rtol=1e-6, atol=1e-8)
    tr.disc_opt.step(None, w)
    # reference update: Adam on the MEAN gradient
    g = sum(gathered) / world
    m = (1 - tr.disc_opt.beta1) * g
    v = (1 - tr.disc_opt.beta2) * g * g
    expect = p_before - tr.disc_opt.lr_t() * m / (v.sqrt() + 1e-8)
    assert torch.allclose(tr.disc_opt.flat_p, expect, rtol=1e-5, atol=1e-8)
    torch.save(dict(p=tr.disc_opt.flat_p, local=local), os.path.join(out_dir, 'rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('script', ['mnist', 'resnet'])
def test_two_rank_gloo_data_parallel(tmp_path, script):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, script, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / 'rank0.pt'), torch.load(tmp_path / 'rank1.pt')
    assert torch.equal(r0['p'], r1['p'])                        # replicas stay bit-identical
    assert not torch.equal(r0['local'], r1['local'])            # shards / random streams differ per rank


def _train_worker(rank, world, port, data_dir, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests import fake_backend
    fake_backend.install(_MP())
    from ctgan_b200 import train as T
    seen = []
    import ctgan_b200.data as D
    orig = D.DeviceFeeder.__next__

    def spy(self):
        out = orig(self)
        seen.append(out[0].clone())
        return out
    D.DeviceFeeder.__next__ = spy
    sess = T.train('cifar', data_dir, iters=2, dev_every=100, out_dir=os.path.join(out_dir, 'run'), batch_size=4, n_examples=80,
                   device='cpu', use_graphs=False, act_dtype=torch.float32)
    torch.save(dict(pd=sess.tr.disc_opt.flat_p, pg=sess.tr.gen_opt.flat_p, seen=torch.stack(seen)),
               os.path.join(out_dir, 'train_rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_loop_shards_the_epoch(tmp_path):
    """ctgan_b200.train under world_size 2: both ranks walk the same epoch order and take alternating batches
    (disjoint shards), gradients are all-reduced every step, replicas stay bit-identical, rank 0 writes the logs."""
    from tests.test_host_utils import write_cifar_dir
    data = write_cifar_dir(str(tmp_path / 'data'), n_per_file=16)
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, data, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / 'train_rank0.pt'), torch.load(tmp_path / 'train_rank1.pt')
    assert torch.equal(r0['pd'], r1['pd']) and torch.equal(r0['pg'], r1['pg'])
    assert r0['seen'].shape == r1['seen'].shape == (10, 4, 3072)
    rows0 = {bytes(r.numpy().tobytes()) for b in r0['seen'] for r in b}
    rows1 = {bytes(r.numpy().tobytes()) for b in r1['seen'] for r in b}
    assert not (rows0 & rows1)                                  # no image is seen by both ranks within the epoch
    assert os.path.exists(tmp_path / 'run' / 'log.pkl')
