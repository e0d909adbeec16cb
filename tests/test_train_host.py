"""Host logic of the training loop, the input feeder and the checkpoints (`-m "not gpu"`, kernels replaced by the
TEST-ONLY stand-ins of tests/fake_backend.py): the reference's iteration schedule (TG/CT_gan_cifar_resnet.py:393-434),
metric names, files written, and an exact save / load round trip."""
import gzip
import os
import pickle

import numpy as np
import pytest
import torch

from tests.test_host_utils import write_cifar_dir, mnist_sets


def test_device_feeder_ring_keeps_the_last_batches_valid():
    from ctgan_b200.data import DeviceFeeder, inf_train_gen

    def epoch():
        for i in range(7):
            yield np.full((4, 8), i, dtype='uint8'), np.full((4,), i, dtype='int64'), 'ignored'

    f = DeviceFeeder(inf_train_gen(epoch), 'cpu', depth=2, hold=3, take=2)
    held = []
    for k in range(20):
        x, y = next(f)
        assert x.dtype == torch.uint8 and y.dtype == torch.int32          # pixels stay bytes, labels become int32
        held = (held + [(k % 7, x, y)])[-3:]
        for v, a, b in held:
            assert int(a[0, 0]) == v and int(b[0]) == v
    assert f.bytes_per_batch == 4 * 8 + 4 * 4
    g = DeviceFeeder(inf_train_gen(epoch), 'cpu', keep_uint8=False, take=1)
    assert next(g)[0].dtype == torch.int32


@pytest.mark.parametrize('script', ['cifar', 'cifar_resnet', 'mnist'])
def test_train_loop_schedule_and_outputs(fake_kernels, tmp_path, script, capsys):
    from ctgan_b200 import train as T
    import ctgan_b200.tflib.plot as plot
    if script == 'mnist':
        data = str(tmp_path / 'mnist.pkl.gz')
        with gzip.open(data, 'wb') as f:
            pickle.dump(tuple(mnist_sets(n=(40, 8, 8))), f, protocol=2)
    else:
        data = write_cifar_dir(str(tmp_path / 'data'), n_per_file=8)
    out = str(tmp_path / 'out')
    sess = T.train(script, data, iters=3, dev_every=2, out_dir=out, dev_batches=1, batch_size=4, n_examples=40,
                   device='cpu', use_graphs=False, act_dtype=torch.float32, checkpoint_every=3, acc_every=2)
    try:
        # iteration 0 has no generator step (:396); every iteration runs N_CRITIC critic steps on fresh batches
        n_critic = sess.n_critic
        assert n_critic == 5
        assert sess.tr.disc_opt.t == 3 * n_critic and sess.tr.gen_opt.t == 2
        assert sess.feeder.batches == 3 * n_critic
        log = pickle.load(open(os.path.join(out, 'log.pkl'), 'rb'))
        cost_name = 'cost' if script == 'cifar_resnet' else 'train disc cost'
        dev_name = 'dev_cost' if script == 'cifar_resnet' else 'dev disc cost'
        assert sorted(log[cost_name]) == [0, 1, 2] and sorted(log['time']) == [0, 1, 2] and sorted(log[dev_name]) == [1]
        if script == 'cifar_resnet':
            assert sorted(log['wgan']) == [0, 1, 2] and sorted(log['acgan']) == [0, 1, 2]
            # 'acc_real' / 'acc_fake' (TG/CT_gan_cifar_resnet.py:410-411): the metrics-only clean pass, every acc_every iterations
            assert sorted(log['acc_real']) == [0, 2] and sorted(log['acc_fake']) == [0, 2]
            assert all(0.0 <= v <= 1.0 for v in list(log['acc_real'].values()) + list(log['acc_fake'].values()))
            # 'wgan' = disc_wgan (TG/CT_gan_cifar_resnet.py:295) = cost - ACGAN_SCALE * acgan (:300)
            assert all(abs(log['wgan'][i] - (log['cost'][i] - log['acgan'][i])) < 1e-4 * max(1., abs(log['cost'][i])) for i in range(3))
        assert all(np.isfinite(v) for v in log[cost_name].values())
        ext = {'cifar': 'jpg', 'cifar_resnet': 'png', 'mnist': 'png'}[script]
        assert os.path.getsize(os.path.join(out, 'samples_1.%s' % ext)) > 0
        assert os.path.exists(os.path.join(out, 'checkpoint.npz'))
        if script == 'cifar':
            para = np.load(os.path.join(out, 'param.pyn.npy'), allow_pickle=True)
            assert len(para) == 8 and para[0].shape == (5, 5, 3, 128)        # 3 convs + Output: Filters/W + Biases/b
        assert 'iter 0\t' in capsys.readouterr().out
        # --resume: the loop continues at the saved iteration (here 3) with the saved optimizer / random state
        blob = np.load(os.path.join(out, 'checkpoint.npz'))
        assert int(blob['iteration']) == 3
        sess2 = T.train(script, data, iters=4, dev_every=2, out_dir=out, dev_batches=1, batch_size=4, n_examples=40,
                        device='cpu', use_graphs=False, act_dtype=torch.float32, resume=os.path.join(out, 'checkpoint.npz'))
        assert sess2.iteration == 4 and sess2.tr.disc_opt.t == 4 * n_critic and sess2.tr.gen_opt.t == 3
    finally:
        plot.output_dir = '.'
        plot.reset()


def test_checkpoint_round_trip_resumes_exactly(fake_kernels, tmp_path):
    import ctgan_b200.gan_cifar as C
    import ctgan_b200.tflib as lib
    from ctgan_b200 import checkpoint
    x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (3, 4, 3072)).astype('int32'))

    def fresh():
        np.random.seed(5)
        return C.Trainer(device='cpu', seed=9, act_dtype=torch.float32, batch_size=4)

    a = fresh()
    a.critic_step(x[0]); a.gen_step()
    path = str(tmp_path / 'ck.npz')
    checkpoint.save(path, a, iteration=7)
    blob = np.load(path)
    assert blob['param/Discriminator.2.Filters'].shape == (5, 5, 128, 256)            # reference name, HWIO
    assert blob['param/Generator.Input.W'].shape == (128, 8192)
    assert a.rng.offset > 0 and int(blob['rng/offset']) == a.rng.offset          # the Philox position is part of the state
    np_next = np.random.RandomState()
    np_next.set_state(np.random.get_state())
    a.critic_step(x[1])
    want = a.disc_opt.flat_p.clone()

    b = fresh()
    np.random.seed(123)                                      # a resumed process starts with some other numpy state
    assert checkpoint.load(path, b) == {'iteration': 7}
    assert b.disc_opt.t == 1 and b.gen_opt.t == 1
    assert b.rng.offset == int(blob['rng/offset'])           # same noise / dropout stream position, no manual fix-up
    assert np.random.randint(1 << 30) == np_next.randint(1 << 30)                 # loaders' shuffles continue too
    b.critic_step(x[1])
    assert torch.equal(b.disc_opt.flat_p, want)
    with pytest.raises(KeyError):
        lib.param('Discriminator.Extra', np.zeros(2, dtype='float32'))
        checkpoint.load(path, None)


def test_train_loop_64x64(fake_kernels, tmp_path, capsys):
    """The CT_gan_64x64.py loop (row N4) over tflib.small_imagenet PNG folders: schedule, metric names, files."""
    from PIL import Image
    from ctgan_b200 import train as T
    import ctgan_b200.tflib.plot as plot
    import ctgan_b200.gan_64x64 as G
    rs = np.random.RandomState(0)
    for sub, n in (('train_64x64', 45), ('valid_64x64', 9)):
        d = tmp_path / 'data' / sub
        d.mkdir(parents=True)
        for i in range(1, n + 1):
            Image.fromarray(rs.randint(0, 256, (64, 64, 3)).astype('uint8'), 'RGB').save(str(d / ('%s.png' % str(i).zfill(len(str(n))))))
    out = str(tmp_path / 'out')
    try:
        sess = T.train('64x64', str(tmp_path / 'data'), iters=2, dev_every=2, out_dir=out, dev_batches=1, batch_size=4,
                       n_examples=(45, 9), device='cpu', use_graphs=False, act_dtype=torch.float32, model_kw=dict(dim=4))
        assert sess.tr.disc_opt.t == 2 * 5 and sess.tr.gen_opt.t == 1
        log = pickle.load(open(os.path.join(out, 'log.pkl'), 'rb'))
        assert sorted(log['train disc cost']) == [0, 1] and sorted(log['dev disc cost']) == [1]
        assert all(np.isfinite(v) for v in log['train disc cost'].values())
        assert os.path.getsize(os.path.join(out, 'samples_1.png')) > 0
    finally:
        plot.output_dir = '.'
        plot.reset()
        G.DIM = 64


def test_train_loop_lsun128(fake_kernels, tmp_path, capsys):
    """The LS/wgan_LSUN_Bedrooms128.py loop (row N4): critic steps first, a generator step in EVERY iteration, 'cost' / 'time'
    metrics, samples before the loop and at iteration % 100 == 0, batches from the image-folder loader (incl. its mirrored views)."""
    from PIL import Image
    from ctgan_b200 import train as T
    import ctgan_b200.tflib.plot as plot
    import ctgan_b200.gan_lsun128 as G
    rs = np.random.RandomState(0)
    d = tmp_path / 'lsun'
    d.mkdir()
    for i in range(60):
        Image.fromarray(rs.randint(0, 256, (128, 128, 3)).astype('uint8'), 'RGB').save(str(d / ('%03d.png' % i)))
    out = str(tmp_path / 'out')
    try:
        np.random.seed(3)
        sess = T.train('lsun128', str(d), iters=2, out_dir=out, batch_size=4, device='cpu', use_graphs=False,
                       act_dtype=torch.float32, model_kw=dict(width=1.0 / 32))
        assert sess.tr.disc_opt.t == 2 * 5 and sess.tr.gen_opt.t == 2            # LS :373-389: no skipped generator step
        log = pickle.load(open(os.path.join(out, 'log.pkl'), 'rb'))
        assert sorted(log['cost']) == [0] and sorted(log['time']) == [0] and sess.iteration == 2   # flushed at iteration % 5 == 0 (LS :398)
        assert all(np.isfinite(v) for v in log['cost'].values())
        assert os.path.getsize(os.path.join(out, 'samples_0.png')) > 0
        assert os.path.exists(os.path.join(out, 'checkpoint.npz'))
    finally:
        plot.output_dir = '.'
        plot.reset()
        G.WIDTH = 1.0
