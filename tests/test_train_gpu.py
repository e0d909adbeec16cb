"""GPU tests of the pieces around the training step: the pinned / asynchronous uint8 input feeder, the training loop in
CUDA-graph mode (ctgan_b200.train) and the checkpoint round trip on the flat device buffers."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests.test_host_utils import write_cifar_dir

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


def test_feeder_delivers_the_generator_sequence_as_bytes(tmp_path):
    _need_gpu()
    import ctgan_b200.tflib.cifar10 as c10
    import ctgan_b200.kernels as K
    from ctgan_b200.data import DeviceFeeder, inf_train_gen
    d = write_cifar_dir(str(tmp_path / 'd'), n_per_file=16)
    np.random.seed(3)
    train, _ = c10.load(8, d, 64)
    want = []
    for _ in range(3):
        want += [(x.copy(), y.copy()) for x, y in train()]
    np.random.seed(3)
    train, _ = c10.load(8, d, 64)
    f = DeviceFeeder(inf_train_gen(train), 'cuda', depth=2, take=2, hold=5)
    held = []
    for k in range(len(want) - 3):
        x, y = next(f)
        assert x.dtype == torch.uint8 and x.is_cuda and y.dtype == torch.int32
        held = (held + [(k, x, y)])[-5:]
        # a consumer kernel on the current stream between next() calls (what a critic step does)
        prep = K.prep_real(x, 256., 0.)
        for j, a, b in held:
            assert np.array_equal(a.cpu().numpy(), want[j][0]) and np.array_equal(b.cpu().numpy(), want[j][1].astype('int32'))
        ref = 2 * (torch.from_numpy(want[k][0].astype('float32')) / 256. - 0.5)
        assert torch.equal(prep.cpu(), ref)
    assert f.bytes_per_batch == 8 * 3072 + 8 * 4


@pytest.mark.parametrize('script', ['cifar_resnet', 'cifar'])
def test_training_loop_in_graph_mode(tmp_path, script, capsys):
    """3 iterations of the reference's schedule with CUDA-graph replays, the batched fake generation (ResNet), the
    feeder's uint8 batches as graph inputs, asynchronous metric fetches, dev cost, sample grid and checkpoint."""
    _need_gpu()
    from ctgan_b200 import train as T, checkpoint
    import ctgan_b200.tflib.plot as plot
    d = write_cifar_dir(str(tmp_path / 'd'), n_per_file=32)
    out = str(tmp_path / 'out')
    try:
        sess = T.train(script, d, iters=3, dev_every=2, out_dir=out, dev_batches=1, batch_size=16, n_examples=160,
                       checkpoint_every=3, acc_every=2)
        torch.cuda.synchronize()
        assert sess.gt is not None and sess.tr.disc_opt.t == 15 and sess.tr.gen_opt.t == 2     # the capture's warm-up steps are undone
        log = pickle.load(open(os.path.join(out, 'log.pkl'), 'rb'))
        name = 'cost' if script == 'cifar_resnet' else 'train disc cost'
        assert sorted(log[name]) == [0, 1, 2] and all(np.isfinite(v) for v in log[name].values())
        if script == 'cifar_resnet':           # the metrics-only clean pass between graph replays (fakes of the last critic step)
            assert sorted(log['acc_real']) == [0, 2] and all(0.0 <= v <= 1.0 for v in log['acc_fake'].values())
        dev = log['dev_cost' if script == 'cifar_resnet' else 'dev disc cost']
        assert sorted(dev) == [1] and np.isfinite(dev[1])
        assert os.path.getsize(os.path.join(out, 'samples_1.%s' % ('png' if script == 'cifar_resnet' else 'jpg'))) > 0
        # checkpoint round trip on the flat device buffers
        before = sess.tr.disc_opt.flat_p.clone()
        sess.tr.disc_opt.flat_p.zero_()
        checkpoint.load(os.path.join(out, 'checkpoint.npz'), sess.tr)
        assert torch.equal(sess.tr.disc_opt.flat_p, before)
        assert 'iter 0\t' in capsys.readouterr().out
    finally:
        plot.output_dir = '.'
        plot.reset()
