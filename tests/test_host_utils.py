"""Host-side utilities around the training step ("next" rows N1 / N3 of SURVEY.md 8(f)): the CIFAR-10 / MNIST batch
generators, the metric logger and the sample-grid writer, each compared with the REFERENCE'S OWN module executed here
(oracle/ref_harness.load_ref_host_module: py2 -> py3 translation of TG/tflib/{cifar10,mnist,plot,save_images}.py) on
synthetic dataset files, and with committed fixtures (tests/golden/host_utils.npz, made by tests/golden/make_golden_host.py
from those reference runs) so that the same checks run where /root/reference does not exist."""
import os
import pickle
import types

import numpy as np
import pytest

from oracle import ref_harness as RH

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'host_utils.npz')


def write_cifar_dir(path, n_per_file=40, seed=0, py2_keys=True):
    """A directory with the CIFAR-10 python-pickle layout: data_batch_1..5 + test_batch, `n_per_file` images each.
    py2_keys: pickled as Python 2 wrote them (bytes keys) -- what the distributed files look like under Python 3."""
    rs = np.random.RandomState(seed)
    os.makedirs(path, exist_ok=True)
    for name in ['data_batch_%d' % i for i in range(1, 6)] + ['test_batch']:
        data = rs.randint(0, 256, (n_per_file, 3072)).astype('uint8')
        labels = [int(v) for v in rs.randint(0, 10, n_per_file)]
        d = {b'data': data, b'labels': labels} if py2_keys else {'data': data, 'labels': labels}
        with open(os.path.join(path, name), 'wb') as f:
            pickle.dump(d, f, protocol=2)
    return path


def mnist_sets(seed=0, n=(120, 40, 40)):
    rs = np.random.RandomState(seed)
    return [(rs.random_sample((k, 784)).astype('float32'), rs.randint(0, 10, k).astype('int64')) for k in n]


def run_epochs(epoch_fn, n_epochs):
    out = []
    for _ in range(n_epochs):
        for batch in epoch_fn():
            out.append(tuple(np.array(a) for a in batch))
    return out


def cifar_sequence(mod, data_dir, batch_size=16, n_examples=100, seed=5, epochs=3):
    np.random.seed(seed)
    train, dev = mod.load(batch_size, data_dir, n_examples)
    return run_epochs(train, epochs), run_epochs(dev, 1)


def mnist_sequence(mod, seed=7, epochs=3):
    sets = mnist_sets()
    np.random.seed(seed)
    train = mod.mnist_generator2(sets[0], 20, None, 100)
    dev = mod.mnist_generator(sets[1], 10, None)
    lab = mod.mnist_generator2(mnist_sets()[0], 20, 30, 60)
    return run_epochs(train, epochs), run_epochs(dev, 1), run_epochs(lab, 1)


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert len(x) == len(y)
        for u, v in zip(x, y):
            assert u.shape == v.shape and u.dtype == v.dtype and np.array_equal(u, v)


def digest(seq):
    """Per-batch fingerprints: (first-column checksum of each array, label vector)."""
    return np.array([[float(np.asarray(a, dtype='float64').sum()) for a in b] for b in seq])


needs_ref = pytest.mark.skipif(not RH.available(), reason='/root/reference not present')


@needs_ref
def test_cifar10_generators_match_reference_module(tmp_path):
    import ctgan_b200.tflib.cifar10 as mine
    ref = RH.load_ref_host_module('cifar10')
    d_py2 = write_cifar_dir(str(tmp_path / 'py2'), py2_keys=True)
    d_str = write_cifar_dir(str(tmp_path / 'str'), py2_keys=False)      # what the translated reference can read
    rt, rd = cifar_sequence(ref, d_str)
    mt, md = cifar_sequence(mine, d_py2)
    assert len(mt) == 3 * (100 // 16) and len(md) == 40 // 16
    same(mt, rt)
    same(md, rd)
    assert mt[0][0].dtype == np.uint8 and mt[0][0].shape == (16, 3072)
    same(cifar_sequence(mine, d_str)[0], rt)                            # str-keyed pickles are accepted as well


@needs_ref
def test_mnist_generators_match_reference_module():
    import ctgan_b200.tflib.mnist as mine
    ref = RH.load_ref_host_module('mnist')
    for a, b in zip(mnist_sequence(mine), mnist_sequence(ref)):
        same(a, b)


def test_generators_match_committed_fixture(tmp_path):
    import ctgan_b200.tflib.cifar10 as c10
    import ctgan_b200.tflib.mnist as mn
    gold = np.load(GOLDEN)
    d = write_cifar_dir(str(tmp_path / 'c'), py2_keys=True)
    t, dv = cifar_sequence(c10, d)
    np.testing.assert_array_equal(digest(t), gold['cifar_train'])
    np.testing.assert_array_equal(digest(dv), gold['cifar_dev'])
    a, b, c = mnist_sequence(mn)
    np.testing.assert_array_equal(digest(a), gold['mnist_train'])
    np.testing.assert_array_equal(digest(b), gold['mnist_dev'])
    np.testing.assert_array_equal(digest(c), gold['mnist_labelled'])


def test_mnist_load_reads_the_pickle(tmp_path):
    import gzip
    import ctgan_b200.tflib.mnist as mn
    path = str(tmp_path / 'mnist.pkl.gz')
    with gzip.open(path, 'wb') as f:
        pickle.dump(tuple(mnist_sets()), f, protocol=2)
    np.random.seed(1)
    train, dev, test = mn.load(20, 10, n_examples=100, filepath=path)
    batches = run_epochs(train, 1)
    assert len(batches) == 5 and batches[0][0].shape == (20, 784) and batches[0][0].dtype == np.float32
    assert len(run_epochs(dev, 1)) == 4 and len(run_epochs(test, 1)) == 4


# ----------------------------------------------------------------------------- plot / save_images
def _plot_session(mod):
    out = []
    for it in range(7):
        mod.plot('train disc cost', 1.0 / (it + 1))
        if it % 2 == 0:
            mod.plot('time', 0.25 * it)
        if it in (2, 6):
            mod.flush()
        mod.tick()
    return out


@needs_ref
def test_plot_matches_reference_module(tmp_path, capsys, monkeypatch):
    import ctgan_b200.tflib.plot as mine
    plt = types.ModuleType('matplotlib.pyplot')
    for fn in ('clf', 'plot', 'xlabel', 'ylabel', 'savefig'):
        setattr(plt, fn, lambda *a, **k: None)
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    mpl.pyplot = plt
    ref = RH.load_ref_host_module('plot', stubs={'matplotlib': mpl, 'matplotlib.pyplot': plt})
    rdir, mdir = tmp_path / 'ref', tmp_path / 'mine'
    rdir.mkdir(); mdir.mkdir()
    monkeypatch.chdir(rdir)
    _plot_session(ref)
    ref_out = capsys.readouterr().out
    mine.reset()
    mine.output_dir = str(mdir)
    try:
        _plot_session(mine)
    finally:
        mine.output_dir = '.'
    assert capsys.readouterr().out == ref_out
    with open(rdir / 'log.pkl', 'rb') as f:
        ref_log = pickle.load(f)
    with open(mdir / 'log.pkl', 'rb') as f:
        my_log = pickle.load(f)
    assert my_log == ref_log and set(my_log) == {'train disc cost', 'time'}
    assert (mdir / 'train_disc_cost.jpg').stat().st_size > 0 and (mdir / 'time.jpg').stat().st_size > 0


@needs_ref
@pytest.mark.parametrize('shape,kind', [((100, 3, 32, 32), 'int'), ((128, 3, 32, 32), 'float'), ((12, 784), 'float'),
                                         ((6, 28, 28), 'float')])
def test_save_images_grid_matches_reference_module(tmp_path, shape, kind):
    import ctgan_b200.tflib.save_images as mine
    captured = {}
    misc = types.ModuleType('scipy.misc')
    misc.imsave = lambda path, img: captured.update(path=path, img=np.array(img))
    sp = types.ModuleType('scipy')
    sp.misc = misc
    ref = RH.load_ref_host_module('save_images', stubs={'scipy': sp, 'scipy.misc': misc})
    rs = np.random.RandomState(3)
    X = rs.randint(0, 256, shape).astype('int32') if kind == 'int' else rs.random_sample(shape).astype('float32')
    ref.save_images(X.copy(), 'unused.png')
    grid = mine.tile_images(X.copy())
    assert grid.shape == captured['img'].shape and np.array_equal(grid, captured['img'])
    path = str(tmp_path / 'samples_0.png')
    mine.save_images(X, path)
    from PIL import Image
    im = np.asarray(Image.open(path))
    assert im.shape == grid.shape and im.dtype == np.uint8
    assert np.array_equal(im, mine._to_uint8(grid))


def test_save_images_fixture():
    import ctgan_b200.tflib.save_images as mine
    gold = np.load(GOLDEN)
    rs = np.random.RandomState(3)
    X = rs.randint(0, 256, (100, 3, 32, 32)).astype('int32')
    grid = mine.tile_images(X)
    assert grid.shape == (320, 320, 3)
    np.testing.assert_array_equal(grid.sum(axis=(1, 2)), gold['grid_row_sums'])


@needs_ref
def test_small_imagenet_generator_matches_reference_module(tmp_path):
    """TG/tflib/small_imagenet.py, including its round-robin buffer quirk (a batch is yielded right after slot 0 was
    overwritten by the next batch's first file)."""
    from PIL import Image
    import ctgan_b200.tflib.small_imagenet as mine
    d = tmp_path / 'train_64x64'
    d.mkdir()
    n_files, bs = 23, 4
    rs = np.random.RandomState(9)
    for i in range(1, n_files + 1):
        Image.fromarray(rs.randint(0, 256, (64, 64, 3)).astype('uint8'), 'RGB').save(str(d / ('%s.png' % str(i).zfill(2))))
    misc = types.ModuleType('scipy.misc')
    misc.imread = lambda p: np.asarray(Image.open(p).convert('RGB'))
    sp = types.ModuleType('scipy')
    sp.misc = misc
    ref = RH.load_ref_host_module('small_imagenet', stubs={'scipy': sp, 'scipy.misc': misc})
    # the reference shuffles a py2 `range` list in place: under py3 the translated module needs a list
    def run(mod):
        gen = mod.make_generator(str(d), n_files, bs)
        return [b[0].copy() for _ in range(2) for b in gen()]
    got, want = run(mine), run(ref)
    assert len(got) == len(want) == 2 * ((n_files - 1) // bs)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype == np.int32 and a.shape == (bs, 3, 64, 64) and np.array_equal(a, b)


@needs_ref
def test_lsun_folder_generator_matches_reference_module(tmp_path):
    """LS/tflib/imagenet.py (the loader of LS/wgan_LSUN_Bedrooms128.py): compounding in-place shuffles of the file list,
    the round-robin buffer yielded right after slot 0 was overwritten, greyscale broadcast, wrong-size images skipped with
    their slot left stale, and the horizontal flips that compound because the reference re-binds its buffer to a reversed view."""
    from PIL import Image
    import ctgan_b200.tflib.imagenet as mine
    d = tmp_path / 'lsun'
    d.mkdir()
    rs = np.random.RandomState(4)
    for i in range(19):
        if i == 5:
            img = Image.fromarray(rs.randint(0, 256, (128, 128)).astype('uint8'), 'L')            # greyscale
        elif i == 11:
            img = Image.fromarray(rs.randint(0, 256, (64, 128, 3)).astype('uint8'), 'RGB')        # wrong size: skipped
        else:
            img = Image.fromarray(rs.randint(0, 256, (128, 128, 3)).astype('uint8'), 'RGB')
        img.save(str(d / ('img_%02d.png' % i)))
    misc = types.ModuleType('scipy.misc')
    sp = types.ModuleType('scipy')
    sp.misc = misc
    pil = types.ModuleType('Image')
    pil.open = Image.open
    ref = RH.load_ref_host_module('imagenet', stubs={'scipy': sp, 'scipy.misc': misc, 'Image': pil}, fork='LSUN_bedrooms')

    def run(mod):
        np.random.seed(77)                                   # the flips use numpy's global RandomState
        gen = mod.make_generator(str(d), 4)
        return [np.array(b[0]) for _ in range(3) for b in gen()]
    got, want = run(mine), run(ref)
    assert len(got) == len(want) == 3 * 4
    for a, b in zip(got, want):
        assert a.dtype == b.dtype == np.int32 and a.shape == (4, 3, 128, 128) and np.array_equal(a, b)
    assert any(not np.array_equal(x, y) for x, y in zip(got[:4], got[4:8]))                       # epochs differ


@needs_ref
@pytest.mark.parametrize('script,ref_file', [('mnist', 'CT_gan_mnist.py'), ('cifar', 'CT_gan_cifar.py'),
                                             ('cifar_resnet', 'CT_gan_cifar_resnet.py'), ('64x64', 'CT_gan_64x64.py')])
def test_training_loop_metric_names_are_the_reference_scripts(script, ref_file):
    """Every metric name ctgan_b200.train logs for a script is one the reference script logs with lib.plot.plot(...)."""
    import inspect
    import re
    from ctgan_b200 import train as T
    with open(os.path.join(RH.REF_ROOT, ref_file)) as f:
        ref_names = set(re.findall(r"lib\.plot\.plot\('([^']+)'", f.read()))
    src = inspect.getsource(T)
    ours = set(re.findall(r"_plot\.plot\('([^']+)'", src))
    for m in re.findall(r"_plot\.plot\('([^']+)' if s\.resnet else '([^']+)'", src):
        ours.update(m)
    resnet_only = {'cost', 'wgan', 'acgan', 'dev_cost', 'acc_real', 'acc_fake'}
    mine = {n for n in ours if (n in resnet_only) == (script == 'cifar_resnet') or n == 'time'}
    assert mine and mine <= ref_names, (sorted(mine - ref_names), sorted(ref_names))
