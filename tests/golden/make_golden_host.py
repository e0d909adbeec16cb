"""Generates tests/golden/host_utils.npz by EXECUTING THE REFERENCE'S OWN host modules (TG/tflib/cifar10.py, mnist.py,
save_images.py through oracle/ref_harness.load_ref_host_module) on the synthetic datasets of tests/test_host_utils.py:

    python tests/golden/make_golden_host.py

Stored: per-batch fingerprints of three training epochs + one dev epoch of each loader, and the row sums of a sample grid."""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH            # noqa: E402
from tests import test_host_utils as T          # noqa: E402


def main():
    blob = {}
    with tempfile.TemporaryDirectory() as tmp:
        d = T.write_cifar_dir(os.path.join(tmp, 'c'), py2_keys=False)
        t, dv = T.cifar_sequence(RH.load_ref_host_module('cifar10'), d)
        blob['cifar_train'], blob['cifar_dev'] = T.digest(t), T.digest(dv)
    a, b, c = T.mnist_sequence(RH.load_ref_host_module('mnist'))
    blob['mnist_train'], blob['mnist_dev'], blob['mnist_labelled'] = T.digest(a), T.digest(b), T.digest(c)
    captured = {}
    misc = types.ModuleType('scipy.misc')
    misc.imsave = lambda path, img: captured.update(img=np.array(img))
    sp = types.ModuleType('scipy')
    sp.misc = misc
    ref = RH.load_ref_host_module('save_images', stubs={'scipy': sp, 'scipy.misc': misc})
    X = np.random.RandomState(3).randint(0, 256, (100, 3, 32, 32)).astype('int32')
    ref.save_images(X, 'unused.png')
    blob['grid_row_sums'] = captured['img'].sum(axis=(1, 2))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_utils.npz')
    np.savez_compressed(path, **blob)
    print({k: v.shape for k, v in blob.items()}, '->', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
