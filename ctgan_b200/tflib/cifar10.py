"""CIFAR-10 batch generators with the reference's semantics (TG/tflib/cifar10.py:8-70), Python 3.

`load(batch_size, data_dir, n_examples)` returns `(train_epoch, dev_epoch)`: calling one yields the
`(images uint8 [batch_size, 3072], labels [batch_size])` batches of ONE epoch.  What the CT scripts rely on
(TG/CT_gan_cifar_resnet.py:362-366, TG/CT_gan_cifar.py:179-183):
  * the training set is the first `n_examples` rows of data_batch_1..5 concatenated (:50-51); the dev set is the
    whole test_batch;
  * every epoch shuffles images and labels IN PLACE with the same permutation, drawn from numpy's GLOBAL
    RandomState (state saved before the first shuffle and restored before the second, :28-31), so the n-th epoch is a
    permutation of the (n-1)-th;
  * `len(images) // batch_size` whole batches per epoch, the ragged tail is dropped (:33).
The files are the python pickles of the CIFAR-10 distribution (written by Python 2: `bytes` keys under Python 3).
Pixels stay uint8 end to end: `ctgan_b200.data.DeviceFeeder` moves them to the GPU as bytes and the input scaling
(TG/CT_gan_cifar_resnet.py:201-202) runs on the device (`ctgan_prep_real_u8`).
"""
import os
import pickle

import numpy as np

TRAIN_FILES = ['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5']
TEST_FILES = ['test_batch']


def unpickle(file):
    with open(file, 'rb') as fo:
        d = pickle.load(fo, encoding='bytes')
    get = lambda k: d[k] if k in d else d[k.encode()]
    return get('data'), get('labels')


def _read(filenames, data_dir):
    parts = [unpickle(os.path.join(data_dir, f)) for f in filenames]
    images = np.concatenate([p[0] for p in parts], axis=0)
    labels = np.concatenate([p[1] for p in parts], axis=0)
    return images, labels


def _epochs(images, labels, batch_size):
    def get_epoch():
        state = np.random.get_state()          # one permutation for both arrays
        np.random.shuffle(images)
        np.random.set_state(state)
        np.random.shuffle(labels)
        for i in range(len(images) // batch_size):
            sl = slice(i * batch_size, (i + 1) * batch_size)
            yield images[sl], labels[sl]
    return get_epoch


def cifar_generator(filenames, batch_size, data_dir):
    images, labels = _read(filenames, data_dir)
    return _epochs(images, labels, batch_size)


def cifar_generator2(filenames, batch_size, data_dir, n_examples):
    images, labels = _read(filenames, data_dir)
    return _epochs(images[0:n_examples, :], labels[0:n_examples], batch_size)


def load(batch_size, data_dir, n_examples):
    return (cifar_generator2(TRAIN_FILES, batch_size, data_dir, n_examples),
            cifar_generator(TEST_FILES, batch_size, data_dir))
