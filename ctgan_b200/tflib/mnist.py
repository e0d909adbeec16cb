"""MNIST batch generators with the reference's semantics (TG/tflib/mnist.py:8-104), Python 3.

`load(batch_size, test_batch_size, n_examples=60000, n_labelled=None)` reads `/tmp/mnist.pkl.gz` (the
deeplearning.net pickle: three `(images float32 [n,784] in [0,1], targets int64 [n])` pairs) and returns
`(train_epoch, dev_epoch, test_epoch)`.  Semantics the CT script relies on (TG/CT_gan_mnist.py:219-224):
  * the training generator keeps the first `n_examples` rows (:51-52) and shuffles them once at construction and once
    more per epoch, images and targets by the same permutation from numpy's GLOBAL RandomState (:53-56, :66-69);
  * an epoch is `images.reshape(-1, batch_size, 784)`: the set size must be a multiple of the batch size (:75-76);
  * batches are COPIES (`numpy.copy`, :83-88);
  * with `n_labelled` a third array is yielded -- the reference yields the WHOLE shuffled `labelled` vector with every
    batch (:83), not its batch slice; kept as is.
The reference downloads the file when it is missing (:92-97); so does this module (it fails without a network).
"""
import gzip
import os
import pickle

import numpy

FILEPATH = '/tmp/mnist.pkl.gz'
URL = 'http://www.iro.umontreal.ca/~lisa/deep/data/mnist/mnist.pkl.gz'


def _shuffle_together(*arrays):
    state = numpy.random.get_state()
    for i, a in enumerate(arrays):
        if i:
            numpy.random.set_state(state)
        numpy.random.shuffle(a)


def _generator(images, targets, batch_size, n_labelled, limit):
    _shuffle_together(images, targets)
    if limit is not None:
        print("WARNING ONLY FIRST {} MNIST DIGITS".format(limit))
        images = images.astype('float32')[:limit]
        targets = targets.astype('int32')[:limit]
    labelled = None
    if n_labelled is not None:
        labelled = numpy.zeros(len(images), dtype='int32')
        labelled[:n_labelled] = 1

    def get_epoch():
        if labelled is None:
            _shuffle_together(images, targets)
        else:
            _shuffle_together(images, targets, labelled)
        image_batches = images.reshape(-1, batch_size, 784)
        target_batches = targets.reshape(-1, batch_size)
        for i in range(len(image_batches)):
            if labelled is None:
                yield numpy.copy(image_batches[i]), numpy.copy(target_batches[i])
            else:
                yield numpy.copy(image_batches[i]), numpy.copy(target_batches[i]), numpy.copy(labelled)
    return get_epoch


def mnist_generator(data, batch_size, n_labelled, limit=None):
    images, targets = data
    return _generator(images, targets, batch_size, n_labelled, limit)


def mnist_generator2(data, batch_size, n_labelled, n_examples, limit=None):
    images, targets = data
    return _generator(images[0:n_examples, :], targets[0:n_examples], batch_size, n_labelled, limit)


def load(batch_size, test_batch_size, n_examples=60000, n_labelled=None, filepath=None):
    filepath = filepath or FILEPATH
    if not os.path.isfile(filepath):
        print("Couldn't find MNIST dataset in /tmp, downloading...")
        import urllib.request
        urllib.request.urlretrieve(URL, filepath)
    with gzip.open(filepath, 'rb') as f:
        train_data, dev_data, test_data = pickle.load(f, encoding='latin1')
    return (mnist_generator2(train_data, batch_size, n_labelled, n_examples),
            mnist_generator(dev_data, test_batch_size, n_labelled),
            mnist_generator(test_data, test_batch_size, n_labelled))
