"""Injected randomness for the oracle (TEST INFRASTRUCTURE).

The reference draws z, dropout masks, interpolation alphas, dequantisation noise and
fake labels inside the TF graph (tf.random_normal / random_uniform / nn.dropout —
TG/CT_gan_cifar_resnet.py:157,202,277-281,319).  For parity the oracle takes every
random tensor from a source keyed by a *tag*; the product records its own device-side
draws under the same tags and the oracle replays them (`ReplayRandom`).
"""
import numpy as np
import torch


class SeededRandom:
    """Draws from numpy RandomState(seed), in call order; remembers every draw."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)
        self.tape = {}

    def _keep(self, tag, arr):
        assert tag not in self.tape, "duplicate random tag %s" % tag
        t = torch.from_numpy(arr)
        self.tape[tag] = t
        return t

    def normal(self, tag, shape):
        return self._keep(tag, self.rs.standard_normal(size=shape).astype('float32'))

    def uniform(self, tag, shape, lo=0., hi=1.):
        return self._keep(tag, (lo + (hi - lo) * self.rs.random_sample(size=shape)).astype('float32'))

    def labels(self, tag, n, n_labels=10):
        # tf.cast(tf.random_uniform([n])*10, tf.int32) -- TG/CT_gan_cifar_resnet.py:319
        u = self.rs.random_sample(size=(n,)).astype('float32')
        return self._keep(tag, (u * n_labels).astype('int32'))


class ReplayRandom:
    """Replays a tape {tag: tensor} recorded elsewhere (e.g. by the CUDA path)."""

    def __init__(self, tape, patterns=None):
        # optional activation patterns recorded on the device (see ct_gan_common.StepMixin._pattern)
        self.patterns = [p.detach().cpu().bool() for p in patterns] if patterns is not None else None
        self.tape = {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else torch.as_tensor(v))
                     for k, v in tape.items()}
        self.used = set()

    def _get(self, tag, shape):
        t = self.tape[tag]
        self.used.add(tag)
        assert tuple(t.shape) == tuple(shape), (tag, tuple(t.shape), tuple(shape))
        return t

    def normal(self, tag, shape):
        return self._get(tag, shape).float()

    def uniform(self, tag, shape, lo=0., hi=1.):
        return self._get(tag, shape).float()

    def labels(self, tag, n, n_labels=10):
        return self._get(tag, (n,)).to(torch.int32)
