"""A stand-in for the ~45 `tf.*` symbols the reference's training step touches, evaluated eagerly on
PyTorch-CPU (float64) -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Purpose: let the reference's OWN source (TG/tflib/*.py and the model / loss sections of the three
CT_gan_*.py scripts, translated py2->py3 on the fly by oracle/ref_harness.py) execute in this
container, where TensorFlow 1.2.1 cannot be installed.  Each function restates the documented
TF-1.x behaviour of the op it stands for (op semantics: oracle/tf_ops.py).
Random ops draw from `rng` (a numpy RandomState) and every draw is recorded in `draws`, in
graph-construction order, so the same numbers can be replayed through the oracle.
"""
import contextlib
import types

import numpy as np
import torch

from .. import tf_ops

DT = torch.float64
__version__ = '1.2.1'
float32 = 'float32'
int32 = 'int32'

rng = np.random.RandomState(0)
draws = []            # [(kind, tensor)] in call order
feeds = []            # values handed out by tf.placeholder, in call order


_variables = []       # every tf.Variable created since the last reset_variables()


def reset(seed, placeholder_values):
    global rng
    rng = np.random.RandomState(seed)
    draws.clear()
    feeds[:] = list(placeholder_values)


def reset_variables():
    _variables.clear()


class _IdentityList(list):
    """`p in tf.trainable_variables()` (LS/tflib/__init__.py:38-39) must compare variables by identity, not by value."""

    def __contains__(self, item):
        return any(item is v for v in self)


def trainable_variables():
    return _IdentityList(v for v in _variables if v.requires_grad)


class _Shape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


# the reference calls x.get_shape(); give torch tensors that method (test process only)
torch.Tensor.get_shape = lambda self: _Shape(int(d) for d in self.shape)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


def Variable(initial_value, name=None, trainable=True, **kw):
    v = torch.tensor(np.asarray(initial_value), dtype=DT, requires_grad=bool(trainable))
    v.tf_name = name
    _variables.append(v)
    return v


def placeholder(dtype, shape=None, **kw):
    return feeds.pop(0)


def constant(value, **kw):
    return _t(value)


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


device = name_scope


def cast(x, dtype):
    x = _t(x)
    if dtype in ('float32', 'float'):
        return x.to(torch.float32) if not x.is_floating_point() else x
    if dtype in ('int32',):
        return torch.floor(x).to(torch.int32) if x.is_floating_point() else x.to(torch.int32)
    raise NotImplementedError(dtype)


def to_int32(x):
    return x.to(torch.int32)


def reshape(x, shape):
    return x.reshape([int(s) for s in shape])


def transpose(x, perm, name=None):
    return x.permute(*perm)


def shape(x):
    return [int(s) for s in x.shape]


def stack(values, **kw):
    return [int(v) for v in values]


def pack(values, **kw):
    raise AttributeError("module 'tensorflow' has no attribute 'pack'")    # TF >= 1.0 (deconv2d.py:93-96)


def unpack(values, **kw):
    return list(values)


def concat(values, axis=0, **kw):
    return torch.cat(list(values), dim=axis)


def split(value, num_or_size_splits, axis=0, **kw):
    return list(torch.chunk(value, num_or_size_splits, dim=axis))


def add_n(inputs):
    out = inputs[0]
    for t in inputs[1:]:
        out = out + t
    return out


def maximum(a, b):
    return torch.maximum(_t(a).to(DT) if not isinstance(a, torch.Tensor) else a,
                         _t(b).to(DT) if not isinstance(b, torch.Tensor) else b)


def square(x):
    return x * x


def sqrt(x):
    return torch.sqrt(x)


def tanh(x):
    return torch.tanh(x)


def identity(x):
    return x


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def ones_like(x):
    return torch.ones_like(x)


def zeros_like(x):
    return torch.zeros_like(x)


def _axes(kw):
    ax = kw.get('axis', kw.get('reduction_indices'))
    return None if ax is None else (list(ax) if isinstance(ax, (list, tuple)) else [ax])


def reduce_mean(x, *a, **kw):
    ax = _axes(kw) if not a else (list(a[0]) if isinstance(a[0], (list, tuple)) else [a[0]])
    return x.mean() if ax is None else x.mean(dim=ax)


def reduce_sum(x, *a, **kw):
    ax = _axes(kw) if not a else (list(a[0]) if isinstance(a[0], (list, tuple)) else [a[0]])
    return x.sum() if ax is None else x.sum(dim=ax)


def argmax(x, dimension=None, axis=None):
    return x.argmax(dim=dimension if dimension is not None else axis)


def equal(a, b):
    return a == b


def matmul(a, b):
    return a.to(b.dtype) @ b


def depth_to_space(x, block_size):
    """NHWC depth_to_space: out[n, h*b+i, w*b+j, c] = x[n, h, w, (i*b + j)*C + c]."""
    N, H, W, D = x.shape
    b = block_size
    C = D // (b * b)
    return x.reshape(N, H, W, b, b, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H * b, W * b, C)


def random_normal(shape, **kw):
    t = torch.from_numpy(rng.standard_normal(size=[int(s) for s in shape]).astype('float32'))
    draws.append(('normal', t))
    return t


def random_uniform(shape, minval=0., maxval=1., **kw):
    u = torch.from_numpy(rng.random_sample(size=[int(s) for s in shape]).astype('float32'))
    t = (np.float32(minval) + np.float32(maxval - minval) * u)
    draws.append(('uniform', t))
    return t


def gradients(ys, xs, **kw):
    y = ys if isinstance(ys, torch.Tensor) else add_n([t.sum() for t in ys])
    xs = list(xs)
    live = [x for x in xs if x.requires_grad]
    got = iter(torch.autograd.grad(y.sum(), live, create_graph=True, allow_unused=True)) if live else iter(())
    # a gradient wrt a pure input (TG/CT_gan_cifar.py:145, a dev-only metric) cannot be taken after the fact in
    # eager mode; it feeds nothing on the training path, so a zero tensor stands in for it
    return [next(got) if x.requires_grad else torch.zeros_like(x) for x in xs]


def _nn_dropout(x, keep_prob, **kw):
    if keep_prob == 1.0 or keep_prob == 1:
        return x                                              # TF returns x itself for keep_prob == 1
    u = torch.from_numpy(rng.random_sample(size=tuple(x.shape)).astype('float32'))
    draws.append(('dropout', u))
    return tf_ops.dropout(x, keep_prob, u)


def _nn_conv2d(input, filter, strides, padding, data_format='NHWC', **kw):
    assert padding == 'SAME' and data_format == 'NCHW' and strides[2] == strides[3]
    return tf_ops.conv2d_same(input.to(filter.dtype), filter, strides[2])


def _nn_conv2d_transpose(value, filter, output_shape, strides, padding, **kw):
    assert padding == 'SAME' and list(strides) == [1, 2, 2, 1]
    y = tf_ops.conv2d_transpose_same2(value.permute(0, 3, 1, 2).to(filter.dtype), filter)
    y = y.permute(0, 2, 3, 1)
    assert list(y.shape) == [int(s) for s in output_shape], (y.shape, output_shape)
    return y


def _nn_bias_add(value, bias, data_format=None, **kw):
    if data_format == 'NCHW':
        return tf_ops.bias_add_nchw(value, bias)
    return value + bias


def _nn_moments(x, axes, keep_dims=False, **kw):
    m, v = tf_ops.moments(x, list(axes))
    if not keep_dims:
        for a in sorted(axes, reverse=True):
            m, v = m.squeeze(a), v.squeeze(a)
    return m, v


def _nn_fused_batch_norm(x, scale, offset, epsilon=0.001, data_format='NHWC', **kw):
    assert data_format == 'NCHW'
    m, v = tf_ops.moments(x, [0, 2, 3])
    y = tf_ops.batch_normalization(x, m, v, offset.view(1, -1, 1, 1), scale.view(1, -1, 1, 1), epsilon)
    return y, m.reshape(-1), v.reshape(-1)


def _nn_xent(logits=None, labels=None, **kw):
    return tf_ops.sparse_softmax_cross_entropy_with_logits(logits, labels)


nn = types.SimpleNamespace(
    conv2d=_nn_conv2d, conv2d_transpose=_nn_conv2d_transpose, bias_add=_nn_bias_add, dropout=_nn_dropout,
    relu=lambda x: torch.relu(x), sigmoid=lambda x: torch.sigmoid(x), moments=_nn_moments,
    batch_normalization=lambda x, mean, variance, offset, scale, variance_epsilon:
        tf_ops.batch_normalization(x, mean, variance, offset, scale, variance_epsilon),
    fused_batch_norm=_nn_fused_batch_norm,
    embedding_lookup=lambda params, ids: params[ids.long()],
    sparse_softmax_cross_entropy_with_logits=_nn_xent,
)
