"""First-principles checks of oracle/tf_ops.py -- the restatement of TensorFlow 1.x op semantics that both the oracle and
the reference-code harness (oracle/tf_shim) stand on.  TensorFlow itself cannot be installed here ("parity unpinned" for
TF's internal kernels), so each op is compared with a brute-force evaluation of the DEFINITION TensorFlow
publishes for it (api_docs: tf.nn.convolution "SAME" padding, tf.nn.conv2d_transpose = gradient of conv2d w.r.t. its
input, tf.nn.dropout, tf.nn.fused_batch_norm, tf.train.AdamOptimizer), written with plain numpy loops that share no code
with the restatement."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops


def brute_conv2d_same(x, w, stride):
    """output[b, k, i, j] = sum_{di, dj, q} input[b, q, s*i + di - pad_top, s*j + dj - pad_left] * filter[di, dj, q, k]
    with out = ceil(in / s), pad_along = max((out - 1) * s + filter - in, 0), pad_top = pad_along // 2 (the extra pixel
    goes to the bottom / right)."""
    B, C, H, W = x.shape
    kh, kw, _, O = w.shape
    Ho, Wo = math.ceil(H / stride), math.ceil(W / stride)
    pt = max((Ho - 1) * stride + kh - H, 0) // 2
    pl = max((Wo - 1) * stride + kw - W, 0) // 2
    y = np.zeros((B, O, Ho, Wo))
    for i in range(Ho):
        for j in range(Wo):
            for di in range(kh):
                for dj in range(kw):
                    h, v = stride * i + di - pt, stride * j + dj - pl
                    if 0 <= h < H and 0 <= v < W:
                        y[:, :, i, j] += x[:, :, h, v] @ w[di, dj]
    return y


# (extent, filter, stride) of every convolution geometry on the hot path, small channel counts
GEOMS = [(32, 5, 2), (16, 5, 2), (8, 5, 2), (28, 5, 2), (14, 5, 2), (7, 5, 2), (32, 3, 1), (8, 3, 1), (4, 3, 1), (16, 1, 1), (64, 3, 1)]


@pytest.mark.parametrize('n,k,s', GEOMS)
def test_conv2d_same_is_the_published_definition(n, k, s):
    rs = np.random.RandomState(n * 100 + k * 10 + s)
    x, w = rs.standard_normal((2, 3, n, n)), rs.standard_normal((k, k, 3, 4))
    got = tf_ops.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), s).numpy()
    ref = brute_conv2d_same(x, w, s)
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-10
    out, before, after = tf_ops.same_pad(n, k, s)
    assert out == math.ceil(n / s) and before <= after <= before + 1


@pytest.mark.parametrize('h', [4, 7, 8, 14, 16])
def test_conv2d_transpose_is_the_gradient_of_the_stride2_conv(h):
    """tf.nn.conv2d_transpose "is actually the transpose (gradient) of conv2d": <conv(x, w), y> == <x, conv_transpose(y, w)>
    for the SAME stride-2 conv mapping 2h -> h, filter [k, k, out, in] (TG/tflib/ops/deconv2d.py:59-67,97-103)."""
    rs = np.random.RandomState(h)
    k, cin, cout = 5, 3, 4                                   # the CONV maps cin(=deconv out) -> cout(=deconv in)
    x = rs.standard_normal((2, cin, 2 * h, 2 * h))
    y = rs.standard_normal((2, cout, h, h))
    w = rs.standard_normal((k, k, cin, cout))                # conv HWIO == deconv [k, k, out, in]
    conv = brute_conv2d_same(x, w, 2)
    dec = tf_ops.conv2d_transpose_same2(torch.from_numpy(y), torch.from_numpy(w)).numpy()
    assert dec.shape == x.shape
    assert abs((conv * y).sum() - (x * dec).sum()) < 1e-9 * max(1.0, abs((conv * y).sum()))


def test_dropout_definition():
    """tf.nn.dropout: "outputs the input element scaled up by 1 / keep_prob, otherwise outputs 0"; the mask is
    floor(keep_prob + uniform[0,1)), i.e. kept with probability keep_prob."""
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((4, 1000)))
    u = torch.from_numpy(rs.random_sample((4, 1000)).astype('float32'))
    for keep in (0.5, 0.8):
        y = tf_ops.dropout(x, keep, u)
        kept = u.numpy() >= np.float32(1.0) - np.float32(keep)
        assert np.array_equal((y != 0).numpy(), kept & (x.numpy() != 0))
        assert np.allclose(y.numpy()[kept], x.numpy()[kept] / keep)
        assert abs(kept.mean() - keep) < 0.03
    assert tf_ops.dropout(x, 1.0, None) is x


def test_fused_batch_norm_training_definition():
    """Training-mode batch norm: per-channel batch mean and BIASED variance over (N, H, W), epsilon 1e-5 inside the root."""
    rs = np.random.RandomState(1)
    x = rs.standard_normal((5, 3, 4, 4)) * 2 + 1
    g, b = rs.standard_normal(3), rs.standard_normal(3)
    ref = np.empty_like(x)
    for c in range(3):
        v = x[:, c]
        ref[:, c] = (v - v.mean()) / np.sqrt(((v - v.mean()) ** 2).mean() + 1e-5) * g[c] + b[c]
    got = tf_ops.fused_batch_norm_training(torch.from_numpy(x), torch.from_numpy(g), torch.from_numpy(b), 1e-5).numpy()
    assert np.abs(got - ref).max() < 1e-12


def test_adam_definition():
    """tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m, v moving averages; p -= lr_t * m / (sqrt(v) + eps)."""
    rs = np.random.RandomState(2)
    p0, g1, g2 = rs.standard_normal(6), rs.standard_normal(6), rs.standard_normal(6)
    lr, b1, b2, eps = 1e-4, 0.5, 0.9, 1e-8
    m = v = np.zeros(6)
    p = p0.copy()
    for t, g in enumerate((g1, g2), start=1):
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        p = p - lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t) * m / (np.sqrt(v) + eps)
    opt = tf_ops.TFAdam(b1, b2, eps)
    named = {'w': torch.from_numpy(p0.copy())}
    opt.apply(named, {'w': torch.from_numpy(g1)}, lr)
    opt.apply(named, {'w': torch.from_numpy(g2), 'absent': None}, lr)
    assert np.abs(named['w'].numpy() - p).max() < 1e-15


def test_losses_and_resampling_definitions():
    rs = np.random.RandomState(3)
    logits, labels = rs.standard_normal((7, 10)), rs.randint(0, 10, 7)
    ref = np.array([-np.log(np.exp(l[y]) / np.exp(l).sum()) for l, y in zip(logits, labels)])
    got = tf_ops.sparse_softmax_cross_entropy_with_logits(torch.from_numpy(logits), torch.from_numpy(labels)).numpy()
    assert np.abs(got - ref).max() < 1e-12
    x = rs.standard_normal((2, 3, 4, 6))
    pool = tf_ops.mean_pool_2x2(torch.from_numpy(x)).numpy()
    assert np.allclose(pool, x.reshape(2, 3, 2, 2, 3, 2).mean(axis=(3, 5)))
    up = tf_ops.upsample_2x(torch.from_numpy(x)).numpy()
    assert up.shape == (2, 3, 8, 12) and all(np.array_equal(up[:, :, i::2, j::2], x) for i in (0, 1) for j in (0, 1))
    a = torch.from_numpy(rs.standard_normal(50))
    assert torch.equal(tf_ops.leaky_relu(a), torch.where(a > 0, a, 0.2 * a))
