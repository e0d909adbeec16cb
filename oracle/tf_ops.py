"""TensorFlow-1.x semantics of the ops on the CT-GAN hot path, restated on PyTorch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function names the `tf.*`
call it restates and the reference call site that uses it.  TG =
/root/reference/CT-GANs/tensorflow_generative_model.
"""
import math

import torch
import torch.nn.functional as F


def same_pad(in_size, k, stride):
    """TF `padding='SAME'`: out=ceil(in/s), total=max((out-1)s+k-in,0), before=total//2.
    (SURVEY.md §8(c) rule 1; used by tf.nn.conv2d at TG/tflib/ops/conv2d.py:106-112.)"""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def conv2d_same(x, w_hwio, stride=1):
    """tf.nn.conv2d(x NCHW, filter HWIO, strides=[1,1,s,s], 'SAME', NCHW)
    -- TG/tflib/ops/conv2d.py:106-112.  Cross-correlation, asymmetric SAME padding."""
    k = w_hwio.shape[0]
    _, pt, pb = same_pad(x.shape[2], k, stride)
    _, pl, pr = same_pad(x.shape[3], k, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), stride=stride)


def conv2d_transpose_same2(x, w_hwoi):
    """tf.nn.conv2d_transpose(value NHWC, filter [k,k,out,in], output 2H x 2W,
    strides 2, 'SAME') -- TG/tflib/ops/deconv2d.py:97-103, here on NCHW `x`.
    It is the gradient-wrt-input of the stride-2 SAME conv whose input is 2H:
    the full transposed conv cropped by that conv's leading pad."""
    k = w_hwoi.shape[0]
    H, W = x.shape[2], x.shape[3]
    _, pt, _ = same_pad(2 * H, k, 2)
    _, pl, _ = same_pad(2 * W, k, 2)
    full = F.conv_transpose2d(x, w_hwoi.permute(3, 2, 0, 1), stride=2)
    # full is (H-1)*2+k; when that is shorter than pad+2H (k=1,2) pad with zeros
    need_h, need_w = pt + 2 * H, pl + 2 * W
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    return full[:, :, pt:pt + 2 * H, pl:pl + 2 * W]


def bias_add_nchw(x, b):
    """tf.nn.bias_add(data_format='NCHW') -- TG/tflib/ops/conv2d.py:120."""
    return x + b.view(1, -1, *([1] * (x.dim() - 2)))


def moments(x, axes):
    """tf.nn.moments(keep_dims=True): mean and BIASED variance
    -- TG/tflib/ops/batchnorm.py:77, cond_batchnorm.py:10."""
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    return mean, var


def batch_normalization(x, mean, var, offset, scale, eps):
    """tf.nn.batch_normalization: (x-mean)*rsqrt(var+eps)*scale+offset
    -- TG/tflib/ops/batchnorm.py:84, cond_batchnorm.py:16."""
    inv = torch.rsqrt(var + eps)
    return (x - mean) * (inv * scale) + offset


def fused_batch_norm_training(x, scale, offset, eps=1e-5):
    """tf.nn.fused_batch_norm(..., epsilon=1e-5, data_format='NCHW') in training mode
    -- TG/tflib/ops/batchnorm.py:29-30: y uses the biased batch variance."""
    mean, var = moments(x, [0, 2, 3])
    return batch_normalization(x, mean, var, offset.view(1, -1, 1, 1), scale.view(1, -1, 1, 1), eps)


def dropout(x, keep_prob, u):
    """tf.nn.dropout(x, keep_prob): x / keep * floor(keep + u), u ~ U[0,1) (TF 1.2
    python/ops/nn_ops.py).  keep_prob == 1 returns x unchanged (no random op).
    `u` is the injected uniform tensor (fp32); the 0/1 pattern is computed in fp32
    exactly as the device does, then cast.  Call sites: TG/CT_gan_cifar.py:86,91,96;
    TG/CT_gan_mnist.py:94,99,104; TG/CT_gan_cifar_resnet.py:173,175,177."""
    if keep_prob == 1.0:
        return x
    binary = torch.floor(torch.tensor(keep_prob, dtype=torch.float32) + u.to(torch.float32))
    return x / keep_prob * binary.to(x.dtype)


def leaky_relu(x, alpha=0.2):
    """tf.maximum(alpha*x, x) -- TG/CT_gan_cifar.py:47-48."""
    return torch.maximum(alpha * x, x)


def sparse_softmax_cross_entropy_with_logits(logits, labels):
    """tf.nn.sparse_softmax_cross_entropy_with_logits -- TG/CT_gan_cifar_resnet.py:247,324."""
    return torch.logsumexp(logits, dim=1) - logits.gather(1, labels.view(-1, 1).long()).view(-1)


def mean_pool_2x2(x):
    """add_n of the four strided slices / 4 -- TG/CT_gan_cifar_resnet.py:91,96."""
    return (x[:, :, ::2, ::2] + x[:, :, 1::2, ::2] + x[:, :, ::2, 1::2] + x[:, :, 1::2, 1::2]) / 4.


def upsample_2x(x):
    """concat x4 on channels + depth_to_space(2) = nearest-neighbour 2x
    -- TG/CT_gan_cifar_resnet.py:102-105."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


class TFAdam:
    """tf.train.AdamOptimizer (TF 1.2 python/training/adam.py):
        lr_t = lr * sqrt(1-b2^t) / (1-b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
        p -= lr_t * m / (sqrt(v) + eps)           (eps OUTSIDE the sqrt, not bias-corrected)
    Variables whose gradient is None are skipped.  One instance per optimizer
    (TG/CT_gan_cifar_resnet.py:333-334)."""

    def __init__(self, beta1, beta2, eps=1e-8):
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def apply(self, params, grads, lr):
        """params/grads: dict name -> tensor (grads may miss names / hold None). In place."""
        self.t += 1
        b1, b2 = self.beta1, self.beta2
        lr_t = lr * math.sqrt(1. - b2 ** self.t) / (1. - b1 ** self.t)
        for name, p in params.items():
            g = grads.get(name)
            if g is None:
                continue
            if name not in self.m:
                self.m[name] = torch.zeros_like(p)
                self.v[name] = torch.zeros_like(p)
            m, v = self.m[name], self.v[name]
            m.mul_(b1).add_(g, alpha=1. - b1)
            v.mul_(b2).addcmul_(g, g, value=1. - b2)
            p.data.sub_(lr_t * m / (v.sqrt() + self.eps))
