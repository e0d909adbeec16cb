"""lib.ops.linear.Linear -- drop-in for TG/tflib/ops/linear.py:24-148.

`<name>.W` is [input_dim, output_dim], `<name>.b` [output_dim].  The `elif` order of the
reference makes initialization=None always Glorot-uniform (linear.py:55-60).
"""
import numpy as np

from ... import tflib as lib
from ... import functional as F

_default_weightnorm = False


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


def disable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = False


_weights_stdev = None


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Linear(name, input_dim, output_dim, inputs, biases=True, initialization=None, weightnorm=None, gain=1.,
           out_dtype=None, input_nhwc=None):
    """
    initialization: None, `lecun`, 'glorot', `he`, 'glorot_he', `orthogonal`, `("uniform", range)`
    out_dtype (extension): dtype of the result; the critic heads keep float32 outputs.
    input_nhwc (extension) = (C, H, W): `inputs` is functional.flat_nhwc(activation) -- the features of a [N, C, H, W] map in
        (h, w, c) order instead of the (c, h, w) order of tf.reshape(output, [-1, C*H*W]) (TG/CT_gan_cifar.py:97).  The parameter
        keeps the reference's row order; its K rows are re-ordered on the fly (one layout kernel over K floats, twice
        differentiable) instead of transposing the batch of activations and, in the backward passes, their gradients.
    """
    def uniform(stdev, size):
        if _weights_stdev is not None:
            stdev = _weights_stdev
        return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')

    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm:
        raise Exception('Unsupported configuration')

    if not lib.has_param(name + '.W'):
        if initialization == 'lecun':
            weight_values = uniform(np.sqrt(1. / input_dim), (input_dim, output_dim))
        elif initialization == 'glorot' or (initialization is None):
            weight_values = uniform(np.sqrt(2. / (input_dim + output_dim)), (input_dim, output_dim))
        elif initialization == 'he':
            weight_values = uniform(np.sqrt(2. / input_dim), (input_dim, output_dim))
        elif initialization == 'glorot_he':
            weight_values = uniform(np.sqrt(4. / (input_dim + output_dim)), (input_dim, output_dim))
        elif initialization == 'orthogonal':
            def sample(shape):
                if len(shape) < 2:
                    raise RuntimeError("Only shapes of length 2 or more are supported.")
                flat_shape = (shape[0], np.prod(shape[1:]))
                a = np.random.normal(0.0, 1.0, flat_shape)
                u, _, v = np.linalg.svd(a, full_matrices=False)
                q = u if u.shape == flat_shape else v
                return q.reshape(shape).astype('float32')
            weight_values = sample((input_dim, output_dim))
        elif initialization[0] == 'uniform':
            weight_values = np.random.uniform(low=-initialization[1], high=initialization[1],
                                              size=(input_dim, output_dim)).astype('float32')
        else:
            raise Exception('Invalid initialization!')
        weight_values *= gain
    else:
        weight_values = None
    weight = lib.param(name + '.W', weight_values)
    b = lib.param(name + '.b', np.zeros((output_dim,), dtype='float32')) if biases else None

    if input_nhwc is not None:
        C, H, W = input_nhwc
        if output_dim != 1 or inputs.dim() != 2 or C * H * W != input_dim:
            raise Exception('Unsupported configuration')
        import torch
        weight = F.derived_from(
            F.to_nhwc(weight.reshape(1, input_dim), C, H, W, torch.float32).permute(0, 2, 3, 1).reshape(input_dim, 1), weight)
    if inputs.dim() == 2:
        return F.linear(inputs, weight, b, out_dtype=out_dtype)
    lead = inputs.shape[:-1]
    result = F.linear(inputs.reshape(-1, input_dim), weight, b, out_dtype=out_dtype)
    return result.reshape(*lead, output_dim)
