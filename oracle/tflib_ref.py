"""Restatement of the reference's `tflib` op surface on PyTorch-CPU (TEST INFRASTRUCTURE).

Follows TG/tflib/__init__.py:8-48 (name-keyed param registry) and
TG/tflib/ops/{conv2d,deconv2d,linear,batchnorm,cond_batchnorm}.py, with
TG = /root/reference/CT-GANs/tensorflow_generative_model.  Initial values are
drawn from the *global* numpy RNG in the same order and with the same formulas
as the reference, so `np.random.seed(s)` before building a model reproduces
the reference's (and the product's) initial weights.
"""
import numpy as np
import torch

from . import tf_ops


class TFLib:
    """One instance == one `tflib` module state (its `_params` dict)."""

    def __init__(self, dtype=torch.float64, style='ct'):
        # style 'lsun': the LSUN fork's names -- conv biases and normalisation offsets are `<name>.b`
        # (LS/tflib/ops/conv2d.py:117, batchnorm.py:24, layernorm.py:15 with LS = TG/LSUN_bedrooms)
        self.conv_bias, self.norm_offset = ('.b', '.b') if style == 'lsun' else ('.Biases', '.offset')
        self.dtype = dtype
        self._params = {}          # TG/tflib/__init__.py:8
        self._trainable = {}

    # -- TG/tflib/__init__.py:10-34 (aliases are unused by the CT scripts)
    def param(self, name, value, trainable=True):
        if name not in self._params:
            t = torch.tensor(np.asarray(value), dtype=self.dtype, requires_grad=trainable)
            self._params[name] = t
            self._trainable[name] = trainable
        return self._params[name]

    # -- TG/tflib/__init__.py:36-37 (substring match, insertion order on py3)
    def params_with_name(self, name):
        return [p for n, p in self._params.items() if name in n]

    def named_params_with_name(self, name, trainable_only=True):
        return {n: p for n, p in self._params.items()
                if name in n and (self._trainable[n] or not trainable_only)}

    def delete_all_params(self):   # TG/tflib/__init__.py:39-40
        self._params.clear()
        self._trainable.clear()

    # ------------------------------------------------------------------ ops
    @staticmethod
    def _uniform(stdev, size):     # conv2d.py:55-60, deconv2d.py:41-46, linear.py:39-46
        return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3),
                                 size=size).astype('float32')

    def Conv2D(self, name, input_dim, output_dim, filter_size, inputs, he_init=True,
               mask_type=None, stride=1, weightnorm=None, biases=True, gain=1.):
        """TG/tflib/ops/conv2d.py:20-123."""
        if mask_type is not None or weightnorm:
            raise Exception('Unsupported configuration')   # branches unused by the CT scripts
        if name + '.Filters' not in self._params:
            fan_in = input_dim * filter_size ** 2
            fan_out = output_dim * filter_size ** 2 / (stride ** 2)      # :63
            stdev = np.sqrt(4. / (fan_in + fan_out)) if he_init else np.sqrt(2. / (fan_in + fan_out))
            filter_values = self._uniform(stdev, (filter_size, filter_size, input_dim, output_dim)) * gain
        else:
            filter_values = None
        filters = self.param(name + '.Filters', filter_values)
        result = tf_ops.conv2d_same(inputs, filters, stride)
        if biases:
            b = self.param(name + self.conv_bias, np.zeros(output_dim, dtype='float32'))
            result = tf_ops.bias_add_nchw(result, b)
        return result

    def Deconv2D(self, name, input_dim, output_dim, filter_size, inputs, he_init=True,
                 weightnorm=None, biases=True, gain=1., mask_type=None):
        """TG/tflib/ops/deconv2d.py:20-115 (fixed stride 2; filter [k,k,out,in])."""
        if mask_type is not None:
            raise Exception('Unsupported configuration')   # deconv2d.py:38-39
        if weightnorm:
            raise Exception('Unsupported configuration')
        if name + '.Filters' not in self._params:
            stride = 2
            fan_in = input_dim * filter_size ** 2 / (stride ** 2)        # :49
            fan_out = output_dim * filter_size ** 2
            stdev = np.sqrt(4. / (fan_in + fan_out)) if he_init else np.sqrt(2. / (fan_in + fan_out))
            filter_values = self._uniform(stdev, (filter_size, filter_size, output_dim, input_dim)) * gain
        else:
            filter_values = None
        filters = self.param(name + '.Filters', filter_values)
        result = tf_ops.conv2d_transpose_same2(inputs, filters)
        if biases:
            b = self.param(name + '.Biases', np.zeros(output_dim, dtype='float32'))
            result = tf_ops.bias_add_nchw(result, b)
        return result

    def Linear(self, name, input_dim, output_dim, inputs, biases=True, initialization=None,
               weightnorm=None, gain=1.):
        """TG/tflib/ops/linear.py:24-148.  NB the `elif` order of the reference makes
        `initialization=None` ALWAYS glorot (linear.py:55-60; the orthogonal branch at
        :76-77 is unreachable for None)."""
        if weightnorm:
            raise Exception('Unsupported configuration')
        if name + '.W' not in self._params:
            if initialization == 'lecun':
                w = self._uniform(np.sqrt(1. / input_dim), (input_dim, output_dim))
            elif initialization == 'glorot' or initialization is None:
                w = self._uniform(np.sqrt(2. / (input_dim + output_dim)), (input_dim, output_dim))
            elif initialization == 'he':
                w = self._uniform(np.sqrt(2. / input_dim), (input_dim, output_dim))
            elif initialization == 'glorot_he':
                w = self._uniform(np.sqrt(4. / (input_dim + output_dim)), (input_dim, output_dim))
            elif initialization[0] == 'uniform':
                w = np.random.uniform(low=-initialization[1], high=initialization[1],
                                      size=(input_dim, output_dim)).astype('float32')
            else:
                raise Exception('Invalid initialization!')
            w = w * gain
        else:
            w = None
        weight = self.param(name + '.W', w)
        if inputs.dim() == 2:
            result = inputs @ weight
        else:
            result = (inputs.reshape(-1, input_dim) @ weight).reshape(*inputs.shape[:-1], output_dim)
        if biases:
            result = result + self.param(name + '.b', np.zeros((output_dim,), dtype='float32'))
        return result

    def Batchnorm(self, name, axes, inputs, is_training=None, stats_iter=None,
                  update_moving_stats=True, fused=True):
        """TG/tflib/ops/batchnorm.py:6-87, `is_training is None` path (the only one the
        CT scripts use).  The fused path also creates the two non-trainable moving
        stats (:26-27) which never receive gradients."""
        if is_training is not None:
            raise Exception('Unsupported configuration')
        if (axes == [0, 2, 3]) and fused:
            C = inputs.shape[1]
            offset = self.param(name + self.norm_offset, np.zeros(C, dtype='float32'))
            scale = self.param(name + '.scale', np.ones(C, dtype='float32'))
            self.param(name + '.moving_mean', np.zeros(C, dtype='float32'), trainable=False)
            self.param(name + '.moving_variance', np.ones(C, dtype='float32'), trainable=False)
            return tf_ops.fused_batch_norm_training(inputs, scale, offset, 1e-5)
        mean, var = tf_ops.moments(inputs, axes)
        shape = list(mean.shape)
        if 0 not in axes:
            shape[0] = 1
        offset = self.param(name + self.norm_offset, np.zeros(shape, dtype='float32'))
        scale = self.param(name + '.scale', np.ones(shape, dtype='float32'))
        return tf_ops.batch_normalization(inputs, mean, var, offset, scale, 1e-5)

    def Layernorm(self, name, norm_axes, inputs):
        """TG/tflib/ops/layernorm.py:6-21: per-sample moments over `norm_axes` (keep_dims), scale / offset per
        'neuron' = the first of norm_axes (channels for BCHW), eps 1e-5."""
        mean, var = tf_ops.moments(inputs, norm_axes)
        n_neurons = inputs.shape[norm_axes[0]]
        offset = self.param(name + self.norm_offset, np.zeros(n_neurons, dtype='float32'))
        scale = self.param(name + '.scale', np.ones(n_neurons, dtype='float32'))
        bshape = [-1] + [1 for _ in range(len(norm_axes) - 1)]
        return tf_ops.batch_normalization(inputs, mean, var, offset.reshape(bshape), scale.reshape(bshape), 1e-5)

    def CondBatchnorm(self, name, axes, inputs, labels=None, n_labels=None):
        """TG/tflib/ops/cond_batchnorm.py:6-17."""
        if axes != [0, 2, 3]:
            raise Exception('unsupported')
        mean, var = tf_ops.moments(inputs, axes)
        C = inputs.shape[1]
        offset_m = self.param(name + '.offset', np.zeros([n_labels, C], dtype='float32'))
        scale_m = self.param(name + '.scale', np.ones([n_labels, C], dtype='float32'))
        offset = offset_m[labels.long()]
        scale = scale_m[labels.long()]
        return tf_ops.batch_normalization(inputs, mean, var, offset[:, :, None, None],
                                          scale[:, :, None, None], 1e-5)
