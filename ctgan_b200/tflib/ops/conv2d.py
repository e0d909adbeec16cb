"""lib.ops.conv2d.Conv2D -- drop-in for TG/tflib/ops/conv2d.py:20-123.

Same signature, parameter names (`<name>.Filters` HWIO, `<name>.Biases`), init formulas
and NCHW-in/NCHW-out contract; the arithmetic is libctgan_sm100's conv kernels
(tcgen05 implicit GEMM when eligible, SIMT otherwise) instead of tf.nn.conv2d.
"""
import numpy as np

from ... import tflib as lib
from ... import functional as F

_default_weightnorm = False


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


_weights_stdev = None


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def _uniform(stdev, size):
    return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')


def Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=True, mask_type=None, stride=1,
           weightnorm=None, biases=True, gain=1., residual=None, relu=False, in_relu=False,
           relu_bwd_fused=False, residual_up2=False, act_dropout=None, out_s2d=False, mean_pool=False):
    """
    inputs: tensor of shape (batch size, num channels, height, width)
    mask_type: one of None, 'a', 'b'  (PixelCNN masks: unused by the CT-GAN scripts -> unsupported)

    returns: tensor of shape (batch size, num channels, height, width)
    """
    if mask_type is not None:
        raise Exception('Unsupported configuration')
    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm:
        raise Exception('Unsupported configuration')

    if not lib.has_param(name + '.Filters'):
        fan_in = input_dim * filter_size ** 2
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)
        if he_init:
            filters_stdev = np.sqrt(4. / (fan_in + fan_out))
        else:  # Normalized init (Glorot & Bengio)
            filters_stdev = np.sqrt(2. / (fan_in + fan_out))
        stdev = _weights_stdev if _weights_stdev is not None else filters_stdev
        filter_values = _uniform(stdev, (filter_size, filter_size, input_dim, output_dim))
        filter_values *= gain
    else:
        filter_values = None
    filters = lib.param(name + '.Filters', filter_values)
    _biases = lib.param(name + lib.CONV_BIAS, np.zeros(output_dim, dtype='float32')) if biases else None

    if not isinstance(inputs, F.S2DAct):
        inputs = F.ensure_nhwc(inputs)
    if inputs.shape[1] != input_dim:
        raise Exception('Conv2D %s: expected %d input channels, got %d' % (name, input_dim, inputs.shape[1]))
    # residual (extension): a tensor of the output's shape added in the conv epilogue (skip connections)
    # relu (extension): the nonlinearity that follows this conv, applied in the epilogue
    # in_relu / relu_bwd_fused (extension): see functional.ConvF -- the ReLU between two convs differentiated inside
    # the second conv's dgrad epilogue
    # residual_up2 (extension): residual at half resolution, added nearest-neighbour upsampled
    # act_dropout (extension) = dict(slope, keep, rng[, next_cout, next_k]): the LeakyReLU + tf.nn.dropout that follow this
    # conv in the DCGAN critics (TG/CT_gan_cifar.py:84-96), applied in the conv epilogue (functional.conv2d_act_dropout)
    # out_s2d (extension): the result (relu fused) is handed to the next layer as a functional.S2DAct -- the producer of
    # mean_pool (extension): this 3x3 conv FOLLOWED BY the 2x2 mean pool of ConvMeanPool (TG/CT_gan_cifar_resnet.py:89-92) as
    # one stride-2 conv over that S2DAct (functional.conv_mean_pool_s2d)
    if mean_pool:
        if filter_size != 3 or stride != 1:
            raise Exception('Conv2D %s: mean_pool fuses a 3x3 / stride-1 conv only' % name)
        return F.conv_mean_pool_s2d(inputs, filters, _biases, residual=residual)
    if act_dropout is not None:
        return F.conv2d_act_dropout(inputs, filters, _biases, filter_size, stride, **act_dropout)
    return F.conv2d(inputs, filters, _biases, filter_size, stride, residual=residual, relu=relu, in_relu=in_relu,
                    relu_bwd_fused=relu_bwd_fused, res_up2=residual_up2, out_s2d=out_s2d)
