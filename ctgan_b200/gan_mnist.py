"""CT-GAN DCGAN for MNIST: the training step of TG/CT_gan_mnist.py (MODE='wgan-CT') on B200.

Hyper-parameters :26-35, Generator :62-87 (no batch norm in this mode; 8x8 -> 7x7 crop at
:77), Discriminator :89-108, loss :146-167, Adam :168-177.  Real data arrives as float32
in [0,1] (:110) and is fed unchanged.
"""
import torch

from . import tflib as lib
from . import functional as F
from . import kernels as K
from .tflib.ops import linear as _linear, conv2d as _conv2d, batchnorm as _batchnorm, deconv2d as _deconv2d
from . import gan_cifar as _dcgan

Factor_M = 0.0  # factor M
LAMBDA_2 = 2.0  # weight factor
n_examples = 1000  # number of examples for training
MODE = 'wgan-CT'
DIM = 64  # Model dimensionality
BATCH_SIZE = 50  # Batch size
CRITIC_ITERS = 5  # For WGAN and WGAN-GP, number of critic iters per gen iter
LAMBDA = 10  # Gradient penalty lambda hyperparameter
ITERS = 50000  # How many generator iterations to train for
OUTPUT_DIM = 784  # Number of pixels in MNIST (28*28)

ACT_DTYPE = torch.bfloat16
RNG = None
HEAD_NHWC = True     # see gan_cifar.HEAD_NHWC


def LeakyReLU(x, alpha=0.2):
    return F.leaky_relu_dropout(x, alpha, 1.0)


def _lrelu_dropout(output, keep):
    return F.leaky_relu_dropout(output, 0.2, keep, **RNG.dropout_args(output))


def _conv_lrelu_dropout(name, input_dim, output_dim, inputs, keep, next_cout=None):
    """lib.ops.conv2d.Conv2D(name, ., ., 5, ., stride=2) -> LeakyReLU -> tf.nn.dropout(keep_prob=keep)"""
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, 5, inputs, stride=2,
                                 act_dropout=dict(slope=0.2, keep=keep, rng=RNG, next_cout=next_cout, next_k=5))


def Generator(n_samples, noise=None):
    if noise is None:
        noise = RNG.normal('z', (n_samples, 128))
    noise = F.cast(noise, ACT_DTYPE)
    output = lib.ops.linear.Linear('Generator.Input', 128, 4 * 4 * 4 * DIM, noise)
    output = F.relu(output)
    output = F.to_nhwc(output, 4 * DIM, 4, 4, ACT_DTYPE)

    output = lib.ops.deconv2d.Deconv2D('Generator.2', 4 * DIM, 2 * DIM, 5, output)
    output = F.relu(output)

    output = F.crop(output, 7, 7)                         # output[:,:,:7,:7]

    output = lib.ops.deconv2d.Deconv2D('Generator.3', 2 * DIM, DIM, 5, output)
    output = F.relu(output)

    output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 1, 5, output)
    # the image leaves the bf16 domain BEFORE the sigmoid: one rounding less on the generator's output (a tiny tensor)
    output = F.to_flat_nchw(output, torch.float32)
    return F.sigmoid(output)


def Discriminator(inputs):
    output = F.to_nhwc(inputs, 1, 28, 28, ACT_DTYPE)
    # Conv2D -> LeakyReLU -> dropout (:92-104), fused into the conv epilogue where the conv runs on the tensor cores
    output = _conv_lrelu_dropout('Discriminator.1', 1, DIM, output, 0.50, next_cout=2 * DIM)  # adding dropout after activators
    output = _conv_lrelu_dropout('Discriminator.2', DIM, 2 * DIM, output, 0.50, next_cout=4 * DIM)
    output = _conv_lrelu_dropout('Discriminator.3', 2 * DIM, 4 * DIM, output, 0.50)
    if HEAD_NHWC:                       # see gan_cifar.Discriminator
        output2 = F.flat_nhwc(output)
        output = lib.ops.linear.Linear('Discriminator.Output', 4 * 4 * 4 * DIM, 1, output2, out_dtype=torch.float32,
                                       input_nhwc=(4 * DIM, 4, 4))
        return output.reshape(-1), output2
    output2 = F.to_flat_nchw(output)  # D_
    output = lib.ops.linear.Linear('Discriminator.Output', 4 * 4 * 4 * DIM, 1, output2, out_dtype=torch.float32)  # D
    return output.reshape(-1), output2


class Trainer(_dcgan.Trainer):
    def prep_real(self, real_data_in):
        if real_data_in.dtype != torch.float32:
            raise RuntimeError('gan_mnist: real data must be float32 in [0,1]')
        return real_data_in.contiguous()
