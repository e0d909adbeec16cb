"""CT-GAN for 128x128 images: the training step of LS/wgan_LSUN_Bedrooms128.py (LS = TG/LSUN_bedrooms).

SURVEY.md 8(f) row N4, second half.  Hyper-parameters :31-57, Normalize :70-74 (layer norm over [1,2,3] in the critic, fused
batch norm in the generator), MeanPoolConv / ScaledUpsampleConv :80-93, ResidualBlock :95-137 ('down': 3x3 conv + STRIDE-2
3x3 conv, shortcut MeanPoolConv; 'up': ScaledUpsampleConv x2, gain 0.5), ResnetGenerator :139-167, ResnetDiscriminator
:169-205, loss graph :211-283, lr decay :285-288, Adam(1e-4 * decay, 0, .9) :289,296, generator cost :291-295.
The script's own copy of tflib names conv biases and normalisation offsets `<name>.b` (lib.set_name_style('lsun')).

Execution follows gan_cifar_resnet.py (the two scripts share the two-device graph): the two stochastic critic calls on
real + fake (:231-232) run as ONE stacked batch [real, fake, real] -- the fake half of the second call feeds nothing (:238-250) --
the gradient-penalty pass is a second stream branch, every critic op is per sample (layer norm, dropout), and the generator's
per-device calls (:216-218, :293) run as one batch with per-device batch-norm statistics.
`WIDTH` scales every DIM_* constant (tests run a narrow model; 1.0 = the reference's widths).
"""
import functools

import torch

from . import tflib as lib
from . import functional as F
from . import kernels as K
from .tflib.ops import linear as _linear, conv2d as _conv2d, batchnorm as _batchnorm, layernorm as _layernorm
from .runtime import DeviceRandom, FlatAdam

N_GPUS = 2
BATCH_SIZE = 64

DIM_G_64 = 64
DIM_G_32 = 128
DIM_G_16 = 256
DIM_G_8 = 512
DIM_G_4 = 512

DIM_D_64 = 128
DIM_D_32 = 256
DIM_D_16 = 512
DIM_D_8 = 1024
DIM_D_4 = 1024

NORMALIZATION_G = True
NORMALIZATION_D = True

ITERS = 200000
LAMBDA_2 = 2.0  # parameter LAMBDA2
Factor_M = 0.0  # factor M
LR = 1e-4
DECAY = True
CRITIC_ITERS = 5
MOMENTUM_G = 0.
MOMENTUM_D = 0.
GEN_BS_MULTIPLE = 1

OUTPUT_DIM = 3 * 128 * 128

WIDTH = 1.0
ACT_DTYPE = torch.bfloat16
RNG = None
BN_GROUPS = 1


def _w(dim):
    return max(1, int(dim * WIDTH))


def nonlinearity(x):
    return F.relu(x)


def Normalize(name, inputs):
    if ('Discriminator' in name) and NORMALIZATION_D:
        return lib.ops.layernorm.Layernorm(name, [1, 2, 3], inputs)
    elif ('Generator' in name) and NORMALIZATION_G:
        return lib.ops.batchnorm.Batchnorm(name, [0, 2, 3], inputs, fused=True, groups=BN_GROUPS)
    return inputs


def ConvMeanPool(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
    return F.mean_pool_2x2(output)


def MeanPoolConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = F.mean_pool_2x2(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)


def ScaledUpsampleConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = F.upsample_2x(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases, gain=0.5)


def ResidualBlock(name, input_dim, output_dim, filter_size, inputs, resample=None):
    """
    resample: None, 'down', or 'up'
    """
    Conv2D = lib.ops.conv2d.Conv2D
    if resample == 'down':
        conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
        conv_2 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim, stride=2)
        conv_shortcut = MeanPoolConv
    elif resample == 'up':
        conv_1 = functools.partial(ScaledUpsampleConv, input_dim=input_dim, output_dim=output_dim)
        conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
        conv_shortcut = ScaledUpsampleConv
    elif resample is None:
        conv_shortcut = Conv2D
        conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim)
        conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
    else:
        raise Exception('invalid resample value')

    if output_dim == input_dim and resample is None:
        shortcut = inputs  # Identity skip-connection
    else:
        shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim, filter_size=1, he_init=False,
                                 biases=True, inputs=inputs)

    output = inputs
    output = Normalize(name + '.N1', output)
    output = nonlinearity(output)
    output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output)
    output = Normalize(name + '.N2', output)
    output = nonlinearity(output)
    output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output)

    return F.add(shortcut, output)


def ResnetGenerator(n_samples, noise=None):
    if noise is None:
        noise = RNG.normal('z', (n_samples, 128))
    noise = F.cast(noise, ACT_DTYPE)

    output = lib.ops.linear.Linear('Generator.Input', 128, 4 * 4 * _w(DIM_G_4), noise)
    output = F.to_nhwc(output, _w(DIM_G_4), 4, 4, ACT_DTYPE)

    output = ResidualBlock('Generator.4_3', _w(DIM_G_4), _w(DIM_G_8), 3, output, resample='up')
    output = ResidualBlock('Generator.8_3', _w(DIM_G_8), _w(DIM_G_16), 3, output, resample='up')
    output = ResidualBlock('Generator.16_3', _w(DIM_G_16), _w(DIM_G_32), 3, output, resample='up')
    output = ResidualBlock('Generator.32_3', _w(DIM_G_32), _w(DIM_G_64), 3, output, resample='up')

    output = Normalize('Generator.OutputN', output)
    output = nonlinearity(output)
    output = ScaledUpsampleConv('Generator.Output', _w(DIM_G_64), 3, 5, output, he_init=False)

    output = F.tanh(output)

    return F.to_flat_nchw(output, torch.float32)


def _dropout(output, keep):
    return F.dropout(output, keep, **RNG.dropout_args(output))


def ResnetDiscriminator(inputs, kp1, kp2, kp3):
    output = F.to_nhwc(inputs, 3, 128, 128, ACT_DTYPE)

    output = lib.ops.conv2d.Conv2D('Discriminator.Input', 3, _w(DIM_D_64), 5, output, he_init=True, stride=2)

    output = ResidualBlock('Discriminator.64_3', _w(DIM_D_64), _w(DIM_D_32), 3, output, resample='down')
    output = ResidualBlock('Discriminator.32_3', _w(DIM_D_32), _w(DIM_D_16), 3, output, resample='down')
    output = ResidualBlock('Discriminator.16_3', _w(DIM_D_16), _w(DIM_D_8), 3, output, resample='down')
    output = _dropout(output, kp1)  # dropout after activator
    output = ResidualBlock('Discriminator.8_1', _w(DIM_D_8), _w(DIM_D_8), 3, output, resample=None)
    output = _dropout(output, kp2)  # dropout after activator
    output = ResidualBlock('Discriminator.8_2', _w(DIM_D_8), _w(DIM_D_8), 3, output, resample=None)
    output = _dropout(output, kp3)  # dropout after activator

    output2 = F.spatial_mean(output)  # tf.reduce_mean(output, axis=[2,3])
    output = lib.ops.linear.Linear('Discriminator.Output', _w(DIM_D_8), 1, output2, out_dtype=torch.float32)

    return output.reshape(-1), output2


def GeneratorAndDiscriminator():
    return ResnetGenerator, ResnetDiscriminator


Generator, Discriminator = GeneratorAndDiscriminator()


class Trainer:
    """Parameters, optimizers and random stream of one training process."""

    def __init__(self, device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=None, record=False,
                 graph_safe_rng=False, width=None):
        global ACT_DTYPE, RNG, WIDTH
        ACT_DTYPE = act_dtype
        if width is not None:
            WIDTH = width
        self.device = torch.device(device)
        self.B = batch_size or BATCH_SIZE
        if self.B % N_GPUS:
            raise Exception('BATCH_SIZE must be a multiple of N_GPUS')
        lib.delete_all_params()
        lib.set_name_style('lsun')        # LS/tflib/ops/conv2d.py:117, batchnorm.py:24, layernorm.py:15
        lib.set_device(self.device)
        self.rng = RNG = DeviceRandom(seed, self.device, record=record, graph_safe=graph_safe_rng)
        with torch.no_grad():
            RNG.scope('build')
            Discriminator(Generator(2), 1.0, 1.0, 1.0)
        self.rng.offset = 0
        self.gen_opt = FlatAdam('Generator', LR, MOMENTUM_G, 0.9)          # :296
        self.disc_opt = FlatAdam('Discriminator.', LR, MOMENTUM_D, 0.9)    # :289
        self.hp = dict(lambda_gp=10.0, lambda2=LAMBDA_2, factor_m=Factor_M, acgan_scale=0.0)   # 10.*mean((slopes-1)^2) :267

    def activate(self):
        global RNG
        RNG = self.rng

    @staticmethod
    def lr(iteration):
        decay = max(0., 1. - (float(iteration) / ITERS)) if DECAY else 1.     # :285-288
        return LR * decay

    def prep_real(self, all_real_data_conv, out=None, out2=None):
        """2*((int/255.)-.5) on the [B, 3, 128, 128] int batch, flattened to [B, OUTPUT_DIM] (:221)."""
        return K.prep_real(all_real_data_conv.reshape(all_real_data_conv.shape[0], OUTPUT_DIM), 255., 0., out=out, out2=out2)

    def _generate(self, n_total):
        """One Generator call per device in the reference (:216-218, :293): one batch with per-device batch-norm statistics."""
        global BN_GROUPS
        h = n_total // N_GPUS
        noise = self.rng.normal_parts([('z.%d' % i, h) for i in range(N_GPUS)], 128)
        BN_GROUPS = N_GPUS
        try:
            return Generator(n_total, noise=noise)
        finally:
            BN_GROUPS = 1

    def critic_forward_backward(self, all_real_data_conv):
        RNG = self.rng
        B = all_real_data_conv.shape[0]
        h = B // N_GPUS
        with torch.no_grad():
            RNG.begin_stack([h] * N_GPUS)
            fake_data = self._generate(B)
            RNG.end_stack()
        # stochastic call ' on real + fake (2B rows) and call '' on the real half (B rows) as ONE critic batch (:231-232)
        stacked = torch.empty((3 * B, OUTPUT_DIM), dtype=torch.float32, device=all_real_data_conv.device)
        all_real_data = self.prep_real(all_real_data_conv, out=stacked[:B], out2=stacked[2 * B:])
        stacked[B:2 * B].copy_(fake_data)
        fork = K.fork_branch(all_real_data)
        RNG.scope_parts([('drop.p1', 2 * B), ('drop.p2', B)])
        RNG.begin_stack([2 * B, B])
        disc_all, disc_all_2 = Discriminator(stacked, 0.8, 0.5, 0.5)
        RNG.end_stack()
        with K.branch(fork):                          # gradient penalty (:260-267) as a second stream branch
            alpha = RNG.uniform('alpha', (B, 1))
            interpolates = K.interpolate(all_real_data, fake_data, alpha).requires_grad_(True)
            RNG.scope('drop.gp')
            d_interp = Discriminator(interpolates, 0.8, 0.5, 0.5)[0]
            with F.no_param_grads():                  # tf.gradients(..., [interpolates]) (:265)
                gradients = torch.autograd.grad(d_interp, interpolates, grad_outputs=torch.ones_like(d_interp),
                                                create_graph=True)[0]
        K.join_branch(fork)
        out = F.CTGPLossStacked.apply(disc_all, disc_all_2, gradients, None, None, self.hp,
                                      dict(real=(0, B), fake=(B, 2 * B), real2=(2 * B, 3 * B)))
        out[0].backward(inputs=self.disc_opt.param_list())
        K.join_branch(fork)
        K.join_side()
        return dict(out=out.detach(), gradients=gradients.detach(), fake_data=fake_data, real_data=all_real_data)

    def critic_step(self, all_real_data_conv, iteration=0, use_device_lr=False):
        self.disc_opt.zero_grad()
        res = self.critic_forward_backward(all_real_data_conv)
        world = self.disc_opt.all_reduce()
        self.disc_opt.step(self.lr(iteration), world, use_device_lr=use_device_lr)
        self.rng.end_step()
        return res

    def gen_forward_backward(self):
        RNG = self.rng
        n = GEN_BS_MULTIPLE * self.B // N_GPUS
        RNG.begin_stack([n] * N_GPUS)
        fake_data = self._generate(n * N_GPUS)
        RNG.scope_parts([('drop.%d' % i, n) for i in range(N_GPUS)])
        disc_fake, _ = Discriminator(fake_data, 0.8, 0.5, 0.5)
        RNG.end_stack()
        gen_cost = F.MeanLoss.apply(disc_fake, -1.0)      # equal splits: the batch mean == the mean of the device means (:295)
        with F.frozen(self.disc_opt.param_list()):
            gen_cost.backward(inputs=self.gen_opt.param_list())
        K.join_side()
        return dict(cost=gen_cost.detach())

    def gen_step(self, iteration=0, use_device_lr=False):
        self.gen_opt.zero_grad()
        res = self.gen_forward_backward()
        world = self.gen_opt.all_reduce()
        self.gen_opt.step(self.lr(iteration), world, use_device_lr=use_device_lr)
        self.rng.end_step()
        return res
