"""CT-GAN for 64x64 images: the training step of TG/CT_gan_64x64.py (MODE='wgan-ct', GoodGenerator / GoodDiscriminator).

SURVEY.md 8(f) row N4: host logic and parity against the oracle on the stand-in backend (tests/test_gan_64x64_host.py), the
layer-norm kernels (csrc/layernorm.cu) and one critic + generator step against the oracle on a B200 (tests/test_64x64_gpu.py,
fp32 and BF16 paths).  bench.py does not time it (not in BASELINE.json's configs).

Hyper-parameters :28-37, Normalize :87-93, ConvMeanPool / MeanPoolConv / UpsampleConv :106-124, ResidualBlock :166-200,
GoodGenerator :204-221, GoodDiscriminator :357-373, loss graph :480-546 (wgan-ct :494-519), Adam(1e-4, 0, .9) :560-564.
The reference splits the batch over N_GPUS = 2 towers and averages the tower costs (:480, :545-546).  Every critic op is
per-sample (layer norm, dropout), so the towers run as ONE batch; the generator's batch norm keeps per-tower statistics
(groups = N_GPUS), exactly what two Generator(BATCH_SIZE/2) calls compute.
"""
import functools

import torch

from . import tflib as lib
from . import functional as F
from . import kernels as K
from .tflib.ops import linear as _linear, conv2d as _conv2d, batchnorm as _batchnorm, layernorm as _layernorm
from .runtime import DeviceRandom, FlatAdam

LAMBDA_2 = 2.0  # parameter LAMBDA2
Factor_M = 0.0  # factor M
MODE = 'wgan-ct'  # dcgan, wgan, wgan-gp, lsgan
DIM = 64  # Model dimensionality
CRITIC_ITERS = 5  # How many iterations to train the critic for
N_GPUS = 2  # Number of GPUs
BATCH_SIZE = 64  # Batch size. Must be a multiple of N_GPUS
ITERS = 200000  # How many iterations to train for
LAMBDA = 10  # Gradient penalty lambda hyperparameter
OUTPUT_DIM = 64 * 64 * 3  # Number of pixels in each iamge

ACT_DTYPE = torch.bfloat16
RNG = None
BN_GROUPS = 1


def Normalize(name, axes, inputs):
    if ('Discriminator' in name) and (MODE == 'wgan-ct'):
        if axes != [0, 2, 3]:
            raise Exception('Layernorm over non-standard axes is unsupported')
        return lib.ops.layernorm.Layernorm(name, [1, 2, 3], inputs)
    return lib.ops.batchnorm.Batchnorm(name, axes, inputs, fused=True, groups=BN_GROUPS)


def ConvMeanPool(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
    return F.mean_pool_2x2(output)


def MeanPoolConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = F.mean_pool_2x2(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)


def UpsampleConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = F.upsample_2x(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)


def ResidualBlock(name, input_dim, output_dim, filter_size, inputs, resample=None, he_init=True):
    """
    resample: None, 'down', or 'up'
    """
    Conv2D = lib.ops.conv2d.Conv2D
    if resample == 'down':
        conv_shortcut = MeanPoolConv
        conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
        conv_2 = functools.partial(ConvMeanPool, input_dim=input_dim, output_dim=output_dim)
    elif resample == 'up':
        conv_shortcut = UpsampleConv
        conv_1 = functools.partial(UpsampleConv, input_dim=input_dim, output_dim=output_dim)
        conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
    elif resample is None:
        conv_shortcut = Conv2D
        conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
        conv_2 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim)
    else:
        raise Exception('invalid resample value')

    if output_dim == input_dim and resample is None:
        shortcut = inputs  # Identity skip-connection
    else:
        shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim, filter_size=1,
                                 he_init=False, biases=True, inputs=inputs)

    output = inputs
    output = Normalize(name + '.BN1', [0, 2, 3], output)
    output = F.relu(output)
    output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output, he_init=he_init, biases=False)
    output = Normalize(name + '.BN2', [0, 2, 3], output)
    output = F.relu(output)
    output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output, he_init=he_init)

    return F.add(shortcut, output)


def GoodGenerator(n_samples, noise=None, dim=None):
    dim = dim or DIM
    if noise is None:
        noise = RNG.normal('z', (n_samples, 128))
    noise = F.cast(noise, ACT_DTYPE)

    output = lib.ops.linear.Linear('Generator.Input', 128, 4 * 4 * 8 * dim, noise)
    output = F.to_nhwc(output, 8 * dim, 4, 4, ACT_DTYPE)

    output = ResidualBlock('Generator.Res1', 8 * dim, 8 * dim, 3, output, resample='up')
    output = ResidualBlock('Generator.Res2', 8 * dim, 4 * dim, 3, output, resample='up')
    output = ResidualBlock('Generator.Res3', 4 * dim, 2 * dim, 3, output, resample='up')
    output = ResidualBlock('Generator.Res4', 2 * dim, 1 * dim, 3, output, resample='up')

    output = Normalize('Generator.OutputN', [0, 2, 3], output)
    output = F.relu(output)
    output = lib.ops.conv2d.Conv2D('Generator.Output', 1 * dim, 3, 3, output)
    output = F.tanh(output)

    return F.to_flat_nchw(output, torch.float32)


def _dropout(output, keep):
    return F.dropout(output, keep, **RNG.dropout_args(output))


def GoodDiscriminator(inputs, dim, kp1, kp2, kp3):
    output = F.to_nhwc(inputs, 3, 64, 64, ACT_DTYPE)
    output = lib.ops.conv2d.Conv2D('Discriminator.Input', 3, dim, 3, output, he_init=False)

    output = ResidualBlock('Discriminator.Res1', dim, 2 * dim, 3, output, resample='down')
    output = ResidualBlock('Discriminator.Res2', 2 * dim, 4 * dim, 3, output, resample='down')
    output = _dropout(output, kp1)  # dropout after activator
    output = ResidualBlock('Discriminator.Res3', 4 * dim, 8 * dim, 3, output, resample='down')
    output = _dropout(output, kp2)  # dropout after activator
    output = ResidualBlock('Discriminator.Res4', 8 * dim, 8 * dim, 3, output, resample='down')
    output = _dropout(output, kp3)  # dropout after activator

    output2 = F.to_flat_nchw(output)  # tf.reshape(output, [-1, 4*4*8*dim])
    output = lib.ops.linear.Linear('Discriminator.Output', 4 * 4 * 8 * dim, 1, output2, out_dtype=torch.float32)

    return output.reshape(-1), output2


def GeneratorAndDiscriminator():
    """The reference's shipped choice (:48)."""
    return GoodGenerator, GoodDiscriminator


Generator, Discriminator = GeneratorAndDiscriminator()


class Trainer:
    """Parameters, optimizers and random stream of one training process."""

    def __init__(self, device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=None, record=False,
                 graph_safe_rng=False, dim=None):
        global ACT_DTYPE, RNG, DIM
        ACT_DTYPE = act_dtype
        if dim is not None:
            DIM = dim
        self.device = torch.device(device)
        self.B = batch_size or BATCH_SIZE
        if self.B % N_GPUS:
            raise Exception('BATCH_SIZE must be a multiple of N_GPUS')
        lib.delete_all_params()
        lib.set_device(self.device)
        self.rng = RNG = DeviceRandom(seed, self.device, record=record, graph_safe=graph_safe_rng)
        with torch.no_grad():
            RNG.scope('build')
            Discriminator(Generator(2), DIM, 1.0, 1.0, 1.0)
        self.rng.offset = 0
        self.gen_opt = FlatAdam('Generator', 1e-4, 0.0, 0.9)          # :560-561
        self.disc_opt = FlatAdam('Discriminator.', 1e-4, 0.0, 0.9)    # :562-563
        self.hp = dict(lambda_gp=float(LAMBDA), lambda2=LAMBDA_2, factor_m=Factor_M, acgan_scale=0.0)

    def activate(self):
        global RNG
        RNG = self.rng

    def prep_real(self, real_data_conv):
        """2*((int/255.)-.5) on the [B, 3, 64, 64] int batch, flattened to [B, OUTPUT_DIM] (:483)."""
        return K.prep_real(real_data_conv.reshape(real_data_conv.shape[0], OUTPUT_DIM), 255., 0.)

    def _generate(self, n_total):
        """One Generator call per tower in the reference (:484): here one batch with per-tower batch-norm statistics."""
        global BN_GROUPS
        h = n_total // N_GPUS
        noise = self.rng.normal_parts([('z.%d' % i, h) for i in range(N_GPUS)], 128)
        BN_GROUPS = N_GPUS
        self.rng.begin_stack([h] * N_GPUS)
        try:
            return Generator(n_total, noise=noise)
        finally:
            self.rng.end_stack()
            BN_GROUPS = 1

    N_RELU_G = 9          # activation sites of GoodGenerator: 2 per residual block + the one before Generator.Output

    def oracle_pattern_order(self, patterns, kind):
        """Parity tests: the recorded activation patterns arrive in device call order -- every stacked call tower-major
        (G t0, G t1, D real' t0, D real' t1, ...) -- the reference graph is built tower by tower (:480-543)."""
        passes = 4 if kind == 'critic' else 1                 # critic calls per tower: real', real'', fake, interpolates
        n_d = (len(patterns) // N_GPUS - self.N_RELU_G) // passes
        sizes = [self.N_RELU_G] * N_GPUS + [n_d] * (passes * N_GPUS)
        groups, k = [], 0
        for n in sizes:
            groups.append(patterns[k:k + n])
            k += n
        assert k == len(patterns)
        out = []
        for t in range(N_GPUS):
            out += groups[t]
            for p_ in range(passes):
                out += groups[N_GPUS + p_ * N_GPUS + t]
        return out

    def critic_forward_backward(self, all_real_data_conv):
        RNG = self.rng
        real_data = self.prep_real(all_real_data_conv)
        B = real_data.shape[0]
        h = B // N_GPUS
        with torch.no_grad():
            fake_data = self._generate(B)
        towers = lambda tag: [('drop.%d.%s' % (i, tag), h) for i in range(N_GPUS)]
        fork = K.fork_branch(real_data)
        stacked = torch.cat([real_data, real_data, fake_data], dim=0)
        RNG.scope_parts(towers('real1') + towers('real2') + towers('fake'))
        RNG.begin_stack([h] * (3 * N_GPUS))
        d_all, f_all = Discriminator(stacked, DIM, 0.8, 0.5, 0.5)
        RNG.end_stack()
        with K.branch(fork):
            alpha = torch.cat([RNG.uniform('alpha.%d' % i, (h, 1)) for i in range(N_GPUS)], dim=0)
            interpolates = K.interpolate(real_data, fake_data, alpha).requires_grad_(True)
            RNG.scope_parts(towers('gp'))
            RNG.begin_stack([h] * N_GPUS)
            d_interp = Discriminator(interpolates, DIM, 0.8, 0.5, 0.5)[0]
            RNG.end_stack()
            with F.no_param_grads():                  # tf.gradients(..., [interpolates]) (:505)
                gradients = torch.autograd.grad(d_interp, interpolates, grad_outputs=torch.ones_like(d_interp),
                                                create_graph=True)[0]
        K.join_branch(fork)
        # equal tower sizes: the mean over the batch == the mean of the tower means (:545-546)
        out = F.CTGPLossStacked.apply(d_all, f_all, gradients, None, None, self.hp,
                                      dict(real=(0, B), real2=(B, 2 * B), fake=(2 * B, 3 * B)))
        out[0].backward(inputs=self.disc_opt.param_list())
        K.join_branch(fork)
        K.join_side()
        return dict(out=out.detach(), gradients=gradients.detach(), fake_data=fake_data, real_data=real_data)

    def critic_step(self, all_real_data_conv, iteration=0, use_device_lr=False):
        self.disc_opt.zero_grad()
        res = self.critic_forward_backward(all_real_data_conv)
        world = self.disc_opt.all_reduce()
        self.disc_opt.step(None, world, use_device_lr=use_device_lr)
        self.rng.end_step()
        return res

    def gen_forward_backward(self):
        RNG = self.rng
        h = self.B // N_GPUS
        fake_data = self._generate(self.B)
        RNG.scope_parts([('drop.%d.fake' % i, h) for i in range(N_GPUS)])
        RNG.begin_stack([h] * N_GPUS)
        disc_fake, _ = Discriminator(fake_data, DIM, 0.8, 0.5, 0.5)
        RNG.end_stack()
        gen_cost = F.MeanLoss.apply(disc_fake, -1.0)
        with F.frozen(self.disc_opt.param_list()):
            gen_cost.backward(inputs=self.gen_opt.param_list())
        K.join_side()
        return dict(cost=gen_cost.detach())

    def gen_step(self, iteration=0, use_device_lr=False):
        self.gen_opt.zero_grad()
        res = self.gen_forward_backward()
        world = self.gen_opt.all_reduce()
        self.gen_opt.step(None, world, use_device_lr=use_device_lr)
        self.rng.end_step()
        return res
