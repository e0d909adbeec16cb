"""CT-GAN ResNet for CIFAR-10: the training step of TG/CT_gan_cifar_resnet.py on B200.

Same hyper-parameters (:33-56), the same builder functions with the same signatures
(Normalize :70-87, ConvMeanPool :89-92, MeanPoolConv :94-98, UpsampleConv :100-107,
ResidualBlock :109-141, OptimizedResBlockDisc1 :143-153, Generator :155-167,
Discriminator :169-186), the same loss graph (:190-330) and Adam settings (:333-338),
restated as two eager functions `critic_step` / `gen_step` (the reference's
`disc_train_op` / `gen_train_op`).  All arithmetic runs in libctgan_sm100 kernels.

Deliberate, result-preserving departures from the reference graph:
  * the second stochastic critic pass runs on the REAL half only -- the fake half of that
    pass feeds nothing (`disc_fake_`, `disc_fake_2_` are never used, :238-242);
  * the metrics-only clean pass (:228) runs only when `with_metrics=True`.
"""
import contextlib
import functools

import torch

from . import tflib as lib
from . import functional as F
from . import kernels as K
from .tflib.ops import linear as _linear, conv2d as _conv2d, batchnorm as _batchnorm, cond_batchnorm as _cond_batchnorm
from .runtime import DeviceRandom, FlatAdam

lib.ops.linear, lib.ops.conv2d, lib.ops.batchnorm, lib.ops.cond_batchnorm = _linear, _conv2d, _batchnorm, _cond_batchnorm

N_GPUS = 1
LAMBDA_2 = 2.0  # parameter LAMBDA2
n_examples = 50000  # Number of examples
Factor_M = 0.0  # factor M
BATCH_SIZE = 64  # Critic batch size
GEN_BS_MULTIPLE = 2  # Generator batch size, as a multiple of BATCH_SIZE
ITERS = 100000  # How many iterations to train for
DIM_G = 128  # Generator dimensionality
DIM_D = 128  # Critic dimensionality
NORMALIZATION_G = True  # Use batchnorm in generator?
NORMALIZATION_D = False  # Use batchnorm (or layernorm) in critic?
OUTPUT_DIM = 3072  # Number of pixels in CIFAR10 (32*32*3)
LR = 2e-4  # Initial learning rate
DECAY = True  # Whether to decay LR over learning
N_CRITIC = 5  # Critic steps per generator steps
CONDITIONAL = True  # Whether to train a conditional or unconditional model
ACGAN = True  # If CONDITIONAL, whether to use ACGAN or "vanilla" conditioning
ACGAN_SCALE = 1.  # How to scale the critic's ACGAN loss relative to WGAN loss
ACGAN_SCALE_G = 0.1  # How to scale generator's ACGAN loss relative to WGAN loss
LAMBDA = 10.0  # gradient penalty weight (literal 10.0 at :286)
N_DEVICES = 2  # len(DEVICES): the reference always builds two sub-graphs (:61-63)

ACT_DTYPE = torch.bfloat16   # activation storage: bfloat16 ("BF16 path") or float32 ("fp32 path")
RNG = None                   # DeviceRandom, set by Trainer
BN_GROUPS = 1                # >1 while the per-device-split generator calls of the reference run as ONE batch


def nonlinearity(x):
    return F.relu(x)


def Normalize(name, inputs, labels=None, relu=False, up2=False):
    """:70-87.  `relu=True` fuses the nonlinearity() that follows every Normalize call; `up2=True` also the nearest-neighbour
    2x upsampling of the UpsampleConv that consumes it (generator blocks, :132-137 with :100-107)."""
    if not CONDITIONAL:
        labels = None
    if CONDITIONAL and ACGAN and ('Discriminator' in name):
        labels = None
    if ('Discriminator' in name) and NORMALIZATION_D:
        raise Exception('Unsupported configuration')      # layernorm call is dead/broken in the reference
    elif ('Generator' in name) and NORMALIZATION_G:
        if labels is not None:
            return lib.ops.cond_batchnorm.Batchnorm(name, [0, 2, 3], inputs, labels=labels, n_labels=10, relu=relu,
                                                    groups=BN_GROUPS, up2=up2)
        else:
            return lib.ops.batchnorm.Batchnorm(name, [0, 2, 3], inputs, fused=True, relu=relu, groups=BN_GROUPS, up2=up2)
    else:
        output = nonlinearity(inputs) if relu else inputs
        return F.upsample_2x(output) if up2 else output


COMMUTE_1X1 = True   # evaluate 1x1 shortcut convs on the low-resolution side of their resampling (same function)
FUSE_SKIP_ADD = True  # shortcut + conv_2(...) inside conv_2's epilogue where conv_2 is not followed by pooling
FUSE_D_ACT = True     # critic: relu in conv_1's epilogue, dropout -> (skip, relu) forks as one node (see functional.Fork*)
# the ReLU backward inside conv_2's dgrad epilogue (ConvF in_relu / relu_bwd_fused; EPI_MASK variant of the lean / pair
# kernels): 8 fewer mask-multiply launches per critic step (993 -> 979 us)
FUSE_RELU_BWD = True
FUSE_POOL_FORK = True   # critic blocks 1, 2: mean pool + skip add + (dropout) + next relu as one kernel (functional.PoolAddFork)
FUSE_BN_UP = True       # generator blocks: Normalize + relu writes its output already upsampled (the UpsampleConv input)


def ConvMeanPool(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True, in_relu=False):
    if in_relu:
        output = lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases,
                                       in_relu=True)
        return F.mean_pool_2x2(output)
    if filter_size == 1 and COMMUTE_1X1:
        # a 1x1 conv commutes with the 2x2 mean (both linear, the bias is constant over the window): pool first,
        # convolve a quarter of the pixels
        return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, F.mean_pool_2x2(inputs), he_init=he_init,
                                     biases=biases)
    output = lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
    return F.mean_pool_2x2(output)


def MeanPoolConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    output = F.mean_pool_2x2(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)


def UpsampleConv(name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
    if filter_size == 1 and COMMUTE_1X1:
        # a 1x1 conv of a nearest-neighbour upsampled map == the upsampled 1x1 conv (every output pixel sees the same
        # input pixel): convolve at the low resolution, bit-identical result for a quarter of the work
        output = lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
        return F.upsample_2x(output)
    output = F.upsample_2x(inputs)
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)


def _plain_relu_block(name):
    """True for blocks whose Normalize() is just the nonlinearity (the critic: NORMALIZATION_D is False)."""
    return FUSE_D_ACT and ('Discriminator' in name) and not NORMALIZATION_D


def ResidualBlock(name, input_dim, output_dim, filter_size, inputs, resample=None, no_dropout=False, labels=None,
                  pre_act=None, fork_keep=None):
    """
    resample: None, 'down', or 'up'
    pre_act (extension): relu(inputs) when the caller already computed it together with `inputs` (F.fork_dropout_relu)
    fork_keep (extension, 'down' critic blocks): return (d, relu(d)) with d = dropout(block output, keep=fork_keep) -- the
        mean pool, the skip add, the dropout that follows the block and the next block's first relu as ONE kernel
    """
    if resample == 'down':
        conv_1 = functools.partial(lib.ops.conv2d.Conv2D, input_dim=input_dim, output_dim=input_dim)
        conv_2 = functools.partial(ConvMeanPool, input_dim=input_dim, output_dim=output_dim)
        conv_shortcut = ConvMeanPool
    elif resample == 'up':
        conv_1 = functools.partial(UpsampleConv, input_dim=input_dim, output_dim=output_dim)
        conv_shortcut = UpsampleConv
        conv_2 = functools.partial(lib.ops.conv2d.Conv2D, input_dim=output_dim, output_dim=output_dim)
    elif resample is None:
        conv_shortcut = lib.ops.conv2d.Conv2D
        conv_1 = functools.partial(lib.ops.conv2d.Conv2D, input_dim=input_dim, output_dim=output_dim)
        conv_2 = functools.partial(lib.ops.conv2d.Conv2D, input_dim=output_dim, output_dim=output_dim)
    else:
        raise Exception('invalid resample value')

    fused_act = _plain_relu_block(name) and resample != 'up'
    if fused_act and pre_act is None:
        # N1 = relu(inputs) and the skip connection both read `inputs`: one node, so that backward is one kernel
        inputs, pre_act = F.fork_relu(inputs)

    if not fused_act and resample == 'up':
        inputs, inputs_main = F.fork2(inputs)            # block input -> (shortcut, N1): explicit fork, adjoint = K.add
    else:
        inputs_main = inputs
    shortcut_up2 = False
    if output_dim == input_dim and resample is None:
        shortcut = inputs  # Identity skip-connection
    elif resample == 'up' and FUSE_SKIP_ADD and COMMUTE_1X1:
        # UpsampleConv shortcut: the 1x1 conv at the low resolution; its nearest-neighbour upsample is never materialised,
        # conv_2's epilogue adds shortcut[h/2, w/2]
        shortcut = lib.ops.conv2d.Conv2D(name + '.Shortcut', input_dim, output_dim, 1, inputs, he_init=False, biases=True)
        shortcut_up2 = True
    else:
        shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim, filter_size=1,
                                 he_init=False, biases=True, inputs=inputs)

    if fused_act:
        # N2 = relu(conv_1(.)) in conv_1's epilogue; its backward ([output > 0]) in conv_2's dgrad epilogue
        if resample == 'down' and _pool_conv_fused(pre_act, input_dim, input_dim, output_dim, filter_size):
            # conv_1 hands relu(conv_1) over in the space-to-depth layout; conv_2 + mean pool + skip add are ONE stride-2 conv
            output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=pre_act, relu=True, relu_bwd_fused=True, out_s2d=True)
            output = lib.ops.conv2d.Conv2D(name + '.Conv2', input_dim, output_dim, filter_size, output, mean_pool=True,
                                           residual=shortcut)
            return _dropout_relu(output, fork_keep) if fork_keep is not None else output
        output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=pre_act, relu=True, relu_bwd_fused=FUSE_RELU_BWD)
        if resample == 'down' and fork_keep is not None:
            full = lib.ops.conv2d.Conv2D(name + '.Conv2', input_dim, output_dim, filter_size, output, in_relu=FUSE_RELU_BWD)
            return _pool_add_fork(full, shortcut, fork_keep)
        if resample != 'down' and FUSE_SKIP_ADD:
            return conv_2(name + '.Conv2', filter_size=filter_size, inputs=output, residual=shortcut, in_relu=FUSE_RELU_BWD)
        output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output, in_relu=FUSE_RELU_BWD)
        return F.add(shortcut, output)
    else:
        output = inputs_main
        if resample == 'up' and FUSE_BN_UP:
            # UpsampleConv(N1(x)): the normalisation kernel writes the 2x nearest-neighbour upsampled activation itself
            output = Normalize(name + '.N1', output, labels=labels, relu=True, up2=True)
            output = lib.ops.conv2d.Conv2D(name + '.Conv1', input_dim, output_dim, filter_size, output)
        else:
            output = Normalize(name + '.N1', output, labels=labels, relu=True)
            output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output)
        output = Normalize(name + '.N2', output, labels=labels, relu=True)
    if resample != 'down' and FUSE_SKIP_ADD:
        # conv_2 is a plain Conv2D at the shortcut's resolution: the skip connection is added in its epilogue
        return conv_2(name + '.Conv2', filter_size=filter_size, inputs=output, residual=shortcut, residual_up2=shortcut_up2)
    output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output)

    return F.add(shortcut, output)


def _pool_conv_fused(x, cin, cmid, cout, filter_size):
    """True when [3x3 conv cin -> cmid, relu, ConvMeanPool(3x3) cmid -> cout] on activation x takes the fused route: the
    first conv writes the space-to-depth layout, the second and its mean pool run as one stride-2 4x4 conv."""
    return (K.config.pool_conv_s2d and FUSE_D_ACT and FUSE_RELU_BWD and FUSE_SKIP_ADD and filter_size == 3 and not isinstance(x, F.S2DAct)
            and F.conv2d_s2d_out_ok(x, cmid, 3) and F.conv_mean_pool_s2d_ok(x, cmid, cout))


def _pool_add_fork(full, shortcut, keep):
    """(d, relu(d)), d = dropout(meanpool(full) + shortcut)"""
    if keep == 1.0:
        return F.pool_add_fork(full, shortcut)
    return F.pool_add_fork(full, shortcut, keep, **RNG.dropout_args(shortcut))


def OptimizedResBlockDisc1(inputs, fork=False):
    conv_1 = functools.partial(lib.ops.conv2d.Conv2D, input_dim=3, output_dim=DIM_D)
    conv_2 = functools.partial(ConvMeanPool, input_dim=DIM_D, output_dim=DIM_D)
    conv_shortcut = MeanPoolConv
    inputs, inputs_main = F.fork2(inputs)                # image -> (shortcut, conv_1): explicit fork (x^ and fakes carry gradients)
    shortcut = conv_shortcut('Discriminator.1.Shortcut', input_dim=3, output_dim=DIM_D, filter_size=1, he_init=False,
                             biases=True, inputs=inputs)

    output = inputs_main
    if FUSE_D_ACT and _pool_conv_fused(output, 3, DIM_D, DIM_D, 3):
        output = conv_1('Discriminator.1.Conv1', filter_size=3, inputs=output, relu=True, relu_bwd_fused=True, out_s2d=True)
        output = lib.ops.conv2d.Conv2D('Discriminator.1.Conv2', DIM_D, DIM_D, 3, output, mean_pool=True, residual=shortcut)
        return F.fork_relu(output) if fork else output
    if FUSE_D_ACT:
        # nonlinearity in conv_1's epilogue, its backward in conv_2's dgrad epilogue
        output = conv_1('Discriminator.1.Conv1', filter_size=3, inputs=output, relu=True, relu_bwd_fused=FUSE_RELU_BWD)
        if fork:      # -> (block output, relu(block output)): pool + skip add + the next block's first relu in one kernel
            full = lib.ops.conv2d.Conv2D('Discriminator.1.Conv2', DIM_D, DIM_D, 3, output, in_relu=FUSE_RELU_BWD)
            return F.pool_add_fork(full, shortcut)
        output = conv_2('Discriminator.1.Conv2', filter_size=3, inputs=output, in_relu=FUSE_RELU_BWD)
    else:
        output = conv_1('Discriminator.1.Conv1', filter_size=3, inputs=output)
        output = nonlinearity(output)
        output = conv_2('Discriminator.1.Conv2', filter_size=3, inputs=output)
    return F.add(shortcut, output)


def Generator(n_samples, labels, noise=None):
    if noise is None:
        noise = RNG.normal(RNG._scope + '.z', (n_samples, 128))
    noise = F.cast(noise, ACT_DTYPE)
    output = lib.ops.linear.Linear('Generator.Input', 128, 4 * 4 * DIM_G, noise)
    output = F.to_nhwc(output, DIM_G, 4, 4, ACT_DTYPE)           # tf.reshape(output, [-1, DIM_G, 4, 4])
    output = ResidualBlock('Generator.1', DIM_G, DIM_G, 3, output, resample='up', labels=labels)
    output = ResidualBlock('Generator.2', DIM_G, DIM_G, 3, output, resample='up', labels=labels)
    output = ResidualBlock('Generator.3', DIM_G, DIM_G, 3, output, resample='up', labels=labels)
    output = Normalize('Generator.OutputN', output, relu=True)
    output = lib.ops.conv2d.Conv2D('Generator.Output', DIM_G, 3, 3, output, he_init=False)
    # the 3-channel image leaves the bf16 domain BEFORE the tanh: the generator's output (the critic's input, and what
    # the parity tests compare) carries one rounding less; the [B, 3072] tensor is tiny
    output = F.to_flat_nchw(output, torch.float32)                # tf.reshape(output, [-1, OUTPUT_DIM])
    return F.tanh(output)


def _dropout(output, keep):
    if keep == 1.0:
        return output
    return F.dropout(output, keep, **RNG.dropout_args(output))


def _dropout_relu(output, keep):
    """(dropout(x), relu(dropout(x)))"""
    if keep == 1.0:
        return F.fork_relu(output)
    return F.fork_dropout_relu(output, keep, **RNG.dropout_args(output))


def Discriminator(inputs, labels, kp1, kp2, kp3):  # three more parameters of keep rate
    output = F.to_nhwc(inputs, 3, 32, 32, ACT_DTYPE)              # tf.reshape(inputs, [-1, 3, 32, 32])
    if FUSE_D_ACT and FUSE_POOL_FORK and not NORMALIZATION_D:
        output, act = OptimizedResBlockDisc1(output, fork=True)
        output, act = ResidualBlock('Discriminator.2', DIM_D, DIM_D, 3, output, resample='down', labels=labels, pre_act=act,
                                    fork_keep=kp1)
    else:
        output = OptimizedResBlockDisc1(output)
        output = ResidualBlock('Discriminator.2', DIM_D, DIM_D, 3, output, resample='down', labels=labels)
        act = None
    if FUSE_D_ACT and not NORMALIZATION_D:
        # dropout -> (skip connection, first relu of the next block) as one node; the last dropout + relu as one kernel
        # (relu and dropout commute: both multiply by a non-negative constant)
        if act is None:
            output, act = _dropout_relu(output, kp1)
        output = ResidualBlock('Discriminator.3', DIM_D, DIM_D, 3, output, resample=None, labels=labels, pre_act=act)
        output, act = _dropout_relu(output, kp2)
        output = ResidualBlock('Discriminator.4', DIM_D, DIM_D, 3, output, resample=None, labels=labels, pre_act=act)
        output = F.relu(output) if kp3 == 1.0 else F.leaky_relu_dropout(output, 0.0, kp3, **RNG.dropout_args(output))
    else:
        output = _dropout(output, kp1)  # dropout after activator
        output = ResidualBlock('Discriminator.3', DIM_D, DIM_D, 3, output, resample=None, labels=labels)
        output = _dropout(output, kp2)  # dropout after activator
        output = ResidualBlock('Discriminator.4', DIM_D, DIM_D, 3, output, resample=None, labels=labels)
        output = _dropout(output, kp3)  # dropout after activator
        output = nonlinearity(output)
    output2 = F.spatial_mean(output)  # corresponding to D_
    # D_ feeds the critic head, the ACGAN head and (returned) the consistency term: explicit forks, so that the three
    # gradients are summed by the library's add kernel
    output2, feat = F.fork2(output2)
    if CONDITIONAL and ACGAN:
        feat, feat_ac = F.fork2(feat)
    output_wgan = lib.ops.linear.Linear('Discriminator.Output', DIM_D, 1, feat, out_dtype=torch.float32)
    output_wgan = output_wgan.reshape(-1)  # conrresponding to D
    if CONDITIONAL and ACGAN:
        output_acgan = lib.ops.linear.Linear('Discriminator.ACGANOutput', DIM_D, 10, feat_ac, out_dtype=torch.float32)
        return output_wgan, output2, output_acgan
    else:
        return output_wgan, output2, None  # two layers' of output


class Trainer:
    """Owns the parameters, both optimizers and the random stream of one training process."""

    def __init__(self, device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=BATCH_SIZE, record=False,
                 graph_safe_rng=False):
        global ACT_DTYPE, RNG
        ACT_DTYPE = act_dtype
        self.device = torch.device(device)
        self.B = batch_size
        lib.delete_all_params()           # one model per process, like the reference's module-level dict
        lib.set_device(self.device)
        self.rng = RNG = DeviceRandom(seed, self.device, record=record, graph_safe=graph_safe_rng)
        # build every parameter in the reference's graph-construction order (G first, then D)
        with torch.no_grad():
            RNG.scope('build')
            labels = torch.zeros(2, dtype=torch.int32, device=self.device)
            fake = Generator(2, labels)
            Discriminator(fake, labels, 1.0, 1.0, 1.0)
        self.rng.offset = 0
        self.gen_opt = FlatAdam('Generator', LR, 0.0, 0.9)          # :333, var_list :336
        self.disc_opt = FlatAdam('Discriminator.', LR, 0.0, 0.9)    # :334, var_list :302
        self.hp = dict(lambda_gp=LAMBDA, lambda2=LAMBDA_2, factor_m=Factor_M,
                       acgan_scale=ACGAN_SCALE if (CONDITIONAL and ACGAN) else 0.0)

    def activate(self):
        global ACT_DTYPE, RNG
        RNG = self.rng

    def _ones_like(self, t):
        """A constant tensor of ones shaped like t, created once (tf.gradients' implicit grad_ys; no fill kernel per step)."""
        key = (tuple(t.shape), t.dtype)
        cache = self.__dict__.setdefault('_ones', {})
        if key not in cache:
            cache[key] = torch.ones_like(t)
        return cache[key]

    @staticmethod
    def lr(iteration):
        decay = max(0., 1. - (float(iteration) / ITERS)) if DECAY else 1.
        return LR * decay

    # ---------------------------------------------------------------- critic (disc_train_op, :190-300,336-338)
    def _generate(self, parts_z, labels, n_total):
        """The reference calls Generator once per device split (:196-199, :318-321); here the splits run as ONE
        batch whose batch-norm statistics are computed per split (groups) -- same numbers, half the launches."""
        global BN_GROUPS
        RNG = self.rng
        noise = RNG.normal_parts(parts_z, 128)
        BN_GROUPS = len(parts_z)
        try:
            return Generator(n_total, labels, noise=noise)
        finally:
            BN_GROUPS = 1

    def generate_fakes(self, labels_steps):
        """Fake batches for S consecutive critic steps in ONE generator forward (the generator does not change
        between the CRITIC_ITERS critic steps of an iteration, :384-394): labels_steps = the S*B real labels,
        step-major.  Batch-norm statistics stay per 32-sample device split (S*N_DEVICES groups), so every step's
        fakes are what its own Generator call (:196-199) would produce from the same noise."""
        S, h = labels_steps.shape[0] // self.B, self.B // N_DEVICES
        with torch.no_grad():
            return self._generate([('z.%d.%d' % (k, i), h) for k in range(S) for i in range(N_DEVICES)], labels_steps,
                                  S * self.B)

    def critic_forward_backward(self, all_real_data_int, all_real_labels, with_metrics=False, fake_data=None):
        # two stream branches share the SMs in this step: sub-wave layers keep one CTA per tile (see kernels.splitk)
        with K.splitk(K.config.critic_splitk or not K.config.branch_streams):
            return self._critic_forward_backward(all_real_data_int, all_real_labels, with_metrics, fake_data)

    def _critic_forward_backward(self, all_real_data_int, all_real_labels, with_metrics=False, fake_data=None):
        RNG = self.rng
        B = all_real_data_int.shape[0]
        h = B // N_DEVICES
        if fake_data is None:
            with torch.no_grad():
                RNG.begin_stack([h] * N_DEVICES)
                fake_data = self._generate([('z.%d' % i, h) for i in range(N_DEVICES)], all_real_labels, B)
                RNG.end_stack()
        # stochastic pass ' on real+fake (2B rows) and pass '' on the real half (B rows) as ONE critic call: same weights,
        # independent dropout draws per row -- the reference's two calls at :226-227.  The stacked input [real, fake, real] is
        # written in place: the input-scaling kernel stores the real batch into both of its row ranges, the fakes are copied
        # into theirs (no concatenation kernel).
        stacked = torch.empty((3 * B, all_real_data_int.shape[1]), dtype=torch.float32, device=all_real_data_int.device)
        all_real_data = stacked[:B]
        if RNG.replay is not None:                                                              # golden-vector tests
            all_real_data.copy_(K.add(K.prep_real(all_real_data_int, 256., 0.), RNG.uniform('dequant', all_real_data_int.shape)))
            stacked[2 * B:].copy_(all_real_data)
        else:
            seed, off, dyn = RNG.stream('dequant', all_real_data_int)
            K.prep_real(all_real_data_int, 256., 1. / 128, seed, off, dyn=dyn, out=all_real_data, out2=stacked[2 * B:])  # :201-202
        stacked[B:2 * B].copy_(fake_data)
        fork = K.fork_branch(all_real_data)          # the gradient-penalty pass below depends on nothing after this point
        stacked_labels = None     # the critic never reads labels (NORMALIZATION_D is False; ACGAN conditions through its loss)
        RNG.scope_parts([('drop.p1', 2 * B), ('drop.p2', B)])
        RNG.begin_stack([2 * B, B])
        with (K.branch(fork) if K.config.branch_stacked else contextlib.nullcontext()):
            disc_all, disc_all_2, disc_all_acgan = Discriminator(stacked, stacked_labels, 0.8, 0.5, 0.5)
        RNG.end_stack()
        use_logits = CONDITIONAL and ACGAN
        rows = dict(real=(0, B), fake=(B, 2 * B), real2=(2 * B, 3 * B))
        decoupled = K.config.decouple_gp and not K.config.branch_stacked
        if decoupled:
            # the WGAN, CT and ACGAN terms depend on the stacked pass only: its backward starts now, on this stream, while the
            # penalty branch is still in its forward / first backward (the two halves of the loss: functional.GPLoss)
            out_a = F.CTGPLossStacked.apply(disc_all, disc_all_2, None, disc_all_acgan if use_logits else None,
                                            all_real_labels if use_logits else None, self.hp, rows)
            out_a[0].backward(gradient=self._ones_like(out_a[0]), inputs=self.disc_opt.param_list())
        metrics = {}
        if with_metrics and CONDITIONAL and ACGAN:
            with torch.no_grad():
                _, _, clean = Discriminator(stacked[:2 * B], None, 1.0, 1.0, 1.0)
                pred = clean.argmax(dim=1).to(torch.int32)
                metrics['acgan_acc'] = (pred[:B] == all_real_labels).float().mean()
                metrics['acgan_fake_acc'] = (pred[B:] == all_real_labels).float().mean()
        # gradient-penalty pass: independent of the stacked pass until the loss, so it runs as a second branch
        # (stream / CUDA-graph branch); autograd replays each branch's backward on the stream of its forward
        with (contextlib.nullcontext() if K.config.branch_stacked else K.branch(fork)):
            alpha = RNG.uniform('alpha', (B, 1))
            interpolates = K.interpolate(all_real_data, fake_data, alpha).requires_grad_(True)     # :277-283
            RNG.scope('drop.gp')
            d_interp = Discriminator(interpolates, all_real_labels, 0.8, 0.5, 0.5)[0]
            with F.no_param_grads():                  # tf.gradients(..., [interpolates]): d/dx^ only
                gradients = torch.autograd.grad(d_interp, interpolates, grad_outputs=self._ones_like(d_interp),
                                                create_graph=True)[0]                               # :284
            if decoupled:
                out_b = F.GPLoss.apply(gradients, self.hp)
                out_b[0].backward(gradient=self._ones_like(out_b[0]), inputs=self.disc_opt.param_list())
        K.join_branch(fork)
        if decoupled:
            out = K.add(out_a.detach(), out_b.detach())        # {cost, wgan, ct, gp, acgan}: the halves fill disjoint terms
        else:
            # WGAN term and CE use pass ' (rows [0, 2B)), the CT term the real halves of ' and '' (:244-300)
            out = F.CTGPLossStacked.apply(disc_all, disc_all_2, gradients, disc_all_acgan if use_logits else None,
                                          all_real_labels if use_logits else None, self.hp, rows)
            out[0].backward(gradient=self._ones_like(out[0]), inputs=self.disc_opt.param_list())
            K.join_branch(fork)
        K.join_side()
        res = dict(out=out.detach(), gradients=gradients.detach(), fake_data=fake_data, real_data=all_real_data)
        res.update(metrics)
        return res

    def acgan_accuracy(self, all_real_data_int, all_real_labels, fake_data):
        """disc_acgan_acc / disc_acgan_fake_acc (:297-298, plotted as 'acc_real' / 'acc_fake' :410-411): the critic WITHOUT
        dropout (keep probabilities 1: no random draws) on the real batch and on `fake_data` (the fakes of a critic step,
        conditioned on the same labels).  Metrics only -- nothing of the training step depends on it; the loop runs it every
        `acc_every` iterations instead of inside every critic step as the reference graph does."""
        B = all_real_data_int.shape[0]
        with torch.no_grad():
            both = torch.empty((2 * B, all_real_data_int.shape[1]), dtype=torch.float32, device=all_real_data_int.device)
            K.prep_real(all_real_data_int, 256., 0., out=both[:B])
            both[B:].copy_(fake_data)
            _, _, clean = Discriminator(both, None, 1.0, 1.0, 1.0)
            pred = clean.argmax(dim=1).to(torch.int32)
            return (pred[:B] == all_real_labels).float().mean(), (pred[B:] == all_real_labels).float().mean()

    def critic_step(self, all_real_data_int, all_real_labels, iteration=0, with_metrics=False, use_device_lr=False,
                    fake_data=None):
        self.disc_opt.zero_grad()
        res = self.critic_forward_backward(all_real_data_int, all_real_labels, with_metrics, fake_data=fake_data)
        self.rng.end_step(side=True)                  # the Philox counter advances next to the optimizer kernels
        world = self.disc_opt.all_reduce()
        self.disc_opt.step(self.lr(iteration), world, use_device_lr=use_device_lr)
        K.join_side()
        return res

    # ---------------------------------------------------------------- generator (gen_train_op, :314-330,335,337)
    def gen_forward_backward(self):
        if K.config.gen_towers and K.config.branch_streams and self.device.type == 'cuda' and N_DEVICES > 1:
            with K.splitk(K.config.gen_splitk):
                return self._gen_towers()
        RNG = self.rng
        n = GEN_BS_MULTIPLE * self.B // N_DEVICES
        fake_labels = RNG.labels_parts([('labels.%d' % i, n) for i in range(N_DEVICES)])
        RNG.begin_stack([n] * N_DEVICES)
        fake = self._generate([('z.%d' % i, n) for i in range(N_DEVICES)], fake_labels, n * N_DEVICES)
        RNG.scope_parts([('drop.%d' % i, n) for i in range(N_DEVICES)])
        disc_fake, _, disc_fake_acgan = Discriminator(fake, fake_labels, 0.8, 0.5, 0.5)
        RNG.end_stack()
        # mean over the stacked batch == (cost_dev0 + cost_dev1) / len(DEVICES)  (equal split sizes)
        gen_cost = F.MeanLoss.apply(disc_fake, -1.0)
        if CONDITIONAL and ACGAN:
            gen_cost = F.AddScaled.apply(gen_cost, F.SoftmaxCE.apply(disc_fake_acgan, fake_labels), ACGAN_SCALE_G)
        with F.frozen(self.disc_opt.param_list()):    # var_list = gen_params (:336): the critic is not updated here
            gen_cost.backward(gradient=self._ones_like(gen_cost), inputs=self.gen_opt.param_list())
        K.join_side()
        return dict(cost=gen_cost.detach())

    def _gen_towers(self):
        """The generator step as the reference builds it (:314-330): one tower per device split -- Generator(n) -> Discriminator
        -> cost -- here on alternating stream branches.  The step is a single dependent chain of mostly sub-wave kernels
        (8x8 / 16x16 layers at 64-128 samples), so two independent towers fill the SMs the way the gradient-penalty branch
        does in the critic step; the filter gradients of both towers are queued and run as one multi-job launch at the end."""
        RNG = self.rng
        n = GEN_BS_MULTIPLE * self.B // N_DEVICES
        # labels, noise and every dropout site draw the numbers the stacked batch would (DeviceRandom.scope_tower)
        all_labels = RNG.labels_parts([('labels.%d' % i, n) for i in range(N_DEVICES)])
        all_noise = RNG.normal_parts([('z.%d' % i, n) for i in range(N_DEVICES)], 128)
        fork = K.fork_branch(all_noise)
        costs, part = [], 1.0 / N_DEVICES
        for i in range(N_DEVICES):
            with (K.branch(fork) if i % 2 else contextlib.nullcontext()):
                labels = all_labels[i * n:(i + 1) * n]
                fake = Generator(n, labels, noise=all_noise[i * n:(i + 1) * n])
                RNG.scope_tower('drop', i, N_DEVICES)
                disc_fake, _, disc_fake_acgan = Discriminator(fake, labels, 0.8, 0.5, 0.5)
                cost = F.MeanLoss.apply(disc_fake, -part)
                if CONDITIONAL and ACGAN:
                    cost = F.AddScaled.apply(cost, F.SoftmaxCE.apply(disc_fake_acgan, labels), ACGAN_SCALE_G * part)
                costs.append(cost)
        K.join_branch(fork)
        gen_cost = costs[0]                            # (cost_dev0 + cost_dev1) / len(DEVICES)  (:330)
        for cost in costs[1:]:
            gen_cost = F.AddScaled.apply(gen_cost, cost, 1.0)
        with F.frozen(self.disc_opt.param_list()):    # var_list = gen_params (:336): the critic is not updated here
            gen_cost.backward(gradient=self._ones_like(gen_cost), inputs=self.gen_opt.param_list())
        K.join_branch(fork)
        K.join_side()
        return dict(cost=gen_cost.detach())

    def gen_step(self, iteration=0, use_device_lr=False):
        self.gen_opt.zero_grad()
        res = self.gen_forward_backward()
        self.rng.end_step(side=True)
        world = self.gen_opt.all_reduce()
        self.gen_opt.step(self.lr(iteration), world, use_device_lr=use_device_lr)
        K.join_side()
        return res
