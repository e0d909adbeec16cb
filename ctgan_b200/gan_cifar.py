"""CT-GAN DCGAN for CIFAR-10: the training step of TG/CT_gan_cifar.py (MODE='wgan-CT') on B200.

Hyper-parameters :34-43, Generator :58-79, Discriminator :81-100, input scaling :102-103,
loss :123-151, Adam :153-154.  Of the four critic calls the reference builds (:107-110)
`disc_fake_2` feeds nothing and is pruned by TF; three stochastic passes (real', real'',
fake) plus the gradient-penalty pass remain.
"""
import torch

from . import tflib as lib
from . import functional as F
from . import kernels as K
from .tflib.ops import linear as _linear, conv2d as _conv2d, batchnorm as _batchnorm, deconv2d as _deconv2d
from .runtime import DeviceRandom, FlatAdam

n_examples = 1000  # number of examples for training
LAMBDA_2 = 2.0  # weight facter
Factor_M = 0.0  # M
MODE = 'wgan-CT'
DIM = 128  # This overfits substantially; you're probably better off with 64
LAMBDA = 10  # Gradient penalty lambda hyperparameter
CRITIC_ITERS = 5  # How many critic iterations per generator iteration
BATCH_SIZE = 64  # Batch size
ITERS = 50000  # How many generator iterations to train for
OUTPUT_DIM = 3072  # Number of pixels in CIFAR10 (3*32*32)
IMG_C, IMG_HW = 3, 32

ACT_DTYPE = torch.bfloat16
RNG = None
HEAD_NHWC = True     # the critic head reads the last feature map in its own (h, w, c) order (Discriminator below)


def LeakyReLU(x, alpha=0.2):
    return F.leaky_relu_dropout(x, alpha, 1.0)


def _lrelu_dropout(output, keep):
    """LeakyReLU (:47-48) followed by tf.nn.dropout(keep_prob) (:86): one fused kernel."""
    return F.leaky_relu_dropout(output, 0.2, keep, **RNG.dropout_args(output))


def _conv_lrelu_dropout(name, input_dim, output_dim, inputs, keep, next_cout=None):
    """lib.ops.conv2d.Conv2D(name, ., ., 5, ., stride=2) -> LeakyReLU -> tf.nn.dropout(keep_prob=keep)"""
    return lib.ops.conv2d.Conv2D(name, input_dim, output_dim, 5, inputs, stride=2,
                                 act_dropout=dict(slope=0.2, keep=keep, rng=RNG, next_cout=next_cout, next_k=5))


def Generator(n_samples, noise=None):
    if noise is None:
        noise = RNG.normal('z', (n_samples, 128))
    noise = F.cast(noise, ACT_DTYPE)
    # the Linear output stays float until it is normalised: rounding it to bf16 BEFORE the mean subtraction of BN1 would be
    # amplified by |mean| / std (64 samples per feature); the [64, 8192] tensor is tiny
    output = lib.ops.linear.Linear('Generator.Input', 128, 4 * 4 * 4 * DIM, noise, out_dtype=torch.float32)
    output = lib.ops.batchnorm.Batchnorm('Generator.BN1', [0], output, relu=True)
    output = F.to_nhwc(output, 4 * DIM, 4, 4, ACT_DTYPE)

    output = lib.ops.deconv2d.Deconv2D('Generator.2', 4 * DIM, 2 * DIM, 5, output)
    output = lib.ops.batchnorm.Batchnorm('Generator.BN2', [0, 2, 3], output, relu=True)

    output = lib.ops.deconv2d.Deconv2D('Generator.3', 2 * DIM, DIM, 5, output)
    output = lib.ops.batchnorm.Batchnorm('Generator.BN3', [0, 2, 3], output, relu=True)

    output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 3, 5, output)
    # the image leaves the bf16 domain BEFORE the tanh: one rounding less on the generator's output (a tiny tensor)
    output = F.to_flat_nchw(output, torch.float32)
    return F.tanh(output)


def Discriminator(inputs):
    output = F.to_nhwc(inputs, 3, 32, 32, ACT_DTYPE)

    # Conv2D -> LeakyReLU -> dropout (:84-96) as ONE kernel per layer: bias + LeakyReLU + Philox dropout in the conv
    # epilogue, the result written directly in the space-to-depth layout the next stride-2 layer reads
    output = _conv_lrelu_dropout('Discriminator.1', 3, DIM, output, 0.50, next_cout=2 * DIM)
    output = _conv_lrelu_dropout('Discriminator.2', DIM, 2 * DIM, output, 0.50, next_cout=4 * DIM)
    output = _conv_lrelu_dropout('Discriminator.3', 2 * DIM, 4 * DIM, output, 0.50)
    if HEAD_NHWC:
        # D_ (the features of the consistency term, a mean over them: :133) and the head's input stay in the activation's own
        # (h, w, c) order; the head re-orders its 4*4*4*DIM weights instead (lib.ops.linear.Linear(input_nhwc=...)) -- no
        # transposition of the activations / their gradients in any of the step's passes
        output2 = F.flat_nhwc(output)
        output = lib.ops.linear.Linear('Discriminator.Output', 4 * 4 * 4 * DIM, 1, output2, out_dtype=torch.float32,
                                       input_nhwc=(4 * DIM, 4, 4))
        return output.reshape(-1), output2
    output2 = F.to_flat_nchw(output)  # corresponding to D_  (tf.reshape(output, [-1, 4*4*4*DIM]))
    output = lib.ops.linear.Linear('Discriminator.Output', 4 * 4 * 4 * DIM, 1, output2, out_dtype=torch.float32)
    return output.reshape(-1), output2


class Trainer:
    """Parameters, optimizers and random stream of one training process (DCGAN scripts)."""
    _module = None     # set below; gan_mnist re-uses this class with its own module globals

    def __init__(self, device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=None, record=False,
                 graph_safe_rng=False):
        m = self._mod()
        m.ACT_DTYPE = act_dtype
        self.device = torch.device(device)
        self.B = batch_size or m.BATCH_SIZE
        lib.delete_all_params()           # one model per process, like the reference's module-level dict
        lib.set_device(self.device)
        self.rng = m.RNG = DeviceRandom(seed, self.device, record=record, graph_safe=graph_safe_rng)
        with torch.no_grad():
            m.RNG.scope('build')
            m.Discriminator(m.Generator(2))
        self.rng.offset = 0
        self.gen_opt = FlatAdam('Generator', 1e-4, 0.5, 0.9)
        self.disc_opt = FlatAdam('Discriminator', 1e-4, 0.5, 0.9)
        self.hp = dict(lambda_gp=float(m.LAMBDA), lambda2=m.LAMBDA_2, factor_m=m.Factor_M, acgan_scale=0.0)

    @classmethod
    def _mod(cls):
        import sys
        return sys.modules[cls.__module__]

    def activate(self):
        self._mod().RNG = self.rng

    def prep_real(self, real_data_in):
        return K.prep_real(real_data_in, 255., 0.)                     # 2*((x/255.)-.5), :102-103

    def critic_forward_backward(self, real_data_in):
        m, RNG = self._mod(), self.rng
        real_data = self.prep_real(real_data_in)
        B = real_data.shape[0]
        with torch.no_grad():
            fake_data = m.Generator(B, noise=RNG.normal('z', (B, 128)))
        # the three stochastic critic calls of the reference (real', real'', fake) as ONE stacked batch:
        # shared weights, independent dropout draws per row
        fork = K.fork_branch(real_data)               # the gradient-penalty pass depends on nothing after this point
        stacked = torch.cat([real_data, real_data, fake_data], dim=0)
        RNG.scope_parts([('drop.real1', B), ('drop.real2', B), ('drop.fake', B)])
        RNG.begin_stack([B, B, B])
        d_all, f_all = m.Discriminator(stacked)
        RNG.end_stack()
        with K.branch(fork):                          # second stream / CUDA-graph branch (see gan_cifar_resnet.py)
            alpha = RNG.uniform('alpha', (B, 1))
            interpolates = K.interpolate(real_data, fake_data, alpha).requires_grad_(True)
            RNG.scope('drop.gp')
            d_interp = m.Discriminator(interpolates)[0]
            with F.no_param_grads():                  # tf.gradients(..., [interpolates]): d/dx^ only (:144)
                gradients = torch.autograd.grad(d_interp, interpolates, grad_outputs=torch.ones_like(d_interp),
                                                create_graph=True)[0]
        K.join_branch(fork)
        out = F.CTGPLossStacked.apply(d_all, f_all, gradients, None, None, self.hp,
                                      dict(real=(0, B), real2=(B, 2 * B), fake=(2 * B, 3 * B)))
        out[0].backward(inputs=self.disc_opt.param_list())
        K.join_branch(fork)
        K.join_side()
        return dict(out=out.detach(), gradients=gradients.detach(), fake_data=fake_data, real_data=real_data)

    def critic_step(self, real_data_in, iteration=0, use_device_lr=False):
        self.disc_opt.zero_grad()
        res = self.critic_forward_backward(real_data_in)
        world = self.disc_opt.all_reduce()
        self.disc_opt.step(None, world, use_device_lr=use_device_lr)
        self.rng.end_step()
        return res

    def gen_forward_backward(self):
        m, RNG = self._mod(), self.rng
        fake_data = m.Generator(self.B, noise=RNG.normal('z', (self.B, 128)))
        RNG.scope('drop.fake')
        disc_fake, _ = m.Discriminator(fake_data)
        gen_cost = F.MeanLoss.apply(disc_fake, -1.0)
        with F.frozen(self.disc_opt.param_list()):    # var_list=gen_params (:153): the critic is differentiated through, not updated
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
