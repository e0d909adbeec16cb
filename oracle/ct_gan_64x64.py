"""Oracle restatement of TG/CT_gan_64x64.py (MODE='wgan-ct', GoodGenerator / GoodDiscriminator) — TEST INFRASTRUCTURE.

TG = /root/reference/CT-GANs/tensorflow_generative_model.  Hyper-parameters :28-37, architecture choice :41-72 (the
shipped default returns GoodGenerator, GoodDiscriminator :48), Normalize :87-93 (layer norm over [1,2,3] in the critic,
fused batch norm in the generator), ConvMeanPool / MeanPoolConv / UpsampleConv :106-124, ResidualBlock :166-200,
GoodGenerator :204-221, GoodDiscriminator :357-373, per-device loss graph :480-543 (wgan-ct branch :494-519), tower
average :545-546, Adam(1e-4, beta1=0, beta2=.9) :560-564.  The batch is split over N_GPUS = 2 towers (:480): each tower
sees BATCH_SIZE/2 reals, draws its own fakes / masks / alphas, and the two tower costs are averaged; batch-norm
statistics of the generator are therefore per 32-sample tower.  SURVEY.md 8(f) row N4.
"""
import functools

import torch

from . import tf_ops
from .tflib_ref import TFLib
from .ct_gan_common import StepMixin, consistency_term, gradient_penalty

LAMBDA_2 = 2.0
Factor_M = 0.0
MODE = 'wgan-ct'
DIM = 64
CRITIC_ITERS = 5
N_GPUS = 2
BATCH_SIZE = 64
ITERS = 200000
LAMBDA = 10
OUTPUT_DIM = 64 * 64 * 3


class Model(StepMixin):
    gen_name, disc_name = 'Generator', 'Discriminator.'    # :561, :563
    adam_args = (0.0, 0.9)                                 # :560-563

    def __init__(self, dtype=torch.float64, batch_size=BATCH_SIZE, dim=DIM, n_gpus=N_GPUS):
        self.lib = TFLib(dtype)
        self.dtype = dtype
        self.B = batch_size
        self.DIM = dim
        self.N_DEVICES = n_gpus
        self._init_opt()

    def lr(self, iteration):
        return 1e-4

    # ------------------------------------------------------------ builders
    def Normalize(self, name, axes, inputs):               # :87-93
        if ('Discriminator' in name) and (MODE == 'wgan-ct'):
            if axes != [0, 2, 3]:
                raise Exception('Layernorm over non-standard axes is unsupported')
            return self.lib.Layernorm(name, [1, 2, 3], inputs)
        return self.lib.Batchnorm(name, axes, inputs, fused=True)

    def ConvMeanPool(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = self.lib.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
        return tf_ops.mean_pool_2x2(output)

    def MeanPoolConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.mean_pool_2x2(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)

    def UpsampleConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.upsample_2x(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)

    def ResidualBlock(self, name, input_dim, output_dim, filter_size, inputs, resample=None, he_init=True):
        Conv2D = self.lib.Conv2D                            # :166-200
        if resample == 'down':
            conv_shortcut = self.MeanPoolConv
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
            conv_2 = functools.partial(self.ConvMeanPool, input_dim=input_dim, output_dim=output_dim)
        elif resample == 'up':
            conv_shortcut = self.UpsampleConv
            conv_1 = functools.partial(self.UpsampleConv, input_dim=input_dim, output_dim=output_dim)
            conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
        elif resample is None:
            conv_shortcut = Conv2D
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
            conv_2 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim)
        else:
            raise Exception('invalid resample value')
        if output_dim == input_dim and resample is None:
            shortcut = inputs
        else:
            shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim, filter_size=1,
                                     he_init=False, biases=True, inputs=inputs)
        output = inputs
        output = self.Normalize(name + '.BN1', [0, 2, 3], output)
        output = self._relu(output)
        output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output, he_init=he_init, biases=False)
        output = self.Normalize(name + '.BN2', [0, 2, 3], output)
        output = self._relu(output)
        output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output, he_init=he_init)
        return shortcut + output

    def Generator(self, n_samples, noise):                 # GoodGenerator :204-221
        dim = self.DIM
        output = self.lib.Linear('Generator.Input', 128, 4 * 4 * 8 * dim, noise)
        output = output.reshape(-1, 8 * dim, 4, 4)
        output = self.ResidualBlock('Generator.Res1', 8 * dim, 8 * dim, 3, output, resample='up')
        output = self.ResidualBlock('Generator.Res2', 8 * dim, 4 * dim, 3, output, resample='up')
        output = self.ResidualBlock('Generator.Res3', 4 * dim, 2 * dim, 3, output, resample='up')
        output = self.ResidualBlock('Generator.Res4', 2 * dim, 1 * dim, 3, output, resample='up')
        output = self.Normalize('Generator.OutputN', [0, 2, 3], output)
        output = self._relu(output)
        output = self.lib.Conv2D('Generator.Output', 1 * dim, 3, 3, output)
        output = torch.tanh(output)
        return output.reshape(-1, OUTPUT_DIM)

    def Discriminator(self, inputs, kp1, kp2, kp3, rnd=None, tag=None):     # GoodDiscriminator :357-373
        dim = self.DIM
        output = inputs.reshape(-1, 3, 64, 64)
        output = self.lib.Conv2D('Discriminator.Input', 3, dim, 3, output, he_init=False)
        output = self.ResidualBlock('Discriminator.Res1', dim, 2 * dim, 3, output, resample='down')
        output = self.ResidualBlock('Discriminator.Res2', 2 * dim, 4 * dim, 3, output, resample='down')
        output = tf_ops.dropout(output, kp1, None if kp1 == 1.0 else rnd.uniform(tag + '.1', output.shape))
        output = self.ResidualBlock('Discriminator.Res3', 4 * dim, 8 * dim, 3, output, resample='down')
        output = tf_ops.dropout(output, kp2, None if kp2 == 1.0 else rnd.uniform(tag + '.2', output.shape))
        output = self.ResidualBlock('Discriminator.Res4', 8 * dim, 8 * dim, 3, output, resample='down')
        output = tf_ops.dropout(output, kp3, None if kp3 == 1.0 else rnd.uniform(tag + '.3', output.shape))
        output2 = output.reshape(-1, 4 * 4 * 8 * dim)
        output = self.lib.Linear('Discriminator.Output', 4 * 4 * 8 * dim, 1, output2)
        return output.reshape(-1), output2

    def build(self):
        with torch.no_grad():
            fake = self.Generator(2, torch.randn(2, 128, dtype=self.dtype))
            self.Discriminator(fake, 1.0, 1.0, 1.0)
        return self

    # ------------------------------------------------------------ graphs
    def prep_real(self, real_data_conv):                   # :483
        return (2 * ((real_data_conv.to(torch.float32) / 255.) - .5)).reshape(real_data_conv.shape[0], OUTPUT_DIM)

    def disc_cost(self, rnd, all_real_data_conv):          # :479-546, wgan-ct
        self._begin(rnd)
        B = all_real_data_conv.shape[0]
        h = B // self.N_DEVICES
        costs, parts = [], []
        for i in range(self.N_DEVICES):
            real_data = self.prep_real(all_real_data_conv[i * h:(i + 1) * h]).to(self.dtype)
            with torch.no_grad():
                fake_data = self.Generator(h, rnd.normal('z.%d' % i, (h, 128)).to(self.dtype))
            disc_real, disc_real_2 = self.Discriminator(real_data, 0.8, 0.5, 0.5, rnd, 'drop.%d.real1' % i)
            disc_real_, disc_real_2_ = self.Discriminator(real_data, 0.8, 0.5, 0.5, rnd, 'drop.%d.real2' % i)
            disc_fake, _ = self.Discriminator(fake_data, 0.8, 0.5, 0.5, rnd, 'drop.%d.fake' % i)
            wgan = disc_fake.mean() - disc_real.mean()
            alpha = rnd.uniform('alpha.%d' % i, (h, 1)).to(self.dtype)
            gp, slopes, gradients = gradient_penalty(
                lambda x: self.Discriminator(x, 0.8, 0.5, 0.5, rnd, 'drop.%d.gp' % i)[0], real_data, fake_data, alpha)
            ct = consistency_term(disc_real, disc_real_, disc_real_2, disc_real_2_, LAMBDA_2, Factor_M)
            costs.append(wgan + ct + LAMBDA * gp)
            parts.append(dict(wgan=wgan, ct=ct, gp=gp, gradients=gradients, fake_data=fake_data, real_data=real_data))
        cost = sum(costs) / self.N_DEVICES                 # :546
        wgan = sum(p['wgan'] for p in parts) / self.N_DEVICES
        return dict(cost=cost, wgan_term=wgan, wgan=wgan,
                    ct=sum(p['ct'] for p in parts) / self.N_DEVICES, gp=sum(p['gp'] for p in parts) / self.N_DEVICES,
                    gradients=torch.cat([p['gradients'] for p in parts], 0),
                    fake_data=torch.cat([p['fake_data'] for p in parts], 0),
                    real_data=torch.cat([p['real_data'] for p in parts], 0))

    def gen_cost(self, rnd):                               # :484,:495,:545
        self._begin(rnd)
        h = self.B // self.N_DEVICES
        costs, fakes = [], []
        for i in range(self.N_DEVICES):
            fake = self.Generator(h, rnd.normal('z.%d' % i, (h, 128)).to(self.dtype))
            disc_fake, _ = self.Discriminator(fake, 0.8, 0.5, 0.5, rnd, 'drop.%d.fake' % i)
            costs.append(-disc_fake.mean())
            fakes.append(fake)
        return dict(cost=sum(costs) / self.N_DEVICES, fake_data=torch.cat(fakes, 0))
