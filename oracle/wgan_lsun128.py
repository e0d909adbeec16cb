"""Oracle restatement of LS/wgan_LSUN_Bedrooms128.py (128x128 ResNet CT-GAN) -- TEST INFRASTRUCTURE.

LS = /root/reference/CT-GANs/tensorflow_generative_model/LSUN_bedrooms.  Hyper-parameters :31-57, Normalize :70-74 (layer norm
over [1,2,3] in the critic, fused batch norm in the generator), ConvMeanPool / MeanPoolConv / ScaledUpsampleConv :76-93,
ResidualBlock :95-137 ('down' = a 3x3 conv followed by a STRIDE-2 3x3 conv, shortcut MeanPoolConv; 'up' = two
ScaledUpsampleConv with gain 0.5), ResnetGenerator :139-167, ResnetDiscriminator :169-205 (dropout after the last three
blocks, global mean pool, Linear), loss graph :211-283 (one critic device: real + both fake splits through TWO stochastic
critic calls; the other: gradient penalty and the consistency term), lr decay :285-288, Adam(1e-4*decay, 0, .9) :289,296,
generator cost :291-295.  The forked tflib names conv biases and normalisation offsets `<name>.b` (TFLib style 'lsun').
SURVEY.md 8(f) row N4.  `width` scales every DIM_* constant (tests run a narrow model; 1.0 = the reference's widths).
"""
import functools

import torch

from . import tf_ops
from .tflib_ref import TFLib
from .ct_gan_common import StepMixin, consistency_term, gradient_penalty

BATCH_SIZE = 64
N_GPUS = 2
DIM_G = dict(d64=64, d32=128, d16=256, d8=512, d4=512)          # DIM_G_64 .. DIM_G_4   :34-38
DIM_D = dict(d64=128, d32=256, d16=512, d8=1024, d4=1024)       # DIM_D_64 .. DIM_D_4   :40-44
ITERS = 200000
LAMBDA_2 = 2.0
Factor_M = 0.0
LR = 1e-4
DECAY = True
CRITIC_ITERS = 5
GEN_BS_MULTIPLE = 1
OUTPUT_DIM = 3 * 128 * 128


class Model(StepMixin):
    gen_name, disc_name = 'Generator', 'Discriminator.'    # :289, :296
    adam_args = (0.0, 0.9)                                 # MOMENTUM_G / MOMENTUM_D = 0., beta2 = .9

    def __init__(self, dtype=torch.float64, batch_size=BATCH_SIZE, width=1.0, n_gpus=N_GPUS):
        self.lib = TFLib(dtype, style='lsun')
        self.dtype = dtype
        self.B = batch_size
        self.N_DEVICES = n_gpus
        self.G = {k: max(1, int(v * width)) for k, v in DIM_G.items()}
        self.D = {k: max(1, int(v * width)) for k, v in DIM_D.items()}
        self._init_opt()

    def lr(self, iteration):                               # :285-288
        return LR * (max(0., 1. - float(iteration) / ITERS) if DECAY else 1.)

    # ------------------------------------------------------------ builders
    def Normalize(self, name, inputs):                     # :70-74
        if 'Discriminator' in name:
            return self.lib.Layernorm(name, [1, 2, 3], inputs)
        return self.lib.Batchnorm(name, [0, 2, 3], inputs, fused=True)

    def MeanPoolConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.mean_pool_2x2(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)

    def ScaledUpsampleConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.upsample_2x(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases, gain=0.5)

    def ResidualBlock(self, name, input_dim, output_dim, filter_size, inputs, resample=None):
        Conv2D = self.lib.Conv2D                            # :95-137
        if resample == 'down':
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
            conv_2 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim, stride=2)
            conv_shortcut = self.MeanPoolConv
        elif resample == 'up':
            conv_1 = functools.partial(self.ScaledUpsampleConv, input_dim=input_dim, output_dim=output_dim)
            conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
            conv_shortcut = self.ScaledUpsampleConv
        elif resample is None:
            conv_shortcut = Conv2D
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim)
            conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
        else:
            raise Exception('invalid resample value')
        if output_dim == input_dim and resample is None:
            shortcut = inputs
        else:
            shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim, filter_size=1,
                                     he_init=False, biases=True, inputs=inputs)
        output = inputs
        output = self.Normalize(name + '.N1', output)
        output = self._relu(output)
        output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output)
        output = self.Normalize(name + '.N2', output)
        output = self._relu(output)
        output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output)
        return shortcut + output

    def Generator(self, n_samples, noise):                 # ResnetGenerator :139-167
        G = self.G
        output = self.lib.Linear('Generator.Input', 128, 4 * 4 * G['d4'], noise)
        output = output.reshape(-1, G['d4'], 4, 4)
        output = self.ResidualBlock('Generator.4_3', G['d4'], G['d8'], 3, output, resample='up')
        output = self.ResidualBlock('Generator.8_3', G['d8'], G['d16'], 3, output, resample='up')
        output = self.ResidualBlock('Generator.16_3', G['d16'], G['d32'], 3, output, resample='up')
        output = self.ResidualBlock('Generator.32_3', G['d32'], G['d64'], 3, output, resample='up')
        output = self.Normalize('Generator.OutputN', output)
        output = self._relu(output)
        output = self.ScaledUpsampleConv('Generator.Output', G['d64'], 3, 5, output, he_init=False)
        output = torch.tanh(output)
        return output.reshape(-1, OUTPUT_DIM)

    def Discriminator(self, inputs, kp1, kp2, kp3, rnd=None, tag=None):     # ResnetDiscriminator :169-205
        D = self.D
        output = inputs.reshape(-1, 3, 128, 128)
        output = self.lib.Conv2D('Discriminator.Input', 3, D['d64'], 5, output, he_init=True, stride=2)
        output = self.ResidualBlock('Discriminator.64_3', D['d64'], D['d32'], 3, output, resample='down')
        output = self.ResidualBlock('Discriminator.32_3', D['d32'], D['d16'], 3, output, resample='down')
        output = self.ResidualBlock('Discriminator.16_3', D['d16'], D['d8'], 3, output, resample='down')
        output = tf_ops.dropout(output, kp1, None if kp1 == 1.0 else rnd.uniform(tag + '.1', output.shape))
        output = self.ResidualBlock('Discriminator.8_1', D['d8'], D['d8'], 3, output, resample=None)
        output = tf_ops.dropout(output, kp2, None if kp2 == 1.0 else rnd.uniform(tag + '.2', output.shape))
        output = self.ResidualBlock('Discriminator.8_2', D['d8'], D['d8'], 3, output, resample=None)
        output = tf_ops.dropout(output, kp3, None if kp3 == 1.0 else rnd.uniform(tag + '.3', output.shape))
        output2 = output.mean(dim=[2, 3])
        output = self.lib.Linear('Discriminator.Output', D['d8'], 1, output2)
        return output.reshape(-1), output2

    def build(self):
        with torch.no_grad():
            fake = self.Generator(2, torch.randn(2, 128, dtype=self.dtype))
            self.Discriminator(fake, 1.0, 1.0, 1.0)
        return self

    # ------------------------------------------------------------ graphs
    def prep_real(self, all_real_data_conv):               # :221
        return (2 * ((all_real_data_conv.to(torch.float32) / 255.) - .5)).reshape(all_real_data_conv.shape[0], OUTPUT_DIM)

    def disc_cost(self, rnd, all_real_data_conv):          # :213-283
        self._begin(rnd)
        B = all_real_data_conv.shape[0]
        h = B // self.N_DEVICES
        with torch.no_grad():
            fake_data_splits = [self.Generator(h, rnd.normal('z.%d' % i, (h, 128)).to(self.dtype)) for i in range(self.N_DEVICES)]
        all_real_data = self.prep_real(all_real_data_conv).to(self.dtype)
        # DEVICES_A (one entry): real + both fake splits through two stochastic critic calls  :228-258
        real_and_fake_data = torch.cat([all_real_data] + fake_data_splits, dim=0)
        disc_all, disc_all_2 = self.Discriminator(real_and_fake_data, 0.8, 0.5, 0.5, rnd, 'drop.p1')
        disc_all_, disc_all_2_ = self.Discriminator(real_and_fake_data, 0.8, 0.5, 0.5, rnd, 'drop.p2')
        disc_real, disc_fake = disc_all[:B], disc_all[B:]
        disc_real_2, disc_real_, disc_real_2_ = disc_all_2[:B], disc_all_[:B], disc_all_2_[:B]
        wgan = disc_fake.mean() - disc_real.mean()
        # DEVICES_B (one entry): gradient penalty on B interpolates + the consistency term  :260-281
        fake_data = torch.cat(fake_data_splits, dim=0)
        alpha = rnd.uniform('alpha', (B, 1)).to(self.dtype)
        gp, slopes, gradients = gradient_penalty(
            lambda x: self.Discriminator(x, 0.8, 0.5, 0.5, rnd, 'drop.gp')[0], all_real_data, fake_data, alpha)
        ct = consistency_term(disc_real, disc_real_, disc_real_2, disc_real_2_, LAMBDA_2, Factor_M)
        cost = wgan + ct + 10. * gp                        # add_n(disc_costs) / len(DEVICES_A)  :283
        return dict(cost=cost, wgan_term=wgan, wgan=wgan, ct=ct, gp=gp, slopes=slopes, gradients=gradients,
                    fake_data=fake_data, real_data=all_real_data)

    def gen_cost(self, rnd):                               # :291-295
        self._begin(rnd)
        h = GEN_BS_MULTIPLE * self.B // self.N_DEVICES
        costs, fakes = [], []
        for i in range(self.N_DEVICES):
            fake = self.Generator(h, rnd.normal('z.%d' % i, (h, 128)).to(self.dtype))
            disc_fake, _ = self.Discriminator(fake, 0.8, 0.5, 0.5, rnd, 'drop.%d' % i)
            costs.append(-disc_fake.mean())
            fakes.append(fake)
        return dict(cost=sum(costs) / self.N_DEVICES, fake_data=torch.cat(fakes, 0))
