"""Oracle restatement of TG/CT_gan_cifar_resnet.py — TEST INFRASTRUCTURE.

TG = /root/reference/CT-GANs/tensorflow_generative_model.  Hyper-parameters
:33-56, Normalize :70-87, ConvMeanPool/MeanPoolConv/UpsampleConv :89-107,
ResidualBlock :109-141, OptimizedResBlockDisc1 :143-153, Generator :155-167,
Discriminator :169-186, critic graph :190-300, LR decay :309-312, generator
graph :314-330, Adam :333-338.  `len(DEVICES) == 2` always (N_GPUS=1 aliases
both entries to gpu:0, :61-63), so DEVICES_A and DEVICES_B have one entry each.
"""
import functools

import torch

from . import tf_ops
from .tflib_ref import TFLib
from .ct_gan_common import StepMixin, consistency_term, gradient_penalty

LAMBDA_2 = 2.0
Factor_M = 0.0
BATCH_SIZE = 64
GEN_BS_MULTIPLE = 2
ITERS = 100000
DIM_G = 128
DIM_D = 128
NORMALIZATION_G = True
NORMALIZATION_D = False
OUTPUT_DIM = 3072
LR = 2e-4
DECAY = True
N_CRITIC = 5
CONDITIONAL = True
ACGAN = True
ACGAN_SCALE = 1.
ACGAN_SCALE_G = 0.1
N_DEVICES = 2


class Model(StepMixin):
    gen_name, disc_name = 'Generator', 'Discriminator.'    # :336, :302
    adam_args = (0.0, 0.9)                                 # :333-334

    def __init__(self, dtype=torch.float64, batch_size=BATCH_SIZE, dim_g=DIM_G, dim_d=DIM_D,
                 conditional=CONDITIONAL, acgan=ACGAN, iters=ITERS):
        self.lib = TFLib(dtype)
        self.dtype = dtype
        self.B = batch_size
        self.DIM_G, self.DIM_D = dim_g, dim_d
        self.CONDITIONAL, self.ACGAN = conditional, acgan
        self.ITERS = iters
        self._init_opt()

    def lr(self, iteration):                               # :309-312 (float32 arithmetic in TF)
        decay = max(0., 1. - (float(iteration) / self.ITERS)) if DECAY else 1.
        return LR * decay

    # ------------------------------------------------------------ builders
    def Normalize(self, name, inputs, labels=None):        # :70-87
        if not self.CONDITIONAL:
            labels = None
        if self.CONDITIONAL and self.ACGAN and ('Discriminator' in name):
            labels = None
        if ('Discriminator' in name) and NORMALIZATION_D:
            raise Exception('Unsupported configuration')   # layernorm path: dead in the reference
        elif ('Generator' in name) and NORMALIZATION_G:
            if labels is not None:
                return self.lib.CondBatchnorm(name, [0, 2, 3], inputs, labels=labels, n_labels=10)
            return self.lib.Batchnorm(name, [0, 2, 3], inputs, fused=True)
        return inputs

    def ConvMeanPool(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = self.lib.Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=he_init, biases=biases)
        return tf_ops.mean_pool_2x2(output)

    def MeanPoolConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.mean_pool_2x2(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)

    def UpsampleConv(self, name, input_dim, output_dim, filter_size, inputs, he_init=True, biases=True):
        output = tf_ops.upsample_2x(inputs)
        return self.lib.Conv2D(name, input_dim, output_dim, filter_size, output, he_init=he_init, biases=biases)

    def ResidualBlock(self, name, input_dim, output_dim, filter_size, inputs, resample=None, labels=None):
        Conv2D = self.lib.Conv2D                            # :109-141
        if resample == 'down':
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=input_dim)
            conv_2 = functools.partial(self.ConvMeanPool, input_dim=input_dim, output_dim=output_dim)
            conv_shortcut = self.ConvMeanPool
        elif resample == 'up':
            conv_1 = functools.partial(self.UpsampleConv, input_dim=input_dim, output_dim=output_dim)
            conv_shortcut = self.UpsampleConv
            conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
        elif resample is None:
            conv_shortcut = Conv2D
            conv_1 = functools.partial(Conv2D, input_dim=input_dim, output_dim=output_dim)
            conv_2 = functools.partial(Conv2D, input_dim=output_dim, output_dim=output_dim)
        else:
            raise Exception('invalid resample value')
        if output_dim == input_dim and resample is None:
            shortcut = inputs
        else:
            shortcut = conv_shortcut(name + '.Shortcut', input_dim=input_dim, output_dim=output_dim,
                                     filter_size=1, he_init=False, biases=True, inputs=inputs)
        output = inputs
        output = self.Normalize(name + '.N1', output, labels=labels)
        output = self._relu(output)
        output = conv_1(name + '.Conv1', filter_size=filter_size, inputs=output)
        output = self.Normalize(name + '.N2', output, labels=labels)
        output = self._relu(output)
        output = conv_2(name + '.Conv2', filter_size=filter_size, inputs=output)
        return shortcut + output

    def OptimizedResBlockDisc1(self, inputs):              # :143-153
        D = self.DIM_D
        shortcut = self.MeanPoolConv('Discriminator.1.Shortcut', input_dim=3, output_dim=D, filter_size=1,
                                     he_init=False, biases=True, inputs=inputs)
        output = self.lib.Conv2D('Discriminator.1.Conv1', 3, D, 3, inputs)
        output = self._relu(output)
        output = self.ConvMeanPool('Discriminator.1.Conv2', D, D, 3, output)
        return shortcut + output

    def Generator(self, n_samples, labels, noise):         # :155-167
        G = self.DIM_G
        output = self.lib.Linear('Generator.Input', 128, 4 * 4 * G, noise)
        output = output.reshape(-1, G, 4, 4)
        output = self.ResidualBlock('Generator.1', G, G, 3, output, resample='up', labels=labels)
        output = self.ResidualBlock('Generator.2', G, G, 3, output, resample='up', labels=labels)
        output = self.ResidualBlock('Generator.3', G, G, 3, output, resample='up', labels=labels)
        output = self.Normalize('Generator.OutputN', output)
        output = self._relu(output)
        output = self.lib.Conv2D('Generator.Output', G, 3, 3, output, he_init=False)
        output = torch.tanh(output)
        return output.reshape(-1, OUTPUT_DIM)

    def Discriminator(self, inputs, labels, kp1, kp2, kp3, rnd=None, tag=None):   # :169-186
        D = self.DIM_D
        output = inputs.reshape(-1, 3, 32, 32)
        output = self.OptimizedResBlockDisc1(output)
        output = self.ResidualBlock('Discriminator.2', D, D, 3, output, resample='down', labels=labels)
        output = tf_ops.dropout(output, kp1, None if kp1 == 1.0 else rnd.uniform(tag + '.1', output.shape))
        output = self.ResidualBlock('Discriminator.3', D, D, 3, output, resample=None, labels=labels)
        output = tf_ops.dropout(output, kp2, None if kp2 == 1.0 else rnd.uniform(tag + '.2', output.shape))
        output = self.ResidualBlock('Discriminator.4', D, D, 3, output, resample=None, labels=labels)
        output = tf_ops.dropout(output, kp3, None if kp3 == 1.0 else rnd.uniform(tag + '.3', output.shape))
        output = self._relu(output)
        output2 = output.mean(dim=[2, 3])
        output_wgan = self.lib.Linear('Discriminator.Output', D, 1, output2).reshape(-1)
        if self.CONDITIONAL and self.ACGAN:
            output_acgan = self.lib.Linear('Discriminator.ACGANOutput', D, 10, output2)
            return output_wgan, output2, output_acgan
        return output_wgan, output2, None

    def build(self):
        with torch.no_grad():
            labels = torch.zeros(2, dtype=torch.int32)
            fake = self.Generator(2, labels, torch.randn(2, 128, dtype=self.dtype))
            self.Discriminator(fake, labels, 1.0, 1.0, 1.0)
        return self

    # ------------------------------------------------------------ graphs
    def prep_real(self, all_real_data_int, dequant):       # :201-202
        x = 2 * ((all_real_data_int.to(torch.float32) / 256.) - .5)
        return x + dequant

    def disc_cost(self, rnd, all_real_data_int, all_real_labels, with_clean=True):   # :190-300
        self._begin(rnd)
        B = all_real_data_int.shape[0]
        h = B // N_DEVICES
        labels_splits = [all_real_labels[:h], all_real_labels[h:]]
        with torch.no_grad():
            fake_data_splits = [
                self.Generator(h, labels_splits[i], rnd.normal('z.%d' % i, (h, 128)).to(self.dtype))
                for i in range(N_DEVICES)]
        all_real_data = self.prep_real(all_real_data_int,
                                       rnd.uniform('dequant', (B, OUTPUT_DIM), 0., 1. / 128)).to(self.dtype)
        # DEVICES_A (one entry): 2B samples through three critic calls  :213-266
        real_and_fake_data = torch.cat([all_real_data, fake_data_splits[0], fake_data_splits[1]], dim=0)
        real_and_fake_labels = torch.cat([all_real_labels, all_real_labels], dim=0)
        disc_all, disc_all_2, disc_all_acgan = self.Discriminator(
            real_and_fake_data, real_and_fake_labels, 0.8, 0.5, 0.5, rnd, 'drop.p1')
        disc_all_, disc_all_2_, _ = self.Discriminator(
            real_and_fake_data, real_and_fake_labels, 0.8, 0.5, 0.5, rnd, 'drop.p2')
        disc_real, disc_fake = disc_all[:B], disc_all[B:]
        disc_real_2 = disc_all_2[:B]
        disc_real_ = disc_all_[:B]
        disc_real_2_ = disc_all_2_[:B]
        wgan = disc_fake.mean() - disc_real.mean()
        out = {}
        if self.CONDITIONAL and self.ACGAN:
            acgan = tf_ops.sparse_softmax_cross_entropy_with_logits(
                disc_all_acgan[:B], real_and_fake_labels[:B]).mean()
            if with_clean:
                with torch.no_grad():
                    _, _, clean = self.Discriminator(real_and_fake_data, real_and_fake_labels, 1.0, 1.0, 1.0)
                pred = clean.argmax(dim=1).to(torch.int32)
                out['acgan_acc'] = (pred[:B] == real_and_fake_labels[:B]).to(torch.float32).mean()
                out['acgan_fake_acc'] = (pred[B:] == real_and_fake_labels[B:]).to(torch.float32).mean()
        else:
            acgan = torch.zeros((), dtype=self.dtype)
        # DEVICES_B (one entry): GP on B interpolates + CT on the real halves  :269-293
        fake_data = torch.cat(fake_data_splits, dim=0)
        alpha = rnd.uniform('alpha', (B, 1)).to(self.dtype)
        gp, slopes, gradients = gradient_penalty(
            lambda x: self.Discriminator(x, all_real_labels, 0.8, 0.5, 0.5, rnd, 'drop.gp')[0],
            all_real_data, fake_data, alpha)
        gradient_penalty_ = 10.0 * gp
        ct = consistency_term(disc_real, disc_real_, disc_real_2, disc_real_2_, LAMBDA_2, Factor_M)
        disc_wgan = wgan + ct + gradient_penalty_          # add_n(disc_costs)/len(DEVICES_A)  :295
        cost = disc_wgan + ACGAN_SCALE * acgan             # :300
        out.update(cost=cost, wgan_term=wgan, disc_wgan=disc_wgan, ct=ct, gp=gp, acgan=acgan,
                   slopes=slopes, gradients=gradients, fake_data=fake_data, real_data=all_real_data)
        return out

    def gen_cost(self, rnd):                               # :314-330
        self._begin(rnd)
        B = self.B
        n_samples = GEN_BS_MULTIPLE * B // N_DEVICES
        gen_costs, gen_acgan_costs = [], []
        fakes = []
        for i in range(N_DEVICES):
            fake_labels = rnd.labels('labels.%d' % i, n_samples)
            fake = self.Generator(n_samples, fake_labels, rnd.normal('z.%d' % i, (n_samples, 128)).to(self.dtype))
            fakes.append(fake)
            disc_fake, _, disc_fake_acgan = self.Discriminator(fake, fake_labels, 0.8, 0.5, 0.5, rnd, 'drop.%d' % i)
            gen_costs.append(-disc_fake.mean())
            if self.CONDITIONAL and self.ACGAN:
                gen_acgan_costs.append(tf_ops.sparse_softmax_cross_entropy_with_logits(
                    disc_fake_acgan, fake_labels).mean())
        cost = sum(gen_costs) / N_DEVICES
        out = dict(gen_wgan=cost)
        if self.CONDITIONAL and self.ACGAN:
            acg = sum(gen_acgan_costs) / N_DEVICES
            cost = cost + ACGAN_SCALE_G * acg
            out['gen_acgan'] = acg
        out['cost'] = cost
        out['fake_data'] = torch.cat(fakes, 0)
        return out
