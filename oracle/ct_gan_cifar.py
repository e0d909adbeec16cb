"""Oracle restatement of TG/CT_gan_cifar.py (MODE='wgan-CT') — TEST INFRASTRUCTURE.

TG = /root/reference/CT-GANs/tensorflow_generative_model.  Hyper-parameters
:34-43, Generator :58-79, Discriminator :81-100, input scaling :102-103,
loss :123-151, Adam :153-154.
"""
import torch

from . import tf_ops
from .tflib_ref import TFLib
from .ct_gan_common import StepMixin, consistency_term, gradient_penalty

LAMBDA_2 = 2.0
Factor_M = 0.0
MODE = 'wgan-CT'
DIM = 128
LAMBDA = 10
CRITIC_ITERS = 5
BATCH_SIZE = 64
OUTPUT_DIM = 3072


class Model(StepMixin):
    gen_name, disc_name = 'Generator', 'Discriminator'     # :112-113
    adam_args = (0.5, 0.9)                                 # :153-154

    def __init__(self, dtype=torch.float64, batch_size=BATCH_SIZE, dim=DIM):
        self.lib = TFLib(dtype)
        self.dtype = dtype
        self.B = batch_size
        self.DIM = dim
        self._init_opt()

    def lr(self, iteration):
        return 1e-4

    def Generator(self, n_samples, noise):                 # :58-79
        lib, DIM = self.lib, self.DIM
        output = lib.Linear('Generator.Input', 128, 4 * 4 * 4 * DIM, noise)
        output = lib.Batchnorm('Generator.BN1', [0], output)
        output = self._relu(output)
        output = output.reshape(-1, 4 * DIM, 4, 4)
        output = lib.Deconv2D('Generator.2', 4 * DIM, 2 * DIM, 5, output)
        output = lib.Batchnorm('Generator.BN2', [0, 2, 3], output)
        output = self._relu(output)
        output = lib.Deconv2D('Generator.3', 2 * DIM, DIM, 5, output)
        output = lib.Batchnorm('Generator.BN3', [0, 2, 3], output)
        output = self._relu(output)
        output = lib.Deconv2D('Generator.5', DIM, 3, 5, output)
        output = torch.tanh(output)
        return output.reshape(-1, OUTPUT_DIM)

    def Discriminator(self, inputs, rnd, tag):             # :81-100
        lib, DIM = self.lib, self.DIM
        output = inputs.reshape(-1, 3, 32, 32)
        output = lib.Conv2D('Discriminator.1', 3, DIM, 5, output, stride=2)
        output = self._lrelu(output)
        output = tf_ops.dropout(output, 0.5, rnd.uniform(tag + '.1', output.shape))
        output = lib.Conv2D('Discriminator.2', DIM, 2 * DIM, 5, output, stride=2)
        output = self._lrelu(output)
        output = tf_ops.dropout(output, 0.5, rnd.uniform(tag + '.2', output.shape))
        output = lib.Conv2D('Discriminator.3', 2 * DIM, 4 * DIM, 5, output, stride=2)
        output = self._lrelu(output)
        output = tf_ops.dropout(output, 0.5, rnd.uniform(tag + '.3', output.shape))
        output2 = output.reshape(-1, 4 * 4 * 4 * DIM)
        output = lib.Linear('Discriminator.Output', 4 * 4 * 4 * DIM, 1, output2)
        return output.reshape(-1), output2

    def build(self):
        class _Z:
            def uniform(self, tag, shape, lo=0., hi=1.):
                return torch.full(tuple(shape), 0.75)
        with torch.no_grad():
            fake = self.Generator(2, torch.randn(2, 128, dtype=self.dtype))
            self.Discriminator(fake, _Z(), 'b')
        return self

    def prep_real(self, real_data_int):                    # :102-103
        return 2 * ((real_data_int.to(torch.float32) / 255.) - .5)

    def disc_cost(self, rnd, real_data_int):               # :123-151
        self._begin(rnd)
        B = real_data_int.shape[0]
        real_data = self.prep_real(real_data_int).to(self.dtype)
        with torch.no_grad():
            fake_data = self.Generator(B, rnd.normal('z', (B, 128)).to(self.dtype))
        disc_real, disc_real_2 = self.Discriminator(real_data, rnd, 'drop.real1')
        disc_real_, disc_real_2_ = self.Discriminator(real_data, rnd, 'drop.real2')
        disc_fake, _ = self.Discriminator(fake_data, rnd, 'drop.fake')
        wgan = disc_fake.mean() - disc_real.mean()
        ct = consistency_term(disc_real, disc_real_, disc_real_2, disc_real_2_, LAMBDA_2, Factor_M)
        alpha = rnd.uniform('alpha', (B, 1)).to(self.dtype)
        gp, slopes, gradients = gradient_penalty(
            lambda x: self.Discriminator(x, rnd, 'drop.gp')[0], real_data, fake_data, alpha)
        cost = wgan + ct + LAMBDA * gp
        return dict(cost=cost, wgan=wgan, ct=ct, gp=gp, slopes=slopes, gradients=gradients,
                    disc_real=disc_real, disc_fake=disc_fake, fake_data=fake_data)

    def gen_cost(self, rnd):                               # :125
        self._begin(rnd)
        B = self.B
        fake_data = self.Generator(B, rnd.normal('z', (B, 128)).to(self.dtype))
        disc_fake, _ = self.Discriminator(fake_data, rnd, 'drop.fake')
        return dict(cost=-disc_fake.mean(), fake_data=fake_data)
