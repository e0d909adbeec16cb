"""Shared step machinery for the three oracle scripts (TEST INFRASTRUCTURE).

Restates the parts of the reference scripts that are identical across them: the
CT + GP loss (TG/CT_gan_mnist.py:146-167, TG/CT_gan_cifar.py:123-151,
TG/CT_gan_cifar_resnet.py:277-293) and the "compute gradients of the cost wrt the
name-selected variables, then tf.train.AdamOptimizer.apply" step
(TG/CT_gan_cifar.py:153-154, TG/CT_gan_cifar_resnet.py:333-338).
"""
import torch

from .tf_ops import TFAdam


def consistency_term(d1, d2, f1, f2, lambda_2, factor_m):
    """CT = l2*(D'-D'')^2 + 0.1*l2*mean_f((D_'-D_'')^2); CT_ = mean(max(CT-M', 0))
    -- TG/CT_gan_cifar.py:131-134."""
    ct = lambda_2 * (d1 - d2) ** 2
    ct = ct + lambda_2 * 0.1 * ((f1 - f2) ** 2).mean(dim=1)
    ct_ = torch.maximum(ct - factor_m, 0.0 * (ct - factor_m))
    return ct_.mean()


def gradient_penalty(disc_fn, real, fake, alpha):
    """interpolates = real + alpha*(fake-real); slopes = ||dD/dx^||_2 per sample (no
    epsilon under the sqrt); returns mean((slopes-1)^2), slopes and the raw gradient
    -- TG/CT_gan_cifar.py:137-150.  Second-order graph kept (create_graph)."""
    differences = fake - real
    interpolates = (real + alpha * differences).detach().requires_grad_(True)
    d = disc_fn(interpolates)
    gradients = torch.autograd.grad(d.sum(), interpolates, create_graph=True)[0]
    slopes = torch.sqrt((gradients ** 2).sum(dim=1))
    return ((slopes - 1.) ** 2).mean(), slopes, gradients


class StepMixin:
    """Needs: self.lib (TFLib), self.disc_cost(real, rnd[, labels]), self.gen_cost(rnd),
    self.gen_name / self.disc_name (substring selectors), self.adam_args, self.lr(iteration)."""

    # Activation sites.  By default the oracle decides the ReLU / LeakyReLU pattern from its own
    # pre-activation (tf.nn.relu, tf.maximum(alpha*x, x)).  For the "pattern-conditioned" parity
    # mode the test hands over the 0/1 patterns the device computed (`patterns`: an iterator of
    # bool tensors in call order): both sides then evaluate the SAME linear region of the
    # piecewise-linear network, which separates arithmetic error from the activation-pattern
    # flips that reduced-precision pre-activations cause at near-zero values.
    _patterns = None

    def _pattern(self, x):
        if self._patterns is None:
            return x > 0
        p = next(self._patterns)
        if p.shape[0] < x.shape[0]:            # device ran this pass on the leading samples only
            pad = torch.ones((x.shape[0] - p.shape[0],) + tuple(p.shape[1:]), dtype=torch.bool)
            p = torch.cat([p, pad], dim=0)
        assert p.shape == x.shape, (tuple(p.shape), tuple(x.shape))
        return p

    def _relu(self, x):
        return x * self._pattern(x).to(x.dtype)

    def _lrelu(self, x, alpha=0.2):
        p = self._pattern(x)
        return x * torch.where(p, torch.ones((), dtype=x.dtype), torch.full((), alpha, dtype=x.dtype))

    def _begin(self, rnd):
        pats = getattr(rnd, 'patterns', None)
        self._patterns = iter(pats) if pats is not None else None

    def _init_opt(self):
        self.gen_opt = TFAdam(*self.adam_args)
        self.disc_opt = TFAdam(*self.adam_args)

    def _grads(self, cost, named):
        names = list(named.keys())
        gs = torch.autograd.grad(cost, [named[n] for n in names], allow_unused=True)
        return {n: g for n, g in zip(names, gs)}

    def critic_step(self, rnd, *inputs, iteration=0, apply=True):
        out = self.disc_cost(rnd, *inputs)
        named = self.lib.named_params_with_name(self.disc_name)
        grads = self._grads(out['cost'], named)
        if apply:
            self.disc_opt.apply(named, grads, self.lr(iteration))
        out = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        out['grads'] = grads
        return out

    def gen_step(self, rnd, iteration=0, apply=True):
        out = self.gen_cost(rnd)
        named = self.lib.named_params_with_name(self.gen_name)
        grads = self._grads(out['cost'], named)
        if apply:
            self.gen_opt.apply(named, grads, self.lr(iteration))
        out = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        out['grads'] = grads
        return out
