"""Pins the oracle (and the product's parameter surface) to the REFERENCE'S OWN CODE.

 * test_oracle_matches_golden: tests/golden/*.npz were produced by executing the reference's tflib ops
   and the model/loss sections of its CT_gan_*.py scripts (oracle/ref_harness.py; script:
   tests/golden/make_golden.py).  The oracle restatement, given the same parameters, inputs and
   random draws, must reproduce disc_cost, gen_cost, the GP gradient and every parameter gradient.
   Runs anywhere (no /root/reference needed).
 * test_oracle_matches_live_reference: the same comparison against a live run of the reference code at
   the scripts' real widths (only where /root/reference exists, i.e. the build container).
 * test_product_parameter_surface_matches_reference: names, shapes, trainability and init ranges of the
   parameters the product creates vs the ones the reference code creates.
TensorFlow itself was never executed: its op semantics come from oracle/tf_ops.py (documented TF-1.x
behaviour), shared by the shim and the oracle, so this pins everything EXCEPT TF's internal kernels.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ct_gan_mnist, ct_gan_cifar, ct_gan_cifar_resnet, ct_gan_64x64, wgan_lsun128, ref_harness
from oracle.rand import ReplayRandom

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODS = {'mnist': ct_gan_mnist, 'cifar': ct_gan_cifar, 'resnet': ct_gan_cifar_resnet, '64x64': ct_gan_64x64, 'lsun128': wgan_lsun128}


def _model(script, B, dim):
    if script == 'lsun128':                      # dim = the divisor of the reference's widths
        m = MODS[script].Model(dtype=torch.float64, batch_size=B, width=1.0 / dim)
    elif script == 'resnet':
        m = MODS[script].Model(dtype=torch.float64, batch_size=B, dim_g=dim, dim_d=dim)
    else:
        m = MODS[script].Model(dtype=torch.float64, batch_size=B, dim=dim)
    np.random.seed(0)
    return m.build()


def _rel(a, b):
    a, b = torch.as_tensor(a).double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _compare(script, m, params, inputs, tape_disc, tape_gen, ref, tol):
    assert set(m.lib._params) == set(params), sorted(set(m.lib._params) ^ set(params))
    for n, p in params.items():
        assert tuple(m.lib._params[n].shape) == tuple(p.shape), n
        m.lib._params[n].data.copy_(torch.as_tensor(p).double())
    inputs = tuple(torch.as_tensor(a) for a in inputs)
    kw = dict(with_clean=False) if script == 'resnet' else {}
    out = m.disc_cost(ReplayRandom(tape_disc), *inputs, **kw)
    named = m.lib.named_params_with_name(m.disc_name)
    grads = m._grads(out['cost'], named)
    assert abs(float(out['cost']) - float(ref['disc_cost'])) <= tol * max(1.0, abs(float(ref['disc_cost'])))
    gp_rows = len(ref['gp_gradients'])          # 64x64: the reference's `gradients` is the LAST tower's (TG/CT_gan_64x64.py:505)
    assert _rel(out['gradients'][-gp_rows:], ref['gp_gradients']) < tol
    floor = 1e-6 * max(float(torch.as_tensor(g).double().norm()) for g in ref['disc_grads'].values())
    for n, g in ref['disc_grads'].items():
        err = float((grads[n].double() - torch.as_tensor(g).double()).norm()) / max(float(torch.as_tensor(g).double().norm()), floor)
        assert err < tol, ('disc grad', n, err)
    assert set(ref['disc_grads']) == set(n for n, g in grads.items() if g is not None)
    out = m.gen_cost(ReplayRandom(tape_gen))
    named = m.lib.named_params_with_name(m.gen_name)
    grads = m._grads(out['cost'], named)
    assert abs(float(out['cost']) - float(ref['gen_cost'])) <= tol * max(1.0, abs(float(ref['gen_cost'])))
    floor = 1e-6 * max(float(torch.as_tensor(g).double().norm()) for g in ref['gen_grads'].values())
    for n, g in ref['gen_grads'].items():
        err = float((grads[n].double() - torch.as_tensor(g).double()).norm()) / max(float(torch.as_tensor(g).double().norm()), floor)
        assert err < tol, ('gen grad', n, err)


def load_golden(script):
    z = np.load(os.path.join(GOLDEN, script + '.npz'))
    pick = lambda pre: {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    ref = dict(disc_cost=z['disc_cost'], gen_cost=z['gen_cost'], gp_gradients=z['gp_gradients'],
               disc_grads=pick('grad_disc.'), gen_grads=pick('grad_gen.'))
    inputs = [z['input.%d' % i] for i in range(len(pick('input.')))]
    return dict(B=int(z['meta.B']), dim=int(z['meta.dim']), params=pick('param.'), trainable=pick('trainable.'),
                inputs=inputs, tape_disc=pick('tape_disc.'), tape_gen=pick('tape_gen.'), ref=ref)


@pytest.mark.parametrize('script', ['mnist', 'cifar', 'resnet', '64x64', 'lsun128'])
def test_oracle_matches_golden(script):
    g = load_golden(script)
    m = _model(script, g['B'], g['dim'])
    _compare(script, m, g['params'], g['inputs'], g['tape_disc'], g['tape_gen'], g['ref'], tol=2e-6)   # fixtures hold fp32 copies


@pytest.mark.skipif(not ref_harness.available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('script,B', [('mnist', 5), ('cifar', 3), ('resnet', 2), ('64x64', 2), ('lsun128', 2)])
def test_oracle_matches_live_reference(script, B):
    from tests.golden.make_golden import inputs_for
    inputs = inputs_for(script, B, 77)
    # the scripts' real widths (the 128x128 LSUN model at 1/4 of them: fp64 on CPU)
    r = ref_harness.run_reference(script, B, 77, inputs, **(dict(width=0.25) if script == 'lsun128' else {}))
    dim = {'mnist': 64, 'cifar': 128, 'resnet': 128, '64x64': 64, 'lsun128': 4}[script]
    m = _model(script, B, dim)
    ref = dict(disc_cost=r['disc_cost'], gen_cost=r['gen_cost'], gp_gradients=r['gp_gradients'],
               disc_grads=r['disc_grads'], gen_grads={k: v for k, v in r['gen_grads'].items() if v is not None})
    _compare(script, m, r['params'], inputs, r['tape_disc'], r['tape_gen'], ref, tol=1e-9)
    # parameter counts the survey derived from the reference
    count = lambda sel: sum(int(np.prod(p.shape)) for n, p in r['params'].items() if sel in n and r['trainable'][n])
    expect = {'mnist': (1030145, 1554177), 'cifar': (4114689, 5179907), 'resnet': (1055115, 1218307), '64x64': None, 'lsun128': None}[script]
    if expect is not None:
        assert (count('Discriminator'), count('Generator')) == expect
    else:
        print('%s parameter counts: D %d, G %d' % (script, count('Discriminator'), count('Generator')))


@pytest.mark.skipif(not ref_harness.available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('script', ['mnist', 'cifar', 'resnet'])
def test_product_parameter_surface_matches_reference(fake_kernels, script):
    import importlib
    from tests import parity
    from tests.golden.make_golden import inputs_for
    r = ref_harness.run_reference(script, 2, 5, inputs_for(script, 2, 5))
    prod = importlib.import_module(parity.SCRIPTS[script][0])
    np.random.seed(5)
    prod.Trainer(device='cpu', seed=1, act_dtype=torch.float32, batch_size=2)
    import ctgan_b200.tflib as lib
    assert list(lib._params) == list(r['params'])                       # same names, same creation order
    for n, p in lib._params.items():
        q = r['params'][n]
        assert tuple(p.shape) == tuple(q.shape), n
        assert bool(p.requires_grad) == r['trainable'][n], n
        if n.endswith(('.Filters', '.W')):                              # uniform(+-sqrt(3)*stdev): same range
            assert abs(float(p.abs().max()) / float(q.abs().max()) - 1.0) < 0.02, n
            assert abs(float(p.std()) / float(q.std()) - 1.0) < 0.05, n
        else:
            assert torch.equal(p.detach().double().cpu(), q.double()), n  # zeros / ones
