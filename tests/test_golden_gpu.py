"""GPU (`-m gpu`): the CUDA path against golden vectors produced by THE REFERENCE'S OWN CODE
(tests/golden/*.npz, see tests/golden/make_golden.py) -- no oracle in between.  The fixtures' parameters are
loaded into the product by reference name, the fixtures' random draws (z, dropout uniforms, alpha, dequant
noise, labels) are injected through DeviceRandom.replay, and disc_cost / gen_cost / the GP gradient / every
parameter gradient are compared.  fp32 path, tolerance 1e-3 (measured ~1e-6; the bound leaves room for a
ReLU tie flip, see tests/test_step_parity_gpu.py)."""
import importlib

import numpy as np
import pytest
import torch

from tests import parity
from tests.test_oracle_vs_reference import load_golden

pytestmark = pytest.mark.gpu
DIM_NAMES = {'mnist': ['DIM'], 'cifar': ['DIM'], 'resnet': ['DIM_G', 'DIM_D']}


@pytest.mark.parametrize('script', ['mnist', 'cifar', 'resnet'])
def test_product_matches_reference_golden(script):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    g = load_golden(script)
    prod = importlib.import_module(parity.SCRIPTS[script][0])
    saved = {k: getattr(prod, k) for k in DIM_NAMES[script]}
    for k in saved:
        setattr(prod, k, g['dim'])
    try:
        np.random.seed(0)
        tr = prod.Trainer(device='cuda', seed=1, act_dtype=torch.float32, batch_size=g['B'])
        import ctgan_b200.tflib as lib
        assert set(lib._params) == set(g['params'])
        with torch.no_grad():
            for n, p in lib._params.items():
                p.copy_(torch.from_numpy(g['params'][n]).to(p.device).reshape(p.shape))
        inputs = tuple(torch.from_numpy(a).cuda() for a in g['inputs'])
        ref = g['ref']
        # ---- critic
        tr.rng.replay = g['tape_disc']
        tr.disc_opt.zero_grad()
        res = tr.critic_forward_backward(*inputs)
        cost = float(res['out'][0])
        assert abs(cost - float(ref['disc_cost'])) <= 1e-3 * max(1.0, abs(float(ref['disc_cost']))), (cost, float(ref['disc_cost']))
        assert parity.rel_err(res['gradients'], torch.from_numpy(ref['gp_gradients'])) < 1e-3
        floor = 1e-4 * max(float(np.linalg.norm(v)) for v in ref['disc_grads'].values())
        for n, q in tr.disc_opt.params.items():
            assert parity.rel_err(q.grad, torch.from_numpy(ref['disc_grads'][n]), floor) < 1e-3, n
        # ---- generator
        tr.rng.replay = g['tape_gen']
        tr.gen_opt.zero_grad()
        res = tr.gen_forward_backward()
        assert abs(float(res['cost']) - float(ref['gen_cost'])) <= 1e-3 * max(1.0, abs(float(ref['gen_cost'])))
        floor = 1e-4 * max(float(np.linalg.norm(v)) for v in ref['gen_grads'].values())
        for n, q in tr.gen_opt.params.items():
            if n in ref['gen_grads']:
                assert parity.rel_err(q.grad, torch.from_numpy(ref['gen_grads'][n]), floor) < 2e-3, n
    finally:
        for k, v in saved.items():
            setattr(prod, k, v)
