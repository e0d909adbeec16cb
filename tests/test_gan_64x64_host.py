"""CT_gan_64x64.py on the stand-in backend: one critic step and one generator step of
ctgan_b200/gan_64x64.py against the oracle restatement (itself pinned to the reference's code: tests/test_oracle_vs_reference.py)
with the same weights and replayed random draws -- loss terms, the GP gradient, every parameter gradient (the critic's
include the second-order layer-norm terms) and the Adam update."""
import numpy as np
import pytest
import torch

from tests import parity


@pytest.mark.parametrize('conditioned', [True, False])
def test_64x64_step_parity_fake_kernels(fake_kernels, conditioned):
    B, dim = 4, 8
    tr, om = parity.build_pair('64x64', 'cpu', torch.float32, B, dim=dim)
    parity.perturb_params(tr, om)
    rep = parity.critic_parity('64x64', tr, om, parity.make_inputs('64x64', B, 11), conditioned=conditioned)
    tol = 5e-4 if conditioned else 1e-2
    assert parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})[0] < tol, 'critic: ' + parity.format_report(rep)
    assert parity.worst(rep, 'adam.')[0] < 2e-3, 'critic: ' + parity.format_report(rep)
    rep = parity.gen_parity('64x64', tr, om, conditioned=conditioned)
    assert parity.worst(rep, 'loss.')[0] < 1e-3 and parity.worst(rep, 'adam.')[0] < 2e-3, 'gen: ' + parity.format_report(rep)
    assert parity.worst(rep, 'grad.')[0] < tol, 'gen: ' + parity.format_report(rep)


def test_64x64_parameter_surface(fake_kernels):
    """Names, creation order and shapes of the product's parameters == the oracle's (== the reference's, pinned in
    tests/test_oracle_vs_reference.py); DIM 64 widths."""
    import ctgan_b200.gan_64x64 as G
    import ctgan_b200.tflib as lib
    from oracle import ct_gan_64x64 as O
    np.random.seed(3)
    G.Trainer(device='cpu', seed=1, act_dtype=torch.float32, batch_size=2, dim=64)
    np.random.seed(3)
    om = O.Model(dtype=torch.float32, batch_size=2).build()
    assert list(lib._params) == list(om.lib._params)
    for n, p in lib._params.items():
        assert tuple(p.shape) == tuple(om.lib._params[n].shape), n
        assert torch.equal(p.detach().cpu(), om.lib._params[n].detach()), n           # same numpy draws, same formulas
    G.DIM = 64
