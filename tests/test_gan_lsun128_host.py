"""LS/wgan_LSUN_Bedrooms128.py (SURVEY.md 8(f) N4, 128x128 ResNet CT-GAN) on the stand-in backend: one critic step and one
generator step of ctgan_b200/gan_lsun128.py against the oracle restatement (itself pinned to the reference's own code and the
fork of tflib that ships with it: tests/test_oracle_vs_reference.py) with the same weights and replayed random draws -- loss
terms, the GP gradient, every parameter gradient (the critic's include the second-order layer-norm terms and the stride-2
3x3 convs of the 'down' blocks) and the Adam update with the script's decayed learning rate."""
import numpy as np
import pytest
import torch

from tests import parity


@pytest.mark.parametrize('conditioned', [True, False])
def test_lsun128_step_parity_fake_kernels(fake_kernels, conditioned):
    B, width = 4, 1.0 / 32
    tr, om = parity.build_pair('lsun128', 'cpu', torch.float32, B, width=width)
    parity.perturb_params(tr, om)
    rep = parity.critic_parity('lsun128', tr, om, parity.make_inputs('lsun128', B, 11), iteration=50000, conditioned=conditioned)
    tol = 5e-4 if conditioned else 1e-2
    assert parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})[0] < tol, 'critic: ' + parity.format_report(rep)
    assert parity.worst(rep, 'adam.')[0] < 2e-3, 'critic: ' + parity.format_report(rep)
    rep = parity.gen_parity('lsun128', tr, om, iteration=50000, conditioned=conditioned)
    assert parity.worst(rep, 'loss.')[0] < 1e-3 and parity.worst(rep, 'adam.')[0] < 2e-3, 'gen: ' + parity.format_report(rep)
    assert parity.worst(rep, 'grad.')[0] < tol, 'gen: ' + parity.format_report(rep)


def test_lsun128_parameter_surface(fake_kernels):
    """Names (the fork's `.b` suffixes), creation order, shapes and initial values of the product's parameters == the oracle's
    (== the reference's, pinned in tests/test_oracle_vs_reference.py), at 1/8 of the reference's widths; building another
    script afterwards gets the CT scripts' names back."""
    import ctgan_b200.gan_lsun128 as G
    import ctgan_b200.gan_mnist as M
    import ctgan_b200.tflib as lib
    from oracle import wgan_lsun128 as O
    np.random.seed(3)
    G.Trainer(device='cpu', seed=1, act_dtype=torch.float32, batch_size=2, width=0.125)
    np.random.seed(3)
    om = O.Model(dtype=torch.float32, batch_size=2, width=0.125).build()
    assert list(lib._params) == list(om.lib._params)
    assert 'Discriminator.Input.b' in lib._params and 'Discriminator.64_3.N1.b' in lib._params and 'Generator.4_3.N1.b' in lib._params
    for n, p in lib._params.items():
        assert tuple(p.shape) == tuple(om.lib._params[n].shape), n
        assert torch.equal(p.detach().cpu(), om.lib._params[n].detach()), n           # same numpy draws, same formulas
    G.WIDTH = 1.0
    M.Trainer(device='cpu', seed=1, act_dtype=torch.float32, batch_size=2)
    assert 'Discriminator.1.Biases' in lib._params and not any(n.endswith('.Filters.b') for n in lib._params)
