"""Parity harness shared by the CPU (fake-kernel) and GPU tests: run one critic step and one
generator step of a ctgan_b200 Trainer, replay the recorded random draws through the CPU
oracle with the SAME weights, and compare loss terms, the GP gradient, every parameter
gradient and the post-Adam parameters."""
import importlib

import numpy as np
import torch

from oracle.rand import ReplayRandom

SCRIPTS = {
    'mnist': ('ctgan_b200.gan_mnist', 'oracle.ct_gan_mnist'),
    'cifar': ('ctgan_b200.gan_cifar', 'oracle.ct_gan_cifar'),
    'resnet': ('ctgan_b200.gan_cifar_resnet', 'oracle.ct_gan_cifar_resnet'),
    '64x64': ('ctgan_b200.gan_64x64', 'oracle.ct_gan_64x64'),
    'lsun128': ('ctgan_b200.gan_lsun128', 'oracle.wgan_lsun128'),      # N4, second half (LS/wgan_LSUN_Bedrooms128.py)
}


def rel_err(a, b, floor=1e-30):
    """norm-relative error ||a-b|| / max(||b||, floor) (b = oracle), in float64.  `floor` guards
    tensors whose true value is zero (e.g. the gradient of a bias that feeds a batch norm)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = b.norm().item()
    return (a - b).norm().item() / max(den, floor)


def _grad_floor(ref_grads, frac=1e-4):
    return frac * max(float(g.detach().double().norm()) for g in ref_grads.values() if g is not None)


def _flat_rel(params, ref_grads):
    """The parameter gradient as ONE vector: ||g - g_ref|| / ||g_ref|| over all parameters of the optimizer."""
    num = den = 0.0
    for n, q in params.items():
        a, b = q.grad.detach().double().cpu().reshape(-1), ref_grads[n].detach().double().cpu().reshape(-1)
        num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def make_inputs(script, B, seed):
    rs = np.random.RandomState(seed)
    if script == 'mnist':
        return (torch.from_numpy(rs.random_sample((B, 784)).astype('float32')),)
    if script == 'cifar':
        return (torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')),)
    if script == '64x64':
        return (torch.from_numpy(rs.randint(0, 256, (B, 3, 64, 64)).astype('int32')),)
    if script == 'lsun128':
        return (torch.from_numpy(rs.randint(0, 256, (B, 3, 128, 128)).astype('int32')),)
    return (torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')),
            torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')))


def build_pair(script, device, act_dtype, B, seed=7, oracle_dtype=torch.float64, **model_kw):
    prod_name, ora_name = SCRIPTS[script]
    prod = importlib.import_module(prod_name)
    ora = importlib.import_module(ora_name)
    np.random.seed(seed)
    tr = prod.Trainer(device=device, seed=seed + 100, act_dtype=act_dtype, batch_size=B, record=True, **model_kw)
    import ctgan_b200.tflib as lib
    np.random.seed(seed)
    om = ora.Model(dtype=oracle_dtype, batch_size=B, **model_kw).build()
    # identical weights: copy product -> oracle by reference name
    assert set(om.lib._params) == set(lib._params), (sorted(set(om.lib._params) ^ set(lib._params)))
    for n, p in lib._params.items():
        assert tuple(om.lib._params[n].shape) == tuple(p.shape), n
        om.lib._params[n].data.copy_(p.detach().cpu().to(oracle_dtype))
    return tr, om


def perturb_params(tr, om, seed=3, scale=0.05):
    """Biases/offsets start at zero in the reference; give every zero-initialised parameter a
    non-trivial value (identically on both sides) so the test exercises them."""
    import ctgan_b200.tflib as lib
    rs = np.random.RandomState(seed)
    for n, p in lib._params.items():
        if not p.requires_grad:
            continue
        if n.endswith(('.Biases', '.b', '.offset')):
            v = torch.from_numpy((scale * rs.standard_normal(tuple(p.shape))).astype('float32'))
        elif n.endswith('.scale'):
            v = torch.from_numpy((1.0 + scale * rs.standard_normal(tuple(p.shape))).astype('float32'))
        else:
            continue
        with torch.no_grad():
            p.copy_(v.to(p.device))
        om.lib._params[n].data.copy_(v.to(om.lib._params[n].dtype))


def _fix_tape(script, tape, B):
    tape = dict(tape)
    if script in ('resnet', 'lsun128'):
        if 'dequant' in tape:
            tape['dequant'] = tape['dequant'].cpu() * (1.0 / 128)     # kernel applies noise_hi * u
        for k in list(tape):
            if k.startswith('drop.p2.'):
                # the product runs pass '' on the real half only; the oracle draws a mask for 2B samples
                t = tape[k].cpu()
                tape[k] = torch.cat([t, torch.zeros_like(t)], dim=0)
    return tape


def critic_parity(script, tr, om, inputs, iteration=0, conditioned=False, floor_frac=1e-4):
    """conditioned=True: the oracle applies the activation patterns the device computed (same linear
    region), isolating arithmetic error from near-zero ReLU flips; False: fully independent oracle."""
    dev = tr.device
    inputs_dev = tuple(t.to(dev) for t in inputs)
    tr.disc_opt.zero_grad()
    tr.rng.begin_recording()
    res = tr.critic_forward_backward(*inputs_dev)
    tr.rng.stop_recording()
    tape = _fix_tape(script, tr.rng.tape, inputs[0].shape[0])
    kw = dict(with_clean=False) if script == 'resnet' else {}
    pats = tr.rng.patterns if conditioned else None
    if pats is not None and hasattr(tr, 'oracle_pattern_order'):
        pats = tr.oracle_pattern_order(pats, 'critic')
    ref = om.disc_cost(ReplayRandom(tape, pats), *inputs, **kw)
    named = om.lib.named_params_with_name(om.disc_name)
    ref_grads = om._grads(ref['cost'], named)
    out = res['out'].cpu()
    report = {}
    if script == 'resnet':
        pairs = dict(cost=(out[0], ref['cost']), wgan=(out[1], ref['wgan_term']), ct=(out[2], ref['ct']),
                     gp=(out[3], ref['gp']), acgan=(out[4], ref['acgan']))
    else:
        pairs = dict(cost=(out[0], ref['cost']), wgan=(out[1], ref['wgan']), ct=(out[2], ref['ct']), gp=(out[3], ref['gp']))
    scale = max(abs(float(ref['cost'].detach())), 1.0)
    for k, (a, b) in pairs.items():
        report['loss.' + k] = abs(float(a) - float(b)) / max(abs(float(b)), 0.05 * scale)
    report['fake_data'] = rel_err(res['fake_data'], ref['fake_data'])
    report['gp_gradient'] = rel_err(res['gradients'], ref['gradients'])
    floor = _grad_floor(ref_grads, floor_frac)
    for n, q in tr.disc_opt.params.items():
        report['grad.' + n] = rel_err(q.grad, ref_grads[n], floor)
    report['gradall'] = _flat_rel(tr.disc_opt.params, ref_grads)
    # optimizer: the SAME gradients (the product's) go through both Adam implementations, so this
    # isolates the update rule (Adam's m/sqrt(v) is sign-like at t=1 and would amplify gradient noise)
    before = {n: q.detach().clone() for n, q in tr.disc_opt.params.items()}
    same_grads = {n: q.grad.detach().cpu().to(named[n].dtype) for n, q in tr.disc_opt.params.items()}
    tr.disc_opt.step(tr.lr(iteration) if hasattr(tr, 'lr') else None, 1)
    ref_before = {n: p.detach().clone() for n, p in named.items()}
    om.disc_opt.apply(named, same_grads, om.lr(iteration))
    for n, q in tr.disc_opt.params.items():
        report['adam.' + n] = rel_err(q.detach() - before[n], named[n].detach() - ref_before[n])
    tr.rng.end_step()
    return report


def gen_parity(script, tr, om, iteration=1, conditioned=False, floor_frac=1e-4):
    tr.gen_opt.zero_grad()
    tr.rng.begin_recording()
    res = tr.gen_forward_backward()
    tr.rng.stop_recording()
    tape = dict(tr.rng.tape)
    pats = tr.rng.patterns if conditioned else None
    if pats is not None and hasattr(tr, 'oracle_pattern_order'):
        pats = tr.oracle_pattern_order(pats, 'gen')
    ref = om.gen_cost(ReplayRandom(tape, pats))
    named = om.lib.named_params_with_name(om.gen_name)
    ref_grads = om._grads(ref['cost'], named)
    report = {'loss.gen_cost': abs(float(res['cost']) - float(ref['cost'])) / max(abs(float(ref['cost'])), 0.05)}
    floor = _grad_floor(ref_grads, floor_frac)
    for n, q in tr.gen_opt.params.items():
        report['grad.' + n] = rel_err(q.grad, ref_grads[n], floor)
    report['gradall'] = _flat_rel(tr.gen_opt.params, ref_grads)
    before = {n: q.detach().clone() for n, q in tr.gen_opt.params.items()}
    same_grads = {n: q.grad.detach().cpu().to(named[n].dtype) for n, q in tr.gen_opt.params.items()}
    tr.gen_opt.step(tr.lr(iteration) if hasattr(tr, 'lr') else None, 1)
    ref_before = {n: p.detach().clone() for n, p in named.items()}
    om.gen_opt.apply(named, same_grads, om.lr(iteration))
    for n, q in tr.gen_opt.params.items():
        report['adam.' + n] = rel_err(q.detach() - before[n], named[n].detach() - ref_before[n])
    tr.rng.end_step()
    return report


def worst(report, prefix=''):
    items = [(v, k) for k, v in report.items() if k.startswith(prefix)]
    return max(items) if items else (0.0, '')


def format_report(report, top=8):
    rows = sorted(((v, k) for k, v in report.items()), reverse=True)[:top]
    return ', '.join('%s=%.2e' % (k, v) for v, k in rows)
