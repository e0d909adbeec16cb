"""Generates tests/golden/<script>.npz by EXECUTING THE REFERENCE'S OWN CODE (oracle/ref_harness.py:
TG/tflib/*.py and the model / loss sections of TG/CT_gan_*.py, py2->py3, `tensorflow` bound to
oracle/tf_shim) on seeded inputs.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Model width is reduced (DIM 16 / 8) so that the fixtures stay small; the code path is the reference's.
Each fixture holds: inputs, every random draw of the graph (tagged), every parameter, disc_cost /
gen_cost, the GP gradient and every parameter gradient (float32 copies of the float64 results)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH   # noqa: E402

CASES = {'mnist': dict(B=6, dim=8, seed=101), 'cifar': dict(B=4, dim=8, seed=202), 'resnet': dict(B=4, dim=16, seed=303),
         '64x64': dict(B=4, dim=4, seed=404),
         'lsun128': dict(B=2, dim=32, seed=505)}          # LS/wgan_LSUN_Bedrooms128.py at 1/32 of its widths (meta.dim = the divisor)


def inputs_for(script, B, seed):
    rs = np.random.RandomState(seed)
    if script == 'mnist':
        return (rs.random_sample((B, 784)).astype('float32'),)
    if script == 'cifar':
        return (rs.randint(0, 256, (B, 3072)).astype('int32'),)
    if script == '64x64':
        return (rs.randint(0, 256, (B, 3, 64, 64)).astype('int32'),)
    if script == 'lsun128':
        return (rs.randint(0, 256, (B, 3, 128, 128)).astype('int32'),)
    return (rs.randint(0, 256, (B, 3072)).astype('int32'), rs.randint(0, 10, (B,)).astype('int32'))


def main(only=None):
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for script, c in CASES.items():
        if only and script not in only:
            continue
        inputs = inputs_for(script, c['B'], c['seed'])
        if script == 'lsun128':
            r = RH.run_reference(script, c['B'], c['seed'], inputs, width=1.0 / c['dim'])
        else:
            r = RH.run_reference(script, c['B'], c['seed'], inputs, dim=c['dim'])
        blob = {'meta.B': np.int64(c['B']), 'meta.dim': np.int64(c['dim']), 'meta.seed': np.int64(c['seed'])}
        for i, a in enumerate(inputs):
            blob['input.%d' % i] = a
        for n, p in r['params'].items():
            blob['param.' + n] = p.numpy().astype('float32')
            blob['trainable.' + n] = np.bool_(r['trainable'][n])
        for kind in ('disc', 'gen'):
            for t, v in r['tape_' + kind].items():
                blob['tape_%s.%s' % (kind, t)] = v.numpy()
            for n, g in r[kind + '_grads'].items():
                if g is not None:
                    blob['grad_%s.%s' % (kind, n)] = g.numpy().astype('float32')
        blob['disc_cost'] = r['disc_cost'].numpy().astype('float64')
        blob['gen_cost'] = r['gen_cost'].numpy().astype('float64')
        blob['gp_gradients'] = r['gp_gradients'].numpy().astype('float32')
        path = os.path.join(out_dir, script + '.npz')
        np.savez_compressed(path, **blob)
        print(script, 'disc_cost %.12f gen_cost %.12f' % (float(blob['disc_cost']), float(blob['gen_cost'])),
              '->', path, '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main(sys.argv[1:])
