"""Generates tests/golden/resnet128.npz: CT_gan_cifar_resnet.py at its SHIPPED width (DIM_G = DIM_D = 128, the shapes the
tcgen05 kernels run), batch 8, evaluated by THE REFERENCE'S OWN CODE (oracle/ref_harness.py) in float64.  Run in the build
container (needs /root/reference):      python tests/golden/make_golden_full.py

Kept small (~350 KB): parameters come from tests/golden/det_params.py (regenerated on both sides), dropout draws are
stored as their keep decisions (1 bit each; the test replays u = 0.999 / 0.0), parameter gradients as their norm plus
every 61st element (tensors of <= 4096 elements in full)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH            # noqa: E402
from tests.golden.det_params import det_param   # noqa: E402

B, DIM, SEED, STRIDE = 8, 128, 909, 61
KEEP = {'1': 0.8, '2': 0.5, '3': 0.5}           # Discriminator(..., kp1, kp2, kp3) = (0.8, 0.5, 0.5)


def main(case='resnet128'):
    """case 'lsun128q': LS/wgan_LSUN_Bedrooms128.py at a QUARTER of its widths (critic 32..256 channels, generator 128..16),
    batch 4 -> tests/golden/lsun128q.npz."""
    global B, DIM, SEED
    if case == 'lsun128q':
        B, DIM, SEED = 4, 4, 919                # meta.dim = the divisor of the reference's widths
    rs = np.random.RandomState(SEED)
    if case == 'lsun128q':
        inputs = (rs.randint(0, 256, (B, 3, 128, 128)).astype('uint8').astype('int32'),)
        r = RH.run_reference('lsun128', B, SEED, inputs, width=1.0 / DIM, param_init=lambda n, v: det_param(n, v, SEED))
    else:
        inputs = (rs.randint(0, 256, (B, 3072)).astype('int32'), rs.randint(0, 10, (B,)).astype('int32'))
        r = RH.run_reference('resnet', B, SEED, inputs, dim=DIM, param_init=lambda n, v: det_param(n, v, SEED))
    for n, p in r['params'].items():            # the reference really ran on the formula's weights
        assert np.array_equal(p.numpy().astype('float32'), det_param(n, p.numpy(), SEED)), n
    blob = {'meta.B': np.int64(B), 'meta.dim': np.int64(DIM), 'meta.seed': np.int64(SEED), 'meta.stride': np.int64(STRIDE)}
    for i, a in enumerate(inputs):
        blob['input.%d' % i] = a.astype('uint8') if case == 'lsun128q' else a       # pixels fit a byte: 4x smaller
    for kind in ('disc', 'gen'):
        for t, v in r['tape_' + kind].items():
            a = v.numpy()
            if t.startswith('drop.'):
                keep = KEEP[t.rsplit('.', 1)[1]]
                blob['keepbits_%s.%s' % (kind, t)] = np.packbits(np.floor(keep + a).astype(bool).reshape(-1))
                blob['keepshape_%s.%s' % (kind, t)] = np.asarray(a.shape, dtype='int64')
            else:
                blob['tape_%s.%s' % (kind, t)] = a
        for n, g in r[kind + '_grads'].items():
            if g is not None:
                f = g.numpy().astype('float64').reshape(-1)
                blob['gnorm_%s.%s' % (kind, n)] = np.float64(np.linalg.norm(f))
                blob['gsample_%s.%s' % (kind, n)] = (f if f.size <= 4096 else f[::STRIDE]).astype('float32')   # small tensors in full
    for k in ('disc_cost', 'gen_cost', 'gradient_penalty', 'CT_', 'disc_wgan', 'disc_acgan'):
        if k in r:
            blob[k] = r[k].numpy().astype('float64')
    gp = r['gp_gradients'].numpy().astype('float64')
    if case == 'lsun128q':                      # 4 x 49152 values: the norm and every 61st element
        blob['gp_gradients_norm'] = np.float64(np.linalg.norm(gp))
        blob['gp_gradients_sample'] = gp.reshape(-1)[::STRIDE].astype('float32')
    else:
        blob['gp_gradients'] = gp.astype('float32')
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), case + '.npz')
    np.savez_compressed(path, **blob)
    print('disc_cost %.12f gen_cost %.12f gp %.6f ct %.6f ->' % (float(blob['disc_cost']), float(blob['gen_cost']),
          float(blob['gradient_penalty']), float(blob['CT_'])), path, '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main(*sys.argv[1:2])
