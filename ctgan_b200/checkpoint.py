"""Checkpoints under the reference's parameter names and layouts.

The reference keeps its weights in TF variables named `<op name>.Filters` (HWIO), `.Biases`, `.W` ([in, out]), `.b`,
`.scale`, `.offset` (TG/tflib/ops/*.py) and dumps the critic's as `np.save("param.pyn", session.run(disc_params))`
(TG/CT_gan_cifar.py:216-222; the LSUN script uses tf.train.Saver, LS/wgan_LSUN_Bedrooms128.py:367,395).  The
parameters of this package live under the SAME names and layouts (float32 masters inside FlatAdam's flat buffers), so
a checkpoint is simply {name: float32 array} -- loadable into a TF graph of the reference by name, and vice versa --
plus, optionally, what an exact resume needs: the Adam moments and step counts of both optimizers, the Philox state (seed,
device counter, host offset), the iteration (the ResNet script's learning-rate decay depends on it) and numpy's global
RandomState (the loaders' epoch shuffles).
"""
import numpy as np
import torch

from . import tflib as lib
from . import kernels as K

_OPT_KEYS = ('gen_opt', 'disc_opt')


def state_dict(trainer=None, iteration=None):
    """{reference name: float32 numpy array} of every registered parameter (+ optimizer / random state of `trainer`)."""
    blob = {'param/' + n: p.detach().cpu().numpy().astype('float32') for n, p in lib._params.items()}
    if trainer is not None:
        for key in _OPT_KEYS:
            opt = getattr(trainer, key)
            blob['opt/%s/t' % key] = np.int64(opt.t)
            for n in opt.params:
                o, sz = opt.offsets[n], opt.sizes[n]
                shape = tuple(opt.params[n].shape)
                blob['opt/%s/m/%s' % (key, n)] = opt.flat_m[o:o + sz].reshape(shape).cpu().numpy()
                blob['opt/%s/v/%s' % (key, n)] = opt.flat_v[o:o + sz].reshape(shape).cpu().numpy()
        rng = trainer.rng
        blob['rng/seed'] = np.uint64(rng.seed)
        blob['rng/offset'] = np.int64(rng.offset)
        blob['rng/dyn'] = np.int64(int(rng.dyn.item()) if rng.dyn is not None else -1)
        st = np.random.get_state()
        blob['np_rng/keys'], blob['np_rng/pos'] = st[1], np.int64(st[2])
        blob['np_rng/gauss'] = np.array([st[3], st[4]], dtype='float64')
    if iteration is not None:
        blob['iteration'] = np.int64(iteration)
    return blob


def save(path, trainer=None, iteration=None):
    np.savez(path, **state_dict(trainer, iteration))


def load(path_or_blob, trainer=None, strict=True):
    """Copy a checkpoint into the registered parameters IN PLACE (they stay views of the optimizers' flat buffers),
    restore optimizer state when `trainer` is given and the checkpoint has it, and re-pack the BF16 filter operands."""
    blob = np.load(path_or_blob) if isinstance(path_or_blob, str) else path_or_blob
    names = {k[len('param/'):] for k in blob.keys() if k.startswith('param/')}
    if strict and names != set(lib._params):
        raise KeyError('checkpoint / model parameter names differ: %s' % sorted(names ^ set(lib._params))[:8])
    with torch.no_grad():
        for n in names & set(lib._params):
            p, v = lib._params[n], np.asarray(blob['param/' + n])
            if tuple(p.shape) != v.shape:
                raise ValueError('%s: checkpoint shape %s, model shape %s' % (n, v.shape, tuple(p.shape)))
            p.copy_(torch.from_numpy(np.ascontiguousarray(v, dtype='float32')).to(p.device))
        if trainer is not None:
            for key in _OPT_KEYS:
                if 'opt/%s/t' % key not in blob:
                    continue
                opt = getattr(trainer, key)
                opt.t = int(blob['opt/%s/t' % key])
                for n in opt.params:
                    o, sz = opt.offsets[n], opt.sizes[n]
                    for which, flat in (('m', opt.flat_m), ('v', opt.flat_v)):
                        a = np.ascontiguousarray(blob['opt/%s/%s/%s' % (key, which, n)], dtype='float32').reshape(-1)
                        flat[o:o + sz].copy_(torch.from_numpy(a).to(flat.device))
        if trainer is not None and 'rng/offset' in blob:
            rng = trainer.rng
            rng.seed = int(blob['rng/seed'])
            rng.offset = int(blob['rng/offset'])
            if rng.dyn is not None and int(blob['rng/dyn']) >= 0:
                rng.dyn.fill_(int(blob['rng/dyn']))
            elif rng.dyn is None and int(blob['rng/dyn']) > 0:
                rng.offset += int(blob['rng/dyn'])        # saved in graph mode, resumed eagerly: one counter
            np.random.set_state(('MT19937', np.asarray(blob['np_rng/keys'], dtype='uint32'), int(blob['np_rng/pos']),
                                 int(blob['np_rng/gauss'][0]), float(blob['np_rng/gauss'][1])))
    K.invalidate_weight_cache(None if trainer is None else (trainer.gen_opt._ptrs | trainer.disc_opt._ptrs))
    if trainer is not None:
        trainer.gen_opt.refresh_packs()
        trainer.disc_opt.refresh_packs()
    return {'iteration': int(blob['iteration'])} if 'iteration' in blob else {}


def save_disc_params_pyn(path='param.pyn', selector='Discriminator'):
    """TG/CT_gan_cifar.py:216-222: `np.save("param.pyn", session.run(disc_params))` -- the critic's variables as one
    object array, in creation order (numpy appends '.npy')."""
    para = [p.detach().cpu().numpy() for p in lib.params_with_name(selector)]
    arr = np.empty(len(para), dtype=object)
    for i, a in enumerate(para):
        arr[i] = a
    np.save(path, arr, allow_pickle=True)
    return para
