"""Runs the reference's OWN source for the hot path in this container -- TEST INFRASTRUCTURE.

The reference is Python 2 + TensorFlow 1.2.1 and cannot be imported here.  This harness
  1. translates the needed reference files py2->py3 on the fly (print statements, xrange, integer `/` in
     the few batch-split expressions) -- the translated text is exec'd / written under oracle/_ref/
     (git-ignored), never committed;
  2. executes TG/tflib/__init__.py and TG/tflib/ops/{conv2d,deconv2d,linear,batchnorm,cond_batchnorm}.py
     unmodified otherwise, with `tensorflow` bound to oracle/tf_shim;
  3. executes the hyper-parameter, model-function and loss-graph SECTIONS of TG/CT_gan_mnist.py,
     TG/CT_gan_cifar.py, TG/CT_gan_cifar_resnet.py (line ranges below) in that namespace, feeding
     `tf.placeholder`s with concrete tensors;
  4. returns the losses, the GP gradient, every parameter and every parameter gradient, plus the random
     draws in graph order mapped onto the oracle's tags.
`tests/test_oracle_vs_reference.py` compares the oracle restatement against these runs and
`tests/golden/make_golden.py` stores them as fixtures.  What this pins: the reference's own graph code, parameter
naming/layout, init formulas, loss formulas.  What it cannot pin: TensorFlow's internal kernels (absent).
"""
import importlib
import importlib.util
import os
import re
import sys
import textwrap
import types

import numpy as np
import torch

REF_ROOT = '/root/reference/CT-GANs/tensorflow_generative_model'
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.isdir(REF_ROOT)


def py2to3(src):
    """The py2 idioms that occur in the reference files on the path."""
    out = []
    for line in src.splitlines():
        m = re.match(r'^(\s*)print\s+(?!\()(.*)$', line)
        if m:
            line = '%sprint(%s)' % (m.group(1), m.group(2))
        line = line.replace('xrange(', 'range(').replace('import cPickle as pickle', 'import pickle')
        # py2 int/int in the batch-split expressions (tensor / int stays a true division)
        line = line.replace('BATCH_SIZE/len(', 'BATCH_SIZE//len(').replace('BATCH_SIZE / len(', 'BATCH_SIZE // len(')
        line = line.replace('len(DEVICES)/2', 'len(DEVICES)//2')
        out.append(line)
    return '\n'.join(out) + '\n'


def _load_ref_tflib(shim, fork=None):
    """Import the reference's tflib package (translated) with `tensorflow` -> shim.  Returns the module.
    fork='LSUN_bedrooms': the copy of the package that ships next to LS/wgan_LSUN_Bedrooms128.py (other parameter names)."""
    pkg_dir = os.path.join(OUT_DIR, 'tflib')
    os.makedirs(os.path.join(pkg_dir, 'ops'), exist_ok=True)
    src_root = os.path.join(REF_ROOT, fork, 'tflib') if fork else os.path.join(REF_ROOT, 'tflib')
    rels = ['__init__.py', 'ops/__init__.py', 'ops/conv2d.py', 'ops/deconv2d.py', 'ops/linear.py',
            'ops/batchnorm.py', 'ops/cond_batchnorm.py', 'ops/layernorm.py']
    stale = os.path.join(pkg_dir, 'debug.py')
    if os.path.exists(stale):
        os.remove(stale)
    if fork:
        rels.append('debug.py')                        # imported by the fork's conv2d.py / batchnorm.py
    for rel in rels:
        path = os.path.join(src_root, rel)
        if rel == 'ops/__init__.py' and not os.path.exists(path):
            src = ''                                   # the fork ships only the compiled ops/__init__.pyc
        else:
            with open(path) as f:
                src = f.read()
            if rel == 'debug.py':                      # only imported; its print_all_stats (a multi-line py2 print) is dropped
                src = src[:src.index('def print_all_stats')]
            src = py2to3(src)
        if rel == '__init__.py':
            src = src.replace("locale.setlocale(locale.LC_ALL, '')", "pass")
        with open(os.path.join(pkg_dir, rel), 'w') as f:
            f.write(src)
    for k in [k for k in sys.modules if k == 'tflib' or k.startswith('tflib.')]:
        del sys.modules[k]
    saved_tf = sys.modules.get('tensorflow')
    sys.modules['tensorflow'] = shim
    sys.path.insert(0, OUT_DIR)
    try:
        lib = importlib.import_module('tflib')
        for m in ('conv2d', 'deconv2d', 'linear', 'batchnorm', 'cond_batchnorm', 'layernorm'):
            importlib.import_module('tflib.ops.' + m)
    finally:
        sys.path.remove(OUT_DIR)
        if saved_tf is None:
            del sys.modules['tensorflow']
        else:
            sys.modules['tensorflow'] = saved_tf
    return lib


def py2to3_host(src):
    """Additional py2 idioms of the host-side utility modules (tflib/cifar10.py, mnist.py, plot.py, save_images.py):
    integer `/` on python ints and list-returning dict views."""
    src = py2to3(src)
    for old, new in [('len(images) / batch_size', 'len(images) // batch_size'),            # cifar10.py:33,60
                     ('np.mean(vals.values())', 'np.mean(list(vals.values()))'),            # plot.py:25
                     ('np.sort(_since_beginning[name].keys())', 'np.sort(list(_since_beginning[name].keys()))'),   # plot.py:28
                     ('n_samples/rows', 'n_samples//rows'),                                 # save_images.py:19
                     ('j = n/nw', 'j = n//nw'),                                             # save_images.py:34
                     ('files = range(n_files)', 'files = list(range(n_files))')]:           # small_imagenet.py:9 (py2 range is a list)
        src = src.replace(old, new)
    return src


def load_ref_host_module(name, stubs=None, fork=None):
    """Import ONE host-side utility module of the reference's tflib (translated), e.g. 'cifar10', 'mnist', 'plot',
    'save_images', as a stand-alone module.  stubs: {module name: module object} bound in sys.modules during the import
    (matplotlib / scipy.misc are not installed here).  fork='LSUN_bedrooms': the copy of tflib next to the LSUN script."""
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(REF_ROOT, fork, 'tflib', name + '.py') if fork else os.path.join(REF_ROOT, 'tflib', name + '.py')) as f:
        src = py2to3_host(f.read())
    if fork:
        src = src.replace('shell=True).split("\\n")', 'shell=True).decode().split("\\n")')     # py3: check_output returns bytes
    path = os.path.join(OUT_DIR, 'ref_host_%s%s.py' % (name, '_' + fork if fork else ''))
    with open(path, 'w') as f:
        f.write(src)
    saved = {}
    for k, v in (stubs or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        spec = importlib.util.spec_from_file_location('ref_host_' + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                del sys.modules[k]
            else:
                sys.modules[k] = v
    return mod


def _section(path, first, last, dedent=False):
    with open(path) as f:
        lines = f.read().splitlines()
    src = '\n'.join(lines[first - 1:last]) + '\n'
    if dedent:
        src = textwrap.dedent(src)
    return py2to3(src)


# (constants, functions, graph) line ranges, 1-based inclusive, of the reference scripts
SECTIONS = {
    'mnist': dict(file='CT_gan_mnist.py', consts=(26, 35), funcs=(39, 108), graph=(110, 167), dedent=False),
    'cifar': dict(file='CT_gan_cifar.py', consts=(34, 43), funcs=(47, 100), graph=(102, 151), dedent=False),
    'resnet': dict(file='CT_gan_cifar_resnet.py', consts=(37, 56), funcs=(67, 186), graph=(190, 330), dedent=True),
    # hyper-parameters, EVERY architecture function of the file (only GoodGenerator / GoodDiscriminator
    # are called), `Generator, Discriminator = GeneratorAndDiscriminator()`, then the two-tower loss graph
    '64x64': dict(file='CT_gan_64x64.py', consts=(28, 37), funcs=(41, 469), graph=(473, 546), dedent=True),
    # N4, second half: LS/wgan_LSUN_Bedrooms128.py with the fork of tflib next to it.  consts = hyper-parameters +
    # GeneratorAndDiscriminator (:31-60) and OUTPUT_DIM / DEVICES (:64-65); graph = placeholders .. disc_cost + decay (:211-288)
    # and the generator cost (:291-295) -- the two AdamOptimizer(...).minimize lines (:289, :296) are not executed
    'lsun128': dict(file='LSUN_bedrooms/wgan_LSUN_Bedrooms128.py', fork='LSUN_bedrooms', consts=[(31, 60), (64, 65)], funcs=(67, 205),
                    graph=[(211, 288), (291, 295)], dedent=True),
}

# random draws of the reference graph, in graph-construction order, mapped onto the oracle's tags
# (None = the draw feeds nothing the training step uses)
def _draw_tags(script):
    D3 = lambda p: [p + '.1', p + '.2', p + '.3']
    if script in ('mnist', 'cifar'):
        tags = ['z']                                   # fake_data = Generator(BATCH_SIZE)
        if script == 'cifar':
            tags += [None]                             # fake_data_2 = Generator(BATCH_SIZE)      (:105, unused)
        tags += D3('drop.real1') + D3('drop.real2') + D3('drop.fake') + [None] * 3     # 4 critic calls, last unused
        if script == 'cifar':
            tags += [None]                             # fake_data_2 = Generator(...) again       (:117, unused)
        tags += ['alpha'] + D3('drop.gp')
        if script == 'cifar':
            tags += [None] * 3                         # Discriminator(real_data) for `gradients2` (:145, dev metric)
        return tags
    if script == 'lsun128':                            # G x2, D(real+fake) x2, alpha, D(interpolates); then per device G, D
        tags = ['z.0', 'z.1'] + D3('drop.p1') + D3('drop.p2') + ['alpha'] + D3('drop.gp')
        tags += ['z.0g', 'drop.0.1', 'drop.0.2', 'drop.0.3', 'z.1g', 'drop.1.1', 'drop.1.2', 'drop.1.3']
        return tags
    if script == '64x64':                              # per tower: G, D(real) x2, D(fake), alpha, D(interpolates)
        tags = []
        for i in range(2):
            tags += ['z.%d' % i] + D3('drop.%d.real1' % i) + D3('drop.%d.real2' % i) + D3('drop.%d.fake' % i)
            tags += ['alpha.%d' % i] + D3('drop.%d.gp' % i)
        return tags
    tags = ['z.0', 'z.1', 'dequant'] + D3('drop.p1') + D3('drop.p2')     # clean pass: keep_prob 1 -> no draw
    tags += ['alpha'] + D3('drop.gp')
    tags += ['labels.0', 'z.0g', 'drop.0.1', 'drop.0.2', 'drop.0.3', 'labels.1', 'z.1g', 'drop.1.1', 'drop.1.2', 'drop.1.3']
    return tags


def _ranges(r):
    return [r] if isinstance(r[0], int) else list(r)


def run_reference(script, batch_size, seed, inputs, dim=None, param_init=None, width=None):
    """Execute the reference's own code for one evaluation of disc_cost / gen_cost.
    inputs: tuple of numpy arrays fed to the script's placeholders (real data[, labels]).
    param_init(name, value) -> value: optional replacement of every parameter's initial value at the moment the
    reference's lib.param() creates it (full-width fixtures regenerate their weights from a formula instead of storing them).
    Returns dict(params, disc, gen, tape_disc, tape_gen, disc_grads, gen_grads, gp_gradients)."""
    from . import tf_shim as shim
    sec = SECTIONS[script]
    path = os.path.join(REF_ROOT, sec['file'])
    shim.reset_variables()
    lib = _load_ref_tflib(shim, sec.get('fork'))
    lib.delete_all_params()
    if param_init is not None:
        ref_param = lib.param

        def param(name, *args, **kwargs):
            if name not in lib._params and args:
                args = (param_init(name, np.asarray(args[0])),) + tuple(args[1:])
            return ref_param(name, *args, **kwargs)
        lib.param = param
    ns = {'tf': shim, 'lib': lib, 'np': np, 'functools': importlib.import_module('functools'), '__name__': 'ref_section'}
    if script == 'lsun128':
        ns['N_GPUS'] = 2                               # :6
    for rng_ in _ranges(sec['consts']):
        exec(compile(_section(path, *rng_), sec['file'] + ':consts', 'exec'), ns)
    ns['BATCH_SIZE'] = batch_size
    if width is not None:                              # narrower model (LSUN script: every DIM_G_* / DIM_D_* constant)
        for k in list(ns):
            if k.startswith(('DIM_G_', 'DIM_D_')):
                ns[k] = max(1, int(ns[k] * width))
    if dim is not None:                                # smaller model for the committed golden fixtures
        for k in ('DIM', 'DIM_G', 'DIM_D'):
            if k in ns:
                ns[k] = dim
    if script == 'resnet':
        ns['N_GPUS'] = 1
        ns['DEVICES'] = ['/gpu:0', '/gpu:0']           # :61-63 with N_GPUS == 1
    exec(compile(_section(path, *sec['funcs']), sec['file'] + ':funcs', 'exec'), ns)
    np.random.seed(seed)                               # the reference draws initial weights from numpy's global RNG
    feeds = [torch.from_numpy(np.asarray(a)) for a in inputs]
    if script in ('resnet', 'lsun128'):
        feeds = [torch.tensor(0, dtype=torch.int32)] + feeds          # _iteration placeholder (:190 / LS :211)
    if script == 'lsun128':
        ns['Generator'], ns['Discriminator'] = ns['GeneratorAndDiscriminator']()      # LS :209
    shim.reset(seed + 1, feeds)
    for rng_ in _ranges(sec['graph']):
        graph_src = _section(path, *rng_, dedent=sec['dedent'])
        exec(compile(graph_src, sec['file'] + ':graph', 'exec'), ns)
    draws = list(shim.draws)
    tags = _draw_tags(script)
    assert len(draws) == len(tags), (len(draws), len(tags), [k for k, _ in draws])
    tape_disc, tape_gen = {}, {}
    for tag, (kind, t) in zip(tags, draws):
        if tag is None:
            continue
        if script in ('resnet', 'lsun128') and (tag.endswith('g') or tag.startswith('labels.') or tag.startswith('drop.0') or tag.startswith('drop.1')):
            t2 = t
            if tag.startswith('labels.'):
                t2 = torch.floor(t * np.float32(10)).to(torch.int32)
            tape_gen[tag[:-1] if tag.endswith('g') else tag] = t2
        else:
            tape_disc[tag] = t
    if script == '64x64':
        tape_gen = {k: v for k, v in tape_disc.items() if k.startswith('z.') or k.endswith(('.fake.1', '.fake.2', '.fake.3'))}
    elif script not in ('resnet', 'lsun128'):
        tape_gen = {k: v for k, v in tape_disc.items() if k == 'z' or k.startswith('drop.fake')}
    params = {n: p for n, p in lib._params.items()}
    disc_sel = 'Discriminator.' if script in ('resnet', '64x64', 'lsun128') else 'Discriminator'
    dnames = [n for n, p in params.items() if disc_sel in n and p.requires_grad]
    gnames = [n for n, p in params.items() if 'Generator' in n and p.requires_grad]
    dgr = torch.autograd.grad(ns['disc_cost'], [params[n] for n in dnames], retain_graph=True, allow_unused=True)
    ggr = torch.autograd.grad(ns['gen_cost'], [params[n] for n in gnames], retain_graph=True, allow_unused=True)
    out = dict(
        params={n: p.detach().clone() for n, p in params.items()},
        trainable={n: bool(p.requires_grad) for n, p in params.items()},
        disc_cost=ns['disc_cost'].detach(), gen_cost=ns['gen_cost'].detach(),
        gp_gradients=ns['gradients'].detach(),
        disc_grads={n: (g.detach() if g is not None else None) for n, g in zip(dnames, dgr)},
        gen_grads={n: (g.detach() if g is not None else None) for n, g in zip(gnames, ggr)},
        tape_disc=tape_disc, tape_gen=tape_gen,
    )
    for k in ('gradient_penalty', 'CT_', 'disc_wgan', 'disc_acgan'):
        if k in ns and isinstance(ns[k], torch.Tensor):
            out[k] = ns[k].detach()
    return out
