"""Drop-in for the reference's `tflib` package (TG/tflib/__init__.py:8-48).

Same contract: a module-level, name-keyed parameter registry.  `param(name, value)`
creates the parameter the first time a name is seen and returns the SAME object on
every later call, which is what makes `Discriminator(x)` called 3-5 times share its
weights (TG/CT_gan_mnist.py:114-115).  `params_with_name(substr)` selects optimizer
variable lists by substring ('Generator', 'Discriminator.').

Differences that follow from eager PyTorch instead of a TF graph:
  * parameters are float32 `torch.nn.Parameter`s on the current CUDA device, stored under
    the reference names (`X.Filters` HWIO, `X.Biases`, `X.W`, `X.b`, `X.scale`, `X.offset`);
  * `param(..., trainable=False)` (batchnorm moving stats) makes a plain buffer;
  * not thread-safe (neither is the reference: one module-level dict).
"""
import numpy as np
import torch

from .. import functional as _F

_params = {}
_param_aliases = {}
_device = None

# Parameter-name suffixes.  The LSUN fork of the package (LS/tflib/ops/conv2d.py:117, batchnorm.py:24, layernorm.py:15, with
# LS = TG/LSUN_bedrooms) names conv biases and normalisation offsets `<name>.b`; the CT scripts' package uses
# `<name>.Biases` / `<name>.offset`.  `set_name_style('lsun')` before building a model selects the fork's names.
CONV_BIAS, NORM_OFFSET = '.Biases', '.offset'


def set_name_style(style):
    global CONV_BIAS, NORM_OFFSET
    if style not in ('ct', 'lsun'):
        raise Exception('unknown name style %r' % (style,))
    CONV_BIAS, NORM_OFFSET = ('.b', '.b') if style == 'lsun' else ('.Biases', '.offset')


def set_device(device):
    """Device new parameters are created on (default: current CUDA device)."""
    global _device
    _device = torch.device(device)


def device():
    if _device is not None:
        return _device
    if not torch.cuda.is_available():
        raise RuntimeError('ctgan_b200.tflib: no CUDA device; parameters live in HBM and there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def param(name, *args, **kwargs):
    """Create-or-reuse (TG/tflib/__init__.py:10-34).  `args[0]` is the initial value
    (numpy array), `trainable=False` mirrors tf.Variable(trainable=False)."""
    if name not in _params:
        value = np.asarray(args[0] if args else kwargs['initial_value'], dtype='float32')
        trainable = kwargs.get('trainable', True)
        t = torch.from_numpy(np.ascontiguousarray(value)).to(device())
        p = torch.nn.Parameter(t, requires_grad=trainable)
        p.param = True
        p.ctgan_name = name
        _params[name] = p
        _F.register_param(p)
    result = _params[name]
    while result in _param_aliases:
        result = _param_aliases[result]
    return result


def has_param(name):
    return name in _params


def params_with_name(name):
    """TG/tflib/__init__.py:36-37: every registered variable whose name contains `name`
    (includes non-trainable moving stats, like the reference)."""
    return [p for n, p in _params.items() if name in n]


def named_params_with_name(name, trainable_only=True):
    return {n: p for n, p in _params.items() if name in n and (p.requires_grad or not trainable_only)}


def delete_all_params():
    """TG/tflib/__init__.py:39-40.  Also drops the aliases (the reference keeps them: a stale alias would redirect a
    name of the next model built in this process to a parameter of the deleted one)."""
    global CONV_BIAS, NORM_OFFSET
    _params.clear()
    _param_aliases.clear()
    _F._param_ptrs.clear()
    CONV_BIAS, NORM_OFFSET = '.Biases', '.offset'          # the next model states its own name style


def alias_params(replace_dict):
    for old, new in replace_dict.items():
        _param_aliases[old] = new


def delete_param_aliases():
    _param_aliases.clear()


def rebind_params(mapping):
    """Replace registered parameters by new Parameter objects (used when an optimizer moves
    them into one flat buffer).  mapping: name -> new Parameter."""
    for n, p in mapping.items():
        p.param = True
        p.ctgan_name = n
        _params[n] = p
        _F.register_param(p)


def print_model_settings(locals_):
    """TG/tflib/__init__.py:101-106."""
    print("Uppercase local vars:")
    all_vars = [(k, v) for (k, v) in locals_.items()
                if (k.isupper() and k != 'T' and k != 'SETTINGS' and k != 'ALL_SETTINGS')]
    for var_name, var_value in sorted(all_vars, key=lambda x: x[0]):
        print("\t{}: {}".format(var_name, var_value))


def print_model_settings_dict(settings):
    print("Settings dict:")
    for var_name, var_value in sorted(settings.items(), key=lambda x: x[0]):
        print("\t{}: {}".format(var_name, var_value))
