"""lib.ops.layernorm.Layernorm -- drop-in for TG/tflib/ops/layernorm.py:6-21.

Per-sample moments over `norm_axes` (the CT scripts only use [1,2,3] on BCHW data: TG/CT_gan_64x64.py:91), biased variance,
eps 1e-5, `<name>.offset` / `<name>.scale` of length n_neurons = the size of the first normalised axis (channels)."""
import numpy as np

from ... import tflib as lib
from ... import functional as F


def Layernorm(name, norm_axes, inputs):
    if list(norm_axes) != [1, 2, 3] or inputs.dim() != 4:
        raise Exception('Layernorm over non-standard axes is unsupported')
    inputs = F.ensure_nhwc(inputs)
    n_neurons = inputs.shape[norm_axes[0]]
    offset = lib.param(name + lib.NORM_OFFSET, np.zeros(n_neurons, dtype='float32'))
    scale = lib.param(name + '.scale', np.ones(n_neurons, dtype='float32'))
    return F.layer_norm(inputs, scale, offset, 1e-5)
