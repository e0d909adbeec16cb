"""lib.ops.cond_batchnorm.Batchnorm -- drop-in for TG/tflib/ops/cond_batchnorm.py:6-17.

Conditional batch norm: batch moments over [0,2,3], per-class `<name>.scale` /
`<name>.offset` tables [n_labels, C] gathered by label inside the kernel.
"""
import numpy as np

from ... import tflib as lib
from ... import functional as F


def Batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True,
              labels=None, n_labels=None, relu=False, groups=1, up2=False):
    """conditional batchnorm (dumoulin et al 2016) for BCHW conv filtermaps"""
    if axes != [0, 2, 3]:
        raise Exception('unsupported')
    inputs = F.ensure_nhwc(inputs)
    C = inputs.shape[1]
    offset_m = lib.param(name + '.offset', np.zeros([n_labels, C], dtype='float32'))
    scale_m = lib.param(name + '.scale', np.ones([n_labels, C], dtype='float32'))
    return F.batch_norm(inputs, scale_m, offset_m, labels, 1e-5, relu, groups, up2)
