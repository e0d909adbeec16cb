"""lib.ops.batchnorm.Batchnorm -- drop-in for TG/tflib/ops/batchnorm.py:6-87.

Training-mode batch norm, biased variance, eps 1e-5, `<name>.scale` / `<name>.offset`.
axes [0,2,3] (fused path, per channel; also creates the non-trainable moving stats the
reference creates at batchnorm.py:26-27) and axes [0] on [batch, features] (params shaped
[1, features], batchnorm.py:78-83).  `is_training` tensors (inference/moving-stat updates)
are never passed by the CT-GAN scripts -> unsupported.  Extensions: `relu=True` fuses the
tf.nn.relu that follows every Batchnorm call in the scripts into the same kernel; `groups=G` computes the
statistics per block of N/G consecutive samples (the reference's per-device-split generator calls run as one batch).
"""
import numpy as np

from ... import tflib as lib
from ... import functional as F


def Batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True,
              relu=False, groups=1, up2=False):
    if is_training is not None:
        raise Exception('Unsupported configuration')
    if ((axes == [0, 2, 3]) or (axes == [0, 2])) and fused:
        if axes == [0, 2]:
            raise Exception('Unsupported configuration')
        inputs = F.ensure_nhwc(inputs)
        C = inputs.shape[1]
        offset = lib.param(name + lib.NORM_OFFSET, np.zeros(C, dtype='float32'))
        scale = lib.param(name + '.scale', np.ones(C, dtype='float32'))
        lib.param(name + '.moving_mean', np.zeros(C, dtype='float32'), trainable=False)
        lib.param(name + '.moving_variance', np.ones(C, dtype='float32'), trainable=False)
        return F.batch_norm(inputs, scale, offset, None, 1e-5, relu, groups, up2)
    if axes == [0] and inputs.dim() == 2:
        shape = [1, inputs.shape[1]]
        offset = lib.param(name + lib.NORM_OFFSET, np.zeros(shape, dtype='float32'))
        scale = lib.param(name + '.scale', np.ones(shape, dtype='float32'))
        return F.batch_norm(inputs, scale, offset, None, 1e-5, relu, groups, up2)
    if axes == [0, 2, 3]:          # unfused spelling of the same statistics (params shaped [1,C,1,1])
        inputs = F.ensure_nhwc(inputs)
        shape = [1, inputs.shape[1], 1, 1]
        offset = lib.param(name + lib.NORM_OFFSET, np.zeros(shape, dtype='float32'))
        scale = lib.param(name + '.scale', np.ones(shape, dtype='float32'))
        return F.batch_norm(inputs, scale, offset, None, 1e-5, relu, 1, up2)
    raise Exception('Unsupported configuration')
