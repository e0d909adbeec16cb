"""Deterministic parameter values for the full-width golden fixture (tests/golden/resnet128.npz): a function of
(name, shape, seed) only, so that the fixture does not have to store 2.3 M weights.  The SAME function feeds the
reference's own code (oracle/ref_harness.run_reference(param_init=...), when the fixture is generated) and the product
(tests/test_golden_gpu.py, on the GPU box).  Scales follow the reference's initialisers: filters / matrices uniform with
the Glorot standard deviation of their shape, biases and offsets small, scales around one."""
import zlib

import numpy as np


def det_param(name, value, seed):
    value = np.asarray(value)
    shape = value.shape
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    if name.endswith('.Filters') and len(shape) == 4:
        k, _, cin, cout = shape
        s = np.sqrt(6.0 / (k * k * (cin + cout)))
        v = rs.uniform(-s, s, size=shape)
    elif name.endswith('.W') and len(shape) == 2:
        s = np.sqrt(6.0 / (shape[0] + shape[1]))
        v = rs.uniform(-s, s, size=shape)
    elif name.endswith('.scale'):
        v = 1.0 + 0.05 * rs.standard_normal(size=shape)
    elif name.endswith(('.moving_mean', '.moving_variance')):
        return value.astype('float32')
    else:                                   # biases, offsets
        v = 0.05 * rs.standard_normal(size=shape)
    return v.astype('float32')
