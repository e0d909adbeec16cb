"""Diagnostic: which conv geometries of a script's step reach the SIMT kernels (ctgan_conv_fprop / _dgrad / _wgrad).
    python tests/which_simt.py cifar|mnist|64x64|lsun128|cifar_resnet"""
import collections
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from ctgan_b200 import _lib
import ctgan_b200.kernels as K

script = sys.argv[1] if len(sys.argv) > 1 else '64x64'
mod = importlib.import_module('ctgan_b200.gan_' + script)
seen = collections.Counter()
orig = _lib.call


def call(name, *a):
    if name in ('ctgan_conv_fprop', 'ctgan_conv_dgrad', 'ctgan_conv_wgrad'):
        d = a[0]._obj
        seen[(name, d.N, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout, d.kh, d.stride, d.x_dtype, d.y_dtype)] += 1
    return orig(name, *a)


_lib.call = call
K.call = call
np.random.seed(0)
B = mod.BATCH_SIZE
tr = mod.Trainer(device='cuda', seed=1, act_dtype=torch.bfloat16, batch_size=B)
rs = np.random.RandomState(0)
if script == '64x64':
    args = (torch.from_numpy(rs.randint(0, 256, (B, 3, 64, 64)).astype('int32')).cuda(),)
elif script == 'lsun128':
    args = (torch.from_numpy(rs.randint(0, 256, (B, 3, 128, 128)).astype('int32')).cuda(),)
elif script == 'mnist':
    args = (torch.from_numpy(rs.random_sample((B, 784)).astype('float32')).cuda(),)
elif script == 'cifar':
    args = (torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda(),)
else:
    args = (torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda(),
            torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda())
tr.critic_step(*args)
tr.gen_step()
torch.cuda.synchronize()
for k, v in sorted(seen.items(), key=lambda kv: -kv[1]):
    print(v, k)
