"""Profiling driver: one eager critic step + one eager generator step of a DCGAN script (CT_gan_cifar.py / CT_gan_mnist.py,
BF16 path) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
Not a benchmark (numbers under ncu are never bench values).     python tests/profile_dcgan.py cifar|mnist|64x64 [batch]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import numpy as np
import torch

script = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
mod = importlib.import_module('ctgan_b200.gan_' + script)
B = int(sys.argv[2]) if len(sys.argv) > 2 else mod.BATCH_SIZE
np.random.seed(1234)
tr = mod.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B)
rs = np.random.RandomState(0)
if script == 'cifar':
    x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
elif script == '64x64':
    x = torch.from_numpy(rs.randint(0, 256, (B, 3, 64, 64)).astype('int32')).cuda()
elif script == 'lsun128':
    x = torch.from_numpy(rs.randint(0, 256, (B, 3, 128, 128)).astype('int32')).cuda()
else:
    x = torch.from_numpy(rs.random_sample((B, 784)).astype('float32')).cuda()
for _ in range(2):
    tr.critic_step(x)
    tr.gen_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.critic_step(x)
tr.gen_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled 1 critic + 1 generator step of', script, 'at batch', B)
