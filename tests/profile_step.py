"""Profiling driver: one eager critic step + one eager generator step of the ResNet CT-GAN
(batch 64, BF16 path) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
and `ncu --set full -k regex:<kernel>` captures.  Not a benchmark (numbers under ncu are never bench values)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import numpy as np
import torch

import ctgan_b200.gan_cifar_resnet as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()
for _ in range(2):
    tr.critic_step(x, y)
    tr.gen_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.critic_step(x, y)
tr.gen_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled 1 critic + 1 generator step at batch', B)
