"""The step's top kernels, launched alone between cudaProfilerStart/Stop, for ONE `ncu --set full` capture:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_top python tests/profile_kernels.py
(profiles/r02_ncu_top_kernels_summary.txt is extracted from that report).  Not a benchmark."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch

import ctgan_b200.kernels as K
from ctgan_b200 import _lib

CL = torch.channels_last
C = 128


def act(n, h, dtype=torch.bfloat16, c=C):
    return torch.randn(n, c, h, h, device='cuda').to(dtype).contiguous(memory_format=CL)


w = (torch.randn(3, 3, C, C, device='cuda') * 0.03).contiguous()
b = torch.zeros(C, device='cuda')
cases = []

# 1. conv_fprop_tc_pair_kernel: Discriminator.1.Conv2 on the stacked pass (192 x 32 x 32)
g = K.same_geom(192, 32, 32, C, C, 3, 1); x = act(192, 32)
cases.append(('pair fprop 192x32x32', lambda g=g, x=x: K.conv_fprop(x, w, b, g)))
# 2. conv_fprop_tc_lean_kernel<1>: Discriminator.2 on the stacked pass (192 x 16 x 16)
g = K.same_geom(192, 16, 16, C, C, 3, 1); x = act(192, 16)
cases.append(('lean<1> fprop 192x16x16', lambda g=g, x=x: K.conv_fprop(x, w, b, g)))
# 3. conv_fprop_tc_lean_kernel<0> (one CTA per tile) and 4. the cluster split-K kernel: 64 x 8 x 8
g = K.same_geom(64, 8, 8, C, C, 3, 1); x = act(64, 8)


def lean0(g=g, x=x):
    with K.splitk(False):
        K.conv_fprop(x, w, b, g)


cases.append(('lean<0> fprop 64x8x8', lean0))
cases.append(('splitk<4> fprop 64x8x8', lambda g=g, x=x: K.conv_fprop(x, w, b, g)))
# 5. conv_wgrad_tc_multi_kernel: the 16 filter gradients of one critic step
jobs = []
for n in (192, 64):
    for (h, k, reps) in ((32, 3, 1), (16, 3, 2), (8, 3, 4), (8, 1, 1)):
        for _ in range(reps):
            jobs.append((act(n, h), act(n, h), K.same_geom(n, h, h, C, C, k, 1), torch.zeros(k, k, C, C, device='cuda')))


def multi():
    for xj, dyj, gj, dwj in jobs:
        K.conv_wgrad(xj, dyj, gj, tuple(dwj.shape), accumulate_into=dwj, defer=True)
    K.flush_wgrads()


cases.append(('wgrad multi (16 jobs)', multi))
# 6. fused batch norm, forward (+ upsample) and backward: Generator.3.N1 on the 128-image generator step (16x16 -> 32x32)
xb = act(128, 16); gam = torch.ones(10, C, device='cuda'); bet = torch.zeros(10, C, device='cuda')
lab = torch.randint(0, 10, (128,), dtype=torch.int32, device='cuda')
yb, mean, invstd = K.bn_fwd(xb, gam, bet, lab, 1e-5, True, 2, up2=True)
dyb = act(128, 32)
cases.append(('bn fwd fused + up2 128x16x16', lambda: K.bn_fwd(xb, gam, bet, lab, 1e-5, True, 2, up2=True)))
cases.append(('bn bwd fused + up2 128x16x16', lambda: K.bn_bwd(dyb, xb, None, gam, bet, lab, mean, invstd, True, 2, up2=True)))
xb2 = act(128, 32)
y2, mean2, invstd2 = K.bn_fwd(xb2, gam, bet, lab, 1e-5, True, 2)
cases.append(('bn fwd fused 128x32x32', lambda: K.bn_fwd(xb2, gam, bet, lab, 1e-5, True, 2)))
cases.append(('bn bwd fused 128x32x32', lambda: K.bn_bwd(dyb, xb2, None, gam, bet, lab, mean2, invstd2, True, 2)))
# 7. kind::tf32: the same Discriminator.1.Conv2 launch on float activations
K.config.tf32 = True
g32 = K.same_geom(192, 32, 32, C, C, 3, 1); x32 = act(192, 32, torch.float32)
cases.append(('tf32 fprop 192x32x32', lambda: K.conv_fprop(x32, w, b, g32)))

for name, fn in cases:          # warm up (packs, attributes)
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in cases:
    fn()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled:', ', '.join(n for n, _ in cases))
