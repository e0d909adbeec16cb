"""ncu target: the dominant tcgen05 kernels at Discriminator.1.Conv2's shape (the stacked critic pass: 192 images, 32x32, 3x3, 128->128).
  ncu --set full --clock-control none --import-source on -k regex:conv_ -o gpurun_out/prof python tests/ncu_target.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K

N, H, C = 192, 32, 128
g = K.same_geom(N, H, H, C, C, 3, 1)
x = torch.randn(N, C, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
dy = torch.randn(N, C, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
w = (torch.randn(3, 3, C, C, device='cuda') * 0.03).contiguous()
b = torch.zeros(C, device='cuda')
for _ in range(3):
    y = K.conv_fprop(x, w, b, g)
    dw = K.conv_wgrad(x, dy, g, tuple(w.shape))
torch.cuda.synchronize()
print('done', float(y.float().abs().mean()), float(dw.abs().mean()))
