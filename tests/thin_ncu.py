"""ncu target: one eager call of each member of the thin (3-channel-side) conv family."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K
for (N, H, Cin, Cout, k) in [(192, 32, 3, 128, 3), (128, 32, 128, 3, 3)]:
    g = K.same_geom(N, H, H, Cin, Cout, k, 1)
    x = torch.randn(N, Cin, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(N, Cout, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(k, k, Cin, Cout, device='cuda') * 0.05).contiguous(); b = torch.zeros(Cout, device='cuda')
    dw = torch.zeros_like(w)
    for i in range(2):
        if i == 1: torch.cuda.synchronize(); torch.cuda.profiler.start()
        K.conv_fprop(x, w, b, g, w_is_param=True)
        K.conv_dgrad(dy, w, g, w_is_param=True)
        K.conv_wgrad(x, dy, g, tuple(w.shape), accumulate_into=dw)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
