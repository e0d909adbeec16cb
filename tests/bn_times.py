"""CUDA-event timings of the fused BF16 batch-norm kernels alone (rotating buffers > L2), as effective HBM GB/s of their
algorithmic bytes: forward 2 reads + 1 write of x-sized tensors (sums pass + apply pass), backward 4 reads + 1 write."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K
CL = torch.channels_last
def act(n, c, h): return torch.randn(n, c, h, h, device='cuda').to(torch.bfloat16).contiguous(memory_format=CL)
def t(fn, reps=40):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (N, C, H, groups) in ((128, 128, 32, 2), (320, 128, 32, 10), (128, 128, 16, 2), (128, 128, 8, 2), (1024, 128, 32, 2), (4096, 128, 32, 2)):
    S = max(2, int(300e6 // (N * C * H * H * 2)) + 1)
    xs = [act(N, C, H) for _ in range(S)]; dys = [act(N, C, H) for _ in range(S)]
    gam, bet = torch.ones(10, C, device='cuda'), torch.zeros(10, C, device='cuda')
    lab = torch.randint(0, 10, (N,), dtype=torch.int32, device='cuda')
    y, mean, invstd = K.bn_fwd(xs[0], gam, bet, lab, 1e-5, True, groups)
    nb = N * C * H * H * 2
    f = t(lambda i: K.bn_fwd(xs[i % S], gam, bet, lab, 1e-5, True, groups))
    b = t(lambda i: K.bn_bwd(dys[i % S], xs[i % S], None, gam, bet, lab, mean, invstd, True, groups))
    print('BN %dx%dx%dx%d groups %d: fwd %.1f us = %.0f GB/s (3 x %.1f MB)   bwd %.1f us = %.0f GB/s (5 x %.1f MB)' % (
        N, H, H, C, groups, f, 3 * nb / f / 1e3, nb / 1e6, b, 5 * nb / b / 1e3, nb / 1e6))
