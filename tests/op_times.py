"""Per-op device time (CUDA events, warm, 20 reps) at the ResNet CT-GAN shapes. Diagnostic, not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K

CL = torch.channels_last
dev = 'cuda'


def act(N, C, H, W, dt=torch.bfloat16):
    return torch.randn(N, C, H, W, device=dev).to(dt).contiguous(memory_format=CL)


def timeit(name, fn, reps=20, flops=None, bytes_=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = ''
    if flops:
        extra += '  %7.1f TFLOP/s' % (flops / us / 1e6)
    if bytes_:
        extra += '  %7.1f GB/s' % (bytes_ / us / 1e3)
    print('%-44s %9.1f us%s' % (name, us, extra), flush=True)


def conv_suite(tag, N, H, Cin, Cout, k):
    g = K.same_geom(N, H, H, Cin, Cout, k, 1)
    x, dy = act(N, Cin, H, H), act(N, Cout, H, H)
    w = (torch.randn(k, k, Cin, Cout, device=dev) * 0.05).contiguous()
    b = torch.zeros(Cout, device=dev)
    fl = 2.0 * N * H * H * Cin * Cout * k * k
    timeit('%s fprop' % tag, lambda: K.conv_fprop(x, w, b, g, w_is_param=False), flops=fl)
    timeit('%s dgrad' % tag, lambda: K.conv_dgrad(dy, w, g), flops=fl)
    timeit('%s wgrad' % tag, lambda: K.conv_wgrad(x, dy, g, tuple(w.shape)), flops=fl)


print('--- conv family (includes filter pack + output alloc per call)')
conv_suite('D.1.Conv2 3x3 128->128 32x32 b128', 128, 32, 128, 128, 3)
conv_suite('D.2.Conv  3x3 128->128 16x16 b128', 128, 16, 128, 128, 3)
conv_suite('D.3.Conv  3x3 128->128  8x8  b128', 128, 8, 128, 128, 3)
conv_suite('D.3.Conv  3x3 128->128  8x8  b64 ', 64, 8, 128, 128, 3)
conv_suite('D.2.Short 1x1 128->128 16x16 b128', 128, 16, 128, 128, 1)
conv_suite('D.1.Conv1 3x3   3->128 32x32 b128', 128, 32, 3, 128, 3)
conv_suite('D.1.Short 1x1   3->128 16x16 b128', 128, 16, 3, 128, 1)
conv_suite('G.Output  3x3 128->3   32x32 b64 ', 64, 32, 128, 3, 3)
conv_suite('G.3.Conv  3x3 128->128 32x32 b64 ', 64, 32, 128, 128, 3)
g = K.ConvGeom(64, 1, 1, 128, 1, 1, 2048, 1, 1, 1, 0, 0)
x2, w2, b2 = torch.randn(64, 128, device=dev).bfloat16(), torch.randn(128, 2048, device=dev) * 0.05, torch.zeros(2048, device=dev)
timeit('G.Input linear 128->2048 b64 fprop', lambda: K.conv_fprop(x2, w2, b2, g))
w = (torch.randn(3, 3, 128, 128, device=dev) * 0.05).contiguous()
timeit('pack_filter 3x3x128x128', lambda: K.pack_filter(w, 0))

print('--- element-wise / norm at [128,128,32,32] bf16 (33.5 MB)')
a, b = act(128, 128, 32, 32), act(128, 128, 32, 32)
nb = a.numel() * 2
timeit('add', lambda: K.add(a, b), bytes_=3 * nb)
timeit('mul', lambda: K.mul(a, b), bytes_=3 * nb)
timeit('act_dropout relu', lambda: K.act_dropout(a, 0.0, 1.0), bytes_=3 * nb)
timeit('act_dropout drop .5', lambda: K.act_dropout(a, 1.0, 0.5, seed=1, offset=0), bytes_=3 * nb)
timeit('pool2x2', lambda: K.pool2x2(a, 0.25), bytes_=1.25 * nb)
s = act(128, 128, 16, 16)
timeit('upsample2x 16->32', lambda: K.upsample2x(s, 1.0), bytes_=1.25 * nb)
timeit('bias_grad', lambda: K.bias_grad(a), bytes_=nb)
gam, bet = torch.ones(10, 128, device=dev), torch.zeros(10, 128, device=dev)
lab = torch.randint(0, 10, (128,), device=dev, dtype=torch.int32)
y, m, i = K.bn_fwd(a, gam, bet, lab, 1e-5, True)
timeit('bn_fwd cond+relu', lambda: K.bn_fwd(a, gam, bet, lab, 1e-5, True), bytes_=3 * nb)
timeit('bn_bwd cond+relu', lambda: K.bn_bwd(b, a, y, gam, bet, lab, m, i, True), bytes_=6 * nb)
print('--- element-wise at [128,128,8,8] bf16 (2 MB)')
a8, b8 = act(128, 128, 8, 8), act(128, 128, 8, 8)
timeit('add 8x8', lambda: K.add(a8, b8))
timeit('act_dropout 8x8', lambda: K.act_dropout(a8, 1.0, 0.5, seed=1, offset=0))
timeit('torch empty_like (allocator)', lambda: torch.empty_like(a8))
p = torch.randn(1055115 + 53, device=dev)[:1055104]
gq, mq, vq = torch.randn_like(p), torch.zeros_like(p), torch.zeros_like(p)
timeit('adam 1.05M params', lambda: K.adam_step(p, gq, mq, vq, 1e-4, 0.0, 0.9, 1e-8), bytes_=28 * p.numel())
