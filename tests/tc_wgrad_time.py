"""Device time of the wgrad_tc kernel variants (direct C-ABI launches, CUDA events, rotating buffers > L2)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K
from ctgan_b200 import _lib

def run(N, H, C, k, variant, reps=40):
    g = K.same_geom(N, H, H, C, C, k, 1)
    nset = max(2, int(300e6 // (N * H * H * C * 2 * 2)) + 1)
    xs = [torch.randn(N, C, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nset)]
    dys = [torch.randn(N, C, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nset)]
    dw = torch.zeros(k, k, C, C, device='cuda')
    d = K._desc(g, _lib.BF16, _lib.BF16)
    _lib.lib.ctgan_set_wgrad_variant(variant)
    def launch(i):
        _lib.call('ctgan_conv_wgrad_tc', ctypes.byref(d), K._p(xs[i % nset]), K._p(dys[i % nset]), K._p(dw), K._stream())
    for i in range(5): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): launch(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * N * H * H * C * C * k * k
    print('wgrad N=%4d %2dx%-2d k=%d variant=%d  %8.1f us  %7.1f TFLOP/s' % (N, H, H, k, variant, us, fl / us / 1e6), flush=True)

for (N, H, k) in [(192, 32, 3), (64, 32, 3), (192, 16, 3), (64, 16, 3), (128, 16, 3), (192, 8, 3), (64, 8, 3), (128, 8, 3)]:
    for variant in (2,):
        run(N, H, 128, k, variant)
_lib.lib.ctgan_set_wgrad_variant(2)
