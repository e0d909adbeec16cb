"""Device time of the fprop_tc kernel variants (direct C-ABI launches, CUDA events, rotating buffers > L2)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K
from ctgan_b200 import _lib

def run(N, H, C, k, variant, halo, reps=40):
    g = K.same_geom(N, H, H, C, C, k, 1)
    nset = max(2, int(300e6 // (N * H * H * C * 2 * 2)) + 1)
    xs = [torch.randn(N, C, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nset)]
    ys = [torch.empty_like(xs[0]) for _ in range(nset)]
    w = (torch.randn(k, k, C, C, device='cuda') * 0.03).contiguous(); b = torch.zeros(C, device='cuda')
    wp = K.pack_filter(w, 0)
    d = K._desc(g, _lib.BF16, _lib.BF16)
    _lib.lib.ctgan_set_fprop_variant(variant); _lib.lib.ctgan_set_fprop_halo(halo)
    def launch(i):
        _lib.call('ctgan_conv_fprop_tc', ctypes.byref(d), K._p(xs[i % nset]), K._p(wp), K._p(b), None, K._p(ys[i % nset]), 0, K._stream())
    for i in range(5): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): launch(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * N * H * H * C * C * k * k
    print('N=%4d %2dx%-2d k=%d variant=%d halo=%d  %8.1f us  %7.1f TFLOP/s' % (N, H, H, k, variant, halo, us, fl / us / 1e6), flush=True)

for (N, H, k) in [(128, 32, 3), (192, 32, 3), (64, 32, 3), (512, 32, 3), (192, 16, 3), (512, 16, 3)]:
    for variant, halo in [(2, 1), (3, 1), (4, 1)]:
        run(N, H, 128, k, variant, halo)
_lib.lib.ctgan_set_fprop_variant(4); _lib.lib.ctgan_set_fprop_halo(1)
