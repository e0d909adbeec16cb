"""Device time of the 3-channel-side convs: im2col tensor-core path vs the SIMT thin kernels (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K

def timeit(fn, reps=10):
    """device time per call: `reps` calls captured in a CUDA graph (no host launch overhead), replayed 5 times"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3

for (N, H, Cin, Cout, k) in [(192, 32, 3, 128, 3), (64, 32, 3, 128, 3), (192, 16, 3, 128, 1), (128, 32, 128, 3, 3), (64, 32, 128, 3, 3)]:
    g = K.same_geom(N, H, H, Cin, Cout, k, 1)
    x = torch.randn(N, Cin, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(N, Cout, H, H, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(k, k, Cin, Cout, device='cuda') * 0.05).contiguous(); b = torch.zeros(Cout, device='cuda')
    dw = torch.zeros_like(w)
    for thin in (False, True):
        K.config.use_thin_tc = thin
        tf = timeit(lambda: K.conv_fprop(x, w, b, g, w_is_param=True))
        td = timeit(lambda: K.conv_dgrad(dy, w, g, w_is_param=True))
        tw = timeit(lambda: K.conv_wgrad(x, dy, g, tuple(w.shape), accumulate_into=dw))
        print('N=%3d %2dx%-2d %3d->%3d k=%d thin_tc=%d  fprop %7.1f us  dgrad %7.1f us  wgrad %7.1f us' % (N, H, H, Cin, Cout, k, thin, tf, td, tw), flush=True)
K.config.use_thin_tc = True
