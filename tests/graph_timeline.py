"""Kernel timeline of one replay of each captured graph (critic / generator / pregen) of the ResNet iteration.

torch.profiler (Kineto/CUPTI activity records) sees the kernels of a cudaGraphLaunch with start time, duration and
stream, so this shows what the launch list under ncu cannot: which kernels overlap, where the stream branches idle and
what the critical path of a replay is.  A diagnostic, not a benchmark (numbers under a profiler are never bench values).

    python tests/graph_timeline.py [B] [out_prefix]   ->  <out_prefix>_{critic,gen,pregen}.csv  (name,start_us,dur_us,stream)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer

import ctgan_b200.kernels as K_
if os.environ.get('CTGAN_POOL_CONV_TILES'):
    K_.config.pool_conv_min_tiles = int(os.environ['CTGAN_POOL_CONV_TILES'])
if os.environ.get('CTGAN_POOL_CONV'):
    K_.config.pool_conv_s2d = bool(int(os.environ['CTGAN_POOL_CONV']))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prefix = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/timeline'
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()
gt = GraphedTrainer(tr, (x, y), pregen_steps=5)
gt.begin_iteration(y.repeat(5))


def critic():
    gt._k = 0
    gt.critic_step(x, y)


def trace(fn, tag):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    rows = []
    for e in ev:
        tr_ = e.time_range
        rows.append((e.name, tr_.start, tr_.end - tr_.start, getattr(e, 'device_resource_id', -1) if hasattr(e, 'device_resource_id') else -1))
    rows.sort(key=lambda r: r[1])
    if not rows:
        print(tag, 'no CUDA events captured')
        return
    t0 = rows[0][1]
    with open('%s_%s.csv' % (prefix, tag), 'w') as f:
        f.write('name,start_us,dur_us,stream\n')
        for n, s, d, st in rows:
            f.write('"%s",%.3f,%.3f,%s\n' % (n.replace('"', "'"), s - t0, d, st))
    span = max(r[1] + r[2] for r in rows) - t0
    busy = sum(r[2] for r in rows)
    print('%s: %d kernels, span %.1f us, sum of durations %.1f us (overlap factor %.2f)' % (tag, len(rows), span, busy, busy / span))


trace(critic, 'critic')
trace(gt.gen_step, 'gen')
trace(lambda: gt.begin_iteration(y.repeat(5)), 'pregen')
