"""Replay time of the captured critic / generator graphs of a DCGAN script (CT_gan_cifar.py or CT_gan_mnist.py) with the
stride-2 5x5 layers on the SIMT kernels vs on the tensor cores (space-to-depth route), CUDA events, inputs resident.

    python tests/dcgan_times.py cifar|mnist [batch]"""
import importlib
import json
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import ctgan_b200.kernels as K
from ctgan_b200.graphs import GraphedTrainer

script = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
mod = importlib.import_module('ctgan_b200.gan_' + script)
B = int(sys.argv[2]) if len(sys.argv) > 2 else mod.BATCH_SIZE
GFLOP = {'cifar': (191.51, 68.95), 'mnist': (32.80, 11.83)}[script]      # critic step, generator step


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for s2d in (False, True):
    K.config.use_s2d = s2d
    np.random.seed(1234)
    tr = mod.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
    rs = np.random.RandomState(0)
    if script == 'cifar':
        x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
    else:
        x = torch.from_numpy(rs.random_sample((B, 784)).astype('float32')).cuda()
    gt = GraphedTrainer(tr, (x,))
    tc = t(lambda: gt.critic_step(x))
    tg = t(gt.gen_step)
    out = gt.critic_step(x).tolist()
    it_us = 5 * tc + tg
    scale = B / mod.BATCH_SIZE
    print(json.dumps({'script': script, 'batch': B, 's2d': s2d, 'critic_us': tc, 'critic_kernels': gt.critic_kernels, 'gen_us': tg,
                      'gen_kernels': gt.gen_kernels, 'iterations_per_s': 1e6 / it_us,
                      'tflops': scale * (5 * GFLOP[0] + GFLOP[1]) * 1e-3 / (it_us * 1e-6), 'losses': out[:4]}), flush=True)
    del gt, tr
    torch.cuda.synchronize()
