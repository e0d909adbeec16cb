"""BASELINE configs[4]: CT-GAN ResNet critic-step sweep over the per-GPU batch (64 -> 4096), BF16 path.
Prints one JSON line per batch: ms per critic step (CUDA events, after warm-up), algorithmic TFLOP/s
(505.37 GFLOP per critic step at batch 64, linear in the batch) and the fraction of the
measured sustained BF16 peak.  Eager launches up to batch 256 would be host-bound, so every point replays a
CUDA graph of the step."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch

import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer

peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else {'bf16_tflops_sustained': 1400.0}
batches = [int(b) for b in sys.argv[1:]] or [64, 128, 256, 512, 1024, 2048, 4096]
for B in batches:
    np.random.seed(1234)
    tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
    y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()
    gt = GraphedTrainer(tr, (x, y), warmup=2)
    for _ in range(3):
        gt.critic_step(x, y)
    torch.cuda.synchronize()
    reps = 10 if B <= 512 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gt.critic_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gflop = 505.37 * B / 64
    tf = gflop / ms
    print(json.dumps({'per_gpu_batch': B, 'critic_step_ms': round(ms, 3), 'algorithmic_gflop': round(gflop, 1),
                      'tflops': round(tf, 1), 'frac_of_sustained_bf16_peak': round(tf / peaks['bf16_tflops_sustained'], 4),
                      'samples_per_s': round(B / ms * 1e3, 1), 'mem_gb': round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    del gt, tr
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
