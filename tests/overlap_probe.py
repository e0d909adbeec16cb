"""Timing probe (not a test, results are meaningless numerically): how much of the per-step fake generation is hidden when
its graph replays on a second stream next to the critic-step graph instead of before it."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer

B = 64
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()
gt = GraphedTrainer(tr, (x, y), pregen_steps=1)
s2 = torch.cuda.Stream()
main = torch.cuda.current_stream()


def run(mode, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for k in range(5):
            if mode == 'seq':
                gt.pregen.replay()
                gt.critic.g.replay()
            elif mode == 'critic_only':
                gt.critic.g.replay()
            else:
                s2.wait_stream(main) if k == 0 else None
                with torch.cuda.stream(s2):
                    gt.pregen.replay()
                gt.critic.g.replay()
        main.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for mode in ('critic_only', 'seq', 'overlap', 'critic_only', 'seq', 'overlap'):
    print('%-12s %8.1f us per 5 critic steps' % (mode, run(mode)))
