"""2-GPU smoke of the data-parallel path (run under torchrun, wrapped in `timeout`): prints a line per stage."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import torch.distributed as dist

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
def say(*a):
    if rank == 0:
        print('[dp_smoke %.1fs]' % (time.time() - T0), *a, flush=True)
T0 = time.time()
t = torch.ones(4, device='cuda') * (rank + 1)
dist.all_reduce(t); torch.cuda.synchronize()
say('stage1 all_reduce ok', t.tolist())
import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234 + rank, act_dtype=torch.bfloat16, batch_size=64, graph_safe_rng=True)
rs = np.random.RandomState(7 + rank)
x = torch.from_numpy(rs.randint(0, 256, (64, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (64,)).astype('int32')).cuda()
tr.disc_opt.set_device_lr(tr.lr(0)); r = tr.critic_step(x, y, use_device_lr=True); torch.cuda.synchronize()
say('stage2 eager DP critic step ok, cost', float(r['out'][0]))
tr.gen_opt.set_device_lr(tr.lr(0)); tr.gen_step(use_device_lr=True); torch.cuda.synchronize()
say('stage2b eager DP gen step ok')
gt = GraphedTrainer(tr, (x, y)); torch.cuda.synchronize()
say('stage3 graphs captured: critic kernels', gt.critic_kernels, 'gen kernels', gt.gen_kernels)
for i in range(3):
    gt.gen_step(); o = gt.critic_step(x, y)
torch.cuda.synchronize()
say('stage4 replays ok, cost', float(o[0]))
# replicas must hold identical parameters
p = tr.disc_opt.flat_p.clone(); q = p.clone(); dist.broadcast(q, 0); torch.cuda.synchronize()
say('stage5 replicas identical:', bool(torch.equal(p, q)))
ok = torch.tensor([1.0 if torch.equal(p, q) else 0.0], device='cuda'); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
say('DP_SMOKE', 'PASS' if ok.item() == 1.0 else 'FAIL')
dist.destroy_process_group()
