"""Data-parallel check on real GPUs (run under torchrun with >= 2 ranks, wrapped in `timeout`; tests/test_dp_gpu.py does that):
   1. one eager critic step and one generator step per rank on DIFFERENT data shards, update through the peer-memory kernel
      (reduce-scatter + Adam + all-gather, csrc/peer.cu): the new parameters equal TF-Adam applied to the MEAN of the ranks'
      local gradients (all-gathered before the update) -- i.e. the N-GPU step is the single-model step on the concatenated
      batch with per-shard batch-norm statistics and per-shard random draws (each shard's gradients are what the single-GPU
      parity tests compare with the oracle) -- and all replicas are bit-identical;
   2. the same through CUDA-graph replays (one graph per step, no NCCL inside);
   3. the NCCL all-reduce fallback gives the same update (to atomics-order noise of the local gradients: same tolerance).
Prints DP_CHECK PASS / FAIL."""
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
T0 = time.time()
fails = []


def say(*a):
    if rank == 0:
        print('[dp_check %.1fs]' % (time.time() - T0), *a, flush=True)


def check(name, cond):
    ok = torch.tensor([1.0 if cond else 0.0], device='cuda')
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    say(name, 'ok' if ok.item() == 1.0 else 'FAILED')
    if ok.item() != 1.0:
        fails.append(name)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


import ctgan_b200.kernels as K
import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer

B = 16


def trainer(peer):
    K.config.peer_update = peer
    np.random.seed(1234)                               # identical initial weights on every rank
    return R.Trainer(device='cuda', seed=1234 + rank, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)


rs = np.random.RandomState(7 + rank)                   # a different data shard per rank
x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()


def expect_update(opt, lr):
    """TF Adam (t = 1) on the mean of the ranks' local gradients, from the parameters before the update."""
    local_g = opt.flat_g.clone()
    gathered = [torch.zeros_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    g = sum(gathered) / world
    m = (1 - opt.beta1) * g
    v = (1 - opt.beta2) * g * g
    lr_t = lr * np.sqrt(1 - opt.beta2) / (1 - opt.beta1)
    return opt.flat_p.clone() - lr_t * m / (v.sqrt() + opt.eps)


results = {}
for peer in (True, False):
    tr = trainer(peer)
    mode = 'peer' if tr.disc_opt.peer is not None else 'nccl'
    say('mode requested', 'peer' if peer else 'nccl', '-> running', mode)
    if peer:
        check('peer buffers available', tr.disc_opt.peer is not None and tr.gen_opt.peer is not None)
    # ---- critic step: local gradients, then the product update
    tr.disc_opt.zero_grad()
    tr.critic_forward_backward(x, y)
    want = expect_update(tr.disc_opt, tr.lr(0))
    w = tr.disc_opt.all_reduce()
    tr.disc_opt.step(tr.lr(0), w)
    tr.rng.end_step()
    torch.cuda.synchronize()
    check('%s: critic update == Adam(mean of the shards\' gradients)  rel %.2e' % (mode, rel(tr.disc_opt.flat_p, want)),
          rel(tr.disc_opt.flat_p, want) < 1e-6 and torch.isfinite(tr.disc_opt.flat_p).all().item())
    # ---- generator step
    tr.gen_opt.zero_grad()
    tr.gen_forward_backward()
    want = expect_update(tr.gen_opt, tr.lr(0))
    w = tr.gen_opt.all_reduce()
    tr.gen_opt.step(tr.lr(0), w)
    tr.rng.end_step()
    torch.cuda.synchronize()
    check('%s: generator update == Adam(mean of the shards\' gradients)' % mode, rel(tr.gen_opt.flat_p, want) < 1e-6)
    for opt in (tr.disc_opt, tr.gen_opt):
        q = opt.flat_p.clone()
        dist.broadcast(q, 0)
        check('%s: replicas bit-identical' % mode, torch.equal(q, opt.flat_p))
    results[mode] = tr.disc_opt.flat_p.clone()
    # ---- CUDA-graph replays
    gt = GraphedTrainer(tr, (x, y), pregen_steps=2)
    say('%s: graphs captured (%d + %d kernels; %s per step)' % (mode, gt.critic_kernels, gt.gen_kernels,
                                                               'one graph' if gt.critic.one_graph else 'two graphs + NCCL'))
    check('%s: one graph per step' % mode, gt.critic.one_graph == (mode == 'peer'))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        if it == 1:
            dist.barrier(); torch.cuda.synchronize(); ev0.record()
        gt.iteration = it
        gt.gen_step()
        gt.begin_iteration(torch.cat([y, y]))
        for k in range(2):
            out = gt.critic_step(x, y)
    ev1.record()
    torch.cuda.synchronize()
    say('%s: 2 iterations (1 gen + 2 critic steps) %.3f ms, cost %.4f' % (mode, ev0.elapsed_time(ev1), float(out[0])))
    for opt in (tr.disc_opt, tr.gen_opt):
        q = opt.flat_p.clone()
        dist.broadcast(q, 0)
        check('%s: replicas bit-identical after graph replays' % mode, torch.equal(q, opt.flat_p) and torch.isfinite(q).all().item())
    del gt, tr
    torch.cuda.empty_cache()
if 'peer' in results and 'nccl' in results:
    # same seeds, same shards: the two exchange paths produce the same first update up to the atomics-order noise of the
    # local gradients (Adam at t = 1 is sign-like: a gradient element that flips sign moves its update by 2 lr)
    d = rel(results['peer'], results['nccl'])
    check('peer update == NCCL update (rel %.2e)' % d, d < 1e-3)
say('DP_CHECK', 'PASS' if not fails else 'FAIL %s' % fails)
dist.destroy_process_group()
