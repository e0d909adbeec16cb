"""bench.py -- CT-GAN training throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-graph]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

The metric counts training ITERATIONS of CT_gan_cifar_resnet.py -- 1 generator step + 5 critic
steps (TG/CT_gan_cifar_resnet.py:393-404) -- at batch 64 per GPU, DIM 128, BF16 tensor-core
path, synthetic CIFAR-shaped data, random-init weights.  One bench "step" (--steps K) is a block of
ITERS_PER_STEP = 16 such iterations, so that the driver's K = 20 gives a timed region of ~2 s (320
iterations) instead of 0.14 s -- long enough for sustained clocks and a few dozen nvidia-smi samples.
`value` is iterations/s whatever the block size.  One JSON line is printed by rank 0.

  value     iterations/s with the real batches already resident in HBM
  e2e       the same loop through the public API with HOST batches: every critic step copies
            its int32 [64,3072] batch + labels from pinned host memory and reads the 8 loss
            scalars back; every generator step reads its loss back
  roofline  the kernel family with the largest share of the step, timed alone with CUDA events over buffers
            that exceed L2; roofline_top3 = the three largest families (shares: profiles/r02_kernel_shares.json)
  other_workloads
            the same measurement (value + e2e) for CT_gan_cifar.py and CT_gan_mnist.py at the same N
  cpu_baseline / --impl reference
            the CPU transcription of the reference graph (oracle/, PyTorch-CPU fp32, all host
            threads) -- TF 1.2.1 cannot be installed here
  --workload cifar|mnist|64x64|lsun128
            the same line for the DCGAN scripts CT_gan_cifar.py / CT_gan_mnist.py (BASELINE configs[1] / [0]: parity
            configurations, not the headline); default = resnet, BASELINE.json's metric
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'CT-GAN train iters/sec (CIFAR ResNet)'
UNIT = 'iterations/s (1 gen + 5 critic steps of batch 64 per GPU; aggregate over GPUs)'
BATCH = 64
N_CRITIC = 5
ITERS_PER_STEP = 16           # training iterations per bench step (see the module docstring)
ITER_GFLOP = 2990.34          # algorithmic FLOPs of one iteration of the REFERENCE graph
# executed: the 1x1 shortcut convs run on the low-resolution side of their resampling (gan_cifar_resnet.COMMUTE_1X1, exact):
# -0.75 * 2*256*128*128 FLOP per critic pass-image (Discriminator.2.Shortcut) over 5*(3*192 + 4*64) + 2*128 image-traversals,
# -0.75 * 2*(64+256+1024)*128*128 per generator pass-image (Generator.{1,2,3}.Shortcut) over 320 + 3*128 image-traversals
ITER_GFLOP_EXECUTED = ITER_GFLOP - (0.75 * 2 * 256 * 128 * 128 * (5 * (3 * 192 + 4 * 64) + 2 * 128)
                                    + 0.75 * 2 * (64 + 256 + 1024) * 128 * 128 * (320 + 3 * 128)) * 1e-9
# ... and ConvMeanPool(3x3) as one stride-2 4x4 conv (kernels.config.pool_conv_s2d, exact in real arithmetic): 16 taps at a
# quarter of the positions instead of 9 at all of them = -5/9 of 2*1024*1152*128 FLOP per image-traversal of
# Discriminator.1.Conv2 (every pass: same traversal count as above) and of 2*256*1152*128 for Discriminator.2.Conv2 where the
# route's size policy takes it (the 192-image stacked pass: forward, dgrad, wgrad of 5 critic steps)
POOL_CONV_GFLOP_SAVED = 5.0 / 9.0 * 2 * 1152 * 128 * (1024 * (5 * (3 * 192 + 4 * 64) + 2 * 128) + 256 * (5 * 3 * 192)) * 1e-9


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], burst=p['bf16_tflops'], sustained=p['bf16_tflops_sustained'], src='measured')
    except Exception:
        return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src='fallback')


# ----------------------------------------------------------------------------- CPU reference arm
REF_DEAD_WORK = ('executes the reference graph as written, including the metrics-only clean critic pass and the fake half of '
                 'the second stochastic pass (TG/CT_gan_cifar_resnet.py:228,238-242: +173.66 GFLOP per critic step) that the GPU '
                 'arm does not execute because nothing consumes them')


def _cpu_model():
    import numpy as np
    import torch
    from oracle import ct_gan_cifar_resnet as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    np.random.seed(1234)
    m = R.Model(dtype=torch.float32, batch_size=BATCH).build()
    rs = np.random.RandomState(1234)
    x = torch.from_numpy(rs.randint(0, 256, (BATCH, 3072)).astype('int32'))
    y = torch.from_numpy(rs.randint(0, 10, (BATCH,)).astype('int32'))
    return m, x, y, torch.get_num_threads()


def time_cpu_reference(max_seconds=150.0, want_steps=1):
    """cpu_baseline of the default line: the oracle (CPU transcription of CT_gan_cifar_resnet.py, fp32, all host threads)
    on a bounded sample: n critic steps + n generator steps at batch 64; an iteration = 5 critic + 1 gen."""
    from oracle.rand import SeededRandom
    m, x, y, cores = _cpu_model()
    t0 = time.time()
    m.critic_step(SeededRandom(1), x, y, iteration=0)            # warm-up (also sizes the sample)
    t_warm = time.time() - t0
    n = max(1, min(want_steps, int(max_seconds / max(2.2 * t_warm, 1e-3))))
    tc = tg = 0.0
    for i in range(n):
        t0 = time.time(); m.critic_step(SeededRandom(10 + i), x, y, iteration=i); tc += time.time() - t0
        t0 = time.time(); m.gen_step(SeededRandom(100 + i), iteration=i); tg += time.time() - t0
    tc, tg = tc / n, tg / n
    it_s = 1.0 / (N_CRITIC * tc + tg)
    return dict(value=it_s, unit=UNIT, cores=cores, kind='port',
                sample='%d critic + %d generator steps at batch 64 (fp32 oracle, PyTorch-CPU); iteration = 5*critic + 1*gen '
                       '= %.2f s; %s' % (n, n, N_CRITIC * tc + tg, REF_DEAD_WORK)), (N_CRITIC * tc + tg)


def run_reference(args):
    """`--impl reference`: the reference's own CPU path (TF 1.2.1 is not installable -> its PyTorch-CPU transcription, oracle/)
    on all host threads.  W warm-up + EXACTLY K timed steps; a step is one full training iteration (1 generator + 5 critic
    steps) when W + K of them fit in ~4 minutes, else a bounded sample of it (1 generator + 1 critic step; the line says so
    and extrapolates value = 1 / (5 * critic + generator))."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if getattr(args, 'workload', 'resnet') != 'resnet':
        script = args.workload
        cb = time_cpu_dcgan(script, max_seconds=60.0)
        unit = 'iterations/s (1 gen + 5 critic steps of batch %d per GPU; aggregate over GPUs)' % DCGAN[script][2]
        cb['unit'] = unit
        print(json.dumps({
            'impl': 'reference', 'metric': 'CT-GAN train iters/sec (%s)' % ('CIFAR DCGAN' if script == 'cifar' else 'MNIST DCGAN'),
            'value': cb['value'], 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 / cb['value'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': DCGAN[script][6] + '; CPU transcription of the reference graph on host cores'},
            'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}), flush=True)
        return
    from oracle.rand import SeededRandom
    m, x, y, cores = _cpu_model()
    K_, W_ = max(1, args.steps), max(0, args.warmup)
    t0 = time.time(); m.critic_step(SeededRandom(1), x, y, iteration=0); tc0 = time.time() - t0
    t0 = time.time(); m.gen_step(SeededRandom(2), iteration=0); tg0 = time.time() - t0
    full = (K_ + W_) * (N_CRITIC * tc0 + tg0) <= 240.0
    n_c = N_CRITIC if full else 1

    def step(i):
        tc = tg = 0.0
        t0 = time.time(); m.gen_step(SeededRandom(100 + i), iteration=i); tg = time.time() - t0
        for k in range(n_c):
            t0 = time.time(); m.critic_step(SeededRandom(1000 + 10 * i + k), x, y, iteration=i); tc += time.time() - t0
        return tc / n_c, tg
    for i in range(W_ if full else min(W_, 1)):
        step(i)
    t_begin = time.time()
    tcs, tgs = [], []
    for i in range(K_):
        tc, tg = step(W_ + i)
        tcs.append(tc); tgs.append(tg)
    wall = time.time() - t_begin
    tc, tg = sum(tcs) / K_, sum(tgs) / K_
    sec_iter = N_CRITIC * tc + tg
    value = 1.0 / sec_iter
    sample = ('%d timed steps, each %s at batch 64 (fp32 PyTorch-CPU transcription of the reference graph, %d threads); '
              'critic step %.3f s, generator step %.3f s, iteration = 5*critic + 1*gen = %.2f s; %s'
              % (K_, 'one full iteration (1 generator + 5 critic steps)' if full else
                 'a bounded sample of one iteration (1 generator + 1 critic step; value extrapolates to 1 + 5)', cores, tc, tg,
                 sec_iter, REF_DEAD_WORK))
    cb = dict(value=value, unit=UNIT, cores=cores, kind='port', sample=sample)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': K_, 'warmup': W_, 'ms_per_step': wall / K_ * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'CT_gan_cifar_resnet.py critic+generator iteration, batch 64, DIM 128 (BASELINE configs[2]); '
                               'CPU transcription of the reference graph on host cores',
                   'step': 'one full training iteration' if full else 'bounded sample: 1 generator + 1 critic step',
                   'dead_work': REF_DEAD_WORK},
        'cpu_baseline': cb,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- our arm
def _time_launches(torch, launch, n_sets, reps):
    for i in range(n_sets):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps            # ms per launch


def _bf16_act(torch, shape):
    return torch.randn(*shape, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def roofline_kernels(torch, peaks):
    """The three kernel families with the largest share of the ResNet iteration (shares: profiles/r02_kernel_shares.json, from
    the ncu launch list of the same build), each timed alone with CUDA events on the launching stream, rotating over buffer
    sets that exceed the 126 MB L2.  `roofline` is the family with the LARGEST share:
      * conv_fprop_tc_pair_kernel: Discriminator.1.Conv2 on the stacked 192-image pass (32x32, 3x3, 128->128), the single
        largest GEMM of the step (fprop and, same kernel with the flipped filter pack, dgrad);
      * the sub-wave 3x3 family: the 8x8 layers (Discriminator.3/.4, fprop / dgrad / double-backward fprop) as the 64-image
        gradient-penalty pass launches them -- 32 tiles on 148 SMs, latency-bound (conv_fprop_tc_lean_kernel<1> with [h][n][w]
        halo boxes inside the two-branch critic step, the cluster split-K kernel elsewhere);
      * conv_wgrad_tc_multi_kernel: every tensor-core filter gradient of one critic step (stacked pass 192 images +
        gradient-penalty pass 64 images; 3x3 at 32x32, 16x16 x2, 8x8 x4 and the 1x1 shortcut, each) as ONE launch."""
    import ctypes
    import ctgan_b200.kernels as K
    from ctgan_b200 import _lib
    C = 128
    peak_src = ('MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)' if peaks['src'] == 'measured' else 'fallback 1590')
    shares = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_kernel_shares.json')) as f:
            shares = json.load(f)
    except Exception:
        pass
    w = (torch.randn(3, 3, C, C, device='cuda') * 0.03).contiguous()
    b = torch.zeros(C, device='cuda')
    wp = K.pack_filter(w, 0)
    out = []

    def entry(key, kernel, flops, ms, traffic=None, extra=None):
        ach = flops / (ms * 1e-3) / 1e12
        e = {'bound': 'tensor', 'kernel': kernel, 'share_of_step': shares.get(key), 'achieved': ach, 'peak': peaks['burst'],
             'peak_source': peak_src, 'unit': 'TFLOP/s', 'frac': ach / peaks['burst'], 'traffic': traffic,
             'flops_per_launch': flops, 'us_per_launch': ms * 1e3}
        if extra:
            e.update(extra)
        out.append(e)

    # (1) pair kernel, 192 x 32 x 32
    N, H = 3 * BATCH, 32
    g = K.same_geom(N, H, H, C, C, 3, 1)
    xs = [_bf16_act(torch, (N, C, H, H)) for _ in range(6)]
    ys = [torch.empty_like(xs[0]) for _ in range(6)]
    d = K._desc(g, _lib.BF16, _lib.BF16)
    ms = _time_launches(torch, lambda i: _lib.call('ctgan_conv_fprop_tc', ctypes.byref(d), K._p(xs[i % 6]), K._p(wp), K._p(b), None,
                                                  K._p(ys[i % 6]), 0, K._stream()), 6, 30)
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            traffic = json.load(f).get('conv_fprop_tc_192x32x32x128_bytes')
    except Exception:
        pass
    entry('conv_fprop_tc_pair_kernel', 'conv_fprop_tc_pair_kernel (3x3, 128->128, %dx32x32)' % N, 2.0 * N * H * H * C * C * 9, ms, traffic)
    del xs, ys

    # (2) sub-wave 3x3 conv, 64 x 8 x 8 (80 buffer sets of 1 MB in + 1 MB out)
    N, H = BATCH, 8
    g = K.same_geom(N, H, H, C, C, 3, 1)
    xs = [_bf16_act(torch, (N, C, H, H)) for _ in range(80)]
    ys = [torch.empty_like(xs[0]) for _ in range(80)]
    d8 = K._desc(g, _lib.BF16, _lib.BF16)
    launch8 = lambda i: _lib.call('ctgan_conv_fprop_tc', ctypes.byref(d8), K._p(xs[i % 80]), K._p(wp), K._p(b), None,
                                  K._p(ys[i % 80]), 0, K._stream())
    with K.splitk(False):                     # one CTA per tile: what the two-branch critic step launches
        ms = _time_launches(torch, launch8, 80, 160)
    entry('sub_wave_3x3', 'conv_fprop_tc_lean_kernel<1> (3x3, 128->128, %dx8x8: 32 tiles on 148 SMs, one CTA per tile of two images; '
          'latency-bound: 1.2 GFLOP per launch)' % N, 2.0 * N * H * H * C * C * 9, ms)
    with K.splitk(True):                      # cluster split-K (4 CTAs per tile): what the generator / DCGAN steps launch
        ms_sk = _time_launches(torch, launch8, 80, 160)
    out[-1]['split_k_cluster_us_per_launch'] = ms_sk * 1e3
    del xs, ys

    # (3) all filter gradients of one critic step in one launch
    jobs, flops = [], 0.0
    for n in (3 * BATCH, BATCH):
        for (h, k, reps) in ((32, 3, 1), (16, 3, 2), (8, 3, 4), (8, 1, 1)):
            for _ in range(reps):
                gj = K.same_geom(n, h, h, C, C, k, 1)
                jobs.append((_bf16_act(torch, (n, C, h, h)), _bf16_act(torch, (n, C, h, h)), gj,
                             torch.zeros(k, k, C, C, device='cuda')))
                flops += 2.0 * n * h * h * C * C * k * k

    def launch_multi(i):
        for xj, dyj, gj, dwj in jobs:
            K.conv_wgrad(xj, dyj, gj, tuple(dwj.shape), accumulate_into=dwj, defer=True)
        K.flush_wgrads()
    # replayed from a CUDA graph: queueing 16 jobs and building their tensor maps takes the host longer than the kernel runs
    launch_multi(0)
    torch.cuda.synchronize()
    how = 'CUDA graph replay'
    try:
        side = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                launch_multi(0)
        torch.cuda.current_stream().wait_stream(side)
        ms = _time_launches(torch, lambda i: gr.replay(), 2, 20)
    except Exception as e:                                      # never lose the bench line to the stand-alone timing
        sys.stderr.write('bench.py: graph capture of the filter-gradient launch failed (%s); timing eager launches\n' % e)
        torch.cuda.synchronize()
        how = 'eager launches (host-bound)'
        ms = _time_launches(torch, launch_multi, 2, 20)
    entry('conv_wgrad_tc_multi_kernel', 'conv_wgrad_tc_multi_kernel (the %d filter gradients of one critic step: 192 + 64 images)' % len(jobs),
          flops, ms, None, {'operand_bytes': int(sum(2 * xj.numel() * 2 for xj, _, _, _ in jobs)), 'timed_as': how})
    out.sort(key=lambda e: -(e['share_of_step'] or 0.0))
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    import ctgan_b200.gan_cifar_resnet as R
    from ctgan_b200 import _lib
    from ctgan_b200.graphs import GraphedTrainer
    np.random.seed(1234)                      # identical initial weights on every rank (reference init formulas)
    tr = R.Trainer(device=dev, seed=1234 + rank, act_dtype=torch.bfloat16, batch_size=BATCH,
                   graph_safe_rng=not args.no_graph)

    # synthetic data: a pool of batches, pinned on the host (e2e) and resident on the device (value)
    pool = 8
    rs = np.random.RandomState(1234 + rank)
    host_x = torch.from_numpy(rs.randint(0, 256, (pool, BATCH, 3072)).astype('int32')).pin_memory()
    host_y = torch.from_numpy(rs.randint(0, 10, (pool, BATCH)).astype('int32')).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    host_out = torch.zeros(N_CRITIC + 1, 8, dtype=torch.float32).pin_memory()

    if args.no_graph:
        gt = None
    else:
        gt = GraphedTrainer(tr, (dev_x[0], dev_y[0]), pregen_steps=0 if args.no_pregen else N_CRITIC)
    state = {'it': 0, 'b': 0}

    def iteration(e2e):
        it = state['it']
        src_x, src_y = (host_x, host_y) if e2e else (dev_x, dev_y)
        if gt is not None:
            gt.iteration = it
            if it > 0 or True:                # the reference skips the very first generator step (:396); timed loops never hit it==0
                g = gt.gen_step()
                if e2e:
                    host_out[N_CRITIC, 0:1].copy_(g.reshape(-1)[:1], non_blocking=True)
            if gt.pregen_steps:                # fakes of the 5 critic steps in one generator forward (their labels first)
                idx = [(state['b'] + 1 + i) % pool for i in range(N_CRITIC)]
                gt.begin_iteration(src_y[idx])
            for i in range(N_CRITIC):
                b = state['b'] = (state['b'] + 1) % pool
                out = gt.critic_step(src_x[b], src_y[b])
                if e2e:
                    host_out[i].copy_(out, non_blocking=True)
        else:
            res = tr.gen_step(iteration=it)
            if e2e:
                host_out[N_CRITIC, 0:1].copy_(res['cost'].reshape(-1)[:1], non_blocking=True)
            for i in range(N_CRITIC):
                b = state['b'] = (state['b'] + 1) % pool
                x = src_x[b].to(dev, non_blocking=True) if e2e else src_x[b]
                y = src_y[b].to(dev, non_blocking=True) if e2e else src_y[b]
                res = tr.critic_step(x, y, iteration=it)
                if e2e:
                    host_out[i].copy_(res['out'], non_blocking=True)
        state['it'] = it + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    IPS = ITERS_PER_STEP

    def timed(e2e, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps * IPS):
            iteration(e2e)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(3, args.warmup) * IPS):
        iteration(False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    k0 = _lib.lib.ctgan_kernel_launches()
    ms = timed(False, args.steps)
    eager_launches = _lib.lib.ctgan_kernel_launches() - k0
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        iteration(True)
    ms_e2e = timed(True, args.steps)
    n_iters = args.steps * IPS
    if gt is not None:
        launches = n_iters * (gt.gen_kernels + N_CRITIC * gt.critic_kernels + gt.pregen_kernels)
    else:
        launches = eager_launches
    pregen_on = gt is not None and bool(gt.pregen_steps)

    # the DCGAN scripts at the same N (north_star: MNIST and CIFAR throughput at 1 / 2 / 4 / 8 GPUs): free the ResNet model
    # first -- tflib keeps ONE model per process (module-level parameter dict, like the reference)
    others = {}
    if not args.no_other_workloads and not args.no_graph:
        del gt, tr
        torch.cuda.empty_cache()
        for script in ('cifar', 'mnist'):
            o = measure_dcgan(args, script, rank, local_rank, world, dev, with_roofline=False)
            if rank == 0:
                others[o['metric']] = {k: o[k] for k in ('value', 'unit', 'ms_per_step', 'e2e', 'gpu_launches', 'step_tensor_utilisation')}
                others[o['metric']]['config'] = o['config']['workload']
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    it_s = world * n_iters / (ms * 1e-3)
    it_s_e2e = world * n_iters / (ms_e2e * 1e-3)
    top3 = roofline_kernels(torch, peaks)
    roof = top3[0]
    import ctgan_b200.kernels as _K
    executed = ITER_GFLOP_EXECUTED - (POOL_CONV_GFLOP_SAVED if _K.config.pool_conv_s2d else 0.0)
    step_tflops = executed * 1e-3 * (n_iters / (ms * 1e-3))              # per GPU, FLOPs actually required by the executed math
    ref_tflops = ITER_GFLOP * 1e-3 * (n_iters / (ms * 1e-3))             # ... the reference graph's FLOPs per second (what a
    #                                                                      non-restructured implementation would have to sustain)
    line = {
        'metric': METRIC, 'value': it_s, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {
            'workload': 'CT_gan_cifar_resnet.py iteration = 1 generator step (2x64 fakes) + 5 critic steps (64 real + 64 fake, '
                        '2 stochastic passes + GP double-backward), batch 64 per GPU, DIM_G=DIM_D=128, conditional+ACGAN '
                        '(BASELINE configs[2]; configs[3] when n_gpus>1: batch-sharded data parallel, NCCL all-reduce of flat grads)',
            'step': '%d training iterations (so that --steps 20 times %d iterations, ~2 s)' % (IPS, 20 * IPS),
            'per_gpu_batch': BATCH, 'critic_iters': N_CRITIC, 'cuda_graphs': not args.no_graph,
            'l2_policy': 'working set per iteration (~1.5 GB of activations, 8 rotating input batches) exceeds the 126 MB L2; no explicit flush',
            'precision': 'BF16 operands / fp32 accumulate (tcgen05), fp32 master weights and optimizer',
        },
        'clocks': clocks,
        'e2e': {'value': it_s_e2e, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': IPS * (N_CRITIC * (BATCH * 3072 * 4 + BATCH * 4 + 4) + 4 +
                                             (N_CRITIC * BATCH * 4 if pregen_on else 0)),
                'd2h_bytes_per_step': IPS * (N_CRITIC * 32 + 4)},
        'gpu_launches': int(launches),
        'roofline': roof,
        'roofline_top3': top3,
        'other_workloads': others,
        'step_tensor_utilisation': {'algorithmic_gflop_per_iteration': ITER_GFLOP, 'executed_gflop_per_iteration': executed,
                                    'achieved_tflops_per_gpu': step_tflops,
                                    'frac_of_sustained_peak': step_tflops / peaks['sustained'], 'peak': peaks['sustained'],
                                    'reference_graph_tflops_per_gpu': ref_tflops,
                                    'reference_graph_frac_of_sustained_peak': ref_tflops / peaks['sustained'],
                                    'note': 'executed = reference graph minus the exact algebraic restructurings (1x1 shortcut convs on '
                                            'the low-resolution side; ConvMeanPool(3x3) as one 4x4 / stride-2 conv)'},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = time_cpu_reference(max_seconds=25.0, want_steps=1)
        line['cpu_baseline'] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- DCGAN workloads (BASELINE configs[0], [1])
DCGAN = {
    # script: (module, oracle module, batch, critic GFLOP, generator GFLOP, input row, config name)
    'cifar': ('ctgan_b200.gan_cifar', 'oracle.ct_gan_cifar', 64, 191.51, 68.95, 3072,
              'CT_gan_cifar.py DCGAN critic/generator, 32x32x3 synthetic batch 64, critic_iters=5 (BASELINE configs[1])'),
    'mnist': ('ctgan_b200.gan_mnist', 'oracle.ct_gan_mnist', 50, 32.80, 11.83, 784,
              'CT_gan_mnist.py DCGAN-style critic/generator, 28x28x1 synthetic batch 50, dropout CT term + GP (BASELINE configs[0])'),
    # SURVEY 8(f) row N4; GFLOP = executed conv / linear GEMM work counted per launcher call (2 FLOP / MAC), not a SURVEY figure
    '64x64': ('ctgan_b200.gan_64x64', 'oracle.ct_gan_64x64', 64, 3445.21, 1120.42, 12288,
              'CT_gan_64x64.py GoodGenerator / GoodDiscriminator (ResNet, layer norm in the critic), MODE wgan-ct, 64x64x3 synthetic batch 64, '
              'DIM 64, critic_iters=5 (SURVEY 8(f) row N4)'),
    # the same row's second script; GFLOP None = counted at run time (_count_gemm_gflop) from the launcher calls of one eager iteration
    'lsun128': ('ctgan_b200.gan_lsun128', 'oracle.wgan_lsun128', 64, None, None, 49152,
                'LSUN_bedrooms/wgan_LSUN_Bedrooms128.py ResnetGenerator / ResnetDiscriminator (layer norm in the critic), 128x128x3 synthetic '
                'batch 64, critic 128..1024 channels, critic_iters=5 (SURVEY 8(f) row N4)'),
}
DCGAN_ITERS_PER_STEP = {'64x64': 2, 'lsun128': 1}         # 64x64: ~46 ms per iteration; lsun128: ~0.4 s


def _dcgan_batches(np, script, pool, B, seed):
    rs = np.random.RandomState(seed)
    if script == 'cifar':
        return rs.randint(0, 256, (pool, B, 3072)).astype('int32')
    if script == '64x64':
        return rs.randint(0, 256, (pool, B, 3, 64, 64)).astype('int32')
    if script == 'lsun128':
        return rs.randint(0, 256, (pool, B, 3, 128, 128)).astype('int32')
    return rs.random_sample((pool, B, 784)).astype('float32')


def time_cpu_dcgan(script, max_seconds=25.0):
    import importlib
    import numpy as np
    import torch
    from oracle.rand import SeededRandom
    _, ora, B = DCGAN[script][:3]
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(1234)
    m = importlib.import_module(ora).Model(dtype=torch.float32, batch_size=B).build()
    x = torch.from_numpy(_dcgan_batches(np, script, 1, B, 1234)[0])
    t0 = time.time(); m.critic_step(SeededRandom(1), x, iteration=0); t_warm = time.time() - t0
    n = max(1, min(5, int(max_seconds / max(2.2 * t_warm, 1e-3))))
    tc = tg = 0.0
    for i in range(n):
        t0 = time.time(); m.critic_step(SeededRandom(10 + i), x, iteration=i); tc += time.time() - t0
        t0 = time.time(); m.gen_step(SeededRandom(100 + i), iteration=i); tg += time.time() - t0
    tc, tg = tc / n, tg / n
    return dict(value=1.0 / (N_CRITIC * tc + tg), unit='iterations/s', cores=torch.get_num_threads(), kind='port',
                sample='%d critic + %d generator steps at batch %d (fp32 oracle, PyTorch-CPU); iteration = 5*critic + 1*gen = %.2f s'
                       % (n, n, B, N_CRITIC * tc + tg))


def roofline_dcgan_kernel(torch, peaks, script, B):
    """The dominant launch of the DCGAN critic step: the second stride-2 5x5 conv on the stacked pass (3B images), run as a 3x3
    tcgen05 conv over the space-to-depth image.  ALGORITHMIC FLOPs = the 5x5/2 conv's (25 taps); the kernel executes 36/25 of them."""
    import ctypes
    import ctgan_b200.kernels as K
    from ctgan_b200 import _lib
    N, H, Cin, Cout = (3 * B, 16, 128, 256) if script == 'cifar' else (3 * B, 14, 64, 128)
    g = K.same_geom(N, H, H, Cin, Cout, 5, 2)
    g3 = K.s2d_geom(g)
    sets = [K.space_to_depth(torch.randn(N, Cin, H, H, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last), g)
            for _ in range(8)]
    w = (torch.randn(5, 5, Cin, Cout, device='cuda') * 0.03).contiguous()
    wp = K.pack_filter_s2d(w, g, 0)
    b = torch.zeros(Cout, device='cuda')
    ys = [torch.empty((N, Cout, g.Ho, g.Wo), dtype=torch.bfloat16, device='cuda').contiguous(memory_format=torch.channels_last) for _ in range(8)]
    d = K._desc(g3, _lib.BF16, _lib.BF16)
    def launch(i):
        _lib.call('ctgan_conv_fprop_tc', ctypes.byref(d), K._p(sets[i % 8]), K._p(wp), K._p(b), None, K._p(ys[i % 8]), 0, K._stream())
    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * N * g.Ho * g.Wo * Cin * Cout * 25
    achieved = flops / (ms * 1e-3) / 1e12
    return {'bound': 'tensor', 'kernel': 'conv_fprop_tc_lean_kernel<0> as the 5x5/2 conv %d->%d on %dx%dx%d (3x3 over the space-to-depth image)'
                                         % (Cin, Cout, N, H, H),
            'achieved': achieved, 'peak': peaks['burst'], 'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)'
            if peaks['src'] == 'measured' else 'fallback 1590', 'unit': 'TFLOP/s', 'frac': achieved / peaks['burst'], 'traffic': None,
            'flops_per_launch': flops, 'executed_flops_per_launch': flops * 36 / 25, 'us_per_launch': ms * 1e3}


def _count_gemm_gflop(tr, x):
    """Executed conv / linear GEMM work (2 FLOP per MAC, from the geometry of every launcher call) of one eager critic step and
    one eager generator step: (critic GFLOP, generator GFLOP).  For workloads SURVEY.md gives no figure for."""
    import ctgan_b200.kernels as K
    tot = {'f': 0.0}
    names = ('conv_fprop', 'conv_dgrad', 'conv_wgrad', 'conv_fprop_actdrop')
    saved = {n: getattr(K, n) for n in names}

    def wrap(fn):
        def w(*a, **k):
            g = next(v for v in a if isinstance(v, K.ConvGeom))
            tot['f'] += 2.0 * g.N * g.Ho * g.Wo * g.Cin * g.Cout * g.kh * g.kw
            return fn(*a, **k)
        return w
    for n in names:
        setattr(K, n, wrap(saved[n]))
    try:
        tot['f'] = 0.0; tr.critic_step(x); c = tot['f']
        tot['f'] = 0.0; tr.gen_step(); g = tot['f']
    finally:
        for n in names:
            setattr(K, n, saved[n])
    return c / 1e9, g / 1e9


def measure_dcgan(args, script, rank, local_rank, world, dev, with_roofline=True):
    """One DCGAN workload (CT_gan_cifar.py / CT_gan_mnist.py) measured like the headline: W warm-up + K timed steps of
    ITERS_PER_STEP iterations, device-resident (`value`) and through host batches (`e2e`); returns the JSON line (rank 0)."""
    import importlib
    import numpy as np
    import torch
    import torch.distributed as dist
    modname, _, B, gf_c, gf_g, row, cfg = DCGAN[script]
    from ctgan_b200.graphs import GraphedTrainer
    mod = importlib.import_module(modname)
    np.random.seed(1234)
    tr = mod.Trainer(device=dev, seed=1234 + rank, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
    pool = 8
    host_x = torch.from_numpy(_dcgan_batches(np, script, pool, B, 1234 + rank)).pin_memory()
    dev_x = host_x.to(dev)
    host_out = torch.zeros(N_CRITIC + 1, 8, dtype=torch.float32).pin_memory()
    if gf_c is None:
        gf_c, gf_g = _count_gemm_gflop(tr, dev_x[0])
    gt = GraphedTrainer(tr, (dev_x[0],))
    state = {'b': 0}
    IPS = DCGAN_ITERS_PER_STEP.get(script, ITERS_PER_STEP)

    def iteration(e2e):
        src = host_x if e2e else dev_x
        g = gt.gen_step()
        if e2e:
            host_out[N_CRITIC, 0:1].copy_(g.reshape(-1)[:1], non_blocking=True)
        for i in range(N_CRITIC):
            b = state['b'] = (state['b'] + 1) % pool
            out = gt.critic_step(src[b])
            if e2e:
                host_out[i].copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps * IPS):
            iteration(e2e)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(3, args.warmup) * IPS):
        iteration(False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(False, args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        iteration(True)
    ms_e2e = timed(True, args.steps)
    n_iters = args.steps * IPS
    launches = int(n_iters * (gt.gen_kernels + N_CRITIC * gt.critic_kernels))
    del gt, tr
    if rank != 0:
        return None
    peaks = load_peaks()
    unit = 'iterations/s (1 gen + 5 critic steps of batch %d per GPU; aggregate over GPUs)' % B
    it_s, it_s_e2e = world * n_iters / (ms * 1e-3), world * n_iters / (ms_e2e * 1e-3)
    gflop = N_CRITIC * gf_c + gf_g
    step_tflops = gflop * 1e-3 * (n_iters / (ms * 1e-3))
    line = {
        'metric': 'CT-GAN train iters/sec (%s)' % {'cifar': 'CIFAR DCGAN', 'mnist': 'MNIST DCGAN', '64x64': 'ImageNet 64x64 ResNet', 'lsun128': 'LSUN 128x128 ResNet'}[script],
        'value': it_s, 'unit': unit,
        'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': cfg, 'step': '%d training iterations' % IPS, 'per_gpu_batch': B, 'critic_iters': N_CRITIC, 'cuda_graphs': True,
                   'l2_policy': '8 rotating input batches; activations of the stacked critic pass exceed L2 only for cifar; no explicit flush',
                   'precision': ('BF16 operands / fp32 accumulate (tcgen05), layer-norm statistics in fp32, fp32 master weights and optimizer'
                                 if script in ('64x64', 'lsun128') else
                                 'BF16 operands / fp32 accumulate (tcgen05; stride-2 5x5 layers as 3x3 convs over the space-to-depth image, '
                                 'bias + LeakyReLU + Philox dropout in the conv epilogue), fp32 master weights and optimizer')},
        'clocks': clocks,
        'e2e': {'value': it_s_e2e, 'unit': unit, 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': IPS * (N_CRITIC * (B * row * 4 + 4) + 4), 'd2h_bytes_per_step': IPS * (N_CRITIC * 32 + 4)},
        'gpu_launches': launches,
        'step_tensor_utilisation': {'algorithmic_gflop_per_iteration': gflop, 'achieved_tflops_per_gpu': step_tflops,
                                    'frac_of_sustained_peak': step_tflops / peaks['sustained'], 'peak': peaks['sustained']},
    }
    if with_roofline and script in ('cifar', 'mnist'):
        line['roofline'] = roofline_dcgan_kernel(torch, peaks, script, B)
    return line


def run_dcgan(args):
    """`--workload cifar|mnist`: the DCGAN scripts as the bench line (BASELINE configs[1] / [0])."""
    import torch
    import torch.distributed as dist
    script = args.workload
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('LOCAL_RANK', '0'), ('WORLD_SIZE', '1')))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    line = measure_dcgan(args, script, rank, local_rank, world, dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = time_cpu_dcgan(script)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly (debug / profiling)')
    ap.add_argument('--no-pregen', action='store_true', help='one generator forward per critic step instead of one per iteration')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-workloads', action='store_true', help='skip the CIFAR-DCGAN / MNIST lines of the default run')
    ap.add_argument('--workload', default='resnet', choices=['resnet', 'cifar', 'mnist', '64x64', 'lsun128'],
                    help="resnet = BASELINE.json's metric (default); cifar / mnist = the DCGAN parity configurations (our arm only)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload != 'resnet':
        run_dcgan(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
