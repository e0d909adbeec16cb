"""The training loops of the three scripts (TG/CT_gan_mnist.py:226-271, TG/CT_gan_cifar.py:186-236,
TG/CT_gan_cifar_resnet.py:391-434) around the B200 training step.

Per iteration, exactly as the reference schedules it: one generator step (skipped at iteration 0), then CRITIC_ITERS
critic steps each on a fresh real batch; `lib.plot` metrics with the reference's names; every `dev_every` iterations
the critic cost over the dev set and a sample grid from fixed noise; `lib.plot.flush()` / `tick()` at the reference's
cadence.  What differs is how it executes: both steps are CUDA-graph replays (graphs.GraphedTrainer), real batches
arrive through data.DeviceFeeder (pinned, asynchronous, uint8), metrics are fetched asynchronously (tflib.plot), so
the host never waits for the device inside an iteration.  Inception score (TG/tflib/inception_score.py: a frozen TF
graph downloaded at run time) is out of scope.

    python -m ctgan_b200.train cifar_resnet --data-dir /path/to/cifar-10-batches-py --iters 1000
"""
import argparse
import importlib
import os
import time

import numpy as np
import torch

from . import checkpoint
from .data import DeviceFeeder, inf_train_gen
from .graphs import GraphedTrainer
from .tflib import plot as _plot, save_images as _save_images, cifar10 as _cifar10, mnist as _mnist, small_imagenet as _imagenet
from .tflib import imagenet as _lsun

SCRIPTS = {'mnist': 'gan_mnist', 'cifar': 'gan_cifar', 'cifar_resnet': 'gan_cifar_resnet',
           '64x64': 'gan_64x64',        # 64x64: SURVEY.md 8(f) N4, see gan_64x64.py
           'lsun128': 'gan_lsun128'}    # LS/wgan_LSUN_Bedrooms128.py (same row): its own schedule, see run_iteration / train


def _loaders(script, mod, batch_size, data_dir, n_examples):
    if script == 'lsun128':                                  # LS/wgan_LSUN_Bedrooms128.py:351-352: one image folder, no dev set
        return _lsun.load(batch_size, data_dir), None
    if script == 'mnist':
        train, dev, _ = _mnist.load(batch_size, batch_size, n_examples, filepath=data_dir)
    elif script == '64x64':                                  # TG/CT_gan_64x64.py:613; n_examples = (n_train, n_valid) files
        n_train, n_valid = n_examples if isinstance(n_examples, (tuple, list)) else (1281149, 49999)
        train, dev = _imagenet.load(batch_size, data_dir, n_train, n_valid)
    else:
        train, dev = _cifar10.load(batch_size, data_dir, n_examples)
    return train, dev


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _EventSeconds:
    """A device-timed duration that resolves when lib.plot flushes (value()): the iteration time without a sync."""

    def __init__(self, ev0, ev1):
        self.ev0, self.ev1 = ev0, ev1

    def value(self):
        self.ev1.synchronize()
        return self.ev0.elapsed_time(self.ev1) * 1e-3


class Session:
    """One training run: model, graphs, feeder, fixed sample noise."""

    def __init__(self, script, data_dir, batch_size=None, n_examples=None, device='cuda', seed=1234, out_dir='.',
                 act_dtype=torch.bfloat16, use_graphs=True, init_seed=1234, model_kw=None, acc_every=0):
        self.script, self.out_dir = script, out_dir
        self.mod = mod = importlib.import_module('ctgan_b200.' + SCRIPTS[script])
        self.B = batch_size or mod.BATCH_SIZE
        self.n_critic = getattr(mod, 'N_CRITIC', None) or mod.CRITIC_ITERS
        self.resnet = script == 'cifar_resnet'
        np.random.seed(init_seed)                               # initial weights: the reference's numpy-global init
        # data parallel: identical initial weights and data order on every rank, rank-dependent random streams / shards
        self.tr = mod.Trainer(device=device, seed=seed + _rank_world()[0], act_dtype=act_dtype, batch_size=self.B,
                              graph_safe_rng=use_graphs, **(model_kw or {}))
        self.train_epoch, self.dev_epoch = _loaders(script, mod, self.B, data_dir, n_examples or getattr(mod, 'n_examples', None))
        take = 2 if self.resnet else 1
        rank, world = _rank_world()
        self.feeder = DeviceFeeder(inf_train_gen(self.train_epoch, rank, world), device, depth=2, take=take, hold=self.n_critic)
        # fixed noise for the sample grids (:341-343 / TG/CT_gan_cifar.py:157-158 / TG/CT_gan_mnist.py:206-207)
        n_fixed = 100 if self.resnet else (self.B if script == '64x64' else (64 if script == 'lsun128' else 128))   # TG/CT_gan_64x64.py:582, LS :341
        self.fixed_noise = torch.from_numpy(np.random.normal(size=(n_fixed, 128)).astype('float32')).to(device)
        self.fixed_labels = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7, 8, 9] * 10, dtype=torch.int32, device=device) if self.resnet else None
        self.gt = None
        if use_graphs:
            first = next(self.feeder)
            self._pending = first
            self.gt = GraphedTrainer(self.tr, tuple(first), pregen_steps=self.n_critic if self.resnet else 0)
        else:
            self._pending = None
        self.iteration = 0
        self.acc_every, self._last_fakes = acc_every, None

    # -- one reference iteration ----------------------------------------------------------------------------------
    def _next_batch(self):
        if self._pending is not None:
            b, self._pending = self._pending, None
            return b
        return next(self.feeder)

    def run_iteration(self):
        it, tr, gt = self.iteration, self.tr, self.gt
        start = time.time()
        if gt is not None:
            gt.iteration = it
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        lsun = self.script == 'lsun128'          # LS/wgan_LSUN_Bedrooms128.py:373-389: the critic steps FIRST, then a generator step in
        if it > 0 and not lsun:                  # every iteration (also iteration 0)
            gt.gen_step() if gt is not None else tr.gen_step(iteration=it)
        batches = [self._next_batch() for _ in range(self.n_critic)]      # all stay valid: feeder hold = n_critic
        if gt is not None and gt.pregen_steps:
            gt.begin_iteration(torch.cat([b[1] for b in batches]))
        out = None
        for b in batches:
            if gt is not None:
                out = gt.critic_step(*b)
            else:
                res = tr.critic_step(*b, iteration=it)
                out, self._last_fakes = res['out'], res.get('fake_data')
        if lsun:
            gt.gen_step() if gt is not None else tr.gen_step(iteration=it)
            _plot.plot('cost', out[0])                                    # LS :391
        elif self.resnet:
            # names and contents of TG/CT_gan_cifar_resnet.py:406-412: 'wgan' is disc_wgan = Wasserstein term + CT + 10*GP
            # (:295), i.e. the cost without its ACGAN part; out = {cost, wgan term, ct, gp, acgan, ...}.  The two clean-pass
            # accuracies ('acc_real', 'acc_fake', :410-411) come from a metrics-only critic pass that the training step does not
            # execute: it runs here every `acc_every` iterations (0 = never) on the last critic batch and its fakes.
            _plot.plot('cost', out[0])
            if self.mod.CONDITIONAL and self.mod.ACGAN:
                _plot.plot('wgan', out[0] - self.mod.ACGAN_SCALE * out[4])
                _plot.plot('acgan', out[4])
                if self.acc_every and it % self.acc_every == 0:
                    fakes = gt.fake_cur if (gt is not None and gt.pregen_steps) else self._last_fakes
                    if fakes is not None:
                        acc_real, acc_fake = tr.acgan_accuracy(batches[-1][0], batches[-1][1], fakes)
                        _plot.plot('acc_real', acc_real)
                        _plot.plot('acc_fake', acc_fake)
        else:
            _plot.plot('train disc cost', out[0])
        # graph mode: replays are asynchronous, so host wall time would only measure the enqueue; time the iteration on the
        # device instead (events resolved lazily in flush(), like the other metrics)
        if gt is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            _plot.plot('time', _EventSeconds(ev0, ev1))
        else:
            _plot.plot('time', time.time() - start)
        self.iteration = it + 1
        return out

    # -- the every-100-iterations block -----------------------------------------------------------------------------
    def dev_cost(self, max_batches=None):
        """Mean critic cost over one dev epoch (fresh dropout / interpolation draws, no parameter update)."""
        tr, costs = self.tr, []
        for k, batch in enumerate(self.dev_epoch()):
            if max_batches is not None and k >= max_batches:
                break
            arrays = [torch.from_numpy(np.ascontiguousarray(a if a.dtype == np.uint8 else
                                                           (a.astype('int32') if a.dtype.kind in 'iu' else a.astype('float32'))))
                      .to(tr.device) for a in (batch[:2] if self.resnet else batch[:1])]
            tr.disc_opt.zero_grad()
            costs.append(tr.critic_forward_backward(*arrays)['out'][0:1].clone())
            tr.rng.end_step()
        tr.disc_opt.zero_grad()
        return float(torch.cat(costs).mean().item()) if costs else float('nan')

    def generate_image(self, frame):
        mod = self.mod
        with torch.no_grad():
            if self.resnet:
                samples = mod.Generator(self.fixed_noise.shape[0], self.fixed_labels, noise=self.fixed_noise)
            elif self.script == '64x64':                     # one Generator call per tower (:584-586): per-tower BN statistics
                mod.BN_GROUPS = mod.N_GPUS
                try:
                    samples = mod.Generator(self.fixed_noise.shape[0], noise=self.fixed_noise)
                finally:
                    mod.BN_GROUPS = 1
            else:
                samples = mod.Generator(self.fixed_noise.shape[0], noise=self.fixed_noise)
        samples = samples.float().cpu().numpy()
        if self.script == 'mnist':
            path = os.path.join(self.out_dir, 'samples_{}.png'.format(frame))
            _save_images.save_images(samples.reshape((-1, 28, 28)), path)
        elif self.script in ('64x64', 'lsun128'):
            samples = ((samples + 1.) * (255.99 / 2)).astype('int32')               # :593 / LS :345
            path = os.path.join(self.out_dir, 'samples_{}.png'.format(frame))
            side = 64 if self.script == '64x64' else 128
            _save_images.save_images(samples.reshape((-1, 3, side, side)), path)
        else:
            samples = ((samples + 1.) * (255. / 2)).astype('int32')
            ext = 'png' if self.resnet else 'jpg'
            path = os.path.join(self.out_dir, 'samples_{}.{}'.format(frame, ext))
            _save_images.save_images(samples.reshape((-1, 3, 32, 32)), path)
        return path


def train(script, data_dir, iters=None, dev_every=None, out_dir='.', checkpoint_every=None, dev_batches=None, resume=None, **kw):
    """Run up to iteration `iters` (default: the script's ITERS).  resume: a checkpoint written by checkpoint.save() -- weights,
    Adam state, the Philox counters and the iteration (learning-rate decay) continue from it.  Returns the Session."""
    os.makedirs(out_dir, exist_ok=True)
    s = Session(script, data_dir, out_dir=out_dir, **kw)
    if resume:
        s.iteration = checkpoint.load(resume, s.tr).get('iteration', 0)
    dev_every = dev_every or (200 if script == '64x64' else 100)          # TG/CT_gan_64x64.py:656
    _plot.reset()
    _plot.output_dir = out_dir
    iters = iters if iters is not None else s.mod.ITERS
    flush_early = 500 if s.resnet else 5                     # :431 `iteration < 500`; DCGAN scripts: `iteration < 5`
    flush_every = 1000 if s.resnet else (200 if script == '64x64' else 100)
    writer = _rank_world()[0] == 0                           # replicas are identical: rank 0 writes the files
    if script == 'lsun128':
        # LS/wgan_LSUN_Bedrooms128.py:365,373-400: samples before the loop and every 100 iterations (at iteration % 100 == 0),
        # params.ckpt every 1000, flush every 5 (its fork of plot.flush also prints standard deviations: not reproduced), no dev set
        if writer and s.iteration == 0:
            s.generate_image(0)
        for iteration in range(s.iteration, iters):
            s.run_iteration()
            if writer and iteration % 100 == 0:
                s.generate_image(iteration)
            every = checkpoint_every or 1000
            if writer and iteration % every == 0:
                checkpoint.save(os.path.join(out_dir, 'checkpoint.npz'), s.tr, iteration=iteration + 1)
            if iteration % 5 == 0:
                if writer:
                    _plot.flush()
                else:
                    _plot._since_last_flush.clear()
            _plot.tick()
        return s
    for iteration in range(s.iteration, iters):
        s.run_iteration()
        if iteration % dev_every == dev_every - 1:
            _plot.plot('dev_cost' if s.resnet else 'dev disc cost', s.dev_cost(dev_batches))
            if writer:
                s.generate_image(iteration)
                if script == 'cifar':
                    checkpoint.save_disc_params_pyn(os.path.join(out_dir, 'param.pyn'))   # TG/CT_gan_cifar.py:216-222
        if writer and checkpoint_every and iteration % checkpoint_every == checkpoint_every - 1:
            checkpoint.save(os.path.join(out_dir, 'checkpoint.npz'), s.tr, iteration=iteration + 1)
        if iteration < flush_early or iteration % flush_every == flush_every - 1:
            if writer:
                _plot.flush()
            else:
                _plot._since_last_flush.clear()
        _plot.tick()
    return s


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('script', choices=sorted(SCRIPTS))
    ap.add_argument('--data-dir', required=True, help='cifar-10-batches-py directory, or the mnist.pkl.gz path')
    ap.add_argument('--iters', type=int, default=None)
    ap.add_argument('--out-dir', default='.')
    ap.add_argument('--batch-size', type=int, default=None)
    ap.add_argument('--n-examples', type=int, default=None)
    ap.add_argument('--checkpoint-every', type=int, default=None)
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--resume', default=None, help='checkpoint.npz of an earlier run to continue from')
    ap.add_argument('--acc-every', type=int, default=0,
                    help="cifar_resnet: plot 'acc_real' / 'acc_fake' (the reference's metrics-only critic pass) every N iterations; 0 = never")
    a = ap.parse_args()
    train(a.script, a.data_dir, iters=a.iters, out_dir=a.out_dir, batch_size=a.batch_size, n_examples=a.n_examples,
          checkpoint_every=a.checkpoint_every, use_graphs=not a.no_graphs, resume=a.resume, acc_every=a.acc_every)


if __name__ == '__main__':
    main()
