"""CUDA-graph capture of the critic step and the generator step.

At batch 64 a CT-GAN iteration is several hundred short kernels (hard part 1):
launched eagerly through Python/ctypes/autograd it is host-bound.  Both steps are therefore
captured once (forward, first- and second-order backward, Adam, RNG counter advance) and
replayed; everything that changes between replays lives in device memory:
  * the real batch / labels      -> static input buffers (copied into before each replay),
  * random numbers               -> Philox streams offset by a device counter (runtime.DeviceRandom),
  * the learning rate lr_t       -> FlatAdam.lr_t_dev, uploaded from pinned memory per replay.

Data parallel (world_size > 1): with peer buffers (runtime.PeerBuffers, the default) the gradient exchange is part of
the update kernel (csrc/peer.cu: reduce-scatter + Adam + all-gather over NVLink peer memory), so a step is ONE graph
exactly like on a single GPU.  NCCL fallback: each step is captured as TWO graphs -- (zero_grad, forward, backward) and
(Adam, RNG advance) -- with the all-reduce of the flat gradient bucket issued eagerly between them on the same stream.
"""
import torch

from . import kernels as K


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class _Step:
    """One training step (critic or generator) as one graph (single GPU) or two graphs + all-reduce."""

    def __init__(self, opt, fwd_bwd, rng, pool):
        self.opt, self.rng, self.world = opt, rng, _world()
        self.one_graph = self.world == 1 or getattr(opt, 'peer', None) is not None
        self.kernels = 0

        def whole():
            opt.zero_grad()
            out = fwd_bwd()
            rng.end_step(side=True)               # the Philox counter advances on the side stream, next to the optimizer kernels
            opt.step(None, 1, use_device_lr=True)
            K.join_side()
            return out

        def part_a():
            opt.zero_grad()
            return fwd_bwd()

        def part_b():
            opt.step(None, self.world, use_device_lr=True)
            rng.end_step()

        k0 = K._lib.lib.ctgan_kernel_launches()
        if self.one_graph:
            self.g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g, pool=pool):
                self.out = whole()
            self.pool = self.g.pool()
        else:
            self.ga, self.gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.ga, pool=pool):
                self.out = part_a()
            self.pool = self.ga.pool()
            offset = rng.offset                   # consumed by part A; folded into the device counter by part B
            with torch.cuda.graph(self.gb, pool=self.pool):
                rng.offset = offset
                part_b()
        self.kernels = K._lib.lib.ctgan_kernel_launches() - k0

    def replay(self):
        if self.one_graph:
            self.g.replay()
        else:
            self.ga.replay()
            self.opt.all_reduce()
            self.gb.replay()
        return self.out


class GraphedTrainer:
    """Wraps a Trainer (gan_cifar_resnet / gan_cifar / gan_mnist) built with graph_safe_rng=True.

    pregen_steps = S > 0 (trainers with generate_fakes): the fake batches of the next S critic steps come from ONE
    generator forward, replayed by begin_iteration(labels of those S steps) -- the generator is constant between two
    generator steps, and one 5x64-sample forward costs far less than five 64-sample forwards of latency-bound kernels."""

    def __init__(self, trainer, example_inputs, warmup=3, pregen_steps=0):
        self.tr = tr = trainer
        self.pregen_steps = pregen_steps if hasattr(trainer, 'generate_fakes') else 0
        self.static_inputs = tuple(torch.empty_like(t) for t in example_inputs)
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t)
        self.iteration = 0
        if tr.rng.dyn is None:
            raise RuntimeError('GraphedTrainer needs a Trainer created with graph_safe_rng=True')
        tr.rng.record = False

        S = self.pregen_steps
        if S:
            dev, B = example_inputs[0].device, example_inputs[0].shape[0]
            self.labels_all = example_inputs[1].repeat(S).contiguous()
            self.fake_cur = torch.zeros(B, example_inputs[0].shape[1], dtype=torch.float32, device=dev)
            self._k = 0

        def critic_fb():
            if S:
                return tr.critic_forward_backward(*self.static_inputs, fake_data=self.fake_cur)['out']
            return tr.critic_forward_backward(*self.static_inputs)['out']

        def gen_fb():
            return tr.gen_forward_backward()['cost']

        # The warm-up below runs REAL steps (allocator warm-up, lazy operand packs, NCCL buffers): snapshot everything they
        # change -- parameters, Adam moments and step counts, the Philox counters -- and restore it after the capture, so
        # that iteration 0 starts from the initial model exactly like the eager / reference loop (TG/CT_gan_cifar_resnet.py
        # :393-404 runs no generator step and no update before the first critic step)
        opts = (tr.disc_opt, tr.gen_opt)
        snap = [(o.flat_p.clone(), o.flat_m.clone(), o.flat_v.clone(), o.t) for o in opts]
        rng_snap = (tr.rng.dyn.clone(), tr.rng.offset)

        # warm-up on a side stream (allocator + lazy init), eager, including the all-reduce
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                tr.disc_opt.set_device_lr(self._lr())
                tr.critic_step(*self.static_inputs, use_device_lr=True)
                tr.gen_opt.set_device_lr(self._lr())
                tr.gen_step(use_device_lr=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        self.critic = _Step(tr.disc_opt, critic_fb, tr.rng, None)
        self.gen = _Step(tr.gen_opt, gen_fb, tr.rng, self.critic.pool)
        self.critic_kernels, self.gen_kernels = self.critic.kernels, self.gen.kernels
        self.pregen_kernels = 0
        if S:
            k0 = K._lib.lib.ctgan_kernel_launches()
            self.pregen = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.pregen, pool=self.critic.pool):
                fakes = tr.generate_fakes(self.labels_all)
                tr.rng.end_step()
            self.pregen_kernels = K._lib.lib.ctgan_kernel_launches() - k0
            # The three graphs share one memory pool: a tensor produced by one graph may be overwritten by the NEXT replay of
            # another (its block is free while that graph is captured).  The fakes are read across S critic replays, so they
            # are copied out of the pool right after the generator replay (one 4 MB copy per iteration).
            self._fakes_pool = fakes
            self.fakes_all = torch.empty_like(fakes)

        # undo the warm-up steps (captures execute nothing): initial weights, zero moments, t = 0, Philox counters
        torch.cuda.synchronize()
        for o, (p0, m0, v0, t0) in zip(opts, snap):
            o.flat_p.copy_(p0); o.flat_m.copy_(m0); o.flat_v.copy_(v0)
            o.t = t0
            K.invalidate_weight_cache(o._ptrs)
            o.refresh_packs()                       # bf16 operand copies, re-packed in place at the captured addresses
        tr.rng.dyn.copy_(rng_snap[0])
        tr.rng.offset = rng_snap[1]
        torch.cuda.synchronize()

    def _lr(self):
        return self.tr.lr(self.iteration) if hasattr(self.tr, 'lr') else None

    def begin_iteration(self, labels_all, non_blocking=True):
        """pregen mode: labels_all = the real labels of the next pregen_steps critic batches (step-major, device or
        pinned host); replays the batched generator forward."""
        self.labels_all.copy_(labels_all.reshape(-1), non_blocking=non_blocking)
        self.pregen.replay()
        self.fakes_all.copy_(self._fakes_pool)
        self._k = 0

    def critic_step(self, *inputs, non_blocking=True):
        """inputs: tensors (device or pinned host) copied into the static buffers, then one replay.
        Returns the static float[8] loss tensor {cost, wgan, ct, gp, acgan, ...} (device)."""
        for s, t in zip(self.static_inputs, inputs):
            s.copy_(t, non_blocking=non_blocking)
        if self.pregen_steps:
            if self._k >= self.pregen_steps:
                raise RuntimeError('GraphedTrainer: begin_iteration() must precede every %d critic steps' % self.pregen_steps)
            B = self.fake_cur.shape[0]
            self.fake_cur.copy_(self.fakes_all[self._k * B:(self._k + 1) * B])
            self._k += 1
        self.tr.disc_opt.set_device_lr(self._lr())
        return self.critic.replay()

    def gen_step(self):
        self.tr.gen_opt.set_device_lr(self._lr())
        return self.gen.replay()
