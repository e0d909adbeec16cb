"""Step-level runtime: device-side random streams, the flat-buffer Adam optimizer and the
data-parallel gradient exchange.

  * `DeviceRandom` replaces the TF graph's random ops (tf.random_normal / random_uniform /
    nn.dropout, TG/CT_gan_cifar_resnet.py:157,202,277-281,319): one Philox4x32-10 stream per
    training process, every draw takes a disjoint [offset, offset+n) slice.  With
    `record=True` each draw is also materialised under a tag so that the CPU oracle can
    replay the very same numbers (parity tests).
  * `FlatAdam` replaces tf.train.AdamOptimizer(...).minimize(cost, var_list=...)
    (TG/CT_gan_cifar.py:153-154, TG/CT_gan_cifar_resnet.py:333-338): parameters, gradients
    and both moments live in four flat float buffers, one fused kernel per update.
  * Data parallel: one process per GPU.  Default: the flat gradient bucket and the flat parameter
    buffer live in peer-visible memory (CUDA IPC over NVLink / NVSwitch) and ONE kernel per rank does reduce-scatter +
    Adam on the owned slice + all-gather of the new parameters (csrc/peer.cu) -- no NCCL call inside a step, so a
    data-parallel step is one CUDA graph.  Fallback (kernels.config.peer_update = False, or IPC unavailable): one NCCL
    all-reduce of the bucket, the 1/world factor folded into the Adam kernel.
"""
import math

import torch

from . import kernels as K
from . import tflib as lib

CL = torch.channels_last


class DeviceRandom:
    def __init__(self, seed, device, record=False, graph_safe=False):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device)
        self.offset = 0
        self.record = record
        self.tape = {}
        self._scope = ''
        self._parts = None           # [(tag prefix, rows)] when several reference passes run as one stacked batch
        self._tower = None           # (i, towers) while the passes run one after the other on their own rows (scope_tower)
        self._site = 0
        self.patterns = []
        self._stack = None
        # replay: {tag: tensor} of externally supplied draws (golden-vector tests): every draw is taken from here
        # instead of the Philox stream; dropout sites then receive explicit uniform tensors
        self.replay = None
        # graph_safe: kernels add a device-resident base offset, advanced by a kernel at step end
        self.dyn = torch.zeros(1, dtype=torch.int64, device=self.device) if graph_safe else None

    # -- stream bookkeeping ---------------------------------------------------
    def _take(self, n):
        off = self.offset
        self.offset += (int(n) + 3) // 4 * 4        # keep every slice 4-aligned (vectorised Philox)
        return off

    def end_step(self, side=False):
        """Graph-safe mode: fold the offsets consumed by this step into the device counter.
        side: launch the counter update on the side stream (the caller joins it: K.join_side()) -- nothing of the current
        step reads the counter any more once its backward pass is queued, so it overlaps the optimizer kernels."""
        if self.dyn is not None and self.offset:
            n, dyn = self.offset, self.dyn
            if side:
                K.on_side(lambda: K.counter_add(dyn, n), dyn)
            else:
                K.counter_add(dyn, n)
            self.offset = 0

    def begin_recording(self):
        from . import functional as F
        self.record = True
        self.tape = {}
        self.patterns = []
        self._stack = None
        F.pattern_recorder = self.patterns.append

    def stop_recording(self):
        from . import functional as F
        self.record = False
        F.pattern_recorder = None

    def scope(self, name):
        self._scope, self._parts, self._site, self._tower = name, None, 0, None

    def scope_tower(self, prefix, i, towers):
        """Dropout sites of tower i of `towers` equal passes that run one after the other (each on its own rows): site k of
        every tower draws from ONE Philox slice, tower i taking the i-th part of it -- exactly the numbers a single pass over
        the towers' batches stacked along dim 0 (scope_parts) draws, so the two execution orders are interchangeable.
        Tower 0 must run first; tags are '<prefix>.<i>.<site>' as for separate passes."""
        self._scope, self._parts, self._site, self._tower = '%s.%d' % (prefix, i), None, 0, (i, towers)
        if i == 0:
            self._tower_offs = []

    def scope_parts(self, parts):
        """The next dropout sites act on a batch that stacks several reference passes along dim 0:
        parts = [(tag prefix, rows), ...].  One Philox slice per site; when recording, each part's rows are
        exported under '<prefix>.<site>' exactly as if the passes had run separately."""
        self._scope, self._parts, self._site, self._tower = parts[0][0], list(parts), 0, None

    def begin_stack(self, rows):
        """Activation patterns recorded until end_stack() belong to a stacked batch with these row counts."""
        self._stack = (list(rows), len(self.patterns))

    def end_stack(self):
        if self._stack is None:
            return
        rows, start = self._stack
        self._stack = None
        seg = self.patterns[start:]
        if not seg:
            return
        out, r0 = [], 0
        for n in rows:                       # part-major: every site of part 0, then every site of part 1, ...
            out += [p[r0:r0 + n] for p in seg]
            r0 += n
        self.patterns[start:] = out

    def _split_keep(self, parts, t):
        if self.record:
            r0 = 0
            for tag, n in parts:
                self._keep(tag, t[r0:r0 + n])
                r0 += n

    def normal_parts(self, parts, cols):
        if self.replay is not None:
            return torch.cat([self._replayed(tag, n) for tag, n in parts], dim=0)
        rows = sum(n for _, n in parts)
        t = K.philox_normal((rows, cols), self.device, self.seed, self._take(2 * rows * cols), dyn=self.dyn)
        self._split_keep(parts, t)
        return t

    def labels_parts(self, parts, n_labels=10):
        if self.replay is not None:
            return torch.cat([self._replayed(tag, n, torch.int32) for tag, n in parts], dim=0)
        rows = sum(n for _, n in parts)
        t = K.philox_labels(rows, self.device, n_labels, self.seed, self._take(rows), dyn=self.dyn)
        self._split_keep(parts, t)
        return t

    def dropout_args(self, like):
        """Keyword arguments for F.dropout / F.leaky_relu_dropout at the next dropout site of the current scope:
        a Philox slice (seed, offset, dyn) or, in replay mode, the explicit uniform tensor `u`."""
        if self.replay is not None:
            self._site += 1
            parts = self._parts if self._parts is not None else [(self._scope, like.shape[0])]
            u = torch.cat([self._replayed('%s.%d' % (p, self._site), n) for p, n in parts], dim=0)
            if like.dim() == 4:
                u = u.contiguous(memory_format=CL)
            return dict(u=u)
        seed, off, dyn = self.dropout_stream(like)
        return dict(seed=seed, offset=off, dyn=dyn)

    def dropout_stream(self, like):
        """Philox slice for the next dropout site of the current scope; returns (seed, offset, dyn)."""
        self._site += 1
        if self._tower is not None:
            i, towers = self._tower
            n = like.numel()
            if i == 0:
                self._tower_offs.append(self._take(towers * n))
            off = self._tower_offs[self._site - 1] + i * n
            if self.record:
                mf = CL if (like.dim() == 4 and like.is_contiguous(memory_format=CL)) else None
                self._keep('%s.%d' % (self._scope, self._site),
                           K.philox_uniform(tuple(like.shape), self.device, self.seed, off, memory_format=mf, dyn=self.dyn))
            return self.seed, off, self.dyn
        if self._parts is None:
            return self.stream('%s.%d' % (self._scope, self._site), like)
        off = self._take(like.numel())
        if self.record:
            mf = CL if (like.dim() == 4 and like.is_contiguous(memory_format=CL)) else None
            u = K.philox_uniform(tuple(like.shape), self.device, self.seed, off, memory_format=mf, dyn=self.dyn)
            self._split_keep([('%s.%d' % (p, self._site), n) for p, n in self._parts], u)
        return self.seed, off, self.dyn

    def _keep(self, tag, t):
        if self.record:
            assert tag not in self.tape, 'duplicate random tag %s' % tag
            self.tape[tag] = t

    # -- draws ------------------------------------------------------------------
    def _replayed(self, tag, rows=None, dtype=torch.float32):
        t = torch.as_tensor(self.replay[tag])
        if rows is not None:
            t = t[:rows]                     # the reference ran this pass on more rows than the product needs
        return t.to(self.device).to(dtype).contiguous()

    def normal(self, tag, shape):
        if self.replay is not None:
            return self._replayed(tag)
        n = int(math.prod(shape))
        t = K.philox_normal(tuple(shape), self.device, self.seed, self._take(2 * n), dyn=self.dyn)
        self._keep(tag, t)
        return t

    def uniform(self, tag, shape, lo=0., hi=1.):
        if self.replay is not None:
            return self._replayed(tag)
        n = int(math.prod(shape))
        t = K.philox_uniform(tuple(shape), self.device, self.seed, self._take(n), lo, hi, dyn=self.dyn)
        self._keep(tag, t)
        return t

    def labels(self, tag, n, n_labels=10):
        if self.replay is not None:
            return self._replayed(tag, dtype=torch.int32)
        t = K.philox_labels(int(n), self.device, n_labels, self.seed, self._take(n), dyn=self.dyn)
        self._keep(tag, t)
        return t

    def stream(self, tag, like):
        """Reserve a slice for an in-kernel consumer (fused dropout / dequantisation noise) shaped
        like `like`; returns (seed, offset, dyn).  When recording, the same slice is materialised
        with the memory layout of `like`, i.e. element i of the buffer == stream element offset+i."""
        off = self._take(like.numel())
        if self.record:
            mf = CL if (like.dim() == 4 and like.is_contiguous(memory_format=CL)) else None
            u = K.philox_uniform(tuple(like.shape), self.device, self.seed, off, memory_format=mf, dyn=self.dyn)
            self._keep(tag, u)
        return self.seed, off, self.dyn

    def next_dropout_tag(self):
        self._site += 1
        return '%s.%d' % (self._scope, self._site)


def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class PeerBuffers:
    """Flat parameter and gradient buffers of one optimizer in peer-visible memory, plus the mappings of every other
    rank's buffers and flag blocks (CUDA IPC, one node): the operands of K.peer_reduce_adam."""

    @classmethod
    def create(cls, n, dev):
        import torch.distributed as dist
        import warnings
        self = cls()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        ok = torch.ones(1, device=dev)
        try:
            self.p, p_ptr = K.peer_alloc_floats(n, dev)
            self.g, g_ptr = K.peer_alloc_floats(n, dev)
            self.flags, f_ptr = K.peer_alloc_floats(K._lib.lib.ctgan_peer_flag_bytes() // 4, dev)
            mine = (K.ipc_handle(p_ptr), K.ipc_handle(g_ptr), K.ipc_handle(f_ptr))
        except Exception as e:                                  # e.g. IPC not permitted in this container
            warnings.warn('ctgan_b200: peer buffers unavailable (%s); falling back to NCCL all-reduce' % e)
            ok.zero_()
            mine = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)               # all ranks take the same path
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        if ok.item() == 0:
            return None
        self.p_ptrs, self.g_ptrs, self.flag_ptrs = [], [], []
        try:
            for r, h in enumerate(everyone):
                if r == self.rank:
                    ptrs = (p_ptr, g_ptr, f_ptr)
                else:
                    with torch.cuda.device(dev):
                        ptrs = tuple(K.ipc_open(x) for x in h)
                self.p_ptrs.append(ptrs[0]); self.g_ptrs.append(ptrs[1]); self.flag_ptrs.append(ptrs[2])
        except Exception as e:
            warnings.warn('ctgan_b200: could not map the peers\' buffers (%s); falling back to NCCL all-reduce' % e)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            return None
        dist.barrier()
        return self


class FlatAdam:
    """TF-semantics Adam over the parameters selected by name (substring), held flat.

    On construction the selected tflib parameters are MOVED into one flat float buffer
    (each tensor becomes a view, padded to 64 floats), their `.grad`s become views of a
    flat gradient buffer, and the registry is re-bound to the new Parameter objects."""

    PAD = 64

    def __init__(self, selector, lr, beta1, beta2, eps=1e-8):
        named = lib.named_params_with_name(selector, trainable_only=True)
        if not named:
            raise RuntimeError('FlatAdam: no trainable parameter matches %r' % selector)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        dev = next(iter(named.values())).device
        sizes = {n: p.numel() for n, p in named.items()}
        offs, total = {}, 0
        for n in named:
            offs[n] = total
            total += (sizes[n] + self.PAD - 1) // self.PAD * self.PAD
        self.n = total
        self.peer = None
        if dev.type == 'cuda' and K.config.peer_update and _dist_world() > 1:
            self.peer = PeerBuffers.create(total, dev)
        if self.peer is not None:
            self.flat_p, self.flat_g = self.peer.p, self.peer.g
        else:
            self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.lr_t_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        # lr_t staging: a RING of pinned slots -- the host runs many replays ahead of the device, so a single slot
        # would be overwritten before its asynchronous upload has executed
        self._lr_ring = torch.zeros(64, dtype=torch.float32).pin_memory() if dev.type == 'cuda' else None
        self._lr_ring_ev, self._lr_ring_i = [None] * 64, 0
        self.params = {}
        new = {}
        old_ptrs = {p.data_ptr() for p in named.values()}
        for n, p in named.items():
            o, sz = offs[n], sizes[n]
            self.flat_p[o:o + sz].copy_(p.detach().reshape(-1))
            q = torch.nn.Parameter(self.flat_p[o:o + sz].view(p.shape), requires_grad=True)
            q.grad = self.flat_g[o:o + sz].view(p.shape)
            new[n] = q
            self.params[n] = q
        lib.rebind_params(new)
        self.offsets, self.sizes = offs, sizes
        self._ptrs = set(q.data_ptr() for q in self.params.values())
        # only THIS optimizer's tensors moved: a blanket invalidation would also drop the packs another optimizer of the
        # same model published a moment ago (the generator's, when the critic's optimizer is built second)
        K.invalidate_weight_cache(old_ptrs, forget=True)
        self.packer = K.FilterPacker(self.flat_p, self.params, offs) if dev.type == 'cuda' else None
        # publish the one-launch operand packs NOW: a filter first used before the first optimizer step would otherwise
        # get its own lazily created pack, re-packed by one extra launch per filter and layout after every step
        # (16 + 18 tail launches per ResNet critic / generator step in the round-1 timelines)
        self.refresh_packs()

    def param_list(self):
        return list(self.params.values())

    def zero_grad(self):
        K.zero_(self.flat_g) if self.flat_g.is_cuda else self.flat_g.zero_()
        for q in self.params.values():       # keep .grad bound to the flat views
            if q.grad is None or q.grad.data_ptr() != self.flat_g.data_ptr() + 4 * self.offsets[q.ctgan_name]:
                q.grad = self.flat_g[self.offsets[q.ctgan_name]:self.offsets[q.ctgan_name] + q.numel()].view(q.shape)

    def lr_t(self, lr=None):
        lr = self.lr if lr is None else lr
        t = self.t
        return lr * math.sqrt(1. - self.beta2 ** t) / (1. - self.beta1 ** t)

    def all_reduce(self):
        """Sum the flat gradient bucket over ranks (NCCL); the 1/world scale is applied in step().  With peer buffers the
        exchange happens inside step() (one kernel: reduce-scatter + Adam + all-gather): nothing to do here."""
        import torch.distributed as dist
        K.join_side()
        if self.peer is not None:
            return self.peer.world
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g)
            return dist.get_world_size()
        return 1

    def step(self, lr=None, world=1, use_device_lr=False):
        """One update.  use_device_lr: read lr_t from self.lr_t_dev (set with set_device_lr) so the
        launch is replayable inside a CUDA graph."""
        K.join_side()                        # filter / bias gradients accumulated on the side stream
        if not use_device_lr:
            self.t += 1                      # graph mode: t advances in set_device_lr() at replay time
        if self.peer is not None:
            pb = self.peer
            K.peer_reduce_adam(pb.world, pb.rank, pb.g_ptrs, pb.p_ptrs, pb.flag_ptrs, self.flat_m, self.flat_v, self.n,
                               self.lr_t(lr) if not use_device_lr else 0.0, self.beta1, self.beta2, self.eps,
                               grad_scale=1.0 / pb.world, lr_t_dev=self.lr_t_dev if use_device_lr else None)
        else:
            K.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr_t(lr) if not use_device_lr else 0.0,
                        self.beta1, self.beta2, self.eps, grad_scale=1.0 / world,
                        lr_t_dev=self.lr_t_dev if use_device_lr else None)
        K.invalidate_weight_cache(self._ptrs)
        self.refresh_packs()

    def refresh_packs(self):
        """BF16 operand copies of all filters of this optimizer for the new parameter values (one launch)."""
        if self.packer is not None:
            self.packer.refresh()

    def set_device_lr(self, lr=None, advance=True):
        """Host side of a graph replay: bump t and upload lr_t (async, pinned)."""
        if advance:
            self.t += 1
        i = self._lr_ring_i % 64
        self._lr_ring_i += 1
        if self._lr_ring_ev[i] is not None:
            self._lr_ring_ev[i].synchronize()           # the upload that last used this slot has executed
        self._lr_ring[i] = self.lr_t(lr)
        self.lr_t_dev.copy_(self._lr_ring[i:i + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._lr_ring_ev[i] = ev

    def state_dict(self):
        return {'t': self.t, 'p': {n: q.detach().clone() for n, q in self.params.items()},
                'm': self.flat_m.clone(), 'v': self.flat_v.clone()}
