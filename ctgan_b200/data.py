"""Host -> HBM input pipeline for the training step.

The reference feeds every critic step through `feed_dict` (TG/CT_gan_cifar_resnet.py:400-402): a synchronous
host -> device copy of an int32 [64, 3072] batch, 4 bytes per pixel, inside `session.run`.  Here the epoch generators of
`tflib.cifar10` / `tflib.mnist` are wrapped in `inf_train_gen` (the scripts' `while True: for images, labels in
train_gen()` loop, :363-366) and a `DeviceFeeder`:
  * batches are staged in a ring of PINNED host buffers (pixels stay uint8: 1 byte/pixel over PCIe) and copied with
    `cudaMemcpyAsync` on a dedicated copy stream, `depth` batches ahead of the consumer;
  * the consumer's stream waits on the copy's event only; a slot is refilled only after the work that read it was
    queued (event recorded at the next `next()`), so neither side ever blocks the host;
  * the input scaling 2*(x/256 - .5) + U[0, 1/128) (:201-202) runs on the device (`ctgan_prep_real_u8`, inside the
    captured critic graph).
Label vectors are converted to int32 on the host (the placeholder dtype, :191-192).
"""
import numpy as np
import torch


def inf_train_gen(epoch_fn, rank=0, world=1):
    """TG/CT_gan_cifar_resnet.py:363-366 / TG/CT_gan_mnist.py:220-223.  Data parallel (world > 1): every rank runs the
    SAME generator (same numpy seed => same shuffles) and keeps batches rank, rank + world, ... -- disjoint shards of one
    epoch order, the batch-sharding the reference's tower loop does inside one process (TG/CT_gan_64x64.py:475-480)."""
    k = 0
    while True:
        for batch in epoch_fn():
            if k % world == rank:
                yield batch
            k += 1


def _host_dtype(a, keep_uint8):
    a = np.asarray(a)
    if a.dtype == np.uint8:
        return a if keep_uint8 else a.astype('int32')
    if a.dtype.kind in 'iu':
        return a.astype('int32')
    return a.astype('float32')


class DeviceFeeder:
    """Iterator of device-resident batches over an (infinite) host batch iterator.

    it: iterator yielding tuples of numpy arrays of FIXED shapes (e.g. `inf_train_gen(train_epoch)`);
    take: how many leading arrays of each tuple to ship (CIFAR ResNet: images + labels = 2; DCGAN scripts: 1);
    depth: batches copied ahead of the consumer; hold: how many of the most recently returned batches stay valid
    (the ResNet iteration looks at the labels of its 5 critic batches before the first critic step);
    keep_uint8: ship uint8 pixels as bytes (the step's input preparation accepts them) instead of widening to int32.
    """

    def __init__(self, it, device, depth=2, take=None, keep_uint8=True, hold=1):
        self.it, self.device = iter(it), torch.device(device)
        self.take, self.keep_uint8 = take, keep_uint8
        self.cuda = self.device.type == 'cuda'
        self.depth, self.hold = max(1, int(depth)), max(1, int(hold))
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.slots = None                    # ring of (pinned host tuple, device tuple)
        self.ready = []                      # per slot: event recorded after its H2D copy
        self.head = 0                        # next slot to fill
        self.tail = 0                        # next slot to hand out
        self.bytes_per_batch = 0
        self.batches = 0
        for _ in range(self.depth):
            self._fill(None)

    def _alloc(self, arrays):
        self.slots = []
        for _ in range(self.depth + self.hold):
            host = tuple(torch.from_numpy(np.empty_like(a)) for a in arrays)
            if self.cuda:
                host = tuple(h.pin_memory() for h in host)
                dev = tuple(torch.empty_like(h, device=self.device) for h in host)
            else:
                dev = tuple(torch.empty_like(h) for h in host)
            self.slots.append((host, dev))
        self.ready = [None] * len(self.slots)
        self.bytes_per_batch = sum(a.nbytes for a in arrays)

    def _fill(self, consumed):
        """Stage the next host batch into slot `head`.  consumed: event covering every kernel that may still read the
        slot's device buffers (None while the ring is being primed)."""
        batch = next(self.it)
        arrays = [np.ascontiguousarray(_host_dtype(a, self.keep_uint8)) for a in (batch[:self.take] if self.take else batch)]
        if self.slots is None:
            self._alloc(arrays)
        i = self.head
        host, dev = self.slots[i]
        if self.ready[i] is not None:
            self.ready[i].synchronize()      # the previous H2D copy out of this pinned buffer has executed
        for h, a in zip(host, arrays):
            if tuple(h.shape) != a.shape or h.numpy().dtype != a.dtype:
                raise RuntimeError('DeviceFeeder: batch shapes / dtypes must not change (%s %s vs %s %s)'
                                   % (tuple(h.shape), h.numpy().dtype, a.shape, a.dtype))
            h.numpy()[...] = a
        if self.cuda:
            with torch.cuda.stream(self.copy_stream):
                if consumed is not None:
                    self.copy_stream.wait_event(consumed)
                for h, d in zip(host, dev):
                    d.copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self.ready[i] = ev
        else:
            for h, d in zip(host, dev):
                d.copy_(h)
        self.head = (i + 1) % len(self.slots)

    def __iter__(self):
        return self

    def __next__(self):
        consumed = None
        if self.cuda:
            consumed = torch.cuda.Event()
            consumed.record(torch.cuda.current_stream(self.device))     # everything queued on the consumer's stream so far
        i = self.tail
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        out = self.slots[i][1]
        self.tail = (i + 1) % len(self.slots)
        self.batches += 1
        self._fill(consumed)                 # refills the slot handed out `hold` calls ago
        return out

    next = __next__
