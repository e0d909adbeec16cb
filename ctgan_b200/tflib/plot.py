"""Metric logging with the reference's interface and files (TG/tflib/plot.py:11-41): `plot(name, value)` records a
value for the current iteration, `tick()` advances the iteration counter, `flush()` prints
`iter <n>\\t<name>\\t<mean since last flush>...`, rewrites one `<name>.jpg` curve per metric and dumps the whole history
to `log.pkl` (a `{name: {iteration: value}}` dict, pickle.HIGHEST_PROTOCOL).

Differences:
  * `value` may be a 0-d CUDA tensor (or a 1-element slice of the step's loss vector).  It is copied to pinned host
    memory asynchronously on the current stream and only read in `flush()`, so logging never synchronises the training
    stream (the reference fetches every scalar with `session.run`, TG/CT_gan_cifar_resnet.py:402-412).
  * curves are drawn by a few lines of numpy + Pillow (matplotlib is optional and not required).
  * `output_dir` (module attribute, default '.') is where the files go.
"""
import collections
import os
import pickle

import numpy as np

_since_beginning = collections.defaultdict(lambda: {})
_since_last_flush = collections.defaultdict(lambda: {})

_iter = [0]
output_dir = '.'
write_curves = True


class _Pending:
    """A device scalar on its way to pinned host memory."""
    __slots__ = ('host', 'event')

    def __init__(self, t):
        import torch
        self.host = torch.empty((), dtype=t.dtype).pin_memory()
        self.host.copy_(t.detach().reshape(()), non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def value(self):
        self.event.synchronize()
        return self.host.item()


def _resolve(v):
    """Lazy metrics (device scalars copied asynchronously, device-timed durations) resolve when the log is flushed."""
    return v.value() if callable(getattr(v, 'value', None)) else v


def tick():
    _iter[0] += 1


def plot(name, value):
    if hasattr(value, 'is_cuda'):                       # torch tensor
        value = _Pending(value) if value.is_cuda else value.detach().reshape(()).item()
    _since_last_flush[name][_iter[0]] = value


def reset():
    _since_beginning.clear()
    _since_last_flush.clear()
    _iter[0] = 0


def _draw_curve(x_vals, y_vals, name, path, size=(640, 480)):
    """Minimal line plot: axes box, polyline, axis labels and range annotations."""
    from PIL import Image, ImageDraw
    W, H = size
    L, R, T, B = 70, 20, 20, 50
    img = Image.new('RGB', size, 'white')
    d = ImageDraw.Draw(img)
    x, y = np.asarray(x_vals, dtype='float64'), np.asarray(y_vals, dtype='float64')
    ok = np.isfinite(y)
    x0, x1 = (x.min(), x.max()) if len(x) else (0., 1.)
    y0, y1 = (y[ok].min(), y[ok].max()) if ok.any() else (0., 1.)
    if x1 == x0:
        x1 = x0 + 1
    if y1 == y0:
        y1 = y0 + 1
    d.rectangle([L, T, W - R, H - B], outline='black')
    px = L + (x - x0) / (x1 - x0) * (W - L - R)
    py = (H - B) - (np.where(ok, y, y0) - y0) / (y1 - y0) * (H - T - B)
    pts = [(float(a), float(b)) for a, b, k in zip(px, py, ok) if k]
    if len(pts) > 1:
        d.line(pts, fill=(31, 119, 180), width=2)
    elif pts:
        d.ellipse([pts[0][0] - 2, pts[0][1] - 2, pts[0][0] + 2, pts[0][1] + 2], fill=(31, 119, 180))
    d.text((L, H - B + 8), '%g' % x0, fill='black')
    d.text((W - R - 50, H - B + 8), '%g' % x1, fill='black')
    d.text((W // 2 - 30, H - 22), 'iteration', fill='black')
    d.text((4, H - B - 10), '%.4g' % y0, fill='black')
    d.text((4, T), '%.4g' % y1, fill='black')
    d.text((L + 6, T + 4), name, fill='black')
    img.save(path)


def flush():
    prints = []
    for name, vals in _since_last_flush.items():
        vals = {k: _resolve(v) for k, v in vals.items()}
        prints.append("{}\t{}".format(name, np.mean(list(vals.values()))))
        _since_beginning[name].update(vals)
        if write_curves:
            x_vals = np.sort(list(_since_beginning[name].keys()))
            y_vals = [_since_beginning[name][x] for x in x_vals]
            _draw_curve(x_vals, y_vals, name, os.path.join(output_dir, name.replace(' ', '_') + '.jpg'))
    print("iter {}\t{}".format(_iter[0], "\t".join(prints)))
    _since_last_flush.clear()
    with open(os.path.join(output_dir, 'log.pkl'), 'wb') as f:
        pickle.dump(dict(_since_beginning), f, pickle.HIGHEST_PROTOCOL)
