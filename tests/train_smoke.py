"""A few hundred iterations of the CT_gan_cifar_resnet.py loop (graph mode, BF16, batch 64) on a synthetic CIFAR-format folder:
the loss terms must stay finite and the critic cost must move.  Prints the flushed log lines; not collected by pytest.
    python tests/train_smoke.py [iterations]"""
import os
import pickle
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from tests.test_host_utils import write_cifar_dir
from ctgan_b200 import train as T

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
tmp = tempfile.mkdtemp()
data = write_cifar_dir(os.path.join(tmp, 'd'), n_per_file=2000)
out = os.path.join(tmp, 'out')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sess = T.train('cifar_resnet', data, iters=5, out_dir=out, n_examples=10000, acc_every=50)        # capture + first iterations
torch.cuda.synchronize()
e0.record()
sess = None
sess = T.train('cifar_resnet', data, iters=iters, out_dir=out, n_examples=10000, acc_every=50, dev_batches=2)
e1.record()
torch.cuda.synchronize()
log = pickle.load(open(os.path.join(out, 'log.pkl'), 'rb'))
cost = [log['cost'][k] for k in sorted(log['cost'])]
print('logged iterations: %d; cost first/last: %.3f / %.3f; all finite: %s; wall incl. capture, dev cost and samples: %.1f s'
      % (len(cost), cost[0], cost[-1], bool(np.all(np.isfinite(cost))), e0.elapsed_time(e1) * 1e-3))
for name in ('wgan', 'acgan', 'acc_real', 'acc_fake', 'dev_cost', 'time'):
    v = [log[name][k] for k in sorted(log[name])]
    print('%-9s n=%-4d first %.4f last %.4f min %.4f max %.4f' % (name, len(v), v[0], v[-1], min(v), max(v)))
assert np.all(np.isfinite(cost))
