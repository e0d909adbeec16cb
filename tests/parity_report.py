"""Prints the measured parity of every script / precision path / comparison mode (GPU).
Usage: python tests/parity_report.py [B_resnet]   -> table on stdout (kept under profiles/).
Columns: worst loss term, GP gradient, the whole parameter gradient as one vector (gradall), median and worst
single parameter tensor (floor: see tests/test_step_parity_gpu.py), Adam update."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import torch

from tests import parity
from tests.test_step_parity_gpu import TOL, _path


def summarize(rep):
    g = sorted(v for k, v in rep.items() if k.startswith('grad.'))
    return dict(loss=parity.worst(rep, 'loss.')[0], gp_grad=rep.get('gp_gradient', float('nan')), gradall=rep['gradall'],
                grad_med=g[len(g) // 2], grad_max=g[-1], adam=parity.worst(rep, 'adam.')[0])


def main():
    sizes = {'mnist': 50, 'cifar': 64, 'resnet': int(sys.argv[1]) if len(sys.argv) > 1 else 16}
    print('%-7s %-5s %-5s %-7s | %9s %9s %9s %9s %9s %9s' % ('script', 'B', 'path', 'mode', 'loss', 'gp_grad', 'gradall', 'grad_med',
                                                           'grad_max', 'adam'))
    for script, B in sizes.items():
        for path in ('fp32', 'tf32', 'bf16'):
            if path == 'tf32' and script != 'resnet':
                continue
            for cond in (True, False):
                with _path(path) as dtype:
                    tr, om = parity.build_pair(script, 'cuda', dtype, B, oracle_dtype=torch.float32 if B >= 64 and script == 'resnet' else torch.float64)
                    parity.perturb_params(tr, om)
                    for what in ('critic', 'gen'):
                        ff = TOL[path]['floor']
                        if what == 'critic':
                            rep = parity.critic_parity(script, tr, om, parity.make_inputs(script, B, 11), conditioned=cond, floor_frac=ff)
                        else:
                            rep = parity.gen_parity(script, tr, om, conditioned=cond, floor_frac=ff)
                        s = summarize(rep)
                        print('%-7s %-5d %-5s %-7s | %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e  %s' % (
                            script, B, path, ('cond' if cond else 'indep') + '/' + what[0], s['loss'], s['gp_grad'], s['gradall'],
                            s['grad_med'], s['grad_max'], s['adam'], parity.worst(rep, 'grad.')[1]), flush=True)


if __name__ == '__main__':
    main()
