"""Replay time of each captured graph of the ResNet iteration (CUDA events, 20 replays each, inputs resident)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import ctgan_b200.gan_cifar_resnet as R
from ctgan_b200.graphs import GraphedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
import ctgan_b200.kernels as K
if len(sys.argv) > 2:
    K.config.branch_priority = int(sys.argv[2])
if len(sys.argv) > 3:
    K.config.side_stream = K.config.branch_streams = bool(int(sys.argv[3]))
if len(sys.argv) > 4:
    from ctgan_b200 import _lib
    _lib.lib.ctgan_set_pdl(int(sys.argv[4]))
import os
if os.environ.get('CTGAN_FUSE_RELU_BWD'):
    R.FUSE_RELU_BWD = bool(int(os.environ['CTGAN_FUSE_RELU_BWD']))
if os.environ.get('CTGAN_CRITIC_SPLITK'):
    K.config.critic_splitk = bool(int(os.environ['CTGAN_CRITIC_SPLITK']))
if os.environ.get('CTGAN_BRANCH_STACKED'):
    K.config.branch_stacked = bool(int(os.environ['CTGAN_BRANCH_STACKED']))
if os.environ.get('CTGAN_SPLITK'):
    from ctgan_b200 import _lib as _L2
    _L2.lib.ctgan_set_splitk(int(os.environ['CTGAN_SPLITK']))
if os.environ.get('CTGAN_GEN_TOWERS'):
    K.config.gen_towers = bool(int(os.environ['CTGAN_GEN_TOWERS']))
if os.environ.get('CTGAN_GEN_SPLITK'):
    K.config.gen_splitk = bool(int(os.environ['CTGAN_GEN_SPLITK']))
if os.environ.get('CTGAN_DECOUPLE_GP'):
    K.config.decouple_gp = bool(int(os.environ['CTGAN_DECOUPLE_GP']))
if os.environ.get('CTGAN_POOL_CONV_TILES'):
    K.config.pool_conv_min_tiles = int(os.environ['CTGAN_POOL_CONV_TILES'])
if os.environ.get('CTGAN_POOL_CONV'):
    K.config.pool_conv_s2d = bool(int(os.environ['CTGAN_POOL_CONV']))
if os.environ.get('CTGAN_S2D_SKIP'):
    K.config.s2d_skip = bool(int(os.environ['CTGAN_S2D_SKIP']))
if os.environ.get('CTGAN_NORES'):
    from ctgan_b200 import _lib as _L5
    _L5.lib.ctgan_set_fprop_nores(int(os.environ['CTGAN_NORES']))
if os.environ.get('CTGAN_WGRAD_BALANCE'):
    from ctgan_b200 import _lib as _L6
    _b = [int(v) for v in os.environ['CTGAN_WGRAD_BALANCE'].split(',')]
    _L6.lib.ctgan_set_wgrad_multi_balance(_b[0], _b[1] if len(_b) > 1 else -1)
if os.environ.get('CTGAN_WGRAD_CHUNK'):
    from ctgan_b200 import _lib as _L4
    _L4.lib.ctgan_set_wgrad_multi_chunk(int(os.environ['CTGAN_WGRAD_CHUNK']))
if os.environ.get('CTGAN_HALO'):
    from ctgan_b200 import _lib as _L3
    _L3.lib.ctgan_set_fprop_halo(int(os.environ['CTGAN_HALO']))
if os.environ.get('CTGAN_WGRAD_ITEMS'):
    from ctgan_b200 import _lib as _L
    _L.lib.ctgan_set_wgrad_multi_items_per_sm(int(os.environ['CTGAN_WGRAD_ITEMS']))
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randint(0, 256, (B, 3072)).astype('int32')).cuda()
y = torch.from_numpy(rs.randint(0, 10, (B,)).astype('int32')).cuda()
gt = GraphedTrainer(tr, (x, y), pregen_steps=5)
gt.begin_iteration(y.repeat(5))

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def critic():
    gt._k = 0
    gt.critic_step(x, y)
print('critic graph  %8.1f us  (%d kernels)' % (t(critic), gt.critic_kernels))
print('gen graph     %8.1f us  (%d kernels)' % (t(gt.gen_step), gt.gen_kernels))
print('pregen graph  %8.1f us  (%d kernels)' % (t(lambda: gt.begin_iteration(y.repeat(5))), gt.pregen_kernels))
