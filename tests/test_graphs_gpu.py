"""GPU tests of the CUDA-graph / stream-branch execution of the training step (graphs.GraphedTrainer): replaying the
captured graphs must train exactly like launching the same steps eagerly, with and without the batched pre-generation of
the fake batches, and the stream branches must not change a step's result."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(seed=11, B=16, dtype=torch.bfloat16):
    import ctgan_b200.gan_cifar_resnet as R
    np.random.seed(1234)
    return R.Trainer(device='cuda', seed=seed, act_dtype=dtype, batch_size=B, graph_safe_rng=True)


def _batches(n, B, seed=5):
    rs = np.random.RandomState(seed)
    xs = torch.from_numpy(rs.randint(0, 256, (n, B, 3072)).astype('int32')).cuda()
    ys = torch.from_numpy(rs.randint(0, 10, (n, B)).astype('int32')).cuda()
    return xs, ys


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_generate_fakes_matches_per_step_generation():
    """Rows of step 0 of a batched S-step generator forward == the S=1 forward from the same noise (batch-norm
    statistics are per 32-sample split), and a critic step fed those fakes == the step generating its own."""
    B = 16
    xs, ys = _batches(3, B)
    # one model per process (tflib's module-level parameter dict): build, run and read each trainer in turn
    a = _trainer(B=B); one = a.generate_fakes(ys[0]).clone()
    b = _trainer(B=B); three = b.generate_fakes(ys.reshape(-1)).clone()
    assert three.shape == (3 * B, 3072)
    # not bit-equal: the batch-norm sums are accumulated with red.global (order varies with the grid), which moves a
    # few bf16 roundings of the activations
    assert _rel(three[:B], one) < 2e-2
    assert _rel(three[B:2 * B], three[:B]) > 0.1
    c = _trainer(B=B); c.disc_opt.zero_grad(); r1 = c.critic_forward_backward(xs[0], ys[0])
    g1 = c.disc_opt.flat_g.clone()
    d = _trainer(B=B); fake = d.generate_fakes(ys[0]); d.disc_opt.zero_grad()
    r2 = d.critic_forward_backward(xs[0], ys[0], fake_data=fake)
    g2 = d.disc_opt.flat_g.clone()
    assert _rel(r1['fake_data'], fake) < 2e-2
    assert _rel(r2['out'][:5], r1['out'][:5]) < 2e-2
    assert _rel(g2, g1) < 6e-2       # bf16 ReLU-pattern flips downstream of the few re-rounded fake pixels


@pytest.fixture(params=[False, True], ids=['default', 'pool_conv'])
def pool_conv(request):
    """Second setting: ConvMeanPool(3x3) as one stride-2 conv in every pass (kernels.config.pool_conv_s2d) -- its derived 4x4
    filters must be recomputed / re-packed in place by every replayed optimizer step and their scratch gradients folded."""
    import ctgan_b200.kernels as K
    saved = (K.config.pool_conv_s2d, K.config.pool_conv_min_tiles)
    if request.param:
        K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = True, 1
    yield request.param
    K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = saved


@pytest.mark.parametrize('pregen', [0, 2])
def test_graph_replay_trains_like_eager(pregen, pool_conv):
    """2 iterations (1 generator step + 2 critic steps): GraphedTrainer replays vs the same sequence launched eagerly from
    the same seeds (differences: atomic accumulation order only).  The capture's warm-up steps leave no trace: weights,
    Adam state, step counts and the Philox counters are restored, so both start from the initial model."""
    from ctgan_b200.graphs import GraphedTrainer
    B, NC = 16, 2
    xs, ys = _batches(1 + 2 * NC, B)
    graphed = _trainer(B=B)
    p_init = graphed.disc_opt.flat_p.clone()
    gt = GraphedTrainer(graphed, (xs[0], ys[0]), warmup=3, pregen_steps=pregen)
    assert torch.equal(graphed.disc_opt.flat_p, p_init) and graphed.disc_opt.t == 0 and graphed.gen_opt.t == 0
    assert float(graphed.disc_opt.flat_m.abs().max()) == 0.0 and int(graphed.rng.dyn.item()) == 0
    for it in range(2):
        gt.iteration = it
        gt.gen_step()
        if pregen:
            gt.begin_iteration(ys[1 + it * NC:1 + (it + 1) * NC])
        for k in range(NC):
            out_g = gt.critic_step(xs[1 + it * NC + k], ys[1 + it * NC + k]).clone()
    torch.cuda.synchronize()
    pd_g, pg_g = graphed.disc_opt.flat_p.clone(), graphed.gen_opt.flat_p.clone()

    eager = _trainer(B=B)
    pd0, pg0 = eager.disc_opt.flat_p.clone(), eager.gen_opt.flat_p.clone()
    for it in range(2):
        eager.gen_opt.set_device_lr(eager.lr(it)); eager.gen_step(use_device_lr=True)
        fakes = None
        if pregen:
            fakes = eager.generate_fakes(ys[1 + it * NC:1 + (it + 1) * NC].reshape(-1)); eager.rng.end_step()
        for k in range(NC):
            eager.disc_opt.set_device_lr(eager.lr(it))
            f = fakes[k * B:(k + 1) * B] if pregen else None
            out_e = eager.critic_step(xs[1 + it * NC + k], ys[1 + it * NC + k], use_device_lr=True, fake_data=f)['out']
    torch.cuda.synchronize()
    pd_e, pg_e = eager.disc_opt.flat_p.clone(), eager.gen_opt.flat_p.clone()
    print('update-relative differences: D %.3e  G %.3e' % (_rel(pd_g - pd0, pd_e - pd0), _rel(pg_g - pg0, pg_e - pg0)))
    # GAN dynamics amplify the bf16 / atomic-order noise of two runs to a few percent after 4 critic updates (two eager
    # runs differ by as much); a wrong learning rate, random stream or stale buffer gives uncorrelated updates (relative difference > 1)
    print('last critic step, graph vs eager:', out_g[:5].tolist(), out_e[:5].tolist())
    assert _rel(out_g[:5], out_e[:5]) < 0.1
    # the accumulated UPDATES (4 critic / 2 generator Adam steps) agree; bf16 activation-pattern flips from a different
    # atomic accumulation order move individual sign-like Adam updates, so this is a statistical bound
    assert _rel(pd_g - pd0, pd_e - pd0) < 0.5
    assert _rel(pg_g - pg0, pg_e - pg0) < 0.5


def test_stream_branches_do_not_change_a_step():
    """Side-stream wgrad + gradient-penalty branch vs everything on one stream: same losses and gradients."""
    import ctgan_b200.kernels as K
    B = 16
    xs, ys = _batches(1, B)
    res = {}
    for on in (True, False):
        K.config.side_stream = K.config.branch_streams = on
        try:
            tr = _trainer(B=B)
            tr.disc_opt.zero_grad()
            out = tr.critic_forward_backward(xs[0], ys[0])['out']
            tr.gen_opt.zero_grad()
            tr.gen_forward_backward()
            torch.cuda.synchronize()
            res[on] = (out.clone(), tr.disc_opt.flat_g.clone(), tr.gen_opt.flat_g.clone())
        finally:
            K.config.side_stream = K.config.branch_streams = True
    # not bit-equal: batch-norm sums and filter gradients are accumulated with red.global (order varies run to run)
    assert _rel(res[True][0][:5], res[False][0][:5]) < 1e-3
    assert _rel(res[True][1], res[False][1]) < 6e-2
    assert _rel(res[True][2], res[False][2]) < 6e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_generator_towers_match_the_stacked_batch(dtype):
    """kernels.config.gen_towers: the generator step as two stream-branch towers (the reference's per-device graph) draws the
    same random numbers (DeviceRandom.scope_tower) and computes the same cost and gradients as the one stacked batch."""
    import ctgan_b200.kernels as K
    res = {}
    for towers in (True, False):
        K.config.gen_towers = towers
        try:
            tr = _trainer(B=16, dtype=dtype)
            tr.gen_opt.zero_grad()
            cost = tr.gen_forward_backward()['cost']
            torch.cuda.synchronize()
            res[towers] = (cost.clone(), tr.gen_opt.flat_g.clone(), tr.rng.offset)
        finally:
            K.config.gen_towers = False
    assert res[True][2] == res[False][2]                      # same Philox slices consumed
    assert _rel(res[True][0], res[False][0]) < 1e-3
    # fp32: the same arithmetic in another order.  bf16: batch-norm statistics rounded differently flip a few ReLU patterns at
    # B=16 (the bound test_stream_branches_do_not_change_a_step uses for two runs of the SAME schedule)
    assert _rel(res[True][1], res[False][1]) < (1e-3 if dtype == torch.float32 else 6e-2)


def test_dcgan_graph_replay_trains_like_eager():
    """CT_gan_cifar.py with its stride-2 layers on the space-to-depth tensor-core route: the operand packs of those
    filters are created at first use and re-packed in place after every optimizer step (kernels._s2d_packs), so a
    captured graph must keep reading current weights.  2 x (generator step + 2 critic steps),
    graph replay vs the same sequence launched eagerly from the same seeds."""
    import ctgan_b200.gan_cifar as C
    import ctgan_b200.kernels as K
    from ctgan_b200.graphs import GraphedTrainer
    B, NC = 16, 2
    rs = np.random.RandomState(5)
    xs = torch.from_numpy(rs.randint(0, 256, (1 + 2 * NC, B, 3072)).astype('int32')).cuda()

    def trainer():
        np.random.seed(1234)
        return C.Trainer(device='cuda', seed=11, act_dtype=torch.bfloat16, batch_size=B, graph_safe_rng=True)

    graphed = trainer()
    pd0, pg0 = graphed.disc_opt.flat_p.clone(), graphed.gen_opt.flat_p.clone()
    gt = GraphedTrainer(graphed, (xs[0],), warmup=3)
    assert K._lazy_packs, 'the space-to-depth route was not taken'
    for it in range(2):
        gt.gen_step()
        for k in range(NC):
            out_g = gt.critic_step(xs[1 + it * NC + k]).clone()
    torch.cuda.synchronize()
    pd_g, pg_g = graphed.disc_opt.flat_p.clone(), graphed.gen_opt.flat_p.clone()

    eager = trainer()
    for it in range(2):
        eager.gen_opt.set_device_lr(None); eager.gen_step(use_device_lr=True)
        for k in range(NC):
            eager.disc_opt.set_device_lr(None)
            out_e = eager.critic_step(xs[1 + it * NC + k], use_device_lr=True)['out']
    torch.cuda.synchronize()
    pd_e, pg_e = eager.disc_opt.flat_p.clone(), eager.gen_opt.flat_p.clone()
    print('update-relative differences: D %.3e  G %.3e' % (_rel(pd_g - pd0, pd_e - pd0), _rel(pg_g - pg0, pg_e - pg0)))
    print('last critic step, graph vs eager:', out_g[:4].tolist(), out_e[:4].tolist())
    assert _rel(out_g[:4], out_e[:4]) < 0.1
    # stale operand packs would freeze the critic at its initial weights inside the graph: relative difference ~1 here
    assert _rel(pd_g - pd0, pd_e - pd0) < 0.5
    assert _rel(pg_g - pg0, pg_e - pg0) < 0.5
