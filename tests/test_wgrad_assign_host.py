"""Host logic of the multi-job filter-gradient launch (csrc/conv_wgrad_multi.cu: wg_assign_items, through the C-ABI test entry
ctgan_wgrad_multi_assign): the longest-first assignment of work items to CTAs that the kernel reads as WgradJobTable::order.
Pure host code -- runs without a GPU."""
import ctypes

import numpy as np
import pytest

from ctgan_b200 import _lib


def assign(cost, grid, min_gain_pct=0, rotate=0):
    cost = np.ascontiguousarray(cost, dtype=np.int32)
    order = np.full(1024, -7, dtype=np.int16)
    gain = ctypes.c_int(-1)
    n_slots = _lib.lib.ctgan_wgrad_multi_assign(cost.ctypes.data_as(ctypes.c_void_p), len(cost), grid, min_gain_pct, rotate,
                                                order.ctypes.data_as(ctypes.c_void_p), ctypes.byref(gain))
    return n_slots, order, gain.value


def per_cta(order, n_slots, grid):
    return [[int(order[b + grid * k]) for k in range(n_slots // grid)] for b in range(grid)]


@pytest.mark.parametrize('items,grid,seed', [(296, 148, 0), (450, 148, 1), (37, 12, 2), (600, 148, 3)])
def test_every_item_exactly_once_and_loads_are_balanced(items, grid, seed):
    rs = np.random.RandomState(seed)
    cost = rs.randint(5, 160, items)
    n_slots, order, gain = assign(cost, grid)
    assert n_slots > 0 and n_slots % grid == 0 and n_slots <= 1024
    lists = per_cta(order, n_slots, grid)
    flat = [i for l in lists for i in l if i >= 0]
    assert sorted(flat) == list(range(items))                       # a permutation of the items
    for l in lists:                                                 # the kernel stops at the first negative entry of a list
        seen_pad = False
        for i in l:
            assert not (seen_pad and i >= 0)
            seen_pad = seen_pad or i < 0
    loads = np.array([sum(int(cost[i]) for i in l if i >= 0) for l in lists])
    rr = np.array([int(cost[b::grid].sum()) for b in range(grid)])
    assert loads.max() <= rr.max()                                  # never worse than round robin under the cost model
    assert loads.max() <= cost.sum() / grid + cost.max()            # the list-scheduling bound
    assert gain == int(100 - 100 * int(loads.max()) // int(rr.max()))
    for l in lists:                                                 # longest first within a CTA
        c = [int(cost[i]) for i in l if i >= 0]
        assert c == sorted(c, reverse=True)


def test_round_robin_is_kept_when_it_is_already_balanced_or_the_table_cannot_be_used():
    assert assign(np.full(296, 100), 148, min_gain_pct=10)[0] == 0      # equal items, two per CTA: nothing to gain
    assert assign(np.arange(1, 100), 148)[0] == 0                        # fewer items than CTAs: one each
    assert assign(np.full(1025, 3), 148)[0] == 0                         # more items than the table holds
    assert assign(np.random.RandomState(3).randint(5, 160, 1024), 148)[0] == 0   # 7 slots per CTA x 148 > 1024 entries
    # a few large items among many small ones: one CTA would get 3 large ones under round robin
    cost = np.full(300, 10); cost[[0, 148, 296]] = 200
    n_slots, order, gain = assign(cost, 148, min_gain_pct=10)
    assert n_slots > 0 and gain >= 50


def test_rotation_keeps_each_cta_its_items():
    rs = np.random.RandomState(5)
    cost = rs.randint(5, 160, 400)
    n0, o0, _ = assign(cost, 148, rotate=0)
    n1, o1, _ = assign(cost, 148, rotate=1)
    assert n0 == n1
    a, b = per_cta(o0, n0, 148), per_cta(o1, n1, 148)
    for la, lb in zip(a, b):
        assert sorted(i for i in la if i >= 0) == sorted(i for i in lb if i >= 0)
        k = len([i for i in lb if i >= 0])
        assert all(i >= 0 for i in lb[:k]) and all(i < 0 for i in lb[k:])
    assert a != b


def test_conv_mean_pool_equals_one_stride2_conv_with_the_box_filter():
    """The identity behind kernels.config.pool_conv_s2d (csrc/conv_s2d.cu box_filter_kernel, checked against the kernel in
    tests/test_kernels_gpu.py): mean_pool_2x2(conv3x3_SAME(x, w) + b) == conv4x4_stride2_pad1(x, W4) + b with
    W4[u][v] = 1/4 sum_{a,b in {0,1}} w[u-a][v-b], in float64 on the CPU (TG/CT_gan_cifar_resnet.py:89-92: ConvMeanPool)."""
    import torch
    import torch.nn.functional as TF
    g = torch.Generator().manual_seed(0)
    C, O = 5, 7
    w = torch.randn(3, 3, C, O, generator=g, dtype=torch.float64)          # HWIO, as tflib stores filters
    b = torch.randn(O, generator=g, dtype=torch.float64)
    x = torch.randn(3, C, 12, 10, generator=g, dtype=torch.float64)
    w4 = torch.zeros(4, 4, C, O, dtype=torch.float64)
    for u in range(4):
        for v in range(4):
            for a in (0, 1):
                for c in (0, 1):
                    if 0 <= u - a < 3 and 0 <= v - c < 3:
                        w4[u, v] += 0.25 * w[u - a, v - c]
    ref = TF.avg_pool2d(TF.conv2d(x, w.permute(3, 2, 0, 1), b, padding=1), 2)
    got = TF.conv2d(x, w4.permute(3, 2, 0, 1), b, stride=2, padding=1)
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)
    # 'SAME' padding of a 4x4 / stride-2 conv on an even extent is exactly pad 1 (kernels.same_geom)
    import ctgan_b200.kernels as K
    geo = K.same_geom(3, 12, 10, C, O, 4, 2)
    assert (geo.pad_t, geo.pad_l, geo.Ho, geo.Wo) == (1, 1, 6, 5)
    # live (tap, phase) blocks of the embedded 3x3 filter: 16 of 36 for k = 4, 25 of 36 for k = 5 (CTGAN_EPI_S2D_SKIP)
    for k, pad, live in ((4, 1, 16), (5, 1, 25), (5, 2, 25)):
        n = sum(1 for R in range(3) for S in range(3) for dy in (0, 1) for dx in (0, 1)
                if 0 <= 2 * (R - 1) + dy + pad < k and 0 <= 2 * (S - 1) + dx + pad < k)
        assert n == live, (k, pad, n)
