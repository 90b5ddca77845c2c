"""Host-side planner of the fused MFCC kernel's filter-bank stage (csrc/mfcc_plan.cu; no GPU needed).

The plan must (1) cover every non-zero weight of every filter exactly once -- checked by replaying the kernel's two
stages in numpy against the dense ``amp @ H`` of fbank.py:315-316 -- and (2) keep the 64-bit shared-memory loads of a
half-warp in distinct banks, which is what it exists for (round 1: 25 % of the kernel's wavefronts were conflicts).
"""

import ctypes as C

import numpy as np
import pytest

from diffsptk_b200 import tables

SEG, MAXSEG, PER_CH, PITCH = 8, 128, 8, 264


def build_plan(native_lib, cb, ce, K):
    n = native_lib.dsb200_mfcc_plan_ints(len(cb))
    assert n == 4 + 4 * MAXSEG + len(cb) * PER_CH
    plan = np.zeros(n, dtype=np.int32)
    cb = np.ascontiguousarray(cb, dtype=np.int32)
    ce = np.ascontiguousarray(ce, dtype=np.int32)
    p = C.POINTER(C.c_int32)
    rc = native_lib.dsb200_mfcc_plan_build(cb.ctypes.data_as(p), ce.ctypes.data_as(p), len(cb), K, plan.ctypes.data_as(p))
    assert rc == 0
    return plan


def replay(plan, H, amp):
    """The kernel's filter-bank stage on one amplitude row (numpy): segment sums, then per-channel sums."""
    K, Cn = H.shape
    ns = plan[0]
    start, first, last, ch = (plan[4 + i * MAXSEG: 4 + (i + 1) * MAXSEG] for i in range(4))
    slots = plan[4 + 4 * MAXSEG:].reshape(Cn, PER_CH)
    row = np.zeros(PITCH)
    row[:K] = amp
    segsum = np.zeros(MAXSEG + 1)
    for s in range(ns):
        for i in range(SEG):
            k = start[s] + i
            w = H[k, ch[s]] if (ch[s] >= 0 and first[s] <= k < last[s]) else 0.0
            segsum[s] += w * row[k]
    return np.array([segsum[slots[c]].sum() for c in range(Cn)])


def bank_conflicts(plan):
    """Extra wavefronts of the half-warp 64-bit loads over all rounds and steps."""
    ns = plan[0]
    start = plan[4:4 + MAXSEG]
    extra = 0
    for g in range(ns // 16):
        s = start[16 * g:16 * g + 16]
        for i in range(SEG):
            banks = {}
            for v in s + i:
                banks.setdefault(v % 16, set()).add(v)
            extra += max(len(a) for a in banks.values()) - 1
    return extra


@pytest.mark.parametrize("fft_length,n_channel,sr", [(512, 40, 16000), (512, 24, 16000), (512, 80, 16000),
                                                     (512, 40, 8000), (512, 20, 48000)])
def test_plan_covers_every_weight_once_and_avoids_conflicts(native_lib, fft_length, n_channel, sr):
    H = tables.make_fbank_matrix(fft_length, n_channel, sr).double().numpy()
    K = H.shape[0]
    cb, ce = (t.numpy() for t in tables.column_support(tables.make_fbank_matrix(fft_length, n_channel, sr)))
    plan = build_plan(native_lib, cb, ce, K)
    if plan[0] == 0:
        # only when the pieces cannot fit
        assert sum(-(-(e - b) // SEG) for b, e in zip(cb, ce)) > MAXSEG or max(ce - cb) > SEG * PER_CH
        return
    assert plan[0] % 32 == 0 and plan[1] == n_channel
    amp = np.random.default_rng(0).uniform(0.1, 2.0, K)
    np.testing.assert_allclose(replay(plan, H, amp), amp @ H, rtol=1e-12, atol=1e-12)
    assert bank_conflicts(plan) == plan[2] == 0


def test_in_order_cut_of_round_1_has_the_conflicts_the_plan_removes(native_lib):
    """The yardstick: cutting the supports in order (what the kernel does without a plan) collides on this table."""
    Ht = tables.make_fbank_matrix(512, 40, 16000)
    cb, ce = (t.numpy() for t in tables.column_support(Ht))
    starts = [k for b, e in zip(cb, ce) for k in range(b, e, SEG)]
    naive = np.zeros(4 + 4 * MAXSEG, dtype=np.int32)
    naive[0] = -(-len(starts) // 32) * 32
    naive[4:4 + len(starts)] = starts
    assert bank_conflicts(naive) > 50
    assert bank_conflicts(build_plan(native_lib, cb, ce, 257)) == 0


def test_degenerate_supports(native_lib):
    # empty filters and a filter that is too long for eight segments
    plan = build_plan(native_lib, [0, 5, 5], [0, 9, 5], 257)
    assert plan[0] == 32 and (plan[4 + 4 * MAXSEG:].reshape(3, PER_CH)[0] == MAXSEG).all()
    assert build_plan(native_lib, [0], [70], 257)[0] == 0
    assert build_plan(native_lib, [10], [5], 257)[0] == 0
