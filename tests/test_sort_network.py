"""The 96-word merge-split network of the segmented sort (csrc/segsort.cuh: sort3, warp_mergesplit96), stated in numpy
lane for lane — 32 lanes, a sorted triple per lane, the 15 stages of the 32-lane bitonic network as merge-splits — and
checked to sort. This is the model the device function was written from; on the GPU the function itself is covered by
the sorted_key / sorted_index taps of every parity test whose scene has paths of 65..96 fragments (the synthetic scenes:
40 % of their paths)."""
import numpy as np


def sort3(a, b, c):
    mn = np.minimum(np.minimum(a, b), c)          # VIMNMX3
    mx = np.maximum(np.maximum(a, b), c)
    return mn, a ^ b ^ c ^ mn ^ mx, mx            # the middle by exclusive-or (a multiset identity: ties are fine)


def mergesplit96(words):
    """words uint32 [96] in the striped order the kernel loads them (word r * 32 + lane in register r of the lane);
    returns them in the blocked order the network leaves them (lane L: ranks 3L, 3L+1, 3L+2)."""
    lane = np.arange(32)
    e = list(sort3(*words.reshape(3, 32)))
    k = 2
    while k <= 32:
        j = k >> 1
        while j > 0:
            keep_min = ((lane & k) == 0) == ((lane & j) == 0)
            partner = lane ^ j                     # __shfl_xor_sync: the partner's triple, reversed
            got = (e[2][partner], e[1][partner], e[0][partner])
            e = list(sort3(*[np.where(keep_min, np.minimum(x, y), np.maximum(x, y)) for x, y in zip(e, got)]))
            j >>= 1
        k <<= 1
    out = np.empty(96, np.uint32)
    for r in range(3):
        out[3 * lane + r] = e[r]
    return out


def packed(keys):
    """(row|x relative to the path minimum) above the fragment's position, as warp_sort_segment32 packs them; the pad
    words of a path shorter than 96 are all ones."""
    n = len(keys)
    w = np.full(96, 0xFFFFFFFF, np.uint32)
    w[:n] = (keys.astype(np.uint32) << np.uint32(7)) | np.arange(n, dtype=np.uint32)
    return w


def test_sorts_random_paths_with_ties():
    rng = np.random.default_rng(5)
    for t in range(1500):
        n = int(rng.integers(65, 97))
        keys = rng.integers(0, rng.choice([3, 50, 1 << 20]), n)
        w = packed(keys)
        assert np.array_equal(mergesplit96(w), np.sort(w))


def test_sorts_ordered_reversed_and_constant_input():
    for n in (65, 80, 96):
        for keys in (np.arange(n), np.arange(n)[::-1], np.zeros(n, np.int64), np.arange(n) % 2):
            w = packed(keys)
            assert np.array_equal(mergesplit96(w), np.sort(w))


def test_order_is_key_then_position():
    """Equal keys keep their fragment order: the reference's (key, index) comparison (naive_seg_sort_pairs.comp:69-75)."""
    keys = np.array([5, 1, 5, 1, 5] * 16)
    out = mergesplit96(packed(keys))[:80]
    pos = out & np.uint32(127)
    assert np.array_equal(out >> np.uint32(7), np.sort(keys))
    assert np.all(np.diff(pos[:32].astype(int)) > 0) and np.all(np.diff(pos[32:].astype(int)) > 0)


def test_zero_one_principle_on_block_boundaries():
    """0-1 inputs with every number of ones (the zero-one principle restricted to one permutation family per count,
    plus random placements)."""
    rng = np.random.default_rng(9)
    for ones in range(97):
        base = np.zeros(96, np.uint32)
        base[:ones] = 1
        for _ in range(6):
            w = rng.permutation(base)
            assert np.array_equal(mergesplit96(w), np.sort(w))
