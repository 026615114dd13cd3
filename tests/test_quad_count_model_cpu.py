"""A numpy model of the lane arithmetic of sliced_entry_quad_kernel (pf_sliced.cuh): after the four cooperative loads of
a round lane l holds, for tile l & 3, the masks of k-mers q, 8 + q, 16 + q, 24 + q (q = l >> 2); sl_quad_count adds those
four in place (three bit planes) and reduce-scatters over lane bits 2..4 with full adders, which must leave lane l with the
per-column count, over the round's 32 k-mers, of word sl_word_of_lane<8>(q) of its tile -- and the shuffle that hands a
tile's words to sl_reach (source lane 4 (l & 7) + t) must deliver word sl_word_of_lane<8>(l) to lane l.  The CUDA code is
exercised on the GPU (tests/test_gpu_lines.py); this pins the index arithmetic it relies on, without a GPU."""
import numpy as np
import pytest


def word_of_lane8(lane):  # sl_word_of_lane<8>: the low three lane bits, reversed
    return ((lane & 1) << 2) | (lane & 2) | ((lane >> 2) & 1)


def xor3(a, b, c):
    return a ^ b ^ c


def maj(a, b, c):
    return (a & b) | (c & (a | b))


def quad_count(x):
    """x[lane][j][w] -> planes[lane][6] (word owned by the lane), mirroring sl_quad_count stage by stage."""
    a = np.zeros((32, 6, 8), dtype=np.uint32)
    for lane in range(32):
        for w in range(8):
            x0, x1, x2, x3 = (x[lane, j, w] for j in range(4))
            s1, c1 = xor3(x0, x1, x2), maj(x0, x1, x2)
            c2 = s1 & x3
            a[lane, 0, w], a[lane, 1, w], a[lane, 2, w] = s1 ^ x3, c1 ^ c2, c1 & c2
    for st in range(3):
        P, H, bit = 3 + st, 8 >> (st + 1), 2 + st
        new = a.copy()
        for lane in range(32):
            hi, partner = (lane >> bit) & 1, lane ^ (1 << bit)
            for w in range(H):
                carry = np.uint32(0)
                for pl in range(P):
                    keep = a[lane, pl, w + H] if hi else a[lane, pl, w]
                    recv = a[partner, pl, w] if (partner >> bit) & 1 else a[partner, pl, w + H]  # what the partner sends
                    new[lane, pl, w] = xor3(keep, recv, carry)
                    carry = maj(keep, recv, carry)
                new[lane, P, w] = carry
        a = new
    return a[:, :, 0]


@pytest.mark.parametrize("seed", range(4))
def test_quad_count_matches_popcount_per_column(seed):
    rng = np.random.default_rng(seed)
    density = [0.5, 0.05, 0.95, 1.0][seed]
    masks = np.zeros((32, 4, 8), dtype=np.uint32)  # [k-mer][tile][word]
    bits = rng.random((32, 4, 8, 32)) < density
    for b in range(32):
        masks |= bits[..., b].astype(np.uint32) << np.uint32(b)
    x = np.zeros((32, 4, 8), dtype=np.uint32)
    for lane in range(32):
        for j in range(4):
            x[lane, j] = masks[8 * j + (lane >> 2), lane & 3]  # load j brings k-mer 8 j + q, slot lane & 3
    planes = quad_count(x)
    for lane in range(32):
        tile, word = lane & 3, word_of_lane8(lane >> 2)
        got = sum(((planes[lane, pl].astype(np.int64) >> np.arange(32)) & 1) << pl for pl in range(6))
        want = bits[:, tile, word, :].sum(axis=0)
        assert (got == want).all(), (lane, tile, word)


def test_every_tile_word_has_one_owner_and_reach_gets_its_words():
    owners = {(lane & 3, word_of_lane8(lane >> 2)) for lane in range(32)}
    assert owners == {(t, w) for t in range(4) for w in range(8)}
    for t in range(4):
        for lane in range(32):
            src = 4 * (lane & 7) + t  # pass_t = shfl(alive, 4 (lane & 7) + t)
            assert src & 3 == t and word_of_lane8(src >> 2) == word_of_lane8(lane)


def test_word_block_walk_of_the_fused_hash():
    """Lanes hold the read's 64-bit words wb .. wb + 31; a round needs words wi and wi + 1 with wi = base / 32; the block is
    refreshed when wi reaches a multiple of 31, so both are always inside it."""
    for n_k in (1, 31, 32, 33, 131, 992, 993, 1024, 9981, 69981):
        n_w = ((n_k - 1) >> 5) + 2
        wb_loaded = 0
        for base in range(0, n_k, 32):
            wi = base >> 5
            wb = (wi // 31) * 31
            if wi == wb and wi != 0:
                wb_loaded = wb
            assert wb_loaded == wb and 0 <= wi - wb and wi - wb + 1 <= 31 and wi + 1 < n_w
