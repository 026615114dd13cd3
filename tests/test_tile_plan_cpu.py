"""The tiling of the bit-sliced path (pf_sliced.cu: plan_tiles / tile_tree) is host-only logic: checked here without a GPU
through pf_plan_tiles on synthetic level-ordered trees.  Structural invariants that the kernels rely on:
every node is either skipped or owns exactly one (tile, column); a column's parent comes before it in the same tile, or sits
in the tile's parent tile (all roots of a tile hang under ONE parent tile), or is skipped (entry tile); skipped nodes are
interior, connected to the root, and have a fully verified subtree; widths are 32/64/128/256 and fit the columns."""
import ctypes as C

import numpy as np
import pytest

from phagefilter_b200 import _lib

NONE = 0xFFFFFFFF


def random_tree(rng, n_leaves, unbalanced):
    """Level-ordered binary tree grown by splitting random leaves; returns left, right, parent, leaf (DFS index)."""
    children = {0: None}
    leaves = [0]
    nxt = 1
    while len(leaves) < n_leaves:
        i = int(rng.integers(0, len(leaves))) if not unbalanced else min(len(leaves) - 1, int(rng.exponential(1.5)))
        u = leaves.pop(i)
        children[u] = (nxt, nxt + 1)
        children[nxt] = children[nxt + 1] = None
        leaves += [nxt, nxt + 1]
        nxt += 2
    # renumber in level order
    order, q = [], [0]
    while q:
        nq = []
        for u in q:
            order.append(u)
            if children[u]:
                nq += list(children[u])
        q = nq
    new = {u: i for i, u in enumerate(order)}
    n = len(order)
    left = np.full(n, NONE, dtype=np.uint32)
    right = np.full(n, NONE, dtype=np.uint32)
    parent = np.full(n, -1, dtype=np.int64)
    for u in order:
        if children[u]:
            left[new[u]], right[new[u]] = new[children[u][0]], new[children[u][1]]
            parent[new[children[u][0]]] = parent[new[children[u][1]]] = new[u]
    leaf = np.full(n, -1, dtype=np.int32)
    st, k = [0], 0
    while st:
        u = st.pop()
        if left[u] == NONE:
            leaf[u] = k
            k += 1
        else:
            st += [int(right[u]), int(left[u])]
    return left, right, parent, leaf


def plan(left, right, leaf, pop, mono, m, K, theta, n_k, handover):
    L = _lib.lib()
    n = len(left)
    skip = np.zeros(n, dtype=np.uint8)
    tile = np.zeros(n, dtype=np.int32)
    col = np.zeros(n, dtype=np.uint32)
    cap = n + 1
    tparent = np.zeros(cap, dtype=np.int32)
    twidth = np.zeros(cap, dtype=np.uint32)
    nt, ne = C.c_uint64(0), C.c_uint64(0)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _lib.check(L.pf_plan_tiles(n, p(left, C.c_uint32), p(right, C.c_uint32), p(leaf, C.c_int32), p(pop, C.c_uint64),
                               p(mono, C.c_uint8), m, K, C.c_float(theta), n_k, handover, p(skip, C.c_uint8), p(tile, C.c_int32),
                               p(col, C.c_uint32), p(tparent, C.c_int32), p(twidth, C.c_uint32), cap, C.byref(nt), C.byref(ne)))
    return skip.astype(bool), tile, col, tparent[: nt.value], twidth[: nt.value], int(ne.value)


@pytest.mark.parametrize("seed", range(12))
def test_tiling_invariants(seed):
    rng = np.random.default_rng(seed)
    n_leaves = int(rng.choice([1, 2, 7, 100, 700, 3000]))
    left, right, parent, leaf = random_tree(rng, n_leaves, unbalanced=bool(seed % 2))
    n = len(left)
    m, K = 14_377_587, 10
    # fills like a gSBT's: grow with the leaves below a node
    below = np.zeros(n, dtype=np.int64)
    for u in range(n - 1, -1, -1):
        below[u] = 1 if left[u] == NONE else below[left[u]] + below[right[u]]
    pop = (m * (1.0 - np.exp(-K * below * 50_000 / m))).astype(np.uint64)
    mono = (rng.random(n) < (1.0 if seed % 3 else 0.97)).astype(np.uint8)
    mono[left == NONE] = 0
    theta = float(rng.choice([0.3, 0.8, 1.0]))
    for handover in (0, 1):
        skip, tile, col, tparent, twidth, n_entry = plan(left, right, leaf, pop, mono, m, K, theta, 131, handover)
        n_tiles = len(tparent)
        assert n_entry >= 1 and (tparent[:] >= -1).all() and int((tparent == -1).sum()) == n_entry
        # verified-subtree flag, bottom-up
        vb = np.ones(n, dtype=bool)
        for u in range(n - 1, -1, -1):
            if left[u] != NONE:
                vb[u] = bool(mono[u]) and vb[left[u]] and vb[right[u]]
        seen = set()
        for u in range(n):
            if skip[u]:
                assert left[u] != NONE and vb[u], "skipped: interior with a verified subtree"
                assert u == 0 or skip[parent[u]], "skipped nodes are connected to the root"
                if handover == 0:
                    assert tile[u] == -1
                continue
            if handover == 1 and tile[u] < 0:
                # tiles for the cut only: everything else lies below a column of an entry tile
                a = parent[u]
                while a >= 0 and tile[a] < 0:
                    assert not skip[a]
                    a = parent[a]
                assert a >= 0 and tparent[tile[a]] == -1
                continue
            t, c = int(tile[u]), int(col[u])
            assert 0 <= t < n_tiles and 0 <= c < twidth[t] <= 256 and twidth[t] in (32, 64, 128, 256)
            assert (t, c) not in seen
            seen.add((t, c))
            p = int(parent[u])
            if p < 0 or skip[p]:
                assert tparent[t] == -1, "below a skipped node (or the root itself): entry tile"
            elif tile[p] == t:
                assert col[p] < c, "parents come first inside a tile"
            else:
                assert tile[p] == tparent[t], "all roots of a tile hang under one parent tile"
        if handover == 0:
            assert len(seen) == int((~skip).sum())
        # tile tree is ordered: a tile's parent has a smaller index
        assert all(tparent[t] < t for t in range(n_tiles))
