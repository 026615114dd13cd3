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


def layout(left, right, leaf, pop, mono, m, K, theta, n_k, handover, tile_cols):
    L = _lib.lib()
    n = len(left)
    cap = n + 1
    off = np.zeros(cap, dtype=np.uint64)
    stride = np.zeros(cap, dtype=np.uint32)
    words = np.zeros(cap, dtype=np.uint32)
    group = np.zeros(cap, dtype=np.int32)
    pre = np.zeros(cap, dtype=np.uint32)
    fo = np.zeros(cap, dtype=np.uint8)
    order = np.zeros(cap, dtype=np.uint32)
    nt, ne, nl, tw, rows = (C.c_uint64(0) for _ in range(5))
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _lib.check(L.pf_plan_tile_layout(n, p(left, C.c_uint32), p(right, C.c_uint32), p(leaf, C.c_int32), p(pop, C.c_uint64),
                                     p(mono, C.c_uint8), m, K, C.c_float(theta), n_k, handover, tile_cols, cap, p(off, C.c_uint64),
                                     p(stride, C.c_uint32), p(words, C.c_uint32), p(group, C.c_int32), p(pre, C.c_uint32),
                                     p(fo, C.c_uint8), p(order, C.c_uint32), C.byref(nt), C.byref(ne), C.byref(nl), C.byref(tw),
                                     C.byref(rows)))
    t = int(nt.value)
    return dict(off=off[:t], stride=stride[:t], words=words[:t], group=group[:t], pre=pre[:t], filter_only=fo[:t],
                order=order[: int(ne.value)], n_line=int(nl.value), table_words=int(tw.value), rows=int(rows.value))


LINE_CASES = []


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("tile_cols,force_g", [(32, None), (256, None), (32, 3), (32, 5), (64, 4), (128, 5), (256, 6)])
def test_table_layout_invariants(seed, tile_cols, force_g, monkeypatch):
    """Where the tables lie (layout_tables): entry tiles that share 128-byte lines come first in entry order, four to a group,
    each in its own 32-byte slot of a 128-byte-aligned table with rows 32 words apart; every other tile has a table of its
    own with rows row_words apart; no two tables overlap and together they take exactly table_words."""
    if force_g is not None:  # a given cut ("skip every verified node with more than G leaves"): many entry tiles
        monkeypatch.setenv("PF_SLICED_FORCE_G", str(force_g))
    rng = np.random.default_rng(100 + seed)
    n_leaves = int(rng.choice([40, 300, 1500, 4000] if force_g is None else [600, 1500, 4000]))
    left, right, parent, leaf = random_tree(rng, n_leaves, unbalanced=bool(seed % 2))
    n = len(left)
    m, K = 1_000_003, 7
    below = np.zeros(n, dtype=np.int64)
    for u in range(n - 1, -1, -1):
        below[u] = 1 if left[u] == NONE else below[left[u]] + below[right[u]]
    pop = (m * (1.0 - np.exp(-K * below * 3_000 / m))).astype(np.uint64)
    # an unverified node anywhere keeps the whole top of the tree in play (cut at the root): only without forced cuts
    mono = (rng.random(n) < (1.0 if seed % 3 or force_g is not None else 0.98)).astype(np.uint8)
    mono[left == NONE] = 0
    theta = float(rng.choice([0.5, 0.8, 1.0]))
    for handover in (0, 1):
        lay = layout(left, right, leaf, pop, mono, m, K, theta, 131, handover, tile_cols)
        skip, tile, col, tparent, twidth, n_entry = plan(left, right, leaf, pop, mono, m, K, theta, 131, handover) \
            if tile_cols == 256 else (None,) * 6
        rows, nl = lay["rows"], lay["n_line"]
        LINE_CASES.append((seed, tile_cols, force_g, handover, len(lay["order"]), nl))
        assert rows % 1024 == 0 and rows >= m
        order = lay["order"]
        assert len(set(order.tolist())) == len(order)
        if tile_cols == 256:
            assert len(order) == n_entry and (tparent[order] == -1).all()
            assert (lay["words"] * 32 == twidth).all()
        assert nl <= len(order) and nl % 4 != 1  # a tile alone in its group keeps a table of its own
        spans = []
        for e, t in enumerate(order[:nl]):
            g, j = divmod(e, 4)
            assert lay["group"][t] == g and lay["stride"][t] == 32
            base = lay["off"][order[4 * g]]
            assert base % 32 == 0 and lay["off"][t] == base + 8 * j  # slot j of a 128-byte line
            if j == 0:
                spans.append((int(base), int(base) + rows * 32))
        for t in range(len(lay["off"])):
            if lay["group"][t] >= 0:
                assert t in order[:nl]
                continue
            assert lay["stride"][t] == lay["words"][t] and lay["words"][t] in (1, 2, 4, 8)
            assert lay["off"][t] % 32 == 0
            spans.append((int(lay["off"][t]), int(lay["off"][t]) + rows * int(lay["words"][t])))
        spans.sort()
        assert spans[0][0] == 0 and spans[-1][1] == lay["table_words"]
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0  # contiguous, no overlap
        # tiles that share lines can serve as pure filters: whenever the plan gives them a pre-test it is all they do
        for t in order[:nl]:
            assert lay["pre"][t] == 0 or lay["filter_only"][t] == 1


def test_layout_cases_had_line_groups():
    """The invariants above are only worth something if line groups (also several, and incomplete last ones) occurred."""
    with_lines = [c for c in LINE_CASES if c[5] >= 2]
    assert len(with_lines) >= 30, LINE_CASES
    assert any(c[5] > 4 for c in with_lines) and any(c[5] % 4 for c in with_lines) and any(c[4] > c[5] for c in with_lines)
