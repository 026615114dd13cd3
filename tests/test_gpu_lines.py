"""Entry tiles that share 128-byte lines (sliced_entry_quad_kernel): small trees are cut into narrow tiles
(pf_db_set_tile_cols) so that the cut spans several entry tiles -- one group, several groups, a last group that is not
full, a lone tile left over -- and every case must give the oracle's hit sets and counters, with the line kernel
actually used (line_loads > 0)."""
import os

import numpy as np
import pytest

from tests.util import gpu_query, oracle_build_db, random_genomes, sample_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def db(oracle, tmp_path_factory):
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.synth import make_genomes
    genomes = make_genomes(40, 10, seed=4242, len_lo=2500, len_hi=3500)  # 400 leaves, 799 nodes
    d = str(tmp_path_factory.mktemp("lines") / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=5000)
    gt = BloomTree.load(d)
    yield genomes, ot, gt
    gt.close()


def _reads(genomes, n, length, seed, background=0.7):
    from phagefilter_b200.synth import simulate_reads
    reads, _ = simulate_reads(genomes, n, length, seed=seed, error_rates=(0.0, 0.01, 0.05), background_frac=background)
    return [r.tobytes() for r in reads]


USED = {}  # case -> line loads


def _check(ot, gt, reads, theta, expect_lines=True, case=None):
    from phagefilter_b200.query import get_leaf_counts
    ot.reset_counts()
    want = ot.query_batch(reads, theta)
    ws = want.hit_sets(len(reads))
    for handover in (0, 1):
        gt.set_mode(2)
        gt.set_handover(handover)
        gt.reset_counts()
        gt.reset_stats()
        got = gpu_query(gt, reads, theta)
        assert got == ws, (theta, handover)
        assert get_leaf_counts(gt) == ot.leaf_counts(), (theta, handover)
        st = gt.stats()
        assert st.sliced_blocks == 1
        if case is not None:
            USED[f"{case} theta={theta} handover={handover}"] = int(st.line_loads)
    gt.set_handover(-1)
    return want


@pytest.fixture(autouse=True)
def _knobs():
    yield
    for k in ("PF_SLICED_FORCE_PRE", "PF_SLICED_FORCE_G"):
        os.environ.pop(k, None)


@pytest.mark.parametrize("cols,G", [(32, 4), (32, 16), (64, 2), (128, 8), (32, None)])
@pytest.mark.parametrize("theta", [0.8, 1.0, 0.5])
def test_line_groups_match_oracle(db, cols, G, theta):
    """G forces the cut ("skip every verified node with more than G leaves"): many narrow entry tiles, several groups."""
    genomes, ot, gt = db
    gt.set_tile_cols(cols)
    if G is not None:
        os.environ["PF_SLICED_FORCE_G"] = str(G)
        os.environ["PF_SLICED_FORCE_PRE"] = "1"
    reads = _reads(genomes, 6000, 150, seed=cols) + [b"", b"ACGT", b"ACGTACGTACGTACGTACGT", genomes[7][1][:20],
                                                   genomes[3][1][100:121], b"ACGTNACGTACGTACGTACGTACGTAAC" * 3,
                                                   genomes[5][1][:300].lower()]
    want = _check(ot, gt, reads, theta, case=f"cols={cols} G={G}")
    assert len(want.hits) > 500
    gt.set_tile_cols(256)
    gt.set_mode(0)


def test_line_groups_two_step_pretest(db):
    """A second probe step for the whole group as soon as one tile's plan asks for it."""
    genomes, ot, gt = db
    gt.set_tile_cols(32)
    os.environ["PF_SLICED_FORCE_PRE"] = "2"
    os.environ["PF_SLICED_FORCE_G"] = "8"
    _check(ot, gt, _reads(genomes, 4000, 150, seed=9), 0.75, case="two-step")
    _check(ot, gt, _reads(genomes, 500, 700, seed=10), 0.85, case="two-step long")
    gt.set_tile_cols(256)
    gt.set_mode(0)


def test_line_groups_long_and_mixed_reads(db):
    """More than 255 k-mers per read (16 and 32 count planes), mixed with short reads in one block."""
    genomes, ot, gt = db
    gt.set_tile_cols(32)
    os.environ["PF_SLICED_FORCE_G"] = "6"
    os.environ["PF_SLICED_FORCE_PRE"] = "1"
    g = genomes[11][1]
    reads = _reads(genomes, 300, 1000, seed=3) + _reads(genomes, 300, 150, seed=4) + [g[:2900], g[50:2000], b"ACGT" * 2000]
    _check(ot, gt, reads, 0.9, case="mixed lengths")
    _check(ot, gt, reads, 0.3, case="mixed lengths")
    long_read = (g * 30)[:70_000]  # 69,981 k-mers: 32 planes
    rng = np.random.default_rng(1)
    noise = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=70_000)].tobytes()
    _check(ot, gt, [long_read, noise, g[:500]] + _reads(genomes, 50, 150, seed=5), 0.6, case="70 kb read")
    gt.set_tile_cols(256)
    gt.set_mode(0)


def test_line_groups_chunked_batch(db):
    """The batch in several hash-cache chunks, and chunks cut again because the (read, entry tile) pairs outgrow a tiny pair
    index: the entry kernel hashes on the fly chunk by chunk and only each chunk's survivors get cached values."""
    genomes, ot, gt = db
    gt.set_tile_cols(32)
    os.environ["PF_SLICED_FORCE_G"] = "4"
    os.environ["PF_SLICED_FORCE_PRE"] = "1"
    reads = _reads(genomes, 5000, 150, seed=21) + [b"ACGTN" * 40, genomes[9][1][:200].lower(), b"", b"ACGT"]
    gt.set_hash_cache_bytes(12 * 131 * 700)  # ~700 reads per chunk
    try:
        _check(ot, gt, reads, 0.8, case="chunked")
        gt.set_frontier_cap(3000)            # fewer pairs than a chunk's reads x entry tiles
        _check(ot, gt, reads, 0.8, case="chunked, tiny pair index")
        _check(ot, gt, reads, 1.0, case="chunked, tiny pair index")
    finally:
        gt.set_frontier_cap(0xFF000000)
        gt.set_hash_cache_bytes(16 << 30)
    gt.set_tile_cols(256)
    gt.set_mode(0)


def test_threshold_edges(db):
    """theta = 0 (every node passes), theta > 1 (nothing passes), with the line layout in place."""
    genomes, ot, gt = db
    gt.set_tile_cols(64)
    os.environ["PF_SLICED_FORCE_G"] = "4"
    os.environ["PF_SLICED_FORCE_PRE"] = "1"
    reads = _reads(genomes, 200, 150, seed=6) + [b"", b"ACG"]
    _check(ot, gt, reads, 0.0, expect_lines=False)
    _check(ot, gt, reads, 1.2, expect_lines=False)
    _check(ot, gt, reads, 0.05, expect_lines=False)
    gt.set_tile_cols(256)
    gt.set_mode(0)


@pytest.mark.parametrize("k", [17, 24, 25, 32, 16, 33])
def test_line_groups_other_kmer_sizes(oracle, tmp_path, k):
    """The entry kernel hashes on the fly for 17 <= k <= 32 (every byte offset of the hash's suffix words) and reads
    cached values otherwise."""
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.synth import make_genomes
    genomes = make_genomes(40, 10, seed=4242, len_lo=2000, len_hi=3000)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, k, d, largest=4000)
    gt = BloomTree.load(d)
    gt.set_tile_cols(32)
    os.environ["PF_SLICED_FORCE_G"] = "4"
    os.environ["PF_SLICED_FORCE_PRE"] = "1"
    reads = _reads(genomes, 3000, 120, seed=k) + [genomes[0][1][:k], genomes[1][1][5:5 + k + 1], b"ACGTN" * 30,
                                                 genomes[2][1][:33], genomes[2][1][:63], genomes[2][1][:64], genomes[2][1][:65]]
    for theta in (0.9, 0.6):
        _check(ot, gt, reads, theta, case=f"k={k}")
    gt.close()


def test_line_kernel_was_exercised():
    """The cases above are only worth something if the line kernel took part in most of them."""
    import json
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/lines_cases.json", "w") as f:
        json.dump(USED, f, indent=1)
    used = [k for k, v in USED.items() if v > 0]
    assert len(used) >= 30, USED
    assert sum(1 for k in used if k.startswith("k=")) >= 12, USED
