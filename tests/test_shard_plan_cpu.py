"""pf_shard_plan (host only): the partition of a gSBT into a replicated top and per-rank subtrees."""
import numpy as np
import pytest

from tests.util import oracle_build_db, random_genomes


@pytest.fixture(scope="module")
def tree_dir(oracle, tmp_path_factory):
    rng = np.random.default_rng(9)
    genomes = random_genomes(rng, 40, 600, 900)
    d = str(tmp_path_factory.mktemp("plan") / "db")
    ot = oracle_build_db(oracle, genomes, 15, d, largest=1000)
    return d, ot


def _levels(ot):
    """level-order ids as libpfgpu numbers them: BFS, left before right."""
    pre = ot.preorder()  # (name, is_leaf, depth) in pre-order
    return pre


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
@pytest.mark.parametrize("cut", [None, 1, 3, 50])
def test_partition_properties(tree_dir, nranks, cut):
    from phagefilter_b200.shard import shard_plan
    d, ot = tree_dir
    pre = ot.preorder()
    n_nodes, depth = len(pre), max(dd for _, _, dd in pre)
    cut_level, owner = shard_plan(d, nranks, cut_level=cut)
    assert len(owner) == n_nodes
    assert 1 <= cut_level <= depth + 1
    if cut is not None:
        assert cut_level == min(cut, depth + 1)
    # nodes per level (BFS numbering = by depth, stable)
    per_level = np.bincount([dd for _, _, dd in pre])
    start = np.concatenate([[0], np.cumsum(per_level)])
    top = int(start[min(cut_level, depth + 1)])
    assert (owner[:top] == -1).all() and (owner[top:] >= 0).all() and (owner[top:] < nranks).all()
    if cut_level <= depth:
        lvl = owner[start[cut_level]:start[cut_level + 1]]
        assert (np.diff(lvl) >= 0).all()  # owners take contiguous ranges of the cut level
        if len(lvl) >= nranks:
            assert len(set(lvl.tolist())) == nranks  # nobody idle when there are enough subtrees
    if cut is None and nranks > 1 and cut_level <= depth:
        loads = np.bincount(owner[top:], minlength=nranks)
        assert top + loads.max() < n_nodes  # sharding really lowers the per-rank resident filters


def test_plan_errors(tmp_path):
    from phagefilter_b200 import _lib
    from phagefilter_b200.shard import shard_plan
    with pytest.raises(_lib.PfError) as e:
        shard_plan(str(tmp_path), 2)
    assert e.value.status == 2  # PF_ERR_IO: no tree.bin
