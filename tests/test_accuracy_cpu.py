"""The accuracy bookkeeping (phagefilter_b200/accuracy.py) against the expectations of the reference harness's own
unit tests (benchmarking/bench/tests/test_utils.py:10-143)."""
import pytest

from phagefilter_b200 import accuracy as A


def test_compute_metrics():  # test_utils.py:10-29
    for bad in ((-1, 0, 0), (0, -1, 0), (0, 0, -1)):
        with pytest.raises(AssertionError):
            A.compute_metrics(*bad)
    assert A.compute_metrics(0, 0, 0) == {"recall": 0, "precision": 0}
    assert A.compute_metrics(0, 5, 7) == {"recall": 0, "precision": 0}
    m = A.compute_metrics(3, 2, 1)
    assert (m["recall"], m["precision"]) == (0.75, 0.60)


def test_filter_metric_counts():  # test_utils.py:31-53
    true_map = {"t1": 5729, "t2": 6233, "t3": 5720, "t4": 682}
    out_map = {"t1": 5729, "t2": 6233 - 500, "t3": 5720 + 500, "xyz": 5262}
    c = A.get_filter_metric_counts(true_map, out_map)
    assert c == {"TP": 5729 + 5733 + 5720, "FP": 500 + 5262, "FN": 500 + 682}


def test_classification_metric_counts():  # test_utils.py:55-77
    true_map = {"t1": 5000, "t2": 5000, "t3": 5000, "t4": 5000}
    out_map = {"t1": 1, "t2": 5000, "t3": 25000, "xyz": 5000, "abc": 537}
    assert A.get_classification_metric_counts(true_map, out_map) == {"TP": 3, "FP": 2, "FN": 1}
    assert A.get_classification_metrics(true_map, out_map) == (3 / 4, 3 / 5)  # :101-120


def test_filter_metrics():  # test_utils.py:79-99
    true_map = {"t1": 100, "t2": 90, "t3": 80, "t4": 250}
    out_map = {"t1": 100, "t2": 80, "t3": 100, "xyz": 760}
    assert A.get_filter_metrics(true_map, out_map) == (0.5, 0.25)


def test_readcount_metrics():  # test_utils.py:122-143
    true_map = {"t1": 5729, "t2": 6233, "t3": 5720, "t4": 682}
    out_map = {"t1": 5729, "t2": 5733, "t3": 6220, "xyz": 5262}
    assert A.get_readcount_metrics(true_map, out_map) == [0, 500, 500]


def test_parsers(tmp_path):
    fq = tmp_path / "r.fq"
    fq.write_bytes(b"@NC_1.1_0\nACGT\n+\n@III\n@NC_1.1_1\nAC\n+\nII\n@NC_22.3_0 d\nA\n+\nI\n")
    assert A.get_true_maps(str(fq)) == {"NC_1.1": 2, "NC_22.3": 1}  # the '@' quality line is not a read
    csv = tmp_path / "CLASSIFICATION.csv"
    csv.write_text("NC_1.1,995\nNC_22.3,5\nNC_9,6\n")
    assert A.parse_classification(str(csv)) == {"NC_1.1": 995, "NC_9": 6}  # 5 is not > 0.5 % of 1006
    pos = tmp_path / "POS_FILTERING.fq"
    pos.write_bytes(b"@NC_1.1_0 |NC_1.1,NC_9\nACGT\n+\nIIII\n@NC_22.3_0 |NC_9\nA\n+\nI\n")
    assert A.parse_pos_filtering(str(pos)) == {"NC_1.1": 1, "NC_22.3": 1}
