"""Accuracy bookkeeping of the reference's benchmarking harness, for runs of the `phage_filter` host driver.

Mirrors benchmarking/bench/utils.py (truth maps :193-212, metric counts and metrics :229-337) and the output
parser of the PhageFilter adapter (benchmarking/bench/tools/phage_filter.py:27-66) so that a table produced with
this driver can be laid beside the reference's `res_*.csv` files.  Host-side arithmetic on dictionaries only.
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Tuple


def genome_of_read(read_id: str) -> str:
    """Simulated reads are named `<genome id>_<n>` (benchmarking/bench/simulate_reads.py); the truth is the
    prefix before the last underscore (utils.py:209)."""
    return "_".join(read_id.split("_")[:-1])


def get_true_maps(reads_path: str) -> Dict[str, int]:
    """Genome -> number of simulated reads (utils.py:193-212).  The reference counts every line that starts with
    '@'; a quality line may start with '@' too, so records are walked four lines at a time here."""
    from .file_parser import read_records
    out: Counter = Counter()
    for rec in read_records(reads_path):
        out[genome_of_read(rec.id)] += 1
    return dict(out)


def parse_classification(csv_path: str, cutoff: float = 0.005) -> Dict[str, int]:
    """CLASSIFICATION.csv -> genome -> reads, keeping genomes with more than `cutoff` of all classified reads
    (phage_filter.py:52-66)."""
    counts: Dict[str, int] = {}
    with open(csv_path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            name, c = line.rsplit(",", 1)
            counts[name] = int(c)
    total = sum(counts.values())
    return {k: v for k, v in counts.items() if v > cutoff * total}


def parse_pos_filtering(path: str) -> Dict[str, int]:
    """POS_FILTERING.{fa,fq} -> genome (from the read id) -> reads kept (phage_filter.py:40-50)."""
    from .file_parser import read_records
    out: Counter = Counter()
    for rec in read_records(path):
        out[genome_of_read(rec.id)] += 1
    return dict(out)


def compute_metrics(tp: int, fp: int, fn: int) -> Dict[str, float]:
    """utils.py:229-243"""
    assert tp >= 0 and fp >= 0 and fn >= 0, "counts cannot be negative"
    return {"recall": tp / (tp + fn) if tp + fn else 0, "precision": tp / (tp + fp) if tp + fp else 0}


def get_filter_metric_counts(true_map: Dict[str, int], out_map: Dict[str, int]) -> Dict[str, int]:
    """Read-level counts: hits beyond a genome's true count are false positives (utils.py:245-273)."""
    tp = sum(min(out_map.get(g, 0), n) for g, n in true_map.items())
    fp = sum(max(0, n - true_map.get(g, 0)) for g, n in out_map.items())
    fn = sum(max(0, n - out_map.get(g, 0)) for g, n in true_map.items())
    return {"TP": tp, "FP": fp, "FN": fn}


def get_filter_metrics(true_map: Dict[str, int], out_map: Dict[str, int]) -> Tuple[float, float]:
    c = get_filter_metric_counts(true_map, out_map)
    m = compute_metrics(c["TP"], c["FP"], c["FN"])
    return m["recall"], m["precision"]


def get_classification_metric_counts(true_map: Dict[str, int], out_map: Dict[str, int]) -> Dict[str, int]:
    """Genome-level counts: a genome is classified when at least one read maps to it (utils.py:285-300)."""
    t, o = set(true_map), set(out_map)
    return {"TP": len(t & o), "FP": len(o - t), "FN": len(t - o)}


def get_classification_metrics(true_map: Dict[str, int], out_map: Dict[str, int]) -> Tuple[float, float]:
    c = get_classification_metric_counts(true_map, out_map)
    m = compute_metrics(c["TP"], c["FP"], c["FN"])
    return m["recall"], m["precision"]


def get_readcount_metrics(true_map: Dict[str, int], out_map: Dict[str, int]) -> List[int]:
    """|predicted - true| reads for every predicted genome that is real (utils.py:320-337)."""
    return [abs(n - true_map[g]) for g, n in out_map.items() if g in true_map]
