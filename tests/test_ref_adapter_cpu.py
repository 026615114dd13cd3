"""The reference's own benchmarking adapter against this driver.  tests/golden/ref_adapter_vectors.json holds
what the reference's Python code produces when imported (tests/golden/make_ref_adapter_vectors.py): the command lines of
benchmarking/bench/tools/phage_filter.py:68-118 and the results of its output parser and of bench/utils.py's metric
helpers on fixed inputs.  Here: the fixture is current, `phage_filter` accepts every one of those command lines (query
runs end to end on the CPU with the preloaded stub of the GPU entry points), and phagefilter_b200/accuracy.py reproduces
the reference's parser and metric results."""
import json
import os
import subprocess
import sys

import pytest

from phagefilter_b200 import accuracy as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
VEC = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_adapter_vectors.json")))


@pytest.mark.skipif(not os.path.isdir("/root/reference/benchmarking"), reason="reference not present on this machine")
def test_fixture_is_what_the_reference_produces(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_ref_adapter_vectors.py")],
                         capture_output=True, text=True, cwd=str(tmp_path), env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0, out.stderr
    # the generator rewrites the committed file in place: it must not have changed
    assert json.load(open(os.path.join(ROOT, "tests", "golden", "ref_adapter_vectors.json"))) == VEC


def test_accuracy_module_equals_reference_parsers_and_metrics(tmp_path):
    for i, c in enumerate(VEC["parse"]):
        d = tmp_path / str(i)
        d.mkdir()
        (d / "CLASSIFICATION.csv").write_text(c["classification_csv"])
        (d / "POS_FILTERING.fa").write_text(c["pos_filtering_fa"])
        assert A.parse_classification(str(d / "CLASSIFICATION.csv")) == c["classification"]
        assert A.parse_pos_filtering(str(d / "POS_FILTERING.fa")) == c["filter"]
    for c in VEC["metrics"]:
        assert list(A.get_classification_metrics(c["true"], c["out"])) == pytest.approx(c["classification"])
        assert list(A.get_filter_metrics(c["true"], c["out"])) == pytest.approx(c["filter"])
        assert list(A.get_readcount_metrics(c["true"], c["out"])) == c["readcount"]


def test_driver_accepts_the_adapters_command_lines(tmp_path):
    from tests.test_cli_driver_stub_cpu import STUB_SRC
    if not os.path.exists(BIN):
        pytest.skip("phage_filter binary not built")
    stub = str(tmp_path / "stub_pfgpu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", stub, STUB_SRC], check=True)
    reads = tmp_path / "reads.fa"
    reads.write_text("".join(f">NC_{i % 5}.1_{i}\n{'ACGT' * 12}\n" for i in range(300)))
    genomes = tmp_path / "genomes"
    genomes.mkdir()
    (genomes / "g.fa").write_text(">NC_1.1\n" + "ACGTTGCA" * 40 + "\n")
    for n, c in enumerate(VEC["commands"]):
        sub = {"{DB}": str(tmp_path / "db"), "{GENOMES}": str(genomes), "{READS}": str(reads), "{OUT}": str(tmp_path / f"out{n}")}
        for cmd in c["run"]:
            argv = [BIN if a == "./target/release/phage_filter" else sub.get(a, a) for a in cmd]
            p = subprocess.run(argv, capture_output=True, env=dict(os.environ, LD_PRELOAD=stub), timeout=120)
            assert p.returncode == 0, (argv, p.stderr.decode())
            out = sub["{OUT}"]
            assert A.parse_classification(os.path.join(out, "CLASSIFICATION.csv"))
            assert os.path.exists(os.path.join(out, "POS_FILTERING.fa")) == c["filter_reads"]
        for cmd in c["build"]:
            # build needs the GPU builder: without a device it must fail because of that, never because of an argument
            argv = [BIN if a == "./target/release/phage_filter" else sub.get(a, a) for a in cmd]
            p = subprocess.run(argv, capture_output=True, timeout=120)
            err = p.stderr.decode().lower()
            assert p.returncode == 0 or ("cuda" in err or "device" in err), (argv, err)
            assert "unrecognized" not in err and "unexpected" not in err and "invalid value" not in err, err
