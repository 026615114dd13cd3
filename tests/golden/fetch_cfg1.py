"""Copies BASELINE.json config #1 inputs (the reference's examples/: 107 viral genomes, 9 x 10,000 x 100 bp
simulated reads) from /root/reference into tests/_cfg1_data/ (git-ignored: reference data is not committed,
but the directory travels with the repo snapshot to the GPU box, where /root/reference does not exist)."""
import os
import shutil
import sys

SRC = "/root/reference/examples"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DST = os.path.join(ROOT, "tests", "_cfg1_data")


def main():
    if not os.path.isdir(SRC):
        sys.exit(f"{SRC} not present")
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(os.path.join(SRC, "genomes", "viral_genome_dir"), os.path.join(DST, "viral_genome_dir"))
    shutil.copytree(os.path.join(SRC, "test_reads"), os.path.join(DST, "test_reads"))
    print("copied to", DST)


if __name__ == "__main__":
    main()
