"""ResultMap -- read id -> set of genome ids (reference: src/result_map.rs:9-46)."""
from __future__ import annotations

from typing import Dict, Set


class ResultMap:
    def __init__(self) -> None:
        self.read_map: Dict[str, Set[str]] = {}

    def add_read_map(self, read_id: str, genome_id: str) -> None:
        """result_map.rs:20-22"""
        self.read_map.setdefault(read_id, set()).add(genome_id)

    def get_ext_id(self, read_id: str) -> str:
        """result_map.rs:24-37: "{id} |{g1,g2,...}" (genome order unspecified in the reference;
        sorted here so the output is deterministic)."""
        genomes = ",".join(sorted(self.read_map.get(read_id, ())))
        return f"{read_id} |{genomes}"

    def read_mapped(self, read_id: str) -> bool:
        """result_map.rs:39-41"""
        return read_id in self.read_map

    def empty_read_map(self) -> None:
        """result_map.rs:43-45"""
        self.read_map.clear()
