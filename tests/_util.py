"""Shared helpers for the tests: golden-case loading through the ORACLE's parsers (test infrastructure)."""
from __future__ import annotations

import json
import os
import shutil
import tempfile

import numpy as np

from oracle import coracle, frontend as fe, probe

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ratite43", "tiny8"]


class Case:
    def __init__(self, name: str, extra_params: dict | None = None):
        self.name = name
        src = os.path.join(GOLD, name)
        self.dir = tempfile.mkdtemp(prefix=f"pa_case_{name}_")
        for f in ("sim.mod", "sim.fa", "sim.bed", "sim.id"):
            shutil.copy(os.path.join(src, f), os.path.join(self.dir, f))
        os.makedirs(os.path.join(self.dir, "out"), exist_ok=True)
        with open(os.path.join(src, "params.cfg")) as f:
            body = f.read()
        if extra_params:
            keep = [l for l in body.split("\n") if l and l.split()[0] not in extra_params]
            body = "\n".join(keep + [f"{k} {v}" for k, v in extra_params.items()]) + "\n"
        self.cfg = os.path.join(self.dir, "sim.cfg")
        with open(self.cfg, "w") as f:
            f.write(f"PHYTREE_FILE {self.dir}/sim.mod\nSEG_FILE {self.dir}/sim.bed\nALIGN_FILE {self.dir}/sim.fa\n")
            f.write(f"ID_FILE {self.dir}/sim.id\nRESULT_FOLDER {self.dir}/out\n")
            f.write(body)
        self.P = fe.load_params(self.cfg)
        self.tree = fe.parse_mod(self.P.PHYTREE_FILE)
        names, X = fe.load_fasta(self.P.ALIGN_FILE)
        self.elem_names, pos = fe.load_bed(self.P.SEG_FILE)
        self.starts, self.lens = fe.element_layout(pos)
        self.model = fe.build_model(self.tree, self.P)
        self.X = fe.reorder_alignment(self.tree, names, X)
        self.C = len(self.elem_names)
        self._els = {}

    def element(self, c: int) -> fe.Element:
        if c not in self._els:
            self._els[c] = fe.element_frontend(self.model, self.X, int(self.starts[c]), int(self.lens[c]))
        return self._els[c]

    def handle(self, c: int, canonical_sites: bool = False) -> coracle.ElementHandle:
        el = self.element(c)
        order = fe.compress_patterns(el.masks).site_order if canonical_sites else None
        return coracle.ElementHandle(self.model, el, order)

    def golden_probe(self):
        with open(os.path.join(GOLD, self.name, "probe.txt")) as f:
            return probe.parse_probe(f.read())

    def golden_requests(self):
        with open(os.path.join(GOLD, self.name, "requests.json")) as f:
            return json.load(f)

    def golden_refruns(self):
        out = []
        src = os.path.join(GOLD, self.name)
        for fn in sorted(os.listdir(src)):
            if fn.startswith("refrun_seed") and fn.endswith(".json.gz"):
                import gzip

                with gzip.open(os.path.join(src, fn), "rt") as f:
                    out.append(json.load(f))
            elif fn.startswith("refrun_seed"):
                with open(os.path.join(src, fn)) as f:
                    out.append(json.load(f))
        return out

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def parse_table(text: str):
    lines = [l for l in text.split("\n") if l]
    header = lines[0].split("\t")
    rows = [l.split("\t") for l in lines[1:]]
    return header, rows


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
