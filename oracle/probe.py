"""oracle/probe.py -- TEST INFRASTRUCTURE.  Runs oracle/_ref/ref_probe (the unmodified reference
translation units + our dumping driver) and parses its hex-float text dump."""
from __future__ import annotations

import os
import subprocess
from typing import Dict, List

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(HERE, "_ref", "ref_probe")
REF_BIN = os.path.join(HERE, "_ref", "PhyloAcc-ST-ref")


def hexf(tokens):
    return np.array([float.fromhex(t) for t in tokens], dtype=np.float64)


def parse_probe(text: str) -> Dict:
    out: Dict = {"elements": []}
    cur = None
    for line in text.split("\n"):
        t = line.split()
        if not t:
            continue
        k = t[0]
        if k == "tree":
            out.update(S=int(t[2]), N=int(t[4]), C=int(t[6]), G=int(t[8]))
        elif k == "names":
            out["names"] = t[1:]
        elif k == "parent":
            out["parent"] = np.array(t[1:], dtype=np.int32)
        elif k == "children":
            out["children"] = np.array(t[1:], dtype=np.int32).reshape(-1, 2)
        elif k == "distances":
            out["distances"] = hexf(t[1:])
        elif k == "pi":
            out["pi"] = hexf(t[1:])
        elif k == "Q":
            out["Q"] = hexf(t[1:]).reshape(4, 4)
        elif k == "upper":
            out["upper"] = set(map(int, t[1:]))
        elif k == "upper_conserve":
            out["upper_conserve"] = set(map(int, t[1:]))
        elif k == "element_size":
            out["element_size"] = np.array(t[1:], dtype=np.int64)
        elif k == "logTMnull":
            out.setdefault("logTMnull", {})[int(t[1])] = hexf(t[2:]).reshape(4, 4)
        elif k == "element":
            cur = dict(c=int(t[1]), filter=bool(int(t[3])), GG=int(t[5]), masks=[], tg=[], states=[], histories=[], eval2=[])
            out["elements"].append(cur)
        elif k == "root":
            cur["root"] = int(t[1])
        elif k == "missing":
            cur["missing"] = np.array(t[1:], dtype=np.int8).astype(bool)
        elif k == "nodes":
            cur["nodes"] = list(map(int, t[1:]))
        elif k == "children2":
            cur["children2"] = np.array(t[1:], dtype=np.int32).reshape(-1, 2)
        elif k == "parent2":
            cur["parent2"] = np.array(t[1:], dtype=np.int32)
        elif k == "distances2":
            cur["distances2"] = hexf(t[1:])
        elif k == "upper_c":
            cur["upper_c"] = list(map(int, t[1:]))
        elif k == "upper_conserve_c":
            cur["upper_conserve_c"] = list(map(int, t[1:]))
        elif k == "col":
            i = t.index("masks")
            j = t.index("tg")
            cur["masks"].append(list(map(int, t[i + 1:j])))
            cur["tg"].append(list(map(int, t[j + 1:])))
        elif k == "loglik":
            cur["states"].append(dict(resZ=int(t[1]), loglik=float.fromhex(t[2])))
        elif k == "persite":
            cur["states"][-1]["persite"] = hexf(t[1:])
        elif k == "fixZ":
            cur["states"][-1]["fixZ"] = np.array(t[1:], dtype=np.int8)
        elif k == "emission":
            if int(t[1]) == 0:
                cur["histories"].append(dict(emission={}, back={}))
            cur["histories"][-1]["emission"][int(t[1])] = hexf(t[2:])
        elif k == "back":
            cur["histories"][-1]["back"][int(t[1])] = hexf(t[2:])
        elif k == "eval2":
            cur["eval2"].append(dict(resZ=int(t[1]), logml=float.fromhex(t[2]), maxloglik=float.fromhex(t[4]), failure=int(t[6])))
        elif k == "maxZ":
            cur["eval2"][-1]["maxZ"] = np.array(t[1:], dtype=np.int8)
    for e in out["elements"]:
        if not e["filter"]:
            e["masks"] = np.array(e["masks"], dtype=np.uint8)
            e["tg"] = np.array(e["tg"], dtype=np.int8)
    return out


def run_probe(cfg: str, script_text: str, workdir: str) -> Dict:
    sp = os.path.join(workdir, "probe_script.txt")
    with open(sp, "w") as f:
        f.write(script_text)
    r = subprocess.run([PROBE, cfg, sp], capture_output=True, text=True, check=True)
    return parse_probe(r.stdout)
