"""oracle/frontend.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-Python/numpy restatement of the reference's *integer / parsing* stages for the
PhyloAcc-ST hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this package; the product path (phyloacc_b200/) never does.

Each function cites the reference file:line it follows (paths relative to
/root/reference/src/PhyloAcc-ST/).  Parity status: PINNED -- tests/test_oracle_pinning.py
checks every function here against dumps of the unmodified reference code
(oracle/_ref/ref_probe, fixtures under tests/golden/).

The one stage with no reference counterpart is site-pattern compression (the reference
loops over raw columns); it is *defined* here (``compress_patterns``) and
validated by the property that expanding patterns by their weights reproduces the
reference's per-column tip encoding.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Set

import numpy as np

NB = 4  # num_base (the 5-state indel model is dead code: INDEL defaults to 0, bpp.hpp:247-253)


# --------------------------------------------------------------------------- parameters
@dataclasses.dataclass
class Params:
    """Engine parameters with the binary's defaults (main.cpp:49-90, LoadParams main.cpp:100-287)."""

    PHYTREE_FILE: str = ""
    ALIGN_FILE: str = ""
    SEG_FILE: str = ""
    ID_FILE: str = ""
    BATCH: int = -1
    RESULT_FOLDER: str = ""
    PREFIX: str = "test"
    SEED: int = 5
    INIT_CONSERVE_RATE: float = 0.5
    INIT_ACCE_RATE: float = 1.0
    CONSERVE_PRIOR_A: float = 5.0
    CONSERVE_PRIOR_B: float = 0.04
    ACCE_PRIOR_A: float = 10.0
    ACCE_PRIOR_B: float = 0.2
    ROPT: int = 1
    CUB: float = 1.0
    NLB: float = 0.6
    BURNIN: int = 200
    MCMC: int = 800
    ADAPT_FREQ: int = 500
    INIT_LRATE: float = 0.3
    INIT_LRATE2: float = 0.1
    INIT_GRATE: float = 0.5
    HYPER_LRATE_A: float = 1.0
    HYPER_LRATE_B: float = 9.0
    HYPER_GRATE_A: float = 3.0
    HYPER_GRATE_B: float = 1.0
    HYPER_LRATE2_A: float = 1.0
    HYPER_LRATE2_B: float = 9.0
    CHAIN: int = 0
    OUTGROUP: str = ""
    TARGETSPECIES: str = ""
    CONSERVE: str = ""
    CONSERVE_PROP: float = 0.8
    GAP_PROP: float = 0.8
    CONSTOMIS: float = 0.01
    GAPCHAR: str = "-"
    PRUNE_TREE: bool = False
    TRIM_GAP_PERCENT: float = 1.0
    MIN_LEN: int = 50
    SAMPLE_HYPER: bool = False
    VERBOSE: bool = False
    NUM_THREAD: int = 1


_INT_KEYS = {"BATCH", "SEED", "ROPT", "BURNIN", "MCMC", "ADAPT_FREQ", "CHAIN", "MIN_LEN", "NUM_THREAD"}
_BOOL_KEYS = {"PRUNE_TREE", "SAMPLE_HYPER", "VERBOSE"}
_STR_KEYS = {"PHYTREE_FILE", "ALIGN_FILE", "SEG_FILE", "ID_FILE", "RESULT_FOLDER", "PREFIX", "OUTGROUP",
             "TARGETSPECIES", "CONSERVE"}


def load_params(path: str) -> Params:
    """main.cpp:134-287: whitespace-split words, '#' comments and blank lines skipped, exactly two words
    otherwise error, unknown keys skipped; HYPER_LRATE2_A == 0 forces INIT_LRATE2 = 0 (main.cpp:286)."""
    p = Params()
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            words = line.split()
            if not words or words[0][0] == "#":
                continue
            if len(words) != 2:
                raise ValueError(f"Line {ln} in the parameter file is not formatted correctly: {line}")
            k, v = words
            if k == "GAPCHAR":
                p.GAPCHAR = v[0]
            elif k in _INT_KEYS:
                setattr(p, k, int(v))
            elif k in _BOOL_KEYS:
                setattr(p, k, v in ("true", "1"))
            elif k in _STR_KEYS:
                setattr(p, k, v.strip(' "\t\n') if k.endswith("_FILE") or k == "RESULT_FOLDER" else v)
            elif hasattr(p, k):
                setattr(p, k, float(v))
            # INDEL / INDEL2 / SAMPLE_INDEL / unknown keys: ignored
    if p.HYPER_LRATE2_A == 0:
        p.INIT_LRATE2 = 0.0
    return p


# --------------------------------------------------------------------------- .mod / Newick
@dataclasses.dataclass
class Tree:
    S: int
    N: int
    parent: np.ndarray  # int32 [N]; -1 for root
    children: np.ndarray  # int32 [N,2]; -1 for leaves; ordered by node id (bpp.cpp:105-117)
    dist: np.ndarray  # float64 [N], values passed through float32 (newick.cpp:185,300,325)
    names: List[str]
    pi: np.ndarray
    Q: np.ndarray  # [4,4] rows i->j


def _parse_newick(s: str):
    """Recursive-descent equivalent of parseTree (newick.cpp:145-331): returns nested
    (children, label, dist) with dist = float32(atof(text)) or 0 when absent."""
    pos = 0

    def atof32(txt: str) -> float:
        # atof semantics: longest valid numeric prefix, 0.0 if none
        import re
        m = re.match(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", txt)
        v = float(m.group(0)) if m else 0.0
        return float(np.float32(v))

    def node():
        nonlocal pos
        kids = []
        if s[pos] == "(":
            pos += 1
            while True:
                kids.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("bad newick at %d" % pos)
        start = pos
        while pos < len(s) and s[pos] not in ",):;":
            pos += 1
        label = s[start:pos]
        dist = 0.0
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",);":
                pos += 1
            dist = atof32(s[start:pos])
        return (kids, label, dist)

    root = node()
    return root


def parse_mod(path: str) -> Tree:
    """LoadPhyloTree (newick.cpp:431-517) + InitPhyloTree (bpp.cpp:84-117)."""
    pi = np.zeros(4)
    Q = np.zeros((4, 4))
    newick = None
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        line = lines[i]
        i += 1
        if line == "":
            break  # newick.cpp:457: parsing stops at the first empty line
        line = line.strip(" \t\n")
        sp = line.split(" ")
        if sp[0] == "BACKGROUND:":
            vals = [x for x in sp[1:] if x != ""]
            for k, v in enumerate(vals[:4]):
                pi[k] = float(v)
        elif sp[0] == "RATE_MAT:":
            for r in range(4):
                vals = [x for x in lines[i].strip(" \t\n").split(" ") if x != ""]
                i += 1
                for k, v in enumerate(vals[:4]):
                    Q[r, k] = float(v)
        elif sp[0] == "TREE:":
            newick = sp[1].strip(" \t\n")
    if newick is None:
        raise ValueError("no TREE: line")
    root = _parse_newick(newick)

    # TravelTree1/2/3 (newick.cpp:364-428): leaves numbered in DFS order, internals in post-order
    S = 0

    def count(n):
        nonlocal S
        if not n[0]:
            S += 1
        for k in n[0]:
            count(k)

    count(root)
    N = 2 * S - 1
    leaf = [-1]
    inner = [-1]
    ids: Dict[int, int] = {}

    def number(n):
        for k in n[0]:
            number(k)
        if not n[0]:
            leaf[0] += 1
            ids[id(n)] = leaf[0]
        else:
            inner[0] += 1
            ids[id(n)] = inner[0] + S

    number(root)
    names = [""] * N
    dist = np.zeros(N)
    dag = np.zeros((N, N), dtype=bool)

    def fill(n):
        me = ids[id(n)]
        for k in n[0]:
            dag[me, ids[id(k)]] = True
            fill(k)
        dist[me] = n[2]
        if n[1] == "":
            raise ValueError("Some nodes in the tree aren't labeled")  # newick.cpp:416-419
        names[me] = n[1]

    fill(root)
    children = -np.ones((N, 2), dtype=np.int32)
    parent = -np.ones(N, dtype=np.int32)
    for a in range(N):
        p = -1
        for b in range(N):
            if dag[a, b]:
                p += 1
                children[a, p] = b  # a third child would overflow in the reference; binary trees only
                parent[b] = a
    return Tree(S=S, N=N, parent=parent, children=children, dist=dist, names=names, pi=pi, Q=Q)


# --------------------------------------------------------------------------- FASTA / BED / ids
def load_fasta(path: str):
    """LoadPhyloProfiles (profile.cpp:36-65): names after '>', sequences concatenated and lower-cased."""
    names: List[str] = []
    seqs: List[bytes] = []
    cur: List[bytes] = []
    with open(path, "rb") as f:
        for raw in f.read().split(b"\n"):
            line = raw.strip(b" \t\n")
            if line.startswith(b">"):
                if names:
                    seqs.append(b"".join(cur).lower())
                names.append(line[1:].strip(b" \t\n").decode())
                cur = []
            else:
                cur.append(line)
    seqs.append(b"".join(cur).lower())
    G = len(seqs[0])
    assert all(len(s) == G for s in seqs)
    X = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), G).copy()
    return names, X


def load_bed(path: str):
    """profile.cpp:79-92: tab split, stop at the first line with < 3 fields, col0 = name, atoi(col1), atoi(col2)."""
    names, pos = [], []

    def atoi(t):
        import re
        m = re.match(r"\s*[-+]?\d+", t)
        return int(m.group(0)) if m else 0

    with open(path) as f:
        for raw in f.read().split("\n"):
            line = raw.strip(" \t\n")
            sp = line.split("\t")
            if len(sp) < 3:
                break
            names.append(sp[0])
            pos.append((atoi(sp[1]), atoi(sp[2])))
    return names, np.asarray(pos, dtype=np.int64).reshape(-1, 2)


def element_layout(pos: np.ndarray):
    """bpp.hpp:257-274: elements are assumed concatenated in file order; only end-start is used."""
    lens = (pos[:, 1] - pos[:, 0]).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return starts, lens


def load_ids(path: str, C: int, batch: int = -1) -> List[int]:
    """main.cpp:406-439."""
    if path == "":
        if batch == -1:
            return list(range(C))
        import math
        temp = int(math.ceil(C // 3))
        return [c for c in range(batch * temp, (batch + 1) * temp) if c < C]
    ids = []
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w:
                continue
            import re
            m = re.match(r"[-+]?\d+", w[0])
            ids.append(int(m.group(0)) if m else 0)
    return ids


# --------------------------------------------------------------------------- batch model (BPP)
@dataclasses.dataclass
class Model:
    tree: Tree
    params: Params
    species_names: List[str]
    outgroup: List[int]
    conservedgroup: List[int]
    upper: Set[int]
    upper_conserve: Set[int]


def _upper(tree: Tree, group: Sequence[int]) -> Set[int]:
    """getUppertree (bpp.cpp:486-504): ancestors-or-self of each species, plus the root."""
    root = tree.N - 1
    out: Set[int] = set()
    for p in group:
        while p != root:
            out.add(int(p))
            p = tree.parent[p]
    out.add(root)
    return out


def build_model(tree: Tree, params: Params) -> Model:
    """BPP::BPP (bpp.hpp:230-243,303-310).  Species names not found map to index S
    (std::find returns end()), as in the reference."""
    names = tree.names[: tree.S]

    def ids(txt: str) -> List[int]:
        txt = txt.strip(" \t\n")
        if txt == "":
            return []
        out = []
        for nm in txt.split(";"):
            out.append(names.index(nm) if nm in names else tree.S)
        return out

    outgroup = ids(params.OUTGROUP)
    cons = ids(params.CONSERVE)
    return Model(tree=tree, params=params, species_names=names, outgroup=outgroup, conservedgroup=cons,
                 upper=_upper(tree, outgroup), upper_conserve=_upper(tree, cons))


def reorder_alignment(tree: Tree, fasta_names: List[str], X: np.ndarray) -> np.ndarray:
    """MatchProfAndTree (bpp.cpp:133-184): rows reordered to tree leaf order when every tree species is found."""
    idx = []
    for nm in tree.names[: tree.S]:
        if nm not in fasta_names:
            return X  # reference falls back to file order
        idx.append(fasta_names.index(nm))
    return X[idx]


# --------------------------------------------------------------------------- per-element front-end (BPP_C ctor)
_MASK = {ord("a"): 1, ord("c"): 2, ord("g"): 4, ord("t"): 8,
         ord("r"): 1 | 4, ord("y"): 2 | 8, ord("k"): 4 | 8, ord("m"): 1 | 2, ord("s"): 2 | 4, ord("w"): 1 | 8}
_TG = {ord("a"): 0, ord("c"): 1, ord("g"): 2, ord("t"): 3}


@dataclasses.dataclass
class Element:
    filter: bool
    GG: int = 0
    masks: Optional[np.ndarray] = None  # uint8 [GG, S]: bit b set <=> base b allowed (lambda == 0)
    tg: Optional[np.ndarray] = None  # int8 [GG, N]: tips 0..3 / -1 ambiguous / 4 gap / 5 other; internal -1 or 4
    missing: Optional[np.ndarray] = None  # bool [N]
    root: int = -1
    nodes: Optional[List[int]] = None  # post-order of the pruned tree
    internal_nodes: Optional[List[int]] = None
    children2: Optional[np.ndarray] = None
    parent2: Optional[np.ndarray] = None
    distances2: Optional[np.ndarray] = None
    upper_c: Optional[List[int]] = None
    upper_conserve_c: Optional[List[int]] = None


def element_frontend(model: Model, X: np.ndarray, start: int, length: int) -> Element:
    """BPP_C::BPP_C (bpp_c.hpp:143-476), integer part."""
    tree, P = model.tree, model.params
    S, N = tree.S, tree.N
    GG = int(length)
    if GG < P.MIN_LEN:  # bpp_c.hpp:158
        return Element(filter=True)
    gap = ord(P.GAPCHAR)
    cols = X[:, start:start + GG]
    masks = np.zeros((GG, S), dtype=np.uint8)
    tg = -np.ones((GG, N), dtype=np.int8)
    for s in range(S):  # bpp_c.hpp:208-264
        for g in range(GG):
            ch = int(cols[s, g])
            if ch in _MASK:
                masks[g, s] = _MASK[ch]
                tg[g, s] = _TG.get(ch, -1)
            else:
                masks[g, s] = 15
                tg[g, s] = 4 if ch == gap else 5
    if P.TRIM_GAP_PERCENT < 1:  # bpp_c.hpp:268-301
        mis = (tg[:, :S] >= 4).sum(axis=1)
        drop = mis > S * P.TRIM_GAP_PERCENT
        if GG - int(drop.sum()) < P.MIN_LEN:
            return Element(filter=True)
        masks = masks[~drop]
        tg = tg[~drop]
        GG = masks.shape[0]
    num_missing = (tg[:, :S] == 4).sum(axis=0)  # bpp_c.hpp:304-314

    children2 = tree.children.copy()
    parent2 = tree.parent.copy()
    distances2 = tree.dist.copy()
    missing = np.zeros(N, dtype=bool)
    for s in range(S, N):  # bpp_c.hpp:333-351 (node ids ascending: children before parents)
        for chi in children2[s]:
            if chi < S and num_missing[chi] > P.GAP_PROP * GG:
                missing[chi] = True
        if missing[children2[s, 0]] and missing[children2[s, 1]]:
            missing[s] = True
    ct = sum(1 for c in model.conservedgroup if c < N and missing[c])  # bpp_c.hpp:361-373
    if ct > P.CONSERVE_PROP * len(model.conservedgroup):
        return Element(filter=True)

    state = {"root": N - 1}
    upper = set(range(N)) if P.PRUNE_TREE else model.upper

    def subtree_missing(ss: int, child: int):  # getSubtree_missing (bpp_c.cpp:64-113)
        if ss not in upper or ss < S:
            return
        s = ss
        while s in upper and s >= S:
            p0, p1 = int(children2[s, 0]), int(children2[s, 1])
            if missing[p0] and not missing[p1]:
                keep, lose = p1, p0
            elif not missing[p0] and missing[p1]:
                keep, lose = p0, p1
            else:
                subtree_missing(p0, 0)
                subtree_missing(p1, 1)
                return
            parent2[lose] = -1
            distances2[keep] += distances2[s]
            if s == state["root"]:
                state["root"] = keep
            else:
                children2[parent2[ss], child] = keep
                parent2[keep] = parent2[ss]
            s = keep

    subtree_missing(N - 1, -1)
    root = state["root"]
    parent2[root] = N  # bpp_c.hpp:390

    nodes: List[int] = []

    def post(j):  # getSubtree (bpp_c.cpp:44-60)
        for chi in children2[j]:
            if chi != -1:
                post(int(chi))
        nodes.append(j)

    post(root)
    internal = [n for n in nodes if n >= S]
    upper_c = [n for n in nodes if n in model.upper]
    upper_conserve_c = [n for n in nodes if n in model.upper_conserve]
    for n in internal:  # bpp_c.hpp:417-428
        c0, c1 = children2[n]
        both = (tg[:, c0] >= NB) & (tg[:, c1] >= NB)
        tg[both, n] = NB
    return Element(filter=False, GG=GG, masks=masks, tg=tg, missing=missing, root=root, nodes=nodes,
                   internal_nodes=internal, children2=children2, parent2=parent2, distances2=distances2,
                   upper_c=upper_c, upper_conserve_c=upper_conserve_c)


def fixz_for_model(el: Element, N: int, resZ: int) -> np.ndarray:
    """initMCMC (bpp_c.cpp:155-161,240-289): fixZ per model (resZ 0 = M0, 2 = M1, 1 = M2)."""
    fixZ = -np.ones(N, dtype=np.int8)
    fixZ[el.upper_c] = 1
    fixZ[el.root] = 1
    if resZ == 2:
        fixZ[el.upper_conserve_c] = 1
    elif resZ == 0:
        fixZ[el.nodes[:-1]] = 1
    return fixZ


# --------------------------------------------------------------------------- site-pattern compression (new stage)
@dataclasses.dataclass
class Patterns:
    P: int
    codes: np.ndarray  # uint8 [P, S] nibble masks, rows sorted lexicographically
    weights: np.ndarray  # int32 [P]
    site_pattern: np.ndarray  # int32 [GG]: pattern of each (trimmed) column, original column order
    site_order: np.ndarray  # int32 [GG]: columns sorted by (pattern, column) -- the canonical site order


def compress_patterns(masks: np.ndarray) -> Patterns:
    """Unique S-nibble columns, sorted lexicographically by (species 0, species 1, ...)."""
    codes, inv, counts = np.unique(masks, axis=0, return_inverse=True, return_counts=True)
    inv = np.asarray(inv).reshape(-1)
    order = np.lexsort((np.arange(masks.shape[0]), inv)).astype(np.int32)
    return Patterns(P=codes.shape[0], codes=codes.astype(np.uint8), weights=counts.astype(np.int32),
                    site_pattern=inv.astype(np.int32), site_order=order)
