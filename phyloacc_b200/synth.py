"""Synthetic PhyloAcc-ST inputs: a ratite-shaped labelled tree (.mod),
REV rate matrix, and alignments simulated down the tree with per-element conserved /
accelerated rates.  Used by tests, golden-fixture generation and bench.py; there is no
reference data in the container (the reference's test data lives in another repo,
README.md:40-42 of the reference).

Everything is seeded and vectorised over elements so that 10^5-element batches are
generated in seconds.  Outputs are either in-memory arrays (``SynthBatch``) or the
reference's on-disk formats (``write_files``: .mod / FASTA / BED / id / cfg,
formats per SURVEY.md App. B).
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence

import numpy as np

BASES = np.frombuffer(b"acgt", dtype=np.uint8)


@dataclasses.dataclass
class SynthTree:
    """Binary tree in the reference's numbering (newick.cpp:382-400): leaves 0..S-1 in
    left-to-right order, internal nodes S..N-1 in post-order, root = N-1."""

    S: int
    N: int
    parent: np.ndarray  # int32 [N], -1 for the root
    children: np.ndarray  # int32 [N,2], ordered by node id (bpp.cpp:105-117), -1 for leaves
    dist: np.ndarray  # float64 [N] (already rounded to 6 decimals as written to the .mod)
    names: List[str]
    newick: str
    outgroup: List[int]
    targets: List[int]
    conserve: List[int]
    pi: np.ndarray
    Q: np.ndarray  # rows i->j, rows sum to 0


def _yule_topology(rng: np.random.Generator, n_tips: int):
    """Random Yule topology as nested tuples of tip placeholders."""
    lineages = [("tip",) for _ in range(n_tips)]
    # coalesce random pairs until one lineage remains (equivalent in distribution to Yule)
    while len(lineages) > 1:
        i, j = sorted(rng.choice(len(lineages), size=2, replace=False))
        a, b = lineages[i], lineages[j]
        lineages = [x for k, x in enumerate(lineages) if k not in (i, j)]
        lineages.append((a, b))
    return lineages[0]


def make_rev_matrix(rng: np.random.Generator, pi=(0.28, 0.22, 0.22, 0.28)):
    pi = np.asarray(pi, dtype=np.float64)
    ex = rng.uniform(0.5, 2.0, size=6)
    R = np.zeros((4, 4))
    k = 0
    for i in range(4):
        for j in range(i + 1, 4):
            R[i, j] = R[j, i] = ex[k]
            k += 1
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    scale = -(pi * np.diag(Q)).sum()
    Q /= scale  # one expected substitution per site per unit time (phyloFit convention)
    # the .mod file carries 6 decimals; round here so that in-memory == on-disk
    Q = np.round(Q, 6)
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return pi, Q


def make_tree(S: int = 43, seed: int = 43, n_out: int = 7, n_target: int = 9,
              mean_bl: float = 0.05) -> SynthTree:
    """Yule tree with an ``n_out``-taxon outgroup clade sister to the rest at the root,
    ``n_target`` target tips in (up to) 3 clades, the remaining tips conserved."""
    rng = np.random.default_rng(seed)
    n_out = min(n_out, max(1, S // 4))
    ingroup = _yule_topology(rng, S - n_out)
    outgroup = _yule_topology(rng, n_out)
    top = (outgroup, ingroup)

    names: List[Optional[str]] = []
    parent: List[int] = []
    dist: List[float] = []

    # first pass: count leaves to know S, assign ids in the reference's order
    leaf_ctr = [0]
    int_ctr = [0]
    records = {}  # id -> (children ids)

    def bl():
        return float(np.round(np.clip(rng.exponential(mean_bl), 0.002, 0.3), 6))

    def walk(node, is_root=False):
        if node == ("tip",):
            nid = leaf_ctr[0]
            leaf_ctr[0] += 1
            records[nid] = ()
            return nid, f"sp{nid:03d}", bl()
        kids = [walk(ch) for ch in node]
        nid = S + int_ctr[0]
        int_ctr[0] += 1
        records[nid] = tuple(k[0] for k in kids)
        return nid, f"anc{nid:03d}", (0.0 if is_root else bl()), kids

    def render(rec):
        if len(rec) == 3:
            nid, name, d = rec
            return f"{name}:{d:.6f}"
        nid, name, d, kids = rec
        inner = ",".join(render(k) for k in kids)
        return f"({inner}){name}:{d:.6f}"

    root_rec = walk(top, is_root=True)
    N = 2 * S - 1
    names_arr = [""] * N
    dist_arr = np.zeros(N)
    parent_arr = -np.ones(N, dtype=np.int32)
    children_arr = -np.ones((N, 2), dtype=np.int32)

    def fill(rec, par):
        nid, name, d = rec[0], rec[1], rec[2]
        names_arr[nid] = name
        dist_arr[nid] = d
        parent_arr[nid] = par
        if len(rec) == 4:
            kid_ids = sorted(k[0] for k in rec[3])  # children ordered by node id
            children_arr[nid, :] = kid_ids
            for k in rec[3]:
                fill(k, nid)

    fill(root_rec, -1)
    # root written without a branch length -> parses as 0
    nid, name, d, kids = root_rec
    newick = "(" + ",".join(render(k) for k in kids) + ")" + name + ";"

    # the reference parses branch lengths through (float)atof (newick.cpp:185,300,325)
    dist_arr = dist_arr.astype(np.float32).astype(np.float64)

    def tips_under(n):
        if children_arr[n, 0] < 0:
            return [n]
        return tips_under(children_arr[n, 0]) + tips_under(children_arr[n, 1])

    root = N - 1
    c0, c1 = children_arr[root]
    t0, t1 = tips_under(c0), tips_under(c1)
    out_tips = t0 if len(t0) == n_out else t1
    in_root = c1 if len(t0) == n_out else c0
    # targets: pick clades of the ingroup greedily until n_target tips are covered
    in_tips = tips_under(in_root)
    cands = [n for n in range(S, N) if n != root and set(tips_under(n)) <= set(in_tips)]
    cands = [n for n in cands if 2 <= len(tips_under(n)) <= max(2, n_target // 2)]
    rng.shuffle(cands)
    targets: List[int] = []
    for n in cands:
        tt = tips_under(n)
        if set(tt) & set(targets):
            continue
        if len(targets) + len(tt) <= n_target:
            targets += tt
        if len(targets) >= n_target:
            break
    rest = [t for t in in_tips if t not in targets]
    while len(targets) < n_target and rest:
        targets.append(rest.pop(int(rng.integers(len(rest)))))
    conserve = [t for t in in_tips if t not in targets]

    pi, Q = make_rev_matrix(rng)
    return SynthTree(S=S, N=N, parent=parent_arr, children=children_arr, dist=dist_arr,
                     names=names_arr, newick=newick, outgroup=sorted(out_tips),
                     targets=sorted(targets), conserve=sorted(conserve), pi=pi, Q=Q)


def mod_text(tree: SynthTree) -> str:
    """.mod file in the phyloFit layout the reference reads (newick.cpp:454-489)."""
    lines = ["ALPHABET: A C G T ", "ORDER: 0", "SUBST_MOD: REV", "TRAINING_LNL: -1.0",
             "BACKGROUND: " + " ".join(f"{x:.6f}" for x in tree.pi) + " ", "RATE_MAT:"]
    for i in range(4):
        lines.append("  " + "   ".join(f"{x:.6f}" for x in tree.Q[i]) + " ")
    lines.append("TREE: " + tree.newick)
    return "\n".join(lines) + "\n"


def _transition_probs(Q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """P(t) for an array of times, via the real eigensystem of a reversible Q."""
    lam, U = np.linalg.eig(Q)
    lam, U = lam.real, U.real
    Uinv = np.linalg.inv(U)
    E = np.exp(np.multiply.outer(t, lam))  # [..., 4]
    P = np.einsum("ik,...k,kj->...ij", U, E, Uinv)
    return np.clip(P, 0.0, 1.0)


def _gbgc_probs(Q: np.ndarray, sel: np.ndarray, bias: np.ndarray, t: float) -> np.ndarray:
    """P(t) [c, 4, 4] under the GC-biased model of PhyloAcc-ST-GBGC for per-element selection coefficients and biases: every
    off-diagonal rate i -> j is multiplied by f(s + B_ij), f(x) = x / (1 - exp(-x)), B = +b for weak -> strong (A, T -> C, G), -b for
    strong -> weak (BPP_C::getRate, src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:13-30)."""
    strong = np.array([0, 1, 1, 0])
    sign = (strong[None, :] - strong[:, None]).astype(np.float64)  # [i, j]: +1 weak -> strong, -1 strong -> weak
    x = sel[:, None, None] + bias[:, None, None] * sign[None]
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(np.abs(x) < 1e-12, 1.0, x / (1.0 - np.exp(-x)))
    Qs = Q[None] * f
    idx = np.arange(4)
    Qs[:, idx, idx] = 0.0
    Qs[:, idx, idx] = -Qs.sum(axis=2)
    lam, U = np.linalg.eig(Qs)
    E = np.exp(lam * t)
    P = np.einsum("cik,ck,ckj->cij", U, E, np.linalg.inv(U)).real
    return np.clip(P, 0.0, 1.0)


@dataclasses.dataclass
class SynthBatch:
    tree: SynthTree
    X: np.ndarray  # uint8 [S, G] lower-case characters, elements concatenated
    elem_len: np.ndarray  # int32 [C]
    elem_start: np.ndarray  # int64 [C]
    c_rate: np.ndarray
    n_rate: np.ndarray
    scenario: np.ndarray  # 0 conserved, 1 accelerated on targets, 2 accelerated on a random clade
    acc_node: np.ndarray  # for scenario 2 the clade root, else -1

    @property
    def C(self) -> int:
        return int(self.elem_len.shape[0])


def simulate(tree: SynthTree, C: int = 500, length=200, seed: int = 1, gap_frac: float = 0.01,
             missing_frac: float = 0.03, ambig_frac: float = 0.0, lengths: Optional[Sequence[int]] = None,
             chunk: int = 20000, gbgc: Optional[dict] = None) -> SynthBatch:
    """Simulate ``C`` elements.  ``length`` is an int (fixed) or the string ``"cfg2"`` for
    lognormal(median 150, sigma 0.8) clipped to [50, 2000].
    ``gbgc`` switches to the GC-biased model: dict(shape=4, scale=0.8, frac=0.2) draws a bias b ~ Gamma(shape, scale)
    per element acting on a random `frac` of the branches; conserved / accelerated branches then carry SELECTION coefficients
    (-Gamma(5, 0.4) and Gamma(5, 0.4) - 2, the variant's priors) instead of rate multipliers."""
    rng = np.random.default_rng(seed)
    S, N = tree.S, tree.N
    if lengths is not None:
        elem_len = np.asarray(lengths, dtype=np.int32)
        C = len(elem_len)
    elif length == "cfg2":
        elem_len = np.clip(np.round(150.0 * np.exp(0.8 * rng.standard_normal(C))), 50, 2000).astype(np.int32)
    else:
        elem_len = np.full(C, int(length), dtype=np.int32)
    elem_start = np.concatenate([[0], np.cumsum(elem_len)[:-1]]).astype(np.int64)
    G = int(elem_len.sum())

    c_rate = rng.gamma(5.0, 0.04, size=C)
    n_rate = rng.gamma(10.0, 0.2, size=C)
    scenario = rng.choice(3, size=C, p=[0.6, 0.2, 0.2]).astype(np.int8)

    # per-node accelerated flags for each scenario
    def subtree_nodes(n):
        if tree.children[n, 0] < 0:
            return [n]
        return [n] + subtree_nodes(tree.children[n, 0]) + subtree_nodes(tree.children[n, 1])

    target_set = set(tree.targets)
    tips_of = {}

    def tips(n):
        if n in tips_of:
            return tips_of[n]
        if tree.children[n, 0] < 0:
            tips_of[n] = {n}
        else:
            tips_of[n] = tips(tree.children[n, 0]) | tips(tree.children[n, 1])
        return tips_of[n]

    acc_target = np.array([tips(n) <= target_set for n in range(N)])
    in_nodes = [n for n in range(S, N - 1) if not (tips(n) & set(tree.outgroup)) and len(tips(n)) <= 8]
    acc_node = np.full(C, -1, dtype=np.int32)
    clade_masks = {}
    for n in in_nodes:
        m = np.zeros(N, dtype=bool)
        m[subtree_nodes(n)] = True
        clade_masks[n] = m
    pick = rng.choice(in_nodes, size=C)
    acc_node[scenario == 2] = pick[scenario == 2]

    if gbgc is not None:
        g_bias = rng.gamma(gbgc.get("shape", 4.0), gbgc.get("scale", 0.8), size=C)
        g_on = rng.random((C, N)) < gbgc.get("frac", 0.2)
        g_selc = -rng.gamma(5.0, 0.4, size=C)
        g_seln = rng.gamma(5.0, 0.4, size=C) - 2.0
    X = np.empty((S, G), dtype=np.uint8)
    site_elem_all = np.repeat(np.arange(C, dtype=np.int64), elem_len)
    pi_cdf = np.cumsum(tree.pi)
    order = []  # pre-order traversal

    def pre(n):
        order.append(n)
        if tree.children[n, 0] >= 0:
            pre(tree.children[n, 0])
            pre(tree.children[n, 1])

    pre(N - 1)
    lam, U = np.linalg.eig(tree.Q)
    lam, U = lam.real, U.real
    Uinv = np.linalg.inv(U)

    # chunk over elements to bound memory
    for e0 in range(0, C, chunk):
        e1 = min(C, e0 + chunk)
        g0 = int(elem_start[e0])
        g1 = int(elem_start[e1 - 1] + elem_len[e1 - 1])
        site_elem = site_elem_all[g0:g1] - e0
        ng = g1 - g0
        states = np.empty((N, ng), dtype=np.int8)
        states[N - 1] = np.searchsorted(pi_cdf, rng.random(ng) * pi_cdf[-1]).clip(0, 3)
        for n in order[1:]:
            par = tree.parent[n]
            rate = c_rate[e0:e1].copy()
            is_acc = np.zeros(e1 - e0, dtype=bool)
            if acc_target[n]:
                is_acc |= scenario[e0:e1] == 1
            for cn in np.unique(acc_node[e0:e1]):
                if cn >= 0 and clade_masks[cn][n]:
                    is_acc |= acc_node[e0:e1] == cn
            rate[is_acc] = n_rate[e0:e1][is_acc]
            if gbgc is not None:
                P = _gbgc_probs(tree.Q, np.where(is_acc, g_seln[e0:e1], g_selc[e0:e1]), np.where(g_on[e0:e1, n], g_bias[e0:e1], 0.0), float(tree.dist[n]))
            else:
                E = np.exp(np.multiply.outer(rate * tree.dist[n], lam))  # [c,4]
                P = np.einsum("ik,ck,kj->cij", U, E, Uinv).clip(0.0, 1.0)  # [c,4,4] parent i -> child j
            cdf = np.cumsum(P, axis=2)
            row = cdf[site_elem, states[par]]  # [ng,4]
            u = rng.random(ng) * row[:, 3]
            states[n] = (u[:, None] >= row[:, :3]).sum(axis=1)
        chars = BASES[states[:S]]
        if ambig_frac > 0:
            amb = np.frombuffer(b"rykmsw", dtype=np.uint8)
            mask = rng.random(chars.shape) < ambig_frac
            chars[mask] = amb[rng.integers(0, 6, size=int(mask.sum()))]
        if gap_frac > 0:
            mask = rng.random(chars.shape) < gap_frac
            chars[mask] = ord("-")
            maskn = rng.random(chars.shape) < gap_frac * 0.1
            chars[maskn] = ord("n")
        if missing_frac > 0:
            miss = rng.random((S, e1 - e0)) < missing_frac
            chars[miss[:, site_elem]] = ord("-")
        X[:, g0:g1] = chars
    return SynthBatch(tree=tree, X=X, elem_len=elem_len, elem_start=elem_start, c_rate=c_rate,
                      n_rate=n_rate, scenario=scenario, acc_node=acc_node)


DEFAULT_PARAMS = dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1)


def write_files(batch: SynthBatch, outdir: str, prefix: str = "sim", ids: Optional[Sequence[int]] = None,
                params: Optional[dict] = None, upper: bool = False) -> str:
    """Write .mod / FASTA / BED / id / cfg in the reference's formats; returns the cfg path."""
    os.makedirs(outdir, exist_ok=True)
    tree = batch.tree
    mod = os.path.join(outdir, f"{prefix}.mod")
    fa = os.path.join(outdir, f"{prefix}.fa")
    bed = os.path.join(outdir, f"{prefix}.bed")
    idf = os.path.join(outdir, f"{prefix}.id")
    cfg = os.path.join(outdir, f"{prefix}.cfg")
    res = os.path.join(outdir, "out")
    os.makedirs(res, exist_ok=True)
    with open(mod, "w") as f:
        f.write(mod_text(tree))
    with open(fa, "wb") as f:
        for s in range(tree.S):
            f.write(b">" + tree.names[s].encode() + b"\n")
            row = batch.X[s].tobytes()
            f.write((row.upper() if upper else row) + b"\n")
    with open(bed, "w") as f:
        for c in range(batch.C):
            s0 = int(batch.elem_start[c])
            f.write(f"{c}\t{s0}\t{s0 + int(batch.elem_len[c])}\telem{c}\n")
    if ids is None:
        ids = range(batch.C)
    with open(idf, "w") as f:
        for i in ids:
            f.write(f"{i}\n")
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    with open(cfg, "w") as f:
        f.write(f"PHYTREE_FILE {mod}\nSEG_FILE {bed}\nALIGN_FILE {fa}\nID_FILE {idf}\n")
        f.write(f"RESULT_FOLDER {res}\nPREFIX {prefix}\n")
        for k, v in p.items():
            f.write(f"{k} {v}\n")
        f.write("TARGETSPECIES " + ";".join(tree.names[t] for t in tree.targets) + "\n")
        f.write("OUTGROUP " + ";".join(tree.names[t] for t in tree.outgroup) + "\n")
        f.write("CONSERVE " + ";".join(tree.names[t] for t in tree.conserve) + "\n")
    return cfg
