"""TEST INFRASTRUCTURE -- numpy restatement of what PhyloAcc-GT evaluates for one element at a given gene tree (SURVEY.md §8 f2 groundwork):
the coalescent prior of a gene tree embedded in the species tree, log P(missing | Z, gene tree), and the log-likelihood of the alignment on
that gene tree.  Pinned against oracle/_ref/gt_probe (the unmodified reference) through tests/golden/gt12.

The reference walks the embedded tree species by species (GTree::Update_Tg, src/PhyloAcc-GT/genetree.cpp:1330-1418): a gene-tree branch is
cut where it crosses a species boundary and every piece evolves at the rate of the species branch it lies in (1, c_rate or n_rate by Z).
All pieces share one rate matrix, so their transition matrices commute and compose: the branch as a whole has P(sum_k rate_k * len_k).
The restatement therefore prunes the gene tree ONCE with these effective lengths -- the form the device kernel uses too."""
from __future__ import annotations

import numpy as np


def parse_probe(text: str):
    hdr, elements = {}, []
    cur_el, st = None, None
    for line in text.split("\n"):
        f = line.split()
        if not f:
            continue
        k = f[0]
        if k == "gt":
            hdr["N"], hdr["S"] = int(f[2]), int(f[4])
        elif k in ("heights", "thetas", "logpi", "Q") and st is None and cur_el is None:
            hdr[k] = np.array([float.fromhex(x) for x in f[1:]])
        elif k == "parent":
            hdr["parent"] = np.array([int(x) for x in f[1:]])
        elif k == "element":
            cur_el = dict(c=int(f[1]), filter=int(f[3]), GG=int(f[5]), states=[])
            elements.append(cur_el)
        elif k == "state":
            st = dict(resZ=int(f[2]), coal=int(f[4]), root=int(f[6]), cross={}, coalesce={})
        elif st is not None:
            if k == "Z":
                st["Z"] = np.array([int(x) for x in f[1:]])
            elif k == "rates":
                st["n_rate"], st["c_rate"] = float.fromhex(f[1]), float.fromhex(f[2])
            elif k == "consmis":
                st["cons_to_mis"], st["ncons_to_mis"] = float.fromhex(f[1]), float.fromhex(f[2])
            elif k == "gchild":
                st["gchild"] = np.array([int(x) for x in f[1:]]).reshape(-1, 2)
            elif k == "gparent":
                st["gparent"] = np.array([int(x) for x in f[1:]])
            elif k == "gheights":
                st["gheights"] = np.array([float.fromhex(x) for x in f[1:]])
            elif k == "gmissing":
                st["gmissing"] = np.array([int(x) for x in f[1:]], dtype=bool)
            elif k == "cross":  # species -> [(gene lineage, where it goes: itself = leaves the species uncoalesced, else the coalescence node)]
                v = [int(x) for x in f[2:]]
                st["cross"][int(f[1])] = list(zip(v[0::2], v[1::2]))
            elif k == "coal":
                st["coalesce"][int(f[1])] = [int(x) for x in f[2:]]
            elif k == "tg":  # sampled history: species -> {gene lineage: base at the bottom of its piece in that species, per column}
                st.setdefault("tg", {}).setdefault(int(f[1]), {})[int(f[2])] = np.array([int(x) for x in f[3:]], dtype=np.int8)
            elif k == "amb":  # an ambiguity-coded tip column: the allowed bases
                st.setdefault("amb", {})[(int(f[1]), int(f[2]))] = np.array([int(x) for x in f[3:7]])
            elif k == "emission":
                st.setdefault("emission", {})[int(f[1])] = np.array([float.fromhex(x) for x in f[2:5]])
            elif k in ("loglik", "logpz", "prior"):
                st[k] = float.fromhex(f[1])
            elif k == "end":
                cur_el["states"].append(st)
                st = None
    return hdr, elements


def species_heights(parent, dist, N):
    """BPP::InitPhyloTree (PhyloAcc-GT/bpp.cpp:81-84): the root at 0, every node below its parent by its branch length."""
    h = np.zeros(N)
    for i in range(N - 2, -1, -1):
        h[i] = h[parent[i]] - dist[i]
    return h


def prior_tree(st, hdr):
    """GTree::priorTree (genetree.cpp:560-585): the multispecies-coalescent density of the coalescence times inside every species branch."""
    N, S, h, th, par = hdr["N"], hdr["S"], hdr["heights"], hdr["thetas"], hdr["parent"]
    logp = 0.0
    for s in range(S, N):
        lin = st["cross"].get(s, [])
        co = st["coalesce"].get(s, [])
        nn = len(lin) - len(co)
        if nn == 0:
            continue
        cur = h[s]
        for g in co:
            logp -= (st["gheights"][g] - cur) * (nn - 1) * nn / th[s]
            logp += np.log(2) - np.log(th[s])
            nn -= 1
            cur = st["gheights"][g]
        if nn > 1 and par[s] != N:
            logp -= (h[par[s]] - cur) * (nn - 1) * nn / th[s]
    return logp


def logp_missing(st, hdr):
    """GTree::get_logpZ1 (genetree.cpp:1156-1190): every lineage that starts inside a species branch is lost there with probability
    consToMis (conserved) or nconsToMis (otherwise)."""
    N, S, h = hdr["N"], hdr["S"], hdr["heights"]
    Z, mis, gh = st["Z"], st["gmissing"], st["gheights"]
    lc, ln = st["cons_to_mis"], st["ncons_to_mis"]

    def term(s, m):
        p = lc if Z[s] == 1 else ln
        return np.log(p) if m else np.log(1 - p)

    v = sum(term(s, mis[s]) for s in range(S))
    for s in range(S, N):
        for first, _ in st["cross"].get(s, []):
            if gh[first] >= h[s]:
                v += term(s, mis[first])
    return v


def effective_lengths(st, hdr):
    """sum over the species a gene-tree branch crosses of rate(Z[species]) x (length of the piece inside it): the segments of
    GTree::Update_Tg (genetree.cpp:1352-1358 inside a species up to a coalescence, 1385-1388 up to the species' top)."""
    N, h, par = hdr["N"], hdr["heights"], hdr["parent"]
    gh, Z = st["gheights"], st["Z"]
    eff = np.zeros(N)
    for ss, lin in st["cross"].items():
        rate = st["n_rate"] if Z[ss] == 2 else (st["c_rate"] if Z[ss] == 1 else 1.0)
        for gs in st["coalesce"].get(ss, []):
            for chi in st["gchild"][gs]:
                eff[chi] += rate * (gh[gs] - max(gh[chi], h[ss]))
        sp = par[ss]
        if sp < N:
            for first, second in lin:
                if first == second:
                    eff[first] += rate * (h[sp] - max(gh[first], h[ss]))
    return eff


def transition(Q, t):
    """exp(Q t) through the eigensystem of Q (real spectrum: the .mod matrices are reversible to print precision)."""
    w, V = np.linalg.eig(Q)
    return np.real((V * np.exp(w * t)) @ np.linalg.inv(V))


def replay_loglik(st, hdr, masks):
    """log P(X | gene tree, Z, rates): pruning over the gene tree with the effective lengths.  masks uint8 [GG, S]: bit b set <=> base b
    allowed at that tip (all four set for a gap / N: the child is skipped, i.e. contributes 1)."""
    N, S = hdr["N"], hdr["S"]
    Q, pi = hdr["Q"].reshape(4, 4), np.exp(hdr["logpi"])
    eff = effective_lengths(st, hdr)
    gc, mis = st["gchild"], st["gmissing"]
    GG = masks.shape[0]
    part = {}

    def up(x):
        if x < S:
            m = masks[:, x].astype(np.int64)
            return np.stack([(m >> b) & 1 for b in range(4)], axis=1).astype(float)
        out = np.ones((GG, 4))
        for chi in gc[x]:
            if chi < 0 or mis[chi]:
                continue
            if chi not in part:
                part[chi] = up(chi)
            out *= part[chi] @ transition(Q, eff[chi]).T
        return out

    root = up(st["root"])
    return float(np.sum(np.log(root @ pi)))


def _lse(v):
    m = np.max(v)
    return -np.inf if m == -np.inf else float(np.log(np.sum(np.exp(v - m))) + m)


def emission(st, hdr, GG):
    """BPP_C::getEmission of PhyloAcc-GT (bpp_c.cpp:1043-1216): log P(what happened inside species branch ss | Z[ss] = k) for k = background,
    conserved, accelerated, GIVEN one sampled history (st["tg"]): the loss / no-loss terms of the lineages that start in the branch, and for
    every column the probability of the sampled states at the top of the branch given the states at its bottom, with the coalescence nodes
    INSIDE the branch summed out by a small pruning pass.  GG = number of columns.  Returns {species: [3]}."""
    N, S, h, par = hdr["N"], hdr["S"], hdr["heights"], hdr["parent"]
    Q = hdr["Q"].reshape(4, 4)
    gh, mis, gc, tg = st["gheights"], st["gmissing"], st["gchild"], st["tg"]
    rates = (1.0, st["c_rate"], st["n_rate"])
    lc, ln = st["cons_to_mis"], st["ncons_to_mis"]
    amb = st.get("amb", {})  # (tip, column) -> allowed bases, in the history's own column order
    out = {}

    def log_tm(length, rate):  # [child base i, parent base j] = log P(j -> i), as BPP::getlogTMc lays it out
        if length < 0:
            length = 1e-9
        with np.errstate(divide="ignore"):
            return np.log(transition(Q, rate * length)).T

    for ss in sorted(st["cross"]):
        lin = st["cross"][ss]
        sp = int(par[ss])
        em = np.zeros(3)
        for first, _ in lin:  # bpp_c.cpp:1083-1117
            if first < S or gh[first] >= h[ss]:
                if mis[first]:
                    em += np.log([ln, lc, ln])
                else:
                    em += np.log([1 - ln, 1 - lc, 1 - ln])
        base = {}
        for gs, _ in lin:  # lineages that enter the branch from below: indicator of their sampled state (or the tip's own vector)
            if gh[gs] <= h[ss]:
                tl = np.zeros((GG, 4))
                if not mis[gs]:
                    t = tg[ss][gs]
                    for g in range(GG):
                        if t[g] >= 4:
                            continue
                        tl[g] = -np.inf
                        if t[g] == -1:
                            with np.errstate(divide="ignore"):
                                tl[g] = np.log(amb[(gs, g)].astype(float))
                        elif ss >= S:
                            tl[g, t[g]] = 0.0
                base[gs] = tl
        for rr in range(3):
            e = em[rr]
            templ = dict(base)
            for gs in st["coalesce"].get(ss, []):
                acc = np.zeros((GG, 4))
                for chi in gc[gs]:
                    if mis[chi]:
                        continue
                    tm = log_tm(gh[gs] - max(gh[chi], h[ss]), rates[rr])
                    t = tg[ss][chi]
                    for g in range(GG):
                        if t[g] >= 4:
                            continue
                        x = tm + templ[chi][g][:, None]
                        m = x.max()
                        acc[g] += np.full(4, -np.inf) if m == -np.inf else np.log(np.exp(x - m).sum(axis=0)) + m
                templ[gs] = acc
            if sp < N:
                for first, second in lin:
                    if first != second or mis[first]:
                        continue
                    tm = log_tm(h[sp] - max(gh[first], h[ss]), rates[rr])
                    tb, tt = tg[ss][first], tg[sp][first]
                    for g in range(GG):
                        if tb[g] >= 4:
                            continue
                        if ss >= S or tb[g] == -1:
                            e += _lse(tm[:, tt[g]] + templ[first][g])
                        else:
                            e += tm[tb[g], tt[g]]
            else:
                root = st["root"]
                if not mis[ss]:
                    t = tg[ss][root]
                    for g in range(GG):
                        if t[g] < 4:
                            e += templ[root][g][t[g]]
            em[rr] = e
        out[ss] = em
    return out
