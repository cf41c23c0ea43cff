"""oracle/gbgc.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the per-branch matrices of PhyloAcc-ST-GBGC, pinned against the unmodified reference through
oracle/_ref/gbgc_probe (fixtures: tests/golden/gbgc43/probe.txt, generator tests/golden/make_gbgc_golden.py):

  get_rate(Q0, s, b)            BPP_C::getRate        (src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:13-30)
  log_transition(Q0, s, b, t)   BPP_C::getTransition  (src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:1405-1427)

Bases are ordered A, C, G, T; Q[i, j] is the rate i -> j; the bias matrix B is +b for weak -> strong (A, T -> C, G), -b for
strong -> weak and 0 otherwise; every off-diagonal rate is multiplied by f(s + B) with f(x) = x / (1 - exp(-x)), f(0) = 1."""
from __future__ import annotations

import numpy as np


def get_rate(Q0: np.ndarray, s: float, b: float) -> np.ndarray:
    if s == 0 and b == 0:
        return Q0.copy()  # bpp_c.cpp:15
    B = np.zeros((4, 4))
    B[0, 1] = B[0, 2] = B[3, 1] = B[3, 2] = b
    B[1, 0] = B[1, 3] = B[2, 0] = B[2, 3] = -b
    B = B + s
    with np.errstate(divide="ignore", invalid="ignore"):
        F = np.where(B == 0, 1.0, B / (1 - np.exp(-B)))
    Q = Q0 * F
    # the reference subtracts the row sums of the WHOLE matrix (old diagonal included) from the diagonal (bpp_c.cpp:27-28)
    c = Q.sum(axis=1)
    Q[np.diag_indices(4)] -= c
    return Q


def log_transition(Q0: np.ndarray, s: float, b: float, t: float) -> np.ndarray:
    """log of (eigeninv * diag(exp(lambda t)) * eigenvec) with eigenvec = V^T: element [i, j] = log P(j -> i | t)."""
    Q = get_rate(Q0, s, b)
    lam, V = np.linalg.eig(Q)
    lam, V = lam.real, V.real  # conv_to<mat>::from keeps the real parts (bpp_c.cpp:1412-1414); the spectrum is real (reversible Q)
    eigenvec = V.T
    eigeninv = np.linalg.inv(eigenvec)
    x = eigenvec * np.exp(lam * t)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(eigeninv @ x)


def transition_uniformised(Q0: np.ndarray, s: float, b: float, t: float) -> np.ndarray:
    """P(t) [i, j] = P(i -> j) by uniformisation + squaring: every term is non-negative, so small entries keep their relative
    accuracy -- the algorithm of the CUDA kernel (pa_gbgc.cu), here as an independent cross-check of the eigen route."""
    Q = get_rate(Q0, s, b)
    mu = max(-Q[i, i] for i in range(4))
    if mu <= 0 or t <= 0:
        return np.eye(4)
    k = 0
    while mu * t / (1 << k) > 0.5:
        k += 1
    tau = t / (1 << k)
    A = np.eye(4) + Q / mu
    term = np.eye(4) * np.exp(-mu * tau)
    P = term.copy()
    for j in range(1, 20):
        term = term @ A * (mu * tau / j)
        P = P + term
    for _ in range(k):
        P = P @ P
    return P


def replay_loglik(Q0, pi, el, Z, B, sel_c, sel_n, gbc, uniformised=False):
    """Replayed log-likelihood of one element under the GC-biased model: per-branch P(t) from the six rate matrices of
    BPP_C::initMCMC (GBGC/SRC/bpp_c.cpp:196-204: class 0 / 1 / 2 -> s = 0 / sel_c / sel_n, B -> bias gbc) through init_cache_TM
    (:1563-1621), the upward pass Update_Tg(g, ., ., false) (:477-528) column by column, sum_g log sum_b pi_b L_root(b).
    `el` is the oracle's element front-end (oracle/frontend.py: masks [GG][S], nodes, children2, distances2, missing)."""
    N = len(Z)
    P = {}
    for s in el.nodes:
        if s == el.root:
            continue
        sel = sel_n if Z[s] == 2 else (0.0 if Z[s] == 0 else sel_c)
        bias = gbc if B[s] else 0.0
        if uniformised:
            P[s] = transition_uniformised(Q0, sel, bias, float(el.distances2[s]))
        else:
            with np.errstate(over="ignore"):
                P[s] = np.exp(log_transition(Q0, sel, bias, float(el.distances2[s]))).T  # [i, j] = P(i -> j)
    S = el.masks.shape[1]
    total = 0.0
    for g in range(el.masks.shape[0]):
        L = {}
        for s in el.nodes:
            if s < S:
                m = int(el.masks[g, s])
                L[s] = np.array([(m >> b) & 1 for b in range(4)], dtype=np.float64)
                continue
            if el.missing[s]:
                continue
            v = np.ones(4)
            for ch in el.children2[s]:
                if ch < 0 or el.missing[ch]:
                    continue
                v = v * (P[ch] @ L[ch])
            L[s] = v
        total += np.log(float(np.dot(pi, L[el.root])))
    return total


def parse_probe(text: str):
    out, cur = [], None
    Q0 = None
    for line in text.split("\n"):
        f = line.split()
        if not f:
            continue
        if f[0] == "Q":
            Q0 = np.array([float.fromhex(x) for x in f[1:]]).reshape(4, 4)
        elif f[0] == "tm":
            cur = dict(node=int(f[1]), sel=float.fromhex(f[2]), bias=float.fromhex(f[3]), t=float.fromhex(f[5]))
            out.append(cur)
        elif f[0] == "rate":
            cur["rate"] = np.array([float.fromhex(x) for x in f[1:]]).reshape(4, 4)
        elif f[0] == "logtm":
            cur["logtm"] = np.array([float.fromhex(x) for x in f[1:]]).reshape(4, 4)
    return Q0, out


def parse_integ(text: str):
    """The `integ` dumps of gbgc_probe, in script order: dict(resZ, elbo, loglik, Z, fixZ, post {s: 4x4}, emis {s: 6}, tmint {s: 3x3},
    rootback [6], consts (nGBprior_c, consToMis, nconsToMis))."""
    out, cur = [], None
    hx = float.fromhex
    for line in text.split("\n"):
        f = line.split()
        if not f:
            continue
        if f[0] == "integ" and len(f) >= 6:
            cur = dict(resZ=int(f[1]), elbo=hx(f[3]), loglik=hx(f[5]), post={}, emis={}, tmint={})
            out.append(cur)
        elif cur is None:
            continue
        elif f[0] == "Zused":
            cur["Z"] = [int(x) for x in f[1:]]
        elif f[0] == "fixZ":
            cur["fixZ"] = [int(x) for x in f[1:]]
        elif f[0] == "post":
            cur["post"][int(f[1])] = np.array([hx(x) for x in f[2:]]).reshape(4, 4)
        elif f[0] == "emis":
            cur["emis"][int(f[1])] = np.array([hx(x) for x in f[2:]])
        elif f[0] == "tmint":
            cur["tmint"][int(f[1])] = np.array([hx(x) for x in f[2:]]).reshape(3, 3)
        elif f[0] == "rootback":
            cur["rootback"] = np.array([hx(x) for x in f[1:]])
        elif f[0] == "consts":
            cur["consts"] = tuple(hx(x) for x in f[1:])
    return out


def parse_logliks(text: str):
    """Replayed log-likelihoods of the `state` commands, in script order."""
    return [float.fromhex(line.split()[1]) for line in text.split("\n") if line.startswith("loglik ")]


def expected_pairs(Q0, pi, el, Z, B, sel_c, sel_n, gbc):
    """Expected (child base i, parent base j) pairs per branch summed over the columns, given the data and the state: the downward pass
    of Update_Tg(g, ., ., true) (GBGC/SRC/bpp_c.cpp:530-583) in linear space.  With L_s the upward partial of node s, m_s = P_s L_s its
    message and q_p the posterior of the parent's base, the joint is q_p(j) P_s(j -> i) L_s(i) / m_s(j).  Returns {node: 4x4 [i, j]}."""
    P = {}
    for s in el.nodes:
        if s == el.root:
            continue
        sel = sel_n if Z[s] == 2 else (0.0 if Z[s] == 0 else sel_c)
        P[s] = transition_uniformised(Q0, sel, gbc if B[s] else 0.0, float(el.distances2[s]))
    S = el.masks.shape[1]
    post = {s: np.zeros((4, 4)) for s in el.nodes if s != el.root}
    order = list(el.nodes)
    for g in range(el.masks.shape[0]):
        L, M = {}, {}
        for s in order:
            if s < S:
                m = int(el.masks[g, s])
                L[s] = np.array([(m >> b) & 1 for b in range(4)], dtype=np.float64)
            elif not el.missing[s]:
                v = np.ones(4)
                for ch in el.children2[s]:
                    if ch < 0 or el.missing[ch]:
                        continue
                    v = v * M[ch]
                L[s] = v
            if s != el.root and not el.missing[s]:
                M[s] = P[s] @ L[s]
        q = {}
        r = pi * L[el.root]
        q[el.root] = r / r.sum()
        for s in reversed(order[:-1]):
            if el.missing[s]:
                continue
            p = int(el.parent2[s])
            joint = (q[p] / M[s])[None, :] * P[s].T * L[s][:, None]  # [i (child), j (parent)]
            post[s] += joint
            q[s] = joint.sum(axis=1)
    return post


def integrate_zb(post, P_logs, el, fixZ, tmint, consts):
    """initial_prob_back (GBGC/SRC/bpp_c2.cpp:354-407) given the expected pairs: emissions of the six (Z, B) classes of every branch
    (class k = Z + 3 B: accu(log TM_k % pairs) + log prior of B + log(1 - missing rate)), the 6-state upward messages through the
    3 x 3 Z transition matrices, and log_prob_back[root][0].  P_logs[s][k] is log P^T of class k ((child, parent) layout).
    Returns (emissions {s: [6]}, root message [6])."""
    c_gb, c2m, nc2m = consts
    back = {s: np.zeros(6) for s in el.nodes}
    emis = {}
    for s in el.nodes:
        if s == el.root:
            continue
        e = np.zeros(6)
        if el.missing[s]:
            if not el.missing[int(el.parent2[s])]:  # initMCMC, GBGC/SRC/bpp_c.cpp:219-229
                e[[0, 2, 3, 5]] = np.log(nc2m)
                e[[1, 4]] = np.log(c2m)
            if fixZ[s] == 1:
                e[[2, 5]] = -np.inf
        else:
            for k in range(6):
                if fixZ[s] == 1 and k % 3 == 2:
                    e[k] = -np.inf
                    continue
                with np.errstate(invalid="ignore"):
                    prod = np.where(post[s] == 0, 0.0, P_logs[s][k] * post[s])
                e[k] = prod.sum() + (np.log(1 - c_gb) if k < 3 else np.log(c_gb)) + (np.log(1 - c2m) if k % 3 == 1 else np.log(1 - nc2m))
        emis[s] = e
        back[s] = e.copy()
    for s in el.nodes[:-1]:
        p = int(el.parent2[s])
        t = np.vstack([tmint[s], tmint[s]])  # rows: the child's six (Z, B) classes, columns: the parent's Z
        x = t + back[s][:, None]
        mx = x.max()
        msg = np.full(3, -np.inf) if mx == -np.inf else np.log(np.exp(x - mx).sum(axis=0)) + mx
        back[p][:3] += msg
        back[p][3:] += msg
    return emis, back[el.root]
