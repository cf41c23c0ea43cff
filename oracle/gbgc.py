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


def parse_logliks(text: str):
    """Replayed log-likelihoods of the `state` commands, in script order."""
    return [float.fromhex(line.split()[1]) for line in text.split("\n") if line.startswith("loglik ")]
