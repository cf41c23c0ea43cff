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


# ------------------------------------------------------------------------------------------------------------------------------
# The PhyloAcc-ST-GBGC chain (BPP_C::Gibbs, GBGC/SRC/bpp_c.cpp:259-345) restated on the counter-based draw schedule of the CUDA kernel
# (pa_gbgc_chain.cu), so that the two can be compared draw for draw.  The reference's own random STREAM (GSL mt19937) is not
# reproduced; its DISTRIBUTIONS are (tests/test_gbgc_chain.py compares with runs of the unmodified binary across seeds).

ST_GBGC = 4  # Philox stream of the GBGC chain: (idx, sub, iteration, ST_GBGC); sub 0 = rate moves, 1 = (Z, B) draws, 2 = the final accept


def philox(seed, key, idx, sub, it, stream):
    c = [idx & 0xFFFFFFFF, sub & 0xFFFFFFFF, it & 0xFFFFFFFF, stream & 0xFFFFFFFF]
    k0, k1 = seed & 0xFFFFFFFF, key & 0xFFFFFFFF
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c


def u01(w):
    return (float(w) + 0.5) * 2.3283064365386963e-10


def log_gamma_pdf(x, a, b):
    """log(gsl_ran_gamma_pdf(x, a, b)): the density first, then its log (GSL randist/gamma.c), -inf outside the support."""
    import math

    if x < 0:
        return -math.inf
    if x == 0:
        return math.log(1 / b) if a == 1 else -math.inf
    if a == 1:
        d = math.exp(-x / b) / b
    else:
        d = math.exp((a - 1) * math.log(x / b) - x / b - math.lgamma(a)) / b
    return math.log(d) if d > 0 else -math.inf


class GbgcParams:
    """Defaults of GBGC/SRC/main.cpp:50-86 and bpp.hpp:201, bpp.cpp:217-219, bpp_c.hpp:358-361."""

    def __init__(self, **kw):
        self.sel_c0, self.sel_n0, self.gbc0 = -2.0, 0.0, 3.0  # CONSERVE_RATE, ACCE_RATE, cur_nGB
        self.cprior_a, self.cprior_b = 5.0, 0.4
        self.nprior_a, self.nprior_b, self.nprior_c = 5.0, 0.4, 2.0
        self.gbprior_a, self.gbprior_b, self.gbprior_c = 4.0, 0.8, 0.2
        self.cons_to_mis, self.ncons_to_mis = 0.01, 0.1
        self.grate, self.lrate = 0.8, 0.1  # PREP_GRATE, PREP_LRATE
        self.prop_c, self.prop_n, self.prop_gb = 1.2, 1.2, 0.2
        self.num_burn, self.num_mcmc = 200, 800
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


def gbgc_fixz(el, N, S, resZ, targets):
    """BPP_C::initMCMC (GBGC/SRC/bpp_c.cpp:143-180): fixZ per model.  resZ 0 = M0 (every branch), 2 = M1 (outgroup lineage and every tip
    but the target species), 1 = M2 (outgroup lineage); the root is 0."""
    fixZ = -np.ones(N, dtype=np.int8)
    fixZ[el.upper_c] = 1
    fixZ[el.root] = 0
    if resZ == 2:
        fixZ[:S] = 1
        fixZ[list(targets)] = -1
    elif resZ == 0:
        fixZ[el.nodes[:-1]] = 1
    return fixZ


def _tm_int(el, N, upper, grate, lrate):
    """log TM_Int [child Z, parent Z] of every node (GBGC/SRC/bpp.hpp:278-324, bpp_c.hpp:367-377)."""
    out = {}
    with np.errstate(divide="ignore"):
        for s in el.nodes:
            T = np.zeros((3, 3))
            T[0, 0], T[1, 0] = 1 - grate, grate
            T[1, 1], T[2, 1] = 1 - lrate, lrate
            T[2, 2] = 1.0
            if s in upper:
                T[1, 1], T[2, 1] = 1.0, 0.0
            out[s] = np.log(T)
    return out


def _uniformised_batch(Q, t):
    """exp(Q[m] t[m]) for M matrices at once, the same algorithm as transition_uniformised (and pa_gbgc.cuh: expm_uniformised)."""
    M = Q.shape[0]
    mu = np.max(-Q[:, np.arange(4), np.arange(4)], axis=1)
    ok = (mu > 0) & (t > 0)
    mu_s = np.where(ok, mu, 1.0)
    x = mu_s * np.where(ok, t, 0.0)
    k = np.zeros(M, dtype=int)
    while True:
        big = x > 0.5
        if not big.any():
            break
        x = np.where(big, x * 0.5, x)
        k += big
    A = Q / mu_s[:, None, None] + np.eye(4)[None]
    term = np.eye(4)[None] * np.exp(-x)[:, None, None]
    P = term.copy()
    for j in range(1, 20):
        term = np.einsum("mij,mjk->mik", term, A) * (x / j)[:, None, None]
        P = P + term
    for q in range(int(k.max()) if M else 0):
        sq = np.einsum("mij,mjk->mik", P, P)
        P = np.where((k > q)[:, None, None], sq, P)
    P[~ok] = np.eye(4)
    return P


class GbgcChain:
    """One (element, model) chain.  Class k = Z + 3 B of a branch: selection 0 / sel_c / sel_n by k % 3, bias gbc if k >= 3."""

    def __init__(self, Q0, pi, el, N, S, resZ, targets, upper, params, seed, key):
        self.Q0, self.pi, self.el, self.N, self.S, self.resZ, self.par = np.asarray(Q0, float), np.asarray(pi, float), el, N, S, resZ, params
        self.seed, self.key = seed, key
        self.nodes = list(el.nodes)
        self.br = self.nodes[:-1]  # branches = non-root nodes, post-order
        self.local = {s: i for i, s in enumerate(self.nodes)}
        self.fixZ = gbgc_fixz(el, N, S, resZ, targets)
        self.tmint = _tm_int(el, N, upper, params.grate, params.lrate)
        self.t = np.array([float(el.distances2[s]) for s in self.br])
        m = el.masks.astype(np.int64)  # [GG, S]
        self.tipL = {s: np.stack([(m[:, s] >> b) & 1 for b in range(4)], axis=1).astype(float) for s in self.nodes if s < S}

    # ---- matrices
    def class_logs(self, sel_c, sel_n, gbc, classes):
        """{class k: log P^T [branch index][child, parent]} for the listed classes."""
        out = {}
        for k in classes:
            sel = (0.0, sel_c, sel_n)[k % 3]
            Q = get_rate(self.Q0, sel, gbc if k >= 3 else 0.0)
            P = _uniformised_batch(np.repeat(Q[None], len(self.br), axis=0), self.t)
            with np.errstate(divide="ignore"):
                out[k] = np.log(np.transpose(P, (0, 2, 1)))
        return out

    # ---- exact likelihood given (Z, B) and the cached matrices: Update_Tg(g, ., ., false), all columns at once
    def upward(self, Z, B, logs):
        el, L, M = self.el, {}, {}
        for s in self.nodes:
            if s < self.S:
                L[s] = self.tipL[s]
            elif not el.missing[s]:
                v = np.ones((el.masks.shape[0], 4))
                for ch in el.children2[s]:
                    if ch < 0 or el.missing[ch]:
                        continue
                    v = v * M[ch]
                L[s] = v
            if s != el.root and not el.missing[s]:
                P = np.exp(logs[int(Z[s]) + 3 * int(B[s])][self.local[s]]).T  # [parent j, child i]
                M[s] = L[s] @ P.T  # [g, j] = sum_i P(j -> i) L(i)
        ll = float(np.log(L[el.root] @ self.pi).sum())
        return ll, L, M

    # ---- expected (child, parent) pairs: Update_Tg(g, ., ., true)
    def downward(self, Z, B, logs, L, M):
        el = self.el
        post = {s: np.zeros((4, 4)) for s in self.br}
        r = L[el.root] * self.pi[None]
        q = {el.root: r / r.sum(axis=1, keepdims=True)}
        for s in reversed(self.br):
            if el.missing[s]:
                continue
            p = int(el.parent2[s])
            P = np.exp(logs[int(Z[s]) + 3 * int(B[s])][self.local[s]])  # [child i, parent j]
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = np.where(M[s] > 0, q[p] / M[s], 0.0)  # [g, j]
            joint = ratio[:, None, :] * P[None] * L[s][:, :, None]  # [g, i, j]
            post[s] = joint.sum(axis=0)
            q[s] = joint.sum(axis=2)
        return post

    def emissions(self, post, logs, base=None):
        """Six-class emissions of every branch (initial_prob_back / getEmission_update): classes missing from `logs` keep `base`."""
        par, el = self.par, self.el
        out = {}
        for bi, s in enumerate(self.br):
            e = np.zeros(6) if base is None else base[s].copy()
            if el.missing[s]:
                if base is None:
                    if not el.missing[int(el.parent2[s])]:
                        e[[0, 2, 3, 5]] = np.log(par.ncons_to_mis)
                        e[[1, 4]] = np.log(par.cons_to_mis)
                    if self.fixZ[s] == 1:
                        e[[2, 5]] = -np.inf
            else:
                for k in logs:
                    if self.fixZ[s] == 1 and k % 3 == 2:
                        e[k] = -np.inf
                        continue
                    lp = logs[k][bi]
                    with np.errstate(invalid="ignore"):
                        e[k] = np.where(post[s] == 0, 0.0, lp * post[s]).sum()
                    e[k] += (np.log(1 - par.gbprior_c) if k < 3 else np.log(par.gbprior_c))
                    e[k] += (np.log(1 - par.cons_to_mis) if k % 3 == 1 else np.log(1 - par.ncons_to_mis))
            out[s] = e
        return out

    def messages(self, emis):
        el = self.el
        back = {s: (emis[s].copy() if s != el.root else np.zeros(6)) for s in self.nodes}
        for s in self.br:
            p = int(el.parent2[s])
            x = np.vstack([self.tmint[s], self.tmint[s]]) + back[s][:, None]
            mx = x.max()
            with np.errstate(divide="ignore"):
                msg = np.full(3, -np.inf) if mx == -np.inf else np.log(np.exp(x - mx).sum(axis=0)) + mx
            back[p][:3] += msg
            back[p][3:] += msg
        return back

    def elbo(self, Z, B, post, logs):
        tot = 0.0
        for bi, s in enumerate(self.br):
            if self.el.missing[s]:
                continue
            lp = logs[int(Z[s]) + 3 * int(B[s])][bi]
            with np.errstate(invalid="ignore"):
                tot += float(np.where(post[s] == 0, 0.0, lp * post[s]).sum())
        return tot

    def prior_z(self, Z):
        return float(sum(self.tmint[s][int(Z[s]), int(Z[int(self.el.parent2[s])])] for s in reversed(self.br)))

    def draw_zb(self, Z, B, back, it):
        """Update_ZB's top-down draws (GBGC/SRC/bpp_c.cpp:1700-1762)."""
        el = self.el
        Z, B = Z.copy(), B.copy()
        for s in reversed(self.br):
            p = int(el.parent2[s])
            u = u01(philox(self.seed, self.key, self.local[s], 1, it, ST_GBGC)[0])
            if Z[p] != 2:
                lw = np.concatenate([self.tmint[s][:, int(Z[p])]] * 2) + back[s]
                w = np.exp(lw - lw.max())
                nn = _categorical(w, u)
                Z[s] = nn % 3
                if not el.missing[s]:
                    B[s] = 1 if nn > 2 else 0
            else:
                Z[s] = 2
                lw = np.array([back[s][2], back[s][5]])
                w = np.exp(lw - lw.max())
                nn = _categorical(w, u)
                if not el.missing[s]:
                    B[s] = 1 if nn > 0 else 0
        return Z, B

    def full_loglik(self, loglik, logp_z, Z, B, sel_c, sel_n, gbc):
        """MonitorChain (GBGC/SRC/bpp_c.cpp:615-652)."""
        par, el = self.par, self.el
        v = loglik + logp_z + log_gamma_pdf(-sel_c, par.cprior_a, par.cprior_b)
        if self.resZ != 0:
            v += log_gamma_pdf(sel_n + par.nprior_c, par.nprior_a, par.nprior_b)
        mz = [0, 0, 0, 0]
        gb = ungb = 0
        for s in self.br:
            if el.missing[int(el.parent2[s])] if int(el.parent2[s]) < self.N else False:
                continue
            mz[2 * int(el.missing[s]) + int(Z[s] == 1)] += 1
            if el.missing[s]:
                continue
            if B[s]:
                gb += 1
            else:
                ungb += 1
        v += np.log(1 - par.ncons_to_mis) * mz[0] + np.log(1 - par.cons_to_mis) * mz[1] + np.log(par.ncons_to_mis) * mz[2] + np.log(par.cons_to_mis) * mz[3]
        v += np.log(1 - par.gbprior_c) * ungb + log_gamma_pdf(gbc, par.gbprior_a, par.gbprior_b) + np.log(par.gbprior_c) * gb
        return float(v)

    def run(self, n_iter=None):
        par, el, N = self.par, self.el, self.N
        T = par.num_burn + par.num_mcmc if n_iter is None else n_iter
        Z = -np.ones(N, dtype=np.int8)
        Z[self.nodes] = 0
        B = np.zeros(N, dtype=np.int8)
        sel_c, sel_n, gbc = par.sel_c0, par.sel_n0, par.gbc0
        logs = self.class_logs(sel_c, sel_n, gbc, range(6))
        # log P(constraints): the 6-state pass of initMCMC (GBGC/SRC/bpp_c.cpp:232-254) with "- log 2" per message
        back0 = {s: np.zeros(6) for s in self.nodes}
        for s in self.nodes:
            if self.fixZ[s] == 1:
                back0[s][[2, 5]] = -np.inf
        for s in self.br:
            p = int(el.parent2[s])
            x = np.vstack([self.tmint[s], self.tmint[s]]) + back0[s][:, None]
            mx = x.max()
            with np.errstate(divide="ignore"):
                msg = np.log(np.exp(x - mx).sum(axis=0)) + mx - np.log(2)
            back0[p][:3] += msg
            back0[p][3:] += msg
        logp_zc = float(back0[el.root][0])
        loglik, L, M = self.upward(Z, B, logs)
        logp_z = self.prior_z(Z) - logp_zc
        tr = dict(loglik=np.zeros(T), full=np.zeros(T), c=np.zeros(T), n=np.zeros(T), gb=np.zeros(T), Z=np.zeros((T, N), np.int8), B=np.zeros((T, N), np.int8),
                  accept_rates=0, accept_zb=0, logp_zc=logp_zc)
        move = True
        post = post_next = None
        full = 0.0
        for m in range(T):
            if move:
                full = self.full_loglik(loglik, logp_z, Z, B, sel_c, sel_n, gbc)
            tr["loglik"][m], tr["full"][m], tr["c"][m], tr["n"][m], tr["gb"][m] = loglik, full, sel_c, sel_n, gbc
            tr["Z"][m], tr["B"][m] = Z, B
            if m == T - 1:
                break
            if move:
                post_next = self.downward(Z, B, logs, L, M)
            if m == 0:
                post = post_next
            emis = self.emissions(post, logs)
            back = self.messages(emis)
            loglik_old = float(back[el.root][0])
            elbo_old = self.elbo(Z, B, post, logs)
            cur = (sel_c, sel_n, gbc)
            wlogs = dict(logs)
            moved = False
            for mm in range(3):
                w = philox(self.seed, self.key, mm, 0, m, ST_GBGC)
                new_gb = cur[2] + (u01(w[0]) - 0.5) * par.prop_gb * 2
                if new_gb < 0.01:
                    new_gb = cur[2]
                new_n = cur[1]
                if self.resZ != 0:
                    u = u01(w[1]) * (par.prop_n - 1 / par.prop_n) + 1 / par.prop_n
                    new_n = (cur[1] + par.nprior_c) * u - par.nprior_c
                u = u01(w[2]) * (par.prop_c - 1 / par.prop_c) + 1 / par.prop_c
                new_c = cur[0] * u
                plogs = self.class_logs(new_c, new_n, new_gb, range(1, 6))
                emis_p = self.emissions(post, plogs, base=emis)
                back_p = self.messages(emis_p)
                ll_p = float(back_p[el.root][0])
                ratio = ll_p - loglik_old + log_gamma_pdf(new_gb, par.gbprior_a, par.gbprior_b) - log_gamma_pdf(cur[2], par.gbprior_a, par.gbprior_b)
                if self.resZ != 0:
                    ratio += log_gamma_pdf(new_n + par.nprior_c, par.nprior_a, par.nprior_b) - log_gamma_pdf(cur[1] + par.nprior_c, par.nprior_a, par.nprior_b)
                with np.errstate(invalid="ignore", divide="ignore"):
                    ratio += log_gamma_pdf(-new_c, par.cprior_a, par.cprior_b) - log_gamma_pdf(-cur[0], par.cprior_a, par.cprior_b) - np.log(-new_c) + np.log(-cur[0])
                if np.log(u01(w[3])) < ratio:
                    cur = (new_c, new_n, new_gb)
                    loglik_old = ll_p
                    emis = emis_p
                    wlogs.update(plogs)
                    moved = True
            tr["accept_rates"] += moved
            back = self.messages(emis)
            Zn, Bn = self.draw_zb(Z, B, back, m)
            ll_new, Ln, Mn = self.upward(Zn, Bn, wlogs)
            elbo_new = self.elbo(Zn, Bn, post, wlogs)
            ratio = ll_new - loglik + elbo_old - elbo_new
            if np.log(u01(philox(self.seed, self.key, 0, 2, m, ST_GBGC)[0])) < ratio:
                Z, B, logs, L, M, loglik = Zn, Bn, wlogs, Ln, Mn, ll_new
                sel_c, sel_n, gbc = cur
                logp_z = self.prior_z(Z) - logp_zc
                move = True
                tr["accept_zb"] += 1
            else:
                move = False
            post = post_next
        return tr


def _categorical(w, u):
    """first index k with u * total < cumsum_k (weights summed in index order), else the last positive weight"""
    tot = 0.0
    for v in w:
        tot += float(v)
    x, acc = u * tot, 0.0
    for k, v in enumerate(w):
        acc += float(v)
        if x < acc:
            return k
    pos = [k for k, v in enumerate(w) if v > 0]
    return pos[-1] if pos else -1
