"""TEST INFRASTRUCTURE -- oracle restatement of BPP::sample_hyperparam (src/PhyloAcc-ST/bpp.cpp:324-477).

Pinned against the reference itself: tests/golden/ratite43/hyper.json holds the MH ratios, proposals and accepted values
printed by the reference's own sample_hyperparam (run through oracle/_ref/ref_probe, generator tests/golden/make_golden.py);
`mh_rate` / `mh_beta` reproduce those ratios from the same inputs, `sweep` replays a whole sweep given the reference's
proposals and acceptance uniforms, and `sweep_scheduled` draws proposals / uniforms from the Philox schedule the product
uses (stream 4, key 3) so that the CUDA engine's host code can be compared value for value.
"""
from __future__ import annotations

import math

import numpy as np

from . import coracle

ST_HYPER, KEY_HYPER = 4, 3
FIELDS = ("nprior_a", "nprior_b", "cprior_a", "cprior_b", "prior_l_a", "prior_l_b", "prior_g_a", "prior_g_b", "prior_l2_a", "prior_l2_b")


def _log_gamma_pdf(x, a, b):
    return math.log(coracle.lib().pa_or_gamma_pdf(float(x), float(a), float(b)))


def mh_rate(rates, a, b, a_prop, b_prop, va=100.0, vb=100.0):
    """(M_ratio, H_ratio) of the Gamma-prior step for one rate class (bpp.cpp:341-361 / 366-384)."""
    p, r, q, s = 1.0, 1.0, 0.1, 0.1  # bpp.cpp:328-329
    n = float(len(rates))
    sum_r, log_prod_r = 0.0, 0.0
    for x in rates:  # same accumulation order as the reference
        sum_r += float(x)
        log_prod_r += math.log(float(x))
    M = (a_prop - 1) * (math.log(p) + log_prod_r) - (q + sum_r) / b_prop - (n + r) * math.lgamma(a_prop) - math.log(b_prop) * a_prop * (s + n)
    M -= (a - 1) * (math.log(p) + log_prod_r) - (q + sum_r) / b - (n + r) * math.lgamma(a) - math.log(b) * a * (s + n)
    H = (_log_gamma_pdf(a, va, a_prop / va) - _log_gamma_pdf(a_prop, va, a / va) + _log_gamma_pdf(b, vb, b_prop / vb) - _log_gamma_pdf(b_prop, vb, b / vb))
    return M, H


def mh_beta(rates, a, b, a_prop, b_prop):
    """M_ratio + H_ratio of the Beta-prior step for l, l2 or g (bpp.cpp:389-414)."""
    n = float(len(rates))
    log_p, log_pc = 0.0, 0.0
    with np.errstate(divide="ignore"):
        for x in rates:
            log_p += float(np.log(np.float64(x)))
            log_pc += float(np.log(np.float64(1) - np.float64(x)))
    L = coracle.lib()
    M = (a_prop - a) * (log_p - 1) + (b_prop - b) * (log_pc - 1)
    M += n * (L.pa_or_lnbeta(float(a), float(b)) - L.pa_or_lnbeta(float(a_prop), float(b_prop)))
    with np.errstate(divide="ignore", invalid="ignore"):
        H = float(np.log(np.float64(a)) - np.log(np.float64(a_prop)) + np.log(np.float64(b)) - np.log(np.float64(b_prop)))
    return M + H


def sweep(hy: dict, rates: dict, props: dict, unis: dict) -> dict:
    """One sweep with explicit proposals (`props[k] = (a_prop, b_prop)`) and acceptance uniforms (`unis[k]`), k in n, c, l, l2, g."""
    hy = dict(hy)
    for k, (fa, fb) in (("n", ("nprior_a", "nprior_b")), ("c", ("cprior_a", "cprior_b"))):
        M, H = mh_rate(rates[k], hy[fa], hy[fb], *props[k])
        if math.log(unis[k]) < M + H:
            hy[fa], hy[fb] = props[k]
    for k, (fa, fb) in (("l", ("prior_l_a", "prior_l_b")), ("l2", ("prior_l2_a", "prior_l2_b")), ("g", ("prior_g_a", "prior_g_b"))):
        R = mh_beta(rates[k], hy[fa], hy[fb], *props[k])
        if math.log(unis[k]) < R:  # NaN (Dollo: prior_l2_a == 0) never accepts
            hy[fa], hy[fb] = props[k]
    return hy


def sweep_scheduled(hy: dict, rates: dict, seed: int, it: int) -> dict:
    """The sweep with proposals / uniforms from the product's Philox schedule, in the reference's order of consumption."""
    L = coracle.lib()
    n_draw = [0]

    def gam(a, b):
        v = L.pa_or_gamma(seed & 0xFFFFFFFF, KEY_HYPER, n_draw[0], 0, it, ST_HYPER, float(a), float(b))
        n_draw[0] += 1
        return v

    def uni():
        import ctypes as C
        out = (C.c_uint32 * 4)()
        L.pa_or_philox(seed & 0xFFFFFFFF, KEY_HYPER, n_draw[0], 0, it, ST_HYPER, out)
        n_draw[0] += 1
        return (float(out[0]) + 0.5) * 2.3283064365386963e-10

    hy = dict(hy)
    va = 100.0
    na_p = gam(va, hy["nprior_a"] / va)
    ca_p = gam(va, hy["cprior_a"] / va)
    nb_p = gam(va, hy["nprior_b"] / va)
    cb_p = gam(va, hy["cprior_b"] / va)
    for k, fa, fb, pr in (("n", "nprior_a", "nprior_b", (na_p, nb_p)), ("c", "cprior_a", "cprior_b", (ca_p, cb_p))):
        M, H = mh_rate(rates[k], hy[fa], hy[fb], *pr)
        if math.log(uni()) < M + H:
            hy[fa], hy[fb] = pr
    for k, fa, fb in (("l", "prior_l_a", "prior_l_b"), ("l2", "prior_l2_a", "prior_l2_b"), ("g", "prior_g_a", "prior_g_b")):
        a_p = hy[fa] * (uni() * (1.3 - 0.7) + 0.7)
        b_p = hy[fb] * (uni() * (1.3 - 0.7) + 0.7)
        R = mh_beta(rates[k], hy[fa], hy[fb], a_p, b_p)
        if math.log(uni()) < R:
            hy[fa], hy[fb] = a_p, b_p
    return hy
