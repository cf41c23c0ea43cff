"""oracle/coracle.py -- TEST INFRASTRUCTURE.  ctypes binding of oracle/_build/liboracle.so (oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from . import frontend as fe

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)


class OrElement(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("S", C.c_int32), ("GG", C.c_int32),
        ("tg", _i8p), ("masks", _u8p), ("missing", _u8p),
        ("root", C.c_int32), ("n_nodes", C.c_int32), ("nodes", _i32p),
        ("children2", _i32p), ("parent2", _i32p), ("distances2", _f64p), ("distances", _f64p),
        ("n_upper_c", C.c_int32), ("upper_c", _i32p), ("n_upper_conserve_c", C.c_int32), ("upper_conserve_c", _i32p),
        ("eigenval", C.c_double * 4), ("eigenvec", C.c_double * 16), ("eigeninv", C.c_double * 16), ("log_pi", C.c_double * 4),
        ("cprior_a", C.c_double), ("cprior_b", C.c_double), ("nprior_a", C.c_double), ("nprior_b", C.c_double),
        ("prior_l_a", C.c_double), ("prior_l_b", C.c_double), ("prior_l2_a", C.c_double), ("prior_l2_b", C.c_double),
        ("prior_g_a", C.c_double), ("prior_g_b", C.c_double),
        ("init_crate", C.c_double), ("init_nrate", C.c_double), ("init_lrate", C.c_double), ("init_lrate2", C.c_double),
        ("init_grate", C.c_double),
        ("cub", C.c_double), ("nlb", C.c_double), ("consToMis", C.c_double),
        ("ropt", C.c_int32), ("num_burn", C.c_int32), ("num_mcmc", C.c_int32), ("adaptive_freq", C.c_int32),
    ]


class OrTrace(C.Structure):
    _fields_ = [
        ("loglik", _f64p), ("full_loglik", _f64p), ("n_rate", _f64p), ("c_rate", _f64p), ("g_rate", _f64p),
        ("l_rate", _f64p), ("l2_rate", _f64p), ("Z", _i8p), ("max_Z", _i8p),
        ("max_loglik", C.c_double), ("max_m", C.c_int32), ("accept_n", C.c_int32), ("accept_c", C.c_int32),
        ("failure", C.c_int32),
    ]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "oracle.c")):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.pa_or_loglik.restype = C.c_double
        _lib.pa_or_loglik.argtypes = [C.POINTER(OrElement), C.c_int, _i8p, C.c_double, C.c_double, _f64p]
        _lib.pa_or_eval2.restype = C.c_double
        _lib.pa_or_eval2.argtypes = [C.POINTER(OrElement), C.c_int, C.POINTER(OrTrace), _i32p]
        _lib.pa_or_chain.restype = C.c_int
        _lib.pa_or_chain.argtypes = [C.POINTER(OrElement), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(OrTrace)]
        _lib.pa_or_emission_back.restype = None
        _lib.pa_or_emission_back.argtypes = [C.POINTER(OrElement), C.c_int, _i8p] + [C.c_double] * 5 + [_i8p, _f64p, _f64p]
        _lib.pa_or_log_tm.restype = None
        _lib.pa_or_log_tm.argtypes = [C.POINTER(OrElement), C.c_double, C.c_double, _f64p]
        _lib.pa_or_philox.restype = None
        _lib.pa_or_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        _lib.pa_or_gamma.restype = C.c_double
        _lib.pa_or_gamma.argtypes = [C.c_uint32] * 6 + [C.c_double, C.c_double]
        _lib.pa_or_gamma_pdf.restype = C.c_double
        _lib.pa_or_gamma_pdf.argtypes = [C.c_double] * 3
        _lib.pa_or_lnbeta.restype = C.c_double
        _lib.pa_or_lnbeta.argtypes = [C.c_double] * 2
    return _lib


def eigensystem(Q: np.ndarray):
    """bpp.hpp:357-362: eig_gen(Q) -> eigenval, eigenvec = V^T, eigeninv = inv(V^T) (real parts)."""
    lam, V = np.linalg.eig(Q)
    lam = lam.real
    VT = np.ascontiguousarray(V.real.T)
    return lam, VT, np.linalg.inv(VT)


class ElementHandle:
    """Keeps the numpy buffers alive behind an OrElement."""

    def __init__(self, model: fe.Model, el: fe.Element, site_order: Optional[np.ndarray] = None):
        tree, P = model.tree, model.params
        assert not el.filter
        self.N, self.S = tree.N, tree.S
        tg, masks = el.tg, el.masks
        if site_order is not None:  # canonical (pattern-sorted) site order
            tg, masks = tg[site_order], masks[site_order]
        self.GG = tg.shape[0]
        self._keep = dict(
            tg=np.ascontiguousarray(tg, dtype=np.int8), masks=np.ascontiguousarray(masks, dtype=np.uint8),
            missing=np.ascontiguousarray(el.missing, dtype=np.uint8), nodes=np.asarray(el.nodes, dtype=np.int32),
            children2=np.ascontiguousarray(el.children2, dtype=np.int32), parent2=np.ascontiguousarray(el.parent2, dtype=np.int32),
            distances2=np.ascontiguousarray(el.distances2, dtype=np.float64), distances=np.ascontiguousarray(tree.dist, dtype=np.float64),
            upper_c=np.asarray(el.upper_c, dtype=np.int32), upper_conserve_c=np.asarray(el.upper_conserve_c, dtype=np.int32))
        k = self._keep
        s = OrElement()
        s.N, s.S, s.GG = tree.N, tree.S, self.GG
        s.tg = k["tg"].ctypes.data_as(_i8p)
        s.masks = k["masks"].ctypes.data_as(_u8p)
        s.missing = k["missing"].ctypes.data_as(_u8p)
        s.root, s.n_nodes = el.root, len(el.nodes)
        s.nodes = k["nodes"].ctypes.data_as(_i32p)
        s.children2 = k["children2"].ctypes.data_as(_i32p)
        s.parent2 = k["parent2"].ctypes.data_as(_i32p)
        s.distances2 = k["distances2"].ctypes.data_as(_f64p)
        s.distances = k["distances"].ctypes.data_as(_f64p)
        s.n_upper_c = len(el.upper_c)
        s.upper_c = k["upper_c"].ctypes.data_as(_i32p)
        s.n_upper_conserve_c = len(el.upper_conserve_c)
        s.upper_conserve_c = k["upper_conserve_c"].ctypes.data_as(_i32p)
        lam, VT, inv = eigensystem(tree.Q)
        s.eigenval[:] = list(lam)
        s.eigenvec[:] = list(VT.reshape(-1))
        s.eigeninv[:] = list(inv.reshape(-1))
        s.log_pi[:] = list(np.log(tree.pi))
        s.cprior_a, s.cprior_b, s.nprior_a, s.nprior_b = P.CONSERVE_PRIOR_A, P.CONSERVE_PRIOR_B, P.ACCE_PRIOR_A, P.ACCE_PRIOR_B
        s.prior_l_a, s.prior_l_b = P.HYPER_LRATE_A, P.HYPER_LRATE_B
        s.prior_l2_a, s.prior_l2_b = P.HYPER_LRATE2_A, P.HYPER_LRATE2_B
        s.prior_g_a, s.prior_g_b = P.HYPER_GRATE_A, P.HYPER_GRATE_B
        s.init_crate, s.init_nrate = P.INIT_CONSERVE_RATE, P.INIT_ACCE_RATE
        s.init_lrate, s.init_lrate2, s.init_grate = P.INIT_LRATE, P.INIT_LRATE2, P.INIT_GRATE
        s.cub, s.nlb, s.consToMis = P.CUB, P.NLB, P.CONSTOMIS
        s.ropt, s.num_burn, s.num_mcmc, s.adaptive_freq = P.ROPT, P.BURNIN, P.MCMC, P.ADAPT_FREQ
        self.struct = s

    # ---- deterministic pieces
    def loglik(self, resZ: int, Z: np.ndarray, n_rate: float, c_rate: float):
        Z = np.ascontiguousarray(Z, dtype=np.int8)
        per = np.zeros(self.GG)
        v = lib().pa_or_loglik(C.byref(self.struct), resZ, Z.ctypes.data_as(_i8p), n_rate, c_rate, per.ctypes.data_as(_f64p))
        return v, per

    def emission_back(self, resZ, Z, n_rate, c_rate, g, l, l2, tg_hist):
        Z = np.ascontiguousarray(Z, dtype=np.int8)
        tg_hist = np.ascontiguousarray(tg_hist, dtype=np.int8)
        em = np.zeros((self.N, 3))
        bk = np.zeros((self.N, 3))
        lib().pa_or_emission_back(C.byref(self.struct), resZ, Z.ctypes.data_as(_i8p), n_rate, c_rate, g, l, l2,
                                  tg_hist.ctypes.data_as(_i8p), em.ctypes.data_as(_f64p), bk.ctypes.data_as(_f64p))
        return em, bk

    def log_tm(self, t: float, rate: float) -> np.ndarray:
        out = np.zeros(16)
        lib().pa_or_log_tm(C.byref(self.struct), t, rate, out.ctypes.data_as(_f64p))
        return out.reshape(4, 4)

    # ---- chain / eval2
    def new_trace(self):
        T = self.struct.num_burn + self.struct.num_mcmc
        bufs = dict(loglik=np.zeros(T), full_loglik=np.zeros(T), n_rate=np.zeros(T), c_rate=np.zeros(T), g_rate=np.zeros(T),
                    l_rate=np.zeros(T), l2_rate=np.zeros(T), Z=np.ones((T, self.N), dtype=np.int8),
                    max_Z=np.zeros(self.N, dtype=np.int8))
        tr = OrTrace()
        for k in ("loglik", "full_loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
            setattr(tr, k, bufs[k].ctypes.data_as(_f64p))
        tr.Z = bufs["Z"].ctypes.data_as(_i8p)
        tr.max_Z = bufs["max_Z"].ctypes.data_as(_i8p)
        return tr, bufs

    def chain(self, resZ: int, seed: int, chain_key: int, accept_n: int = 0, accept_c: int = 0):
        tr, bufs = self.new_trace()
        tr.accept_n, tr.accept_c = accept_n, accept_c
        lib().pa_or_chain(C.byref(self.struct), resZ, seed, chain_key, C.byref(tr))
        bufs.update(max_loglik=tr.max_loglik, max_m=tr.max_m, accept_n=tr.accept_n, accept_c=tr.accept_c, failure=tr.failure,
                    _trace=tr)
        return bufs

    def eval2(self, resZ: int, bufs) -> float:
        tr = bufs.get("_trace")
        if tr is None:
            tr, nb = self.new_trace()
            for k in ("loglik", "full_loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate", "Z", "max_Z"):
                nb[k][...] = bufs[k]
            bufs = dict(bufs)
            bufs["_keep"] = nb
        fail = C.c_int32(0)
        return lib().pa_or_eval2(C.byref(self.struct), resZ, C.byref(tr), C.byref(fail))


def full_loglik_terms(P: fe.Params, resZ: int, c_rate, n_rate):
    """MonitorChain's prior terms (bpp_c.cpp:784,801)."""
    L = lib()
    v = np.array([np.log(L.pa_or_gamma_pdf(float(c), P.CONSERVE_PRIOR_A, P.CONSERVE_PRIOR_B)) for c in np.atleast_1d(c_rate)])
    if resZ != 0:
        v = v + np.array([np.log(L.pa_or_gamma_pdf(float(n), P.ACCE_PRIOR_A, P.ACCE_PRIOR_B)) for n in np.atleast_1d(n_rate)])
    return v
