"""phyloacc_b200 -- B200-native engine for PhyloAcc's per-element MCMC hot path (PhyloAcc-ST).

This package is a thin ctypes binding of the C ABI in ``include/phyloacc_b200.h`` (libphyloacc_b200.so, built
in-tree by ``phyloacc_b200.build``).  All arithmetic of the hot path runs in hand-written sm_100a CUDA kernels;
there is no CPU fallback and importing this module never touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PA_LIB", _build.LIB)  # PA_LIB: load an alternative build of the same library (A/B timing)
BIN_PATH = _build.BIN
BIN_GBGC_PATH = _build.BIN_GBGC

PA_OK = 0
M0, M1, M2 = 0, 1, 2
MODEL_RESZ = {M0: 0, M1: 2, M2: 1}  # model slot -> reference's resZ (main.cpp:466-494)

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)


class PaTree(C.Structure):
    _fields_ = [("S", C.c_int32), ("N", C.c_int32), ("parent", _i32p), ("children", _i32p), ("distances", _f64p)]


class PaSubst(C.Structure):
    _fields_ = [("pi", C.c_double * 4), ("Q", C.c_double * 16)]


class PaParams(C.Structure):
    _fields_ = [
        ("seed", C.c_int32), ("init_conserve_rate", C.c_double), ("init_acce_rate", C.c_double),
        ("conserve_prior_a", C.c_double), ("conserve_prior_b", C.c_double), ("acce_prior_a", C.c_double), ("acce_prior_b", C.c_double),
        ("ropt", C.c_int32), ("cub", C.c_double), ("nlb", C.c_double),
        ("burnin", C.c_int32), ("mcmc", C.c_int32), ("adapt_freq", C.c_int32),
        ("init_lrate", C.c_double), ("init_lrate2", C.c_double), ("init_grate", C.c_double),
        ("hyper_lrate_a", C.c_double), ("hyper_lrate_b", C.c_double), ("hyper_grate_a", C.c_double), ("hyper_grate_b", C.c_double),
        ("hyper_lrate2_a", C.c_double), ("hyper_lrate2_b", C.c_double),
        ("conserve_prop", C.c_double), ("gap_prop", C.c_double), ("constomis", C.c_double),
        ("gapchar", C.c_char), ("prune_tree", C.c_int32), ("trim_gap_percent", C.c_double), ("min_len", C.c_int32),
        ("n_outgroup", C.c_int32), ("outgroup", _i32p), ("n_conserve", C.c_int32), ("conserve", _i32p),
        ("num_thread", C.c_int32),
    ]


class PaResults(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("filtered", _u8p), ("loglik", _f64p), ("loglik_max", _f64p),
                ("max_Z", _i8p), ("median_rates", _f64p), ("post_Z", _f64p)]


class PaTrace(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("loglik", _f64p), ("n_rate", _f64p), ("c_rate", _f64p),
                ("g_rate", _f64p), ("l_rate", _f64p), ("l2_rate", _f64p), ("Z", _i8p)]


class PaElementInfo(C.Structure):
    _fields_ = [("cap_P", C.c_int32), ("cap_GG", C.c_int32),
                ("filtered", C.c_int32), ("GG", C.c_int32), ("P", C.c_int32), ("root", C.c_int32), ("n_nodes", C.c_int32),
                ("missing", _u8p), ("nodes", _i32p), ("children2", _i32p), ("parent2", _i32p), ("distances2", _f64p),
                ("pattern_codes", _u8p), ("pattern_weights", _i32p), ("site_pattern", _i32p), ("fixZ", _u8p), ("n_slots", C.c_int32)]


HYPER_NSUMS = 11  # PA_HYPER_NSUMS


class PaHyper(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("nprior_a", "nprior_b", "cprior_a", "cprior_b", "prior_l_a", "prior_l_b", "prior_g_a", "prior_g_b",
                                          "prior_l2_a", "prior_l2_b")]

    def as_dict(self):
        return {k: float(getattr(self, k)) for k, _ in self._fields_}


class PaGbgcConsts(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("grate", "lrate", "prior_gb", "cons_to_mis", "ncons_to_mis")]


class PaGbgcParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("sel_c0", "sel_n0", "gbc0", "cprior_a", "cprior_b", "nprior_a", "nprior_b", "nprior_c", "gbprior_a",
                                          "gbprior_b", "gbprior_c", "cons_to_mis", "ncons_to_mis", "grate", "lrate", "prop_c", "prop_n",
                                          "prop_gb")] + [("num_burn", C.c_int32), ("num_mcmc", C.c_int32), ("seed", C.c_uint32), ("pad_", C.c_int32)]


class PaGbgcResults(C.Structure):
    _fields_ = [("n_ids", C.c_int32), ("N", C.c_int32), ("T", C.c_int32), ("filtered", _u8p), ("loglik", _f64p), ("loglik_max", _f64p),
                ("med", _f64p), ("post_b", _f64p), ("post_z", _f64p), ("max_z", _i8p), ("max_b", _i8p), ("accept", _i32p),
                ("trace", _f64p), ("trace_z", _i8p), ("trace_b", _i8p), ("kernel_ms", C.c_double),
                ("node_pattern_updates", C.c_uint64), ("chain_iterations", C.c_uint64), ("pt_builds", C.c_uint64),
                ("branch_pattern_updates", C.c_uint64), ("d2h_bytes", C.c_uint64), ("phase_cycles", C.c_uint64 * 12), ("kernel_launches", C.c_uint64)]


class PaCounters(C.Structure):
    _fields_ = [("node_pattern_updates", C.c_uint64), ("root_evals", C.c_uint64), ("pt_builds", C.c_uint64),
                ("history_draws", C.c_uint64), ("chain_iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("node_site_updates", C.c_uint64), ("d2h_bytes", C.c_uint64), ("host_eval_chains", C.c_uint64)]


class PaGtStates(C.Structure):  # pa_gt_states
    _fields_ = [("n_states", C.c_int32), ("N", C.c_int32), ("elem", C.POINTER(C.c_int32)), ("gchild", C.POINTER(C.c_int32)),
                ("groot", C.POINTER(C.c_int32)), ("gmissing", C.POINTER(C.c_uint8)), ("gheights", C.POINTER(C.c_double)),
                ("Z", C.POINTER(C.c_int8)), ("n_rate", C.POINTER(C.c_double)), ("c_rate", C.POINTER(C.c_double)),
                ("lin_off", C.POINTER(C.c_int64)), ("lin_species", C.POINTER(C.c_int32)), ("lin_node", C.POINTER(C.c_int32)),
                ("lin_next", C.POINTER(C.c_int32)), ("heights", C.POINTER(C.c_double)), ("thetas", C.POINTER(C.c_double)),
                ("cons_to_mis", C.c_double), ("ncons_to_mis", C.c_double)]


EXPORTS = [
    "pa_last_error", "pa_version", "pa_params_default", "pa_engine_create", "pa_engine_destroy", "pa_engine_set_stream",
    "pa_engine_load_alignment", "pa_engine_num_elements", "pa_engine_element_info", "pa_engine_loglik_replay",
    "pa_engine_sample", "pa_engine_collect", "pa_engine_run", "pa_engine_get_trace", "pa_engine_counters", "pa_engine_phase_cycles",
    "pa_engine_create_from_config", "pa_engine_get_tree", "pa_engine_node_name", "pa_engine_element_name",
    "pa_engine_get_ids", "pa_engine_get_params", "pa_run_from_config", "pa_gbgc_run_from_config", "pa_device_fp64_peak",
    "pa_engine_sample_ex", "pa_engine_get_hyper", "pa_engine_set_hyper", "pa_engine_sample_hyper", "pa_engine_hyper_sums",
    "pa_engine_sample_hyper_sums", "pa_engine_element_blob", "pa_gbgc_transition", "pa_engine_gbgc_replay", "pa_engine_gbgc_integrate",
    "pa_gbgc_params_default", "pa_engine_gbgc_run", "pa_engine_gt_replay", "pa_gbgc_eval2", "pa_gbgc_output_init",
]

_lib = None


class PhyloAccError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"phyloacc_b200 error {code}: {msg}")
        self.code = code


def load_library(build_if_missing: bool = True):
    """dlopen the native library; fails loudly (no fallback) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ImportError(f"{LIB_PATH} is missing: run `python -m phyloacc_b200.build` (there is no CPU fallback)")
        _build.build_native()
    lib = C.CDLL(LIB_PATH)
    lib.pa_last_error.restype = C.c_char_p
    lib.pa_version.restype = C.c_char_p
    lib.pa_params_default.argtypes = [C.POINTER(PaParams)]
    lib.pa_params_default.restype = None
    lib.pa_engine_create.argtypes = [C.POINTER(PaTree), C.POINTER(PaSubst), C.POINTER(PaParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.pa_engine_destroy.argtypes = [C.c_void_p]
    lib.pa_engine_destroy.restype = None
    lib.pa_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pa_engine_load_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, _i64p, _i32p]
    lib.pa_engine_num_elements.argtypes = [C.c_void_p]
    lib.pa_engine_element_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(PaElementInfo)]
    lib.pa_engine_loglik_replay.argtypes = [C.c_void_p, C.c_int32, _i32p, _i8p, _f64p, _f64p, _f64p]
    lib.pa_engine_sample.argtypes = [C.c_void_p, _i32p, C.c_int32]
    lib.pa_engine_collect.argtypes = [C.c_void_p, C.POINTER(PaResults)]
    lib.pa_engine_run.argtypes = [C.c_void_p, _i32p, C.c_int32, C.POINTER(PaResults)]
    lib.pa_engine_sample_ex.argtypes = [C.c_void_p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.pa_engine_get_hyper.argtypes = [C.c_void_p, C.POINTER(PaHyper)]
    lib.pa_engine_set_hyper.argtypes = [C.c_void_p, C.POINTER(PaHyper)]
    lib.pa_engine_sample_hyper.argtypes = [C.c_void_p, C.c_int32, _i32p, C.c_int32, C.POINTER(PaHyper)]
    lib.pa_engine_hyper_sums.argtypes = [C.c_void_p, _i32p, C.c_int32, _f64p]
    lib.pa_engine_element_blob.argtypes = [C.c_void_p, C.c_int32, _u8p, C.c_int64, C.POINTER(C.c_int64)]
    lib.pa_engine_sample_hyper_sums.argtypes = [C.c_void_p, C.c_int32, _f64p, C.POINTER(PaHyper)]
    lib.pa_engine_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(PaTrace)]
    lib.pa_engine_counters.argtypes = [C.c_void_p, C.POINTER(PaCounters)]
    lib.pa_engine_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.pa_engine_create_from_config.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.pa_engine_get_tree.argtypes = [C.c_void_p, C.POINTER(PaTree), C.POINTER(PaSubst)]
    lib.pa_engine_node_name.argtypes = [C.c_void_p, C.c_int32]
    lib.pa_engine_node_name.restype = C.c_char_p
    lib.pa_engine_element_name.argtypes = [C.c_void_p, C.c_int32]
    lib.pa_engine_element_name.restype = C.c_char_p
    lib.pa_engine_get_ids.argtypes = [C.c_void_p, C.POINTER(_i32p), _i32p]
    lib.pa_engine_get_params.argtypes = [C.c_void_p, C.POINTER(PaParams)]
    lib.pa_run_from_config.argtypes = [C.c_char_p]
    lib.pa_gbgc_run_from_config.argtypes = [C.c_char_p]
    lib.pa_device_fp64_peak.argtypes = [C.c_int, C.c_int, _f64p]
    lib.pa_engine_gbgc_integrate.argtypes = [C.c_void_p, C.c_int32, _i32p, _i8p, _i8p, _i8p, _f64p, _f64p, _f64p, C.POINTER(PaGbgcConsts), _f64p, _f64p, _f64p]
    lib.pa_engine_gbgc_replay.argtypes = [C.c_void_p, C.c_int32, _i32p, _i8p, _i8p, _f64p, _f64p, _f64p, _f64p]
    lib.pa_gbgc_transition.argtypes = [C.c_int, C.POINTER(PaSubst), C.c_int32, _f64p, _f64p, _f64p, _f64p, _f64p]
    lib.pa_gbgc_params_default.argtypes = [C.POINTER(PaGbgcParams)]
    lib.pa_gbgc_params_default.restype = None
    lib.pa_engine_gbgc_run.argtypes = [C.c_void_p, _i32p, C.c_int32, _i32p, C.c_int32, C.POINTER(PaGbgcParams), C.c_int, C.POINTER(PaGbgcResults)]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != PA_OK:
        raise PhyloAccError(rc, load_library().pa_last_error().decode(errors="replace"))


def gbgc_eval2(full, c_rate, n_rate, gbc, Z, B, nodes, missing, burn: int, resz: int):
    """The GBGC variant's Eval2 on a given chain (pa_gbgc_eval2, host-only): returns (estimate of log P(X | model), MaxLoglik)."""
    lib = load_library()
    full, c_rate, n_rate, gbc = (np.ascontiguousarray(x, dtype=np.float64) for x in (full, c_rate, n_rate, gbc))
    T = len(full)
    Z = np.ascontiguousarray(Z, dtype=np.int8).reshape(T, -1)
    N = Z.shape[1]
    B = np.ascontiguousarray(B, dtype=np.int8).reshape(T, N)
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    missing = np.ascontiguousarray(missing, dtype=np.uint8)
    out, mx = C.c_double(), C.c_double()
    _check(lib.pa_gbgc_eval2(N, burn, T - burn, resz, full.ctypes.data_as(_f64p), c_rate.ctypes.data_as(_f64p), n_rate.ctypes.data_as(_f64p),
                             gbc.ctypes.data_as(_f64p), Z.ctypes.data_as(_i8p), B.ctypes.data_as(_i8p), nodes.ctypes.data_as(_i32p), len(nodes),
                             missing.ctypes.data_as(_u8p), C.byref(out), C.byref(mx)))
    return out.value, mx.value


def gbgc_output_init(n_rate, c_rate, gbc, Z, B, missing, burn: int):
    """The GBGC variant's Output_init summaries on a given chain (pa_gbgc_output_init, host-only): (med [3], P(B) [N], P(Z) [N, 4])."""
    lib = load_library()
    n_rate, c_rate, gbc = (np.ascontiguousarray(x, dtype=np.float64) for x in (n_rate, c_rate, gbc))
    T = len(n_rate)
    Z = np.ascontiguousarray(Z, dtype=np.int8).reshape(T, -1)
    N = Z.shape[1]
    B = np.ascontiguousarray(B, dtype=np.int8).reshape(T, N)
    missing = np.ascontiguousarray(missing, dtype=np.uint8)
    med, pb, pz = np.zeros(3), np.zeros(N), np.zeros((N, 4))
    _check(lib.pa_gbgc_output_init(N, burn, T - burn, n_rate.ctypes.data_as(_f64p), c_rate.ctypes.data_as(_f64p), gbc.ctypes.data_as(_f64p),
                                   Z.ctypes.data_as(_i8p), B.ctypes.data_as(_i8p), missing.ctypes.data_as(_u8p), med.ctypes.data_as(_f64p),
                                   pb.ctypes.data_as(_f64p), pz.ctypes.data_as(_f64p)))
    return med, pb, pz


def gbgc_transition(Q, sel, bias, t, device: int = 0):
    """PhyloAcc-ST-GBGC's per-branch matrices on the device (pa_gbgc.cu): returns (rate [n, 4, 4], P [n, 4, 4]) for n
    (selection coefficient, gBGC bias, branch length) triples; rate i -> j, P(i -> j | t)."""
    lib = load_library()
    s = PaSubst()
    s.Q[:] = list(map(float, np.asarray(Q).reshape(-1)))
    sel, bias, t = (np.ascontiguousarray(x, dtype=np.float64) for x in (sel, bias, t))
    n = len(sel)
    rate, P = np.zeros((n, 4, 4)), np.zeros((n, 4, 4))
    _check(lib.pa_gbgc_transition(device, C.byref(s), n, sel.ctypes.data_as(_f64p), bias.ctypes.data_as(_f64p), t.ctypes.data_as(_f64p),
                                  rate.ctypes.data_as(_f64p), P.ctypes.data_as(_f64p)))
    return rate, P


def default_params() -> PaParams:
    p = PaParams()
    load_library().pa_params_default(C.byref(p))
    return p


class Results:
    """numpy view of pa_results for n elements on an N-node tree."""

    def __init__(self, n: int, N: int):
        self.n, self.N = n, N
        self.filtered = np.zeros(n, dtype=np.uint8)
        self.loglik = np.zeros((n, 3))
        self.loglik_max = np.zeros((n, 3))
        self.max_Z = np.zeros((n, 3, N), dtype=np.int8)
        self.median_rates = np.zeros((n, 3, 5))
        self.post_Z = np.zeros((n, 3, N, 4))
        s = PaResults()
        s.n, s.N = n, N
        s.filtered = self.filtered.ctypes.data_as(_u8p)
        s.loglik = self.loglik.ctypes.data_as(_f64p)
        s.loglik_max = self.loglik_max.ctypes.data_as(_f64p)
        s.max_Z = self.max_Z.ctypes.data_as(_i8p)
        s.median_rates = self.median_rates.ctypes.data_as(_f64p)
        s.post_Z = self.post_Z.ctypes.data_as(_f64p)
        self.struct = s

    @property
    def logBF1(self):  # bpp.cpp:579
        return self.loglik[:, M1] - self.loglik[:, M0]

    @property
    def logBF2(self):
        return self.loglik[:, M1] - self.loglik[:, M2]


class Engine:
    """One pa_engine handle.  Mirrors the reference's BPP (batch model) + the element loop of main.cpp:452-500."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        self._lib = load_library()
        t, s = PaTree(), PaSubst()
        _check(self._lib.pa_engine_get_tree(self._h, C.byref(t), C.byref(s)))
        self.S, self.N = t.S, t.N
        self.parent = np.ctypeslib.as_array(t.parent, shape=(self.N,)).copy()
        self.children = np.ctypeslib.as_array(t.children, shape=(self.N, 2)).copy()
        self.distances = np.ctypeslib.as_array(t.distances, shape=(self.N,)).copy()
        self.pi = np.array(list(s.pi))
        self.Q = np.array(list(s.Q)).reshape(4, 4)

    # ---- construction
    @classmethod
    def create(cls, parent, children, distances, pi, Q, params: Optional[PaParams] = None, outgroup: Sequence[int] = (),
               conserve: Sequence[int] = (), device: int = 0) -> "Engine":
        lib = load_library()
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        children = np.ascontiguousarray(children, dtype=np.int32)
        distances = np.ascontiguousarray(distances, dtype=np.float64)
        N = parent.shape[0]
        t = PaTree((N + 1) // 2, N, parent.ctypes.data_as(_i32p), children.ctypes.data_as(_i32p), distances.ctypes.data_as(_f64p))
        s = PaSubst()
        s.pi[:] = list(map(float, pi))
        s.Q[:] = list(map(float, np.asarray(Q).reshape(-1)))
        p = params if params is not None else default_params()
        og = np.ascontiguousarray(outgroup, dtype=np.int32)
        cg = np.ascontiguousarray(conserve, dtype=np.int32)
        p.n_outgroup, p.outgroup = len(og), og.ctypes.data_as(_i32p)
        p.n_conserve, p.conserve = len(cg), cg.ctypes.data_as(_i32p)
        h = C.c_void_p()
        _check(lib.pa_engine_create(C.byref(t), C.byref(s), C.byref(p), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_config(cls, cfg_path: str, device: int = 0) -> "Engine":
        lib = load_library()
        h = C.c_void_p()
        _check(lib.pa_engine_create_from_config(cfg_path.encode(), device, C.byref(h)))
        return cls(h.value)

    def close(self):
        if self._h:
            self._lib.pa_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data
    def set_stream(self, cuda_stream: int):
        _check(self._lib.pa_engine_set_stream(self._h, C.c_void_p(cuda_stream)))

    def load_alignment(self, X: np.ndarray, elem_start, elem_len):
        """X: uint8 [S, G] lower-case characters in tree-leaf order."""
        X = np.ascontiguousarray(X, dtype=np.uint8)
        es = np.ascontiguousarray(elem_start, dtype=np.int64)
        el = np.ascontiguousarray(elem_len, dtype=np.int32)
        assert X.shape[0] == self.S
        _check(self._lib.pa_engine_load_alignment(self._h, X.ctypes.data_as(C.c_void_p), X.shape[1], len(el),
                                                  es.ctypes.data_as(_i64p), el.ctypes.data_as(_i32p)))

    @property
    def num_elements(self) -> int:
        return self._lib.pa_engine_num_elements(self._h)

    def node_names(self) -> List[str]:
        return [self._lib.pa_engine_node_name(self._h, i).decode() for i in range(self.N)]

    def element_name(self, c: int) -> str:
        return self._lib.pa_engine_element_name(self._h, c).decode()

    def ids(self) -> np.ndarray:
        p, n = _i32p(), C.c_int32()
        _check(self._lib.pa_engine_get_ids(self._h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.int32)

    def params(self) -> PaParams:
        p = PaParams()
        _check(self._lib.pa_engine_get_params(self._h, C.byref(p)))
        return p

    def element_info(self, c: int, gg_max: int | None = None):
        """Integer front-end of element ``c``.  The pattern / column buffers are sized from the element itself (a first call with
        null arrays returns GG and P); ``gg_max`` is accepted for compatibility and ignored."""
        N, S = self.N, self.S
        probe = PaElementInfo()
        _check(self._lib.pa_engine_element_info(self._h, c, C.byref(probe)))
        gg_max = max(1, int(probe.GG), int(probe.P))
        bufs = dict(missing=np.zeros(N, np.uint8), nodes=np.zeros(N, np.int32), children2=np.zeros((N, 2), np.int32),
                    parent2=np.zeros(N, np.int32), distances2=np.zeros(N), pattern_codes=np.zeros((gg_max, S), np.uint8),
                    pattern_weights=np.zeros(gg_max, np.int32), site_pattern=np.zeros(gg_max, np.int32), fixZ=np.zeros((3, N), np.uint8))
        info = PaElementInfo()
        info.cap_P = info.cap_GG = gg_max
        info.missing = bufs["missing"].ctypes.data_as(_u8p)
        info.nodes = bufs["nodes"].ctypes.data_as(_i32p)
        info.children2 = bufs["children2"].ctypes.data_as(_i32p)
        info.parent2 = bufs["parent2"].ctypes.data_as(_i32p)
        info.distances2 = bufs["distances2"].ctypes.data_as(_f64p)
        info.pattern_codes = bufs["pattern_codes"].ctypes.data_as(_u8p)
        info.pattern_weights = bufs["pattern_weights"].ctypes.data_as(_i32p)
        info.site_pattern = bufs["site_pattern"].ctypes.data_as(_i32p)
        info.fixZ = bufs["fixZ"].ctypes.data_as(_u8p)
        _check(self._lib.pa_engine_element_info(self._h, c, C.byref(info)))
        out = dict(filtered=bool(info.filtered), GG=info.GG, P=info.P, root=info.root, n_nodes=info.n_nodes, n_slots=info.n_slots)
        if not out["filtered"]:
            out.update(missing=bufs["missing"].astype(bool), nodes=bufs["nodes"][:info.n_nodes].tolist(), children2=bufs["children2"],
                       parent2=bufs["parent2"], distances2=bufs["distances2"], pattern_codes=bufs["pattern_codes"][:info.P],
                       pattern_weights=bufs["pattern_weights"][:info.P], site_pattern=bufs["site_pattern"][:info.GG], fixZ=bufs["fixZ"])
        return out

    # ---- compute (CUDA only)
    def loglik_replay(self, elem, Z, n_rate, c_rate) -> np.ndarray:
        elem = np.ascontiguousarray(elem, dtype=np.int32)
        Z = np.ascontiguousarray(Z, dtype=np.int8).reshape(len(elem), self.N)
        nr = np.ascontiguousarray(n_rate, dtype=np.float64)
        cr = np.ascontiguousarray(c_rate, dtype=np.float64)
        out = np.zeros(len(elem))
        _check(self._lib.pa_engine_loglik_replay(self._h, len(elem), elem.ctypes.data_as(_i32p), Z.ctypes.data_as(_i8p),
                                                 nr.ctypes.data_as(_f64p), cr.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p)))
        return out

    def sample(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._last_ids = ids
        _check(self._lib.pa_engine_sample(self._h, ids.ctypes.data_as(_i32p), len(ids)))

    def sample_ex(self, ids, model_mask: int = 7, warm_start: bool = False, chain_iter: int = 0):
        """One outer CHAIN iteration (main.cpp:445-500): models in `model_mask`, optionally continuing every chain from its last state."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._last_ids = ids
        _check(self._lib.pa_engine_sample_ex(self._h, ids.ctypes.data_as(_i32p), len(ids), model_mask, int(bool(warm_start)), chain_iter))

    def hyper(self) -> dict:
        h = PaHyper()
        _check(self._lib.pa_engine_get_hyper(self._h, C.byref(h)))
        return h.as_dict()

    def set_hyper(self, **kw):
        h = PaHyper()
        _check(self._lib.pa_engine_get_hyper(self._h, C.byref(h)))
        for k, v in kw.items():
            setattr(h, k, float(v))
        _check(self._lib.pa_engine_set_hyper(self._h, C.byref(h)))

    def sample_hyper(self, it: int, ids) -> dict:
        """BPP::sample_hyperparam (bpp.cpp:324-477) from the current states of the full-model chains of `ids`."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        h = PaHyper()
        _check(self._lib.pa_engine_sample_hyper(self._h, it, ids.ctypes.data_as(_i32p), len(ids), C.byref(h)))
        return h.as_dict()

    def element_blob(self, c: int) -> np.ndarray:
        """The packed device image of element `c` (ElemHeader + arrays, csrc/pa_internal.h); empty for a filtered element."""
        size = C.c_int64(0)
        _check(self._lib.pa_engine_element_blob(self._h, int(c), None, 0, C.byref(size)))
        buf = np.zeros(size.value, dtype=np.uint8)
        if size.value:
            _check(self._lib.pa_engine_element_blob(self._h, int(c), buf.ctypes.data_as(_u8p), size.value, C.byref(size)))
        return buf

    def hyper_sums(self, ids) -> np.ndarray:
        """This rank's partial sums of the hyper-parameter sweep over its own `ids` (pa_engine_hyper_sums)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        sums = np.zeros(HYPER_NSUMS, dtype=np.float64)
        _check(self._lib.pa_engine_hyper_sums(self._h, ids.ctypes.data_as(_i32p), len(ids), sums.ctypes.data_as(_f64p)))
        return sums

    def sample_hyper_from_sums(self, it: int, sums) -> dict:
        """The sweep of sample_hyper from sums added over all ranks (pa_engine_sample_hyper_sums)."""
        sums = np.ascontiguousarray(sums, dtype=np.float64)
        assert sums.shape == (HYPER_NSUMS,)
        h = PaHyper()
        _check(self._lib.pa_engine_sample_hyper_sums(self._h, it, sums.ctypes.data_as(_f64p), C.byref(h)))
        return h.as_dict()

    def collect(self) -> Results:
        res = Results(len(self._last_ids), self.N)
        _check(self._lib.pa_engine_collect(self._h, C.byref(res.struct)))
        return res

    def run(self, ids) -> Results:
        self.sample(ids)
        return self.collect()

    def gt_replay(self, states, heights, thetas, cons_to_mis=0.5, ncons_to_mis=0.5):
        """PhyloAcc-GT groundwork (pa_engine_gt_replay): for every state -- a dict with elem, gchild [N, 2], root, gmissing [N], gheights [N],
        Z [N], n_rate, c_rate and `cross` {species: [(gene lineage, where it goes)]} (GTree::parent_gene2) -- the log-likelihood on the gene
        tree, the coalescent prior of the tree and log P(missing | Z).  Returns three arrays of length len(states)."""
        n, N = len(states), self.N
        elem = np.array([st["elem"] for st in states], dtype=np.int32)
        gchild = np.ascontiguousarray([st["gchild"] for st in states], dtype=np.int32).reshape(n, N, 2)
        groot = np.array([st["root"] for st in states], dtype=np.int32)
        gmis = np.ascontiguousarray([st["gmissing"] for st in states], dtype=np.uint8).reshape(n, N)
        gh = np.ascontiguousarray([st["gheights"] for st in states], dtype=np.float64).reshape(n, N)
        Z = np.ascontiguousarray([st["Z"] for st in states], dtype=np.int8).reshape(n, N)
        nr = np.array([st["n_rate"] for st in states], dtype=np.float64)
        cr = np.array([st["c_rate"] for st in states], dtype=np.float64)
        off, sp, nd, nx = [0], [], [], []
        for st in states:
            for s_, lin in st["cross"].items():
                for a_, b_ in lin:
                    sp.append(s_); nd.append(a_); nx.append(b_)
            off.append(len(sp))
        off = np.array(off, dtype=np.int64)
        sp, nd, nx = (np.array(x, dtype=np.int32) for x in (sp, nd, nx))
        heights = np.ascontiguousarray(heights, dtype=np.float64)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        gs = PaGtStates(n, N, elem.ctypes.data_as(_i32p), gchild.ctypes.data_as(_i32p), groot.ctypes.data_as(_i32p), gmis.ctypes.data_as(_u8p),
                        gh.ctypes.data_as(_f64p), Z.ctypes.data_as(_i8p), nr.ctypes.data_as(_f64p), cr.ctypes.data_as(_f64p),
                        off.ctypes.data_as(C.POINTER(C.c_int64)), sp.ctypes.data_as(_i32p), nd.ctypes.data_as(_i32p), nx.ctypes.data_as(_i32p),
                        heights.ctypes.data_as(_f64p), thetas.ctypes.data_as(_f64p), cons_to_mis, ncons_to_mis)
        ll, prior, lpz = np.zeros(n), np.zeros(n), np.zeros(n)
        _check(self._lib.pa_engine_gt_replay(self._h, C.byref(gs), ll.ctypes.data_as(_f64p), prior.ctypes.data_as(_f64p), lpz.ctypes.data_as(_f64p)))
        return ll, prior, lpz

    def gbgc_replay(self, elem, Z, B, sel_n, sel_c, gbc) -> np.ndarray:
        """PhyloAcc-ST-GBGC groundwork: replayed log-likelihoods under the GC-biased model (see pa_engine_gbgc_replay)."""
        elem = np.ascontiguousarray(elem, dtype=np.int32)
        Z = np.ascontiguousarray(Z, dtype=np.int8).reshape(len(elem), self.N)
        B = np.ascontiguousarray(B, dtype=np.int8).reshape(len(elem), self.N)
        sel_n, sel_c, gbc = (np.ascontiguousarray(x, dtype=np.float64) for x in (sel_n, sel_c, gbc))
        out = np.zeros(len(elem))
        _check(self._lib.pa_engine_gbgc_replay(self._h, len(elem), elem.ctypes.data_as(_i32p), Z.ctypes.data_as(_i8p), B.ctypes.data_as(_i8p),
                                               sel_n.ctypes.data_as(_f64p), sel_c.ctypes.data_as(_f64p), gbc.ctypes.data_as(_f64p),
                                               out.ctypes.data_as(_f64p)))
        return out

    def gbgc_integrate(self, elem, Z, B, fixZ, sel_n, sel_c, gbc, grate=0.8, lrate=0.1, prior_gb=0.2, cons_to_mis=0.01, ncons_to_mis=0.1):
        """PhyloAcc-ST-GBGC groundwork (pa_engine_gbgc_integrate): expected (child, parent) base pairs [n, N, 4, 4], the six (Z, B)
        emissions [n, N, 6] and [n, 8] = log_prob_back[root][0..5], ELBO term, upward log-likelihood."""
        elem = np.ascontiguousarray(elem, dtype=np.int32)
        n = len(elem)
        Z, B, fixZ = (np.ascontiguousarray(x, dtype=np.int8).reshape(n, self.N) for x in (Z, B, fixZ))
        sel_n, sel_c, gbc = (np.ascontiguousarray(x, dtype=np.float64) for x in (sel_n, sel_c, gbc))
        k = PaGbgcConsts(grate, lrate, prior_gb, cons_to_mis, ncons_to_mis)
        post, emis, rootv = np.zeros((n, self.N, 4, 4)), np.zeros((n, self.N, 6)), np.zeros((n, 8))
        _check(self._lib.pa_engine_gbgc_integrate(self._h, n, elem.ctypes.data_as(_i32p), Z.ctypes.data_as(_i8p), B.ctypes.data_as(_i8p),
                                                  fixZ.ctypes.data_as(_i8p), sel_n.ctypes.data_as(_f64p), sel_c.ctypes.data_as(_f64p),
                                                  gbc.ctypes.data_as(_f64p), C.byref(k), post.ctypes.data_as(_f64p), emis.ctypes.data_as(_f64p),
                                                  rootv.ctypes.data_as(_f64p)))
        return post, emis, rootv

    def gbgc_run(self, ids, targets, traces: bool = False, model_mask: int = 7, **params) -> dict:
        """The PhyloAcc-ST-GBGC chain (pa_engine_gbgc_run): M0 / M1 / M2 for the listed elements.  `targets` = node ids of TARGETSPECIES;
        keyword arguments override pa_gbgc_params_default (num_burn, num_mcmc, seed, sel_c0, ...).  Returns a dict of arrays indexed
        [element, model, ...] (loglik, loglik_max, med = (n_rate, c_rate, gBC), post_b, post_z, max_z, max_b, accept, filtered), with
        trace [.., 5, T], trace_z / trace_b [.., T, N] when traces=True, and kernel_ms / node_pattern_updates / chain_iterations."""
        prm = PaGbgcParams()
        self._lib.pa_gbgc_params_default(C.byref(prm))
        for k, v in params.items():
            if not hasattr(prm, k):
                raise KeyError(k)
            setattr(prm, k, v)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        targets = np.ascontiguousarray(targets, dtype=np.int32)
        n, N, T = len(ids), self.N, prm.num_burn + prm.num_mcmc
        out = dict(filtered=np.zeros(n, np.uint8), loglik=np.zeros((n, 3)), loglik_max=np.zeros((n, 3)), med=np.zeros((n, 3, 3)),
                   post_b=np.zeros((n, 3, N)), post_z=np.zeros((n, 3, N, 4)), max_z=np.zeros((n, 3, N), np.int8), max_b=np.zeros((n, 3, N), np.int8),
                   accept=np.zeros((n, 3, 2), np.int32))
        if traces:
            out.update(trace=np.zeros((n, 3, 5, T)), trace_z=np.zeros((n, 3, T, N), np.int8), trace_b=np.zeros((n, 3, T, N), np.int8))
        r = PaGbgcResults()
        r.n_ids, r.N, r.T = n, N, T
        for k, t in (("filtered", _u8p), ("loglik", _f64p), ("loglik_max", _f64p), ("med", _f64p), ("post_b", _f64p), ("post_z", _f64p),
                     ("max_z", _i8p), ("max_b", _i8p), ("accept", _i32p), ("trace", _f64p), ("trace_z", _i8p), ("trace_b", _i8p)):
            if k in out:
                setattr(r, k, out[k].ctypes.data_as(t))
        _check(self._lib.pa_engine_gbgc_run(self._h, ids.ctypes.data_as(_i32p), n, targets.ctypes.data_as(_i32p), len(targets), C.byref(prm),
                                            model_mask, C.byref(r)))
        out.update(kernel_ms=float(r.kernel_ms), node_pattern_updates=int(r.node_pattern_updates), chain_iterations=int(r.chain_iterations), pt_builds=int(r.pt_builds),
                   branch_pattern_updates=int(r.branch_pattern_updates), d2h_bytes=int(r.d2h_bytes),
                   phase_cycles=[int(x) for x in r.phase_cycles], kernel_launches=int(r.kernel_launches))
        return out

    def trace(self, i: int, model: int):
        p = self.params()
        T, N = p.burnin + p.mcmc, self.N
        b = dict(loglik=np.zeros(T), n_rate=np.zeros(T), c_rate=np.zeros(T), g_rate=np.zeros(T), l_rate=np.zeros(T),
                 l2_rate=np.zeros(T), Z=np.zeros((T, N), np.int8))
        t = PaTrace()
        t.T, t.N = T, N
        for k in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
            setattr(t, k, b[k].ctypes.data_as(_f64p))
        t.Z = b["Z"].ctypes.data_as(_i8p)
        _check(self._lib.pa_engine_get_trace(self._h, i, model, C.byref(t)))
        return b

    PHASES = ["init", "trace", "history", "emission", "z_update", "z_dirty_prune", "n_rate", "c_rate", "c_pass_only", "transitions"]

    def phase_cycles(self) -> dict:
        buf = (C.c_uint64 * 10)()
        _check(self._lib.pa_engine_phase_cycles(self._h, buf))
        return dict(zip(self.PHASES, [int(v) for v in buf]))

    def d2h_bytes(self) -> int:
        """Bytes copied device -> host for the last run so far (summary records, counters, traces that were asked for)."""
        return int(self.counters()["d2h_bytes"])

    def counters(self) -> dict:
        c = PaCounters()
        _check(self._lib.pa_engine_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in PaCounters._fields_}


def fp64_peak_tflops(device: int = 0, millis: int = 300) -> float:
    """Measured FP64 FMA throughput of the device (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)."""
    v = C.c_double(0.0)
    _check(load_library().pa_device_fp64_peak(device, millis, C.byref(v)))
    return v.value


def run_from_config(cfg_path: str) -> int:
    """main() of the reference (main.cpp:324-528)."""
    return load_library().pa_run_from_config(cfg_path.encode())
