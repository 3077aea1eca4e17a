"""ctypes front-end of the CPU restatement (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under minimap2_rs_b200/ imports this module.
PARITY UNPINNED: the reference ships no golden vectors and cannot be built here (see mm2_oracle.hpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MINI_DT = np.dtype([("key_span", "<u8"), ("rid_pos_strand", "<u8")])
ANCHOR_DT = np.dtype([("x", "<u8"), ("y", "<u8")])


class ChainParams(C.Structure):
    # mirrors lchain.rs:36-52 field for field
    _fields_ = [("max_dist_x", C.c_int32), ("max_dist_y", C.c_int32), ("bw", C.c_int32), ("max_chain_iter", C.c_int32),
                ("min_chain_score", C.c_int32), ("min_cnt", C.c_int32), ("chn_pen_gap", C.c_float),
                ("chn_pen_skip", C.c_float), ("max_chain_skip", C.c_int32), ("max_drop", C.c_int32),
                ("bw_long", C.c_int32), ("rmq_rescue_size", C.c_int32), ("rmq_rescue_ratio", C.c_float)]


class AlignOpts(C.Structure):
    _fields_ = [("w", C.c_int32), ("k", C.c_int32), ("frac_top_repetitive", C.c_float), ("max_gap", C.c_int32),
                ("bw", C.c_int32), ("bw_long", C.c_int32), ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32),
                ("mask_level", C.c_float), ("pri_ratio", C.c_float), ("best_n", C.c_int32)]

    @classmethod
    def default(cls, w=10, k=15):
        return cls(w, k, 2e-4, 5000, -1, -1, 3, 40, 0.5, 0.8, 5)


class AlignStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_reads", "n_bases", "n_minimizers", "n_minimizers_kept", "n_anchors",
                                          "cells", "n_rescued", "n_lines", "n_panic")] + [("seconds", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mm2_oracle.cpp", "oracle_capi.cpp", "mm2_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, sz, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
        L.orc_free.argtypes = [vp]
        L.orc_sketch.argtypes = [vp, sz, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(vp), C.POINTER(sz)]
        L.orc_index_build.restype = vp
        L.orc_index_build.argtypes = [vp, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_index_build_fasta.restype = vp
        L.orc_index_build_fasta.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_index_free.argtypes = [vp]
        L.orc_index_save_mmi.argtypes = [vp, C.c_char_p]
        L.orc_index_load_mmi.restype = vp
        L.orc_index_load_mmi.argtypes = [C.c_char_p]
        L.orc_index_save_native.argtypes = [vp, C.c_char_p]
        L.orc_index_load_native.restype = vp
        L.orc_index_load_native.argtypes = [C.c_char_p]
        L.orc_index_stats.argtypes = [vp, u64p, C.POINTER(C.c_double), C.POINTER(C.c_double), u64p]
        L.orc_index_calc_mid_occ.restype = C.c_int32
        L.orc_index_calc_mid_occ.argtypes = [vp, C.c_float]
        L.orc_index_params.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
        L.orc_index_get.argtypes = [vp, C.c_uint64, vp, sz, C.POINTER(sz)]
        L.orc_filter_query_minimizers.restype = sz
        L.orc_filter_query_minimizers.argtypes = [vp, sz, C.c_int32, C.c_float]
        L.orc_build_anchors_filtered.argtypes = [vp, vp, sz, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(sz)]
        L.orc_chain_dp_all.argtypes = [vp, sz, C.POINTER(ChainParams), vp, vp, vp, u64p, C.POINTER(sz),
                                       C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.orc_default_chain_params.argtypes = [C.c_int32, C.POINTER(ChainParams)]
        L.orc_chain_pen_lut.argtypes = [C.c_float, C.c_int32, vp]
        L.orc_align_batch.argtypes = [vp, vp, vp, vp, vp, sz, C.POINTER(AlignOpts), C.c_int, C.POINTER(vp),
                                      C.POINTER(sz), C.POINTER(AlignStats)]
        _LIB = L
    return _LIB


def _take(ptr, n, dtype):
    """copy n records out of a malloc'd buffer and free it"""
    if n:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).copy()
    else:
        arr = np.zeros(0, dtype=dtype)
    lib().orc_free(ptr)
    return arr


def as_u8(seq):
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def sketch(seq, w, k, rid=0, is_hpc=False):
    s = as_u8(seq)
    out, n = C.c_void_p(), C.c_size_t()
    lib().orc_sketch(s.ctypes.data, s.size, w, k, rid, int(is_hpc), C.byref(out), C.byref(n))
    return _take(out, n.value, MINI_DT)


def cat_names(names):
    bs = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8), offs


class Index:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index handle is NULL")
        self.h = C.c_void_p(handle)

    @classmethod
    def build(cls, cat, offs, names, w=10, k=15, b=14, flag=0, threads=1):
        cat = as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ncat, noffs = cat_names(names)
        return cls(lib().orc_index_build(cat.ctypes.data, offs.ctypes.data, ncat.ctypes.data, noffs.ctypes.data,
                                         len(names), w, k, b, flag, threads))

    @classmethod
    def build_fasta(cls, path, w=10, k=15, b=14, flag=0, threads=1):
        return cls(lib().orc_index_build_fasta(path.encode(), w, k, b, flag, threads))

    @classmethod
    def load_mmi(cls, path):
        return cls(lib().orc_index_load_mmi(path.encode()))

    @classmethod
    def load_native(cls, path):
        return cls(lib().orc_index_load_native(path.encode()))

    def save_mmi(self, path):
        if lib().orc_index_save_mmi(self.h, path.encode()) != 0:
            raise IOError(path)

    def save_native(self, path):
        if lib().orc_index_save_native(self.h, path.encode()) != 0:
            raise IOError(path)

    def stats(self):
        nk, tl, ao, sp = C.c_uint64(), C.c_uint64(), C.c_double(), C.c_double()
        lib().orc_index_stats(self.h, C.byref(nk), C.byref(ao), C.byref(sp), C.byref(tl))
        return nk.value, ao.value, sp.value, tl.value

    def calc_mid_occ(self, frac=2e-4):
        return lib().orc_index_calc_mid_occ(self.h, frac)

    def params(self):
        a = (C.c_int32 * 4)()
        n = C.c_uint32()
        lib().orc_index_params(self.h, a, C.byref(n))
        return dict(w=a[0], k=a[1], b=a[2], flag=a[3], n_seq=n.value)

    def get(self, minier, cap=1 << 16):
        occ = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t()
        r = lib().orc_index_get(self.h, int(minier), occ.ctypes.data, cap, C.byref(n))
        return r, occ[:min(n.value, cap)].copy()

    def anchors(self, mv, qlen, mid_occ):
        mv = np.ascontiguousarray(mv, dtype=MINI_DT)
        out, n = C.c_void_p(), C.c_size_t()
        lib().orc_build_anchors_filtered(self.h, mv.ctypes.data, mv.size, qlen, mid_occ, C.byref(out), C.byref(n))
        return _take(out, n.value, ANCHOR_DT)

    def align_batch(self, cat, offs, names, opts=None, threads=1):
        """returns (list of PAF lines, AlignStats)"""
        cat = as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ncat, noffs = cat_names(names)
        opts = opts or AlignOpts.default()
        out, n, st = C.c_void_p(), C.c_size_t(), AlignStats()
        lib().orc_align_batch(self.h, cat.ctypes.data, offs.ctypes.data, ncat.ctypes.data, noffs.ctypes.data,
                              len(names), C.byref(opts), threads, C.byref(out), C.byref(n), C.byref(st))
        txt = C.string_at(out.value, n.value).decode()
        lib().orc_free(out)
        return (txt.split("\n")[:-1] if txt else []), st

    def close(self):
        if self.h:
            lib().orc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def filter_query_minimizers(mv, q_occ_max=10, q_occ_frac=0.01):
    mv = np.ascontiguousarray(mv, dtype=MINI_DT).copy()
    n = lib().orc_filter_query_minimizers(mv.ctypes.data, mv.size, q_occ_max, q_occ_frac)
    return mv[:n].copy()


def default_chain_params(k):
    p = ChainParams()
    lib().orc_default_chain_params(k, C.byref(p))
    return p


def chain_dp_all(anchors, p):
    """returns dict(f, v, pprev, cells, chains=[np.array], scores)"""
    a = np.ascontiguousarray(anchors, dtype=ANCHOR_DT)
    n = a.size
    f = np.zeros(n, dtype=np.int32)
    v = np.zeros(n, dtype=np.int32)
    pp = np.full(n, -1, dtype=np.int64)
    cells, nch = C.c_uint64(), C.c_size_t()
    co, ci, sc = C.c_void_p(), C.c_void_p(), C.c_void_p()
    lib().orc_chain_dp_all(a.ctypes.data, n, C.byref(p), f.ctypes.data, v.ctypes.data, pp.ctypes.data, C.byref(cells),
                           C.byref(nch), C.byref(co), C.byref(ci), C.byref(sc))
    m = nch.value
    offs = _take(co, m + 1, np.dtype("<u8"))
    tot = int(offs[-1]) if m else 0
    idx = _take(ci, tot, np.dtype("<u8")) if tot else (lib().orc_free(ci), np.zeros(0, dtype=np.uint64))[1]
    scores = _take(sc, m, np.dtype("<i4")) if m else (lib().orc_free(sc), np.zeros(0, dtype=np.int32))[1]
    chains = [idx[int(offs[i]):int(offs[i + 1])].astype(np.int64) for i in range(m)]
    return dict(f=f, v=v, pprev=pp, cells=cells.value, chains=chains, scores=scores)


def chain_pen_lut(chn_pen_gap, n):
    out = np.zeros(n, dtype=np.int32)
    lib().orc_chain_pen_lut(chn_pen_gap, n, out.ctypes.data)
    return out
