"""ctypes front-end of tools/libmm2gen.so — deterministic synthetic genomes/reads."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmm2gen.so")
    src = os.path.join(_HERE, "gen.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, sz = C.c_void_p, C.c_size_t
        L.mm2gen_genome.argtypes = [C.c_uint64, vp, sz, C.c_double, C.c_double]
        L.mm2gen_repeat_genome.argtypes = [C.c_uint64, vp, sz, C.c_double, C.c_double]
        L.mm2gen_reads.argtypes = [C.c_uint64, vp, vp, sz, sz, sz, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp]
        L.mm2gen_reads_from.argtypes = [C.c_uint64, vp, vp, sz, sz, sz, sz, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def genome(seed, length, n_run_rate=0.0, n_run_mean=50.0, out=None):
    g = out if out is not None else np.empty(length, dtype=np.uint8)
    lib().mm2gen_genome(seed, g.ctypes.data, length, n_run_rate, n_run_mean)
    return g


def repeat_genome(seed, length, tandem_frac=0.4, dispersed_frac=0.2):
    g = np.empty(length, dtype=np.uint8)
    lib().mm2gen_repeat_genome(seed, g.ctypes.data, length, tandem_frac, dispersed_frac)
    return g


def reads(seed, genome_cat, seq_offs, nreads, read_len, p_sub, p_ins, p_del, out=None, truth=False, first=0):
    """returns (cat uint8[nreads*read_len], offs uint64[nreads+1]) (+ (src_seq, src_pos, src_rev) if truth); `first` = index of
    the first read within the seeded read set (reads [first, first + nreads) of it)"""
    seq_offs = np.ascontiguousarray(seq_offs, dtype=np.uint64)
    cat = out if out is not None else np.empty(nreads * read_len, dtype=np.uint8)
    ss = np.zeros(nreads, dtype=np.uint32)
    sp = np.zeros(nreads, dtype=np.uint64)
    sr = np.zeros(nreads, dtype=np.uint8)
    lib().mm2gen_reads_from(seed, genome_cat.ctypes.data, seq_offs.ctypes.data, seq_offs.size - 1, first, nreads, read_len,
                            p_sub, p_ins, p_del, cat.ctypes.data, ss.ctypes.data, sp.ctypes.data, sr.ctypes.data)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(read_len)
    return (cat, offs, (ss, sp, sr)) if truth else (cat, offs)
