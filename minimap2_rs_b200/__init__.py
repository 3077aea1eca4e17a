"""minimap2_rs_b200 — Python front-end (ctypes) of libmm2b200.so, the B200-native mapping hot path of mm2rs.

The function names mirror the reference crate's public API (src/lib.rs:1-6): sketch_sequence, build_index*,
Index.get / stats / calc_mid_occ / save_to_mmi / load_from_mmi, collect_query_minimizers, filter_query_minimizers,
build_anchors_filtered, chain_dp_all, write_paf; `map_batch` is main.rs:193-219 for every read of a batch.
Everything computes on the GPU through the C ABI in include/mm2b200.h; there is no CPU fallback: importing this
package without the built library, or calling it without an sm_100 device, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmm2b200.so")

MINI_DT = np.dtype([("key_span", "<u8"), ("rid_pos_strand", "<u8")])
ANCHOR_DT = np.dtype([("x", "<u8"), ("y", "<u8")])
PAF_DT = np.dtype([("read_id", "<u4"), ("rid", "<u4"), ("qlen", "<u4"), ("qstart", "<u4"), ("qend", "<u4"),
                   ("tlen", "<u4"), ("tstart", "<u4"), ("tend", "<u4"), ("nm", "<u4"), ("blen", "<u4"), ("cm", "<u4"),
                   ("s1", "<u4"), ("s2", "<u4"), ("dv", "<f4"), ("rl", "<u4"), ("strand", "u1"), ("mapq", "u1"),
                   ("tp", "u1"), ("flags", "u1")])

MM2_OK, MM2_E_ARG, MM2_E_IO, MM2_E_FORMAT, MM2_E_CUDA, MM2_E_OOM, MM2_E_REF_PANIC, MM2_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7


class Mm2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmm2b200 error %d: %s" % (code, msg))
        self.code = code


class ChainParams(C.Structure):  # lchain.rs:36-52
    _fields_ = [("max_dist_x", C.c_int32), ("max_dist_y", C.c_int32), ("bw", C.c_int32), ("max_chain_iter", C.c_int32),
                ("min_chain_score", C.c_int32), ("min_cnt", C.c_int32), ("chn_pen_gap", C.c_float),
                ("chn_pen_skip", C.c_float), ("max_chain_skip", C.c_int32), ("max_drop", C.c_int32),
                ("bw_long", C.c_int32), ("rmq_rescue_size", C.c_int32), ("rmq_rescue_ratio", C.c_float)]


class MapOpts(C.Structure):  # main.rs:55-89
    _fields_ = [("w", C.c_int32), ("k", C.c_int32), ("frac_top_repetitive", C.c_float), ("max_gap", C.c_int32),
                ("bw", C.c_int32), ("bw_long", C.c_int32), ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32),
                ("mask_level", C.c_float), ("pri_ratio", C.c_float), ("best_n", C.c_int32), ("q_occ_max", C.c_int32),
                ("q_occ_frac", C.c_float), ("mid_occ_floor", C.c_int32), ("want_stage_dump", C.c_int32)]


class _Chains(C.Structure):
    _fields_ = [("n_chains", C.c_size_t), ("chain_offs", C.c_void_p), ("chain_idx", C.c_void_p), ("scores", C.c_void_p),
                ("f", C.c_void_p), ("v", C.c_void_p), ("pprev", C.c_void_p)]


class _MapResult(C.Structure):
    _fields_ = [("n_recs", C.c_size_t), ("recs", C.c_void_p), ("n_panic", C.c_size_t), ("panic_reads", C.c_void_p),
                ("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_minimizers", C.c_uint64),
                ("n_minimizers_kept", C.c_uint64), ("n_anchors", C.c_uint64), ("n_rescued", C.c_uint64),
                ("mini_offs", C.c_void_p), ("minis", C.c_void_p), ("mini_keep", C.c_void_p), ("anchor_offs", C.c_void_p),
                ("anchors", C.c_void_p), ("f", C.c_void_p), ("v", C.c_void_p), ("pprev", C.c_void_p)]


# every symbol include/mm2b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "mm2_ctx_create", "mm2_ctx_destroy", "mm2_ctx_set_stream", "mm2_ctx_synchronize", "mm2_last_error", "mm2_free",
    "mm2_host_alloc", "mm2_host_alloc_on", "mm2_host_free", "mm2_sketch", "mm2_sketch_batch",
    "mm2_index_build_fasta", "mm2_index_build_seqs", "mm2_index_save_mmi", "mm2_index_save_mmi_khash", "mm2_index_load_mmi", "mm2_index_save_native",
    "mm2_index_load_native", "mm2_index_load_auto", "mm2_index_free", "mm2_index_get", "mm2_index_stats",
    "mm2_index_calc_mid_occ", "mm2_index_params", "mm2_index_seq", "mm2_index_get_ref_subseq", "mm2_index_build_timings",
    "mm2_filter_query_minimizers", "mm2_build_anchors_filtered", "mm2_chain_dp_all", "mm2_chains_free",
    "mm2_default_chain_params", "mm2_default_map_opts", "mm2_map_batch", "mm2_map_batch_device", "mm2_map_batch_packed", "mm2_pack_reads",
    "mm2_map_result_free",
    "mm2_paf_format", "mm2_paf_format_batch", "mm2_comm_get_unique_id", "mm2_comm_create", "mm2_comm_destroy", "mm2_comm_rank",
    "mm2_comm_barrier", "mm2_index_build_sharded", "mm2_index_build_sharded_emulated", "mm2_index_build_multi",
]

# include/mm2b200_diag.h (diagnostics: bench / profiling, not part of the drop-in boundary)
DIAG_SYMBOLS = ["mm2_ctx_launch_count", "mm2_ctx_last_timings", "mm2_ctx_count_cells", "mm2_ctx_last_cells", "mm2_shard_plan"]

_LIB = None


def lib():
    """load libmm2b200.so; raises if it has not been built (there is no fallback implementation)"""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("MM2_LIB_PATH", LIB_PATH)   # A/B builds of the same library (tools/, never a fallback)
    if not os.path.exists(path):
        raise ImportError("libmm2b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C minimap2_rs_b200/csrc`); minimap2_rs_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, sz, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
    L.mm2_last_error.restype = C.c_char_p
    L.mm2_host_alloc.restype = vp
    L.mm2_host_alloc.argtypes = [sz]
    L.mm2_host_alloc_on.restype = vp
    L.mm2_host_alloc_on.argtypes = [C.c_int, sz]
    L.mm2_host_free.argtypes = [vp]
    L.mm2_free.argtypes = [vp]
    L.mm2_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.mm2_ctx_destroy.argtypes = [vp]
    L.mm2_ctx_set_stream.argtypes = [vp, vp]
    L.mm2_ctx_synchronize.argtypes = [vp]
    L.mm2_ctx_launch_count.restype = C.c_uint64
    L.mm2_ctx_launch_count.argtypes = [vp]
    L.mm2_ctx_last_timings.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int)]
    L.mm2_ctx_count_cells.argtypes = [vp, C.c_int]
    L.mm2_ctx_last_cells.restype = C.c_uint64
    L.mm2_ctx_last_cells.argtypes = [vp]
    L.mm2_sketch.argtypes = [vp, vp, sz, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.mm2_sketch_batch.argtypes = [vp, vp, vp, sz, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.mm2_index_build_fasta.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mm2_index_build_seqs.argtypes = [vp, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    for nm in ("mm2_index_save_mmi", "mm2_index_save_native", "mm2_index_save_mmi_khash"):
        getattr(L, nm).argtypes = [vp, C.c_char_p]
    for nm in ("mm2_index_load_mmi", "mm2_index_load_native"):
        getattr(L, nm).argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.mm2_index_load_auto.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mm2_index_free.argtypes = [vp]
    L.mm2_index_get.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.POINTER(sz), C.POINTER(C.c_int)]
    L.mm2_index_stats.argtypes = [vp, u64p, C.POINTER(C.c_double), C.POINTER(C.c_double), u64p]
    L.mm2_index_calc_mid_occ.argtypes = [vp, C.c_float, C.POINTER(C.c_int32)]
    L.mm2_index_params.argtypes = [vp] + [C.POINTER(C.c_int32)] * 4 + [C.POINTER(C.c_uint32)]
    L.mm2_index_seq.argtypes = [vp, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32)]
    L.mm2_index_get_ref_subseq.argtypes = [vp, C.c_uint32, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(sz)]
    L.mm2_index_build_timings.argtypes = [vp, C.POINTER(C.c_float), u64p, u64p]
    L.mm2_filter_query_minimizers.argtypes = [vp, vp, C.POINTER(sz), C.c_int32, C.c_float]
    L.mm2_build_anchors_filtered.argtypes = [vp, vp, vp, sz, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(sz)]
    L.mm2_chain_dp_all.argtypes = [vp, vp, sz, C.POINTER(ChainParams), C.POINTER(_Chains)]
    L.mm2_chains_free.argtypes = [C.POINTER(_Chains)]
    L.mm2_default_chain_params.argtypes = [C.c_int32, C.POINTER(ChainParams)]
    L.mm2_default_chain_params.restype = None
    L.mm2_default_map_opts.argtypes = [C.POINTER(MapOpts)]
    L.mm2_default_map_opts.restype = None
    L.mm2_map_batch.argtypes = [vp, vp, vp, vp, sz, C.POINTER(MapOpts), C.POINTER(_MapResult)]
    L.mm2_map_batch_device.argtypes = [vp, vp, vp, vp, vp, sz, C.POINTER(MapOpts), C.POINTER(_MapResult)]
    L.mm2_map_batch_packed.argtypes = [vp, vp, vp, vp, sz, vp, sz, C.POINTER(MapOpts), C.POINTER(_MapResult)]
    L.mm2_pack_reads.argtypes = [vp, C.c_uint64, vp, vp, sz, C.POINTER(sz)]
    L.mm2_map_result_free.argtypes = [C.POINTER(_MapResult)]
    L.mm2_paf_format.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, sz]
    L.mm2_paf_format_batch.argtypes = [vp, C.POINTER(_MapResult), vp, C.POINTER(vp), C.POINTER(sz)]
    L.mm2_comm_get_unique_id.argtypes = [vp]
    L.mm2_comm_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.mm2_comm_destroy.argtypes = [vp]
    L.mm2_comm_destroy.restype = None
    L.mm2_comm_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mm2_comm_barrier.argtypes = [vp]
    L.mm2_index_build_sharded.argtypes = [vp, vp, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mm2_index_build_sharded_emulated.argtypes = [vp, C.c_int, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mm2_index_build_multi.argtypes = [vp, C.c_int, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.mm2_shard_plan.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    _LIB = L
    return L


def _check(rc):
    if rc != MM2_OK:
        raise Mm2Error(rc, lib().mm2_last_error().decode(errors="replace"))


def _copy_out(ptr, n, dtype, free=True):
    """copy n records from a library-owned buffer into a numpy array (and release the buffer)"""
    dtype = np.dtype(dtype)
    if ptr and n:
        arr = np.frombuffer((C.c_char * (n * dtype.itemsize)).from_address(ptr), dtype=dtype, count=n).copy()
    else:
        arr = np.zeros(0, dtype=dtype)
    if free and ptr:
        lib().mm2_free(ptr)
    return arr


def as_u8(seq):
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    if isinstance(seq, str):
        return np.frombuffer(seq.encode(), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


class PinnedBuffer:
    """page-locked host memory (mm2_host_alloc) exposed as numpy arrays; keep the object alive while arrays are in use"""

    def __init__(self, nbytes, device=None):
        self.nbytes = max(1, int(nbytes))
        self.p = lib().mm2_host_alloc(self.nbytes) if device is None else lib().mm2_host_alloc_on(int(device), self.nbytes)
        if not self.p:
            raise Mm2Error(MM2_E_OOM, lib().mm2_last_error().decode())
        self._raw = (C.c_char * self.nbytes).from_address(self.p)

    def array(self, dtype=np.uint8, count=None, offset=0):
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset) // dtype.itemsize
        return np.frombuffer(self._raw, dtype=dtype, count=count, offset=offset)

    def free(self):
        if self.p:
            self._raw = None
            lib().mm2_host_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def default_chain_params(k):
    """main.rs:105 default_chain_params"""
    p = ChainParams()
    lib().mm2_default_chain_params(k, C.byref(p))
    return p


def default_map_opts(w=10, k=15):
    o = MapOpts()
    lib().mm2_default_map_opts(C.byref(o))
    o.w, o.k = w, k
    return o


def apply_preset(preset, w, k):
    """main.rs:125 apply_preset -> (w, k)"""
    if preset == "map-ont":
        return 10, 15
    if preset in ("map-hifi", "lr:hq"):
        return 10, 19
    if preset == "sr":
        return 11, 21
    return w, k


class Context:
    """one GPU + stream + scratch arenas (mm2_ctx_t)"""

    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        _check(lib().mm2_ctx_create(device, C.byref(self.h)))
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream):
        _check(lib().mm2_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        _check(lib().mm2_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(lib().mm2_ctx_launch_count(self.h))

    def count_cells(self, on=True):
        """diagnostic: make the chaining kernels count DP cells (lchain.rs:80 iterations); read with last_cells"""
        _check(lib().mm2_ctx_count_cells(self.h, 1 if on else 0))

    @property
    def last_cells(self):
        return int(lib().mm2_ctx_last_cells(self.h))

    def last_timings(self):
        names, ms, n = C.c_void_p(), C.c_void_p(), C.c_int()
        _check(lib().mm2_ctx_last_timings(self.h, C.byref(names), C.byref(ms), C.byref(n)))
        out = {}
        if n.value:
            vals = np.frombuffer((C.c_char * (4 * n.value)).from_address(ms.value), dtype=np.float32, count=n.value)
            addr = names.value
            for i in range(n.value):
                s = C.string_at(addr)
                out[s.decode()] = out.get(s.decode(), 0.0) + float(vals[i])
                addr += len(s) + 1
        return out

    # ---- sketch.rs:29 ------------------------------------------------------------------------------------
    def sketch_sequence(self, seq, w, k, rid=0, is_hpc=False):
        s = as_u8(seq)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().mm2_sketch(self.h, s.ctypes.data, s.size, w, k, rid, int(is_hpc), C.byref(out), C.byref(n)))
        return _copy_out(out.value, n.value, MINI_DT)

    def sketch_batch(self, cat, offs, w, k, rid_base=0, rid_step=0, is_hpc=False):
        cat = as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        out, oo = C.c_void_p(), C.c_void_p()
        _check(lib().mm2_sketch_batch(self.h, cat.ctypes.data, offs.ctypes.data, offs.size - 1, w, k, rid_base, rid_step,
                                      int(is_hpc), C.byref(out), C.byref(oo)))
        o = _copy_out(oo.value, offs.size, np.uint64)
        return _copy_out(out.value, int(o[-1]), MINI_DT), o

    # ---- seeds.rs ------------------------------------------------------------------------------------------
    def collect_query_minimizers(self, seq, w, k):
        """seeds.rs:7: rid 0, no HPC"""
        return self.sketch_sequence(seq, w, k, 0, False)

    def filter_query_minimizers(self, mv, q_occ_max=10, q_occ_frac=0.01):
        mv = np.ascontiguousarray(mv, dtype=MINI_DT).copy()
        n = C.c_size_t(mv.size)
        _check(lib().mm2_filter_query_minimizers(self.h, mv.ctypes.data, C.byref(n), q_occ_max, q_occ_frac))
        return mv[:n.value].copy()

    def build_anchors_filtered(self, idx, mv, qlen, mid_occ):
        mv = np.ascontiguousarray(mv, dtype=MINI_DT)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().mm2_build_anchors_filtered(self.h, idx.h, mv.ctypes.data, mv.size, qlen, mid_occ, C.byref(out), C.byref(n)))
        return _copy_out(out.value, n.value, ANCHOR_DT)

    def build_anchors(self, idx, mv, qlen):
        """seeds.rs:38"""
        return self.build_anchors_filtered(idx, mv, qlen, 0x7FFFFFFF)

    # ---- lchain.rs ------------------------------------------------------------------------------------------
    def chain_dp_all(self, anchors, p):
        a = np.ascontiguousarray(anchors, dtype=ANCHOR_DT)
        ch = _Chains()
        _check(lib().mm2_chain_dp_all(self.h, a.ctypes.data, a.size, C.byref(p), C.byref(ch)))
        try:
            m = ch.n_chains
            offs = _copy_out(ch.chain_offs, m + 1, np.uint64, free=False)
            idx = _copy_out(ch.chain_idx, int(offs[-1]) if m else 0, np.uint64, free=False)
            res = dict(chains=[idx[int(offs[i]):int(offs[i + 1])].astype(np.int64) for i in range(m)],
                       scores=_copy_out(ch.scores, m, np.int32, free=False),
                       f=_copy_out(ch.f, a.size, np.int32, free=False), v=_copy_out(ch.v, a.size, np.int32, free=False),
                       pprev=_copy_out(ch.pprev, a.size, np.int64, free=False))
        finally:
            lib().mm2_chains_free(C.byref(ch))
        return res

    def chain_dp(self, anchors, p):
        """lchain.rs:54"""
        r = self.chain_dp_all(anchors, p)
        return r["chains"][0] if r["chains"] else np.zeros(0, dtype=np.int64)

    # ---- main.rs:193-219 over a batch ---------------------------------------------------------------------------
    def map_batch(self, idx, cat, offs, opts=None, device_ptrs=None):
        """cat/offs: host arrays (H2D inside), or device_ptrs=(d_cat, d_offs) for reads already in HBM.
        returns MapResult"""
        opts = opts or default_map_opts(idx.w, idx.k)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        res = _MapResult()
        if device_ptrs is None:
            cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 and cat.flags.c_contiguous else as_u8(cat)
            _check(lib().mm2_map_batch(self.h, idx.h, cat.ctypes.data, offs.ctypes.data, offs.size - 1, C.byref(opts), C.byref(res)))
        else:
            _check(lib().mm2_map_batch_device(self.h, idx.h, C.c_void_p(device_ptrs[0]), C.c_void_p(device_ptrs[1]),
                                              offs.ctypes.data, offs.size - 1, C.byref(opts), C.byref(res)))
        return MapResult(idx, res, offs.size - 1)

    def map_batch_packed(self, idx, packed, n_pos, offs, opts=None):
        """reads as 2-bit codes + positions of the non-ACGT bases (pack_reads): a quarter of the H2D bytes of map_batch"""
        opts = opts or default_map_opts(idx.w, idx.k)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n_pos = np.ascontiguousarray(n_pos, dtype=np.uint64)
        assert isinstance(packed, np.ndarray) and packed.dtype == np.uint8 and packed.flags.c_contiguous
        res = _MapResult()
        _check(lib().mm2_map_batch_packed(self.h, idx.h, packed.ctypes.data, n_pos.ctypes.data if n_pos.size else None, n_pos.size,
                                          offs.ctypes.data, offs.size - 1, C.byref(opts), C.byref(res)))
        return MapResult(idx, res, offs.size - 1)

    def close(self):
        if getattr(self, "h", None):
            lib().mm2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_reads(cat, out=None):
    """mm2_pack_reads: ASCII bases -> (packed 2-bit codes, positions of the bases that are not ACGTacgt).
    out: optional uint8 buffer of at least (n + 3) // 4 + 64 bytes (e.g. page-locked)"""
    cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 and cat.flags.c_contiguous else as_u8(cat)
    n = cat.size
    need = (n + 3) // 4 + 64
    packed = out if out is not None else np.zeros(need, dtype=np.uint8)
    assert packed.size >= need
    cap = 1024
    while True:
        n_pos = np.zeros(cap, dtype=np.uint64)
        nn = C.c_size_t()
        _check(lib().mm2_pack_reads(cat.ctypes.data, n, packed.ctypes.data, n_pos.ctypes.data, cap, C.byref(nn)))
        if nn.value <= cap:
            return packed, n_pos[:nn.value].copy()
        cap = nn.value


class Comm:
    """the library's NCCL communicator for one context (mm2_comm_t): one rank per GPU"""

    def __init__(self, ctx, unique_id, nranks, rank):
        self.h = C.c_void_p()
        self.ctx = ctx
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(lib().mm2_comm_create(ctx.h, buf, nranks, rank, C.byref(self.h)))
        self.nranks, self.rank = nranks, rank

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        _check(lib().mm2_comm_get_unique_id(buf))
        return bytes(buf.raw)

    @classmethod
    def from_torch(cls, ctx, dist, group=None):
        """collective over a torch.distributed group: rank 0's unique id is broadcast through the group's store"""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(ctx, box[0], world, rank)

    def barrier(self):
        _check(lib().mm2_comm_barrier(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().mm2_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_plan(offs, w, k, flag, nranks, rank):
    """diagnostic (host only): the share of `rank` in the sharded index build"""
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    _check(lib().mm2_shard_plan(offs.ctypes.data, offs.size - 1, w, k, flag, nranks, rank, out.ctypes.data))
    keys = ("tile_path", "lo", "hi", "sketch_byte_lo", "sketch_byte_hi", "word_lo", "word_hi", "upload_bytes")
    return dict(zip(keys, (int(x) for x in out)))


class MapResult:
    def __init__(self, idx, res, nreads):
        self._res = res
        self._idx = idx
        self.nreads = nreads
        self.n_recs = int(res.n_recs)
        self._recs = None            # copied out of the library's buffer on first use (6.4 MB per 100k reads)
        self.panic_reads = _copy_out(res.panic_reads, res.n_panic, np.uint32, free=False)
        self.stats = dict(n_reads=res.n_reads, n_bases=res.n_bases, n_minimizers=res.n_minimizers,
                          n_minimizers_kept=res.n_minimizers_kept, n_anchors=res.n_anchors, n_rescued=res.n_rescued)
        self.stage = None
        if res.mini_offs:
            mo = _copy_out(res.mini_offs, nreads + 1, np.uint64, free=False)
            ao = _copy_out(res.anchor_offs, nreads + 1, np.uint64, free=False)
            nm, na = int(mo[-1]), int(ao[-1])
            self.stage = dict(mini_offs=mo, anchor_offs=ao, minis=_copy_out(res.minis, nm, MINI_DT, free=False),
                              mini_keep=_copy_out(res.mini_keep, nm, np.uint8, free=False),
                              anchors=_copy_out(res.anchors, na, ANCHOR_DT, free=False),
                              f=_copy_out(res.f, na, np.int32, free=False), v=_copy_out(res.v, na, np.int32, free=False),
                              pprev=_copy_out(res.pprev, na, np.int32, free=False))

    @property
    def recs(self):
        """the PAF records (PAF_DT), one per kept chain, in read order"""
        if self._recs is None:
            if self._res is None:
                raise Mm2Error(MM2_E_ARG, "MapResult is closed")
            self._recs = _copy_out(self._res.recs, self._res.n_recs, PAF_DT, free=False)
        return self._recs

    def paf_lines(self, qnames=None):
        """paf.rs:238 write_paf_many_with_scores for the whole batch -> list of lines"""
        arr = None
        keep = None
        if qnames is not None:
            keep = [n.encode() if isinstance(n, str) else bytes(n) for n in qnames]
            arr = (C.c_char_p * len(keep))(*keep)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().mm2_paf_format_batch(self._idx.h, C.byref(self._res), arr, C.byref(out), C.byref(n)))
        txt = C.string_at(out.value, n.value).decode()
        lib().mm2_free(out)
        return txt.split("\n")[:-1] if txt else []

    def close(self):
        if self._res is not None:
            lib().mm2_map_result_free(C.byref(self._res))
            self._res = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """index.rs:33 `Index`, resident in HBM"""

    def __init__(self, handle):
        self.h = handle
        w, k, b, f, n = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint32()
        _check(lib().mm2_index_params(self.h, C.byref(w), C.byref(k), C.byref(b), C.byref(f), C.byref(n)))
        self.w, self.k, self.b, self.flag, self.n_seq = w.value, k.value, b.value, f.value, n.value

    @classmethod
    def build(cls, ctx, cat, offs, names, w=10, k=15, b=14, flag=0):
        """index.rs:427 build_index_from_fasta, from records in host memory"""
        cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 else as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        h = C.c_void_p()
        _check(lib().mm2_index_build_seqs(ctx.h, cat.ctypes.data, offs.ctypes.data, arr, len(enc), w, k, b, flag, C.byref(h)))
        return cls(h)

    @classmethod
    def build_sharded(cls, ctx, comm, cat, offs, names, w=10, k=15, b=14, flag=0):
        """collective over `comm` (mm2_index_build_sharded): every rank passes the same host arrays and gets its replica"""
        cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 else as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        h = C.c_void_p()
        _check(lib().mm2_index_build_sharded(ctx.h, comm.h, cat.ctypes.data, offs.ctypes.data, arr, len(enc), w, k, b, flag, C.byref(h)))
        return cls(h)

    @classmethod
    def build_sharded_emulated(cls, ctx, nranks, cat, offs, names, w=10, k=15, b=14, flag=0):
        """the sharded build's phases for `nranks` virtual ranks on one GPU (single-GPU test of that path)"""
        cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 else as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        h = C.c_void_p()
        _check(lib().mm2_index_build_sharded_emulated(ctx.h, nranks, cat.ctypes.data, offs.ctypes.data, arr, len(enc), w, k, b, flag, C.byref(h)))
        return cls(h)

    @classmethod
    def build_multi(cls, ctxs, cat, offs, names, w=10, k=15, b=14, flag=0):
        """one process, several GPUs (mm2_index_build_multi): -> one replica per context"""
        cat = cat if isinstance(cat, np.ndarray) and cat.dtype == np.uint8 else as_u8(cat)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        out = (C.c_void_p * len(ctxs))()
        _check(lib().mm2_index_build_multi(hs, len(ctxs), cat.ctypes.data, offs.ctypes.data, arr, len(enc), w, k, b, flag, out))
        return [cls(C.c_void_p(o)) for o in out]

    @classmethod
    def build_from_fasta(cls, ctx, path, w=10, k=15, b=14, flag=0):
        h = C.c_void_p()
        _check(lib().mm2_index_build_fasta(ctx.h, path.encode(), w, k, b, flag, C.byref(h)))
        return cls(h)

    @classmethod
    def load_from_mmi(cls, ctx, path):
        h = C.c_void_p()
        _check(lib().mm2_index_load_mmi(ctx.h, path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def load_from_file(cls, ctx, path):
        h = C.c_void_p()
        _check(lib().mm2_index_load_native(ctx.h, path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def load_auto(cls, ctx, path, w=10, k=15, b=14, flag=0):
        """main.rs:135 load_index_auto"""
        h = C.c_void_p()
        _check(lib().mm2_index_load_auto(ctx.h, path.encode(), w, k, b, flag, C.byref(h)))
        return cls(h)

    def save_to_mmi(self, path):
        _check(lib().mm2_index_save_mmi(self.h, path.encode()))

    def save_to_mmi_khash(self, path):
        """.mmi with every bucket's entries in C minimap2's khash slot order (mm_idx_dump)"""
        _check(lib().mm2_index_save_mmi_khash(self.h, path.encode()))

    def save_to_file(self, path):
        _check(lib().mm2_index_save_native(self.h, path.encode()))

    def get(self, minier):
        """index.rs:143 -> (kind, occurrences): kind 0 None, 1 Single, 2 Multi"""
        occ, n, kind = C.c_void_p(), C.c_size_t(), C.c_int()
        _check(lib().mm2_index_get(self.h, int(minier), C.byref(occ), C.byref(n), C.byref(kind)))
        return kind.value, _copy_out(occ.value, n.value, np.uint64)

    def stats(self):
        nk, tl, ao, sp = C.c_uint64(), C.c_uint64(), C.c_double(), C.c_double()
        _check(lib().mm2_index_stats(self.h, C.byref(nk), C.byref(ao), C.byref(sp), C.byref(tl)))
        return nk.value, ao.value, sp.value, tl.value

    def calc_mid_occ(self, frac=2e-4):
        v = C.c_int32()
        _check(lib().mm2_index_calc_mid_occ(self.h, frac, C.byref(v)))
        return v.value

    def seq(self, rid):
        name, ln = C.c_char_p(), C.c_uint32()
        _check(lib().mm2_index_seq(self.h, rid, C.byref(name), C.byref(ln)))
        return name.value.decode(), ln.value

    def get_ref_subseq(self, rid, st, en):
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().mm2_index_get_ref_subseq(self.h, rid, st, en, C.byref(out), C.byref(n)))
        return _copy_out(out.value, n.value, np.uint8).tobytes()

    def build_timings(self):
        ms = (C.c_float * 5)()
        nb, nm = C.c_uint64(), C.c_uint64()
        _check(lib().mm2_index_build_timings(self.h, ms, C.byref(nb), C.byref(nm)))
        return dict(sketch_ms=ms[0], sort_ms=ms[1], bucket_build_ms=ms[2], pack_ms=ms[3], total_ms=ms[4], n_bases=nb.value,
                    n_minimizers=nm.value)

    def close(self):
        if getattr(self, "h", None):
            lib().mm2_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_paf(rec, qname, tname):
    """paf.rs:224 write_paf for one PAF_DT record"""
    r = np.ascontiguousarray(rec, dtype=PAF_DT).reshape(1)
    buf = C.create_string_buffer(len(qname) + len(tname) + 256)
    n = lib().mm2_paf_format(r.ctypes.data, qname.encode(), tname.encode(), buf, len(buf))
    if n < 0:
        _check(n)
    return buf.raw[:n].decode()
