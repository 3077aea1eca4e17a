/* mm2b200.h — C ABI of libmm2b200.so: the B200-native (sm_100a) mapping hot path of mm2rs
 * (sketch -> .mmi index build -> seed lookup / anchors -> chaining -> PAF record).
 *
 * This is the drop-in boundary a Rust `extern "C"` shim in the reference crate would bind.
 * Each entry point names the reference function it replaces (file:line under the reference's src/).
 * Plain pointers and sizes only; every function returns an int status (0 = MM2_OK, < 0 = error, message through
 * mm2_last_error()); nothing throws or aborts across the boundary.  There is NO CPU fallback: every compute entry
 * point runs CUDA kernels and fails with MM2_E_CUDA when no sm_100 device / context is available.
 *
 * Ownership: arrays returned through `**out` parameters are host memory owned by the library and are released
 * with mm2_free().  Inputs are never retained past the call.  A context is used by one host thread at a time;
 * an index is immutable after build/load and may be shared by several contexts on the same device.
 */
#ifndef MM2B200_H
#define MM2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM2_OK 0
#define MM2_E_ARG (-1)       /* bad argument; also the inputs sketch.rs:30-32 asserts on (empty seq, w, k out of range) */
#define MM2_E_IO (-2)        /* file could not be opened / read / written (anyhow io errors, index.rs:156,233,309,361,427) */
#define MM2_E_FORMAT (-3)    /* bad magic / truncated index (index.rs:312,366) */
#define MM2_E_CUDA (-4)      /* CUDA runtime error or no usable device */
#define MM2_E_OOM (-5)       /* host or device allocation failed */
#define MM2_E_REF_PANIC (-6) /* input on which the reference panics (odd-rid anchors, seeds.rs:64-71 + paf.rs:148-150) */
#define MM2_E_UNSUPPORTED (-7)

typedef struct mm2_ctx mm2_ctx_t;     /* one device + stream + scratch arenas */
typedef struct mm2_index mm2_index_t; /* index.rs:33-42 `Index`, resident in HBM */

/* sketch.rs:15-19 `Minimizer` */
typedef struct { uint64_t key_span; uint64_t rid_pos_strand; } mm2_mini_t;
/* seeds.rs:4-5 `Anchor` */
typedef struct { uint64_t x; uint64_t y; } mm2_anchor_t;
/* lchain.rs:36-52 `ChainParams`, field for field */
typedef struct {
  int32_t max_dist_x, max_dist_y, bw, max_chain_iter, min_chain_score, min_cnt;
  float chn_pen_gap, chn_pen_skip;
  int32_t max_chain_skip, max_drop, bw_long, rmq_rescue_size;
  float rmq_rescue_ratio;
} mm2_chain_params_t;

/* ---- context ------------------------------------------------------------------------------------------- */
int mm2_ctx_create(int device, mm2_ctx_t** out);
void mm2_ctx_destroy(mm2_ctx_t* ctx);
/* run all work of this context on a caller-owned cudaStream_t (e.g. the current PyTorch stream); NULL = own stream */
int mm2_ctx_set_stream(mm2_ctx_t* ctx, void* cuda_stream);
int mm2_ctx_synchronize(mm2_ctx_t* ctx);
const char* mm2_last_error(void);
void mm2_free(void* p);               /* releases any `**out` array */
void* mm2_host_alloc(size_t bytes);   /* page-locked host buffer (fast H2D/D2H); NULL on failure */
void* mm2_host_alloc_on(int device, size_t bytes);   /* the same, on the NUMA node next to that GPU (multi-rank uploads) */
void mm2_host_free(void* p);
/* (launch counters, stage timers and the DP-cell counter used by bench.py are diagnostics: see mm2b200_diag.h) */

/* ---- sketch (sketch.rs:29 sketch_sequence) -------------------------------------------------------------- */
int mm2_sketch(mm2_ctx_t* ctx, const uint8_t* seq, size_t len, int w, int k, uint32_t rid, int is_hpc,
               mm2_mini_t** out, size_t* n);
/* many sequences in one launch: cat = concatenated bases, offs[nseq+1]; rid of sequence i = rid_base + i*rid_step.
 * out_offs[nseq+1] delimits each sequence's minimizers in *out. */
int mm2_sketch_batch(mm2_ctx_t* ctx, const uint8_t* cat, const uint64_t* offs, size_t nseq, int w, int k,
                     uint32_t rid_base, uint32_t rid_step, int is_hpc, mm2_mini_t** out, uint64_t** out_offs);

/* ---- index (index.rs) ------------------------------------------------------------------------------------ */
/* index.rs:427 build_index_from_fasta */
int mm2_index_build_fasta(mm2_ctx_t* ctx, const char* path, int w, int k, int b, int flag, mm2_index_t** out);
/* same, from records already in host memory (names[i] may be NULL) */
int mm2_index_build_seqs(mm2_ctx_t* ctx, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                         size_t nseq, int w, int k, int b, int flag, mm2_index_t** out);
int mm2_index_save_mmi(const mm2_index_t* idx, const char* path);               /* index.rs:233 save_to_mmi */
/* same layout, the hash entries of every bucket in the slot order of C minimap2's khash table (what `minimap2 -d` dumps)
 * instead of ascending key order (index.rs:279-287 writes std HashMap iteration order, which is random per process) */
int mm2_index_save_mmi_khash(const mm2_index_t* idx, const char* path);
/* index.rs:361 load_from_mmi.  Entries may come in any order; an index flagged MM_I_NO_SEQ (2) has no sequence array */
int mm2_index_load_mmi(mm2_ctx_t* ctx, const char* path, mm2_index_t** out);
int mm2_index_save_native(const mm2_index_t* idx, const char* path);            /* index.rs:156 save_to_file */
int mm2_index_load_native(mm2_ctx_t* ctx, const char* path, mm2_index_t** out); /* index.rs:309 load_from_file */
/* main.rs:135-145 load_index_auto: ".mmi" suffix -> MMI; else native format; else build from FASTA */
int mm2_index_load_auto(mm2_ctx_t* ctx, const char* path, int w, int k, int b, int flag, mm2_index_t** out);
void mm2_index_free(mm2_index_t* idx);
/* index.rs:143 Index::get.  *kind: 0 = None, 1 = Single, 2 = Multi; *occ is an mm2_free()-able copy (n entries) */
int mm2_index_get(const mm2_index_t* idx, uint64_t minier, uint64_t** occ, size_t* n, int* kind);
/* index.rs:111 Index::stats */
int mm2_index_stats(const mm2_index_t* idx, uint64_t* n_keys, double* avg_occ, double* avg_spacing, uint64_t* total_len);
/* index.rs:124 Index::calc_mid_occ */
int mm2_index_calc_mid_occ(const mm2_index_t* idx, float frac, int32_t* out);
/* Index fields w,k,b,flag,n_seq (index.rs:33-38) */
int mm2_index_params(const mm2_index_t* idx, int32_t* w, int32_t* k, int32_t* b, int32_t* flag, uint32_t* n_seq);
/* IndexSeq name/len (index.rs:28-29); name points into the index (valid until mm2_index_free), "*" if absent */
int mm2_index_seq(const mm2_index_t* idx, uint32_t rid, const char** name, uint32_t* len);
/* index.rs:53 get_ref_subseq: ASCII bases [st,en) of sequence rid decoded from the 4-bit S array */
int mm2_index_get_ref_subseq(const mm2_index_t* idx, uint32_t rid, int32_t st, int32_t en, uint8_t** out, size_t* n);
/* device-time breakdown of the build that produced idx (ms): sketch, sort, bucket build, pack, total; genome bases */
int mm2_index_build_timings(const mm2_index_t* idx, float* ms5, uint64_t* n_bases, uint64_t* n_minimizers);

/* ---- multi-GPU index build (index.rs:427-475 split over the GPUs of one box) ------------------------------
 * One rank per GPU.  The genome is sketched in shares (split INSIDE sequences, by tiles with a 2w+k halo), the minimizers
 * are sorted locally and routed to the owner of their hash-prefix bucket (all-to-all over NCCL / NVLink), every rank builds
 * the buckets it owns (index.rs:74-109), and the finished ranges are all-gathered so that every GPU ends up with the whole
 * index (mapping shards the READS, not the index).  The result is byte-identical to the single-GPU / CPU build.
 * The communicator lives in the library: NCCL is bound at run time (dlopen), the 128-byte unique id travels by whatever
 * the caller has (torch.distributed, MPI, a file). */
typedef struct mm2_comm mm2_comm_t;
#define MM2_COMM_ID_BYTES 128
int mm2_comm_get_unique_id(void* id128);                 /* rank 0: ncclGetUniqueId */
int mm2_comm_create(mm2_ctx_t* ctx, const void* id128, int nranks, int rank, mm2_comm_t** out);   /* collective */
void mm2_comm_destroy(mm2_comm_t* comm);
int mm2_comm_rank(const mm2_comm_t* comm, int* rank, int* nranks);
int mm2_comm_barrier(mm2_comm_t* comm);                  /* device work of every rank's context is complete on return */
/* collective: every rank passes the same host arrays (as mm2_index_build_seqs) and gets its replica of the index */
int mm2_index_build_sharded(mm2_ctx_t* ctx, mm2_comm_t* comm, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                            size_t nseq, int w, int k, int b, int flag, mm2_index_t** out);
/* the same phases for `nranks` virtual ranks on ONE GPU (the exchange steps are device copies): single-GPU test of the path */
int mm2_index_build_sharded_emulated(mm2_ctx_t* ctx, int nranks, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                                     size_t nseq, int w, int k, int b, int flag, mm2_index_t** out);
/* one process, several GPUs: ctxs[i] on distinct devices; a private host thread per context runs the sharded build over a
 * communicator of its own; out[i] = the replica on ctxs[i]'s device (`mm2rs index --gpus N`) */
int mm2_index_build_multi(mm2_ctx_t* const* ctxs, int nctx, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                          size_t nseq, int w, int k, int b, int flag, mm2_index_t** out);

/* ---- seeds (seeds.rs) -------------------------------------------------------------------------------------- */
/* seeds.rs:13 filter_query_minimizers: in place, *n updated */
int mm2_filter_query_minimizers(mm2_ctx_t* ctx, mm2_mini_t* mv, size_t* n, int32_t q_occ_max, float q_occ_frac);
/* seeds.rs:42 build_anchors_filtered (seeds.rs:38 build_anchors = mid_occ INT32_MAX): sorted by (x,y) */
int mm2_build_anchors_filtered(mm2_ctx_t* ctx, const mm2_index_t* idx, const mm2_mini_t* mv, size_t n, int32_t qlen,
                               int32_t mid_occ, mm2_anchor_t** out, size_t* n_out);

/* ---- chaining (lchain.rs) ---------------------------------------------------------------------------------- */
typedef struct {
  size_t n_chains;
  uint64_t* chain_offs; /* n_chains+1 */
  uint64_t* chain_idx;  /* anchor indices, chain after chain (lchain.rs:59 returns Vec<Vec<usize>>) */
  int32_t* scores;      /* n_chains */
  /* forward-DP trace (lchain.rs:67-92), n_anchors each; for stage-level parity checks */
  int32_t* f; int32_t* v; int64_t* pprev;
} mm2_chains_t;
/* lchain.rs:59 chain_dp_all (lchain.rs:54 chain_dp = first chain) */
int mm2_chain_dp_all(mm2_ctx_t* ctx, const mm2_anchor_t* a, size_t n, const mm2_chain_params_t* p, mm2_chains_t* out);
void mm2_chains_free(mm2_chains_t* c);
/* main.rs:105 default_chain_params */
void mm2_default_chain_params(int32_t k, mm2_chain_params_t* p);

/* ---- batched mapping: main.rs:193-219 for every read of a batch ---------------------------------------------- */
typedef struct {            /* the `align` flag surface (main.rs:55-89) after preset resolution */
  int32_t w, k;             /* -w/-k or preset (main.rs:125-133) */
  float frac_top_repetitive;/* -f */
  int32_t max_gap;          /* -g */
  int32_t bw, bw_long;      /* -r NUM[,NUM]; < 0 = not given */
  int32_t min_cnt;          /* -n */
  int32_t min_chain_score;  /* -m */
  float mask_level, pri_ratio; /* -M, -p */
  int32_t best_n;           /* -N */
  int32_t q_occ_max;        /* main.rs:195 hard-codes 10 */
  float q_occ_frac;         /* main.rs:195 hard-codes 0.01 */
  int32_t mid_occ_floor;    /* main.rs:197 hard-codes 10 */
  int32_t want_stage_dump;  /* != 0: also return minimizers / anchors / DP trace of every read (parity harness) */
} mm2_map_opts_t;
void mm2_default_map_opts(mm2_map_opts_t* o);

/* paf.rs:4-24 `PafRecord` as POD (names resolved through qname table / mm2_index_seq) */
typedef struct {
  uint32_t read_id;         /* index of the read in the batch */
  uint32_t rid;             /* target sequence */
  uint32_t qlen, qstart, qend; /* qstart/qend as stored in PafRecord (flipped only at print time, paf.rs:225-227) */
  uint32_t tlen, tstart, tend;
  uint32_t nm, blen;
  uint32_t cm, s1, s2;
  float dv;
  uint32_t rl;
  uint8_t strand;           /* '+' or '-' */
  uint8_t mapq;             /* 60 (paf.rs:213) */
  uint8_t tp;               /* 'P' or 'S' */
  uint8_t flags;            /* bit0: rescue_long_join reran the DP (lchain.rs:321-330) */
} mm2_paf_rec_t;

typedef struct {
  size_t n_recs;
  mm2_paf_rec_t* recs;      /* in read order; reads with no anchors produce no record (main.rs:211-213) */
  size_t n_panic;           /* reads on which the reference would panic (F5); they produce no record */
  uint32_t* panic_reads;
  /* work counters of this call */
  uint64_t n_reads, n_bases, n_minimizers, n_minimizers_kept, n_anchors, n_rescued;
  /* stage dump (want_stage_dump): per-read offsets (n_reads+1) into the flat arrays */
  uint64_t* mini_offs; mm2_mini_t* minis; uint8_t* mini_keep;
  uint64_t* anchor_offs; mm2_anchor_t* anchors;
  int32_t* f; int32_t* v; int32_t* pprev; /* of the DP that produced the reported chain (rescue rerun if it ran) */
} mm2_map_result_t;

/* host buffers: cat/offs as in mm2_sketch_batch.  H2D of the reads and D2H of the records happen inside. */
int mm2_map_batch(mm2_ctx_t* ctx, const mm2_index_t* idx, const uint8_t* cat, const uint64_t* offs, size_t nreads,
                  const mm2_map_opts_t* opts, mm2_map_result_t* out);
/* same with the reads already resident in HBM (d_cat/d_offs are device pointers on ctx's device; h_offs is the
 * host copy of the offsets).  Only the records travel D2H. */
int mm2_map_batch_device(mm2_ctx_t* ctx, const mm2_index_t* idx, const void* d_cat, const void* d_offs,
                         const uint64_t* h_offs, size_t nreads, const mm2_map_opts_t* opts, mm2_map_result_t* out);
/* The same with the reads as 2-bit codes: a quarter of the H2D bytes (the link, not the kernels, bounds mm2_map_batch from
 * host memory, and eight GPUs behind one host share it).  packed: base i of the concatenation in bits 2*(i%4) of byte
 * i/4, codes A0 C1 G2 T3 (nt4.rs:2-10); n_pos: ascending positions of the bases that are not ACGTacgt (nt4 code 4; they
 * are restored as 'N', which sketch.rs cannot tell from the original letter).  offs as in mm2_map_batch, in bases;
 * offs[0] must be a multiple of 16.  Results are identical to mm2_map_batch on the ASCII reads. */
int mm2_map_batch_packed(mm2_ctx_t* ctx, const mm2_index_t* idx, const uint8_t* packed, const uint64_t* n_pos, size_t n_n,
                         const uint64_t* offs, size_t nreads, const mm2_map_opts_t* opts, mm2_map_result_t* out);
/* host helper: ASCII -> the two arrays above.  packed holds (n_bases + 3) / 4 bytes (round the allocation up to a multiple
 * of 4 bytes plus 64: the upload moves whole 16-base words); *n_n is the number of non-ACGT bases, of which the first
 * n_cap positions were written. */
int mm2_pack_reads(const uint8_t* cat, uint64_t n_bases, uint8_t* packed, uint64_t* n_pos, size_t n_cap, size_t* n_n);
void mm2_map_result_free(mm2_map_result_t* r);

/* paf.rs:224 write_paf: formats one record; returns the number of bytes written (no NUL counted), < 0 on error */
int mm2_paf_format(const mm2_paf_rec_t* rec, const char* qname, const char* tname, char* buf, size_t cap);
/* paf.rs:238 write_paf_many_with_scores over a whole result: '\n'-terminated lines in one mm2_free()-able buffer.
 * qnames: nreads C strings (NULL -> "*") */
int mm2_paf_format_batch(const mm2_index_t* idx, const mm2_map_result_t* res, const char* const* qnames, char** out,
                         size_t* out_len);

#ifdef __cplusplus
}
#endif
#endif /* MM2B200_H */
