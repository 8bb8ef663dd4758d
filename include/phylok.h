/* phylok.h -- C ABI of the B200-native phyloK2 tree-kernel path (libphylok.so).
 *
 * The reference (ArtPoon/kamphir) has no FFI for this path: the boundary is the Python class
 * `PhyloKernel` (phyloK2.py:9-236) that `Kamphir` inherits (kamphir.py:34,54).  Every entry point below
 * names the reference interface it replaces (paths relative to /root/reference).  The Python mirror of the
 * class (kamphir_b200/phylokernel.py) binds these with ctypes; INTEGRATION.md shows the stub a kamphir
 * maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns a pk_status (0 = ok, < 0 = error) and
 * leaves a thread-local message retrievable with pk_last_error().  Input buffers are borrowed for the call,
 * output buffers are caller-owned, opaque handles are library-owned until the matching *_free/_destroy.
 * There is no CPU fallback: device entry points fail with PK_ERR_NODEVICE/PK_ERR_CUDA when no sm_100 GPU is usable.
 */
#ifndef PHYLOK_H
#define PHYLOK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 1

typedef enum pk_status {
    PK_OK = 0,
    PK_ERR_INVALID = -1,   /* bad argument or tree the kernel is not defined for (non-binary, NaN length, empty) -> ValueError */
    PK_ERR_PARSE = -2,     /* Newick syntax error                                                                 -> ValueError */
    PK_ERR_CUDA = -3,      /* CUDA runtime error                                                                  -> RuntimeError */
    PK_ERR_NOMEM = -4,
    PK_ERR_NODEVICE = -5   /* no usable GPU: the product path never falls back to the CPU                         -> RuntimeError */
} pk_status;

/* normalize_tree modes, phyloK2.py:101-130 ('none' = skipped, kamphir.py:616) */
enum { PK_NORM_NONE = 0, PK_NORM_MEAN = 1, PK_NORM_MEDIAN = 2 };

/* tree preparation flags: the caller-side prep sequence of kamphir.py:153-156 / :279-282 */
enum {
    PK_PREP_ZERO_ROOT = 1,  /* tree.root.branch_length = 0          kamphir.py:153,279 */
    PK_PREP_LADDERIZE = 2,  /* tree.ladderize()  (ascending #tips, stable)  kamphir.py:154,280; phyloK2.py:77-78 */
    PK_PREP_KAMPHIR = 3
};

typedef struct pk_tree pk_tree;       /* host: one flattened, annotated tree (annotate_tree, phyloK2.py:132-146) */
typedef struct pk_ctx pk_ctx;         /* one GPU + stream + scratch pool; create lazily, never before fork() (kamphir.py:638) */
typedef struct pk_forest pk_forest;   /* device-resident packed SoA of a set of trees */

/* kernel parameters = PhyloKernel.__init__ keywords, phyloK2.py:10-22 */
typedef struct pk_params {
    double decay;       /* decayFactor, lambda   (kamphir -kdecay, kamphir.py:608) */
    double gauss;       /* gaussFactor, tau      (kamphir -tau,    kamphir.py:611) */
    double sigma;       /* sigma, additive constant in prod(sigma + C(child)), phyloK2.py:194-199 */
    int32_t precision;  /* 64: fp64 end to end (1e-9 rel.)   32: fp32 cells, fp64 accumulation (1e-5 rel.) */
    int32_t reserved;
} pk_params;

typedef struct pk_tree_info {
    int32_t n_internal;   /* len(get_nonterminals()) */
    int32_t n_tips;       /* len(get_terminals()) */
    int32_t n_levels;     /* wavefront depth: 1 + max height of an internal node above the cherries */
    int32_t n_classes;    /* production classes: 3 unlabelled, 1 + 2S + S*S with S tip states */
    int32_t n_states;     /* 0 = unlabelled */
    int32_t reserved;
    double height;        /* max root-to-tip depth before normalisation (kamphir.py:118-119,255-256) */
    double scale;         /* divisor applied by normalize_tree (1 when PK_NORM_NONE) */
} pk_tree_info;

/* ---- library ------------------------------------------------------------------------------------- */
int pk_abi_version(void);
const char *pk_last_error(void);
int pk_device_count(void);                       /* CUDA devices visible, 0 when none / no driver */

/* ---- host side: Newick -> flattened tree (replaces Bio.Phylo parse + kamphir prep + annotate_tree) ----
 * pk_tree_from_newick : Phylo.read (kamphir.py:323) + prep (kamphir.py:279-282) + annotate (phyloK2.py:132-146)
 * pk_trees_from_newick: Phylo.parse over a multi-tree handle (kamphir.py:112, phyloK2.py:74); returns an array
 *                       allocated by the library (free each tree, then pk_tree_array_free).
 * n_states = S > 0 turns on tip-state productions: state = integer after the last '_' of the tip name, 0 .. S-1 or
 *                       rcolgem's 1 .. S (rcolgem.py:274,376,588; S counts as 0); extension with no reference
 *                       implementation.
 * nthreads: worker threads for the multi-tree form (<= 0: hardware concurrency). */
int pk_tree_from_newick(const char *text, size_t len, int prep_flags, int normalize, int n_states, pk_tree **out);
int pk_trees_from_newick(const char *text, size_t len, int prep_flags, int normalize, int n_states, int nthreads,
                         pk_tree ***out, int32_t *n_out);
/* same, for n separate one-tree strings (what a simulator returns, kamphir.py:318-326): no concatenation, no split pass */
int pk_trees_from_newick_list(const char *const *texts, const size_t *lens, int32_t n, int prep_flags, int normalize,
                              int n_states, int nthreads, pk_tree ***out, int32_t *n_out);
void pk_tree_array_free(pk_tree **arr);

/* pk_tree_from_arrays: a tree already walked by the caller (the Python mirror walks Bio.Phylo-like objects in
 * get_nonterminals(order='postorder') order, phyloK2.py:166-167).  cidx[2*i+s] = post-order index of the internal
 * child in slot s or -1 for a tip; bl[2*i+s] = that child's branch length (node.bl, phyloK2.py:144-145);
 * cls = NULL (production = #tip children + 1, phyloK2.py:141-142) or explicit class ids in [0, n_classes). */
int pk_tree_from_arrays(int32_t n_internal, const int32_t *cidx, const double *bl, const uint8_t *cls,
                        int32_t n_classes, pk_tree **out);
void pk_tree_free(pk_tree *t);
/* node heights (any order), un-normalised tree height and normalisation divisor of a tree built with pk_tree_from_arrays
 * (what kcoal, kamphir.py:255-269, needs and the arrays do not carry); the Python mirror uses it to rebuild pickled trees */
int pk_tree_set_heights(pk_tree *t, const double *heights, int32_t n, double height, double scale);
int pk_tree_get_info(const pk_tree *t, pk_tree_info *info);
/* post-order arrays as annotate_tree defines them (for parity tests): prod[n], cprod[2n], cidx[2n], bl[2n] */
int pk_tree_export(const pk_tree *t, uint8_t *prod, uint8_t *cprod, int32_t *cidx, double *bl);
/* ascending heights of the internal nodes above the deepest tip, un-normalised (kamphir.py:255-258,147-149) */
int pk_tree_node_heights(const pk_tree *t, double *heights /* n_internal */);
/* exact work of one DP: counts = {node pairs, matched cells M, child-cell reads R} */
int pk_pair_counts(const pk_tree *a, const pk_tree *b, int64_t counts[3]);
/* kcoal of kamphir.py:260-269: cosine of the two ascending node-height vectors (sim indexed into target).
 * PK_ERR_INVALID where the reference raises: more internal nodes than the target (IndexError, :264), zero norm
 * (ZeroDivisionError, :269). */
int pk_kcoal(const pk_tree *sim, const pk_tree *target, double *out);
/* the same for every replicate of one proposal (kamphir.py:434-450 loops over them): out[s] = kcoal(sims[s], target) */
int pk_kcoal_batch(const pk_tree *const *sims, int32_t nsims, const pk_tree *target, double *out);

/* ---- device side ------------------------------------------------------------------------------------ */
int pk_ctx_create(int device, pk_ctx **out);
void pk_ctx_destroy(pk_ctx *ctx);
int pk_ctx_set_stream(pk_ctx *ctx, void *cuda_stream);   /* run on a caller-owned cudaStream_t (NULL = own stream) */
int pk_ctx_sync(pk_ctx *ctx);
/* tuning knobs (0 = default): threads per CTA (multiple of 32, >= widest class), CTAs per SM for the HBM-slab path,
 * C storage: 0 = automatic (shared memory only when it costs no resident CTAs), 1 = HBM slabs, 2 = shared memory */
int pk_ctx_configure(pk_ctx *ctx, int32_t threads, int32_t ctas_per_sm, int32_t force_hbm);
/* named test / tuning knobs: "unstaged" = 1 stages both tree parts in global memory even when they fit the shared memory
 * of an SM (the path trees beyond ~3 500 tips take by themselves), 2 = COLUMN part in global memory wherever possible (the
 * path 1 600-3 000-tip trees take by themselves: two CTAs per SM instead of one), 3 = never, 0 = automatic */
int pk_ctx_set_option(pk_ctx *ctx, const char *name, int32_t value);
/* the most recent DP kernel launch on this ctx: ms = its device time (CUDA events recorded around that launch on the
 * ctx stream; blocks until it finished), launches = DP kernels launched since ctx creation, path = 0 (C in shared
 * memory) / 1 (C in HBM slabs) / 2 (HBM slabs, tree parts staged in global memory: trees beyond the shared-memory limit) / 3 (HBM slabs, ROW part
 * in shared memory, COLUMN part in global memory), launch_cfg = {grid, threads per CTA, dynamic shared memory bytes}. Any may be NULL. */
int pk_ctx_last_stats(pk_ctx *ctx, double *ms, int64_t *launches, int32_t *path, int32_t launch_cfg[3]);

/* Results that came out non-finite since the last reset, counted by the DP kernel itself.  In fp32 mode a cell beyond
 * 3e38 overflows (decay factors above ~0.25 on deep, similar trees); the host-resident entry points (pk_kernel_pairs,
 * pk_score_batch, pk_score_newick_batch, pk_gram) re-evaluate such pairs in fp64 themselves, callers of the
 * device-resident forms check this counter (GramRunner.fetch raises).  Blocks until the stream is idle. */
int pk_ctx_nonfinite(pk_ctx *ctx, int64_t *count, int reset);

int pk_forest_upload(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, int32_t precision, pk_forest **out);
/* Multi-GPU replication of a forest without packing it W times: every rank lays the whole forest out (same offsets,
 * same meta records, same host-side statistics) but packs and uploads only the trees [part_lo, part_hi); the ranks then
 * exchange their byte ranges device-to-device over NVLink (pk_forest_device_range gives the pointer and size of a tree
 * range; kamphir_b200/gram.py broadcasts them with NCCL).  The forest must not be used before the exchange is done. */
int pk_forest_upload_part(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, int32_t precision, int32_t part_lo,
                          int32_t part_hi, pk_forest **out);
int pk_forest_device_range(const pk_forest *f, int32_t tree_lo, int32_t tree_hi, void **d_ptr, int64_t *bytes);
void pk_forest_free(pk_forest *f);
int pk_forest_bytes(const pk_forest *f, int64_t *bytes);
/* exact work of a pair list / of one Gram part: counts = {DPs, node pairs, matched cells M, child reads R} */
int pk_forest_pair_stats(const pk_forest *fa, const pk_forest *fb, const int32_t *pairs, int64_t npairs, int64_t counts[4]);
int pk_forest_gram_stats(const pk_forest *f, int32_t tile, int32_t part, int32_t nparts, int64_t counts[4]);

/* PhyloKernel.kernel (phyloK2.py:158-209) for a list of pairs: out[p] = K(A[pairs[2p]], B[pairs[2p+1]]).
 * Host-resident form: uploads, launches, copies back (this is what PhyloKernel.kernel/kernel_batch call). */
int pk_kernel_pairs(pk_ctx *ctx, const pk_tree *const *A, int32_t nA, const pk_tree *const *B, int32_t nB,
                    const int32_t *pairs, int32_t npairs, const pk_params *params, double *out_k);
/* Device-resident form: forests already in HBM; pairs and out are DEVICE pointers; asynchronous on the ctx stream. */
int pk_forest_kernel_pairs(pk_ctx *ctx, const pk_forest *fa, const pk_forest *fb, const int32_t *d_pairs,
                           int32_t npairs, const pk_params *params, double *d_out_k);

/* Device-resident forests, HOST pair list and HOST results: one launch, one synchronisation (what PhyloKernel.kernel
 * uses for trees whose single-tree forests it keeps on the device between calls). */
int pk_forest_kernel_pairs_host(pk_ctx *ctx, const pk_forest *fa, const pk_forest *fb, const int32_t *pairs,
                                int32_t npairs, const pk_params *params, double *out_k);

/* The ABC objective's inner loop, kamphir.py:288-306 + :434-450, for one target tree and nsims replicates:
 *   k[s] = K(obs, sim_s), denom[s] = K(sim_s, sim_s), khat[s] = exp(log k - (log ref_denom + log denom)/2).
 * *ref_denom = K(obs,obs) (kamphir.py:157) is computed when passed as NaN or <= 0 and returned. Any output may be NULL.
 * Where the reference's math.log(0) raises (K = 0: decayFactor 0, or labelled trees without a common production class)
 * khat[s] is NaN or 0 and the call still succeeds; the Python mirror turns k <= 0 into the reference's ValueError. */
int pk_score_batch(pk_ctx *ctx, const pk_tree *obs, const pk_tree *const *sims, int32_t nsims, const pk_params *params,
                   double *ref_denom, double *out_k, double *out_denom, double *out_khat);

/* One ABC proposal in one call, from the simulators' Newick strings (kamphir.py:318 hands them over as text) to the
 * scores of kamphir.py:288-306 and their mean (:450): parse + prep (prep_flags / normalize / n_states as in
 * pk_trees_from_newick_list) on the worker pool, one DP launch (pk_score_batch), kcoal (unless kadj is NaN) and
 * score[s] = khat[s] * kcoal[s]^kadj.  obs is the target tree, flattened once per run.  Any output may be NULL. */
int pk_score_newick_batch(pk_ctx *ctx, const pk_tree *obs, const char *const *sims, const size_t *lens, int32_t nsims,
                          int prep_flags, int normalize, int n_states, int nthreads, const pk_params *params,
                          double *ref_denom, double kadj, double *out_k, double *out_denom, double *out_khat,
                          double *out_kcoal, double *out_score, double *out_mean);

/* FORWARD NOTE on the tile partition: pk_gram_part_pairs / pk_gram_part_tiles below describe the uniform partition (tile t
 * to part t % nparts); a forest-aware call (pk_forest_gram_part, pk_forest_gram_stats, pk_forest_gram_part_tiles) deals the
 * tiles by cost = node pairs of the tile, which for equal-sized trees differs only in where the half-filled
 * diagonal tiles go.  Either way the parts are disjoint and cover the upper triangle. */

/* All-pairs Gram matrix, the working form of PhyloKernel.compute_matrix (phyloK2.py:148-156): out is N x N row-major,
 * symmetric; normalised = 0: raw K (reference semantics); 1: K / sqrt(K_ii K_jj).  Host-resident form. */
int pk_gram(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, const pk_params *params, int normalised, double *out);
/* Device-resident, shardable form.  The upper triangle is cut into tile x tile blocks, dealt to the parts by cost
 * (see the note above); this call computes the blocks of `part` and writes both (i,j) and (j,i) into d_out (N x N, leading
 * dimension ld, DEVICE pointer -- may be a peer/IPC mapping of another GPU's buffer so results cross NVLink once,
 * written by the DP kernel itself).  d_diag (N doubles, device) receives K_ii; it is computed by every part.
 * Asynchronous on the ctx stream. */
int pk_forest_gram_part(pk_ctx *ctx, const pk_forest *f, const pk_params *params, int normalised, int32_t tile,
                        int32_t part, int32_t nparts, double *d_out, int64_t ld, double *d_diag);
/* DPs that pk_forest_gram_part(part, nparts) evaluates (excluding the N diagonal DPs every part repeats) */
int pk_gram_part_pairs(int32_t n, int32_t tile, int32_t part, int32_t nparts, int64_t *npairs);
/* the (row-block, column-block) coordinates of that part's tiles; pass tiles = NULL to query the count (host only) */
int pk_gram_part_tiles(int32_t n, int32_t tile, int32_t part, int32_t nparts, int32_t *tiles, int64_t *ntiles);

/* the tiles pk_forest_gram_part(part, nparts) evaluates for THIS forest (cost-weighted deal); tiles = NULL queries the count */
int pk_forest_gram_part_tiles(const pk_forest *f, int32_t tile, int32_t part, int32_t nparts, int32_t *tiles, int64_t *ntiles);

/* cross-process sharing of a device buffer (one result matrix on rank 0, written by all ranks over NVLink) */
/* page-locked host buffers for results fetched every step (GramRunner.fetch) */
int pk_host_alloc(pk_ctx *ctx, size_t bytes, void **h_ptr);
int pk_host_free(pk_ctx *ctx, void *h_ptr);
int pk_device_alloc(pk_ctx *ctx, size_t bytes, void **d_ptr);
int pk_device_free(pk_ctx *ctx, void *d_ptr);
int pk_ipc_export(pk_ctx *ctx, void *d_ptr, unsigned char handle[64]);
int pk_ipc_open(pk_ctx *ctx, const unsigned char handle[64], void **d_ptr);
int pk_ipc_close(pk_ctx *ctx, void *d_ptr);
int pk_memcpy_d2h(pk_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int pk_memcpy_h2d(pk_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int pk_memset_d(pk_ctx *ctx, void *d_dst, int value, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PHYLOK_H */
