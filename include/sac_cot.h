/*
 * sac_cot.h — C ABI of the SAC-COT registration hot path.
 *
 * Boundary provenance. The reference repository (ytuhzq/SAC-COT) ships no code and no
 * FFI surface: /root/reference/README.md:1-2 is the whole repo.  The boundary declared
 * here is therefore the one BASELINE.json `north_star` dictates —
 *     sac_cot_register(src, dst, N, params, &R, &t, &inliers)
 * — fleshed out as SURVEY.md §8b describes.  Every entry point below "replaces" the
 * (absent) reference interface README.md:2 promises ("method code for the paper").
 *
 * The same header is implemented twice:
 *   - sac_cot_b200/lib/libsaccot.so    the B200 product (hand-written sm_100a CUDA, no CPU path)
 *   - oracle/libsaccot_oracle.so       the from-paper CPU oracle (test infrastructure only)
 * so one ctypes harness drives both and parity tests read identically for both.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the ABI; every function returns a status
 *     (0 = OK, <0 = invalid argument / unsupported, >0 = CUDA runtime failure).
 *   - points are fp32, row-major N x 3 (x,y,z); correspondence n is (src[n], dst[n]).
 *   - R is row-major 3x3 with dst ~= R*src + t.
 *   - the caller owns every in/out buffer; nothing is retained after return.
 *   - a ctx is single-threaded (one caller at a time); distinct ctxs are independent.
 *   - "no valid triangle" (graph has no 3-clique among the selected edges) is a *result*:
 *     status 0, R = I, t = 0, inliers = 0.
 */
#ifndef SAC_COT_H_
#define SAC_COT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SAC_COT_API
#else
#define SAC_COT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ limits */
#define SAC_COT_MAX_N            65535   /* node ids and per-edge counts are packed in 16 bit */
#define SAC_COT_MAX_EDGES        4096    /* K_e  */
#define SAC_COT_MAX_APEX         8       /* m    */
#define SAC_COT_MAX_HYPOTHESES   32768   /* K_e*m; hypothesis ids are packed in 16 bit      */
#define SAC_COT_MAX_DESC_DIM     256     /* descriptor width of the matching front end       */
#define SAC_COT_MAX_KEYPOINTS    1048576 /* keypoints per cloud of the matching front end    */

/* ------------------------------------------------------------------ status */
enum {
  SAC_COT_OK             = 0,
  SAC_COT_E_NULL         = -1,  /* a required pointer is NULL                              */
  SAC_COT_E_SIZE         = -2,  /* N < 3, N > SAC_COT_MAX_N, B < 0                         */
  SAC_COT_E_PARAMS       = -3,  /* struct_size mismatch or a field out of range            */
  SAC_COT_E_NODEVICE     = -4,  /* GPU library only: no usable CUDA device / ext. missing  */
  SAC_COT_E_UNSUPPORTED  = -5,  /* entry point not provided by this implementation         */
  SAC_COT_E_WHICH        = -6,  /* debug_get: unknown selector or pair index               */
  SAC_COT_E_CAPACITY     = -7,  /* debug_get: caller buffer too small (written = needed)   */
  SAC_COT_E_NOMEM        = -8,  /* workspace could not be allocated                        */
  SAC_COT_E_COMM         = -9   /* GPU library only: no communicator on the ctx, NCCL not
                                   loadable, or an NCCL call failed                        */
  /* > 0 : cudaError_t of the failing CUDA call (GPU library only)                          */
};

/* ------------------------------------------------------------------ params */
enum { SAC_COT_SCORE_INLIER_COUNT = 0, SAC_COT_SCORE_TRUNCATED_RESIDUAL = 1 };
/* Compatibility graph the ranking runs on:
 *   FIRST_ORDER   A_ij  = | |s_i - s_j| - |d_i - d_j| | < tau_c                                  (S1, the default)
 *   SECOND_ORDER  A2_ij = A_ij and C_ij >= so_min_common, where C_ij = popc(row_i(A) & row_j(A)) is the number of
 *                 correspondences compatible with BOTH i and j — the second-order (SC^2-style) measure, (A.A) o A.
 *                 Triangle counts, edge ranking, apex selection (S2, S3) then run on A2 instead of A; hypotheses,
 *                 scoring and refit (S4-S7) are unchanged.  so_min_common = 0 makes A2 = A. */
enum { SAC_COT_COMPAT_FIRST_ORDER = 0, SAC_COT_COMPAT_SECOND_ORDER = 1 };

typedef struct sac_cot_params {
  uint32_t struct_size;    /* = sizeof(sac_cot_params); ABI version check (see below)                */
  float    tau_compat;     /* tau_c : |len_src - len_dst| < tau_c  => compatible  (S1)        */
  float    tau_inlier;     /* tau_in: |R s + t - d|^2 < tau_in^2   => inlier      (S5)        */
  int32_t  num_edges;      /* K_e   : top-ranked edges used as triangle bases     (S3)        */
  int32_t  apex_per_edge;  /* m     : apexes per edge; K = K_e*m hypotheses       (S3)        */
  int32_t  score_mode;     /* SAC_COT_SCORE_*                                     (S5)        */
  int32_t  refit;          /* 1: fp64 Kabsch over the winner's inliers (S7); 0: return winner */
  int32_t  compat_mode;    /* SAC_COT_COMPAT_*  (version 1 of the struct called this `reserved`, must be 0) */
  /* ---- version 2 (struct_size 40) ---- */
  int32_t  so_min_common;  /* SECOND_ORDER: smallest C_ij an edge of A must reach to stay in A2, 0 .. 65535 */
  int32_t  reserved;       /* must be 0                                                       */
} sac_cot_params;
/* struct_size versioning: both libraries accept 32 (version 1: the fields up to compat_mode, which must then be 0)
 * and 40 (version 2).  A caller compiled against version 1 keeps working unchanged. */
#define SAC_COT_PARAMS_SIZE_V1 32u

/* Fills *p with the defaults (tau 0.1/0.1, K_e 1024, m 4, inlier count, refit on, first-order graph). */
SAC_COT_API int sac_cot_params_default(sac_cot_params* p);

/* ------------------------------------------------------------------ one pair */
/* Uses a lazily created process-global ctx (GPU library: device 0, its own stream). */
SAC_COT_API int sac_cot_register(const float* src, const float* dst, int32_t N,
                                 const sac_cot_params* params,
                                 float R[9], float t[3], int32_t* inliers);

/* ------------------------------------------------------------------ contexts */
typedef struct sac_cot_ctx sac_cot_ctx;

/* device: CUDA ordinal (ignored by the oracle).  stream: a cudaStream_t on that device on
 * which every kernel and copy of this ctx is enqueued, or NULL for a private non-blocking
 * stream.  Passing the caller's stream lets the caller bracket the work with its own events;
 * to name the legacy default stream (handle 0, e.g. torch's default stream) pass
 * cudaStreamLegacy ((void*)0x1), since NULL means "private". */
SAC_COT_API int sac_cot_ctx_create(sac_cot_ctx** out, int32_t device, void* stream);
SAC_COT_API int sac_cot_ctx_destroy(sac_cot_ctx* ctx);

/* Tunables/inspection by name; unknown names return SAC_COT_E_WHICH.
 *   set: "keep_debug" (0/1: retain per-pair intermediates for sac_cot_debug_get; forces
 *        chunk = whole batch), "chunk_pairs" (pairs per kernel wave, 0 = auto), "lanes" (1..4
 *        internal streams the chunks of a batch are dealt to, default 3; GPU only),
 *        "triangle_path" (0 = POPC bitset kernels, 1 = tensor-core dense kernel, 2 = chosen per chunk
 *        from the measured edge density (default); GPU only),
 *        "triangle_prune" (tensor-core path, default 1: keep only edge keys whose count reaches
 *        a per-pair threshold proven to lie at or below the K_e-th largest count; results are
 *        unchanged, SAC_COT_DBG_EDGE_KEYS / _HIST then cover the kept edges only; GPU only),
 *        "node_prune" (tensor-core path, default 1: a pair whose selectable edges provably join few
 *        high-degree nodes — degree >= that threshold + 1, e.g. an inlier clique in a sparse outlier graph —
 *        counts triangles for those nodes' rows only; results are unchanged; off while "keep_debug" is set,
 *        because SAC_COT_DBG_T_NODE then covers the kept nodes only; 2 = on whenever the kept list fits, also
 *        with keep_debug (tests); 0 = off; GPU only), "node_prune_cost" (default 200: a pair is pruned if
 *        (sum of the kept nodes' degrees) x cost <= Npad^2; GPU only), "node_prune_rect" (default 1: the kept rows of
 *        pairs with N >= 1921 and at most 1024 kept nodes run on the tensor cores when the chunk holds >= 4 pairs,
 *        2 = whatever the chunk size, 0 = always the POPC kept-row kernel; "rect_pairs" tells how many pairs of the
 *        latest chunk did; GPU only), "node_prune_probe" (default 30: a ctx
 *        whose last two calls tried and pruned nothing skips the attempt — five near-empty launches per chunk —
 *        for this many calls, then tries again; results never depend on it; GPU only),
 *        "stage_timing" (0/1: bracket every pipeline stage with CUDA events on the ctx
 *        stream; setting it also clears the accumulated times; GPU only),
 *        test switches that never change a result (GPU only): "tile_runs" (tensor-core path, default 1:
 *        tiles dealt to the CTA pairs in runs; 0 = one at a time), "apex_path" (0 = shared-memory
 *        kernel with the per-pair rank list (default), 1 = same kernel, exhaustive scan of every edge,
 *        2 = the global-lookup kernel that N > 51200 falls back to), "triangle_dbg" (experiments)
 *   get: "triangle_path_used" (0/1: which S2 kernels the latest chunk ran; synchronises; GPU only),
 *        "pruned_pairs" / "kept_nodes" (node pruning in the latest chunk: pairs that took the kept-row kernel and
 *        the nodes they kept in total; synchronises; GPU only), "node_prune_trying" (1: the next call attempts the
 *        node pruning, 0: the ctx is inside a back-off; GPU only),
 *        "launches" (kernels launched since ctx creation), "workspace_bytes",
 *        "device", "sm_count", "retries" (workspace-growth re-runs),
 *        "last_status" (deferred status of the SAC_COT_LOC_DEVICE calls since the previous query: SAC_COT_E_NOMEM
 *        if a chunk of any of them ran out of key-pool space — its outputs read "no result" and the pool has been
 *        grown for the next call; synchronises),
 *        "probe_mxf4_gflops" (measures, now, the dense rate of the tensor-core triangle kernel's MMA shape with
 *        nothing else running: the roofline denominator bench.py reports against; synchronises; GPU only),
 *        "stage_us_<s>" / "stage_calls_<s>" with <s> in pack, graph, scan, theta, triangles, triangles_kept, select,
 *        apex, kabsch, score, finalize, exchange1 (sharded: record + all-gather + merge), exchange2
 *        (sharded: all-reduce), match_prep, match_sweep, match_exact (front end): accumulated device
 *        microseconds / launches,
 *        "comm_rank", "comm_world" (0 = no communicator)
 * Diagnostics: with SAC_COT_TRACE set in the environment the GPU library synchronises the device after every
 * kernel launcher and names, on stderr, the first one whose kernels failed.               */
SAC_COT_API int sac_cot_ctx_set(sac_cot_ctx* ctx, const char* name, int64_t value);
SAC_COT_API int sac_cot_ctx_get(sac_cot_ctx* ctx, const char* name, int64_t* value);

/* ------------------------------------------------------------------ batches */
/* B independent pairs, host pointer arrays (SURVEY.md §8b signature).  Outputs: R[B*9],
 * t[B*3], inliers[B] (host). */
SAC_COT_API int sac_cot_register_batch(sac_cot_ctx* ctx,
                                       const float* const* src, const float* const* dst,
                                       const int32_t* N, int32_t B,
                                       const sac_cot_params* params,
                                       float* R, float* t, int32_t* inliers);

/* B independent pairs packed back to back: pair b owns points [offsets[b], offsets[b+1])
 * of src/dst (each total x 3 floats).  `offsets` (B+1 entries) is always a host array.
 * location = SAC_COT_LOC_HOST: src/dst/R/t/inliers are host buffers (pinned or pageable);
 * the H2D and D2H copies are part of the call.
 * location = SAC_COT_LOC_DEVICE (GPU library only): they are device buffers on the ctx
 * device; the call only enqueues work on the ctx stream and returns without
 * synchronising (status reflects enqueue errors; results are ready when the stream is). */
enum { SAC_COT_LOC_HOST = 0, SAC_COT_LOC_DEVICE = 1 };
SAC_COT_API int sac_cot_register_packed(sac_cot_ctx* ctx,
                                        const float* src, const float* dst,
                                        const int64_t* offsets, int32_t B,
                                        const sac_cot_params* params,
                                        float* R, float* t, int32_t* inliers,
                                        int32_t location);

/* ------------------------------------------------------------------ correspondence front end */
/* Nearest-neighbour matching of local descriptors (33-D FPFH-like): the stage that produces the N
 * putative correspondences sac_cot_register* consumes.  For every source keypoint i of pair b
 *     D_ij  = sum_c (f_ic - g_jc)^2 in fp32:  D = 0; for c = 0 .. dim-1: e = f_ic - g_jc; D = fma(e, e, D)
 *     nn[i] = argmin_j D_ij over the pair's target keypoints, ties -> lowest j   (index local to the pair)
 *     corr_src[i] = xyz_src[i],  corr_dst[i] = xyz_dst[nn[i]]
 * so pair b gets Ns_b correspondences at offs_src[b]: corr_src / corr_dst / offs_src are exactly the src / dst /
 * offsets of sac_cot_register_packed (with SAC_COT_LOC_DEVICE the hand-off never leaves the device).
 * Descriptors must be finite.  desc_*: rows x dim, xyz_*: rows x 3, row-major, packed over the pairs; offs_* (B+1
 * entries) are host arrays.  location as in sac_cot_register_packed (DEVICE: enqueue only on the ctx stream).
 * GPU library: dim <= 40 runs the search on the tensor cores (bf16x3 split operands, tcgen05.mma, candidates within a
 * proven margin) and decides with the specified fp32 chain — results are bit-identical to the oracle's brute force;
 * wider descriptors are scanned exhaustively with that chain.  ctx knob "match_path" = 0 forces the exhaustive scan. */
SAC_COT_API int sac_cot_match_packed(sac_cot_ctx* ctx,
                                     const float* desc_src, const float* xyz_src, const int64_t* offs_src,
                                     const float* desc_dst, const float* xyz_dst, const int64_t* offs_dst,
                                     int32_t B, int32_t dim,
                                     int32_t* nn, float* corr_src, float* corr_dst, int32_t location);
/* Mutual-nearest-neighbour filter (optional second step).  nn is the result of matching source -> target, nn_back the
 * result of a second sac_cot_match_packed call with the two clouds swapped (target -> source); correspondence i of pair
 * b survives iff nn_back[nn[i]] == i.  The survivors are written, in order and packed over the pairs, to out_src /
 * out_dst (room for offs_src[B] x 3 floats each) and out_offsets (B + 1 entries: pair b owns [out_offsets[b],
 * out_offsets[b+1])) — again exactly the src / dst / offsets of sac_cot_register_packed, whose offsets must be a host
 * array: with SAC_COT_LOC_DEVICE out_offsets is a device array the caller copies back.  A pair may end up with fewer
 * than three correspondences: leave it out of the registration call. */
SAC_COT_API int sac_cot_match_mutual(sac_cot_ctx* ctx, const int32_t* nn, const int32_t* nn_back,
                                     const float* corr_src, const float* corr_dst,
                                     const int64_t* offs_src, const int64_t* offs_dst, int32_t B,
                                     float* out_src, float* out_dst, int64_t* out_offsets, int32_t location);
/* One pair, host buffers, process-global ctx (as sac_cot_register). */
SAC_COT_API int sac_cot_match(const float* desc_src, const float* xyz_src, int32_t Ns,
                              const float* desc_dst, const float* xyz_dst, int32_t Nd, int32_t dim,
                              int32_t* nn, float* corr_src, float* corr_dst);

/* ------------------------------------------------------------------ several GPUs of one box, batched pairs */
/* A group owns one ctx per listed CUDA device and an enqueueing thread for each.  A batch is dealt round-robin —
 * pair b to member (b mod G) — with no communication between the devices (independent pairs
 * shard naturally); each member gathers its share straight from the caller's host arrays (a strided 2-D copy when
 * the pairs are of equal size) and scatters its results into the caller's R / t / inliers.  Host buffers only
 * (pinned memory makes the copies asynchronous); the call returns when every device has finished.  Results are
 * those of sac_cot_register_packed on one device, bit for bit.  GPU library only. */
typedef struct sac_cot_group sac_cot_group;
SAC_COT_API int sac_cot_group_create(sac_cot_group** out, const int32_t* devices, int32_t n_devices);
SAC_COT_API int sac_cot_group_destroy(sac_cot_group* group);
SAC_COT_API int32_t sac_cot_group_size(const sac_cot_group* group);
/* member ctx `index` (owned by the group): for sac_cot_ctx_get / per-device knobs */
SAC_COT_API sac_cot_ctx* sac_cot_group_ctx(sac_cot_group* group, int32_t index);
/* sac_cot_ctx_set on every member */
SAC_COT_API int sac_cot_group_set(sac_cot_group* group, const char* name, int64_t value);
SAC_COT_API int sac_cot_group_register_packed(sac_cot_group* group,
                                              const float* src, const float* dst,
                                              const int64_t* offsets, int32_t B,
                                              const sac_cot_params* params,
                                              float* R, float* t, int32_t* inliers);

/* ------------------------------------------------------------------ one large pair, sharded */
/* A single pair whose triangle-count work and hypothesis ranges are split over `world` ranks, one
 * process (or thread) per GPU, every rank holding the same src/dst (BASELINE.json
 * configs[4]).  Exactly two exchanges:
 *
 *   part 1   graph (full, local) + triangle counts of the cells this rank owns
 *            -> partial node sums (u64[N]; summed over the ranks they equal 2*t_i)
 *            -> this rank's top-K_e edge keys (descending, 0-padded)
 *   exchange #1: all-gather of both
 *   part 2   merge (sum; top-K_e of the union), apexes, Kabsch, scoring of this rank's hypothesis
 *            range -> best key of the range
 *   exchange #2: all-reduce(max) of the packed (score, hypothesis id) key
 *   part 3   inlier mask + refit of the global winner; every rank returns the same (R, t, inliers),
 *            bit-identical to the unsharded call.
 *
 * "S2 partition" (normative): edge (i < j) lies in cell (cb, ic) = (j / 1920, i / 256) and cell
 * (cb, ic) belongs to rank (257*cb + ic) mod world; hypothesis h belongs to rank h / ceil(K/world).
 *
 * Two ways to run it:
 *
 * (a) sac_cot_register_sharded — the GPU library does everything, collectives included: the ctx
 *     holds an NCCL communicator (created once by sac_cot_ctx_comm_init, or adopted with
 *     sac_cot_ctx_set_comm) and ncclAllGather / ncclAllReduce(ncclUint64, ncclMax) are enqueued
 *     on the ctx's own stream between the kernels, device buffer to device buffer, with no host
 *     synchronisation between the parts.  Collective call: every rank of the communicator calls it
 *     with the same src/dst/N/params.
 *
 * (b) sac_cot_sharded_phase{1,2,3} — the three parts with host arrays in between, for callers
 *     that bring their own transport (MPI, gloo, a test harness emulating the ranks).            */

/* ---- (a) in-library collectives (GPU library only; NCCL is loaded with dlopen("libnccl.so.2")
 * on first use, so the library itself has no link-time NCCL dependency) */
#define SAC_COT_COMM_ID_BYTES 128   /* sizeof(ncclUniqueId) */
/* Rank 0 creates an id (ncclGetUniqueId) and hands the 128 bytes to every rank by any means
 * (torch.distributed broadcast, MPI_Bcast, a file); then every rank calls comm_init
 * (ncclCommInitRank on the ctx device; collective, blocks until all ranks have joined). */
SAC_COT_API int sac_cot_comm_unique_id(void* id_out /* SAC_COT_COMM_ID_BYTES */);
SAC_COT_API int sac_cot_ctx_comm_init(sac_cot_ctx* ctx, const void* id, int32_t rank, int32_t world);
/* Adopts an existing ncclComm_t whose device is the ctx device (the caller keeps ownership and
 * must keep it alive while the ctx uses it); comm = NULL detaches. */
SAC_COT_API int sac_cot_ctx_set_comm(sac_cot_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world);
/* location = SAC_COT_LOC_HOST: host buffers, the call returns when (R, t, inliers) are written; a key
 * pool that proves too small on ANY rank is grown on every rank and the pair re-run, in step.
 * location = SAC_COT_LOC_DEVICE: device buffers on the ctx device, enqueue only (status through
 * sac_cot_ctx_get("last_status"), the same on every rank).
 * Without a communicator: SAC_COT_E_COMM.  The oracle implements it for world = 1 only. */
SAC_COT_API int sac_cot_register_sharded(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N,
                                         const sac_cot_params* params,
                                         float R[9], float t[3], int32_t* inliers, int32_t location);

/* ---- (b) the parts, host arrays in and out; t_all is world x N, cand_all is world x K_e */
SAC_COT_API int sac_cot_sharded_phase1(sac_cot_ctx* ctx, const float* src, const float* dst,
                                       int32_t N, const sac_cot_params* params,
                                       int32_t rank, int32_t world,
                                       uint64_t* t_partial, uint64_t* cand);
SAC_COT_API int sac_cot_sharded_phase2(sac_cot_ctx* ctx, const uint64_t* t_all,
                                       const uint64_t* cand_all, uint64_t* best_key);
SAC_COT_API int sac_cot_sharded_phase3(sac_cot_ctx* ctx, uint64_t best_key_global,
                                       float R[9], float t[3], int32_t* inliers);

/* ------------------------------------------------------------------ parity/debug access */
/* Intermediates of pair `pair` of the most recent call on ctx (requires keep_debug = 1).
 * Copies into `out` (host) and reports the byte count in *written.                      */
enum {
  SAC_COT_DBG_ADJ        = 0, /* u32[N][stride_words]: bit j&31 of word j>>5 of row i = A_ij;
                                 stride_words = ceil(N/128)*4; pad bits 0                    */
  SAC_COT_DBG_T_NODE     = 1, /* u32[N]  t_i  = #triangles through node i                    */
  SAC_COT_DBG_NUM_EDGES  = 2, /* u64     E    = #undirected edges                            */
  SAC_COT_DBG_EDGE_KEYS  = 3, /* u64[E]  all edge keys, ORDER UNSPECIFIED (sort to compare)   */
  SAC_COT_DBG_TOP_EDGES  = 4, /* u64[K_e'] selected edge keys, descending; K_e'=min(K_e,E)   */
  SAC_COT_DBG_TRIANGLES  = 5, /* i32[K][3] (i,j,k) per hypothesis id, -1,-1,-1 if invalid    */
  SAC_COT_DBG_HYP_RT     = 6, /* f32[K][12] R (9, row-major) then t (3); zeros if invalid    */
  SAC_COT_DBG_HYP_SCORE  = 7, /* u64[K]  packed selection key per hypothesis (0 = invalid)   */
  SAC_COT_DBG_BEST_KEY   = 8, /* u64     max of HYP_SCORE                                    */
  SAC_COT_DBG_MASK       = 9, /* u32[ceil(N/32)] inlier bits of the winning hypothesis       */
  SAC_COT_DBG_HIST       = 10, /* u32[4096] histogram of (T_ij >> 4) over all edges           */
  SAC_COT_DBG_ADJ_FIRST  = 11  /* SECOND_ORDER mode: the first-order graph A, same layout as ADJ (ADJ is then A2,
                                  the graph every later stage ran on); FIRST_ORDER mode: identical to ADJ       */
};
SAC_COT_API int sac_cot_debug_get(sac_cot_ctx* ctx, int32_t pair, int32_t which,
                                  void* out, size_t cap, size_t* written);

/* Edge key layout (u64):  T_ij << 32 | (0xFFFF - i) << 16 | (0xFFFF - j),  i < j.
 *   larger key = better edge: more triangles first, then smaller i, then smaller j.
 * Hypothesis key layout (u64):  score << 16 | (0xFFFF - h).
 *   mode 0: score = inlier_count + 1;  mode 1: score = N*2^20 - sum_fixed + 1;  0 = invalid. */

SAC_COT_API const char* sac_cot_strerror(int status);
SAC_COT_API const char* sac_cot_version(void); /* "sac-cot-b200 <ver> (cuda sm_100a)" / "... (oracle)" */

#ifdef __cplusplus
}
#endif
#endif /* SAC_COT_H_ */
