// api.cu — host library behind the C ABI of include/sac_cot.h: contexts, workspace planning,
// the chunked kernel pipeline, key-pool growth/retry, the three sharded phases and the parity
// getter.  CUDA only: there is no CPU path; without a usable device every entry point fails
// with SAC_COT_E_NODEVICE or the cudaError_t.
//
// The reference ships no host layer to follow (/root/reference/README.md:1-2 is the repo); the
// contract implemented here is SURVEY.md §8b.
#include "../../include/sac_cot.h"
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the functions are resolved with dlopen/dlsym (see NcclApi)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace saccot;

namespace {

#define CU_TRY(expr)                                    \
  do {                                                  \
    const cudaError_t e__ = (expr);                     \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)

// SAC_COT_TRACE=1 in the environment (diagnostics): the device is synchronised after every launcher and the first
// one whose kernels fail is named on stderr.
static const bool g_trace = std::getenv("SAC_COT_TRACE") != nullptr;
#define KL_TRY(expr)                                                                          \
  do {                                                                                        \
    const int n__ = (expr);                                                                   \
    if (n__ < 0) {                                                                            \
      if (g_trace) std::fprintf(stderr, "sac_cot trace: launch failed (%d): %s\n", -n__, #expr); \
      return -n__;                                                                            \
    }                                                                                         \
    ctx->launches += n__;                                                                     \
    if (g_trace) {                                                                            \
      const cudaError_t e__ = cudaDeviceSynchronize();                                        \
      if (e__ != cudaSuccess) {                                                               \
        std::fprintf(stderr, "sac_cot trace: %s after %s\n", cudaGetErrorString(e__), #expr); \
        return static_cast<int>(e__);                                                         \
      }                                                                                       \
    }                                                                                         \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Sub-buffers of the arena for one chunk of pairs.
struct Layout {
  int pairs = 0;
  int max_n = 0, max_npad = 0, max_nblk = 0, max_stride = 0;
  size_t sum_n = 0, sum_npad = 0;
  int Ke = 0, m = 0, K = 0;
  // zero region (one memset): state | chunk | hist | t2 | ucount
  PairDev* state = nullptr;
  ChunkDev* chunk = nullptr;
  uint32_t* hist = nullptr;
  unsigned long long* t2 = nullptr;
  uint32_t* ucount = nullptr;  // edges per triangle work unit, [pairs][unit_pitch]
  size_t zero_bytes = 0;
  uint32_t* ubase = nullptr;   // key offset per unit inside the pair's key slice
  int unit_pitch = 0;
  // the rest
  PairDesc* desc = nullptr;
  float* in_src = nullptr;
  float* in_dst = nullptr;
  float* soa = nullptr;
  uint32_t* adj = nullptr;
  uint32_t* panel = nullptr;   // K-panel copy of the adjacency (tensor-core triangle path only)
  // second-order compatibility mode: A2 (+ its K-panel copy) and the saved [state | chunk] of the first pass
  uint32_t* adj2 = nullptr;
  uint32_t* panel2 = nullptr;
  unsigned char* pass1 = nullptr;
  size_t pass1_bytes = 0, adj_bytes = 0, panel_bytes = 0;
  uint32_t* adj_rank = nullptr;  // the graph S2/S3 ran on: adj, or adj2 in second-order mode
  uint32_t* theta = nullptr;   // per-pair pruning threshold (tensor-core triangle path only)
  void* theta_ws = nullptr;    // scratch of the multi-CTA threshold kernels (calls with very few pairs)
  uint2* tile_tab = nullptr;   // tensor-core path: (pair, row block << 16 | column block) per tile
  int total_tiles = 0;         // tensor-core path: tiles of the whole chunk
  // exact node pruning (kernels_prune.cu): plan per pair (in the zero region), degrees, kept lists, the tile list
  // without the pruned pairs' tiles and its length
  NodePlan* nplan = nullptr;
  unsigned short* deg = nullptr;
  unsigned short* kept = nullptr;
  uint32_t* keptbits = nullptr;  // [sum of Npad / 32], indexed like the inlier masks (PairDesc::mask_off)
  uint2* tile_tab2 = nullptr;
  int* tile_total = nullptr;     // int[4]: see launch_node_plan
  // RECT instance of the tensor-core kernel (node-pruned pairs): its tile list and the compact panel copies
  uint2* rect_tiles = nullptr;
  int rect_max_tiles = 0;
  uint32_t* kpanel = nullptr;
  long long kpanel_pair_words = 0;
  int max_npanel = 0;
  unsigned long long* sel = nullptr;
  unsigned long long* tie = nullptr;
  unsigned long long* top = nullptr;
  int32_t* tri = nullptr;
  float* rt = nullptr;
  unsigned long long* hyp_key = nullptr;
  uint32_t* mask = nullptr;
  float* outR = nullptr;
  float* outT = nullptr;
  int32_t* outInl = nullptr;
  unsigned long long* best_override = nullptr;
  // sharded single pair with in-library collectives: this rank's record, the gathered records, the merge summary
  unsigned long long* xsend = nullptr;
  unsigned long long* xrecv = nullptr;
  unsigned long long* xsum = nullptr;
  size_t xrec_len = 0;  // u64 entries per record = Npad + Ke + 2
  size_t total_bytes = 0;
  unsigned long long key_guess = 0;  // initial key-pool demand estimate
};

unsigned long long key_guess_for(int N) {
  const unsigned long long P = static_cast<unsigned long long>(N) * (N - 1) / 2;
  unsigned long long g = P / 8;  // 12.5 % edge density; grows on demand
  if (g < 4096) g = 4096;
  return g < P ? g : P;
}

size_t pair_bytes_estimate(int N, int K, int Ke, bool tensor_path, bool second_order) {
  const size_t npad = align_up(static_cast<size_t>(N), 128);
  return npad * (npad / 32) * 4 * (tensor_path ? 2 : 1) * (second_order ? 2 : 1) + key_guess_for(N) * 8 + npad * (6 * 4 + 8 + 24) + static_cast<size_t>(K) * (12 + 48 + 8) +
         static_cast<size_t>(Ke) * 16 + static_cast<size_t>(kTieCap) * 8 + kHistBins * 4 + 4096;
}

// Optional per-stage timing with CUDA events on the ctx stream ("stage_timing" knob).
enum Stage { ST_PACK = 0, ST_GRAPH, ST_SCAN, ST_THETA, ST_TRIANGLES, ST_KEPT, ST_SELECT, ST_APEX, ST_KABSCH, ST_SCORE, ST_FINALIZE, ST_EXCH1, ST_EXCH2,
             ST_MATCH_PREP, ST_MATCH_SWEEP, ST_MATCH_EXACT, ST_COUNT };
const char* const kStageNames[ST_COUNT] = {"pack", "graph", "scan", "theta", "triangles", "triangles_kept", "select", "apex", "kabsch", "score", "finalize", "exchange1", "exchange2",
                                           "match_prep", "match_sweep", "match_exact"};

// NCCL, bound at run time: the library carries no link-time dependency on it (a process that already loaded
// libnccl.so.2 — torch does — gets that copy; otherwise the system one).
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // SAC_COT_NCCL_LIB names the file to use (a process holds one library per soname: an application that will load
    // another libnccl.so.2 later — torch bundles its own — should point this at that copy, or load it first)
    const char* wanted = std::getenv("SAC_COT_NCCL_LIB");
    for (const char* name : {wanted, "libnccl.so.2", "libnccl.so"}) {
      if (!name || !*name) continue;
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce;
  });
  return api.ok ? &api : nullptr;
}
static_assert(sizeof(ncclUniqueId) == SAC_COT_COMM_ID_BYTES, "SAC_COT_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

struct StageTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int stage; cudaEvent_t a, b; };
  std::vector<Rec> pending;
  double acc_ms[ST_COUNT] = {0};
  int64_t calls[ST_COUNT] = {0};

  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      pool.push_back(e);
    }
    return pool[used++];
  }
  void resolve() {
    for (const Rec& r : pending) {
      float ms = 0.0f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        acc_ms[r.stage] += ms;
        ++calls[r.stage];
      }
    }
    pending.clear();
    used = 0;
  }
  void reset() {
    resolve();
    for (int k = 0; k < ST_COUNT; ++k) { acc_ms[k] = 0; calls[k] = 0; }
  }
  void destroy() {
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
  }
};

}  // namespace

// A lane = one internal stream with its own workspace.  Chunks of a batch are dealt to the lanes
// round-robin, so the tail of one chunk's kernels (and, in host mode, its H2D / D2H copies)
// overlaps the next chunk's work.  Lane streams fork from / join into the ctx stream with events,
// so a caller bracketing the call with events on the ctx stream still times everything.
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  unsigned long long* keys = nullptr;
  unsigned long long key_cap = 0;
  Layout lay;                     // layout of the chunk most recently enqueued on this lane
  std::vector<PairDesc> descs;
  std::vector<uint2> tile_tab;    // host copy of the tensor-core path's tile list
  std::vector<uint2> tile_scratch;
  // What the lane's arena holds descriptors and a tile list for.  A chunk with the same shapes, parameters and
  // partition as the previous one on this lane (every step of a steady workload) re-uses both: no plan(), no upload.
  std::vector<int32_t> plan_Ns;
  int plan_sig[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  // pinned staging of the descriptor / tile-list uploads (a pageable source would be staged by the driver and could
  // serialise the lanes); `uploaded` guards its re-use
  unsigned char* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  cudaEvent_t uploaded = nullptr;
};
constexpr int kMaxLanes = 4;

struct sac_cot_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // the caller-visible stream
  bool own_stream = false;
  int sm_count = 0;
  bool keep_debug = false;
  int chunk_pairs = 0;
  int n_lanes = 3;  // 3 chunks of ~86 pairs per 256-pair call: the uploads of later chunks hide behind the kernels of earlier ones
  int tri_path = 2;   // 0 = POPC bitset kernels, 1 = tensor-core (tcgen05 mxf4) kernel, 2 = by edge density per chunk
  int apex_path = 0;  // tests: 0 rank list, 1 exhaustive scan, 2 global-lookup kernel (kernels_select.cu)
  int tile_runs = 1;  // tensor-core path: deal tiles to the CTA pairs in runs (0: one at a time; experiments)
  int tri_dbg = 0;    // experiments only (bit 0: skip the tensor-core kernel's epilogue work; results are void)
  int tri_prune = 1;  // tensor-core path: drop edges below the per-pair threshold (exact for the selection)
  // tensor-core path: pairs whose selectable edges join few high-degree nodes count triangles for those nodes only
  // (kernels_prune.cu; exact).  0 off, 1 on unless keep_debug is set (the T_NODE / EDGE_KEYS / HIST dumps then cover
  // the kept nodes only), 2 on whenever the kept list fits (tests)
  int node_prune = 1;
  int node_prune_cost = 200;  // a pair is pruned if (sum of kept degrees) x cost <= Npad^2
  int node_prune_rect = 1;    // kept rows of long pairs: 1 = the tensor-core kernel's RECT instance for chunks of >= 4 pairs, 2 = always, 0 = POPC kept-row kernel
  // Trying costs five near-empty launches per chunk when nothing can be pruned (indoor-scale graphs).  The ctx
  // therefore watches the device-side count of pruned pairs: after two calls in a row that tried and pruned nothing
  // it stops trying for node_prune_probe calls, then tries again.  Results never depend on it.
  int node_prune_probe = 30;
  int np_idle = 0;        // calls in a row that tried and pruned nothing
  int np_skip_left = 0;   // calls that will not try
  bool np_try = true;     // the current call tries
  bool np_tried = false;  // some call since the last verdict enqueued the pruning kernels
  uint32_t np_seen = 0;   // device counter at the last verdict
  int64_t launches = 0;
  int64_t retries = 0;
  int deferred_status = 0;  // device-location calls: status discovered after the fact

  Lane lanes[kMaxLanes];
  cudaEvent_t fork_event = nullptr;
  std::vector<ChunkDev*> h_chunks;  // pinned read-backs of the chunk headers (host-location calls)
  StickyDev* d_sticky = nullptr;    // device: overflow record that survives across calls
  StickyDev* h_sticky = nullptr;    // pinned [3]: the record before / after the device-location calls whose verdict is open; [2]: after a host-location call
  cudaEvent_t sticky_event = nullptr;
  bool sticky_pending = false;

  // what is resident in lane 0's workspace (debug_get / sharded phases)
  sac_cot_params prm{};
  bool ws_valid = false;
  int sh_rank = 0, sh_world = 1, sh_N = 0;
  bool sh_valid = false;

  // host staging for the pointer-array batch entry point
  std::vector<float> stage_src, stage_dst;
  StageTimer timer;

  // correspondence front end (sac_cot_match_packed): its own grow-only workspace on the ctx stream
  unsigned char* match_arena = nullptr;
  size_t match_arena_bytes = 0;
  unsigned char* h_match = nullptr;   // pinned staging of the pair table
  size_t h_match_bytes = 0;
  cudaEvent_t match_uploaded = nullptr;
  std::vector<int64_t> match_sig;     // shapes the resident pair table was built for
  int match_dbg = 0;                  // experiments: the sweep kernel prints its barrier wait cycles (CTA 0)
  int match_path = 1;                 // 1: tensor-core sweep + exact decision (dim <= kMatchMaxDim); 0: exhaustive exact scan

  // sharded single pair with in-library collectives
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  int comm_rank = 0, comm_world = 0;   // world 0 = no communicator
  unsigned long long* h_xsum = nullptr;  // pinned: merge summary {any rank overflowed, largest key-pool demand}
};

namespace {

// struct_size versioning: version 1 (32 bytes) ends with compat_mode (then called `reserved`, must be 0)
int check_params(const sac_cot_params* p) {
  if (!p) return SAC_COT_E_NULL;
  if (p->struct_size != sizeof(sac_cot_params) && p->struct_size != SAC_COT_PARAMS_SIZE_V1) return SAC_COT_E_PARAMS;
  if (!(p->tau_compat > 0.0f) || !(p->tau_inlier > 0.0f)) return SAC_COT_E_PARAMS;
  if (p->num_edges < 1 || p->num_edges > SAC_COT_MAX_EDGES) return SAC_COT_E_PARAMS;
  if (p->apex_per_edge < 1 || p->apex_per_edge > SAC_COT_MAX_APEX) return SAC_COT_E_PARAMS;
  if (p->num_edges * p->apex_per_edge > SAC_COT_MAX_HYPOTHESES) return SAC_COT_E_PARAMS;
  if (p->score_mode != 0 && p->score_mode != 1) return SAC_COT_E_PARAMS;
  if (p->refit != 0 && p->refit != 1) return SAC_COT_E_PARAMS;
  if (p->struct_size == SAC_COT_PARAMS_SIZE_V1) return p->compat_mode == 0 ? SAC_COT_OK : SAC_COT_E_PARAMS;
  if (p->compat_mode != SAC_COT_COMPAT_FIRST_ORDER && p->compat_mode != SAC_COT_COMPAT_SECOND_ORDER) return SAC_COT_E_PARAMS;
  if (p->so_min_common < 0 || p->so_min_common > 65535) return SAC_COT_E_PARAMS;
  if (p->reserved != 0) return SAC_COT_E_PARAMS;
  return SAC_COT_OK;
}
// the caller's (validated) struct of either version as the current one
sac_cot_params normalized(const sac_cot_params* in) {
  sac_cot_params out{};
  std::memcpy(&out, in, std::min<size_t>(in->struct_size, sizeof(out)));
  out.struct_size = sizeof(out);
  return out;
}

// Builds the descriptors and the arena layout for pairs with the given sizes.
void plan(const int32_t* Ns, int pairs, const sac_cot_params& prm, bool need_input_copy, bool tensor_path, int xworld,
          std::vector<PairDesc>& descs, std::vector<uint2>& tile_tab, Layout& L) {
  L = Layout();
  L.pairs = pairs;
  L.Ke = prm.num_edges;
  L.m = prm.apex_per_edge;
  L.K = L.Ke * L.m;
  descs.resize(pairs);
  size_t pt = 0, soa = 0, adj = 0, node = 0, mask = 0, panel = 0;
  int tiles = 0;
  tile_tab.clear();
  for (int b = 0; b < pairs; ++b) {
    PairDesc& d = descs[b];
    d.N = Ns[b];
    d.Npad = static_cast<int32_t>(align_up(static_cast<size_t>(Ns[b]), 128));
    d.stride = d.Npad / 32;
    d.nblk = d.Npad / 128;
    d.pt_off = static_cast<int64_t>(pt);
    d.soa_off = static_cast<int64_t>(soa);
    d.adj_off = static_cast<int64_t>(adj);
    d.node_off = static_cast<int64_t>(node);
    d.mask_off = static_cast<int64_t>(mask);
    d.panel_off = static_cast<int64_t>(panel);
    d.npanel = (d.Npad + 255) / 256;
    d.tile_base = tiles;
    if (tensor_path) {
      panel += static_cast<size_t>(d.npanel) * d.Npad * 8;
      const int nJ = (d.N + kMmaTileN - 1) / kMmaTileN;
      for (int jq = 0; jq < nJ; ++jq)
        for (int ib = 0, c = mma_tiles_of_jblock(d.N, jq); ib < c; ++ib)
          tile_tab.push_back(make_uint2(static_cast<unsigned>(b), (static_cast<unsigned>(ib) << 16) | static_cast<unsigned>(jq)));
      tiles = static_cast<int>(tile_tab.size());
    }
    pt += d.N;
    soa += static_cast<size_t>(6) * d.Npad;
    adj += static_cast<size_t>(d.Npad) * d.stride;
    node += d.Npad;
    mask += d.Npad / 32;
    L.max_n = std::max(L.max_n, d.N);
    L.max_npad = std::max(L.max_npad, d.Npad);
    L.max_nblk = std::max(L.max_nblk, d.nblk);
    L.max_stride = std::max(L.max_stride, d.stride);
    L.key_guess += key_guess_for(d.N);
  }
  L.sum_n = pt;
  L.sum_npad = node;
  L.total_tiles = tiles;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  // zero region first
  const size_t o_state = take(sizeof(PairDev) * pairs);
  const size_t o_chunk = take(sizeof(ChunkDev));
  L.pass1_bytes = off;  // [state | chunk]: what the second-order mode saves of its first pass
  const size_t o_hist = take(sizeof(uint32_t) * kHistBins * pairs);
  const size_t o_t2 = take(sizeof(unsigned long long) * node);
  L.unit_pitch = static_cast<int>(unit_count(static_cast<unsigned int>(L.max_nblk)));
  const size_t o_ucount = take(sizeof(uint32_t) * L.unit_pitch * pairs);
  const size_t o_nplan = take(sizeof(NodePlan) * pairs);
  const size_t o_ttotal = take(4 * sizeof(int));
  L.zero_bytes = off;
  const size_t o_ubase = take(sizeof(uint32_t) * L.unit_pitch * pairs);
  const size_t o_desc = take(sizeof(PairDesc) * pairs);
  const size_t o_insrc = need_input_copy ? take(sizeof(float) * 3 * pt) : 0;
  const size_t o_indst = need_input_copy ? take(sizeof(float) * 3 * pt) : 0;
  const size_t o_soa = take(sizeof(float) * soa);
  const size_t o_adj = take(sizeof(uint32_t) * adj);
  const size_t o_panel = tensor_path ? take(sizeof(uint32_t) * panel) : 0;
  const bool so = prm.compat_mode == SAC_COT_COMPAT_SECOND_ORDER;
  L.adj_bytes = sizeof(uint32_t) * adj;
  L.panel_bytes = tensor_path ? sizeof(uint32_t) * panel : 0;
  const size_t o_adj2 = so ? take(L.adj_bytes) : 0;
  const size_t o_panel2 = so && tensor_path ? take(L.panel_bytes) : 0;
  const size_t o_pass1 = so ? take(L.pass1_bytes) : 0;
  const size_t o_theta = tensor_path ? take(sizeof(uint32_t) * pairs) : 0;
  const size_t theta_ws_bytes = tensor_path ? theta_scratch_bytes(pairs) : 0;
  const size_t o_theta_ws = theta_ws_bytes ? take(theta_ws_bytes) : 0;
  const size_t o_tiles = tensor_path ? take(sizeof(uint2) * std::max(1, tiles)) : 0;
  const size_t o_tiles2 = tensor_path ? take(sizeof(uint2) * std::max(1, tiles)) : 0;
  const size_t o_deg = tensor_path ? take(sizeof(unsigned short) * node) : 0;
  const size_t o_kept = tensor_path ? take(sizeof(unsigned short) * kNodeKeepMax * pairs) : 0;
  const size_t o_keptbits = tensor_path ? take(sizeof(uint32_t) * mask) : 0;
  // RECT path: only for chunks whose rows are long enough and short enough for the node pruning at all
  const bool rect = tensor_path && L.max_npad >= kRectMinNpad && L.max_npad <= kNodePruneMaxNpad;
  L.max_npanel = (L.max_npad + 255) / 256;
  L.rect_max_tiles = rect ? pairs * (kRectRows / kMmaTileM) * ((L.max_n + kMmaTileN - 1) / kMmaTileN) : 0;
  L.kpanel_pair_words = rect ? static_cast<long long>((L.max_npanel + 1) & ~1) * kRectRows * 8 : 0;
  const size_t o_rect = rect ? take(sizeof(uint2) * static_cast<size_t>(L.rect_max_tiles)) : 0;
  const size_t o_kpanel = rect ? take(sizeof(uint32_t) * static_cast<size_t>(L.kpanel_pair_words) * pairs) : 0;
  const size_t o_sel = take(sizeof(unsigned long long) * L.Ke * pairs);
  const size_t o_tie = take(sizeof(unsigned long long) * kTieCap * pairs);
  const size_t o_top = take(sizeof(unsigned long long) * L.Ke * pairs);
  const size_t o_tri = take(sizeof(int32_t) * 3 * L.K * pairs);
  const size_t o_rt = take(sizeof(float) * 12 * L.K * pairs);
  const size_t o_hk = take(sizeof(unsigned long long) * L.K * pairs);
  const size_t o_mask = take(sizeof(uint32_t) * mask);
  const size_t o_R = take(sizeof(float) * 9 * pairs);
  const size_t o_T = take(sizeof(float) * 3 * pairs);
  const size_t o_inl = take(sizeof(int32_t) * pairs);
  const size_t o_bo = take(sizeof(unsigned long long));
  L.xrec_len = xworld > 0 ? node + static_cast<size_t>(L.Ke) + 2 : 0;
  const size_t o_xsend = xworld > 0 ? take(sizeof(unsigned long long) * L.xrec_len) : 0;
  const size_t o_xrecv = xworld > 0 ? take(sizeof(unsigned long long) * L.xrec_len * xworld) : 0;
  const size_t o_xsum = xworld > 0 ? take(sizeof(unsigned long long) * 2) : 0;
  L.total_bytes = off;
  // stash offsets as pointers relative to nullptr; bind() adds the arena base
  L.state = reinterpret_cast<PairDev*>(o_state);
  L.chunk = reinterpret_cast<ChunkDev*>(o_chunk);
  L.hist = reinterpret_cast<uint32_t*>(o_hist);
  L.t2 = reinterpret_cast<unsigned long long*>(o_t2);
  L.ucount = reinterpret_cast<uint32_t*>(o_ucount);
  L.ubase = reinterpret_cast<uint32_t*>(o_ubase);
  L.desc = reinterpret_cast<PairDesc*>(o_desc);
  L.in_src = need_input_copy ? reinterpret_cast<float*>(o_insrc) : nullptr;
  L.in_dst = need_input_copy ? reinterpret_cast<float*>(o_indst) : nullptr;
  L.soa = reinterpret_cast<float*>(o_soa);
  L.adj = reinterpret_cast<uint32_t*>(o_adj);
  L.panel = tensor_path ? reinterpret_cast<uint32_t*>(o_panel) : nullptr;
  L.adj2 = so ? reinterpret_cast<uint32_t*>(o_adj2) : nullptr;
  L.panel2 = so && tensor_path ? reinterpret_cast<uint32_t*>(o_panel2) : nullptr;
  L.pass1 = so ? reinterpret_cast<unsigned char*>(o_pass1) : nullptr;
  L.theta = tensor_path ? reinterpret_cast<uint32_t*>(o_theta) : nullptr;
  L.theta_ws = theta_ws_bytes ? reinterpret_cast<void*>(o_theta_ws + 1) : nullptr;  // +1: offset 0 must not read as "absent"
  L.tile_tab = tensor_path ? reinterpret_cast<uint2*>(o_tiles) : nullptr;
  L.nplan = reinterpret_cast<NodePlan*>(o_nplan);
  L.tile_tab2 = tensor_path ? reinterpret_cast<uint2*>(o_tiles2) : nullptr;
  L.tile_total = reinterpret_cast<int*>(o_ttotal);
  L.deg = tensor_path ? reinterpret_cast<unsigned short*>(o_deg) : nullptr;
  L.kept = tensor_path ? reinterpret_cast<unsigned short*>(o_kept) : nullptr;
  L.keptbits = tensor_path ? reinterpret_cast<uint32_t*>(o_keptbits) : nullptr;
  L.rect_tiles = rect ? reinterpret_cast<uint2*>(o_rect + 1) : nullptr;      // +1: offset 0 must not read as "absent"
  L.kpanel = rect ? reinterpret_cast<uint32_t*>(o_kpanel + 1) : nullptr;
  L.sel = reinterpret_cast<unsigned long long*>(o_sel);
  L.tie = reinterpret_cast<unsigned long long*>(o_tie);
  L.top = reinterpret_cast<unsigned long long*>(o_top);
  L.tri = reinterpret_cast<int32_t*>(o_tri);
  L.rt = reinterpret_cast<float*>(o_rt);
  L.hyp_key = reinterpret_cast<unsigned long long*>(o_hk);
  L.mask = reinterpret_cast<uint32_t*>(o_mask);
  L.outR = reinterpret_cast<float*>(o_R);
  L.outT = reinterpret_cast<float*>(o_T);
  L.outInl = reinterpret_cast<int32_t*>(o_inl);
  L.best_override = reinterpret_cast<unsigned long long*>(o_bo);
  L.xsend = xworld > 0 ? reinterpret_cast<unsigned long long*>(o_xsend) : nullptr;
  L.xrecv = xworld > 0 ? reinterpret_cast<unsigned long long*>(o_xrecv) : nullptr;
  L.xsum = xworld > 0 ? reinterpret_cast<unsigned long long*>(o_xsum) : nullptr;
}

template <typename T>
inline T* rebase(T* rel, unsigned char* base, bool present = true) {
  return present ? reinterpret_cast<T*>(base + reinterpret_cast<size_t>(rel)) : nullptr;
}

void bind(Layout& L, unsigned char* base, bool has_input, bool tensor_path) {
  L.state = rebase(L.state, base);
  L.chunk = rebase(L.chunk, base);
  L.hist = rebase(L.hist, base);
  L.t2 = rebase(L.t2, base);
  L.ucount = rebase(L.ucount, base);
  L.ubase = rebase(L.ubase, base);
  L.desc = rebase(L.desc, base);
  L.in_src = rebase(L.in_src, base, has_input);
  L.in_dst = rebase(L.in_dst, base, has_input);
  L.soa = rebase(L.soa, base);
  L.adj = rebase(L.adj, base);
  L.panel = rebase(L.panel, base, tensor_path);
  const bool so = L.pass1 != nullptr;
  L.adj2 = rebase(L.adj2, base, so);
  L.panel2 = rebase(L.panel2, base, so && tensor_path);
  L.pass1 = rebase(L.pass1, base, so);
  L.adj_rank = so ? L.adj2 : L.adj;
  L.theta = rebase(L.theta, base, tensor_path);
  L.theta_ws = L.theta_ws ? static_cast<void*>(base + (reinterpret_cast<size_t>(L.theta_ws) - 1)) : nullptr;
  L.tile_tab = rebase(L.tile_tab, base, tensor_path);
  L.nplan = rebase(L.nplan, base);
  L.tile_tab2 = rebase(L.tile_tab2, base, tensor_path);
  L.tile_total = rebase(L.tile_total, base);
  L.deg = rebase(L.deg, base, tensor_path);
  L.kept = rebase(L.kept, base, tensor_path);
  L.keptbits = rebase(L.keptbits, base, tensor_path);
  L.rect_tiles = L.rect_tiles ? reinterpret_cast<uint2*>(base + (reinterpret_cast<size_t>(L.rect_tiles) - 1)) : nullptr;
  L.kpanel = L.kpanel ? reinterpret_cast<uint32_t*>(base + (reinterpret_cast<size_t>(L.kpanel) - 1)) : nullptr;
  L.sel = rebase(L.sel, base);
  L.tie = rebase(L.tie, base);
  L.top = rebase(L.top, base);
  L.tri = rebase(L.tri, base);
  L.rt = rebase(L.rt, base);
  L.hyp_key = rebase(L.hyp_key, base);
  L.mask = rebase(L.mask, base);
  L.outR = rebase(L.outR, base);
  L.outT = rebase(L.outT, base);
  L.outInl = rebase(L.outInl, base);
  L.best_override = rebase(L.best_override, base);
  const bool x = L.xrec_len != 0;
  L.xsend = rebase(L.xsend, base, x);
  L.xrecv = rebase(L.xrecv, base, x);
  L.xsum = rebase(L.xsum, base, x);
}

int sync_all(sac_cot_ctx* ctx) {
  for (int l = 0; l < kMaxLanes; ++l)
    if (ctx->lanes[l].stream) CU_TRY(cudaStreamSynchronize(ctx->lanes[l].stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ensure_arena(sac_cot_ctx* ctx, Lane& ln, size_t bytes) {
  if (bytes <= ln.arena_bytes) return 0;
  if (int rc = sync_all(ctx)) return rc;
  if (ln.arena) CU_TRY(cudaFree(ln.arena));
  ln.arena = nullptr;
  ln.arena_bytes = 0;
  ln.plan_sig[0] = -1;  // descriptors and tile list went with the arena
  const size_t want = align_up(bytes + bytes / 8, 1 << 20);
  cudaError_t e = cudaMalloc(&ln.arena, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    e = cudaMalloc(&ln.arena, bytes);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return SAC_COT_E_NOMEM;
    }
    ln.arena_bytes = bytes;
  } else {
    ln.arena_bytes = want;
  }
  if (&ln == &ctx->lanes[0]) {
    ctx->ws_valid = false;
    ctx->sh_valid = false;
  }
  return 0;
}

int ensure_keys(sac_cot_ctx* ctx, Lane& ln, unsigned long long cap) {
  if (cap <= ln.key_cap) return 0;
  if (int rc = sync_all(ctx)) return rc;
  if (ln.keys) CU_TRY(cudaFree(ln.keys));
  ln.keys = nullptr;
  ln.key_cap = 0;
  if (cudaMalloc(&ln.keys, cap * sizeof(unsigned long long)) != cudaSuccess) {
    (void)cudaGetLastError();
    return SAC_COT_E_NOMEM;
  }
  ln.key_cap = cap;
  return 0;
}

// Deferred status of device-location calls.  The device-side sticky record (overflowing chunks so far, largest
// demand) is copied to pinned memory on the ctx stream BEFORE the first call whose verdict is still open and AFTER
// the latest one, with an event behind the second copy.  The pinned copies are read only once that event has
// completed; a changed overflow count means some chunk of those calls ran out of key-pool space: their outputs are
// void, "last_status" reports SAC_COT_E_NOMEM once, and every lane's pool is grown to the recorded demand so the
// caller's next call succeeds.  block = false never waits (a verdict that is not in yet is merely late).
// node pruning: what the calls since the last verdict achieved (counter = pairs pruned so far on this ctx)
void node_prune_verdict(sac_cot_ctx* ctx, uint32_t counter) {
  if (ctx->np_tried) {
    if (counter == ctx->np_seen) {
      if (++ctx->np_idle >= 2) {
        ctx->np_skip_left = ctx->node_prune_probe;
        ctx->np_idle = 0;
      }
    } else {
      ctx->np_idle = 0;
    }
    ctx->np_tried = false;
  }
  ctx->np_seen = counter;
}

int sticky_before(sac_cot_ctx* ctx) {
  if (ctx->sticky_pending) return 0;  // an open verdict keeps its "before" snapshot and will cover this call too
  CU_TRY(cudaMemcpyAsync(&ctx->h_sticky[0], ctx->d_sticky, sizeof(StickyDev), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}
int sticky_after(sac_cot_ctx* ctx) {
  CU_TRY(cudaMemcpyAsync(&ctx->h_sticky[1], ctx->d_sticky, sizeof(StickyDev), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaEventRecord(ctx->sticky_event, ctx->stream));
  ctx->sticky_pending = true;
  return 0;
}
int resolve_pending(sac_cot_ctx* ctx, bool block) {
  if (block)
    if (int rc = sync_all(ctx)) return rc;
  if (!ctx->sticky_pending) return 0;
  if (!block) {
    const cudaError_t q = cudaEventQuery(ctx->sticky_event);
    if (q == cudaErrorNotReady) return 0;
    if (q != cudaSuccess) return static_cast<int>(q);
  }
  ctx->sticky_pending = false;
  node_prune_verdict(ctx, ctx->h_sticky[1].pruned_total);
  if (ctx->h_sticky[1].overflow_count != ctx->h_sticky[0].overflow_count) {
    ctx->deferred_status = SAC_COT_E_NOMEM;
    const unsigned long long demand = ctx->h_sticky[1].max_total_edges;
    ++ctx->retries;
    for (int l = 0; l < ctx->n_lanes; ++l)
      if (int rc = ensure_keys(ctx, ctx->lanes[l], demand + demand / 8 + 1024)) return rc;
  }
  return 0;
}

// Enqueues the whole pipeline on lane `ln` for the pairs described by ln.descs / ln.lay.
//   d_src/d_dst : device AoS inputs for this chunk (pt_off relative to them)
//   dR/dT/dInl  : device outputs for this chunk
int enqueue_pipeline(sac_cot_ctx* ctx, Lane& ln, const float* d_src, const float* d_dst, float* dR, float* dT,
                     int32_t* dInl, int rank, int world, bool stop_after_edges) {
  Layout& L = ln.lay;
  const sac_cot_params& prm = ctx->prm;
  LaunchCtx lc{ln.stream, ctx->sm_count};
  StageTimer& tm = ctx->timer;
  if (tm.enabled && tm.pending.size() > 8192) tm.resolve();
  cudaEvent_t ev_prev = nullptr;
  auto mark = [&](int stage) {  // closes `stage`: records an event and pairs it with the previous one
    if (!tm.enabled) return;
    cudaEvent_t e = tm.next();
    if (!e) return;
    cudaEventRecord(e, ln.stream);
    if (stage >= 0 && ev_prev) tm.pending.push_back({stage, ev_prev, e});
    ev_prev = e;
  };
  CU_TRY(cudaMemsetAsync(L.state, 0, L.zero_bytes, ln.stream));
  mark(-1);
  KL_TRY(launch_pack_soa(lc, L.desc, L.pairs, L.max_npad, d_src, d_dst, L.soa));
  mark(ST_PACK);
  // planned with the tensor-core buffers: mode 1 (forced) or 2 (the key scan decides from the edge density);
  // both triangle kernels are then enqueued and the one not chosen returns at once
  const int tri_mode = L.panel != nullptr ? ctx->tri_path : 0;
  KL_TRY(launch_graph(lc, L.desc, L.pairs, L.max_nblk, L.soa, L.adj, L.panel, L.ucount, L.unit_pitch, prm.tau_compat));
  mark(ST_GRAPH);
  KL_TRY(launch_unit_scan(lc, L.desc, L.pairs, L.state, L.ucount, L.ubase, L.unit_pitch, rank, world));
  KL_TRY(launch_key_scan(lc, L.desc, L.pairs, L.state, L.chunk, ctx->d_sticky, ln.key_cap, tri_mode, nullptr));
  mark(ST_SCAN);
  const uint32_t* adj_use = L.adj;
  const uint32_t* panel_use = L.panel;
  if (L.adj2 != nullptr) {
    // ---- second-order compatibility: a first counting pass over A leaves C_ij = T_ij(A) of every edge that reaches
    //      so_min_common in the key list (the tensor-core kernel's pruning threshold IS that cut; the POPC kernels
    //      keep every key and the scatter filters); the list becomes A2 and everything below runs on A2 ----
    const uint32_t cmin = static_cast<uint32_t>(prm.so_min_common);
    if (tri_mode != 0) {
      KL_TRY(launch_fill_u32(lc, L.theta, cmin | 0x80000000u, L.pairs));  // bit 31: fixed, never raised
      KL_TRY(launch_triangles_mma(lc, L.desc, L.tile_tab, L.total_tiles, nullptr, nullptr, L.adj, L.panel, L.state, L.chunk, ln.keys, L.theta,
                                  L.hist, L.t2, L.Ke, 0, ctx->tri_dbg));
    }
    if (tri_mode != 1)
      KL_TRY(launch_triangles(lc, L.desc, L.pairs, L.max_nblk, L.max_stride, L.adj, L.state, L.chunk, ln.keys, L.ubase,
                              L.unit_pitch, L.hist, L.t2, rank, world));
    CU_TRY(cudaMemcpyAsync(L.pass1, L.state, L.pass1_bytes, cudaMemcpyDeviceToDevice, ln.stream));
    CU_TRY(cudaMemsetAsync(L.state, 0, L.zero_bytes, ln.stream));
    CU_TRY(cudaMemsetAsync(L.adj2, 0, L.adj_bytes, ln.stream));
    if (L.panel2) CU_TRY(cudaMemsetAsync(L.panel2, 0, L.panel_bytes, ln.stream));
    const PairDev* state1 = reinterpret_cast<const PairDev*>(L.pass1);
    const ChunkDev* chunk1 = reinterpret_cast<const ChunkDev*>(L.pass1 + (reinterpret_cast<unsigned char*>(L.chunk) - reinterpret_cast<unsigned char*>(L.state)));
    KL_TRY(launch_second_order_scatter(lc, L.desc, L.pairs, state1, chunk1, ln.keys, cmin, L.adj2, L.panel2, L.ucount, L.unit_pitch));
    KL_TRY(launch_unit_scan(lc, L.desc, L.pairs, L.state, L.ucount, L.ubase, L.unit_pitch, rank, world));
    KL_TRY(launch_key_scan(lc, L.desc, L.pairs, L.state, L.chunk, ctx->d_sticky, ln.key_cap, tri_mode, chunk1));
    adj_use = L.adj2;
    panel_use = L.panel2;
    mark(ST_GRAPH);  // the first pass and the rebuild count as graph construction
  }
  // exact node pruning: first-order graph, rows short enough for the kept-row kernel; not for the sharded phases
  // (stop_after_edges: their later parts rank apexes from complete node sums, also with world = 1)
  const bool node_prune = tri_mode != 0 && ctx->tri_prune && ctx->node_prune != 0 && L.adj2 == nullptr && world <= 1 &&
                          !stop_after_edges && ctx->np_try &&
                          L.max_npad <= kNodePruneMaxNpad && L.total_tiles > 0 && ctx->apex_path != 2 &&
                          (ctx->node_prune >= 2 || !ctx->keep_debug);
  if (tri_mode != 0) {
    KL_TRY(launch_tri_theta(lc, L.desc, L.pairs, L.max_npad, adj_use, L.chunk, L.theta, L.theta_ws, L.Ke, ctx->tri_prune));
    if (node_prune) ctx->np_tried = true;
    // RECT instance for the kept rows of long pairs — from four pairs per chunk on (a single pair's 40-80 tiles do not
    // fill the 74 CTA pairs: 66 us against 56 us for the kept-row POPC kernel at N = 10 000); node_prune_rect = 2 forces it
    const bool use_rect = node_prune && L.kpanel != nullptr && (ctx->node_prune_rect == 2 || (ctx->node_prune_rect == 1 && L.pairs >= 4));
    if (node_prune)
      KL_TRY(launch_node_plan(lc, L.desc, L.pairs, L.max_npad, adj_use, L.chunk, ctx->d_sticky, L.state, L.theta, L.deg, L.nplan, L.kept, L.keptbits, L.tile_tab,
                              L.total_tiles, L.tile_tab2, L.tile_total, ctx->node_prune_cost, ctx->node_prune, L.rect_tiles, panel_use,
                              use_rect ? L.kpanel : nullptr, L.kpanel_pair_words, L.max_npanel));
    mark(ST_THETA);
    KL_TRY(launch_triangles_mma(lc, L.desc, L.tile_tab, L.total_tiles, node_prune ? L.tile_total : nullptr, L.tile_tab2,
                                adj_use, panel_use, L.state, L.chunk, ln.keys, L.theta, L.hist, L.t2, L.Ke, ctx->tri_prune, ctx->tri_dbg));
    if (node_prune) mark(ST_TRIANGLES);
    if (use_rect)
      KL_TRY(launch_triangles_mma_rect(lc, L.desc, L.rect_tiles, L.rect_max_tiles, L.tile_total + 2, adj_use, panel_use, L.state,
                                       L.chunk, ln.keys, L.theta, L.hist, L.t2, L.Ke, ctx->tri_prune, ctx->tri_dbg, L.nplan, L.kept,
                                       L.kpanel, L.kpanel_pair_words));
    if (node_prune)
      KL_TRY(launch_triangles_kept(lc, L.desc, L.pairs, L.max_n, L.max_stride, adj_use, L.nplan, L.kept, L.keptbits, L.chunk, L.state, ln.keys,
                                   L.hist, L.t2));
    if (node_prune) mark(ST_KEPT);
  }
  if (tri_mode != 1)
    KL_TRY(launch_triangles(lc, L.desc, L.pairs, L.max_nblk, L.max_stride, adj_use, L.state, L.chunk, ln.keys, L.ubase,
                            L.unit_pitch, L.hist, L.t2, rank, world));
  mark(ST_TRIANGLES);
  KL_TRY(launch_select_edges(lc, L.pairs, L.state, L.chunk, ln.keys, L.hist, L.sel, L.tie, L.top, L.Ke));
  mark(ST_SELECT);
  if (stop_after_edges) return 0;
  const float tau2 = prm.tau_inlier * prm.tau_inlier;
  KL_TRY(launch_select_apex(lc, L.desc, L.pairs, L.max_npad, adj_use, L.t2, L.top, L.tri, L.Ke, L.m, ctx->apex_path,
                            node_prune ? L.nplan : nullptr, node_prune ? L.deg : nullptr, node_prune ? L.keptbits : nullptr));
  mark(ST_APEX);
  KL_TRY(launch_kabsch(lc, L.desc, L.pairs, L.soa, L.tri, L.rt, L.K));
  mark(ST_KABSCH);
  KL_TRY(launch_score(lc, L.desc, L.pairs, L.max_n, L.soa, L.tri, L.rt, L.hyp_key, L.state, tau2, L.K, 0, L.K,
                      prm.score_mode));
  mark(ST_SCORE);
  KL_TRY(launch_finalize(lc, L.desc, L.pairs, L.soa, L.rt, L.state, nullptr, L.mask, dR, dT, dInl, tau2, L.K,
                         prm.refit));
  mark(ST_FINALIZE);
  return 0;
}

// lane streams start after everything already enqueued on the ctx stream ...
int fork_lanes(sac_cot_ctx* ctx, int n) {
  CU_TRY(cudaEventRecord(ctx->fork_event, ctx->stream));
  for (int l = 0; l < n; ++l) CU_TRY(cudaStreamWaitEvent(ctx->lanes[l].stream, ctx->fork_event, 0));
  return 0;
}
// ... and the ctx stream continues after everything enqueued on the lanes
int join_lanes(sac_cot_ctx* ctx, int n) {
  for (int l = 0; l < n; ++l) {
    CU_TRY(cudaEventRecord(ctx->lanes[l].done, ctx->lanes[l].stream));
    CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->lanes[l].done, 0));
  }
  return 0;
}

// Tensor-core path: cluster c of the kernel walks entries c, c + ncl, c + 2 ncl, ... of the tile list.  plan()
// emits the list sorted by (pair, column block, row block); dealt out one tile at a time, every cluster
// would touch every pair (a pair change costs the epilogue a histogram flush and two barriers, and its
// first window load a dependent descriptor fetch).  Reorder it so that a cluster gets runs of B consecutive
// tiles: few pair changes per cluster, while the pairs in flight at any time (about ncl * B / tiles per
// pair) still fit the L2 with their K-panel copies.  The tail that does not fill a whole round of runs is
// dealt out tile by tile, so the clusters' tile counts differ by at most one, exactly as before.
void interleave_tile_runs(std::vector<uint2>& tab, std::vector<uint2>& scratch, const std::vector<PairDesc>& descs,
                          int sm_count) {
  const int T = static_cast<int>(tab.size());
  const int ncl = mma_clusters(T, sm_count);
  if (ncl <= 1 || T < 2 * ncl) return;
  size_t panel_bytes = 0;
  for (const PairDesc& d : descs) panel_bytes += static_cast<size_t>(d.npanel) * d.Npad * 32;
  const double pairs = static_cast<double>(descs.size());
  const double pairs_in_flight = std::max(1.0, 32.0e6 / (static_cast<double>(panel_bytes) / pairs));
  const int rounds = T / ncl;
  int B = static_cast<int>(std::min<double>(rounds, std::max(1.0, pairs_in_flight * (T / pairs) / ncl)));
  B = std::min(B, 256);
  if (B < rounds) {  // a run length near B that leaves the smallest tail
    int best = B;
    for (int c = B; c >= std::max(1, (3 * B) / 4); --c)
      if (T % (c * ncl) < T % (best * ncl)) best = c;
    B = best;
  }
  if (B <= 1) return;
  const int per_cluster = (T / (B * ncl)) * B, bulk = per_cluster * ncl;
  scratch.resize(tab.size());
  for (int o = 0; o < bulk; ++o) {
    const int r = o / B;
    scratch[static_cast<size_t>(r % ncl) + static_cast<size_t>((r / ncl) * B + o % B) * ncl] = tab[o];
  }
  for (int o = bulk; o < T; ++o) scratch[static_cast<size_t>((o - bulk) % ncl) + static_cast<size_t>(per_cluster + (o - bulk) / ncl) * ncl] = tab[o];
  tab.swap(scratch);
}

// Makes the lane ready for a chunk of `pairs` pairs of sizes Ns: layout planned and bound to the arena, descriptors
// and (tensor-core path) tile list resident on the device.  rank / world: sharded single pair, only the tiles of the
// cells this rank owns are listed.  xworld > 0 adds the exchange buffers of the in-library collectives.
// A chunk with the same shapes, parameters and partition as the lane's previous one costs nothing here.
int prepare_lane(sac_cot_ctx* ctx, Lane& ln, const int32_t* Ns, int pairs, const sac_cot_params& prm, bool host,
                 bool tensor, int rank, int world, int xworld) {
  const int sig[8] = {pairs, prm.num_edges, prm.apex_per_edge, (host ? 1 : 0) | (tensor ? 2 : 0) | (ctx->tile_runs ? 4 : 0),
                      rank, world, xworld, prm.compat_mode};
  if (ln.arena && !std::memcmp(sig, ln.plan_sig, sizeof(sig)) && ln.plan_Ns.size() == static_cast<size_t>(pairs) &&
      std::equal(Ns, Ns + pairs, ln.plan_Ns.begin()))
    return 0;
  ln.plan_sig[0] = -1;
  plan(Ns, pairs, prm, host, tensor, xworld, ln.descs, ln.tile_tab, ln.lay);
  if (tensor) {
    if (world > 1) {  // keep the tiles of this rank's cells (common.cuh: owner_of_cell; a tile never straddles two cells)
      std::vector<uint2>& tab = ln.tile_tab;
      size_t kept = 0;
      for (size_t k = 0; k < tab.size(); ++k) {
        const unsigned int jq = tab[k].y & 0xFFFFu, ib = tab[k].y >> 16;
        if (owner_of_cell(jq * static_cast<unsigned int>(kMmaTileN) / kOwnerCols, ib * static_cast<unsigned int>(kMmaTileM) / kOwnerRows,
                          static_cast<unsigned int>(world)) == static_cast<unsigned int>(rank))
          tab[kept++] = tab[k];
      }
      tab.resize(kept);
      ln.lay.total_tiles = static_cast<int>(kept);
    }
    if (ctx->tile_runs) interleave_tile_runs(ln.tile_tab, ln.tile_scratch, ln.descs, ctx->sm_count);
  }
  if (int rc = ensure_arena(ctx, ln, ln.lay.total_bytes)) return rc;
  if (int rc = ensure_keys(ctx, ln, ln.lay.key_guess)) return rc;
  bind(ln.lay, ln.arena, host, tensor);
  Layout& L = ln.lay;
  // uploads go through the lane's pinned staging buffer; the previous upload must have left it
  const size_t desc_bytes = sizeof(PairDesc) * static_cast<size_t>(pairs);
  const size_t tile_bytes = tensor ? sizeof(uint2) * ln.tile_tab.size() : 0;
  if (ln.h_stage) CU_TRY(cudaEventSynchronize(ln.uploaded));
  if (desc_bytes + tile_bytes > ln.h_stage_bytes) {
    if (ln.h_stage) CU_TRY(cudaFreeHost(ln.h_stage));
    ln.h_stage = nullptr;
    ln.h_stage_bytes = 0;
    const size_t want = align_up(desc_bytes + tile_bytes + (desc_bytes + tile_bytes) / 4, 4096);
    if (cudaMallocHost(&ln.h_stage, want) != cudaSuccess) {
      (void)cudaGetLastError();
      return SAC_COT_E_NOMEM;
    }
    ln.h_stage_bytes = want;
  }
  std::memcpy(ln.h_stage, ln.descs.data(), desc_bytes);
  CU_TRY(cudaMemcpyAsync(L.desc, ln.h_stage, desc_bytes, cudaMemcpyHostToDevice, ln.stream));
  if (tile_bytes) {
    std::memcpy(ln.h_stage + desc_bytes, ln.tile_tab.data(), tile_bytes);
    CU_TRY(cudaMemcpyAsync(L.tile_tab, ln.h_stage + desc_bytes, tile_bytes, cudaMemcpyHostToDevice, ln.stream));
  }
  CU_TRY(cudaEventRecord(ln.uploaded, ln.stream));
  std::memcpy(ln.plan_sig, sig, sizeof(sig));
  ln.plan_Ns.assign(Ns, Ns + pairs);
  return 0;
}

// The pairs a call works on: pair k of the call is pair first + k * stride of the caller's arrays (stride 1 = all of
// them, in order; stride G = the round-robin share of device `first` in a group of G).
struct PairSel {
  int first, stride, count;
  int id(int k) const { return first + k * stride; }
};

// Enqueues one chunk [k0,k1) of the selection on a lane.  offsets are absolute (whole call).
int enqueue_chunk(sac_cot_ctx* ctx, Lane& ln, const float* src, const float* dst, const int64_t* offsets, PairSel sel,
                  int k0, int k1, const sac_cot_params& prm, float* R, float* t, int32_t* inliers, bool host,
                  ChunkDev* h_chunk) {
  const int pairs = k1 - k0;
  std::vector<int32_t> Ns(pairs);
  for (int k = 0; k < pairs; ++k) Ns[k] = static_cast<int32_t>(offsets[sel.id(k0 + k) + 1] - offsets[sel.id(k0 + k)]);
  const bool tensor = ctx->tri_path != 0;
  if (int rc = prepare_lane(ctx, ln, Ns.data(), pairs, prm, host, tensor, 0, 1, 0)) return rc;
  Layout& L = ln.lay;
  const int id0 = sel.id(k0);
  const int64_t p0 = offsets[id0];
  if (host) {
    if (sel.stride == 1) {
      CU_TRY(cudaMemcpyAsync(L.in_src, src + 3 * p0, sizeof(float) * 3 * L.sum_n, cudaMemcpyHostToDevice, ln.stream));
      CU_TRY(cudaMemcpyAsync(L.in_dst, dst + 3 * p0, sizeof(float) * 3 * L.sum_n, cudaMemcpyHostToDevice, ln.stream));
    } else {
      // the chunk's pairs lie `stride` pairs apart in the caller's arrays; equal sizes (the usual batch) make that a
      // constant pitch: one 2-D copy per side gathers them into the chunk's contiguous input block
      bool uniform = true;
      const int64_t pitch = pairs > 1 ? offsets[sel.id(k0 + 1)] - p0 : 0;
      for (int k = 0; k < pairs && uniform; ++k)
        uniform = Ns[k] == Ns[0] && offsets[sel.id(k0 + k)] - p0 == pitch * k;
      if (uniform) {
        const size_t width = sizeof(float) * 3 * static_cast<size_t>(Ns[0]);
        const size_t spitch = pairs > 1 ? sizeof(float) * 3 * static_cast<size_t>(pitch) : width;
        CU_TRY(cudaMemcpy2DAsync(L.in_src, width, src + 3 * p0, spitch, width, pairs, cudaMemcpyHostToDevice, ln.stream));
        CU_TRY(cudaMemcpy2DAsync(L.in_dst, width, dst + 3 * p0, spitch, width, pairs, cudaMemcpyHostToDevice, ln.stream));
      } else {
        for (int k = 0; k < pairs; ++k) {
          const int64_t pk = offsets[sel.id(k0 + k)];
          const size_t bytes = sizeof(float) * 3 * static_cast<size_t>(Ns[k]);
          CU_TRY(cudaMemcpyAsync(L.in_src + 3 * ln.descs[k].pt_off, src + 3 * pk, bytes, cudaMemcpyHostToDevice, ln.stream));
          CU_TRY(cudaMemcpyAsync(L.in_dst + 3 * ln.descs[k].pt_off, dst + 3 * pk, bytes, cudaMemcpyHostToDevice, ln.stream));
        }
      }
    }
    if (int rc = enqueue_pipeline(ctx, ln, L.in_src, L.in_dst, L.outR, L.outT, L.outInl, 0, 1, false)) return rc;
    CU_TRY(cudaMemcpyAsync(h_chunk, L.chunk, sizeof(ChunkDev), cudaMemcpyDeviceToHost, ln.stream));
    if (sel.stride == 1) {
      CU_TRY(cudaMemcpyAsync(R + 9 * static_cast<size_t>(id0), L.outR, sizeof(float) * 9 * pairs, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpyAsync(t + 3 * static_cast<size_t>(id0), L.outT, sizeof(float) * 3 * pairs, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpyAsync(inliers + id0, L.outInl, sizeof(int32_t) * pairs, cudaMemcpyDeviceToHost, ln.stream));
    } else {  // scatter: results of pair k go to slot first + k * stride of the caller's arrays
      const size_t st = static_cast<size_t>(sel.stride);
      CU_TRY(cudaMemcpy2DAsync(R + 9 * static_cast<size_t>(id0), 36 * st, L.outR, 36, 36, pairs, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpy2DAsync(t + 3 * static_cast<size_t>(id0), 12 * st, L.outT, 12, 12, pairs, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpy2DAsync(inliers + id0, 4 * st, L.outInl, 4, 4, pairs, cudaMemcpyDeviceToHost, ln.stream));
    }
  } else {
    if (int rc = enqueue_pipeline(ctx, ln, src + 3 * p0, dst + 3 * p0, R + 9 * static_cast<size_t>(id0),
                                  t + 3 * static_cast<size_t>(id0), inliers + id0, 0, 1, false))
      return rc;
  }
  return 0;
}

int ensure_chunk_headers(sac_cot_ctx* ctx, size_t n) {
  while (ctx->h_chunks.size() < n) {
    ChunkDev* h = nullptr;
    CU_TRY(cudaMallocHost(&h, sizeof(ChunkDev)));
    ctx->h_chunks.push_back(h);
  }
  return 0;
}

// The selected pairs of a packed batch on one ctx (arguments already validated).
int run_selected(sac_cot_ctx* ctx, const float* src, const float* dst, const int64_t* offsets, PairSel sel,
                 const sac_cot_params* params, float* R, float* t, int32_t* inliers, int32_t location) {
  const int B = sel.count;
  if (B == 0) return SAC_COT_OK;
  if (location == SAC_COT_LOC_DEVICE && sel.stride != 1) return SAC_COT_E_UNSUPPORTED;
  CU_TRY(cudaSetDevice(ctx->device));
  if (int rc = resolve_pending(ctx, false)) return rc;
  const bool host = location == SAC_COT_LOC_HOST;
  ctx->np_try = ctx->node_prune >= 2 || ctx->np_skip_left == 0;
  if (!ctx->np_try) --ctx->np_skip_left;
  // chunking: bounded workspace per wave of kernels (keep_debug keeps the whole batch resident on
  // lane 0); chunks are dealt round-robin to the lanes
  const int K = params->num_edges * params->apex_per_edge;
  int chunk = B;
  int lanes = ctx->keep_debug ? 1 : ctx->n_lanes;
  if (!ctx->keep_debug) {
    if (ctx->chunk_pairs > 0) {
      chunk = std::min(B, ctx->chunk_pairs);
    } else {
      size_t total = 0;
      for (int k = 0; k < B; ++k)
        total += pair_bytes_estimate(static_cast<int>(offsets[sel.id(k) + 1] - offsets[sel.id(k)]), K, params->num_edges, ctx->tri_path != 0,
                                     params->compat_mode == SAC_COT_COMPAT_SECOND_ORDER);
      const size_t budget = static_cast<size_t>(3) << 30;  // ~3 GB of workspace per chunk (larger chunks measured faster)
      int nchunks = static_cast<int>((total + budget - 1) / budget);
      // give every lane something to overlap — but not chunks so small that their kernels under-fill the device:
      // measured at N = 5000, 32 pairs run 11 % faster as one chunk (1.77 ms) than as three of 11 (1.98 ms), 64 pairs
      // 1.5 % faster as two or three; a chunk of 32 such pairs is ~350 MB of workspace
      const int worth = static_cast<int>(std::min<size_t>(static_cast<size_t>(lanes), std::max<size_t>(1, total / (static_cast<size_t>(350) << 20))));
      if (nchunks < worth && B >= 2 * worth) nchunks = worth;
      chunk = (B + nchunks - 1) / std::max(1, nchunks);
    }
  }
  if (chunk > 32768) {  // pairs index blockIdx.y (<= 65535)
    if (ctx->keep_debug) return SAC_COT_E_SIZE;
    chunk = 32768;
  }
  const int nchunks = (B + chunk - 1) / chunk;
  lanes = std::min(lanes, nchunks);
  ctx->prm = *params;
  ctx->ws_valid = false;
  ctx->sh_valid = false;
  if (host)
    if (int rc = ensure_chunk_headers(ctx, static_cast<size_t>(nchunks))) return rc;

  std::vector<int> todo(nchunks);
  for (int c = 0; c < nchunks; ++c) todo[c] = c;
  if (!host)
    if (int rc = sticky_before(ctx)) return rc;
  for (int attempt = 0; attempt < 3 && !todo.empty(); ++attempt) {
    if (int rc = fork_lanes(ctx, lanes)) return rc;
    for (size_t k = 0; k < todo.size(); ++k) {
      const int c = todo[k];
      const int k0 = c * chunk, k1 = std::min(B, k0 + chunk);
      Lane& ln = ctx->lanes[k % lanes];
      if (int rc = enqueue_chunk(ctx, ln, src, dst, offsets, sel, k0, k1, *params, R, t, inliers, host,
                                 host ? ctx->h_chunks[c] : nullptr))
        return rc;
    }
    if (int rc = join_lanes(ctx, lanes)) return rc;
    if (!host) {
      // enqueue only; an overflow (if any) surfaces through resolve_pending()
      if (int rc = sticky_after(ctx)) return rc;
      ctx->ws_valid = nchunks == 1;
      return SAC_COT_OK;
    }
    if (ctx->np_tried && !ctx->sticky_pending)
      CU_TRY(cudaMemcpyAsync(&ctx->h_sticky[2], ctx->d_sticky, sizeof(StickyDev), cudaMemcpyDeviceToHost, ctx->stream));
    const bool np_read = ctx->np_tried && !ctx->sticky_pending;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (np_read) node_prune_verdict(ctx, ctx->h_sticky[2].pruned_total);
    // chunks whose edge count exceeded the lane's key pool: grow every lane to the measured demand
    // and re-run just those chunks
    std::vector<int> again;
    unsigned long long want = 0;
    for (int c : todo)
      if (ctx->h_chunks[c]->overflow) {
        again.push_back(c);
        want = std::max(want, ctx->h_chunks[c]->total_edges + ctx->h_chunks[c]->total_edges / 16 + 1024);
      }
    if (!again.empty()) {
      ++ctx->retries;
      for (int l = 0; l < lanes; ++l)
        if (int rc = ensure_keys(ctx, ctx->lanes[l], want)) return rc;
    }
    todo.swap(again);
  }
  if (!todo.empty()) return SAC_COT_E_NOMEM;
  ctx->ws_valid = nchunks == 1;
  return SAC_COT_OK;
}

int check_packed_args(const void* handle, const float* src, const float* dst, const int64_t* offsets, int32_t B,
                      const sac_cot_params* params, const float* R, const float* t, const int32_t* inliers, int32_t location) {
  if (!handle || !offsets) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST && location != SAC_COT_LOC_DEVICE) return SAC_COT_E_UNSUPPORTED;
  if (B < 0) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  if (B > 0 && (!src || !dst || !R || !t || !inliers)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t n = offsets[b + 1] - offsets[b];
    if (n < 3 || n > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  }
  return SAC_COT_OK;
}

int run_packed(sac_cot_ctx* ctx, const float* src, const float* dst, const int64_t* offsets, int32_t B,
               const sac_cot_params* params, float* R, float* t, int32_t* inliers, int32_t location) {
  if (int rc = check_packed_args(ctx, src, dst, offsets, B, params, R, t, inliers, location)) return rc;
  const sac_cot_params prm = normalized(params);
  return run_selected(ctx, src, dst, offsets, PairSel{0, 1, B}, &prm, R, t, inliers, location);
}

std::mutex g_mutex;
sac_cot_ctx* g_ctx = nullptr;

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int sac_cot_params_default(sac_cot_params* p) {
  if (!p) return SAC_COT_E_NULL;
  p->struct_size = sizeof(sac_cot_params);
  p->tau_compat = 0.1f;
  p->tau_inlier = 0.1f;
  p->num_edges = 1024;
  p->apex_per_edge = 4;
  p->score_mode = SAC_COT_SCORE_INLIER_COUNT;
  p->refit = 1;
  p->compat_mode = SAC_COT_COMPAT_FIRST_ORDER;
  p->so_min_common = 0;
  p->reserved = 0;
  return SAC_COT_OK;
}

int sac_cot_ctx_create(sac_cot_ctx** out, int32_t device, void* stream) {
  if (!out) return SAC_COT_E_NULL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    (void)cudaGetLastError();
    return SAC_COT_E_NODEVICE;
  }
  if (device < 0 || device >= count) return SAC_COT_E_NODEVICE;
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return SAC_COT_E_NODEVICE;  // built for sm_100a only
  sac_cot_ctx* ctx = new (std::nothrow) sac_cot_ctx();
  if (!ctx) return SAC_COT_E_NOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = static_cast<cudaStream_t>(stream);  // cudaStreamLegacy / cudaStreamPerThread handles work too
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return static_cast<int>(e); }
    ctx->own_stream = true;
  }
  cudaError_t e = cudaMallocHost(&ctx->h_sticky, 3 * sizeof(StickyDev));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->sticky_event, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_sticky, sizeof(StickyDev));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_sticky, 0, sizeof(StickyDev));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming);
  for (int l = 0; l < kMaxLanes && e == cudaSuccess; ++l) {
    e = cudaStreamCreateWithFlags(&ctx->lanes[l].stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].uploaded, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_xsum, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->match_uploaded, cudaEventDisableTiming);
  if (e != cudaSuccess) { sac_cot_ctx_destroy(ctx); return static_cast<int>(e); }
  std::memset(ctx->h_sticky, 0, 3 * sizeof(StickyDev));
  int rc = triangles_configure();
  if (rc >= 0) rc = triangles_mma_configure();
  if (rc >= 0) rc = select_configure();
  if (rc >= 0) rc = node_prune_configure();
  if (rc >= 0) rc = match_configure();
  if (rc < 0) { sac_cot_ctx_destroy(ctx); return -rc; }
  *out = ctx;
  return SAC_COT_OK;
}

int sac_cot_ctx_destroy(sac_cot_ctx* ctx) {
  if (!ctx) return SAC_COT_OK;
  cudaSetDevice(ctx->device);
  (void)sync_all(ctx);
  for (int l = 0; l < kMaxLanes; ++l) {
    Lane& ln = ctx->lanes[l];
    if (ln.arena) cudaFree(ln.arena);
    if (ln.keys) cudaFree(ln.keys);
    if (ln.h_stage) cudaFreeHost(ln.h_stage);
    if (ln.done) cudaEventDestroy(ln.done);
    if (ln.uploaded) cudaEventDestroy(ln.uploaded);
    if (ln.stream) cudaStreamDestroy(ln.stream);
  }
  if (ctx->comm && ctx->own_comm)
    if (NcclApi* nc = nccl_api()) nc->CommDestroy(ctx->comm);
  if (ctx->h_xsum) cudaFreeHost(ctx->h_xsum);
  if (ctx->match_arena) cudaFree(ctx->match_arena);
  if (ctx->h_match) cudaFreeHost(ctx->h_match);
  if (ctx->match_uploaded) cudaEventDestroy(ctx->match_uploaded);
  for (ChunkDev* h : ctx->h_chunks) cudaFreeHost(h);
  if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
  if (ctx->h_sticky) cudaFreeHost(ctx->h_sticky);
  if (ctx->sticky_event) cudaEventDestroy(ctx->sticky_event);
  if (ctx->d_sticky) cudaFree(ctx->d_sticky);
  ctx->timer.destroy();
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SAC_COT_OK;
}

int sac_cot_ctx_set(sac_cot_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return SAC_COT_E_NULL;
  if (!std::strcmp(name, "keep_debug")) { ctx->keep_debug = value != 0; return SAC_COT_OK; }
  if (!std::strcmp(name, "chunk_pairs")) {
    if (value < 0) return SAC_COT_E_SIZE;
    ctx->chunk_pairs = static_cast<int>(std::min<int64_t>(value, 65535));
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "lanes")) {
    if (value < 1 || value > kMaxLanes) return SAC_COT_E_SIZE;
    ctx->n_lanes = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "triangle_path")) {
    if (value < 0 || value > 2) return SAC_COT_E_UNSUPPORTED;
    ctx->tri_path = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "triangle_prune")) {
    ctx->tri_prune = value != 0;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "apex_path")) {
    if (value < 0 || value > 2) return SAC_COT_E_UNSUPPORTED;
    ctx->apex_path = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune")) {
    if (value < 0 || value > 2) return SAC_COT_E_UNSUPPORTED;
    ctx->node_prune = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune_cost")) {
    if (value < 1 || value > 1000000) return SAC_COT_E_SIZE;
    ctx->node_prune_cost = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune_rect")) {
    if (value < 0 || value > 2) return SAC_COT_E_UNSUPPORTED;
    ctx->node_prune_rect = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune_probe")) {
    if (value < 0 || value > 1000000) return SAC_COT_E_SIZE;
    ctx->node_prune_probe = static_cast<int>(value);
    ctx->np_skip_left = 0;
    ctx->np_idle = 0;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "tile_runs")) {
    ctx->tile_runs = value != 0;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "match_dbg")) { ctx->match_dbg = static_cast<int>(value); return SAC_COT_OK; }
  if (!std::strcmp(name, "match_path")) {
    if (value < 0 || value > 1) return SAC_COT_E_UNSUPPORTED;
    ctx->match_path = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "triangle_dbg")) {
    ctx->tri_dbg = static_cast<int>(value);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "stage_timing")) {
    cudaSetDevice(ctx->device);
    ctx->timer.reset();
    ctx->timer.enabled = value != 0;
    return SAC_COT_OK;
  }
  return SAC_COT_E_WHICH;
}

int sac_cot_ctx_get(sac_cot_ctx* ctx, const char* name, int64_t* value) {
  if (!ctx || !name || !value) return SAC_COT_E_NULL;
  if (!std::strcmp(name, "launches")) { *value = ctx->launches; return SAC_COT_OK; }
  if (!std::strcmp(name, "workspace_bytes")) {
    int64_t total = 0;
    for (int l = 0; l < kMaxLanes; ++l)
      total += static_cast<int64_t>(ctx->lanes[l].arena_bytes + ctx->lanes[l].key_cap * sizeof(unsigned long long));
    *value = total;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "retries")) { *value = ctx->retries; return SAC_COT_OK; }
  if (!std::strcmp(name, "device")) { *value = ctx->device; return SAC_COT_OK; }
  if (!std::strcmp(name, "sm_count")) { *value = ctx->sm_count; return SAC_COT_OK; }
  if (!std::strcmp(name, "lanes")) { *value = ctx->n_lanes; return SAC_COT_OK; }
  if (!std::strcmp(name, "triangle_path")) { *value = ctx->tri_path; return SAC_COT_OK; }
  if (!std::strcmp(name, "triangle_prune")) { *value = ctx->tri_prune; return SAC_COT_OK; }
  if (!std::strcmp(name, "match_path")) { *value = ctx->match_path; return SAC_COT_OK; }
  if (!std::strcmp(name, "triangle_path_used")) {
    // which S2 kernels the most recent chunk on lane 0 ran (0 = POPC, 1 = tensor core); synchronises
    Lane& ln = ctx->lanes[0];
    if (!ln.arena || !ln.lay.chunk || ln.lay.pairs == 0) return SAC_COT_E_WHICH;
    cudaSetDevice(ctx->device);
    if (int rc = sync_all(ctx)) return rc;
    ChunkDev c;
    CU_TRY(cudaMemcpy(&c, ln.lay.chunk, sizeof(c), cudaMemcpyDeviceToHost));
    *value = c.use_tensor;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune")) { *value = ctx->node_prune; return SAC_COT_OK; }
  if (!std::strcmp(name, "node_prune_cost")) { *value = ctx->node_prune_cost; return SAC_COT_OK; }
  if (!std::strcmp(name, "node_prune_probe")) { *value = ctx->node_prune_probe; return SAC_COT_OK; }
  if (!std::strcmp(name, "node_prune_rect")) { *value = ctx->node_prune_rect; return SAC_COT_OK; }
  if (!std::strcmp(name, "rect_pairs")) {
    // pairs of the most recent chunk on lane 0 whose kept rows went through the RECT instance; synchronises
    Lane& ln = ctx->lanes[0];
    if (!ln.arena || !ln.lay.nplan || ln.lay.pairs == 0) return SAC_COT_E_WHICH;
    cudaSetDevice(ctx->device);
    if (int rc = sync_all(ctx)) return rc;
    std::vector<NodePlan> pl(ln.lay.pairs);
    CU_TRY(cudaMemcpy(pl.data(), ln.lay.nplan, sizeof(NodePlan) * pl.size(), cudaMemcpyDeviceToHost));
    int64_t n = 0;
    for (const NodePlan& q : pl) n += q.pruned == 2u ? 1 : 0;
    *value = n;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "node_prune_trying")) { *value = ctx->np_skip_left == 0 ? 1 : 0; return SAC_COT_OK; }
  if (!std::strcmp(name, "pruned_pairs") || !std::strcmp(name, "kept_nodes")) {
    // exact node pruning in the most recent chunk on lane 0: pairs that took the kept-row kernel / the nodes they
    // kept in total; synchronises
    Lane& ln = ctx->lanes[0];
    if (!ln.arena || !ln.lay.nplan || ln.lay.pairs == 0) return SAC_COT_E_WHICH;
    cudaSetDevice(ctx->device);
    if (int rc = sync_all(ctx)) return rc;
    std::vector<NodePlan> pl(ln.lay.pairs);
    CU_TRY(cudaMemcpy(pl.data(), ln.lay.nplan, sizeof(NodePlan) * pl.size(), cudaMemcpyDeviceToHost));
    int64_t np = 0, nk = 0;
    for (const NodePlan& q : pl)
      if (q.pruned) {
        ++np;
        nk += q.n_keep;
      }
    *value = name[0] == 'p' ? np : nk;
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "threads")) { *value = 0; return SAC_COT_OK; }
  if (!std::strcmp(name, "comm_rank")) { *value = ctx->comm_rank; return SAC_COT_OK; }
  if (!std::strcmp(name, "comm_world")) { *value = ctx->comm_world; return SAC_COT_OK; }
  if (!std::strncmp(name, "stage_us_", 9) || !std::strncmp(name, "stage_calls_", 12)) {
    const bool us = name[6] == 'u';
    const char* stage = name + (us ? 9 : 12);
    for (int k = 0; k < ST_COUNT; ++k)
      if (!std::strcmp(stage, kStageNames[k])) {
        cudaSetDevice(ctx->device);
        ctx->timer.resolve();
        *value = us ? static_cast<int64_t>(ctx->timer.acc_ms[k] * 1000.0 + 0.5) : ctx->timer.calls[k];
        return SAC_COT_OK;
      }
    return SAC_COT_E_WHICH;
  }
  if (!std::strcmp(name, "probe_mxf4_gflops")) {
    // dense rate of the triangle kernel's MMA (mxf4, cta_group::2, M 256 x N 240 x K 64) with nothing else
    // running: three timed launches on every CTA pair of the device, best of three; synchronises
    CU_TRY(cudaSetDevice(ctx->device));
    if (int rc = sync_all(ctx)) return rc;
    cudaEvent_t a = nullptr, b = nullptr;
    CU_TRY(cudaEventCreate(&a));
    cudaError_t ce = cudaEventCreate(&b);
    if (ce != cudaSuccess) { cudaEventDestroy(a); return static_cast<int>(ce); }
    LaunchCtx lc{ctx->stream, ctx->sm_count};
    const int clusters = ctx->sm_count / 2, stage_pairs = 2048;
    double best_ms = 0.0;
    int rc = 0;
    for (int rep = 0; rep < 4 && rc == 0; ++rep) {  // rep 0 warms up
      cudaEventRecord(a, ctx->stream);
      const int n = launch_mma_peak_probe(lc, clusters, stage_pairs);
      if (n < 0) { rc = -n; break; }
      ctx->launches += n;
      cudaEventRecord(b, ctx->stream);
      ce = cudaEventSynchronize(b);
      if (ce != cudaSuccess) { rc = static_cast<int>(ce); break; }
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && (best_ms == 0.0 || ms < best_ms)) best_ms = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (rc) return rc;
    *value = static_cast<int64_t>(mma_peak_probe_flops(clusters, stage_pairs) / (best_ms * 1e-3) / 1e9);
    return SAC_COT_OK;
  }
  if (!std::strcmp(name, "last_status")) {
    // status of the most recent device-location call; synchronises with it
    cudaSetDevice(ctx->device);
    const int rc = resolve_pending(ctx, true);
    *value = rc ? rc : ctx->deferred_status;
    ctx->deferred_status = 0;
    return SAC_COT_OK;
  }
  return SAC_COT_E_WHICH;
}

int sac_cot_register_packed(sac_cot_ctx* ctx, const float* src, const float* dst, const int64_t* offsets, int32_t B,
                            const sac_cot_params* params, float* R, float* t, int32_t* inliers, int32_t location) {
  try {
    return run_packed(ctx, src, dst, offsets, B, params, R, t, inliers, location);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

int sac_cot_register_batch(sac_cot_ctx* ctx, const float* const* src, const float* const* dst, const int32_t* N,
                           int32_t B, const sac_cot_params* params, float* R, float* t, int32_t* inliers) {
  if (!ctx) return SAC_COT_E_NULL;
  if (B < 0) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  if (B > 0 && (!src || !dst || !N || !R || !t || !inliers)) return SAC_COT_E_NULL;
  try {
    std::vector<int64_t> offsets(static_cast<size_t>(B) + 1, 0);
    for (int b = 0; b < B; ++b) {
      if (!src[b] || !dst[b]) return SAC_COT_E_NULL;
      if (N[b] < 3 || N[b] > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
      offsets[b + 1] = offsets[b] + N[b];
    }
    ctx->stage_src.resize(static_cast<size_t>(offsets[B]) * 3);
    ctx->stage_dst.resize(static_cast<size_t>(offsets[B]) * 3);
    for (int b = 0; b < B; ++b) {
      std::memcpy(&ctx->stage_src[static_cast<size_t>(offsets[b]) * 3], src[b], sizeof(float) * 3 * N[b]);
      std::memcpy(&ctx->stage_dst[static_cast<size_t>(offsets[b]) * 3], dst[b], sizeof(float) * 3 * N[b]);
    }
    return run_packed(ctx, ctx->stage_src.data(), ctx->stage_dst.data(), offsets.data(), B, params, R, t, inliers,
                      SAC_COT_LOC_HOST);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

int sac_cot_register(const float* src, const float* dst, int32_t N, const sac_cot_params* params, float R[9],
                     float t[3], int32_t* inliers) {
  if (!src || !dst || !R || !t || !inliers) return SAC_COT_E_NULL;
  if (N < 3 || N > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g_ctx) {
    const int rc = sac_cot_ctx_create(&g_ctx, 0, nullptr);
    if (rc) return rc;
  }
  const int64_t offsets[2] = {0, N};
  return sac_cot_register_packed(g_ctx, src, dst, offsets, 1, params, R, t, inliers, SAC_COT_LOC_HOST);
}

// ---- correspondence front end: descriptor nearest-neighbour matching -------
}  // extern "C"

namespace {

int run_match(sac_cot_ctx* ctx, const float* desc_src, const float* xyz_src, const int64_t* offs_src, const float* desc_dst,
              const float* xyz_dst, const int64_t* offs_dst, int32_t B, int32_t dim, int32_t* nn, float* corr_src,
              float* corr_dst, int32_t location) {
  if (!ctx || !offs_src || !offs_dst) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST && location != SAC_COT_LOC_DEVICE) return SAC_COT_E_UNSUPPORTED;
  if (B < 0 || B > 65535) return SAC_COT_E_SIZE;
  if (dim < 1 || dim > SAC_COT_MAX_DESC_DIM) return SAC_COT_E_SIZE;
  if (B > 0 && (!desc_src || !xyz_src || !desc_dst || !xyz_dst || !nn || !corr_src || !corr_dst)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t ns = offs_src[b + 1] - offs_src[b], nd = offs_dst[b + 1] - offs_dst[b];
    if (ns < 1 || nd < 1 || ns > SAC_COT_MAX_KEYPOINTS || nd > SAC_COT_MAX_KEYPOINTS) return SAC_COT_E_SIZE;
  }
  if (B == 0) return SAC_COT_OK;
  CU_TRY(cudaSetDevice(ctx->device));
  const bool host = location == SAC_COT_LOC_HOST;
  const bool tensor = ctx->match_path == 1 && dim <= kMatchMaxDim;
  const int chunks = match_chunks(dim);
  const int64_t tot_s = offs_src[B] - offs_src[0], tot_d = offs_dst[B] - offs_dst[0];

  // ---- pair table + workspace layout (offsets relative to the arena) ----
  std::vector<MatchPair> tab(static_cast<size_t>(B));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_tab = take(sizeof(MatchPair) * B);
  const size_t o_bmax = take(sizeof(uint32_t) * B);
  size_t norm_rows = 0, img_bytes = 0;
  int max_s_tiles = 0, max_d_tiles = 0, max_ns = 0;
  for (int b = 0; b < B; ++b) {
    MatchPair& m = tab[b];
    m.Ns = static_cast<int32_t>(offs_src[b + 1] - offs_src[b]);
    m.Nd = static_cast<int32_t>(offs_dst[b + 1] - offs_dst[b]);
    m.s_tiles = (m.Ns + kMatchTileM - 1) / kMatchTileM;
    m.d_tiles = (m.Nd + kMatchTileN - 1) / kMatchTileN;
    m.s_off = offs_src[b] - offs_src[0];
    m.d_off = offs_dst[b] - offs_dst[0];
    m.s_norm = static_cast<int64_t>(norm_rows);
    m.cand_off = m.s_off;
    m.s_img = static_cast<int64_t>(img_bytes);
    img_bytes += static_cast<size_t>(m.s_tiles) * kMatchTileM * chunks * 16;
    m.d_img = static_cast<int64_t>(img_bytes);
    img_bytes += static_cast<size_t>(m.d_tiles) * kMatchTileN * chunks * 16;
    norm_rows += static_cast<size_t>(m.s_tiles) * kMatchTileM;
    max_s_tiles = std::max(max_s_tiles, m.s_tiles);
    max_d_tiles = std::max(max_d_tiles, m.d_tiles);
    max_ns = std::max(max_ns, m.Ns);
  }
  const size_t o_norm = tensor ? take(sizeof(float) * norm_rows) : 0;
  const size_t o_img = tensor ? take(img_bytes) : 0;
  const size_t o_cand = tensor ? take(sizeof(int32_t) * kMatchUnion * tot_s) : 0;
  const size_t o_cnt = tensor ? take(sizeof(int32_t) * tot_s) : 0;
  const size_t o_work = tensor ? take(sizeof(int32_t) * 2 * tot_s + 16) : 0;  // work list of the scan kernel + its counter
  const size_t o_ds = host ? take(sizeof(float) * dim * tot_s) : 0;
  const size_t o_dd = host ? take(sizeof(float) * dim * tot_d) : 0;
  const size_t o_xs = host ? take(sizeof(float) * 3 * tot_s) : 0;
  const size_t o_xd = host ? take(sizeof(float) * 3 * tot_d) : 0;
  const size_t o_nn = host ? take(sizeof(int32_t) * tot_s) : 0;
  const size_t o_cs = host ? take(sizeof(float) * 3 * tot_s) : 0;
  const size_t o_cd = host ? take(sizeof(float) * 3 * tot_s) : 0;
  if (off > ctx->match_arena_bytes) {
    if (int rc = sync_all(ctx)) return rc;
    if (ctx->match_arena) CU_TRY(cudaFree(ctx->match_arena));
    ctx->match_arena = nullptr;
    ctx->match_arena_bytes = 0;
    ctx->match_sig.clear();
    const size_t want = align_up(off + off / 8, 1 << 20);
    if (cudaMalloc(&ctx->match_arena, want) != cudaSuccess) {
      (void)cudaGetLastError();
      return SAC_COT_E_NOMEM;
    }
    ctx->match_arena_bytes = want;
  }
  unsigned char* base = ctx->match_arena;
  // images hold offsets relative to the image block
  MatchPair* d_tab = reinterpret_cast<MatchPair*>(base + o_tab);
  uint32_t* d_bmax = reinterpret_cast<uint32_t*>(base + o_bmax);
  float* d_norm = reinterpret_cast<float*>(base + o_norm);
  unsigned char* d_img = base + o_img;
  int32_t* d_cand = tensor ? reinterpret_cast<int32_t*>(base + o_cand) : nullptr;
  int32_t* d_cnt = tensor ? reinterpret_cast<int32_t*>(base + o_cnt) : nullptr;
  cudaStream_t stream = ctx->stream;
  // ---- pair table upload (skipped when the shapes repeat) ----
  std::vector<int64_t> sig;
  sig.reserve(2 * static_cast<size_t>(B) + 4);
  sig.push_back(B);
  sig.push_back(dim);
  sig.push_back((host ? 1 : 0) | (tensor ? 2 : 0));
  for (int b = 0; b < B; ++b) {
    sig.push_back(tab[b].Ns);
    sig.push_back(tab[b].Nd);
  }
  if (sig != ctx->match_sig) {
    const size_t bytes = sizeof(MatchPair) * B;
    CU_TRY(cudaEventSynchronize(ctx->match_uploaded));
    if (bytes > ctx->h_match_bytes) {
      if (ctx->h_match) CU_TRY(cudaFreeHost(ctx->h_match));
      ctx->h_match = nullptr;
      ctx->h_match_bytes = 0;
      if (cudaMallocHost(&ctx->h_match, align_up(bytes * 2, 4096)) != cudaSuccess) {
        (void)cudaGetLastError();
        return SAC_COT_E_NOMEM;
      }
      ctx->h_match_bytes = align_up(bytes * 2, 4096);
    }
    std::memcpy(ctx->h_match, tab.data(), bytes);
    CU_TRY(cudaMemcpyAsync(d_tab, ctx->h_match, bytes, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaEventRecord(ctx->match_uploaded, stream));
    ctx->match_sig.swap(sig);
  }
  const float* ds = desc_src + static_cast<size_t>(offs_src[0]) * dim;
  const float* dd = desc_dst + static_cast<size_t>(offs_dst[0]) * dim;
  const float* xs = xyz_src + static_cast<size_t>(offs_src[0]) * 3;
  const float* xd = xyz_dst + static_cast<size_t>(offs_dst[0]) * 3;
  int32_t* o_nn_p = nn + offs_src[0];
  float* o_cs_p = corr_src + static_cast<size_t>(offs_src[0]) * 3;
  float* o_cd_p = corr_dst + static_cast<size_t>(offs_src[0]) * 3;
  if (host) {
    CU_TRY(cudaMemcpyAsync(base + o_ds, ds, sizeof(float) * dim * tot_s, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_dd, dd, sizeof(float) * dim * tot_d, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_xs, xs, sizeof(float) * 3 * tot_s, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_xd, xd, sizeof(float) * 3 * tot_d, cudaMemcpyHostToDevice, stream));
    ds = reinterpret_cast<const float*>(base + o_ds);
    dd = reinterpret_cast<const float*>(base + o_dd);
    xs = reinterpret_cast<const float*>(base + o_xs);
    xd = reinterpret_cast<const float*>(base + o_xd);
  }
  int32_t* k_nn = host ? reinterpret_cast<int32_t*>(base + o_nn) : o_nn_p;
  float* k_cs = host ? reinterpret_cast<float*>(base + o_cs) : o_cs_p;
  float* k_cd = host ? reinterpret_cast<float*>(base + o_cd) : o_cd_p;
  LaunchCtx lc{stream, ctx->sm_count};
  StageTimer& tm = ctx->timer;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto span_begin = [&]() {
    if (tm.enabled && (e0 = tm.next())) cudaEventRecord(e0, stream);
  };
  auto span_end = [&](int stage) {
    if (tm.enabled && e0 && (e1 = tm.next())) {
      cudaEventRecord(e1, stream);
      tm.pending.push_back({stage, e0, e1});
    }
    e0 = nullptr;
  };
  if (tensor) {
    span_begin();
    CU_TRY(cudaMemsetAsync(d_bmax, 0, sizeof(uint32_t) * B, stream));
    KL_TRY(launch_match_prep(lc, d_tab, B, max_s_tiles * kMatchTileM, ds, 0, dim, d_img, d_norm, d_bmax));
    KL_TRY(launch_match_prep(lc, d_tab, B, max_d_tiles * kMatchTileN, dd, 1, dim, d_img, d_norm, d_bmax));
    span_end(ST_MATCH_PREP);
    span_begin();
    KL_TRY(launch_match_mma(lc, d_tab, B, max_s_tiles, dim, d_img, d_norm, d_bmax, d_cand, d_cnt, ctx->match_dbg));
    span_end(ST_MATCH_SWEEP);
  }
  span_begin();
  uint32_t* d_work_count = tensor ? reinterpret_cast<uint32_t*>(base + o_work) : nullptr;
  int32_t* d_work_list = tensor ? reinterpret_cast<int32_t*>(base + o_work + 16) : nullptr;
  KL_TRY(launch_match_exact(lc, d_tab, B, max_ns, ds, dd, xs, xd, dim, d_cand, d_cnt, k_nn, k_cs, k_cd, d_work_list, d_work_count));
  span_end(ST_MATCH_EXACT);
  if (host) {
    CU_TRY(cudaMemcpyAsync(o_nn_p, k_nn, sizeof(int32_t) * tot_s, cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaMemcpyAsync(o_cs_p, k_cs, sizeof(float) * 3 * tot_s, cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaMemcpyAsync(o_cd_p, k_cd, sizeof(float) * 3 * tot_s, cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
  }
  return SAC_COT_OK;
}

int run_match_mutual(sac_cot_ctx* ctx, const int32_t* nn, const int32_t* nn_back, const float* corr_src, const float* corr_dst,
                     const int64_t* offs_src, const int64_t* offs_dst, int32_t B, float* out_src, float* out_dst,
                     int64_t* out_offsets, int32_t location) {
  if (!ctx || !offs_src || !offs_dst || !out_offsets) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST && location != SAC_COT_LOC_DEVICE) return SAC_COT_E_UNSUPPORTED;
  if (B < 0 || B > 65535) return SAC_COT_E_SIZE;
  if (B > 0 && (!nn || !nn_back || !corr_src || !corr_dst || !out_src || !out_dst)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t ns = offs_src[b + 1] - offs_src[b], nd = offs_dst[b + 1] - offs_dst[b];
    if (ns < 1 || nd < 1 || ns > SAC_COT_MAX_KEYPOINTS || nd > SAC_COT_MAX_KEYPOINTS) return SAC_COT_E_SIZE;
  }
  const bool host = location == SAC_COT_LOC_HOST;
  if (B == 0) {
    if (host) out_offsets[0] = 0;
    return SAC_COT_OK;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  const int64_t tot_s = offs_src[B] - offs_src[0], tot_d = offs_dst[B] - offs_dst[0];
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_tab = take(sizeof(MutualPair) * B);
  const size_t o_kept = take(sizeof(unsigned long long) * B);
  const size_t o_offs = take(sizeof(long long) * (static_cast<size_t>(B) + 1));
  const size_t o_nn = host ? take(sizeof(int32_t) * tot_s) : 0;
  const size_t o_nb = host ? take(sizeof(int32_t) * tot_d) : 0;
  const size_t o_cs = host ? take(sizeof(float) * 3 * tot_s) : 0;
  const size_t o_cd = host ? take(sizeof(float) * 3 * tot_s) : 0;
  const size_t o_os = host ? take(sizeof(float) * 3 * tot_s) : 0;
  const size_t o_od = host ? take(sizeof(float) * 3 * tot_s) : 0;
  if (off > ctx->match_arena_bytes) {
    if (int rc = sync_all(ctx)) return rc;
    if (ctx->match_arena) CU_TRY(cudaFree(ctx->match_arena));
    ctx->match_arena = nullptr;
    ctx->match_arena_bytes = 0;
    const size_t want = align_up(off + off / 8, 1 << 20);
    if (cudaMalloc(&ctx->match_arena, want) != cudaSuccess) {
      (void)cudaGetLastError();
      return SAC_COT_E_NOMEM;
    }
    ctx->match_arena_bytes = want;
  }
  ctx->match_sig.clear();  // the arena no longer holds the matching call's pair table
  unsigned char* base = ctx->match_arena;
  cudaStream_t stream = ctx->stream;
  std::vector<MutualPair> tab(static_cast<size_t>(B));
  for (int b = 0; b < B; ++b) {
    tab[b].Ns = static_cast<int32_t>(offs_src[b + 1] - offs_src[b]);
    tab[b].Nd = static_cast<int32_t>(offs_dst[b + 1] - offs_dst[b]);
    tab[b].s_off = offs_src[b] - offs_src[0];
    tab[b].d_off = offs_dst[b] - offs_dst[0];
  }
  const size_t tab_bytes = sizeof(MutualPair) * B;
  CU_TRY(cudaEventSynchronize(ctx->match_uploaded));
  if (tab_bytes > ctx->h_match_bytes) {
    if (ctx->h_match) CU_TRY(cudaFreeHost(ctx->h_match));
    ctx->h_match = nullptr;
    ctx->h_match_bytes = 0;
    if (cudaMallocHost(&ctx->h_match, align_up(tab_bytes * 2, 4096)) != cudaSuccess) {
      (void)cudaGetLastError();
      return SAC_COT_E_NOMEM;
    }
    ctx->h_match_bytes = align_up(tab_bytes * 2, 4096);
  }
  std::memcpy(ctx->h_match, tab.data(), tab_bytes);
  CU_TRY(cudaMemcpyAsync(base + o_tab, ctx->h_match, tab_bytes, cudaMemcpyHostToDevice, stream));
  CU_TRY(cudaEventRecord(ctx->match_uploaded, stream));
  const int32_t* k_nn = nn + offs_src[0];
  const int32_t* k_nb = nn_back + offs_dst[0];
  const float* k_cs = corr_src + static_cast<size_t>(offs_src[0]) * 3;
  const float* k_cd = corr_dst + static_cast<size_t>(offs_src[0]) * 3;
  float* k_os = out_src;
  float* k_od = out_dst;
  long long* k_off = reinterpret_cast<long long*>(out_offsets);
  if (host) {
    CU_TRY(cudaMemcpyAsync(base + o_nn, k_nn, sizeof(int32_t) * tot_s, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_nb, k_nb, sizeof(int32_t) * tot_d, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_cs, k_cs, sizeof(float) * 3 * tot_s, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(base + o_cd, k_cd, sizeof(float) * 3 * tot_s, cudaMemcpyHostToDevice, stream));
    k_nn = reinterpret_cast<const int32_t*>(base + o_nn);
    k_nb = reinterpret_cast<const int32_t*>(base + o_nb);
    k_cs = reinterpret_cast<const float*>(base + o_cs);
    k_cd = reinterpret_cast<const float*>(base + o_cd);
    k_os = reinterpret_cast<float*>(base + o_os);
    k_od = reinterpret_cast<float*>(base + o_od);
    k_off = reinterpret_cast<long long*>(base + o_offs);
  }
  LaunchCtx lc{stream, ctx->sm_count};
  KL_TRY(launch_match_mutual(lc, reinterpret_cast<const MutualPair*>(base + o_tab), B, k_nn, k_nb, k_cs, k_cd,
                             reinterpret_cast<unsigned long long*>(base + o_kept), k_off, k_os, k_od));
  if (host) {
    CU_TRY(cudaMemcpyAsync(out_offsets, k_off, sizeof(long long) * (static_cast<size_t>(B) + 1), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    const size_t kept = static_cast<size_t>(out_offsets[B]);
    if (kept) {
      CU_TRY(cudaMemcpyAsync(out_src, k_os, sizeof(float) * 3 * kept, cudaMemcpyDeviceToHost, stream));
      CU_TRY(cudaMemcpyAsync(out_dst, k_od, sizeof(float) * 3 * kept, cudaMemcpyDeviceToHost, stream));
      CU_TRY(cudaStreamSynchronize(stream));
    }
  }
  return SAC_COT_OK;
}

}  // namespace

extern "C" {

int sac_cot_match_mutual(sac_cot_ctx* ctx, const int32_t* nn, const int32_t* nn_back, const float* corr_src,
                         const float* corr_dst, const int64_t* offs_src, const int64_t* offs_dst, int32_t B, float* out_src,
                         float* out_dst, int64_t* out_offsets, int32_t location) {
  try {
    return run_match_mutual(ctx, nn, nn_back, corr_src, corr_dst, offs_src, offs_dst, B, out_src, out_dst, out_offsets, location);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

int sac_cot_match_packed(sac_cot_ctx* ctx, const float* desc_src, const float* xyz_src, const int64_t* offs_src,
                         const float* desc_dst, const float* xyz_dst, const int64_t* offs_dst, int32_t B, int32_t dim,
                         int32_t* nn, float* corr_src, float* corr_dst, int32_t location) {
  try {
    return run_match(ctx, desc_src, xyz_src, offs_src, desc_dst, xyz_dst, offs_dst, B, dim, nn, corr_src, corr_dst, location);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

int sac_cot_match(const float* desc_src, const float* xyz_src, int32_t Ns, const float* desc_dst, const float* xyz_dst,
                  int32_t Nd, int32_t dim, int32_t* nn, float* corr_src, float* corr_dst) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g_ctx) {
    const int rc = sac_cot_ctx_create(&g_ctx, 0, nullptr);
    if (rc) return rc;
  }
  const int64_t os[2] = {0, Ns}, od[2] = {0, Nd};
  return sac_cot_match_packed(g_ctx, desc_src, xyz_src, os, desc_dst, xyz_dst, od, 1, dim, nn, corr_src, corr_dst, SAC_COT_LOC_HOST);
}

// ---- device groups: one batch over several GPUs of the box, no communication ------------------
}  // extern "C"

// One worker thread per member device: enqueueing a chunk costs the host ~0.1 ms and a step of the headline batch on
// eight GPUs lasts ~1.5 ms, so a single enqueueing thread would be the bottleneck.  Workers sleep on a condition
// variable between calls.
struct sac_cot_group {
  std::vector<sac_cot_ctx*> ctxs;
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  // epoch / pending are also polled without the mutex: a worker spins for a short while after finishing (a caller in
  // a loop finds it awake), the caller spins while the devices work; both fall back to the condition variables
  std::atomic<uint64_t> epoch{0};
  std::atomic<int> pending{0};
  bool stop = false;
  // the call in flight
  const float* src = nullptr;
  const float* dst = nullptr;
  const int64_t* offsets = nullptr;
  int32_t B = 0;
  const sac_cot_params* params = nullptr;
  sac_cot_params prm{};  // the caller's params, normalised to the current struct version
  float* R = nullptr;
  float* t = nullptr;
  int32_t* inliers = nullptr;
  std::vector<int> rc;
};

namespace {
void group_worker(sac_cot_group* g, int idx) {
  uint64_t seen = 0;
  const int G = static_cast<int>(g->ctxs.size());
  for (;;) {
    {
      // ~200 us of polling before sleeping: waking through the condition variable costs tens of microseconds, a
      // step of the headline batch on eight GPUs lasts two milliseconds
      const auto t0 = std::chrono::steady_clock::now();
      while (g->epoch.load(std::memory_order_acquire) == seen &&
             std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      std::unique_lock<std::mutex> lk(g->m);
      g->cv_go.wait(lk, [&] { return g->stop || g->epoch.load() != seen; });
      if (g->stop) return;
      seen = g->epoch.load();
    }
    int rc;
    try {
      const int count = g->B > idx ? (g->B - idx + G - 1) / G : 0;
      rc = run_selected(g->ctxs[idx], g->src, g->dst, g->offsets, PairSel{idx, G, count}, g->params, g->R, g->t, g->inliers,
                        SAC_COT_LOC_HOST);
    } catch (const std::bad_alloc&) {
      rc = SAC_COT_E_NOMEM;
    }
    {
      std::lock_guard<std::mutex> lk(g->m);
      g->rc[idx] = rc;
      if (g->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) g->cv_done.notify_all();
    }
  }
}
}  // namespace

extern "C" {

int sac_cot_group_destroy(sac_cot_group* g) {
  if (!g) return SAC_COT_OK;
  {
    std::lock_guard<std::mutex> lk(g->m);
    g->stop = true;
  }
  g->cv_go.notify_all();
  for (std::thread& th : g->workers)
    if (th.joinable()) th.join();
  for (sac_cot_ctx* c : g->ctxs) sac_cot_ctx_destroy(c);
  delete g;
  return SAC_COT_OK;
}

int sac_cot_group_create(sac_cot_group** out, const int32_t* devices, int32_t n_devices) {
  if (!out || !devices) return SAC_COT_E_NULL;
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) return SAC_COT_E_SIZE;
  sac_cot_group* g = new (std::nothrow) sac_cot_group();
  if (!g) return SAC_COT_E_NOMEM;
  try {
    for (int k = 0; k < n_devices; ++k) {
      sac_cot_ctx* c = nullptr;
      const int rc = sac_cot_ctx_create(&c, devices[k], nullptr);
      if (rc) {
        sac_cot_group_destroy(g);
        return rc;
      }
      g->ctxs.push_back(c);
    }
    g->rc.assign(static_cast<size_t>(n_devices), 0);
    for (int k = 0; k < n_devices; ++k) g->workers.emplace_back(group_worker, g, k);
  } catch (...) {
    sac_cot_group_destroy(g);
    return SAC_COT_E_NOMEM;
  }
  *out = g;
  return SAC_COT_OK;
}

int32_t sac_cot_group_size(const sac_cot_group* g) { return g ? static_cast<int32_t>(g->ctxs.size()) : 0; }

sac_cot_ctx* sac_cot_group_ctx(sac_cot_group* g, int32_t index) {
  return (g && index >= 0 && index < static_cast<int32_t>(g->ctxs.size())) ? g->ctxs[index] : nullptr;
}

int sac_cot_group_set(sac_cot_group* g, const char* name, int64_t value) {
  if (!g || !name) return SAC_COT_E_NULL;
  for (sac_cot_ctx* c : g->ctxs)
    if (int rc = sac_cot_ctx_set(c, name, value)) return rc;
  return SAC_COT_OK;
}

int sac_cot_group_register_packed(sac_cot_group* g, const float* src, const float* dst, const int64_t* offsets, int32_t B,
                                  const sac_cot_params* params, float* R, float* t, int32_t* inliers) {
  if (int rc = check_packed_args(g, src, dst, offsets, B, params, R, t, inliers, SAC_COT_LOC_HOST)) return rc;
  if (B == 0) return SAC_COT_OK;
  std::unique_lock<std::mutex> lk(g->m);
  g->src = src;
  g->dst = dst;
  g->offsets = offsets;
  g->B = B;
  g->prm = normalized(params);
  g->params = &g->prm;
  g->R = R;
  g->t = t;
  g->inliers = inliers;
  g->pending.store(static_cast<int>(g->ctxs.size()));
  g->epoch.fetch_add(1, std::memory_order_release);
  g->cv_go.notify_all();
  lk.unlock();
  {
    const auto t0 = std::chrono::steady_clock::now();
    while (g->pending.load(std::memory_order_acquire) != 0 && std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(20)) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  lk.lock();
  g->cv_done.wait(lk, [&] { return g->pending.load() == 0; });
  for (int rc : g->rc)
    if (rc) return rc;
  return SAC_COT_OK;
}

// ---- sharded single pair ------------------------------------------------------------------
}  // extern "C"

namespace {

// Part 1 on lane 0: inputs -> full graph -> triangle counts of this rank's cells -> its top-K_e edge keys.
// host_inputs: src/dst are host arrays (copied into the arena), else device arrays on the ctx device.
int shard_part1(sac_cot_ctx* ctx, Lane& ln, const float* src, const float* dst, int32_t N, const sac_cot_params& prm,
                int rank, int world, int xworld, bool host_inputs) {
  // S2 path as in the unsharded call (tensor cores for dense graphs, decided by the density of the WHOLE graph, the
  // same on every rank); a rank counts the edges of the cells it owns (common.cuh: owner_of_cell)
  const bool tensor = ctx->tri_path != 0;
  if (int rc = prepare_lane(ctx, ln, &N, 1, prm, host_inputs, tensor, rank, world, xworld)) return rc;
  ctx->prm = prm;
  Layout& L = ln.lay;
  if (int rc = fork_lanes(ctx, 1)) return rc;
  if (host_inputs) {
    CU_TRY(cudaMemcpyAsync(L.in_src, src, sizeof(float) * 3 * N, cudaMemcpyHostToDevice, ln.stream));
    CU_TRY(cudaMemcpyAsync(L.in_dst, dst, sizeof(float) * 3 * N, cudaMemcpyHostToDevice, ln.stream));
    src = L.in_src;
    dst = L.in_dst;
  }
  return enqueue_pipeline(ctx, ln, src, dst, nullptr, nullptr, nullptr, rank, world, true);
}

int run_sharded(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N, const sac_cot_params* params,
                float* R, float* t, int32_t* inliers, int32_t location) {
  if (!ctx || !src || !dst || !R || !t || !inliers) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST && location != SAC_COT_LOC_DEVICE) return SAC_COT_E_UNSUPPORTED;
  if (N < 3 || N > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  const sac_cot_params prm_n = normalized(params);
  params = &prm_n;
  if (params->compat_mode != SAC_COT_COMPAT_FIRST_ORDER) return SAC_COT_E_UNSUPPORTED;  // A2 would need every rank's counts
  NcclApi* nc = nccl_api();
  if (!ctx->comm || ctx->comm_world < 1 || !nc) return SAC_COT_E_COMM;
  CU_TRY(cudaSetDevice(ctx->device));
  if (int rc = resolve_pending(ctx, false)) return rc;
  const bool host = location == SAC_COT_LOC_HOST;
  const int rank = ctx->comm_rank, world = ctx->comm_world;
  ctx->sh_valid = false;
  ctx->ws_valid = false;
  Lane& ln = ctx->lanes[0];
  StageTimer& tm = ctx->timer;
  if (!host)
    if (int rc = sticky_before(ctx)) return rc;
  for (int attempt = 0; attempt < 3; ++attempt) {
    if (int rc = shard_part1(ctx, ln, src, dst, N, *params, rank, world, world, host)) return rc;
    Layout& L = ln.lay;
    LaunchCtx lc{ln.stream, ctx->sm_count};
    const int npad = ln.descs[0].Npad, Ke = L.Ke;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto span_begin = [&]() {
      if (tm.enabled && (e0 = tm.next())) cudaEventRecord(e0, ln.stream);
    };
    auto span_end = [&](int stage) {
      if (tm.enabled && e0 && (e1 = tm.next())) {
        cudaEventRecord(e1, ln.stream);
        tm.pending.push_back({stage, e0, e1});
      }
      e0 = nullptr;
    };
    // ---- exchange #1: every rank's record (partial node sums | its top-K_e | overflow flag, demand), device to
    //      device on the lane stream, then the merge: summed node sums, global top-K_e
    span_begin();
    KL_TRY(launch_shard_pack(lc, L.t2, L.top, L.chunk, L.xsend, npad, Ke));
    if (nc->AllGather(L.xsend, L.xrecv, L.xrec_len, ncclUint64, ctx->comm, ln.stream) != ncclSuccess) return SAC_COT_E_COMM;
    CU_TRY(cudaMemsetAsync(L.top, 0, sizeof(unsigned long long) * Ke, ln.stream));
    CU_TRY(cudaMemsetAsync(&L.state[0].best_key, 0, sizeof(unsigned long long), ln.stream));
    CU_TRY(cudaMemsetAsync(L.hyp_key, 0, sizeof(unsigned long long) * L.K, ln.stream));
    KL_TRY(launch_shard_merge(lc, L.xrecv, world, npad, Ke, L.t2, L.top, L.state, L.chunk, host ? nullptr : ctx->d_sticky,
                              L.xsum));
    span_end(ST_EXCH1);
    // ---- part 2: apexes and hypotheses everywhere (cheap, identical), scores of this rank's hypothesis range
    const float tau2 = params->tau_inlier * params->tau_inlier;
    span_begin();
    KL_TRY(launch_select_apex(lc, L.desc, 1, L.max_npad, L.adj, L.t2, L.top, L.tri, L.Ke, L.m, ctx->apex_path, nullptr, nullptr, nullptr));
    span_end(ST_APEX);
    span_begin();
    KL_TRY(launch_kabsch(lc, L.desc, 1, L.soa, L.tri, L.rt, L.K));
    span_end(ST_KABSCH);
    const int per = (L.K + world - 1) / world;
    const int h0 = std::min(L.K, rank * per), h1 = std::min(L.K, h0 + per);
    span_begin();
    KL_TRY(launch_score(lc, L.desc, 1, L.max_n, L.soa, L.tri, L.rt, L.hyp_key, L.state, tau2, L.K, h0, h1, params->score_mode));
    span_end(ST_SCORE);
    // ---- exchange #2: all-reduce(max) of the packed (score, hypothesis id) key
    span_begin();
    if (nc->AllReduce(&L.state[0].best_key, L.best_override, 1, ncclUint64, ncclMax, ctx->comm, ln.stream) != ncclSuccess)
      return SAC_COT_E_COMM;
    CU_TRY(cudaMemcpyAsync(&L.state[0].best_key, L.best_override, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ln.stream));
    span_end(ST_EXCH2);
    // ---- part 3: inlier mask + refit of the global winner, on every rank
    span_begin();
    KL_TRY(launch_finalize(lc, L.desc, 1, L.soa, L.rt, L.state, L.best_override, L.mask, host ? L.outR : R, host ? L.outT : t,
                           host ? L.outInl : inliers, tau2, L.K, params->refit));
    span_end(ST_FINALIZE);
    if (!host) {
      if (int rc = join_lanes(ctx, 1)) return rc;
      if (int rc = sticky_after(ctx)) return rc;
      ctx->ws_valid = true;
      return SAC_COT_OK;
    }
    CU_TRY(cudaMemcpyAsync(R, L.outR, sizeof(float) * 9, cudaMemcpyDeviceToHost, ln.stream));
    CU_TRY(cudaMemcpyAsync(t, L.outT, sizeof(float) * 3, cudaMemcpyDeviceToHost, ln.stream));
    CU_TRY(cudaMemcpyAsync(inliers, L.outInl, sizeof(int32_t), cudaMemcpyDeviceToHost, ln.stream));
    CU_TRY(cudaMemcpyAsync(ctx->h_xsum, L.xsum, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ln.stream));
    if (int rc = join_lanes(ctx, 1)) return rc;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (!ctx->h_xsum[0]) {
      ctx->ws_valid = true;  // the pair is resident: debug_get(pair 0) works
      return SAC_COT_OK;
    }
    // some rank ran out of key-pool space; every rank saw the same summary: grow and re-run in step
    ++ctx->retries;
    const unsigned long long demand = ctx->h_xsum[1];
    if (int rc = ensure_keys(ctx, ln, demand + demand / 16 + 1024)) return rc;
  }
  return SAC_COT_E_NOMEM;
}

}  // namespace

extern "C" {

int sac_cot_comm_unique_id(void* id_out) {
  if (!id_out) return SAC_COT_E_NULL;
  NcclApi* nc = nccl_api();
  if (!nc) return SAC_COT_E_COMM;
  ncclUniqueId id;
  if (nc->GetUniqueId(&id) != ncclSuccess) return SAC_COT_E_COMM;
  std::memcpy(id_out, &id, sizeof(id));
  return SAC_COT_OK;
}

int sac_cot_ctx_set_comm(sac_cot_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world) {
  if (!ctx) return SAC_COT_E_NULL;
  if (nccl_comm && (world < 1 || world > 64 || rank < 0 || rank >= world)) return SAC_COT_E_SIZE;
  NcclApi* nc = nccl_api();
  if (nccl_comm && !nc) return SAC_COT_E_COMM;
  cudaSetDevice(ctx->device);
  (void)sync_all(ctx);
  if (ctx->comm && ctx->own_comm && nc) nc->CommDestroy(ctx->comm);
  ctx->comm = static_cast<ncclComm_t>(nccl_comm);
  ctx->own_comm = false;
  ctx->comm_rank = nccl_comm ? rank : 0;
  ctx->comm_world = nccl_comm ? world : 0;
  return SAC_COT_OK;
}

int sac_cot_ctx_comm_init(sac_cot_ctx* ctx, const void* id, int32_t rank, int32_t world) {
  if (!ctx || !id) return SAC_COT_E_NULL;
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return SAC_COT_E_SIZE;
  NcclApi* nc = nccl_api();
  if (!nc) return SAC_COT_E_COMM;
  if (int rc = sac_cot_ctx_set_comm(ctx, nullptr, 0, 0)) return rc;
  CU_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  if (nc->CommInitRank(&comm, world, uid, rank) != ncclSuccess) return SAC_COT_E_COMM;
  ctx->comm = comm;
  ctx->own_comm = true;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return SAC_COT_OK;
}

int sac_cot_register_sharded(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N,
                             const sac_cot_params* params, float R[9], float t[3], int32_t* inliers, int32_t location) {
  try {
    return run_sharded(ctx, src, dst, N, params, R, t, inliers, location);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

int sac_cot_sharded_phase1(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N,
                           const sac_cot_params* params, int32_t rank, int32_t world, uint64_t* t_partial,
                           uint64_t* cand) {
  if (!ctx || !src || !dst || !t_partial || !cand) return SAC_COT_E_NULL;
  if (N < 3 || N > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  if (world < 1 || rank < 0 || rank >= world) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  const sac_cot_params prm_n = normalized(params);
  params = &prm_n;
  if (params->compat_mode != SAC_COT_COMPAT_FIRST_ORDER) return SAC_COT_E_UNSUPPORTED;  // A2 would need every rank's counts
  CU_TRY(cudaSetDevice(ctx->device));
  if (int rc = resolve_pending(ctx, false)) return rc;
  ctx->sh_valid = false;
  ctx->ws_valid = false;
  Lane& ln = ctx->lanes[0];
  try {
    if (int rc = ensure_chunk_headers(ctx, 1)) return rc;
    ChunkDev* h_chunk = ctx->h_chunks[0];
    for (int attempt = 0; attempt < 3; ++attempt) {
      if (int rc = shard_part1(ctx, ln, src, dst, N, *params, rank, world, 0, true)) return rc;
      Layout& L = ln.lay;
      CU_TRY(cudaMemcpyAsync(h_chunk, L.chunk, sizeof(ChunkDev), cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpyAsync(t_partial, L.t2, sizeof(uint64_t) * N, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaMemcpyAsync(cand, L.top, sizeof(uint64_t) * params->num_edges, cudaMemcpyDeviceToHost, ln.stream));
      CU_TRY(cudaStreamSynchronize(ln.stream));
      if (!h_chunk->overflow) {
        ctx->sh_rank = rank;
        ctx->sh_world = world;
        ctx->sh_N = N;
        ctx->sh_valid = true;
        return SAC_COT_OK;
      }
      ++ctx->retries;
      if (int rc = ensure_keys(ctx, ln, h_chunk->total_edges + h_chunk->total_edges / 16 + 1024)) return rc;
    }
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_E_NOMEM;
}

int sac_cot_sharded_phase2(sac_cot_ctx* ctx, const uint64_t* t_all, const uint64_t* cand_all, uint64_t* best_key) {
  if (!ctx || !t_all || !cand_all || !best_key) return SAC_COT_E_NULL;
  if (!ctx->sh_valid) return SAC_COT_E_SIZE;
  CU_TRY(cudaSetDevice(ctx->device));
  Lane& ln = ctx->lanes[0];
  try {
    Layout& L = ln.lay;
    const int N = ctx->sh_N, world = ctx->sh_world, Ke = L.Ke;
    // exchange #1 results: node sums add up over ranks; candidates merge to the global top-K_e
    std::vector<unsigned long long> t2(static_cast<size_t>(ln.descs[0].Npad), 0ull);
    for (int g = 0; g < world; ++g)
      for (int i = 0; i < N; ++i) t2[i] += t_all[static_cast<size_t>(g) * N + i];
    std::vector<unsigned long long> all;
    all.reserve(static_cast<size_t>(world) * Ke);
    for (size_t k = 0; k < static_cast<size_t>(world) * Ke; ++k)
      if (cand_all[k]) all.push_back(cand_all[k]);
    std::sort(all.begin(), all.end(), [](unsigned long long a, unsigned long long b) { return a > b; });
    all.resize(static_cast<size_t>(Ke), 0ull);
    CU_TRY(cudaMemcpyAsync(L.t2, t2.data(), sizeof(unsigned long long) * t2.size(), cudaMemcpyHostToDevice, ln.stream));
    CU_TRY(cudaMemcpyAsync(L.top, all.data(), sizeof(unsigned long long) * Ke, cudaMemcpyHostToDevice, ln.stream));
    CU_TRY(cudaMemsetAsync(&L.state[0].best_key, 0, sizeof(unsigned long long), ln.stream));
    CU_TRY(cudaMemsetAsync(L.hyp_key, 0, sizeof(unsigned long long) * L.K, ln.stream));
    LaunchCtx lc{ln.stream, ctx->sm_count};
    const float tau2 = ctx->prm.tau_inlier * ctx->prm.tau_inlier;
    KL_TRY(launch_select_apex(lc, L.desc, 1, L.max_npad, L.adj, L.t2, L.top, L.tri, L.Ke, L.m, ctx->apex_path, nullptr, nullptr, nullptr));
    KL_TRY(launch_kabsch(lc, L.desc, 1, L.soa, L.tri, L.rt, L.K));
    const int per = (L.K + world - 1) / world;
    const int h0 = std::min(L.K, ctx->sh_rank * per), h1 = std::min(L.K, h0 + per);
    KL_TRY(launch_score(lc, L.desc, 1, L.max_n, L.soa, L.tri, L.rt, L.hyp_key, L.state, tau2, L.K, h0, h1,
                        ctx->prm.score_mode));
    unsigned long long best = 0;
    CU_TRY(cudaMemcpyAsync(&best, &L.state[0].best_key, sizeof(best), cudaMemcpyDeviceToHost, ln.stream));
    CU_TRY(cudaStreamSynchronize(ln.stream));
    *best_key = best;
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_OK;
}

int sac_cot_sharded_phase3(sac_cot_ctx* ctx, uint64_t best_key_global, float R[9], float t[3], int32_t* inliers) {
  if (!ctx || !R || !t || !inliers) return SAC_COT_E_NULL;
  if (!ctx->sh_valid) return SAC_COT_E_SIZE;
  CU_TRY(cudaSetDevice(ctx->device));
  Lane& ln = ctx->lanes[0];
  Layout& L = ln.lay;
  const unsigned long long key = best_key_global;
  CU_TRY(cudaMemcpyAsync(L.best_override, &key, sizeof(key), cudaMemcpyHostToDevice, ln.stream));
  CU_TRY(cudaMemcpyAsync(&L.state[0].best_key, &key, sizeof(key), cudaMemcpyHostToDevice, ln.stream));
  LaunchCtx lc{ln.stream, ctx->sm_count};
  const float tau2 = ctx->prm.tau_inlier * ctx->prm.tau_inlier;
  KL_TRY(launch_finalize(lc, L.desc, 1, L.soa, L.rt, L.state, L.best_override, L.mask, L.outR, L.outT, L.outInl, tau2,
                         L.K, ctx->prm.refit));
  CU_TRY(cudaMemcpyAsync(R, L.outR, sizeof(float) * 9, cudaMemcpyDeviceToHost, ln.stream));
  CU_TRY(cudaMemcpyAsync(t, L.outT, sizeof(float) * 3, cudaMemcpyDeviceToHost, ln.stream));
  CU_TRY(cudaMemcpyAsync(inliers, L.outInl, sizeof(int32_t), cudaMemcpyDeviceToHost, ln.stream));
  CU_TRY(cudaStreamSynchronize(ln.stream));
  ctx->ws_valid = true;  // the single pair is fully resident: debug_get(pair = 0 or -1) works
  return SAC_COT_OK;
}

// ---- parity getter ------------------------------------------------------------------------
int sac_cot_debug_get(sac_cot_ctx* ctx, int32_t pair, int32_t which, void* out, size_t cap, size_t* written) {
  if (!ctx || !written) return SAC_COT_E_NULL;
  if (pair == -1) pair = 0;
  Lane& ln = ctx->lanes[0];
  if (!(ctx->ws_valid || ctx->sh_valid) || pair < 0 || pair >= ln.lay.pairs) return SAC_COT_E_WHICH;
  CU_TRY(cudaSetDevice(ctx->device));
  if (int rc = resolve_pending(ctx, true)) return rc;
  const Layout& L = ln.lay;
  const PairDesc& d = ln.descs[pair];
  PairDev st;
  CU_TRY(cudaMemcpy(&st, L.state + pair, sizeof(st), cudaMemcpyDeviceToHost));
  const void* dsrc = nullptr;
  size_t bytes = 0;
  uint64_t scalar = 0;
  std::vector<uint32_t> tmp32;
  switch (which) {
    case SAC_COT_DBG_ADJ: dsrc = L.adj_rank + d.adj_off; bytes = static_cast<size_t>(d.N) * d.stride * 4; break;
    case SAC_COT_DBG_ADJ_FIRST: dsrc = L.adj + d.adj_off; bytes = static_cast<size_t>(d.N) * d.stride * 4; break;
    case SAC_COT_DBG_T_NODE: bytes = static_cast<size_t>(d.N) * 4; break;
    case SAC_COT_DBG_NUM_EDGES: scalar = st.num_edges; bytes = 8; break;
    case SAC_COT_DBG_EDGE_KEYS: dsrc = ln.keys + st.key_base; bytes = static_cast<size_t>(st.key_count) * 8; break;
    case SAC_COT_DBG_TOP_EDGES: dsrc = L.top + static_cast<size_t>(pair) * L.Ke; bytes = static_cast<size_t>(st.n_sel) * 8; break;
    case SAC_COT_DBG_TRIANGLES: dsrc = L.tri + static_cast<size_t>(pair) * L.K * 3; bytes = static_cast<size_t>(L.K) * 12; break;
    case SAC_COT_DBG_HYP_RT: dsrc = L.rt + static_cast<size_t>(pair) * L.K * 12; bytes = static_cast<size_t>(L.K) * 48; break;
    case SAC_COT_DBG_HYP_SCORE: dsrc = L.hyp_key + static_cast<size_t>(pair) * L.K; bytes = static_cast<size_t>(L.K) * 8; break;
    case SAC_COT_DBG_BEST_KEY: scalar = st.best_key; bytes = 8; break;
    case SAC_COT_DBG_MASK: dsrc = L.mask + d.mask_off; bytes = static_cast<size_t>((d.N + 31) / 32) * 4; break;
    case SAC_COT_DBG_HIST: dsrc = L.hist + static_cast<size_t>(pair) * kHistBins; bytes = kHistBins * 4; break;
    default: return SAC_COT_E_WHICH;
  }
  *written = bytes;
  if (bytes > cap) return SAC_COT_E_CAPACITY;
  if (bytes && !out) return SAC_COT_E_NULL;
  if (!bytes) return SAC_COT_OK;
  if (which == SAC_COT_DBG_NUM_EDGES || which == SAC_COT_DBG_BEST_KEY) {
    std::memcpy(out, &scalar, 8);
    return SAC_COT_OK;
  }
  if (which == SAC_COT_DBG_T_NODE) {
    try {
      std::vector<unsigned long long> t2(static_cast<size_t>(d.N));
      CU_TRY(cudaMemcpy(t2.data(), L.t2 + d.node_off, sizeof(unsigned long long) * d.N, cudaMemcpyDeviceToHost));
      uint32_t* o = static_cast<uint32_t*>(out);
      for (int i = 0; i < d.N; ++i) o[i] = static_cast<uint32_t>(t2[i] / 2);
    } catch (const std::bad_alloc&) {
      return SAC_COT_E_NOMEM;
    }
    return SAC_COT_OK;
  }
  CU_TRY(cudaMemcpy(out, dsrc, bytes, cudaMemcpyDeviceToHost));
  return SAC_COT_OK;
}

const char* sac_cot_strerror(int status) {
  switch (status) {
    case SAC_COT_OK: return "ok";
    case SAC_COT_E_NULL: return "null pointer argument";
    case SAC_COT_E_SIZE: return "size out of range (3 <= N <= 65535, B >= 0) or call out of sequence";
    case SAC_COT_E_PARAMS: return "invalid sac_cot_params";
    case SAC_COT_E_NODEVICE: return "no usable sm_100 CUDA device";
    case SAC_COT_E_UNSUPPORTED: return "not supported by this implementation";
    case SAC_COT_E_WHICH: return "unknown selector / index, or nothing resident";
    case SAC_COT_E_CAPACITY: return "output buffer too small";
    case SAC_COT_E_NOMEM: return "out of device memory / workspace overflow";
    case SAC_COT_E_COMM: return "no communicator on the ctx, NCCL not loadable, or an NCCL call failed";
    default: return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown status";
  }
}

const char* sac_cot_version(void) { return "sac-cot-b200 0.2 (cuda sm_100a)"; }

}  // extern "C"
