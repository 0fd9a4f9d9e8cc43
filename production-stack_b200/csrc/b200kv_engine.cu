// b200kv_engine.cu — per-GPU runtime behind the C ABI (include/b200kv.h).
//
// Stands in for LMCacheEngine.store / .retrieve as driven by vLLM's adapter
// (vllm/.../lmcache_integration/vllm_v1_adapter.py:1115-1123 store, :882-889 retrieve) and for
// the NIXL peer channel (helm/templates/deployment-vllm-multi.yaml:296-324).
//
// Streams (all created on the engine's device):
//   s_gather   store kernels          (lowest priority: shares SMs with decode)
//   s_d2h      staging -> pinned pool (copy engine)
//   s_h2d      pinned pool -> staging (copy engine)
//   s_scatter  load / pull kernels    (highest priority: on the TTFT critical path)
// The caller's compute stream only ever waits for a *kernel* (gather done / scatter done),
// never for a PCIe transfer of a store.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "b200kv.h"
#include "b200kv_kernels.cuh"

using namespace b200kv;

namespace {

thread_local char g_err[512] = {0};

void set_err(const char* what, cudaError_t e, int line) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s (%s) at b200kv_engine.cu:%d", what,
                cudaGetErrorName(e), cudaGetErrorString(e), line);
}

#define CU_TRY(expr)                          \
  do {                                        \
    cudaError_t _e = (expr);                  \
    if (_e != cudaSuccess) {                  \
      set_err(#expr, _e, __LINE__);           \
      return B200KV_ENODEV;                   \
    }                                         \
  } while (0)

constexpr int kTableSlots = 16;
constexpr size_t kTableBytes = 1u << 20;
constexpr int kMaxPeers = 64;
constexpr uint32_t kStageMax = 32u << 10;

struct StageSlot {
  cudaEvent_t free_ev = nullptr;  // recorded when the last consumer of this slot is done
  bool used = false;
};

struct TableSlot {
  uint8_t* host = nullptr;  // pinned
  uint8_t* dev = nullptr;
  cudaEvent_t ev = nullptr;  // H2D of the table finished (host side reusable)
  cudaEvent_t done_ev = nullptr;  // kernels reading the device table finished
  bool used = false;
};

struct Op {
  uint64_t id = 0;
  b200kv_pool* pool = nullptr;
  std::vector<uint64_t> commit_keys;   // store: commit after D2H
  std::vector<uint64_t> release_keys;  // load: unpin after scatter
  cudaEvent_t done = nullptr;
  std::atomic<int> host_done{0};
  // layer-wise loads: group_ev[g] fires when layers [g*G, (g+1)*G) of every chunk are in the pages
  std::vector<cudaEvent_t> group_ev;
  int layers_per_group = 0;
};

void destroy_op_events(Op* op) {
  if (op->done) cudaEventDestroy(op->done);
  op->done = nullptr;
  for (cudaEvent_t e : op->group_ev) cudaEventDestroy(e);
  op->group_ev.clear();
}

void CUDART_CB op_host_cb(void* p) {
  Op* op = static_cast<Op*>(p);
  if (op->pool) {
    for (uint64_t k : op->commit_keys) b200kv_pool_commit(op->pool, k);
    for (uint64_t k : op->release_keys) b200kv_pool_release(op->pool, k);
  }
  op->host_done.store(1, std::memory_order_release);
}

// Undoes what a store / load had taken when it leaves early (argument the table cannot hold, a failed
// CUDA call): reserved-but-uncommitted pool slots are aborted (otherwise later stores of those keys see
// -EEXIST until the stale-writer reclaim), pinned chunks are released (otherwise they can never be
// evicted), and the op's events are destroyed.  The engine's streams are drained first, so no copy into
// or out of those slots is still in flight.
struct OpGuard {
  b200kv_ctx* ctx;
  cudaStream_t s[4];
  std::unique_ptr<Op> op;
  std::vector<uint64_t> reserved;   // store: keys reserved in the pool, not yet handed to the commit callback
  bool cb_enqueued = false;         // the host callback owns commit / release from here on
  bool armed = true;
  OpGuard(b200kv_ctx* c, cudaStream_t a, cudaStream_t b, cudaStream_t d, cudaStream_t e) : ctx(c), s{a, b, d, e} {}
  ~OpGuard() {
    if (!armed || (!op && reserved.empty())) return;
    for (cudaStream_t st : s) if (st) cudaStreamSynchronize(st);
    cudaGetLastError();
    if (!cb_enqueued) {
      for (uint64_t k : reserved) b200kv_pool_abort(ctx_pool(), k);
      if (op) for (uint64_t k : op->release_keys) b200kv_pool_release(ctx_pool(), k);
    }
    if (op) destroy_op_events(op.get());
  }
  b200kv_pool* ctx_pool() const;
};

struct Peer {
  bool valid = false;
  int device = -1;
  uint64_t* d_bases = nullptr;
  uint64_t block_stride = 0;
  uint64_t n_blocks = 0;
  std::vector<void*> opened;  // cudaIpcOpenMemHandle mappings to close
};

struct Geometry {
  uint32_t L, H, D, elem, bs, C, planes;
  uint32_t token_bytes;   // H*D*elem
  uint64_t slab_bytes;    // C*token_bytes (RAW) or C*H*D (FP8)
  uint64_t chunk_bytes;
  uint64_t scales_off;    // FP8 only
  uint32_t fmt_token_bytes;
  bool hnd;               // [H][bs][D] inside a block (B200KV_LAYOUT_HND)
  uint32_t row_bytes;     // D*elem
};

int make_geometry(const b200kv_engine_config* c, Geometry* g) {
  if (c->n_layers <= 0 || c->n_kv_heads <= 0 || c->head_dim <= 0 || c->block_tokens <= 0 ||
      c->chunk_tokens <= 0 || c->chunk_tokens % c->block_tokens)
    return B200KV_EINVAL;
  if (c->elem_bytes != 1 && c->elem_bytes != 2 && c->elem_bytes != 4) return B200KV_EINVAL;
  g->L = c->n_layers;
  g->H = c->n_kv_heads;
  g->D = c->head_dim;
  g->elem = c->elem_bytes;
  g->bs = c->block_tokens;
  g->C = c->chunk_tokens;
  g->planes = 2u * g->L;
  g->token_bytes = g->H * g->D * g->elem;
  g->row_bytes = g->D * g->elem;
  if (c->kv_layout != B200KV_LAYOUT_NHD && c->kv_layout != B200KV_LAYOUT_HND) return B200KV_EINVAL;
  g->hnd = c->kv_layout == B200KV_LAYOUT_HND;
  if (g->token_bytes % 16) return B200KV_EINVAL;  // 16-byte vectors / bulk-copy granularity
  if (g->hnd && g->row_bytes % 16) return B200KV_EINVAL;
  if (c->format == B200KV_FMT_RAW) {
    g->fmt_token_bytes = g->token_bytes;
    g->slab_bytes = static_cast<uint64_t>(g->C) * g->token_bytes;
    g->scales_off = 0;
    g->chunk_bytes = g->slab_bytes * g->planes;
  } else if (c->format == B200KV_FMT_FP8) {
    if (g->elem != 2) return B200KV_ENOTSUP;  // source must be bf16
    if (g->hnd && (g->C / kCluster) % g->bs) return B200KV_ENOTSUP;  // a cluster window = whole tiles
    if (g->H > kMaxHeads) return B200KV_ENOTSUP;
    if (g->C % kCluster) return B200KV_EINVAL;
    if ((g->D * 2) % 16) return B200KV_EINVAL;
    g->fmt_token_bytes = g->token_bytes / 2;
    g->slab_bytes = static_cast<uint64_t>(g->C) * g->fmt_token_bytes;
    g->scales_off = g->slab_bytes * g->planes;
    g->chunk_bytes = g->scales_off + static_cast<uint64_t>(g->planes) * g->H * sizeof(float);
    g->chunk_bytes = (g->chunk_bytes + 255) / 256 * 256;
  } else if (c->format == B200KV_FMT_Q4) {
    if (g->elem != 2) return B200KV_ENOTSUP;
    if (g->D % 32) return B200KV_EINVAL;
    g->fmt_token_bytes = g->H * g->D / 2 + g->H * (g->D / 32) * 2;   // one token record of one plane
    g->slab_bytes = static_cast<uint64_t>(g->C) * g->fmt_token_bytes;
    g->scales_off = 0;
    g->chunk_bytes = (g->slab_bytes * g->planes + 255) / 256 * 256;
  } else {
    return B200KV_EINVAL;
  }
  return B200KV_OK;
}

}  // namespace

struct b200kv_ctx {
  b200kv_engine_config cfg{};
  Geometry g{};
  b200kv_pool* pool = nullptr;
  bool pool_registered = false;
  std::string numa_placement = "no pool";   // what place_pool_pages() did when this engine pinned the pool
  void* pool_base = nullptr;
  int sm_count = 148;
  std::mutex mu;

  cudaStream_t s_gather = nullptr, s_d2h = nullptr, s_h2d = nullptr, s_scatter = nullptr;
  uint64_t* d_bases = nullptr;
  bool kv_registered = false;
  std::vector<const void*> h_k, h_v;

  uint8_t* d_staging = nullptr;
  std::vector<StageSlot> stage;
  // The ring is split in two halves — stores allocate from [0, n_store), loads from [n_store, n):
  // a TTFT-critical load must never queue behind the D2H of an unrelated store.
  uint32_t n_store_slots = 0, store_next = 0, load_next = 0;

  TableSlot tables[kTableSlots];
  uint32_t table_next = 0;

  Peer peers[kMaxPeers];

  uint64_t next_ticket = 1;
  std::map<uint64_t, std::unique_ptr<Op>> ops;

  // timing of the most recent kernel batch of each kind (CUDA events on the launching stream)
  struct Timing {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t used = 0;
  } timing[3];

  b200kv_engine_stats stats{};

  // bulk-kernel launch shape
  int S = 2, LAG = 1, ctas_per_sm = 1;  // swept on B200: profiles/sweep_r01.txt
  int fp8_threads = 256;                // B200KV_FP8_THREADS: CTA width of the FP8 store kernel
  bool fp8_two_pass = false;            // B200KV_FP8_2PASS=1: smem-free two-pass store kernel (experimental, slower)
  int fp8_store_ver = 3;                // B200KV_FP8_STORE: 3 = persistent warp-specialised kernel (default), 1 = cluster kernel
  void* d_tmaps = nullptr;              // one CUtensorMap per plane (NHD pages: strided [bs][D] boxes for the FP8 store)
  std::string tmap_status = "not built";
  uint32_t piece_tokens = 0, pieces = 0, stage_bytes = 0;

  void* tier_base = nullptr;            // device chunk tier (b200kv_tier_create)
  uint32_t tier_slots = 0;
  std::vector<void*> tier_imports;      // peers' tiers opened over CUDA IPC
};

b200kv_pool* OpGuard::ctx_pool() const { return ctx->pool; }

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// ---- run construction (host) -----------------------------------------------------------------
// slot_mapping[i] for op-relative token i (vllm_v1_adapter.py:368-375).  A run never crosses a
// vLLM block or a chunk boundary, so both sides of every run are contiguous byte ranges.
// For HND tiles a run additionally stops at every chunk-side tile boundary, and runs are split
// into `full` (one whole aligned block on both sides: a contiguous tile, moved by the bulk kernel
// exactly like NHD) and `partial` (moved per head by kv_hnd_partial_kernel).  NHD: all `full`.
struct RunSink {
  std::vector<Run>* full;
  std::vector<Run>* partial;
  bool hnd;
  int32_t bs;
  bool both_paged;  // peer pull: `b` is a paged slot too
  void push(const Run& r) const {
    if (!hnd || !partial) { full->push_back(r); return; }  // FP8 kernels take one sorted list
    const bool whole = r.n == bs && (r.a % bs) == 0 && (r.b % bs) == 0;
    (whole ? full : partial)->push_back(r);
  }
};

int build_runs(const b200kv_ctx* ctx, const int64_t* slots, int64_t tok_begin, int64_t tok_end,
               int64_t b_shift, std::vector<Run>* full, std::vector<Run>* partial) {
  const Geometry& g = ctx->g;
  const RunSink sink{full, partial, g.hnd, static_cast<int32_t>(g.bs), false};
  const int64_t max_slot = static_cast<int64_t>(ctx->cfg.n_blocks) * g.bs;
  Run cur{0, 0, 0};
  for (int64_t i = tok_begin; i < tok_end; ++i) {
    const int64_t s = slots[i];
    if (s < 0 || s >= max_slot) return B200KV_EINVAL;
    const bool extend = cur.n > 0 && s == static_cast<int64_t>(cur.a) + cur.n && (s % g.bs) != 0 &&
                        ((i - b_shift) % g.C) != 0 && !(g.hnd && ((i - b_shift) % g.bs) == 0);
    if (extend) {
      ++cur.n;
    } else {
      if (cur.n) sink.push(cur);
      cur.a = static_cast<int32_t>(s);
      cur.b = static_cast<int32_t>(i - b_shift);
      cur.n = 1;
    }
  }
  if (cur.n) sink.push(cur);
  return B200KV_OK;
}

int build_pull_runs(const b200kv_ctx* ctx, const Peer& peer, const int64_t* src, const int64_t* dst,
                    int64_t n, std::vector<Run>* full, std::vector<Run>* partial) {
  const Geometry& g = ctx->g;
  const RunSink sink{full, partial, g.hnd, static_cast<int32_t>(g.bs), true};
  const int64_t max_dst = static_cast<int64_t>(ctx->cfg.n_blocks) * g.bs;
  const int64_t max_src = static_cast<int64_t>(peer.n_blocks) * g.bs;
  Run cur{0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    const int64_t s = src[i], d = dst[i];
    if (s < 0 || s >= max_src || d < 0 || d >= max_dst) return B200KV_EINVAL;
    const bool extend = cur.n > 0 && s == static_cast<int64_t>(cur.a) + cur.n &&
                        d == static_cast<int64_t>(cur.b) + cur.n && (s % g.bs) != 0 && (d % g.bs) != 0;
    if (extend) {
      ++cur.n;
    } else {
      if (cur.n) sink.push(cur);
      cur.a = static_cast<int32_t>(s);
      cur.b = static_cast<int32_t>(d);
      cur.n = 1;
    }
  }
  if (cur.n) sink.push(cur);
  return B200KV_OK;
}

// ---- table ring ------------------------------------------------------------------------------
struct TableView {
  TableSlot* slot;
  size_t runs_off, addrs_off, offs_off, bytes;
};

int table_acquire(b200kv_ctx* ctx, size_t n_runs, size_t n_chunks, TableView* tv) {
  const size_t runs_bytes = (n_runs * sizeof(Run) + 15) / 16 * 16;
  const size_t addrs_bytes = (n_chunks * 8 + 15) / 16 * 16;
  const size_t offs_bytes = ((n_chunks + 1) * 4 + 15) / 16 * 16;
  const size_t total = runs_bytes + addrs_bytes + offs_bytes;
  if (total > kTableBytes) return B200KV_EINVAL;  // op too large: caller must split
  TableSlot& t = ctx->tables[ctx->table_next];
  ctx->table_next = (ctx->table_next + 1) % kTableSlots;
  if (t.used) {  // host must not overwrite a table whose upload / readers are still in flight
    CU_TRY(cudaEventSynchronize(t.ev));
    CU_TRY(cudaEventSynchronize(t.done_ev));
  }
  tv->slot = &t;
  tv->runs_off = 0;
  tv->addrs_off = runs_bytes;
  tv->offs_off = runs_bytes + addrs_bytes;
  tv->bytes = total;
  return B200KV_OK;
}

int table_upload(b200kv_ctx* ctx, const TableView& tv, cudaStream_t s) {
  CU_TRY(cudaMemcpyAsync(tv.slot->dev, tv.slot->host, tv.bytes, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaEventRecord(tv.slot->ev, s));
  tv.slot->used = true;
  ctx->stats.h2d_bytes += tv.bytes;
  return B200KV_OK;
}

// ---- timing events -----------------------------------------------------------------------------
void timing_reset(b200kv_ctx* ctx, int which) { ctx->timing[which].used = 0; }

int timing_begin(b200kv_ctx* ctx, int which, cudaStream_t s) {
  auto& t = ctx->timing[which];
  if (t.used == t.ev.size()) {
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreate(&a));
    CU_TRY(cudaEventCreate(&b));
    t.ev.emplace_back(a, b);
  }
  CU_TRY(cudaEventRecord(t.ev[t.used].first, s));
  return B200KV_OK;
}
int timing_end(b200kv_ctx* ctx, int which, cudaStream_t s) {
  auto& t = ctx->timing[which];
  CU_TRY(cudaEventRecord(t.ev[t.used].second, s));
  ++t.used;
  return B200KV_OK;
}

// ---- kernel launches ---------------------------------------------------------------------------
template <int MODE, int S, int LAG>
int launch_bulk_t(b200kv_ctx* ctx, const CopyParams& p, cudaStream_t s) {
  const size_t smem = 256 + static_cast<size_t>(S) * ctx->stage_bytes;
  static bool attr_set[8] = {false};  // per device ordinal (<=8 GPUs per box)
  const int dev = ctx->cfg.device & 7;
  if (!attr_set[dev]) {
    CU_TRY(cudaFuncSetAttribute(kv_bulk_copy_kernel<MODE, S, LAG>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  uint32_t grid = static_cast<uint32_t>(ctx->sm_count * ctx->ctas_per_sm);
  grid = std::min(grid, p.total_units);
  if (grid == 0) return B200KV_OK;
  kv_bulk_copy_kernel<MODE, S, LAG><<<grid, 32, smem, s>>>(p, ctx->stage_bytes);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return B200KV_OK;
}

template <int MODE>
int launch_bulk(b200kv_ctx* ctx, const CopyParams& p, cudaStream_t s) {
  switch (ctx->S * 10 + ctx->LAG) {
    case 21: return launch_bulk_t<MODE, 2, 1>(ctx, p, s);
    case 31: return launch_bulk_t<MODE, 3, 1>(ctx, p, s);
    case 32: return launch_bulk_t<MODE, 3, 2>(ctx, p, s);
    case 41: return launch_bulk_t<MODE, 4, 1>(ctx, p, s);
    case 42: return launch_bulk_t<MODE, 4, 2>(ctx, p, s);
    case 43: return launch_bulk_t<MODE, 4, 3>(ctx, p, s);
    case 63: return launch_bulk_t<MODE, 6, 3>(ctx, p, s);
    case 64: return launch_bulk_t<MODE, 6, 4>(ctx, p, s);
    default: return B200KV_EINVAL;
  }
}

template <int MODE>
int launch_copy(b200kv_ctx* ctx, const CopyParams& p, cudaStream_t s) {
  if (ctx->cfg.variant == B200KV_VARIANT_LDG) {
    uint32_t grid = static_cast<uint32_t>(ctx->sm_count * 8);
    grid = std::min(grid, p.total_units);
    if (grid == 0) return B200KV_OK;
    kv_ldg_copy_kernel<MODE><<<grid, 256, 0, s>>>(p);
    CU_TRY(cudaGetLastError());
    ++ctx->stats.n_kernel_launches;
    return B200KV_OK;
  }
  return launch_bulk<MODE>(ctx, p, s);
}

// format tag recorded with a chunk: a pool shared by engines of different tile order or codec
// never hands one engine's bytes to another
uint32_t pool_fmt(const b200kv_ctx* ctx) {
  return static_cast<uint32_t>(ctx->cfg.format) | (static_cast<uint32_t>(ctx->cfg.kv_layout) << 8);
}

PagedSide local_side(const b200kv_ctx* ctx) {
  PagedSide s;
  s.bases = ctx->d_bases;
  s.block_stride = ctx->cfg.block_stride_bytes;
  s.block_tokens = ctx->g.bs;
  s.token_bytes = ctx->g.token_bytes;
  return s;
}

CopyParams make_copy_params(const b200kv_ctx* ctx, const uint8_t* dev_table, const TableView& tv,
                            size_t run_begin, size_t n_runs, uint32_t plane_begin = 0, uint32_t n_planes = 0) {
  CopyParams p{};
  p.paged = local_side(ctx);
  p.peer = p.paged;
  p.chunk.chunk_addrs = reinterpret_cast<const uint64_t*>(dev_table + tv.addrs_off);
  p.chunk.slab_bytes = ctx->g.slab_bytes;
  p.chunk.chunk_tokens = ctx->g.C;
  p.chunk.token_bytes = ctx->g.fmt_token_bytes;
  p.runs = reinterpret_cast<const Run*>(dev_table + tv.runs_off) + run_begin;
  p.n_runs = static_cast<uint32_t>(n_runs);
  p.n_planes = n_planes ? n_planes : ctx->g.planes;
  p.plane_begin = plane_begin;
  p.pieces = ctx->pieces;
  p.piece_tokens = ctx->piece_tokens;
  p.total_units = p.n_runs * p.n_planes * p.pieces;
  return p;
}

template <int MODE>
int launch_hnd_partial(b200kv_ctx* ctx, const CopyParams& cp, const Run* runs, uint32_t n_runs, cudaStream_t s) {
  if (n_runs == 0) return B200KV_OK;
  HndParams p{};
  p.paged = cp.paged;
  p.peer = cp.peer;
  p.chunk = cp.chunk;
  p.runs = runs;
  p.n_runs = n_runs;
  p.n_planes = cp.n_planes;
  p.plane_begin = cp.plane_begin;
  p.n_heads = ctx->g.H;
  p.row_bytes = ctx->g.row_bytes;
  p.total_units = n_runs * p.n_planes * p.n_heads;
  const uint32_t grid = std::min<uint32_t>(p.total_units, static_cast<uint32_t>(ctx->sm_count) * 16u);
  kv_hnd_partial_kernel<MODE><<<grid, 128, 0, s>>>(p);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return B200KV_OK;
}

// runs [run_begin, +n_full) are whole tiles (or anything, for NHD); the next n_partial are HND
// partial-tile runs.  `peer` non-null selects the pull source.
template <int MODE>
int launch_copy_runs(b200kv_ctx* ctx, const uint8_t* dev_table, const TableView& tv, size_t run_begin,
                     size_t n_full, size_t n_partial, cudaStream_t s, const Peer* peer = nullptr,
                     uint32_t plane_begin = 0, uint32_t n_planes = 0) {
  CopyParams p = make_copy_params(ctx, dev_table, tv, run_begin, n_full, plane_begin, n_planes);
  if (peer) {
    p.peer.bases = peer->d_bases;
    p.peer.block_stride = peer->block_stride;
  }
  int rc = B200KV_OK;
  if (n_full) rc = launch_copy<MODE>(ctx, p, s);
  if (rc) return rc;
  return launch_hnd_partial<MODE>(ctx, p, p.runs + n_full, static_cast<uint32_t>(n_partial), s);
}

// Persistent FP8 store (kv_fp8_store3_kernel) when every chunk of the batch is a sequence of whole, block-aligned
// runs (the last possibly short).  Returns 1 if it launched, 0 if the batch is not eligible, < 0 on error.
static int try_launch_fp8_store3(b200kv_ctx* ctx, const uint8_t* dev_table, const TableView& tv, uint32_t n_chunks,
                                 const uint32_t* chunk_ntok, cudaStream_t s) {
  const Geometry& g = ctx->g;
  const uint32_t head_bytes = g.D * 2, rv = head_bytes >> 4;
  const uint64_t unit_bytes = static_cast<uint64_t>(g.C) * head_bytes;
  if (ctx->fp8_store_ver != 3 || g.elem != 2 || unit_bytes > kS3MaxUnitBytes || (head_bytes & 15) || rv == 0 ||
      (kS3GroupThreads % rv) != 0 || (static_cast<uint64_t>(g.bs) * head_bytes) % 128 != 0 || g.C / g.bs > 32 * 8 ||
      (g.hnd && (kS3GroupThreads / rv) % g.bs != 0))
    return 0;
  const bool use_tmap = !g.hnd || env_int("B200KV_FP8_TMAP_HND", 0);
  if (use_tmap && (!ctx->d_tmaps || ctx->tmap_status != "ok")) return 0;
  const Run* runs = reinterpret_cast<const Run*>(tv.slot->host + tv.runs_off);
  const uint32_t* offs = reinterpret_cast<const uint32_t*>(tv.slot->host + tv.offs_off);
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const uint32_t base = c * g.C;
    const uint32_t n_valid = chunk_ntok[c];
    if (n_valid == 0 || n_valid > g.C) return 0;
    const uint32_t nb = (n_valid + g.bs - 1) / g.bs;
    if (offs[c + 1] - offs[c] != nb) return 0;
    for (uint32_t j = 0; j < nb; ++j) {
      const Run& r = runs[offs[c] + j];
      if (r.b != static_cast<int32_t>(base + j * g.bs) || (static_cast<uint32_t>(r.a) % g.bs) != 0 ||
          static_cast<uint32_t>(r.n) != std::min<uint32_t>(g.bs, n_valid - j * g.bs))
        return 0;
    }
  }
  Fp8Store3Params p{};
  p.paged = local_side(ctx);
  p.runs = reinterpret_cast<const Run*>(dev_table + tv.runs_off);
  p.chunk_run_off = reinterpret_cast<const uint32_t*>(dev_table + tv.offs_off);
  p.chunk_addrs = reinterpret_cast<const uint64_t*>(dev_table + tv.addrs_off);
  p.tmaps = use_tmap ? ctx->d_tmaps : nullptr;
  p.n_chunks = n_chunks;
  p.n_planes = g.planes;
  p.chunk_tokens = g.C;
  p.n_tokens = 0;   // unused: a chunk's token count comes from its runs
  p.n_heads = g.H;
  p.head_bytes = head_bytes;
  p.slab_q_bytes = g.slab_bytes;
  p.scales_off = g.scales_off;
  p.hnd = g.hnd ? 1u : 0u;
  p.total_units = n_chunks * g.planes * g.H;
  const size_t smem = static_cast<size_t>(kS3Stages) * unit_bytes;
  static bool attr_set[8] = {false};
  const int dev = ctx->cfg.device & 7;
  if (!attr_set[dev]) {
    CU_TRY(cudaFuncSetAttribute(kv_fp8_store3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Stages * kS3MaxUnitBytes));
    attr_set[dev] = true;
  }
  const uint32_t grid = std::min<uint32_t>(p.total_units, static_cast<uint32_t>(ctx->sm_count));
  if (grid == 0) return 1;
  kv_fp8_store3_kernel<<<grid, kS3Threads, smem, s>>>(p);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return 1;
}

// chunk_ntok[c] = tokens of chunk c of this launch (chunks of several requests may be partial anywhere)
int launch_fp8_store(b200kv_ctx* ctx, const uint8_t* dev_table, const TableView& tv,
                     uint32_t n_chunks, const uint32_t* chunk_ntok, cudaStream_t s) {
  {
    const int rc3 = try_launch_fp8_store3(ctx, dev_table, tv, n_chunks, chunk_ntok, s);
    if (rc3 != 0) return rc3 < 0 ? rc3 : B200KV_OK;
  }
  const uint32_t n_tokens = 0;  // unused by the kernels (see above)
  Fp8StoreParams p{};
  p.paged = local_side(ctx);
  p.runs = reinterpret_cast<const Run*>(dev_table + tv.runs_off);
  p.chunk_run_off = reinterpret_cast<const uint32_t*>(dev_table + tv.offs_off);
  p.chunk_addrs = reinterpret_cast<const uint64_t*>(dev_table + tv.addrs_off);
  p.n_chunks = n_chunks;
  p.n_planes = ctx->g.planes;
  p.chunk_tokens = ctx->g.C;
  p.n_tokens = n_tokens;
  p.n_heads = ctx->g.H;
  p.head_bytes = ctx->g.D * 2;
  p.slab_q_bytes = ctx->g.slab_bytes;
  p.scales_off = ctx->g.scales_off;
  p.hnd = ctx->g.hnd ? 1u : 0u;
  const size_t smem = static_cast<size_t>(ctx->g.C / kCluster) * ctx->g.token_bytes;
  if (smem > 200 * 1024) return B200KV_ENOTSUP;
  const uint32_t grid = kCluster * n_chunks * ctx->g.planes;
  if (grid == 0) return B200KV_OK;
  static bool attr_set[8] = {false};
  const int dev = ctx->cfg.device & 7;
  if (!attr_set[dev]) {
    CU_TRY(cudaFuncSetAttribute(kv_fp8_store_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU_TRY(cudaFuncSetAttribute(kv_fp8_store_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev] = true;
  }
  if (ctx->fp8_two_pass && ctx->g.C / kCluster <= static_cast<uint32_t>(kMaxWindow)) kv_fp8_store2_kernel<<<grid, 256, 0, s>>>(p);
  else if (ctx->fp8_threads == 512) kv_fp8_store_kernel<512><<<grid, 512, smem, s>>>(p);
  else kv_fp8_store_kernel<256><<<grid, 256, smem, s>>>(p);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return B200KV_OK;
}

int launch_fp8_load(b200kv_ctx* ctx, const uint8_t* dev_table, const TableView& tv,
                    size_t run_begin, size_t n_runs, cudaStream_t s, uint32_t plane_begin = 0,
                    uint32_t n_planes = 0) {
  Fp8LoadParams p{};
  p.paged = local_side(ctx);
  p.runs = reinterpret_cast<const Run*>(dev_table + tv.runs_off) + run_begin;
  p.chunk_addrs = reinterpret_cast<const uint64_t*>(dev_table + tv.addrs_off);
  p.n_runs = static_cast<uint32_t>(n_runs);
  p.n_planes = n_planes ? n_planes : ctx->g.planes;
  p.plane_begin = plane_begin;
  p.chunk_tokens = ctx->g.C;
  p.n_heads = ctx->g.H;
  p.head_bytes = ctx->g.D * 2;
  p.slab_q_bytes = ctx->g.slab_bytes;
  p.scales_off = ctx->g.scales_off;
  p.hnd = ctx->g.hnd ? 1u : 0u;
  p.total_units = p.n_runs * p.n_planes;
  const size_t smem = static_cast<size_t>(ctx->g.bs) * ctx->g.fmt_token_bytes;
  if (smem > 48 * 1024) return B200KV_ENOTSUP;
  uint32_t grid = std::min<uint32_t>(p.total_units, static_cast<uint32_t>(ctx->sm_count) * 8u);
  if (grid == 0) return B200KV_OK;
  kv_fp8_load_kernel<<<grid, kFp8Threads, smem, s>>>(p);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return B200KV_OK;
}

// chunk-side copy between staging and the pinned pool; partial chunks move only their tokens.
int launch_q4(b200kv_ctx* ctx, bool store, const uint8_t* dev_table, const TableView& tv, size_t run_begin,
              size_t n_runs, cudaStream_t s, uint32_t plane_begin = 0, uint32_t n_planes = 0) {
  Q4Params p{};
  p.paged = local_side(ctx);
  p.runs = reinterpret_cast<const Run*>(dev_table + tv.runs_off) + run_begin;
  p.chunk_addrs = reinterpret_cast<const uint64_t*>(dev_table + tv.addrs_off);
  p.n_runs = static_cast<uint32_t>(n_runs);
  p.n_planes = n_planes ? n_planes : ctx->g.planes;
  p.plane_begin = plane_begin;
  p.chunk_tokens = ctx->g.C;
  p.n_heads = ctx->g.H;
  p.head_bytes = ctx->g.D * 2;
  p.slab_bytes = ctx->g.slab_bytes;
  p.rec_bytes = ctx->g.fmt_token_bytes;
  p.hnd = ctx->g.hnd ? 1u : 0u;
  p.total_units = p.n_runs * p.n_planes;
  const uint32_t grid = std::min<uint32_t>(p.total_units, static_cast<uint32_t>(ctx->sm_count) * 8u);
  if (grid == 0) return B200KV_OK;
  if (store) kv_q4_store_kernel<<<grid, 256, 0, s>>>(p);
  else kv_q4_load_kernel<<<grid, 256, 0, s>>>(p);
  CU_TRY(cudaGetLastError());
  ++ctx->stats.n_kernel_launches;
  return B200KV_OK;
}

int copy_chunk(b200kv_ctx* ctx, void* dst, const void* src, uint32_t n_tok, cudaMemcpyKind kind,
               cudaStream_t s) {
  const Geometry& g = ctx->g;
  uint64_t moved;
  if (n_tok == g.C) {
    CU_TRY(cudaMemcpyAsync(dst, src, g.chunk_bytes, kind, s));
    moved = g.chunk_bytes;
  } else {
    // HND keeps whole tiles: a ragged tail still occupies its last tile across all heads
    const uint32_t tok_span = (g.hnd && ctx->cfg.format != B200KV_FMT_Q4) ? (n_tok + g.bs - 1) / g.bs * g.bs : n_tok;
    const size_t width = static_cast<size_t>(tok_span) * g.fmt_token_bytes;
    CU_TRY(cudaMemcpy2DAsync(dst, g.slab_bytes, src, g.slab_bytes, width, g.planes, kind, s));
    moved = width * g.planes;
    if (ctx->cfg.format == B200KV_FMT_FP8) {
      const size_t sb = static_cast<size_t>(g.planes) * g.H * sizeof(float);
      CU_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + g.scales_off,
                             static_cast<const uint8_t*>(src) + g.scales_off, sb, kind, s));
      moved += sb;
    }
  }
  if (kind == cudaMemcpyDeviceToHost) ctx->stats.d2h_bytes += moved;
  else ctx->stats.h2d_bytes += moved;
  return B200KV_OK;
}

// planes [pb, pb+np) of one chunk between the pinned pool and staging (layer-wise loads); the
// FP8 scales travel once, with the first group.
int copy_chunk_planes(b200kv_ctx* ctx, void* dst, const void* src, uint32_t n_tok, uint32_t pb, uint32_t np,
                      bool with_scales, cudaMemcpyKind kind, cudaStream_t s) {
  const Geometry& g = ctx->g;
  uint8_t* d = static_cast<uint8_t*>(dst) + static_cast<uint64_t>(pb) * g.slab_bytes;
  const uint8_t* sp = static_cast<const uint8_t*>(src) + static_cast<uint64_t>(pb) * g.slab_bytes;
  uint64_t moved;
  if (n_tok == g.C) {
    moved = static_cast<uint64_t>(np) * g.slab_bytes;
    CU_TRY(cudaMemcpyAsync(d, sp, moved, kind, s));
  } else {
    const uint32_t tok_span = (g.hnd && ctx->cfg.format != B200KV_FMT_Q4) ? (n_tok + g.bs - 1) / g.bs * g.bs : n_tok;
    const size_t width = static_cast<size_t>(tok_span) * g.fmt_token_bytes;
    CU_TRY(cudaMemcpy2DAsync(d, g.slab_bytes, sp, g.slab_bytes, width, np, kind, s));
    moved = width * np;
  }
  if (with_scales && ctx->cfg.format == B200KV_FMT_FP8) {
    const size_t sb = static_cast<size_t>(g.planes) * g.H * sizeof(float);
    CU_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + g.scales_off, static_cast<const uint8_t*>(src) + g.scales_off,
                           sb, kind, s));
    moved += sb;
  }
  if (kind == cudaMemcpyDeviceToHost) ctx->stats.d2h_bytes += moved;
  else ctx->stats.h2d_bytes += moved;
  return B200KV_OK;
}

void reap(b200kv_ctx* ctx) {
  for (auto it = ctx->ops.begin(); it != ctx->ops.end();) {
    Op* op = it->second.get();
    if (cudaEventQuery(op->done) == cudaSuccess && op->host_done.load(std::memory_order_acquire)) {
      destroy_op_events(op);
      it = ctx->ops.erase(it);
    } else {
      ++it;
    }
  }
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int b200kv_abi_version(void) { return B200KV_ABI_VERSION; }

extern "C" const char* b200kv_last_error(void) { return g_err; }

extern "C" const char* b200kv_strerror(int err) {
  switch (err) {
    case B200KV_OK: return "ok";
    case B200KV_EINVAL: return "invalid argument";
    case B200KV_ENOMEM: return "out of memory";
    case B200KV_ENODEV: return "CUDA device unavailable or CUDA call failed";
    case B200KV_ENOENT: return "not found";
    case B200KV_EEXIST: return "already exists";
    case B200KV_ENOSPC: return "pool full";
    case B200KV_ENOTSUP: return "not supported";
    case B200KV_EBUSY: return "busy";
    default: return std::strerror(-err);
  }
}

extern "C" int64_t b200kv_engine_chunk_bytes(const b200kv_engine_config* cfg) {
  if (!cfg) return B200KV_EINVAL;
  Geometry g;
  const int rc = make_geometry(cfg, &g);
  if (rc) return rc;
  return static_cast<int64_t>(g.chunk_bytes);
}

static int engine_create_impl(const b200kv_engine_config* cfg, b200kv_pool* pool, b200kv_ctx** out);

extern "C" int b200kv_engine_destroy(b200kv_ctx* ctx);

extern "C" int b200kv_engine_numa_placement(b200kv_ctx* ctx, char* buf, uint64_t n) {
  if (!ctx || !buf || n == 0) return B200KV_EINVAL;
  std::snprintf(buf, n, "%s", ctx->numa_placement.c_str());
  return B200KV_OK;
}

// ---- host-page placement of the pinned pool --------------------------------------------------
// On a two-socket HGX box every DMA of a GPU whose pool pages sit on the other socket crosses the
// inter-socket link twice as slowly as the PCIe link it came over.  The pages of the pool segment are
// physically allocated when the first engine pins them (cudaHostRegister faults them in), so the
// policy is applied to the mapping right before that: mbind() on the shared mapping (shmem keeps the
// policy with the object, so later attachers inherit it).  Where mbind is filtered (containers without
// CAP_SYS_NICE) the fallback for NUMA_LOCAL is to run the pinning call on the node's CPUs.
static int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
  for (char* p = bus; *p; ++p) *p = static_cast<char>(std::tolower(*p));
  char path[128];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = std::fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (std::fscanf(f, "%d", &node) != 1) node = -1;
  std::fclose(f);
  return node;
}

static int online_numa_nodes(unsigned long* mask, int* max_node) {
  *mask = 0;
  *max_node = -1;
  FILE* f = std::fopen("/sys/devices/system/node/online", "r");
  if (!f) return -1;
  char buf[128] = {0};
  if (!std::fgets(buf, sizeof(buf), f)) { std::fclose(f); return -1; }
  std::fclose(f);
  for (char* p = buf; *p && *p != '\n';) {       // "0-1" or "0,2-3"
    char* e = nullptr;
    long a = std::strtol(p, &e, 10), b = a;
    if (e == p) break;
    if (*e == '-') { p = e + 1; b = std::strtol(p, &e, 10); }
    for (long n = a; n <= b && n < 64; ++n) { *mask |= 1ul << n; *max_node = std::max<int>(*max_node, static_cast<int>(n)); }
    p = (*e == ',') ? e + 1 : e;
  }
  return *max_node >= 0 ? 0 : -1;
}

struct NumaPinScope {       // restores the thread's CPU affinity if the fallback narrowed it
  cpu_set_t saved;
  bool narrowed = false;
  ~NumaPinScope() { if (narrowed) sched_setaffinity(0, sizeof(saved), &saved); }
};

// returns a short description for the log / stats ("mbind preferred node 1", "affinity node 0", ...)
static std::string place_pool_pages(void* base, uint64_t bytes, int device, int policy, NumaPinScope* scope) {
  const char* e = std::getenv("B200KV_POOL_NUMA");       // local | interleave | off overrides the config
  if (e) policy = !std::strcmp(e, "interleave") ? B200KV_NUMA_INTERLEAVE : !std::strcmp(e, "off") ? B200KV_NUMA_OFF : B200KV_NUMA_LOCAL;
  if (policy == B200KV_NUMA_OFF) return "first touch (off)";
  unsigned long online = 0;
  int max_node = -1;
  if (online_numa_nodes(&online, &max_node) != 0 || max_node == 0) return "single NUMA node";
  constexpr int kPreferred = 1, kInterleave = 3;    // MPOL_* of <linux/mempolicy.h>
  if (policy == B200KV_NUMA_INTERLEAVE) {
    const long rc = syscall(SYS_mbind, base, static_cast<unsigned long>(bytes), kInterleave, &online, 65ul, 0u);
    return rc == 0 ? "mbind interleave over online nodes" : std::string("mbind refused (") + std::strerror(errno) + "): first touch";
  }
  const int node = gpu_numa_node(device);
  if (node < 0 || node > max_node) return "GPU NUMA node unknown: first touch";
  unsigned long mask = 1ul << node;
  if (syscall(SYS_mbind, base, static_cast<unsigned long>(bytes), kPreferred, &mask, 65ul, 0u) == 0)
    return "mbind preferred node " + std::to_string(node);
  const int mbind_errno = errno;
  // fallback: fault the pages in from a CPU of that node
  char path[96];
  std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = std::fopen(path, "r");
  cpu_set_t want;
  CPU_ZERO(&want);
  if (f) {
    char buf[512] = {0};
    if (std::fgets(buf, sizeof(buf), f)) {
      for (char* p = buf; *p && *p != '\n';) {
        char* q = nullptr;
        long a = std::strtol(p, &q, 10), b = a;
        if (q == p) break;
        if (*q == '-') { p = q + 1; b = std::strtol(p, &q, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(static_cast<int>(c), &want);
        p = (*q == ',') ? q + 1 : q;
      }
    }
    std::fclose(f);
  }
  if (CPU_COUNT(&want) > 0 && sched_getaffinity(0, sizeof(scope->saved), &scope->saved) == 0) {
    cpu_set_t both;
    CPU_AND(&both, &want, &scope->saved);
    if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0) {
      scope->narrowed = true;
      return std::string("mbind refused (") + std::strerror(mbind_errno) + "): pinned from the CPUs of node " + std::to_string(node);
    }
  }
  return std::string("mbind refused (") + std::strerror(mbind_errno) + "): first touch";
}

extern "C" int b200kv_engine_create(const b200kv_engine_config* cfg, b200kv_pool* pool,
                                    b200kv_ctx** out) {
  if (!cfg || !out) return B200KV_EINVAL;
  *out = nullptr;
  b200kv_ctx* ctx = nullptr;
  const int rc = engine_create_impl(cfg, pool, &ctx);
  if (rc != B200KV_OK) {
    if (ctx) b200kv_engine_destroy(ctx);  // releases whatever had been created (streams, buffers, pinning)
    return rc;
  }
  *out = ctx;
  return B200KV_OK;
}

static int engine_create_impl(const b200kv_engine_config* cfg, b200kv_pool* pool, b200kv_ctx** out) {
  Geometry g;
  int rc = make_geometry(cfg, &g);
  if (rc) return rc;
  if (cfg->block_stride_bytes < static_cast<uint64_t>(g.bs) * g.token_bytes ||
      cfg->block_stride_bytes % 16 || cfg->n_blocks == 0 ||
      cfg->n_blocks * g.bs > 0x7fffffffull)
    return B200KV_EINVAL;

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    std::snprintf(g_err, sizeof(g_err),
                  "no CUDA device: libb200kv has no CPU fallback (the oracle/ directory is test "
                  "infrastructure only)");
    return B200KV_ENODEV;
  }
  if (cfg->device < 0 || cfg->device >= n_dev) return B200KV_EINVAL;
  DeviceGuard dg(cfg->device);
  if (!dg.ok) return B200KV_ENODEV;

  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    std::snprintf(g_err, sizeof(g_err), "device %d is sm_%d%d; this library is built for sm_100a only",
                  cfg->device, prop.major, prop.minor);
    return B200KV_ENOTSUP;
  }

  b200kv_ctx* ctx = new (std::nothrow) b200kv_ctx();
  if (!ctx) return B200KV_ENOMEM;
  *out = ctx;  // from here on the caller destroys it on failure
  ctx->cfg = *cfg;
  ctx->g = g;
  ctx->pool = pool;
  ctx->sm_count = prop.multiProcessorCount;

  // launch shape of the bulk kernel
  ctx->S = cfg->stages > 0 ? cfg->stages : env_int("B200KV_STAGES", 2);
  ctx->LAG = env_int("B200KV_LAG", ctx->S / 2);
  ctx->ctas_per_sm = cfg->ctas_per_sm > 0 ? cfg->ctas_per_sm : env_int("B200KV_CTAS_PER_SM", 1);
  ctx->fp8_threads = env_int("B200KV_FP8_THREADS", 256) == 512 ? 512 : 256;
  ctx->fp8_two_pass = env_int("B200KV_FP8_2PASS", 0) != 0;
  ctx->fp8_store_ver = env_int("B200KV_FP8_STORE", 3);
  const uint32_t stage_max = static_cast<uint32_t>(env_int("B200KV_STAGE_KB", 32)) << 10;
  if (g.token_bytes > stage_max || stage_max > kStageMax * 2) return B200KV_ENOTSUP;
  ctx->piece_tokens = std::min<uint32_t>(g.bs, stage_max / g.token_bytes);
  ctx->pieces = (g.bs + ctx->piece_tokens - 1) / ctx->piece_tokens;
  ctx->stage_bytes = ctx->piece_tokens * g.token_bytes;
  if (256 + static_cast<size_t>(ctx->S) * ctx->stage_bytes > 227u * 1024) return B200KV_EINVAL;

  int lo_pri = 0, hi_pri = 0;
  CU_TRY(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
  CU_TRY(cudaStreamCreateWithPriority(&ctx->s_gather, cudaStreamNonBlocking, lo_pri));
  CU_TRY(cudaStreamCreateWithPriority(&ctx->s_d2h, cudaStreamNonBlocking, lo_pri));
  CU_TRY(cudaStreamCreateWithPriority(&ctx->s_h2d, cudaStreamNonBlocking, hi_pri));
  CU_TRY(cudaStreamCreateWithPriority(&ctx->s_scatter, cudaStreamNonBlocking, hi_pri));

  CU_TRY(cudaMalloc(&ctx->d_bases, sizeof(uint64_t) * g.planes));

  const uint64_t n_stage = cfg->staging_bytes / g.chunk_bytes;
  if (n_stage > 0) {
    CU_TRY(cudaMalloc(&ctx->d_staging, n_stage * g.chunk_bytes));
    ctx->stage.resize(n_stage);
    for (auto& s : ctx->stage) CU_TRY(cudaEventCreateWithFlags(&s.free_ev, cudaEventDisableTiming));
    if (pool && n_stage < 2) return B200KV_EINVAL;  // one half each for stores and loads
    ctx->n_store_slots = static_cast<uint32_t>(n_stage / 2);
  } else if (pool) {
    return B200KV_EINVAL;  // store/load need at least one staging chunk
  }
  for (auto& t : ctx->tables) {
    CU_TRY(cudaHostAlloc(&t.host, kTableBytes, cudaHostAllocDefault));
    CU_TRY(cudaMalloc(&t.dev, kTableBytes));
    CU_TRY(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&t.done_ev, cudaEventDisableTiming));
  }

  if (pool) {
    void* base = nullptr;
    uint64_t bytes = 0;
    rc = b200kv_pool_region(pool, &base, &bytes);
    if (rc) return rc;
    b200kv_pool_stats ps;
    b200kv_pool_get_stats(pool, &ps);
    if (ps.slot_bytes < g.chunk_bytes) return B200KV_EINVAL;
    // Pin the (possibly shared) pool for this process so D2H/H2D run at full PCIe speed and
    // asynchronously.  Portable: every replica on the box registers the same segment.
    NumaPinScope numa_scope;
    ctx->numa_placement = place_pool_pages(base, bytes, cfg->device, cfg->numa_policy, &numa_scope);
    if (env_int("B200KV_VERBOSE", 0)) std::fprintf(stderr, "b200kv: pool pages: %s\n", ctx->numa_placement.c_str());
    const cudaError_t re = cudaHostRegister(base, bytes, cudaHostRegisterPortable);
    if (re == cudaSuccess) {
      ctx->pool_registered = true;
      ctx->pool_base = base;
    } else if (re == cudaErrorHostMemoryAlreadyRegistered) {
      cudaGetLastError();  // another engine of this process (e.g. a second GPU) pinned it already
    } else {
      set_err("cudaHostRegister(pool)", re, __LINE__);
      return B200KV_ENODEV;
    }
  }
  return B200KV_OK;
}

extern "C" int b200kv_wait_all(b200kv_ctx* ctx);

extern "C" int b200kv_engine_destroy(b200kv_ctx* ctx) {
  if (!ctx) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  b200kv_wait_all(ctx);
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    reap(ctx);
    for (auto& kv : ctx->ops) destroy_op_events(kv.second.get());
    ctx->ops.clear();
  }
  if (ctx->pool_registered) cudaHostUnregister(ctx->pool_base);
  for (auto& p : ctx->peers) {
    if (!p.valid) continue;
    for (void* m : p.opened) cudaIpcCloseMemHandle(m);
    cudaFree(p.d_bases);
  }
  for (auto& t : ctx->tables) {
    if (t.host) cudaFreeHost(t.host);
    if (t.dev) cudaFree(t.dev);
    if (t.ev) cudaEventDestroy(t.ev);
    if (t.done_ev) cudaEventDestroy(t.done_ev);
  }
  for (auto& s : ctx->stage)
    if (s.free_ev) cudaEventDestroy(s.free_ev);
  for (auto& t : ctx->timing)
    for (auto& e : t.ev) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
  if (ctx->d_staging) cudaFree(ctx->d_staging);
  for (void* m : ctx->tier_imports) cudaIpcCloseMemHandle(m);
  if (ctx->tier_base) cudaFree(ctx->tier_base);
  if (ctx->d_bases) cudaFree(ctx->d_bases);
  if (ctx->d_tmaps) cudaFree(ctx->d_tmaps);
  for (cudaStream_t s : {ctx->s_gather, ctx->s_d2h, ctx->s_h2d, ctx->s_scatter})
    if (s) cudaStreamDestroy(s);
  delete ctx;
  return B200KV_OK;
}

// One 4-D tensor map per plane over the paged cache, so that TMA can fetch "the bs rows of head h in block b"
// — a [bs][D] box whose rows are a token apart — as ONE copy (SASS UTMALDG).  Dimensions are ordered by
// ascending stride: NHD pages {D, H, bs, NB} (box {D, 1, bs, 1}), HND pages {D, bs, H, NB} (box {D, bs, 1, 1});
// the box always lands as [bs][D] in shared memory.  Failure is not fatal: the FP8 store then runs its
// cluster kernel.
static void build_tensor_maps(b200kv_ctx* ctx, const std::vector<uint64_t>& bases) {
  const Geometry& g = ctx->g;
  if (ctx->cfg.format != B200KV_FMT_FP8 || g.elem != 2) { ctx->tmap_status = "not needed"; return; }
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    ctx->tmap_status = "cuTensorMapEncodeTiled unavailable";
    return;
  }
  if (g.D > 256 || g.bs > 256) { ctx->tmap_status = "box too large"; return; }
  std::vector<CUtensorMap> maps(g.planes);
  const cuuint64_t row = static_cast<cuuint64_t>(g.D) * g.elem, tok = g.token_bytes, blk = ctx->cfg.block_stride_bytes;
  for (uint32_t pl = 0; pl < g.planes; ++pl) {
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    dims[0] = g.D;
    dims[3] = ctx->cfg.n_blocks;
    if (g.hnd) { dims[1] = g.bs; dims[2] = g.H; strides[0] = row; strides[1] = row * g.bs; box[1] = g.bs; box[2] = 1; }
    else       { dims[1] = g.H; dims[2] = g.bs; strides[0] = row; strides[1] = tok;        box[1] = 1; box[2] = g.bs; }
    strides[2] = blk;
    box[0] = g.D;
    box[3] = 1;
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&maps[pl], CU_TENSOR_MAP_DATA_TYPE_UINT16, 4,
                                                      reinterpret_cast<void*>(bases[pl]), dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ctx->tmap_status = "cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")";
      return;
    }
  }
  if (!ctx->d_tmaps && cudaMalloc(&ctx->d_tmaps, sizeof(CUtensorMap) * g.planes) != cudaSuccess) {
    cudaGetLastError();
    ctx->d_tmaps = nullptr;
    ctx->tmap_status = "cudaMalloc failed";
    return;
  }
  if (cudaMemcpy(ctx->d_tmaps, maps.data(), sizeof(CUtensorMap) * g.planes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(ctx->d_tmaps);
    ctx->d_tmaps = nullptr;
    ctx->tmap_status = "cudaMemcpy failed";
    return;
  }
  ctx->tmap_status = "ok";
}

extern "C" int b200kv_register_kv(b200kv_ctx* ctx, const void* const* k_ptrs,
                                  const void* const* v_ptrs) {
  if (!ctx || !k_ptrs || !v_ptrs) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::vector<uint64_t> h(ctx->g.planes);
  ctx->h_k.assign(k_ptrs, k_ptrs + ctx->g.L);
  ctx->h_v.assign(v_ptrs, v_ptrs + ctx->g.L);
  for (uint32_t l = 0; l < ctx->g.L; ++l) {
    const uint64_t k = reinterpret_cast<uint64_t>(k_ptrs[l]);
    const uint64_t v = reinterpret_cast<uint64_t>(v_ptrs[l]);
    if (!k || !v || (k % 16) || (v % 16)) return B200KV_EINVAL;
    h[2 * l] = k;
    h[2 * l + 1] = v;
  }
  CU_TRY(cudaMemcpy(ctx->d_bases, h.data(), sizeof(uint64_t) * h.size(), cudaMemcpyHostToDevice));
  build_tensor_maps(ctx, h);
  if (env_int("B200KV_VERBOSE", 0)) std::fprintf(stderr, "b200kv: tensor maps: %s\n", ctx->tmap_status.c_str());
  ctx->kv_registered = true;
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident gather / scatter
// ------------------------------------------------------------------------------------------------
static int gather_scatter(b200kv_ctx* ctx, const int64_t* slots, int64_t n_tokens, void* dev_chunks,
                          void* stream, bool is_gather, const uint64_t* chunk_ptrs = nullptr) {
  if (!ctx || !slots || n_tokens <= 0 || (!dev_chunks && !chunk_ptrs)) return B200KV_EINVAL;
  if (!ctx->kv_registered) return B200KV_EINVAL;
  if (reinterpret_cast<uint64_t>(dev_chunks) % 16) return B200KV_EINVAL;
  if (chunk_ptrs) {
    const int64_t n = (n_tokens + ctx->g.C - 1) / ctx->g.C;
    for (int64_t c = 0; c < n; ++c)
      if (!chunk_ptrs[c] || chunk_ptrs[c] % 16) return B200KV_EINVAL;
  }
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  const Geometry& g = ctx->g;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint32_t n_chunks = static_cast<uint32_t>((n_tokens + g.C - 1) / g.C);

  std::vector<Run> runs, partial;
  const bool fp8 = ctx->cfg.format != B200KV_FMT_RAW;   // one sorted run list for every transformed format
  int rc = build_runs(ctx, slots, 0, n_tokens, 0, &runs, fp8 ? nullptr : &partial);
  if (rc) return rc;
  const size_t n_full = runs.size(), n_part = partial.size();
  runs.insert(runs.end(), partial.begin(), partial.end());
  TableView tv;
  rc = table_acquire(ctx, runs.size(), n_chunks, &tv);
  if (rc) return rc;
  std::memcpy(tv.slot->host + tv.runs_off, runs.data(), runs.size() * sizeof(Run));
  uint64_t* addrs = reinterpret_cast<uint64_t*>(tv.slot->host + tv.addrs_off);
  uint32_t* offs = reinterpret_cast<uint32_t*>(tv.slot->host + tv.offs_off);
  size_t r = 0;
  for (uint32_t c = 0; c < n_chunks; ++c) {
    addrs[c] = chunk_ptrs ? chunk_ptrs[c]
                          : reinterpret_cast<uint64_t>(dev_chunks) + static_cast<uint64_t>(c) * g.chunk_bytes;
    offs[c] = static_cast<uint32_t>(r);
    while (r < runs.size() && static_cast<uint32_t>(runs[r].b) / g.C == c) ++r;
  }
  offs[n_chunks] = static_cast<uint32_t>(r);
  rc = table_upload(ctx, tv, s);
  if (rc) return rc;

  const int which = is_gather ? 0 : 1;
  timing_reset(ctx, which);
  rc = timing_begin(ctx, which, s);
  if (rc) return rc;
  if (ctx->cfg.format == B200KV_FMT_FP8) {
    std::vector<uint32_t> cn(n_chunks);
    for (uint32_t c = 0; c < n_chunks; ++c)
      cn[c] = static_cast<uint32_t>(std::min<int64_t>(g.C, n_tokens - static_cast<int64_t>(c) * g.C));
    rc = is_gather ? launch_fp8_store(ctx, tv.slot->dev, tv, n_chunks, cn.data(), s)
                   : launch_fp8_load(ctx, tv.slot->dev, tv, 0, runs.size(), s);
  } else if (ctx->cfg.format == B200KV_FMT_Q4) {
    rc = launch_q4(ctx, is_gather, tv.slot->dev, tv, 0, runs.size(), s);
  } else {
    rc = is_gather ? launch_copy_runs<kStore>(ctx, tv.slot->dev, tv, 0, n_full, n_part, s)
                   : launch_copy_runs<kLoad>(ctx, tv.slot->dev, tv, 0, n_full, n_part, s);
  }
  if (rc) return rc;
  rc = timing_end(ctx, which, s);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(tv.slot->done_ev, s));
  if (is_gather) ctx->stats.n_stored_tokens += n_tokens; else ctx->stats.n_loaded_tokens += n_tokens;
  return B200KV_OK;
}

extern "C" int b200kv_gather(b200kv_ctx* ctx, const int64_t* slot_mapping, int64_t n_tokens,
                             void* dev_chunks, void* stream) {
  return gather_scatter(ctx, slot_mapping, n_tokens, dev_chunks, stream, true);
}

extern "C" int b200kv_scatter(b200kv_ctx* ctx, const int64_t* slot_mapping, int64_t n_tokens,
                              const void* dev_chunks, void* stream) {
  return gather_scatter(ctx, slot_mapping, n_tokens, const_cast<void*>(dev_chunks), stream, false);
}

extern "C" int b200kv_gather_chunks(b200kv_ctx* ctx, const int64_t* slot_mapping, int64_t n_tokens,
                                    const uint64_t* chunk_ptrs, void* stream) {
  return gather_scatter(ctx, slot_mapping, n_tokens, nullptr, stream, true, chunk_ptrs);
}

extern "C" int b200kv_scatter_chunks(b200kv_ctx* ctx, const int64_t* slot_mapping, int64_t n_tokens,
                                     const uint64_t* chunk_ptrs, void* stream) {
  return gather_scatter(ctx, slot_mapping, n_tokens, nullptr, stream, false, chunk_ptrs);
}

// ------------------------------------------------------------------------------------------------
// device chunk tier: a buffer of chunk-format slots in HBM, exported to peer replicas over CUDA IPC.
// The index (which key sits in which slot, LRU, pins) is a b200kv_pool in shm owned by the host side.
// ------------------------------------------------------------------------------------------------
extern "C" int b200kv_tier_create(b200kv_ctx* ctx, uint32_t n_slots, uint64_t* base_out) {
  if (!ctx || !n_slots || !base_out) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ctx->tier_base) return B200KV_EEXIST;
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, static_cast<size_t>(n_slots) * ctx->g.chunk_bytes));
  ctx->tier_base = p;
  ctx->tier_slots = n_slots;
  *base_out = reinterpret_cast<uint64_t>(p);
  return B200KV_OK;
}

extern "C" int b200kv_tier_export(b200kv_ctx* ctx, b200kv_ipc_desc* desc_out) {
  if (!ctx || !desc_out || !ctx->tier_base) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::memset(desc_out, 0, sizeof(*desc_out));
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, ctx->tier_base));
  std::memcpy(desc_out->handle, &h, 64);
  desc_out->offset = 0;
  desc_out->alloc_bytes = static_cast<uint64_t>(ctx->tier_slots) * ctx->g.chunk_bytes;
  return B200KV_OK;
}

extern "C" int b200kv_tier_import(b200kv_ctx* ctx, const b200kv_ipc_desc* desc, uint64_t* mapped_base_out) {
  if (!ctx || !desc || !mapped_base_out) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, desc->handle, 64);
  void* base = nullptr;
  CU_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->tier_imports.push_back(base);
  *mapped_base_out = reinterpret_cast<uint64_t>(base) + desc->offset;
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// store: paged HBM -> staging (kernel) -> pinned pool (DMA)
// ------------------------------------------------------------------------------------------------
// Chunk c of the op holds chunk_tokens[c] tokens whose slots are slot_mapping[c*C .. c*C + chunk_tokens[c]): a
// single request (every chunk full but the last) or the chunks of SEVERAL requests back to back, each
// request's tail padded to the chunk size (b200kv_store_batch_async: one table, one launch per staging
// batch and one ticket for a whole engine step).
static int store_impl(b200kv_ctx* ctx, const uint64_t* keys, const int32_t* chunk_tokens, int32_t n_chunks,
                      const int64_t* slot_mapping, void* compute_stream, uint64_t* ticket) {
  if (!ctx || !keys || !chunk_tokens || !slot_mapping || n_chunks <= 0 || !ticket) return B200KV_EINVAL;
  if (!ctx->pool || !ctx->kv_registered || ctx->stage.empty()) return B200KV_EINVAL;
  const Geometry& g = ctx->g;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  reap(ctx);
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream);
  *ticket = 0;

  // 0. validate every slot BEFORE touching the pool: a malformed mapping must not leave chunks
  //    reserved-but-never-committed
  {
    const int64_t max_slot = static_cast<int64_t>(ctx->cfg.n_blocks) * g.bs;
    for (int32_t c = 0; c < n_chunks; ++c) {
      if (chunk_tokens[c] <= 0 || chunk_tokens[c] > static_cast<int32_t>(g.C)) return B200KV_EINVAL;
      const int64_t* sm = slot_mapping + static_cast<int64_t>(c) * g.C;
      for (int32_t i = 0; i < chunk_tokens[c]; ++i)
        if (sm[i] < 0 || sm[i] >= max_slot) return B200KV_EINVAL;
    }
  }
  // 1. reserve pool slots; chunks already present (or not placeable) are skipped
  struct Todo { int32_t c; uint32_t slot; uint32_t n_tok; };
  std::vector<Todo> todo;
  OpGuard guard(ctx, ctx->s_gather, ctx->s_d2h, nullptr, nullptr);   // undoes the reservations on any early return
  for (int32_t c = 0; c < n_chunks; ++c) {
    const uint32_t n_tok = static_cast<uint32_t>(chunk_tokens[c]);
    uint32_t slot = 0;
    const int rc = b200kv_pool_reserve(ctx->pool, keys[c], static_cast<int32_t>(n_tok), pool_fmt(ctx),
                                       ctx->cfg.owner, &slot);
    if (rc == B200KV_OK) { todo.push_back({c, slot, n_tok}); guard.reserved.push_back(keys[c]); }
    else if (rc != B200KV_EEXIST && rc != B200KV_ENOSPC) return rc;
  }
  ++ctx->stats.n_store_ops;
  if (todo.empty()) return B200KV_OK;

  guard.op.reset(new Op());
  Op* op = guard.op.get();
  op->id = ctx->next_ticket++;
  op->pool = ctx->pool;
  CU_TRY(cudaEventCreateWithFlags(&op->done, cudaEventDisableTiming));

  cudaEvent_t ev_compute;
  CU_TRY(cudaEventCreateWithFlags(&ev_compute, cudaEventDisableTiming));
  CU_TRY(cudaEventRecord(ev_compute, cs));
  CU_TRY(cudaStreamWaitEvent(ctx->s_gather, ev_compute, 0));
  CU_TRY(cudaEventDestroy(ev_compute));

  timing_reset(ctx, 0);
  const size_t n_stage = ctx->n_store_slots;
  for (size_t b0 = 0; b0 < todo.size(); b0 += n_stage) {
    const size_t nb = std::min(n_stage, todo.size() - b0);
    std::vector<Run> runs, partial;
    std::vector<uint32_t> offs(nb + 1);
    std::vector<uint64_t> addrs(nb);
    std::vector<uint32_t> sidx(nb), batch_ntok(nb);
    for (size_t i = 0; i < nb; ++i) {
      const Todo& t = todo[b0 + i];
      batch_ntok[i] = t.n_tok;
      offs[i] = static_cast<uint32_t>(runs.size());
      const int64_t tb = static_cast<int64_t>(t.c) * g.C;
      // dense op-relative token index: chunk i of this batch starts at i*C
      int rc = build_runs(ctx, slot_mapping, tb, tb + t.n_tok, tb - static_cast<int64_t>(i) * g.C, &runs,
                          ctx->cfg.format != B200KV_FMT_RAW ? nullptr : &partial);
      if (rc) return rc;
      sidx[i] = ctx->store_next;
      ctx->store_next = (ctx->store_next + 1) % static_cast<uint32_t>(n_stage);
      addrs[i] = reinterpret_cast<uint64_t>(ctx->d_staging) + static_cast<uint64_t>(sidx[i]) * g.chunk_bytes;
      if (ctx->stage[sidx[i]].used) CU_TRY(cudaStreamWaitEvent(ctx->s_gather, ctx->stage[sidx[i]].free_ev, 0));
    }
    offs[nb] = static_cast<uint32_t>(runs.size());
    const size_t n_full = runs.size(), n_part = partial.size();
    runs.insert(runs.end(), partial.begin(), partial.end());
    TableView tv;
    int rc = table_acquire(ctx, runs.size(), nb, &tv);
    if (rc) return rc;
    std::memcpy(tv.slot->host + tv.runs_off, runs.data(), runs.size() * sizeof(Run));
    std::memcpy(tv.slot->host + tv.addrs_off, addrs.data(), nb * 8);
    std::memcpy(tv.slot->host + tv.offs_off, offs.data(), (nb + 1) * 4);
    rc = table_upload(ctx, tv, ctx->s_gather);
    if (rc) return rc;

    rc = timing_begin(ctx, 0, ctx->s_gather);
    if (rc) return rc;
    if (ctx->cfg.format == B200KV_FMT_FP8) {
      rc = launch_fp8_store(ctx, tv.slot->dev, tv, static_cast<uint32_t>(nb), batch_ntok.data(), ctx->s_gather);
    } else if (ctx->cfg.format == B200KV_FMT_Q4) {
      rc = launch_q4(ctx, true, tv.slot->dev, tv, 0, runs.size(), ctx->s_gather);
    } else {
      rc = launch_copy_runs<kStore>(ctx, tv.slot->dev, tv, 0, n_full, n_part, ctx->s_gather);
    }
    if (rc) return rc;
    rc = timing_end(ctx, 0, ctx->s_gather);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(tv.slot->done_ev, ctx->s_gather));
    // the pages may be reused as soon as the gather is done: compute waits for the kernel only
    CU_TRY(cudaStreamWaitEvent(cs, tv.slot->done_ev, 0));
    CU_TRY(cudaStreamWaitEvent(ctx->s_d2h, tv.slot->done_ev, 0));
    for (size_t i = 0; i < nb; ++i) {
      const Todo& t = todo[b0 + i];
      rc = copy_chunk(ctx, b200kv_pool_slot_ptr(ctx->pool, t.slot), reinterpret_cast<void*>(addrs[i]),
                      t.n_tok, cudaMemcpyDeviceToHost, ctx->s_d2h);
      if (rc) return rc;
      CU_TRY(cudaEventRecord(ctx->stage[sidx[i]].free_ev, ctx->s_d2h));
      ctx->stage[sidx[i]].used = true;
      op->commit_keys.push_back(keys[t.c]);
      ctx->stats.n_stored_tokens += t.n_tok;
    }
  }
  // chunks whose D2H was never queued (none on this path) would stay reserved: every todo chunk is in commit_keys
  CU_TRY(cudaLaunchHostFunc(ctx->s_d2h, op_host_cb, op));
  guard.cb_enqueued = true;
  CU_TRY(cudaEventRecord(op->done, ctx->s_d2h));
  *ticket = op->id;
  ctx->ops.emplace(op->id, std::move(guard.op));
  guard.armed = false;
  return B200KV_OK;
}

extern "C" int b200kv_store_async(b200kv_ctx* ctx, const uint64_t* keys, int32_t n_chunks,
                                  const int64_t* slot_mapping, int64_t n_tokens,
                                  void* compute_stream, uint64_t* ticket) {
  if (!ctx || n_tokens <= 0) return B200KV_EINVAL;
  const int64_t C = ctx->g.C;
  if (n_chunks != static_cast<int32_t>((n_tokens + C - 1) / C)) return B200KV_EINVAL;
  std::vector<int32_t> ct(static_cast<size_t>(n_chunks));
  for (int32_t c = 0; c < n_chunks; ++c) ct[c] = static_cast<int32_t>(std::min<int64_t>(C, n_tokens - c * C));
  return store_impl(ctx, keys, ct.data(), n_chunks, slot_mapping, compute_stream, ticket);
}

extern "C" int b200kv_store_batch_async(b200kv_ctx* ctx, const uint64_t* keys, const int32_t* chunk_tokens,
                                        int32_t n_chunks, const int64_t* slot_mapping, void* compute_stream,
                                        uint64_t* ticket) {
  return store_impl(ctx, keys, chunk_tokens, n_chunks, slot_mapping, compute_stream, ticket);
}

// ------------------------------------------------------------------------------------------------
// load: pinned pool -> staging (DMA) -> paged HBM (kernel), pipelined per chunk
// ------------------------------------------------------------------------------------------------
// Chunk layout as in store_impl.  Request r owns chunks [req_first[r], req_first[r+1]); each request is loaded
// up to its first missing chunk (prefix semantics of lmcache_engine.retrieve), independently of the others.
static int load_impl(b200kv_ctx* ctx, const uint64_t* keys, const int32_t* chunk_tokens, int32_t n_chunks,
                     const int32_t* req_first, int32_t n_reqs, const int64_t* slot_mapping, void* compute_stream,
                     uint64_t* ticket, int64_t* req_loaded, int32_t layers_per_group) {
  if (!ctx || !keys || !chunk_tokens || !req_first || !slot_mapping || n_chunks <= 0 || n_reqs <= 0 || !ticket ||
      layers_per_group < 0)
    return B200KV_EINVAL;
  if (!ctx->pool || !ctx->kv_registered || ctx->stage.empty()) return B200KV_EINVAL;
  const Geometry& g = ctx->g;
  if (req_first[0] != 0 || req_first[n_reqs] != n_chunks) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  reap(ctx);
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream);
  *ticket = 0;
  if (req_loaded) std::fill(req_loaded, req_loaded + n_reqs, 0);
  {
    const int64_t max_slot = static_cast<int64_t>(ctx->cfg.n_blocks) * g.bs;
    for (int32_t c = 0; c < n_chunks; ++c) {   // before pinning anything
      if (chunk_tokens[c] <= 0 || chunk_tokens[c] > static_cast<int32_t>(g.C)) return B200KV_EINVAL;
      const int64_t* sm = slot_mapping + static_cast<int64_t>(c) * g.C;
      for (int32_t i = 0; i < chunk_tokens[c]; ++i)
        if (sm[i] < 0 || sm[i] >= max_slot) return B200KV_EINVAL;
    }
    for (int32_t r = 0; r < n_reqs; ++r)
      if (req_first[r] > req_first[r + 1]) return B200KV_EINVAL;
  }

  struct Todo { int32_t c; uint32_t slot; uint32_t n_tok; };
  std::vector<Todo> todo;
  for (int32_t r = 0; r < n_reqs; ++r) {
    for (int32_t c = req_first[r]; c < req_first[r + 1]; ++c) {
      const uint32_t want = static_cast<uint32_t>(chunk_tokens[c]);
      uint32_t slot = 0, fmt = 0;
      int32_t have = 0;
      if (b200kv_pool_acquire(ctx->pool, keys[c], &slot, &have, &fmt) != B200KV_OK) break;
      if (static_cast<uint32_t>(have) != want || fmt != pool_fmt(ctx)) {
        b200kv_pool_release(ctx->pool, keys[c]);
        break;
      }
      todo.push_back({c, slot, want});
      if (req_loaded) req_loaded[r] += want;
    }
  }
  ++ctx->stats.n_load_ops;
  if (todo.empty()) return B200KV_OK;

  OpGuard guard(ctx, ctx->s_h2d, ctx->s_scatter, nullptr, nullptr);
  guard.op.reset(new Op());
  Op* op = guard.op.get();
  op->id = ctx->next_ticket++;
  op->pool = ctx->pool;
  for (const Todo& t : todo) op->release_keys.push_back(keys[t.c]);
  // From here on every early return drops the pins and the op's events through `guard` (not every failure
  // below is fatal: table_acquire refuses an op whose run table does not fit).
  CU_TRY(cudaEventCreateWithFlags(&op->done, cudaEventDisableTiming));

  const bool detached = compute_stream == B200KV_STREAM_DETACHED;
  if (!detached) {
    cudaEvent_t ev_compute;
    CU_TRY(cudaEventCreateWithFlags(&ev_compute, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ev_compute, cs));
    CU_TRY(cudaStreamWaitEvent(ctx->s_scatter, ev_compute, 0));
    CU_TRY(cudaEventDestroy(ev_compute));
  }

  timing_reset(ctx, 1);
  const size_t n_stage = ctx->stage.size() - ctx->n_store_slots;
  int64_t loaded = 0;
  // Layer-wise: every chunk must own a staging slot for the whole op (the groups of all chunks are
  // interleaved); an op larger than the load half of the ring falls back to the chunk-wise path.
  const bool layerwise = layers_per_group > 0 && todo.size() <= n_stage;
  if (layerwise) {
    const uint32_t G = static_cast<uint32_t>(layers_per_group);
    const uint32_t n_groups = (g.L + G - 1) / G;
    const size_t nb = todo.size();
    std::vector<Run> runs;
    std::vector<uint32_t> offs(nb + 1), n_full_of(nb), sidx(nb);
    std::vector<uint64_t> addrs(nb);
    for (size_t i = 0; i < nb; ++i) {
      const Todo& t = todo[i];
      offs[i] = static_cast<uint32_t>(runs.size());
      const int64_t tb = static_cast<int64_t>(t.c) * g.C;
      std::vector<Run> cf, cp;
      int rc = build_runs(ctx, slot_mapping, tb, tb + t.n_tok, tb - static_cast<int64_t>(i) * g.C, &cf,
                          ctx->cfg.format != B200KV_FMT_RAW ? nullptr : &cp);
      if (rc) return rc;
      n_full_of[i] = static_cast<uint32_t>(cf.size());
      runs.insert(runs.end(), cf.begin(), cf.end());
      runs.insert(runs.end(), cp.begin(), cp.end());
      sidx[i] = ctx->n_store_slots + ctx->load_next;
      ctx->load_next = (ctx->load_next + 1) % static_cast<uint32_t>(n_stage);
      addrs[i] = reinterpret_cast<uint64_t>(ctx->d_staging) + static_cast<uint64_t>(sidx[i]) * g.chunk_bytes;
      StageSlot& ss = ctx->stage[sidx[i]];
      if (ss.used) CU_TRY(cudaStreamWaitEvent(ctx->s_h2d, ss.free_ev, 0));
      loaded += t.n_tok;
    }
    offs[nb] = static_cast<uint32_t>(runs.size());
    TableView tv;
    int rc = table_acquire(ctx, runs.size(), nb, &tv);
    if (rc) return rc;
    std::memcpy(tv.slot->host + tv.runs_off, runs.data(), runs.size() * sizeof(Run));
    std::memcpy(tv.slot->host + tv.addrs_off, addrs.data(), nb * 8);
    std::memcpy(tv.slot->host + tv.offs_off, offs.data(), (nb + 1) * 4);
    rc = table_upload(ctx, tv, ctx->s_scatter);
    if (rc) return rc;
    op->layers_per_group = layers_per_group;
    op->group_ev.resize(n_groups, nullptr);
    cudaEvent_t landed;
    CU_TRY(cudaEventCreateWithFlags(&landed, cudaEventDisableTiming));
    for (uint32_t gi = 0; gi < n_groups; ++gi) {
      const uint32_t pb = 2 * gi * G;
      const uint32_t np = std::min<uint32_t>(2 * G, g.planes - pb);
      for (size_t i = 0; i < nb; ++i) {
        rc = copy_chunk_planes(ctx, reinterpret_cast<void*>(addrs[i]), b200kv_pool_slot_ptr(ctx->pool, todo[i].slot),
                               todo[i].n_tok, pb, np, gi == 0, cudaMemcpyHostToDevice, ctx->s_h2d);
        if (rc) return rc;
      }
      CU_TRY(cudaEventRecord(landed, ctx->s_h2d));           // this group of every chunk has landed
      CU_TRY(cudaStreamWaitEvent(ctx->s_scatter, landed, 0));
      rc = timing_begin(ctx, 1, ctx->s_scatter);
      if (rc) return rc;
      if (ctx->cfg.format == B200KV_FMT_FP8) {
        rc = launch_fp8_load(ctx, tv.slot->dev, tv, 0, runs.size(), ctx->s_scatter, pb, np);
      } else if (ctx->cfg.format == B200KV_FMT_Q4) {
        rc = launch_q4(ctx, false, tv.slot->dev, tv, 0, runs.size(), ctx->s_scatter, pb, np);
      } else {
        // one launch per chunk keeps the (full, partial) run split of each chunk
        for (size_t i = 0; i < nb && rc == B200KV_OK; ++i)
          rc = launch_copy_runs<kLoad>(ctx, tv.slot->dev, tv, offs[i], n_full_of[i],
                                       offs[i + 1] - offs[i] - n_full_of[i], ctx->s_scatter, nullptr, pb, np);
      }
      if (rc) return rc;
      rc = timing_end(ctx, 1, ctx->s_scatter);
      if (rc) return rc;
      CU_TRY(cudaEventCreateWithFlags(&op->group_ev[gi], cudaEventDisableTiming));
      CU_TRY(cudaEventRecord(op->group_ev[gi], ctx->s_scatter));
    }
    CU_TRY(cudaEventDestroy(landed));
    for (size_t i = 0; i < nb; ++i) {
      CU_TRY(cudaEventRecord(ctx->stage[sidx[i]].free_ev, ctx->s_scatter));
      ctx->stage[sidx[i]].used = true;
    }
    CU_TRY(cudaEventRecord(tv.slot->done_ev, ctx->s_scatter));
  }
  for (size_t b0 = 0; !layerwise && b0 < todo.size(); b0 += n_stage) {
    const size_t nb = std::min(n_stage, todo.size() - b0);
    std::vector<Run> runs;
    std::vector<uint32_t> offs(nb + 1);
    std::vector<uint32_t> n_full_of(nb);
    std::vector<uint64_t> addrs(nb);
    std::vector<uint32_t> sidx(nb);
    for (size_t i = 0; i < nb; ++i) {
      const Todo& t = todo[b0 + i];
      offs[i] = static_cast<uint32_t>(runs.size());
      const int64_t tb = static_cast<int64_t>(t.c) * g.C;
      std::vector<Run> cf, cp;  // per chunk: whole tiles first, then HND partial-tile runs
      int rc = build_runs(ctx, slot_mapping, tb, tb + t.n_tok, tb - static_cast<int64_t>(i) * g.C, &cf,
                          ctx->cfg.format != B200KV_FMT_RAW ? nullptr : &cp);
      if (rc) return rc;
      n_full_of[i] = static_cast<uint32_t>(cf.size());
      runs.insert(runs.end(), cf.begin(), cf.end());
      runs.insert(runs.end(), cp.begin(), cp.end());
      sidx[i] = ctx->n_store_slots + ctx->load_next;
      ctx->load_next = (ctx->load_next + 1) % static_cast<uint32_t>(n_stage);
      addrs[i] = reinterpret_cast<uint64_t>(ctx->d_staging) + static_cast<uint64_t>(sidx[i]) * g.chunk_bytes;
    }
    offs[nb] = static_cast<uint32_t>(runs.size());
    TableView tv;
    int rc = table_acquire(ctx, runs.size(), nb, &tv);
    if (rc) return rc;
    std::memcpy(tv.slot->host + tv.runs_off, runs.data(), runs.size() * sizeof(Run));
    std::memcpy(tv.slot->host + tv.addrs_off, addrs.data(), nb * 8);
    std::memcpy(tv.slot->host + tv.offs_off, offs.data(), (nb + 1) * 4);
    rc = table_upload(ctx, tv, ctx->s_scatter);
    if (rc) return rc;

    for (size_t i = 0; i < nb; ++i) {
      const Todo& t = todo[b0 + i];
      StageSlot& ss = ctx->stage[sidx[i]];
      if (ss.used) CU_TRY(cudaStreamWaitEvent(ctx->s_h2d, ss.free_ev, 0));
      rc = copy_chunk(ctx, reinterpret_cast<void*>(addrs[i]), b200kv_pool_slot_ptr(ctx->pool, t.slot),
                      t.n_tok, cudaMemcpyHostToDevice, ctx->s_h2d);
      if (rc) return rc;
      CU_TRY(cudaEventRecord(ss.free_ev, ctx->s_h2d));  // reused as "chunk landed" first ...
      CU_TRY(cudaStreamWaitEvent(ctx->s_scatter, ss.free_ev, 0));
      rc = timing_begin(ctx, 1, ctx->s_scatter);
      if (rc) return rc;
      const size_t r0 = offs[i], rn = offs[i + 1] - offs[i];
      if (ctx->cfg.format == B200KV_FMT_FP8) {
        rc = launch_fp8_load(ctx, tv.slot->dev, tv, r0, rn, ctx->s_scatter);
      } else if (ctx->cfg.format == B200KV_FMT_Q4) {
        rc = launch_q4(ctx, false, tv.slot->dev, tv, r0, rn, ctx->s_scatter);
      } else {
        rc = launch_copy_runs<kLoad>(ctx, tv.slot->dev, tv, r0, n_full_of[i], rn - n_full_of[i], ctx->s_scatter);
      }
      if (rc) return rc;
      rc = timing_end(ctx, 1, ctx->s_scatter);
      if (rc) return rc;
      CU_TRY(cudaEventRecord(ss.free_ev, ctx->s_scatter));  // ... then as "slot drained"
      ss.used = true;
      loaded += t.n_tok;
    }
    CU_TRY(cudaEventRecord(tv.slot->done_ev, ctx->s_scatter));
  }
  CU_TRY(cudaLaunchHostFunc(ctx->s_scatter, op_host_cb, op));
  guard.cb_enqueued = true;
  CU_TRY(cudaEventRecord(op->done, ctx->s_scatter));
  // chunk-wise: the forward pass must see the loaded pages; layer-wise: it waits per layer instead
  if (!detached && !layerwise) CU_TRY(cudaStreamWaitEvent(cs, op->done, 0));
  ctx->stats.n_loaded_tokens += loaded;
  *ticket = op->id;
  ctx->ops.emplace(op->id, std::move(guard.op));
  guard.armed = false;
  return B200KV_OK;
}

static int load_single(b200kv_ctx* ctx, const uint64_t* keys, int32_t n_chunks, const int64_t* slot_mapping,
                       int64_t n_tokens, int32_t skip_chunks, void* compute_stream, uint64_t* ticket,
                       int64_t* n_loaded_tokens, int32_t layers_per_group) {
  if (!ctx || !keys || !slot_mapping || n_tokens <= 0 || !ticket || skip_chunks < 0) return B200KV_EINVAL;
  const int64_t C = ctx->g.C;
  if (n_chunks != static_cast<int32_t>((n_tokens + C - 1) / C)) return B200KV_EINVAL;
  if (ticket) *ticket = 0;
  if (n_loaded_tokens) *n_loaded_tokens = 0;
  if (skip_chunks >= n_chunks) {      // nothing after the masked prefix: still validate like the general path
    const int64_t max_slot = static_cast<int64_t>(ctx->cfg.n_blocks) * ctx->g.bs;
    for (int64_t i = 0; i < n_tokens; ++i)
      if (slot_mapping[i] < 0 || slot_mapping[i] >= max_slot) return B200KV_EINVAL;
    return (ctx->pool && ctx->kv_registered && !ctx->stage.empty()) ? B200KV_OK : B200KV_EINVAL;
  }
  {   // slots of the masked prefix are validated too (the old single-request contract)
    const int64_t max_slot = static_cast<int64_t>(ctx->cfg.n_blocks) * ctx->g.bs;
    for (int64_t i = 0; i < static_cast<int64_t>(skip_chunks) * C && i < n_tokens; ++i)
      if (slot_mapping[i] < 0 || slot_mapping[i] >= max_slot) return B200KV_EINVAL;
  }
  const int32_t n = n_chunks - skip_chunks;
  std::vector<int32_t> ct(static_cast<size_t>(n));
  for (int32_t c = 0; c < n; ++c) ct[c] = static_cast<int32_t>(std::min<int64_t>(C, n_tokens - (c + skip_chunks) * C));
  const int32_t first[2] = {0, n};
  int64_t loaded = 0;
  const int rc = load_impl(ctx, keys + skip_chunks, ct.data(), n, first, 1, slot_mapping + static_cast<int64_t>(skip_chunks) * C,
                           compute_stream, ticket, &loaded, layers_per_group);
  if (rc == B200KV_OK && n_loaded_tokens) *n_loaded_tokens = loaded;
  return rc;
}

extern "C" int b200kv_load_async(b200kv_ctx* ctx, const uint64_t* keys, int32_t n_chunks,
                                 const int64_t* slot_mapping, int64_t n_tokens, int32_t skip_chunks,
                                 void* compute_stream, uint64_t* ticket, int64_t* n_loaded_tokens) {
  return load_single(ctx, keys, n_chunks, slot_mapping, n_tokens, skip_chunks, compute_stream, ticket,
                     n_loaded_tokens, 0);
}

extern "C" int b200kv_load_layerwise_async(b200kv_ctx* ctx, const uint64_t* keys, int32_t n_chunks,
                                           const int64_t* slot_mapping, int64_t n_tokens, int32_t skip_chunks,
                                           int32_t layers_per_group, void* compute_stream, uint64_t* ticket,
                                           int64_t* n_loaded_tokens) {
  if (layers_per_group <= 0) return B200KV_EINVAL;
  return load_single(ctx, keys, n_chunks, slot_mapping, n_tokens, skip_chunks, compute_stream, ticket,
                     n_loaded_tokens, layers_per_group);
}

extern "C" int b200kv_load_batch_async(b200kv_ctx* ctx, const uint64_t* keys, const int32_t* chunk_tokens,
                                       int32_t n_chunks, const int32_t* req_first_chunk, int32_t n_reqs,
                                       const int64_t* slot_mapping, int32_t layers_per_group, void* compute_stream,
                                       uint64_t* ticket, int64_t* req_loaded_tokens) {
  return load_impl(ctx, keys, chunk_tokens, n_chunks, req_first_chunk, n_reqs, slot_mapping, compute_stream, ticket,
                   req_loaded_tokens, layers_per_group);
}

extern "C" int b200kv_wait_layer(b200kv_ctx* ctx, uint64_t ticket, int32_t layer, void* compute_stream) {
  if (!ctx || layer < 0) return B200KV_EINVAL;
  if (ticket == 0) return B200KV_OK;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->ops.find(ticket);
  if (it == ctx->ops.end()) return B200KV_OK;  // finished and reaped: the pages are already there
  Op* op = it->second.get();
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream);
  if (op->group_ev.empty()) {  // the op fell back to the chunk-wise path: wait for all of it
    CU_TRY(cudaStreamWaitEvent(cs, op->done, 0));
    return B200KV_OK;
  }
  const size_t gi = std::min<size_t>(static_cast<size_t>(layer / op->layers_per_group), op->group_ev.size() - 1);
  CU_TRY(cudaStreamWaitEvent(cs, op->group_ev[gi], 0));
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// tickets
// ------------------------------------------------------------------------------------------------
extern "C" int b200kv_poll(b200kv_ctx* ctx, uint64_t ticket, int* done) {
  if (!ctx || !done) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ticket == 0) { *done = 1; return B200KV_OK; }
  auto it = ctx->ops.find(ticket);
  if (it == ctx->ops.end()) { *done = 1; return B200KV_OK; }  // already reaped
  Op* op = it->second.get();
  const cudaError_t q = cudaEventQuery(op->done);
  if (q == cudaSuccess && op->host_done.load(std::memory_order_acquire)) {
    destroy_op_events(op);
    ctx->ops.erase(it);
    *done = 1;
  } else if (q == cudaSuccess || q == cudaErrorNotReady) {
    *done = 0;
  } else {
    set_err("cudaEventQuery", q, __LINE__);
    return B200KV_ENODEV;
  }
  return B200KV_OK;
}

extern "C" int b200kv_wait(b200kv_ctx* ctx, uint64_t ticket) {
  if (!ctx) return B200KV_EINVAL;
  if (ticket == 0) return B200KV_OK;
  DeviceGuard dg(ctx->cfg.device);
  cudaEvent_t ev = nullptr;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->ops.find(ticket);
    if (it == ctx->ops.end()) return B200KV_OK;
    ev = it->second->done;
  }
  CU_TRY(cudaEventSynchronize(ev));
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->ops.find(ticket);
  if (it != ctx->ops.end()) {
    while (!it->second->host_done.load(std::memory_order_acquire)) { /* callback precedes event */ }
    destroy_op_events(it->second.get());
    ctx->ops.erase(it);
  }
  return B200KV_OK;
}

extern "C" int b200kv_wait_all(b200kv_ctx* ctx) {
  if (!ctx) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  for (cudaStream_t s : {ctx->s_gather, ctx->s_d2h, ctx->s_h2d, ctx->s_scatter})
    if (s) CU_TRY(cudaStreamSynchronize(s));
  std::lock_guard<std::mutex> lk(ctx->mu);
  reap(ctx);
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// peers
// ------------------------------------------------------------------------------------------------
extern "C" int b200kv_export_ipc(b200kv_ctx* ctx, b200kv_ipc_desc* descs_out, int32_t n_descs) {
  if (!ctx || !descs_out) return B200KV_EINVAL;
  if (!ctx->kv_registered || n_descs != static_cast<int32_t>(ctx->g.planes)) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  typedef CUresult (*GetRange)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CU_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  if (!fn || qr != cudaDriverEntryPointSuccess) return B200KV_ENODEV;
  GetRange get_range = reinterpret_cast<GetRange>(fn);
  for (uint32_t l = 0; l < ctx->g.L; ++l) {
    for (int kv = 0; kv < 2; ++kv) {
      const void* p = kv ? ctx->h_v[l] : ctx->h_k[l];
      CUdeviceptr base = 0;
      size_t size = 0;
      if (get_range(&base, &size, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) return B200KV_ENODEV;
      b200kv_ipc_desc& d = descs_out[kv * ctx->g.L + l];  // K planes then V planes
      std::memset(&d, 0, sizeof(d));
      cudaIpcMemHandle_t h;
      CU_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
      static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
      std::memcpy(d.handle, &h, 64);
      d.offset = reinterpret_cast<uint64_t>(p) - static_cast<uint64_t>(base);
      d.alloc_bytes = size;
    }
  }
  return B200KV_OK;
}

static int install_peer(b200kv_ctx* ctx, int32_t peer_id, int32_t peer_device,
                        const std::vector<uint64_t>& bases, uint64_t stride, uint64_t n_blocks,
                        std::vector<void*>&& opened) {
  Peer& p = ctx->peers[peer_id];
  if (p.valid) return B200KV_EEXIST;
  CU_TRY(cudaMalloc(&p.d_bases, sizeof(uint64_t) * bases.size()));
  CU_TRY(cudaMemcpy(p.d_bases, bases.data(), sizeof(uint64_t) * bases.size(), cudaMemcpyHostToDevice));
  p.device = peer_device;
  p.block_stride = stride;
  p.n_blocks = n_blocks;
  p.opened = std::move(opened);
  p.valid = true;
  return B200KV_OK;
}

extern "C" int b200kv_import_peer(b200kv_ctx* ctx, int32_t peer_id, int32_t peer_device,
                                  const b200kv_ipc_desc* descs, int32_t n_descs,
                                  uint64_t peer_block_stride_bytes, uint64_t peer_n_blocks) {
  if (!ctx || !descs || peer_id < 0 || peer_id >= kMaxPeers) return B200KV_EINVAL;
  if (n_descs != static_cast<int32_t>(ctx->g.planes) || peer_n_blocks == 0 ||
      peer_n_blocks * ctx->g.bs > 0x7fffffffull)
    return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  // one cudaIpcOpenMemHandle per distinct allocation
  std::vector<std::pair<std::string, void*>> seen;
  std::vector<void*> opened;
  std::vector<uint64_t> bases(ctx->g.planes);
  for (uint32_t l = 0; l < ctx->g.L; ++l) {
    for (int kv = 0; kv < 2; ++kv) {
      const b200kv_ipc_desc& d = descs[kv * ctx->g.L + l];
      const std::string key(reinterpret_cast<const char*>(d.handle), 64);
      void* base = nullptr;
      for (auto& s : seen)
        if (s.first == key) base = s.second;
      if (!base) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, d.handle, 64);
        CU_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        seen.emplace_back(key, base);
        opened.push_back(base);
      }
      bases[2 * l + kv] = reinterpret_cast<uint64_t>(base) + d.offset;
    }
  }
  return install_peer(ctx, peer_id, peer_device, bases, peer_block_stride_bytes, peer_n_blocks,
                      std::move(opened));
}

extern "C" int b200kv_import_peer_ptrs(b200kv_ctx* ctx, int32_t peer_id, int32_t peer_device,
                                       const void* const* k_ptrs, const void* const* v_ptrs,
                                       uint64_t peer_block_stride_bytes, uint64_t peer_n_blocks) {
  if (!ctx || !k_ptrs || !v_ptrs || peer_id < 0 || peer_id >= kMaxPeers) return B200KV_EINVAL;
  if (peer_n_blocks == 0 || peer_n_blocks * ctx->g.bs > 0x7fffffffull) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (peer_device != ctx->cfg.device) {
    int can = 0;
    CU_TRY(cudaDeviceCanAccessPeer(&can, ctx->cfg.device, peer_device));
    if (!can) return B200KV_ENOTSUP;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      set_err("cudaDeviceEnablePeerAccess", e, __LINE__);
      return B200KV_ENODEV;
    }
    cudaGetLastError();
  }
  std::vector<uint64_t> bases(ctx->g.planes);
  for (uint32_t l = 0; l < ctx->g.L; ++l) {
    bases[2 * l] = reinterpret_cast<uint64_t>(k_ptrs[l]);
    bases[2 * l + 1] = reinterpret_cast<uint64_t>(v_ptrs[l]);
    if (!bases[2 * l] || !bases[2 * l + 1] || bases[2 * l] % 16 || bases[2 * l + 1] % 16)
      return B200KV_EINVAL;
  }
  return install_peer(ctx, peer_id, peer_device, bases, peer_block_stride_bytes, peer_n_blocks, {});
}

extern "C" int b200kv_peer_pull_async(b200kv_ctx* ctx, int32_t peer_id, const int64_t* src_slots,
                                      const int64_t* dst_slots, int64_t n_tokens,
                                      void* compute_stream, uint64_t* ticket) {
  if (!ctx || !src_slots || !dst_slots || n_tokens <= 0 || !ticket) return B200KV_EINVAL;
  if (peer_id < 0 || peer_id >= kMaxPeers || !ctx->peers[peer_id].valid) return B200KV_ENOENT;
  if (!ctx->kv_registered) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  reap(ctx);
  const Peer& peer = ctx->peers[peer_id];
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream);
  *ticket = 0;

  std::vector<Run> runs, partial;
  int rc = build_pull_runs(ctx, peer, src_slots, dst_slots, n_tokens, &runs, &partial);
  if (rc) return rc;
  const size_t n_full = runs.size(), n_part = partial.size();
  runs.insert(runs.end(), partial.begin(), partial.end());
  TableView tv;
  rc = table_acquire(ctx, runs.size(), 1, &tv);
  if (rc) return rc;
  std::memcpy(tv.slot->host + tv.runs_off, runs.data(), runs.size() * sizeof(Run));

  std::unique_ptr<Op> op(new Op());
  op->id = ctx->next_ticket++;
  CU_TRY(cudaEventCreateWithFlags(&op->done, cudaEventDisableTiming));
  const bool detached = compute_stream == B200KV_STREAM_DETACHED;
  if (!detached) {
    cudaEvent_t ev_compute;
    CU_TRY(cudaEventCreateWithFlags(&ev_compute, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ev_compute, cs));
    CU_TRY(cudaStreamWaitEvent(ctx->s_scatter, ev_compute, 0));
    CU_TRY(cudaEventDestroy(ev_compute));
  }
  rc = table_upload(ctx, tv, ctx->s_scatter);
  if (rc) return rc;

  timing_reset(ctx, 2);
  rc = timing_begin(ctx, 2, ctx->s_scatter);
  if (rc) return rc;
  rc = launch_copy_runs<kPull>(ctx, tv.slot->dev, tv, 0, n_full, n_part, ctx->s_scatter, &peer);
  if (rc) return rc;
  rc = timing_end(ctx, 2, ctx->s_scatter);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(tv.slot->done_ev, ctx->s_scatter));
  CU_TRY(cudaLaunchHostFunc(ctx->s_scatter, op_host_cb, op.get()));
  CU_TRY(cudaEventRecord(op->done, ctx->s_scatter));
  if (!detached) CU_TRY(cudaStreamWaitEvent(cs, op->done, 0));
  ++ctx->stats.n_pull_ops;
  ctx->stats.n_pulled_tokens += n_tokens;
  ctx->stats.p2p_bytes += static_cast<uint64_t>(n_tokens) * ctx->g.token_bytes * ctx->g.planes;
  *ticket = op->id;
  ctx->ops.emplace(op->id, std::move(op));
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------------
extern "C" int b200kv_engine_get_stats(b200kv_ctx* ctx, b200kv_engine_stats* out) {
  if (!ctx || !out) return B200KV_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  *out = ctx->stats;
  return B200KV_OK;
}

extern "C" int b200kv_last_kernel_ms(b200kv_ctx* ctx, int which, float* ms_out) {
  if (!ctx || !ms_out || which < 0 || which > 2) return B200KV_EINVAL;
  DeviceGuard dg(ctx->cfg.device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  float total = 0.f;
  auto& t = ctx->timing[which];
  for (size_t i = 0; i < t.used; ++i) {
    float ms = 0.f;
    CU_TRY(cudaEventSynchronize(t.ev[i].second));
    CU_TRY(cudaEventElapsedTime(&ms, t.ev[i].first, t.ev[i].second));
    total += ms;
  }
  *ms_out = total;
  return B200KV_OK;
}
