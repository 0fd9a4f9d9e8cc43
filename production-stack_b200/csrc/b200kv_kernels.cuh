// b200kv_kernels.cuh — sm_100a kernels of the KV offload / transfer hot path.
//
// What they replace: LMCache's paged-memory GPU connector (`VLLMPagedMemGPUConnectorV2`,
// named at vllm/.../lmcache_integration/vllm_v1_adapter.py:19-23,533-541) — an element-wise
// gather/scatter between vLLM's paged KV cache and a contiguous (L,2,C,H,D) chunk — plus the
// UCX/NIXL peer copy the reference configures for disaggregated prefill
// (helm/templates/deployment-vllm-multi.yaml:296-324).
//
// All of them are HBM-bound byte movers: no tensor cores.  The design rules
// that matter are coalescing, enough bytes in flight per SM to cover DRAM latency, and a grid
// that is a multiple of the 148 SMs.
//
//   kv_bulk_copy_kernel  RAW store / load / peer pull.  One elected thread per CTA drives the
//                        TMA engine: cp.async.bulk global->smem (mbarrier complete_tx) into an
//                        S-stage ring, then cp.async.bulk smem->global (bulk_group).  No
//                        registers touch the payload.  Persistent grid.
//   kv_ldg_copy_kernel   same contract with 128-bit ld.global.nc / st.global (A/B variant).
//   kv_fp8_store_kernel  cluster of 8 CTAs per (chunk, layer, K/V) slab: warp 0 bulk-loads the CTA's 32
//                        tokens into smem (one run per lane), per-head absmax in registers -> smem,
//                        every CTA pushes its maxima into all 8 CTAs' smem with DSMEM atomics
//                        (red.shared::cluster.max) between a split cluster barrier (arrive early,
//                        wait late) and one full one, quantise to e4m3 from smem, coalesced stores.
//                        HBM is read exactly once.  NHD and HND tiles.
//   kv_fp8_load_kernel   bulk-load e4m3 piece -> dequantise -> 16-byte stores into the pages.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200kv {

// One run = consecutive tokens that stay inside one vLLM block and one chunk.
//   store / load : a = paged slot (block*bs + off) ; b = op-relative token index
//   peer pull    : a = peer slot                   ; b = local slot
struct Run {
  int32_t a;
  int32_t b;
  int32_t n;
};

struct PagedSide {
  const uint64_t* bases;  // device array [2*L]: K plane of layer l at [2l], V plane at [2l+1]
  uint64_t block_stride;  // bytes between consecutive blocks of one plane
  uint32_t block_tokens;
  uint32_t token_bytes;   // H*D*elem (NHD: one token's heads are contiguous)
};

struct ChunkSide {
  const uint64_t* chunk_addrs;  // device array [n_chunks]: where chunk c of this op lives
  uint64_t slab_bytes;          // C * token_bytes(format): one (layer, K/V) slab
  uint32_t chunk_tokens;
  uint32_t token_bytes;         // bytes per token per slab in this format (RAW: 2048, FP8: 1024)
};

struct CopyParams {
  PagedSide paged;        // local pages (store src / load dst / pull dst)
  PagedSide peer;         // pull src (peer mapping); unused otherwise
  ChunkSide chunk;
  const Run* runs;
  uint32_t n_runs;
  uint32_t n_planes;      // planes covered by this launch (2*L, or one layer group of it)
  uint32_t plane_begin;   // first plane of the launch (layer-wise loads walk the planes in groups)
  uint32_t pieces;        // pieces per run = ceil(block_tokens / piece_tokens)
  uint32_t piece_tokens;  // tokens per piece (piece bytes <= stage bytes)
  uint32_t total_units;   // n_runs * n_planes * pieces
};

enum CopyMode { kStore = 0, kLoad = 1, kPull = 2 };

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared (this CTA), completion signalled on an mbarrier in bytes.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // <= N groups still reading smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {  // <= N groups not yet complete (writes done)
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_na_v2(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// atomic max into the same shared variable of CTA `rank` of this cluster (DSMEM)
__device__ __forceinline__ void red_dsmem_max_u32(void* local_smem_ptr, uint32_t rank, uint32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  asm volatile("red.relaxed.cluster.shared::cluster.max.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(const void* local_smem_ptr, uint32_t rank) {
  uint32_t remote, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
               : "=r"(remote)
               : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------
// addressing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t paged_addr(const PagedSide& s, uint32_t plane, uint32_t slot) {
  const uint32_t blk = slot / s.block_tokens;
  const uint32_t off = slot - blk * s.block_tokens;
  return __ldg(s.bases + plane) + static_cast<uint64_t>(blk) * s.block_stride +
         static_cast<uint64_t>(off) * s.token_bytes;
}
__device__ __forceinline__ uint64_t chunk_addr(const ChunkSide& s, uint32_t plane, uint32_t tok) {
  const uint32_t c = tok / s.chunk_tokens;
  const uint32_t t = tok - c * s.chunk_tokens;
  return __ldg(s.chunk_addrs + c) + static_cast<uint64_t>(plane) * s.slab_bytes +
         static_cast<uint64_t>(t) * s.token_bytes;
}

struct Unit {
  uint64_t src, dst;
  uint32_t bytes;
};

template <int MODE>
__device__ __forceinline__ Unit decode_unit(const CopyParams& p, uint32_t u) {
  // plane fastest: consecutive units of one CTA stride across planes of the same run.
  const uint32_t plane = p.plane_begin + u % p.n_planes;
  const uint32_t t = u / p.n_planes;
  const uint32_t piece = t % p.pieces;
  const uint32_t r = t / p.pieces;
  const Run run = p.runs[r];
  const int32_t off = static_cast<int32_t>(piece * p.piece_tokens);
  int32_t n = run.n - off;
  n = n < 0 ? 0 : (n > static_cast<int32_t>(p.piece_tokens) ? static_cast<int32_t>(p.piece_tokens) : n);
  Unit out;
  out.bytes = static_cast<uint32_t>(n) * p.paged.token_bytes;
  if (MODE == kStore) {
    out.src = paged_addr(p.paged, plane, static_cast<uint32_t>(run.a + off));
    out.dst = chunk_addr(p.chunk, plane, static_cast<uint32_t>(run.b + off));
  } else if (MODE == kLoad) {
    out.src = chunk_addr(p.chunk, plane, static_cast<uint32_t>(run.b + off));
    out.dst = paged_addr(p.paged, plane, static_cast<uint32_t>(run.a + off));
  } else {
    out.src = paged_addr(p.peer, plane, static_cast<uint32_t>(run.a + off));
    out.dst = paged_addr(p.paged, plane, static_cast<uint32_t>(run.b + off));
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// RAW path, TMA-engine variant.  One warp per CTA, lane 0 does everything:
//   iteration j:  (B) issue bulk load of unit j into stage j%S   (after the store that last used
//                     that stage finished READING smem)
//                 (A) wait for unit j-LAG's bytes, issue its bulk store, commit one group.
// Loads in flight per CTA: LAG; stores in flight: S-LAG.
// ---------------------------------------------------------------------------------------------
template <int MODE, int S, int LAG>
__global__ void __launch_bounds__(32) kv_bulk_copy_kernel(const CopyParams p, uint32_t stage_bytes) {
  static_assert(LAG >= 1 && LAG < S, "need 1 <= LAG < S");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);               // S mbarriers
  uint64_t* dsts = reinterpret_cast<uint64_t*>(smem + 64);          // S pending dst addresses
  uint32_t* lens = reinterpret_cast<uint32_t*>(smem + 64 + 8 * S);  // S pending lengths
  uint8_t* stage0 = smem + 256;

  if (threadIdx.x != 0) return;
#pragma unroll
  for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
  fence_mbar_init();

  const uint32_t first = blockIdx.x;
  const uint32_t step = gridDim.x;
  const uint32_t n = first < p.total_units ? (p.total_units - first + step - 1) / step : 0;

  for (uint32_t j = 0; j < n + LAG; ++j) {
    if (j < n) {
      const uint32_t s = j % S;
      if (j >= S) bulk_wait_read<S - 1 - LAG>();  // store of unit j-S has drained its stage
      const Unit u = decode_unit<MODE>(p, first + j * step);
      dsts[s] = u.dst;
      lens[s] = u.bytes;
      if (u.bytes) bulk_g2s(stage0 + static_cast<size_t>(s) * stage_bytes,
                            reinterpret_cast<const void*>(u.src), u.bytes, &full[s]);
      mbar_arrive_expect_tx(&full[s], u.bytes);
    }
    if (j >= LAG) {
      const uint32_t k = j - LAG;
      const uint32_t s = k % S;
      mbar_wait(&full[s], (k / S) & 1u);
      const uint32_t bytes = lens[s];
      if (bytes) bulk_s2g(reinterpret_cast<void*>(dsts[s]),
                          stage0 + static_cast<size_t>(s) * stage_bytes, bytes);
      bulk_commit();
    }
  }
  bulk_wait_all<0>();  // every store has landed before the grid is considered complete
}

// ---------------------------------------------------------------------------------------------
// RAW path, LSU variant: a CTA copies one unit at a time with 128-bit loads, 4 in flight/thread.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) kv_ldg_copy_kernel(const CopyParams p) {
  for (uint32_t ui = blockIdx.x; ui < p.total_units; ui += gridDim.x) {
    const Unit u = decode_unit<MODE>(p, ui);
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(u.src);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(u.dst);
    const uint32_t nvec = u.bytes >> 4;
    uint32_t i = threadIdx.x;
    for (; i + 3 * 256 < nvec; i += 4 * 256) {
      const uint4 a = ld_nc_v4(src + i), b = ld_nc_v4(src + i + 256);
      const uint4 c = ld_nc_v4(src + i + 512), d = ld_nc_v4(src + i + 768);
      st_na_v4(dst + i, a);
      st_na_v4(dst + i + 256, b);
      st_na_v4(dst + i + 512, c);
      st_na_v4(dst + i + 768, d);
    }
    for (; i < nvec; i += 256) st_na_v4(dst + i, ld_nc_v4(src + i));
  }
}

// ---------------------------------------------------------------------------------------------
// HND tiles ([H][block_tokens][D] inside a block).  A whole-block run is one contiguous tile and
// goes through the kernels above unchanged (the chunk stores tiles verbatim).  Runs that cover
// only part of a tile — the ragged tail of a request, or token-granular mappings — are moved
// per head by this kernel: one unit = (run, plane, head) = n * D * elem contiguous bytes.
// ---------------------------------------------------------------------------------------------
struct HndParams {
  PagedSide paged, peer;
  ChunkSide chunk;       // chunk.token_bytes = H*D*elem; a tile = block_tokens * token_bytes
  const Run* runs;
  uint32_t n_runs, n_planes, n_heads;
  uint32_t plane_begin;
  uint32_t row_bytes;    // D * elem: one (token, head) row
  uint32_t total_units;  // n_runs * n_planes * n_heads
};

__device__ __forceinline__ uint64_t paged_addr_hnd(const PagedSide& s, uint32_t plane, uint32_t slot,
                                                   uint32_t h, uint32_t row_bytes) {
  const uint32_t blk = slot / s.block_tokens;
  const uint32_t off = slot - blk * s.block_tokens;
  return __ldg(s.bases + plane) + static_cast<uint64_t>(blk) * s.block_stride +
         static_cast<uint64_t>(h) * s.block_tokens * row_bytes + static_cast<uint64_t>(off) * row_bytes;
}
__device__ __forceinline__ uint64_t chunk_addr_hnd(const ChunkSide& s, uint32_t plane, uint32_t tok,
                                                   uint32_t h, uint32_t block_tokens, uint32_t row_bytes) {
  const uint32_t c = tok / s.chunk_tokens;
  const uint32_t t = tok - c * s.chunk_tokens;
  const uint32_t tile = t / block_tokens;
  const uint32_t off = t - tile * block_tokens;
  return __ldg(s.chunk_addrs + c) + static_cast<uint64_t>(plane) * s.slab_bytes +
         static_cast<uint64_t>(tile) * block_tokens * s.token_bytes +
         static_cast<uint64_t>(h) * block_tokens * row_bytes + static_cast<uint64_t>(off) * row_bytes;
}

template <int MODE>
__global__ void __launch_bounds__(128) kv_hnd_partial_kernel(const HndParams p) {
  for (uint32_t ui = blockIdx.x; ui < p.total_units; ui += gridDim.x) {
    const uint32_t h = ui % p.n_heads;
    const uint32_t t = ui / p.n_heads;
    const uint32_t plane = p.plane_begin + t % p.n_planes;
    const Run run = p.runs[t / p.n_planes];
    uint64_t src, dst;
    if (MODE == kStore) {
      src = paged_addr_hnd(p.paged, plane, static_cast<uint32_t>(run.a), h, p.row_bytes);
      dst = chunk_addr_hnd(p.chunk, plane, static_cast<uint32_t>(run.b), h, p.paged.block_tokens, p.row_bytes);
    } else if (MODE == kLoad) {
      src = chunk_addr_hnd(p.chunk, plane, static_cast<uint32_t>(run.b), h, p.paged.block_tokens, p.row_bytes);
      dst = paged_addr_hnd(p.paged, plane, static_cast<uint32_t>(run.a), h, p.row_bytes);
    } else {
      src = paged_addr_hnd(p.peer, plane, static_cast<uint32_t>(run.a), h, p.row_bytes);
      dst = paged_addr_hnd(p.paged, plane, static_cast<uint32_t>(run.b), h, p.row_bytes);
    }
    const uint32_t nvec = (static_cast<uint32_t>(run.n) * p.row_bytes) >> 4;
    for (uint32_t i = threadIdx.x; i < nvec; i += 128)
      st_na_v4(reinterpret_cast<uint4*>(dst) + i, ld_nc_v4(reinterpret_cast<const uint4*>(src) + i));
  }
}

// ---------------------------------------------------------------------------------------------
// FP8 store.  grid = (kCluster * n_slabs), cluster (kCluster,1,1); slab = (chunk c, plane).
// CTA `rank` owns tokens [rank*W, rank*W+W) of the chunk, W = C / kCluster.
// ---------------------------------------------------------------------------------------------
constexpr int kCluster = 8;
constexpr int kFp8Threads = 256;
constexpr int kMaxHeads = 64;

struct Fp8StoreParams {
  PagedSide paged;
  const Run* runs;                // sorted by b
  const uint32_t* chunk_run_off;  // [n_chunks+1] run range of each chunk
  const uint64_t* chunk_addrs;    // [n_chunks] destination of each chunk
  uint32_t n_chunks, n_planes;
  uint32_t chunk_tokens;          // C
  uint32_t n_tokens;              // op length
  uint32_t n_heads, head_bytes;   // H, D*2
  uint64_t slab_q_bytes;          // C * H * D   (e4m3 bytes per slab)
  uint64_t scales_off;            // byte offset of the (planes, H) fp32 scales inside a chunk
  uint32_t hnd;                   // tiles are [H][block_tokens][D]; the packed chunk mirrors them
};

__device__ __forceinline__ uint32_t absmax_u16x2(uint32_t acc, uint32_t w) {
  return __vmaxu2(acc, w & 0x7fff7fffu);
}

// e4m3 of 8 bf16 scaled by inv (fp32 multiply, RN, satfinite) -> 8 bytes.
__device__ __forceinline__ uint2 quant8(const uint4& v, float inv) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float a0 = __uint_as_float(w[2 * i] << 16) * inv;
    const float a1 = __uint_as_float(w[2 * i] & 0xffff0000u) * inv;
    const float b0 = __uint_as_float(w[2 * i + 1] << 16) * inv;
    const float b1 = __uint_as_float(w[2 * i + 1] & 0xffff0000u) * inv;
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a0, a1), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(b0, b1), __NV_SATFINITE, __NV_E4M3);
    o[i] = lo | (hi << 16);
  }
  return make_uint2(o[0], o[1]);
}

template <int kThreads>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
    kv_fp8_store_kernel(const Fp8StoreParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_absmax[kMaxHeads];  // this CTA's |x| max per head (bf16 bits: integer order == magnitude)
  __shared__ uint32_t s_all[kMaxHeads];     // cluster-wide max: every CTA pushes its value into every CTA's copy
  __shared__ float s_inv[kMaxHeads];

  const uint32_t rank = cluster_ctarank();
  const uint32_t slab = blockIdx.x / kCluster;
  const uint32_t c = slab / p.n_planes;
  const uint32_t plane = slab - c * p.n_planes;
  const uint32_t W = p.chunk_tokens / kCluster;
  const uint32_t tb = p.paged.token_bytes;
  const uint32_t win_lo = c * p.chunk_tokens + rank * W;  // op-relative token index
  uint32_t n_valid = 0;
  // tokens of this chunk = end of its last run (runs are sorted by b): a batch of several requests may hold
  // partial chunks anywhere, not only at its end
  uint32_t chunk_hi = c * p.chunk_tokens;
  {
    const uint32_t rr0 = p.chunk_run_off[c], rr1 = p.chunk_run_off[c + 1];
    if (rr1 > rr0) chunk_hi = static_cast<uint32_t>(p.runs[rr1 - 1].b + p.runs[rr1 - 1].n);
  }
  if (win_lo < chunk_hi) n_valid = min(W, chunk_hi - win_lo);
  if (win_lo + n_valid > chunk_hi) n_valid = chunk_hi > win_lo ? chunk_hi - win_lo : 0;

  if (threadIdx.x < kMaxHeads) {
    s_absmax[threadIdx.x] = 0;
    s_all[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();  // barrier initialised before any lane can signal it
  if (threadIdx.x < 32) {
    // warp 0 walks the chunk's run list in parallel (one run per lane): the table reads are one
    // L2 round trip instead of a serial chain, and every lane issues its own bulk copy.
    uint32_t mine = 0;
    const uint32_t r0 = p.chunk_run_off[c], r1 = p.chunk_run_off[c + 1];
    const uint32_t bs = p.paged.block_tokens;
    for (uint32_t r = r0 + threadIdx.x; r < r1; r += 32) {
      const Run run = p.runs[r];
      const int32_t lo = max(run.b, static_cast<int32_t>(win_lo));
      const int32_t hi = min(run.b + run.n, static_cast<int32_t>(win_lo + n_valid));
      if (hi <= lo) continue;
      const uint32_t slot = static_cast<uint32_t>(run.a + (lo - run.b));
      const uint32_t rel = static_cast<uint32_t>(lo - static_cast<int32_t>(win_lo));
      const uint32_t n = static_cast<uint32_t>(hi - lo);
      if (!p.hnd || (n == bs && (slot % bs) == 0 && (rel % bs) == 0)) {
        // NHD run, or a whole HND tile: one contiguous piece
        bulk_g2s(smem + static_cast<size_t>(rel) * tb, reinterpret_cast<const void*>(paged_addr(p.paged, plane, slot)),
                 n * tb, &bar);
        mine += n * tb;
      } else {
        // partial HND tile: one piece per head (runs never cross a tile on either side)
        const uint32_t row = p.head_bytes;  // D*2
        uint8_t* tile = smem + static_cast<size_t>(rel / bs) * bs * tb + static_cast<size_t>(rel % bs) * row;
        for (uint32_t h = 0; h < p.n_heads; ++h) {
          bulk_g2s(tile + static_cast<size_t>(h) * bs * row,
                   reinterpret_cast<const void*>(paged_addr_hnd(p.paged, plane, slot, h, row)), n * row, &bar);
          mine += n * row;
        }
      }
    }
    const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
    // complete_tx may land before this expect_tx: the phase cannot complete until the arrive
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, total);
  }
  // Cluster barrier #1, split: arrive now (publishes the zeroed s_all), wait only just before the
  // DSMEM pushes — its latency hides behind the bulk loads that are already in flight.
  cluster_arrive_release();
  mbar_wait(&bar, 0);

  const uint32_t vpt = tb >> 4;  // 16-byte vectors per token
  // HND window = W/bs whole tiles [H][bs][D]: vector idx = ((tile*H + h)*bs + row)*rv + c.
  const uint32_t rv = p.head_bytes >> 4;                 // 16-byte vectors per (token, head) row
  const uint32_t tile_rows = p.paged.block_tokens * rv;  // vectors per (tile, head)
  const uint32_t n_tiles = W / p.paged.block_tokens;
  if (p.hnd) {
    // One warp per head (heads w, w+8, ...): the valid rows of a (tile, head) block are a prefix of it
    // ([bs][D] rows are contiguous), so the warp streams the block with 16-byte loads, folds the lanes
    // with one REDUX and owns s_absmax[h] — no index arithmetic per vector, no atomics.
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t h = threadIdx.x >> 5; h < p.n_heads; h += kThreads >> 5) {
      uint32_t acc = 0;
      for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        const uint32_t t0 = tile * p.paged.block_tokens;
        const uint32_t nv = (n_valid > t0 ? min(p.paged.block_tokens, n_valid - t0) : 0u) * rv;
        const uint8_t* blk = smem + static_cast<size_t>(tile * p.n_heads + h) * tile_rows * 16;
#pragma unroll 8
        for (uint32_t v = lane; v < nv; v += 32) {
          const uint4 x = *reinterpret_cast<const uint4*>(blk + static_cast<size_t>(v) * 16);
          acc = absmax_u16x2(absmax_u16x2(absmax_u16x2(absmax_u16x2(acc, x.x), x.y), x.z), x.w);
        }
      }
      const uint32_t m = __reduce_max_sync(0xffffffffu, max(acc & 0xffffu, acc >> 16));
      if (lane == 0) s_absmax[h] = m;
    }
  } else {
  // ---- per-head absmax over this CTA's tokens: a thread's 16-byte column has a fixed head ----
    // When a token has fewer than 256 vectors the spare threads split the tokens between them.
    const uint32_t groups = vpt < kThreads ? kThreads / vpt : 1;
    const uint32_t grp = threadIdx.x / vpt;
    if (grp < groups) {
      const uint32_t col_step = groups == 1 ? kThreads : vpt;
      for (uint32_t col = threadIdx.x - grp * vpt; col < vpt; col += col_step) {
        uint32_t acc = 0;
        const uint8_t* q = smem + static_cast<size_t>(col) * 16;
        for (uint32_t t = grp; t < n_valid; t += groups) {
          const uint4 v = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(t) * tb);
          acc = absmax_u16x2(absmax_u16x2(absmax_u16x2(absmax_u16x2(acc, v.x), v.y), v.z), v.w);
        }
        const uint32_t m = max(acc & 0xffffu, acc >> 16);
        atomicMax(&s_absmax[(col * 16) / p.head_bytes], m);
      }
    }
  }
  __syncthreads();
  cluster_wait_acquire();  // #1: every CTA of the cluster has zeroed its s_all
  if (threadIdx.x < p.n_heads) {
    const uint32_t mine = s_absmax[threadIdx.x];
#pragma unroll
    for (uint32_t r = 0; r < kCluster; ++r) red_dsmem_max_u32(&s_all[threadIdx.x], r, mine);
  }
  // Cluster barrier #2: all pushes have landed everywhere.  Nothing remote is touched after it, so
  // no trailing barrier is needed to keep shared memory alive.
  cluster_arrive_release();
  cluster_wait_acquire();

  if (threadIdx.x < p.n_heads) {
    const uint32_t m = s_all[threadIdx.x];
    const float amax = __uint_as_float(m << 16);
    const float inv = (m == 0) ? 1.0f : __fdiv_rn(448.0f, amax);
    s_inv[threadIdx.x] = inv;
    if (rank == 0) {
      float* scales = reinterpret_cast<float*>(p.chunk_addrs[c] + p.scales_off);
      scales[plane * p.n_heads + threadIdx.x] = (m == 0) ? 1.0f : __fdiv_rn(amax, 448.0f);
    }
  }
  __syncthreads();

  if (p.hnd) {
    // the packed slab mirrors the tiles at half size: same vector index, 8 bytes each
    uint8_t* outh = reinterpret_cast<uint8_t*>(p.chunk_addrs[c] + static_cast<uint64_t>(plane) * p.slab_q_bytes +
                                               static_cast<uint64_t>(rank * W) * (tb >> 1));
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t h = threadIdx.x >> 5; h < p.n_heads; h += kThreads >> 5) {
      const float inv = s_inv[h];
      for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        const uint32_t t0 = tile * p.paged.block_tokens;
        const uint32_t nv = (n_valid > t0 ? min(p.paged.block_tokens, n_valid - t0) : 0u) * rv;
        const size_t base = static_cast<size_t>(tile * p.n_heads + h) * tile_rows;
#pragma unroll 8
        for (uint32_t v = lane; v < nv; v += 32) {
          const uint4 x = *reinterpret_cast<const uint4*>(smem + (base + v) * 16);
          st_na_v2(outh + (base + v) * 8, quant8(x, inv));
        }
      }
    }
    return;
  }

  // ---- quantise from smem, 8-byte coalesced stores ----
  uint8_t* out = reinterpret_cast<uint8_t*>(p.chunk_addrs[c] + static_cast<uint64_t>(plane) * p.slab_q_bytes +
                                            static_cast<uint64_t>(rank * W) * (tb >> 1));
  const uint32_t nvec = n_valid * vpt;
  if ((kThreads % vpt) == 0) {
    // i % vpt == threadIdx.x % vpt for every i of this thread: one head, one scale, no division
    const float inv = s_inv[((threadIdx.x % vpt) * 16) / p.head_bytes];
    for (uint32_t i = threadIdx.x; i < nvec; i += kThreads) {
      const uint4 v = *reinterpret_cast<const uint4*>(smem + static_cast<size_t>(i) * 16);
      st_na_v2(out + static_cast<size_t>(i) * 8, quant8(v, inv));
    }
  } else {
    for (uint32_t i = threadIdx.x; i < nvec; i += kThreads) {
      const uint32_t col = i % vpt;
      const uint4 v = *reinterpret_cast<const uint4*>(smem + static_cast<size_t>(i) * 16);
      st_na_v2(out + static_cast<size_t>(i) * 8, quant8(v, s_inv[(col * 16) / p.head_bytes]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FP8 store, persistent warp-specialised pipeline (the default for block-structured ops).
//
// Unit of work = (chunk, plane, head): the C x D bf16 elements one scale covers (64 KiB for C=256,
// D=128), i.e. exactly what has to be seen before anything can be quantised.  Nothing is exchanged
// between CTAs: no cluster, no DSMEM, no atomics.  One CTA per SM:
//   warp 0           producer: for every unit one TMA copy per paged block ([bs][D] of head h, 4 KiB) into the
//                    unit buffer of the consumer group that will take it, completion on its `full` mbarrier.
//                    HND tiles: the piece is contiguous -> cp.async.bulk (UBLKCP).  NHD tiles: the piece is
//                    bs rows of D elements strided by the token size -> cp.async.bulk.tensor.4d through a
//                    per-plane tensor map {D, H, bs, NB}, box {D, 1, bs, 1} (UTMALDG).  Either way the unit
//                    lands as [C tokens][D] in shared memory.
//   2 x 8 warps      two consumer groups ping-pong over the units: wait `full`, copy the unit into registers
//                    (16 x 16 B per thread, conflict-free), release the stage at once (`empty`), absmax
//                    (packed u16 max, one REDUX per warp, one named barrier per group), quantise from
//                    registers, 8-byte stores that fill whole 128-byte lines.
// A group releases its buffer as soon as the unit is in registers, so while it reduces and quantises unit i its
// unit i+2 is already streaming in, and so is the other group's: the memory pipe never drains between the load -> absmax -> quantise phases of a unit (the cluster kernel above
// runs at 33 % warps active with exactly that bubble).  HBM is read once and written once.
// Eligibility (checked on the host, else the cluster kernel runs): every chunk is a sequence of whole,
// block-aligned runs, the last one possibly short — what any vLLM block table produces.
// ---------------------------------------------------------------------------------------------
constexpr int kS3Groups = 2;
constexpr int kS3Stages = kS3Groups;   // stage g is filled for, and drained by, consumer group g only: consecutive uses of a
                                       // stage are strictly ordered, so a parity wait can never alias an older phase (with a
                                       // shared ring a group could ask for use k of a stage whose use k-1 — the OTHER group's
                                       // unit — had not landed yet, pass on the parity of use k-2 and corrupt the barrier)
constexpr int kS3GroupThreads = 256;
constexpr int kS3Threads = 32 + kS3Groups * kS3GroupThreads;
constexpr int kS3MaxVec = 16;          // 16-byte vectors per consumer thread per unit
constexpr uint32_t kS3MaxUnitBytes = kS3MaxVec * kS3GroupThreads * 16;   // 64 KiB

struct Fp8Store3Params {
  PagedSide paged;
  const Run* runs;                // run j of chunk c = paged block j of the chunk (host-checked)
  const uint32_t* chunk_run_off;  // [n_chunks+1]
  const uint64_t* chunk_addrs;    // [n_chunks]
  const void* tmaps;              // device array of CUtensorMap, one per plane (NHD); nullptr: plain bulk copies (HND)
  uint32_t n_chunks, n_planes;
  uint32_t chunk_tokens, n_tokens;
  uint32_t n_heads, head_bytes;   // H, D*2
  uint64_t slab_q_bytes, scales_off;
  uint32_t hnd;
  uint32_t total_units;           // n_chunks * n_planes * H
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 4-D tiled TMA load global -> shared (this CTA), completion on an mbarrier.  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int32_t c0, int32_t c1, int32_t c2,
                                            int32_t c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

__global__ void __launch_bounds__(kS3Threads, 1) kv_fp8_store3_kernel(const Fp8Store3Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kS3Stages];
  __shared__ __align__(8) uint64_t empty_bar[kS3Stages];
  __shared__ uint32_t s_max[kS3Groups][2][kS3GroupThreads / 32];

  const uint32_t bs = p.paged.block_tokens;
  const uint32_t piece_bytes = bs * p.head_bytes;            // one block's rows of one head
  const uint32_t unit_bytes = p.chunk_tokens * p.head_bytes;  // stage size (<= 64 KiB, host-checked)
  const uint32_t H = p.n_heads;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kS3Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kS3GroupThreads / 32);   // one arrive per warp of the consuming group
    }
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t n_mine = p.total_units > blockIdx.x ? (p.total_units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x < 32) {
    // ------------------------------- producer -------------------------------
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = 0; i < n_mine; ++i) {
      const uint32_t u = blockIdx.x + i * gridDim.x;
      const uint32_t h = u % H;
      const uint32_t plane = (u / H) % p.n_planes;
      const uint32_t c = u / (H * p.n_planes);
      const uint32_t s = i % kS3Groups, k = i / kS3Groups;    // stage == consumer group; k-th use of that stage
      const uint32_t r0 = __ldg(p.chunk_run_off + c), r1 = __ldg(p.chunk_run_off + c + 1);
      const uint32_t n_blocks = r1 - r0;
      if (k > 0) mbar_wait(&empty_bar[s], (k - 1) & 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], n_blocks * piece_bytes);
      __syncwarp();
      uint8_t* stage = smem + static_cast<size_t>(s) * unit_bytes;
      for (uint32_t j = lane; j < n_blocks; j += 32) {
        const uint32_t block = static_cast<uint32_t>(p.runs[r0 + j].a) / bs;
        uint8_t* dst = stage + static_cast<size_t>(j) * piece_bytes;
        if (p.tmaps) {
          const uint8_t* tm = static_cast<const uint8_t*>(p.tmaps) + static_cast<size_t>(plane) * 128;
          if (p.hnd) tma_load_4d(dst, tm, 0, 0, static_cast<int32_t>(h), static_cast<int32_t>(block), &full_bar[s]);
          else tma_load_4d(dst, tm, 0, static_cast<int32_t>(h), 0, static_cast<int32_t>(block), &full_bar[s]);
        } else {
          const uint64_t src = __ldg(p.paged.bases + plane) + static_cast<uint64_t>(block) * p.paged.block_stride +
                               static_cast<uint64_t>(h) * piece_bytes;   // HND tile [H][bs][D]
          bulk_g2s(dst, reinterpret_cast<const void*>(src), piece_bytes, &full_bar[s]);
        }
      }
    }
    return;
  }

  // ------------------------------- consumers -------------------------------
  const uint32_t g = (threadIdx.x - 32) / kS3GroupThreads;
  const uint32_t t = (threadIdx.x - 32) % kS3GroupThreads;
  const uint32_t warp = t >> 5, lane = t & 31u;
  const uint32_t rv = p.head_bytes >> 4;          // 16-byte vectors per (token, head) row; power of two (host-checked)
  const uint32_t col8 = (t % rv) * 8;             // byte offset of this thread's 8 output bytes inside a row
  const uint32_t tok0 = t / rv;                   // token of vector kk = tok0 + kk * tok_step
  const uint32_t tok_step = kS3GroupThreads / rv;
  const uint32_t out_row = p.head_bytes >> 1;     // D bytes of e4m3 per (token, head)
  uint32_t it = 0;
  for (uint32_t i = g; i < n_mine; i += kS3Groups, ++it) {
    const uint32_t u = blockIdx.x + i * gridDim.x;
    const uint32_t h = u % H;
    const uint32_t plane = (u / H) % p.n_planes;
    const uint32_t c = u / (H * p.n_planes);
    const uint32_t s = g, k = it;
    // run j of the chunk is its paged block j, all full but possibly the last (host-checked)
    const uint32_t cr0 = __ldg(p.chunk_run_off + c), cr1 = __ldg(p.chunk_run_off + c + 1);
    const uint32_t n_valid = (cr1 - cr0 - 1) * bs + static_cast<uint32_t>(p.runs[cr1 - 1].n);
    const uint8_t* stage = smem + static_cast<size_t>(s) * unit_bytes;

    mbar_wait(&full_bar[s], k & 1);
    uint4 x[kS3MaxVec];
    uint32_t acc = 0;
#pragma unroll
    for (int kk = 0; kk < kS3MaxVec; ++kk) {
      const uint32_t tok = tok0 + kk * tok_step;
      if (tok < n_valid) {
        x[kk] = *reinterpret_cast<const uint4*>(stage + (static_cast<size_t>(kk) * kS3GroupThreads + t) * 16);
        acc = absmax_u16x2(absmax_u16x2(absmax_u16x2(absmax_u16x2(acc, x[kk].x), x[kk].y), x[kk].z), x[kk].w);
      } else {
        x[kk] = make_uint4(0, 0, 0, 0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);    // the unit lives in registers now: the stage can be refilled

    const uint32_t wmax = __reduce_max_sync(0xffffffffu, max(acc & 0xffffu, acc >> 16));
    const uint32_t par = it & 1;
    if (lane == 0) s_max[g][par][warp] = wmax;
    named_bar_sync(1 + g, kS3GroupThreads);
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < kS3GroupThreads / 32; ++w) m = max(m, s_max[g][par][w]);
    const float amax = __uint_as_float(m << 16);
    const float inv = (m == 0) ? 1.0f : __fdiv_rn(448.0f, amax);
    const uint64_t chunk_base = __ldg(p.chunk_addrs + c);
    if (t == 0) {
      float* scales = reinterpret_cast<float*>(chunk_base + p.scales_off);
      scales[plane * H + h] = (m == 0) ? 1.0f : __fdiv_rn(amax, 448.0f);
    }
    uint8_t* slab = reinterpret_cast<uint8_t*>(chunk_base + static_cast<uint64_t>(plane) * p.slab_q_bytes);
    // output offset of vector kk is linear in kk: NHD [token][H][D]; HND [tile][H][bs][D] when a step of
    // tok_step tokens is a whole number of tiles (the host checks tok_step % bs == 0 before choosing this kernel)
    uint32_t off0, off_step;
    if (p.hnd) {
      const uint32_t tile0 = tok0 / bs, row0 = tok0 - tile0 * bs;
      off0 = ((tile0 * H + h) * bs + row0) * out_row + col8;
      off_step = (tok_step / bs) * H * bs * out_row;
    } else {
      off0 = (tok0 * H + h) * out_row + col8;
      off_step = tok_step * H * out_row;
    }
    uint8_t* dst = slab + off0;
#pragma unroll
    for (int kk = 0; kk < kS3MaxVec; ++kk) {
      if (tok0 + kk * tok_step < n_valid) st_na_v2(dst + static_cast<size_t>(kk) * off_step, quant8(x[kk], inv));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FP8 store, two-pass variant (B200KV_FP8_2PASS=1; experimental): same grid, cluster and output as
// kv_fp8_store_kernel, but no shared-memory staging.  Pass 1 streams the CTA's window straight from
// the pages into the per-head absmax; after the cluster exchange pass 2 reads the same bytes again —
// an L2 hit, the slab was touched microseconds ago — quantises and stores.  HBM is still read once;
// without the 64 KiB of smem per CTA residency is bounded by registers/threads, not by smem.
// Per token the kernel keeps one source address in smem (NHD: the token's row; HND: row 0 of head 0
// of its tile position), filled by warp 0 from the run list.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxWindow = 64;  // tokens per CTA window (C / kCluster)
constexpr int kBatch = 8;       // 16-byte loads a thread keeps in flight

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(256, 4)
    kv_fp8_store2_kernel(const Fp8StoreParams p) {
  __shared__ uint64_t s_src[kMaxWindow];
  __shared__ uint32_t s_absmax[kMaxHeads];
  __shared__ uint32_t s_all[kMaxHeads];
  __shared__ float s_inv[kMaxHeads];

  const uint32_t rank = cluster_ctarank();
  const uint32_t slab = blockIdx.x / kCluster;
  const uint32_t c = slab / p.n_planes;
  const uint32_t plane = slab - c * p.n_planes;
  const uint32_t W = p.chunk_tokens / kCluster;
  const uint32_t tb = p.paged.token_bytes;
  const uint32_t bs = p.paged.block_tokens;
  const uint32_t win_lo = c * p.chunk_tokens + rank * W;
  uint32_t n_valid = 0;
  // tokens of this chunk = end of its last run (runs are sorted by b): a batch of several requests may hold
  // partial chunks anywhere, not only at its end
  uint32_t chunk_hi = c * p.chunk_tokens;
  {
    const uint32_t rr0 = p.chunk_run_off[c], rr1 = p.chunk_run_off[c + 1];
    if (rr1 > rr0) chunk_hi = static_cast<uint32_t>(p.runs[rr1 - 1].b + p.runs[rr1 - 1].n);
  }
  if (win_lo < chunk_hi) n_valid = min(W, chunk_hi - win_lo);
  if (win_lo + n_valid > chunk_hi) n_valid = chunk_hi > win_lo ? chunk_hi - win_lo : 0;

  if (threadIdx.x < kMaxHeads) {
    s_absmax[threadIdx.x] = 0;
    s_all[threadIdx.x] = 0;
  }
  if (threadIdx.x < 32) {
    const uint32_t r0 = p.chunk_run_off[c], r1 = p.chunk_run_off[c + 1];
    for (uint32_t r = r0 + threadIdx.x; r < r1; r += 32) {
      const Run run = p.runs[r];
      const int32_t lo = max(run.b, static_cast<int32_t>(win_lo));
      const int32_t hi = min(run.b + run.n, static_cast<int32_t>(win_lo + n_valid));
      for (int32_t t = lo; t < hi; ++t) {
        const uint32_t slot = static_cast<uint32_t>(run.a + (t - run.b));
        s_src[t - static_cast<int32_t>(win_lo)] =
            p.hnd ? paged_addr_hnd(p.paged, plane, slot, 0, p.head_bytes) : paged_addr(p.paged, plane, slot);
      }
    }
  }
  __syncthreads();
  cluster_arrive_release();  // split barrier #1: s_all is zeroed everywhere before any push

  const uint32_t vpt = tb >> 4;            // 16-byte vectors per token
  const uint32_t rv = p.head_bytes >> 4;   // ... per (token, head) row
  const uint32_t head_stride = p.hnd ? bs * p.head_bytes : p.head_bytes;  // bytes between heads of one token
  // thread -> fixed column (head h, vector cc) ; token groups share the columns when vpt < 256
  const uint32_t groups = vpt < 256 ? 256 / vpt : 1;
  const uint32_t grp = threadIdx.x / vpt;
  const bool active = grp < groups;
  const uint32_t col0 = threadIdx.x - grp * vpt;
  const uint32_t col_step = groups == 1 ? 256u : vpt;

  // ---- pass 1: absmax straight from the pages ----
  if (active) {
    for (uint32_t col = col0; col < vpt; col += col_step) {
      const uint32_t h = col / rv, cc = col - h * rv;
      const uint64_t off = static_cast<uint64_t>(h) * head_stride + static_cast<uint64_t>(cc) * 16;
      uint32_t acc = 0;
      for (uint32_t t0 = grp; t0 < n_valid; t0 += groups * kBatch) {
        uint4 v[kBatch];  // all loads of a batch are issued before the first one is consumed
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const uint32_t t = t0 + j * groups;
          v[j] = t < n_valid ? ld_nc_v4(reinterpret_cast<const void*>(s_src[t] + off)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
          acc = absmax_u16x2(absmax_u16x2(absmax_u16x2(absmax_u16x2(acc, v[j].x), v[j].y), v[j].z), v[j].w);
      }
      const uint32_t m = max(acc & 0xffffu, acc >> 16);
      if (m) atomicMax(&s_absmax[h], m);
    }
  }
  __syncthreads();
  cluster_wait_acquire();  // #1
  if (threadIdx.x < p.n_heads) {
    const uint32_t mine = s_absmax[threadIdx.x];
#pragma unroll
    for (uint32_t r = 0; r < kCluster; ++r) red_dsmem_max_u32(&s_all[threadIdx.x], r, mine);
  }
  cluster_arrive_release();  // #2: all pushes have landed; nothing remote is touched afterwards
  cluster_wait_acquire();
  if (threadIdx.x < p.n_heads) {
    const uint32_t m = s_all[threadIdx.x];
    const float amax = __uint_as_float(m << 16);
    s_inv[threadIdx.x] = (m == 0) ? 1.0f : __fdiv_rn(448.0f, amax);
    if (rank == 0) {
      float* scales = reinterpret_cast<float*>(p.chunk_addrs[c] + p.scales_off);
      scales[plane * p.n_heads + threadIdx.x] = (m == 0) ? 1.0f : __fdiv_rn(amax, 448.0f);
    }
  }
  __syncthreads();

  // ---- pass 2: the same bytes again (L2), quantise, store ----
  uint8_t* out = reinterpret_cast<uint8_t*>(p.chunk_addrs[c] + static_cast<uint64_t>(plane) * p.slab_q_bytes +
                                            static_cast<uint64_t>(rank * W) * (tb >> 1));
  if (active) {
    for (uint32_t col = col0; col < vpt; col += col_step) {
      const uint32_t h = col / rv, cc = col - h * rv;
      const uint64_t off = static_cast<uint64_t>(h) * head_stride + static_cast<uint64_t>(cc) * 16;
      const float inv = s_inv[h];
      for (uint32_t t0 = grp; t0 < n_valid; t0 += groups * kBatch) {
        uint4 v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const uint32_t t = t0 + j * groups;
          v[j] = t < n_valid ? ld_nc_v4(reinterpret_cast<const void*>(s_src[t] + off)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const uint32_t t = t0 + j * groups;
          if (t >= n_valid) break;
          // NHD chunk: token-major; HND chunk: tiles verbatim, ((tile*H + h)*bs + row)*rv + cc
          const size_t idx = p.hnd ? (static_cast<size_t>((t / bs) * p.n_heads + h) * bs + (t % bs)) * rv + cc
                                   : static_cast<size_t>(t) * vpt + col;
          st_na_v2(out + idx * 8, quant8(v[j], inv));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FP8 load: one CTA per (run, plane) unit; e4m3 * scale -> bf16 (RN) into the pages.
// ---------------------------------------------------------------------------------------------
struct Fp8LoadParams {
  PagedSide paged;
  const Run* runs;
  const uint64_t* chunk_addrs;
  uint32_t n_runs, n_planes;
  uint32_t chunk_tokens;
  uint32_t n_heads, head_bytes;  // head_bytes = D*2 (bf16 side)
  uint64_t slab_q_bytes;
  uint64_t scales_off;
  uint32_t total_units;
  uint32_t hnd;
  uint32_t plane_begin;
};

__device__ __forceinline__ uint4 dequant8(const uint2& q, float scale) {
  const uint32_t w[2] = {q.x, q.y};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __half2_raw h0 = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(w[i] & 0xffffu), __NV_E4M3);
    const __half2_raw h1 = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(w[i] >> 16), __NV_E4M3);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
    const __nv_bfloat162 b0 = __floats2bfloat162_rn(f0.x * scale, f0.y * scale);
    const __nv_bfloat162 b1 = __floats2bfloat162_rn(f1.x * scale, f1.y * scale);
    o[2 * i] = *reinterpret_cast<const uint32_t*>(&b0);
    o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b1);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(kFp8Threads) kv_fp8_load_kernel(const Fp8LoadParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ float s_scale[kMaxHeads];
  const uint32_t tb = p.paged.token_bytes;  // bf16 bytes per token
  const uint32_t qtb = tb >> 1;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t phase = 0;
  for (uint32_t ui = blockIdx.x; ui < p.total_units; ui += gridDim.x, phase ^= 1u) {
    const uint32_t plane = p.plane_begin + ui % p.n_planes;
    const Run run = p.runs[ui / p.n_planes];
    const uint32_t c = static_cast<uint32_t>(run.b) / p.chunk_tokens;
    const uint32_t t = static_cast<uint32_t>(run.b) - c * p.chunk_tokens;
    const uint64_t cbase = __ldg(p.chunk_addrs + c);
    const uint32_t bs = p.paged.block_tokens;
    // NHD: the run's n tokens are contiguous in the packed slab.  HND: fetch the whole packed tile
    // ([H][bs][D] bytes) the run lives in and pick its rows.
    const uint32_t q_off = p.hnd ? (t / bs) * bs * qtb : t * qtb;
    const uint32_t qbytes = p.hnd ? bs * qtb : static_cast<uint32_t>(run.n) * qtb;
    if (threadIdx.x == 0) {
      bulk_g2s(smem, reinterpret_cast<const void*>(cbase + static_cast<uint64_t>(plane) * p.slab_q_bytes + q_off),
               qbytes, &bar);
      mbar_arrive_expect_tx(&bar, qbytes);
    }
    if (threadIdx.x < p.n_heads)
      s_scale[threadIdx.x] =
          reinterpret_cast<const float*>(cbase + p.scales_off)[plane * p.n_heads + threadIdx.x];
    __syncthreads();
    mbar_wait(&bar, phase);
    const uint32_t vpt = tb >> 4;
    if (p.hnd) {
      const uint32_t rv = p.head_bytes >> 4;          // 16-byte output vectors per (token, head) row
      const uint32_t o_src = t % bs;                    // first row inside the packed tile
      const uint32_t n = static_cast<uint32_t>(run.n);
      const uint32_t total = p.n_heads * n * rv;
      for (uint32_t e = threadIdx.x; e < total; e += kFp8Threads) {
        const uint32_t cc = e % rv;
        const uint32_t r = (e / rv) % n;
        const uint32_t h = e / (rv * n);
        const uint2 q = *reinterpret_cast<const uint2*>(smem + (static_cast<size_t>(h) * bs + o_src + r) * (p.head_bytes >> 1) +
                                                        static_cast<size_t>(cc) * 8);
        uint8_t* d = reinterpret_cast<uint8_t*>(paged_addr_hnd(p.paged, plane, static_cast<uint32_t>(run.a), h, p.head_bytes));
        st_na_v4(d + static_cast<size_t>(r) * p.head_bytes + static_cast<size_t>(cc) * 16, dequant8(q, s_scale[h]));
      }
      __syncthreads();
      continue;
    }
    uint8_t* dst = reinterpret_cast<uint8_t*>(paged_addr(p.paged, plane, static_cast<uint32_t>(run.a)));
    const uint32_t nvec = static_cast<uint32_t>(run.n) * vpt;
    if ((kFp8Threads % vpt) == 0) {
      const float sc = s_scale[((threadIdx.x % vpt) * 16) / p.head_bytes];
      for (uint32_t i = threadIdx.x; i < nvec; i += kFp8Threads) {
        const uint2 q = *reinterpret_cast<const uint2*>(smem + static_cast<size_t>(i) * 8);
        st_na_v4(dst + static_cast<size_t>(i) * 16, dequant8(q, sc));
      }
    } else {
      for (uint32_t i = threadIdx.x; i < nvec; i += kFp8Threads) {
        const uint32_t col = i % vpt;
        const uint2 q = *reinterpret_cast<const uint2*>(smem + static_cast<size_t>(i) * 8);
        st_na_v4(dst + static_cast<size_t>(i) * 16, dequant8(q, s_scale[(col * 16) / p.head_bytes]));
      }
    }
    __syncthreads();  // smem + s_scale reusable
  }
}

// ---------------------------------------------------------------------------------------------
// Q4: group-wise 4-bit format (B200KV_FMT_Q4; oracle: q4_pack_chunk / q4_unpack_chunk).  A stored
// token of one plane is a record [H*D/2 code bytes][H*D/32 bf16 scales]; records are token-major in the
// slab whatever the order inside the paged tiles.  One 16-byte vector (8 elements) per thread, a group
// of 32 elements = 4 adjacent lanes: absmax by two shuffles, no shared memory, no cluster.
// EXPERIMENTAL: written against the oracle, not yet run on a GPU.
// ---------------------------------------------------------------------------------------------
struct Q4Params {
  PagedSide paged;
  const Run* runs;
  const uint64_t* chunk_addrs;
  uint32_t n_runs, n_planes, plane_begin;
  uint32_t chunk_tokens;
  uint32_t n_heads, head_bytes;   // head_bytes = D*2 (bf16 side)
  uint64_t slab_bytes;            // C * rec_bytes
  uint32_t rec_bytes;             // H*D/2 + H*D/32*2
  uint32_t total_units;           // n_runs * n_planes
  uint32_t hnd;
};

__device__ __forceinline__ uint64_t q4_src_addr(const Q4Params& p, uint32_t plane, uint32_t slot, uint32_t v) {
  // v = 16-byte vector index inside the token: head h = v / rv, vector c of that head
  if (!p.hnd) return paged_addr(p.paged, plane, slot) + static_cast<uint64_t>(v) * 16;
  const uint32_t rv = p.head_bytes >> 4;
  const uint32_t h = v / rv, c = v - h * rv;
  return paged_addr_hnd(p.paged, plane, slot, h, p.head_bytes) + static_cast<uint64_t>(c) * 16;
}

constexpr int kQ4Unroll = 4;   // 16-byte vectors a thread keeps in flight (the kernels are pure streaming)

__device__ __forceinline__ void q4_split(uint32_t idx, uint32_t vpt, uint32_t vpt_shift, uint32_t* t, uint32_t* v) {
  if (vpt_shift != 0xffffffffu) { *t = idx >> vpt_shift; *v = idx & (vpt - 1); }
  else { *t = idx / vpt; *v = idx - *t * vpt; }
}

// q = rint(x * inv) (exact product, ties to even), inv = fl32(1 / s), s = bf16(absmax_group / 7)
// (oracle/kv_oracle.py q4_pack_chunk): one IEEE division per group, then one FMA per element.
__global__ void __launch_bounds__(256) kv_q4_store_kernel(const Q4Params p) {
  const uint32_t vpt = (p.n_heads * p.head_bytes) >> 4;      // vectors per token (multiple of 4)
  const uint32_t vpt_shift = (vpt & (vpt - 1)) == 0 ? 31u - __clz(vpt) : 0xffffffffu;
  const uint32_t codes_bytes = vpt * 4;
  for (uint32_t ui = blockIdx.x; ui < p.total_units; ui += gridDim.x) {
    const uint32_t plane = p.plane_begin + ui % p.n_planes;
    const Run run = p.runs[ui / p.n_planes];
    const uint32_t c = static_cast<uint32_t>(run.b) / p.chunk_tokens;
    const uint32_t t0 = static_cast<uint32_t>(run.b) - c * p.chunk_tokens;
    uint8_t* slab = reinterpret_cast<uint8_t*>(__ldg(p.chunk_addrs + c) + static_cast<uint64_t>(plane) * p.slab_bytes);
    const uint32_t nvec = static_cast<uint32_t>(run.n) * vpt;
    for (uint32_t base = 0; base < nvec; base += 256 * kQ4Unroll) {   // uniform trip count: every lane shuffles
      uint4 x[kQ4Unroll];
      uint32_t tt[kQ4Unroll], vv[kQ4Unroll];
      bool ok[kQ4Unroll];
#pragma unroll
      for (int j = 0; j < kQ4Unroll; ++j) {                 // all loads first: kQ4Unroll x 16 B in flight per thread
        const uint32_t idx = base + j * 256 + threadIdx.x;   // 256 and vpt are multiples of 4: quads stay whole
        ok[j] = idx < nvec;
        tt[j] = vv[j] = 0;
        if (ok[j]) q4_split(idx, vpt, vpt_shift, &tt[j], &vv[j]);
        x[j] = ok[j] ? ld_nc_v4(reinterpret_cast<const void*>(q4_src_addr(p, plane, static_cast<uint32_t>(run.a) + tt[j], vv[j])))
                     : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < kQ4Unroll; ++j) {
        uint32_t acc = absmax_u16x2(absmax_u16x2(absmax_u16x2(absmax_u16x2(0u, x[j].x), x[j].y), x[j].z), x[j].w);
        uint32_t m = max(acc & 0xffffu, acc >> 16);
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
        const float amax = __uint_as_float(m << 16);
        const __nv_bfloat16 sb = m ? __float2bfloat16_rn(__fdiv_rn(amax, 7.0f)) : __float2bfloat16_rn(1.0f);
        const float inv = __fdiv_rn(1.0f, __bfloat162float(sb));
        // rint(x * inv) for |x * inv| <= 7.03 without a convert: fma(x, inv, 1.5 * 2^23) rounds the EXACT product to
        // the nearest integer (ties to even) into the low mantissa bits, whose low nibble is the two's-complement
        // code.  No clamp is needed: |x| <= absmax and scale >= absmax / 7 * (1 - 2^-9).
        const uint32_t w[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
          const uint32_t ql = __float_as_uint(__fmaf_rn(lo, inv, 12582912.0f)) & 0xFu;
          const uint32_t qh = __float_as_uint(__fmaf_rn(hi, inv, 12582912.0f)) & 0xFu;
          packed |= (ql | (qh << 4)) << (8 * i);
        }
        if (!ok[j]) continue;
        uint8_t* rec = slab + static_cast<uint64_t>(t0 + tt[j]) * p.rec_bytes;
        *reinterpret_cast<uint32_t*>(rec + static_cast<size_t>(vv[j]) * 4) = packed;
        if ((vv[j] & 3u) == 0) *reinterpret_cast<__nv_bfloat16*>(rec + codes_bytes + static_cast<size_t>(vv[j] >> 2) * 2) = sb;
      }
    }
  }
}

__global__ void __launch_bounds__(256) kv_q4_load_kernel(const Q4Params p) {
  const uint32_t vpt = (p.n_heads * p.head_bytes) >> 4;
  const uint32_t vpt_shift = (vpt & (vpt - 1)) == 0 ? 31u - __clz(vpt) : 0xffffffffu;
  const uint32_t codes_bytes = vpt * 4;
  for (uint32_t ui = blockIdx.x; ui < p.total_units; ui += gridDim.x) {
    const uint32_t plane = p.plane_begin + ui % p.n_planes;
    const Run run = p.runs[ui / p.n_planes];
    const uint32_t c = static_cast<uint32_t>(run.b) / p.chunk_tokens;
    const uint32_t t0 = static_cast<uint32_t>(run.b) - c * p.chunk_tokens;
    const uint8_t* slab =
        reinterpret_cast<const uint8_t*>(__ldg(p.chunk_addrs + c) + static_cast<uint64_t>(plane) * p.slab_bytes);
    const uint32_t nvec = static_cast<uint32_t>(run.n) * vpt;
    for (uint32_t base = threadIdx.x; base < nvec; base += 256 * kQ4Unroll) {
      uint32_t packed[kQ4Unroll], tt[kQ4Unroll], vv[kQ4Unroll];
      uint16_t sbits[kQ4Unroll];
#pragma unroll
      for (int j = 0; j < kQ4Unroll; ++j) {                 // all loads first
        const uint32_t idx = base + j * 256;
        packed[j] = 0;
        sbits[j] = 0;
        tt[j] = vv[j] = 0;
        if (idx < nvec) {
          q4_split(idx, vpt, vpt_shift, &tt[j], &vv[j]);
          const uint8_t* rec = slab + static_cast<uint64_t>(t0 + tt[j]) * p.rec_bytes;
          packed[j] = __ldg(reinterpret_cast<const uint32_t*>(rec + static_cast<size_t>(vv[j]) * 4));
          sbits[j] = __ldg(reinterpret_cast<const uint16_t*>(rec + codes_bytes + static_cast<size_t>(vv[j] >> 2) * 2));
        }
      }
#pragma unroll
      for (int j = 0; j < kQ4Unroll; ++j) {
        if (base + j * 256 >= nvec) continue;
        const float s = __uint_as_float(static_cast<uint32_t>(sbits[j]) << 16);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int b = static_cast<int>((packed[j] >> (8 * i)) & 0xffu);
          const int ql = ((b & 0xF) ^ 8) - 8, qh = ((b >> 4) ^ 8) - 8;   // sign-extend the nibbles
          const __nv_bfloat162 r = __floats2bfloat162_rn(static_cast<float>(ql) * s, static_cast<float>(qh) * s);
          o[i] = *reinterpret_cast<const uint32_t*>(&r);
        }
        st_na_v4(reinterpret_cast<void*>(q4_src_addr(p, plane, static_cast<uint32_t>(run.a) + tt[j], vv[j])),
                 make_uint4(o[0], o[1], o[2], o[3]));
      }
    }
  }
}

}  // namespace b200kv
