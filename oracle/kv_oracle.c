/*
 * kv_oracle.c — plain-C CPU restatement of the KV offload hot path.  TEST INFRASTRUCTURE ONLY:
 * linked / loaded solely by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline and
 * --impl reference legs.  Never loaded by production-stack_b200.
 *
 * PARITY STATUS: parity unpinned for KV bytes (see oracle/kv_oracle.py header): the reference's
 * arithmetic is inside the absent third-party wheel lmcache==0.3.11
 * (/root/reference/pyproject.toml:49-52).  This file follows the same executable descriptions as
 * kv_oracle.py and is cross-checked against it in tests/test_oracle.py:
 *   gather : layer.reshape(2, NB*bs, -1)[:, slot_mapping]   vllm/.../v1/example_connector.py:247-248
 *   scatter: dst.reshape(2, NB*bs, -1)[:, slot_mapping] = src            example_connector.py:154-159
 *   layout : chunk = (L, 2, C, H, D)          vllm/.../lmcache_integration/vllm_v1_adapter.py:471-477
 *
 * Threads: pthreads over (chunk, plane) units — "all the host threads it can use" for the
 * reference arm of bench.py.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for (no libgomp in this image) --------------------------------- */
static int g_threads = 0;
int oracle_num_threads(void) {
  if (g_threads <= 0) {
    const char* e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = n < 1 ? 1 : (n > 256 ? 256 : (int)n);
  }
  return g_threads;
}
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }

typedef void (*unit_fn)(int64_t unit, void* arg);
typedef struct { unit_fn fn; void* arg; int64_t n; int tid, nt; } par_job;
static void* par_main(void* p) {
  par_job* j = (par_job*)p;
  for (int64_t u = j->tid; u < j->n; u += j->nt) j->fn(u, j->arg);
  return NULL;
}
static void parallel_for(int64_t n, unit_fn fn, void* arg) {
  int nt = oracle_num_threads();
  if (nt > n) nt = (int)(n > 0 ? n : 1);
  pthread_t th[256];
  par_job jobs[256];
  for (int t = 0; t < nt; ++t) {
    jobs[t].fn = fn; jobs[t].arg = arg; jobs[t].n = n; jobs[t].tid = t; jobs[t].nt = nt;
    if (t > 0) pthread_create(&th[t], NULL, par_main, &jobs[t]);
  }
  par_main(&jobs[0]);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
}

/* planes[2*l] = K plane of layer l (block 0), planes[2*l+1] = V plane. */
static inline const uint8_t* paged(const uint8_t* const* planes, int plane, int64_t slot,
                                   int bs, uint64_t block_stride, uint32_t token_bytes) {
  const int64_t blk = slot / bs, off = slot % bs;
  return planes[plane] + (uint64_t)blk * block_stride + (uint64_t)off * token_bytes;
}

typedef struct {
  uint8_t* const* planes;
  int n_planes, bs, chunk_tokens, n_heads, head_dim;
  uint64_t block_stride, chunk_bytes, scales_off;
  uint32_t token_bytes;
  const int64_t* slot_mapping;
  int64_t n_tokens;
  uint8_t* chunks;
} job_t;

#define UNIT_PROLOGUE                                                                   \
  const job_t* j = (const job_t*)arg;                                                   \
  const int64_t c = unit / j->n_planes;                                                 \
  const int p = (int)(unit % j->n_planes);                                              \
  const int64_t t0 = c * j->chunk_tokens;                                               \
  const int64_t t1 = t0 + j->chunk_tokens < j->n_tokens ? t0 + j->chunk_tokens : j->n_tokens; \
  const uint8_t* const* planes = (const uint8_t* const*)j->planes;                      \
  (void)planes

/* paged -> chunks laid back to back at `chunks` (chunk_bytes apart), RAW format */
static void gather_raw_unit(int64_t unit, void* arg) {
  UNIT_PROLOGUE;
  uint8_t* dst = j->chunks + (uint64_t)c * j->chunk_bytes + (uint64_t)p * j->chunk_tokens * j->token_bytes;
  for (int64_t t = t0; t < t1; ++t)
    memcpy(dst + (uint64_t)(t - t0) * j->token_bytes,
           paged(planes, p, j->slot_mapping[t], j->bs, j->block_stride, j->token_bytes), j->token_bytes);
}
static void scatter_raw_unit(int64_t unit, void* arg) {
  UNIT_PROLOGUE;
  const uint8_t* src = j->chunks + (uint64_t)c * j->chunk_bytes + (uint64_t)p * j->chunk_tokens * j->token_bytes;
  for (int64_t t = t0; t < t1; ++t)
    memcpy((uint8_t*)paged(planes, p, j->slot_mapping[t], j->bs, j->block_stride, j->token_bytes),
           src + (uint64_t)(t - t0) * j->token_bytes, j->token_bytes);
}

void oracle_gather_raw(const uint8_t* const* planes, int n_planes, uint64_t block_stride, int bs,
                       uint32_t token_bytes, const int64_t* slot_mapping, int64_t n_tokens,
                       int chunk_tokens, uint8_t* chunks, uint64_t chunk_bytes) {
  job_t j = {(uint8_t* const*)planes, n_planes, bs, chunk_tokens, 0, 0, block_stride, chunk_bytes, 0,
             token_bytes, slot_mapping, n_tokens, chunks};
  parallel_for((n_tokens + chunk_tokens - 1) / chunk_tokens * n_planes, gather_raw_unit, &j);
}
void oracle_scatter_raw(uint8_t* const* planes, int n_planes, uint64_t block_stride, int bs,
                        uint32_t token_bytes, const int64_t* slot_mapping, int64_t n_tokens,
                        int chunk_tokens, const uint8_t* chunks, uint64_t chunk_bytes) {
  job_t j = {planes, n_planes, bs, chunk_tokens, 0, 0, block_stride, chunk_bytes, 0,
             token_bytes, slot_mapping, n_tokens, (uint8_t*)chunks};
  parallel_for((n_tokens + chunk_tokens - 1) / chunk_tokens * n_planes, scatter_raw_unit, &j);
}

/* ---- e4m3fn codec (round-to-nearest-even, saturate-to-finite) ------------------------------ */
static inline float bf16_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
static inline uint8_t f32_to_e4m3(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint8_t sign = (uint8_t)((u >> 24) & 0x80);
  if (x != x) return (uint8_t)(0x7f | sign);
  double a = fabs((double)x);
  if (a > 448.0) a = 448.0;
  if (a < 0.015625) { /* subnormal grid 2^-9; 8 -> smallest normal 0x08 */
    return (uint8_t)((int)nearbyint(a * 512.0) | sign);
  }
  int ex;
  (void)frexp(a, &ex);
  int e = ex - 1;
  int q = (int)nearbyint(ldexp(a, 3 - e)); /* [8,16], ties-to-even under default rounding mode */
  if (q == 16) { q = 8; ++e; }
  return (uint8_t)((((e + 7) << 3) | (q - 8)) | sign);
}
static inline float e4m3_to_f32(uint8_t c) {
  const int e = (c >> 3) & 0xf, m = c & 7;
  float v;
  if (e == 0xf && m == 7) v = NAN;
  else if (e == 0) v = (float)ldexp((double)m, -9);
  else v = (float)ldexp((double)(8 + m), e - 10);
  return (c & 0x80) ? -v : v;
}

/* paged (bf16) -> FP8 chunks: per slab C*H*D e4m3 bytes, then (planes,H) fp32 scales at
 * scales_off.  One scale per (chunk, plane, head) = absmax/448. */
static void gather_fp8_unit(int64_t unit, void* arg) {
  UNIT_PROLOGUE;
  const int H = j->n_heads, D = j->head_dim;
  uint8_t* base = j->chunks + (uint64_t)c * j->chunk_bytes;
  float* scales = (float*)(base + j->scales_off) + (size_t)p * H;
  uint8_t* out = base + (uint64_t)p * j->chunk_tokens * H * D;
  for (int h = 0; h < H; ++h) {
    uint16_t amax = 0;
    for (int64_t t = t0; t < t1; ++t) {
      const uint16_t* v = (const uint16_t*)paged(planes, p, j->slot_mapping[t], j->bs, j->block_stride,
                                                 j->token_bytes) + (size_t)h * D;
      for (int d = 0; d < D; ++d) {
        const uint16_t m = v[d] & 0x7fff;
        if (m > amax) amax = m;
      }
    }
    const float fa = bf16_to_f32(amax);
    const float inv = amax ? 448.0f / fa : 1.0f;
    scales[h] = amax ? fa / 448.0f : 1.0f;
    for (int64_t t = t0; t < t1; ++t) {
      const uint16_t* v = (const uint16_t*)paged(planes, p, j->slot_mapping[t], j->bs, j->block_stride,
                                                 j->token_bytes) + (size_t)h * D;
      uint8_t* o = out + ((uint64_t)(t - t0) * H + h) * D;
      for (int d = 0; d < D; ++d) o[d] = f32_to_e4m3(bf16_to_f32(v[d]) * inv);
    }
  }
}
static void scatter_fp8_unit(int64_t unit, void* arg) {
  UNIT_PROLOGUE;
  const int H = j->n_heads, D = j->head_dim;
  const uint8_t* base = j->chunks + (uint64_t)c * j->chunk_bytes;
  const float* scales = (const float*)(base + j->scales_off) + (size_t)p * H;
  const uint8_t* in = base + (uint64_t)p * j->chunk_tokens * H * D;
  for (int64_t t = t0; t < t1; ++t) {
    uint16_t* dst = (uint16_t*)paged(planes, p, j->slot_mapping[t], j->bs, j->block_stride, j->token_bytes);
    const uint8_t* q = in + (uint64_t)(t - t0) * H * D;
    for (int h = 0; h < H; ++h)
      for (int d = 0; d < D; ++d)
        dst[h * D + d] = f32_to_bf16_rn(e4m3_to_f32(q[h * D + d]) * scales[h]);
  }
}

void oracle_gather_fp8(const uint8_t* const* planes, int n_planes, uint64_t block_stride, int bs,
                       int n_heads, int head_dim, const int64_t* slot_mapping, int64_t n_tokens,
                       int chunk_tokens, uint8_t* chunks, uint64_t chunk_bytes,
                       uint64_t scales_off) {
  job_t j = {(uint8_t* const*)planes, n_planes, bs, chunk_tokens, n_heads, head_dim, block_stride,
             chunk_bytes, scales_off, (uint32_t)(n_heads * head_dim * 2), slot_mapping, n_tokens, chunks};
  parallel_for((n_tokens + chunk_tokens - 1) / chunk_tokens * n_planes, gather_fp8_unit, &j);
}
void oracle_scatter_fp8(uint8_t* const* planes, int n_planes, uint64_t block_stride, int bs,
                        int n_heads, int head_dim, const int64_t* slot_mapping, int64_t n_tokens,
                        int chunk_tokens, const uint8_t* chunks, uint64_t chunk_bytes,
                        uint64_t scales_off) {
  job_t j = {planes, n_planes, bs, chunk_tokens, n_heads, head_dim, block_stride, chunk_bytes,
             scales_off, (uint32_t)(n_heads * head_dim * 2), slot_mapping, n_tokens, (uint8_t*)chunks};
  parallel_for((n_tokens + chunk_tokens - 1) / chunk_tokens * n_planes, scatter_fp8_unit, &j);
}

/* ---- XXH64 (published algorithm) ------------------------------------------------------------ */
#define XP1 11400714785074694791ULL
#define XP2 14029467366897019727ULL
#define XP3 1609587929392839161ULL
#define XP4 9650029242287828579ULL
#define XP5 2870177450012600261ULL
static inline uint64_t xrotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t xrd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t xrd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t in) {
  acc += in * XP2;
  return xrotl(acc, 31) * XP1;
}
uint64_t oracle_xxh64(const uint8_t* p, size_t len, uint64_t seed) {
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    for (; p + 32 <= end; p += 32) {
      v1 = xround(v1, xrd64(p));
      v2 = xround(v2, xrd64(p + 8));
      v3 = xround(v3, xrd64(p + 16));
      v4 = xround(v4, xrd64(p + 24));
    }
    h = xrotl(v1, 1) + xrotl(v2, 7) + xrotl(v3, 12) + xrotl(v4, 18);
    h = (h ^ xround(0, v1)) * XP1 + XP4;
    h = (h ^ xround(0, v2)) * XP1 + XP4;
    h = (h ^ xround(0, v3)) * XP1 + XP4;
    h = (h ^ xround(0, v4)) * XP1 + XP4;
  } else {
    h = seed + XP5;
  }
  h += (uint64_t)len;
  for (; p + 8 <= end; p += 8) h = xrotl(h ^ xround(0, xrd64(p)), 27) * XP1 + XP4;
  if (p + 4 <= end) { h = xrotl(h ^ ((uint64_t)xrd32(p) * XP1), 23) * XP2 + XP3; p += 4; }
  for (; p < end; ++p) h = xrotl(h ^ ((uint64_t)*p * XP5), 11) * XP1;
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

int oracle_chunk_keys(const int32_t* tokens, int64_t n, int chunk_tokens, uint64_t seed,
                      int include_partial, uint64_t* out) {
  const int64_t full = n / chunk_tokens, rem = n % chunk_tokens;
  int64_t k = 0;
  uint64_t prev = seed;
  for (; k < full; ++k)
    out[k] = prev = oracle_xxh64((const uint8_t*)(tokens + k * chunk_tokens), (size_t)chunk_tokens * 4, prev);
  if (include_partial && rem)
    { out[k] = prev = oracle_xxh64((const uint8_t*)(tokens + k * chunk_tokens), (size_t)rem * 4, prev); ++k; }
  return (int)k;
}
