// b200kv_hash.cpp — chunk keys for the KV pool (CPU only).
//
// Own implementation of the published XXH64 algorithm (Yann Collet, BSD spec) — the same
// function the reference's prefix router uses through the `xxhash` wheel
// (src/vllm_router/prefix/hashtrie.py:56-57).  tests/test_hash.py pins it against that wheel.
// Prefix-chained chunk keys replace LMCache's token-chunk hashing behind
// lookup/store/retrieve.
#include <cstring>

#include "b200kv.h"

namespace {

constexpr uint64_t P1 = 11400714785074694791ull;
constexpr uint64_t P2 = 14029467366897019727ull;
constexpr uint64_t P3 = 1609587929392839161ull;
constexpr uint64_t P4 = 9650029242287828579ull;
constexpr uint64_t P5 = 2870177450012600261ull;

inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;  // little-endian host (x86-64 / aarch64-le)
}
inline uint32_t rd32(const uint8_t* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}
inline uint64_t lane_round(uint64_t acc, uint64_t in) {
  acc += in * P2;
  acc = rotl(acc, 31);
  return acc * P1;
}
inline uint64_t lane_merge(uint64_t h, uint64_t v) {
  h ^= lane_round(0, v);
  return h * P1 + P4;
}

}  // namespace

extern "C" uint64_t b200kv_xxh64(const void* data, size_t len, uint64_t seed) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  const uint8_t* const end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* const limit = end - 32;
    do {
      v1 = lane_round(v1, rd64(p));
      v2 = lane_round(v2, rd64(p + 8));
      v3 = lane_round(v3, rd64(p + 16));
      v4 = lane_round(v4, rd64(p + 24));
      p += 32;
    } while (p <= limit);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    h = lane_merge(h, v1);
    h = lane_merge(h, v2);
    h = lane_merge(h, v3);
    h = lane_merge(h, v4);
  } else {
    h = seed + P5;
  }
  h += static_cast<uint64_t>(len);
  while (p + 8 <= end) {
    h ^= lane_round(0, rd64(p));
    h = rotl(h, 27) * P1 + P4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= static_cast<uint64_t>(rd32(p)) * P1;
    h = rotl(h, 23) * P2 + P3;
    p += 4;
  }
  while (p < end) {
    h ^= static_cast<uint64_t>(*p) * P5;
    h = rotl(h, 11) * P1;
    ++p;
  }
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}

extern "C" int b200kv_chunk_keys(const int32_t* tokens, int64_t n_tokens, int32_t chunk_tokens,
                                 uint64_t seed, int include_partial, uint64_t* keys_out,
                                 int32_t* n_keys_out) {
  if (n_tokens < 0 || chunk_tokens <= 0 || !n_keys_out || (n_tokens > 0 && !tokens))
    return B200KV_EINVAL;
  const int64_t full = n_tokens / chunk_tokens;
  const int64_t rem = n_tokens % chunk_tokens;
  const int64_t n = full + ((include_partial && rem) ? 1 : 0);
  if (n > 0 && !keys_out) return B200KV_EINVAL;
  uint64_t prev = seed;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t cnt = (i < full) ? chunk_tokens : rem;
    prev = b200kv_xxh64(tokens + i * chunk_tokens, static_cast<size_t>(cnt) * 4, prev);
    keys_out[i] = prev;
  }
  *n_keys_out = static_cast<int32_t>(n);
  return B200KV_OK;
}
