// h2d_pack.cpp -- the host loops of the packed ingest (h2d.cu), compiled by g++ with function multiversioning: the library is
// built on one machine and runs on another, so every loop exists as an AVX-512, an AVX2 and a baseline clone and glibc's ifunc
// picks at load time.  Plain streaming loops (range scan, narrowing stores) written so that the vectoriser takes them.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#define SG_MV __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define SG_MV
#endif

extern "C" {

// Decimal128 piece: min / max of the low words, and whether every high word is the sign extension of its low word
SG_MV void sg_scan_dec128(const int64_t* __restrict p, int64_t n, int64_t* mn_out, int64_t* mx_out, uint64_t* bad_out) {
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  uint64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t lo = p[2 * i], hi = p[2 * i + 1];
    bad |= (uint64_t)(hi ^ (lo >> 63));
    mn = lo < mn ? lo : mn;
    mx = lo > mx ? lo : mx;
  }
  *mn_out = mn; *mx_out = mx; *bad_out = bad;
}
SG_MV void sg_scan_i64(const int64_t* __restrict p, int64_t n, int64_t* mn_out, int64_t* mx_out) {
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  for (int64_t i = 0; i < n; ++i) { mn = p[i] < mn ? p[i] : mn; mx = p[i] > mx ? p[i] : mx; }
  *mn_out = mn; *mx_out = mx;
}
SG_MV void sg_scan_i32(const int32_t* __restrict p, int64_t n, int32_t* mn_out, int32_t* mx_out) {
  int32_t mn = INT32_MAX, mx = INT32_MIN;
  for (int64_t i = 0; i < n; ++i) { mn = p[i] < mn ? p[i] : mn; mx = p[i] > mx ? p[i] : mx; }
  *mn_out = mn; *mx_out = mx;
}
SG_MV uint32_t sg_scan_view_maxlen(const uint32_t* __restrict p, int64_t n) {
  uint32_t L = 0;
  for (int64_t i = 0; i < n; ++i) L = p[4 * i] > L ? p[4 * i] : L;
  return L;
}

// deltas against `base`, narrowed to w bytes; `stride` in elements (2 for the low words of Decimal128)
SG_MV void sg_pack_i64(uint8_t* __restrict out, const int64_t* __restrict vals, int64_t stride, int64_t n, int64_t base, int w) {
  const uint64_t b = (uint64_t)base;
  if (stride == 2) {
    switch (w) {
      case 1: for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)((uint64_t)vals[2 * i] - b); break;
      case 2: { uint16_t* o = reinterpret_cast<uint16_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint16_t)((uint64_t)vals[2 * i] - b); break; }
      case 4: { uint32_t* o = reinterpret_cast<uint32_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint32_t)((uint64_t)vals[2 * i] - b); break; }
      default: { uint64_t* o = reinterpret_cast<uint64_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint64_t)vals[2 * i] - b; }
    }
  } else {
    switch (w) {
      case 1: for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)((uint64_t)vals[i] - b); break;
      case 2: { uint16_t* o = reinterpret_cast<uint16_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint16_t)((uint64_t)vals[i] - b); break; }
      case 4: { uint32_t* o = reinterpret_cast<uint32_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint32_t)((uint64_t)vals[i] - b); break; }
      default: { uint64_t* o = reinterpret_cast<uint64_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint64_t)vals[i] - b; }
    }
  }
}
SG_MV void sg_pack_i32(uint8_t* __restrict out, const int32_t* __restrict vals, int64_t n, int32_t base, int w) {
  const uint32_t b = (uint32_t)base;
  if (w == 1) for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)((uint32_t)vals[i] - b);
  else { uint16_t* o = reinterpret_cast<uint16_t*>(out); for (int64_t i = 0; i < n; ++i) o[i] = (uint16_t)((uint32_t)vals[i] - b); }
}
// inline views -> [length byte][L bytes] rows
SG_MV void sg_pack_views(uint8_t* __restrict out, const uint8_t* __restrict views, int64_t n, uint32_t L) {
  if (L == 0) { for (int64_t i = 0; i < n; ++i) out[i] = views[16 * i]; return; }
  if (L == 1) { for (int64_t i = 0; i < n; ++i) { out[2 * i] = views[16 * i]; out[2 * i + 1] = views[16 * i + 4]; } return; }
  const size_t rb = 1 + L;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* v = views + 16 * i;
    uint8_t* o = out + rb * (size_t)i;
    o[0] = v[0];
    for (uint32_t k = 0; k < L; ++k) o[1 + k] = v[4 + k];
  }
}

}  // extern "C"
