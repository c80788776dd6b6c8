// dev_ops.cuh -- small device helpers shared by the interpreter kernel (pipeline.cu) and the runtime of the
// specialised (NVRTC-compiled) pipeline kernels (jit_rt.cuh): wrapping arithmetic, date/LIKE helpers, accumulator
// identities and combiners, 128-bit atomics.
#pragma once
#include "dev_util.cuh"
#include "vm.h"

namespace sg {

struct OpAdd { template <typename T> static __device__ __forceinline__ T f(T a, T b) { return a + b; } };
struct OpSub { template <typename T> static __device__ __forceinline__ T f(T a, T b) { return a - b; } };
struct OpMul { template <typename T> static __device__ __forceinline__ T f(T a, T b) { return a * b; } };
template <> __device__ __forceinline__ int32_t OpAdd::f<int32_t>(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
template <> __device__ __forceinline__ int32_t OpSub::f<int32_t>(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
template <> __device__ __forceinline__ int32_t OpMul::f<int32_t>(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
template <> __device__ __forceinline__ int64_t OpAdd::f<int64_t>(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
template <> __device__ __forceinline__ int64_t OpSub::f<int64_t>(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
template <> __device__ __forceinline__ int64_t OpMul::f<int64_t>(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
template <> __device__ __forceinline__ i128 OpAdd::f<i128>(i128 a, i128 b) { return (i128)((u128)a + (u128)b); }
template <> __device__ __forceinline__ i128 OpSub::f<i128>(i128 a, i128 b) { return (i128)((u128)a - (u128)b); }
template <> __device__ __forceinline__ i128 OpMul::f<i128>(i128 a, i128 b) { return (i128)((u128)a * (u128)b); }

struct CmpEq { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a == b; } };
struct CmpNe { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a != b; } };
struct CmpLt { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a < b; } };
struct CmpLe { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a <= b; } };
struct CmpGt { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a > b; } };
struct CmpGe { template <typename T> static __device__ __forceinline__ bool f(T a, T b) { return a >= b; } };

// days since 1970-01-01 -> civil (year, month, day)
__device__ __forceinline__ void civil_from_days(int32_t z0, int& y, int& m, int& d) {
  int64_t z = (int64_t)z0 + 719468;
  int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  int64_t doe = z - era * 146097;
  int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t yy = yoe + era * 400;
  int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  int64_t mp = (5 * doy + 2) / 153;
  d = (int)(doy - (153 * mp + 2) / 5 + 1);
  m = (int)(mp < 10 ? mp + 3 : mp - 9);
  y = (int)(m <= 2 ? yy + 1 : yy);
}

__device__ __forceinline__ bool like_match(const uint8_t* s, uint32_t n, const uint8_t* p, uint32_t m, int cls) {
  switch (cls) {
    case LIKE_EXACT:
      if (n != m) return false;
      for (uint32_t i = 0; i < m; ++i) if (s[i] != p[i]) return false;
      return true;
    case LIKE_PREFIX:
      if (n < m) return false;
      for (uint32_t i = 0; i < m; ++i) if (s[i] != p[i]) return false;
      return true;
    case LIKE_SUFFIX:
      if (n < m) return false;
      for (uint32_t i = 0; i < m; ++i) if (s[n - m + i] != p[i]) return false;
      return true;
    case LIKE_CONTAINS:
      if (n < m) return false;
      for (uint32_t st = 0; st + m <= n; ++st) {
        uint32_t i = 0;
        while (i < m && s[st + i] == p[i]) ++i;
        if (i == m) return true;
      }
      return false;
    default: {
      // generic %/_ matcher with single backtrack point (pattern bytes: '%' any run, '_' one byte, '\\' escape)
      uint32_t si = 0, pi = 0, star_p = 0xFFFFFFFFu, star_s = 0;
      while (si < n) {
        if (pi < m && p[pi] == '\\' && pi + 1 < m && p[pi + 1] == s[si]) { pi += 2; ++si; }
        else if (pi < m && p[pi] != '%' && p[pi] != '\\' && (p[pi] == '_' || p[pi] == s[si])) { ++pi; ++si; }
        else if (pi < m && p[pi] == '%') { star_p = pi++; star_s = si; }
        else if (star_p != 0xFFFFFFFFu) { pi = star_p + 1; si = ++star_s; }
        else return false;
      }
      while (pi < m && p[pi] == '%') ++pi;
      return pi == m;
    }
  }
}

__device__ __forceinline__ uint64_t load_key_word(const uint8_t* p, int width) {
  switch (width) {
    case 1: return *p;
    case 4: return (uint64_t)(uint32_t)lds<int32_t>(p);   // zero-extended: equality domain only
    default: return lds<uint64_t>(p);
  }
}

struct KeyRegs { uint64_t w[MAX_KEY_WORDS]; };

enum : uint32_t { ST_EMPTY = 0, ST_LOCKED = 1, ST_READY = 2 };

struct AccVal { i128 i; double f; bool valid; };

__device__ __forceinline__ void atomic_add_i128(uint64_t* w, i128 v) {
  unsigned long long lo = (unsigned long long)(u128)v, hi = (unsigned long long)((u128)v >> 64);
  unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(w), lo);
  unsigned long long carry = (old + lo) < old ? 1ull : 0ull;
  if (hi + carry) atomicAdd(reinterpret_cast<unsigned long long*>(w + 1), hi + carry);
}
__device__ __forceinline__ void atomic_minmax_i128(uint64_t* w, i128 v, bool is_min) {
  u128 cur = ((u128)w[1] << 64) | w[0];
  for (;;) {
    i128 c = (i128)cur;
    if (is_min ? (c <= v) : (c >= v)) return;
    u128 prev = atomic_cas_128(w, cur, (u128)v);
    if (prev == cur) return;
    cur = prev;
  }
}
__device__ __forceinline__ void atomic_minmax_f64(uint64_t* w, double v, bool is_min) {
  unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(w);
  for (;;) {
    double c = __longlong_as_double((long long)cur);
    if (is_min ? (c <= v) : (c >= v)) return;
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(w), cur, (unsigned long long)__double_as_longlong(v));
    if (prev == cur) return;
    cur = prev;
  }
}

// identity of accumulator word `word_in_acc` (0 or 1)
__host__ __device__ inline uint64_t acc_identity(int op, int word_in_acc) {
  switch (op) {
    case ACC_MIN_I32: case ACC_MIN_I64: return 0x7FFFFFFFFFFFFFFFull;
    case ACC_MAX_I32: case ACC_MAX_I64: return 0x8000000000000000ull;
    case ACC_MIN_I128: return word_in_acc ? 0x7FFFFFFFFFFFFFFFull : 0xFFFFFFFFFFFFFFFFull;
    case ACC_MAX_I128: return word_in_acc ? 0x8000000000000000ull : 0ull;
    case ACC_MIN_F64: return 0x7FF0000000000000ull;   // +inf
    case ACC_MAX_F64: return 0xFFF0000000000000ull;   // -inf
    default: return 0ull;
  }
}
__host__ __device__ inline int acc_words_of(int op) {
  return (op == ACC_SUM_I128 || op == ACC_MIN_I128 || op == ACC_MAX_I128) ? 2 : 1;
}

// combine value into a (private or CTA-total) accumulator held in plain memory words
__device__ __forceinline__ void acc_combine_words(int op, uint64_t& w0, uint64_t& w1, uint64_t v0, uint64_t v1) {
  switch (op) {
    case ACC_SUM_I64: case ACC_COUNT: w0 += v0; break;
    case ACC_SUM_I128: { uint64_t s = w0 + v0; w1 += v1 + (s < w0 ? 1ull : 0ull); w0 = s; break; }
    case ACC_SUM_F64: w0 = (uint64_t)__double_as_longlong(__longlong_as_double((long long)w0) + __longlong_as_double((long long)v0)); break;
    case ACC_MIN_I32: case ACC_MIN_I64: if ((int64_t)v0 < (int64_t)w0) w0 = v0; break;
    case ACC_MAX_I32: case ACC_MAX_I64: if ((int64_t)v0 > (int64_t)w0) w0 = v0; break;
    case ACC_MIN_I128: { i128 a = (i128)(((u128)w1 << 64) | w0), b = (i128)(((u128)v1 << 64) | v0); if (b < a) { w0 = v0; w1 = v1; } break; }
    case ACC_MAX_I128: { i128 a = (i128)(((u128)w1 << 64) | w0), b = (i128)(((u128)v1 << 64) | v0); if (b > a) { w0 = v0; w1 = v1; } break; }
    case ACC_MIN_F64: if (__longlong_as_double((long long)v0) < __longlong_as_double((long long)w0)) w0 = v0; break;
    case ACC_MAX_F64: if (__longlong_as_double((long long)v0) > __longlong_as_double((long long)w0)) w0 = v0; break;
    default: break;
  }
}

__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  return v;
}

// values whose magnitude is below 2^55 may be summed 128 at a time in 64 bits without overflow
__device__ __forceinline__ bool fits55(i128 v) {
  const int64_t lo = (int64_t)v;
  return (i128)lo == v && ((uint64_t)(lo + (1ll << 55)) >> 56) == 0;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t fold32(uint64_t v) { return (uint32_t)v ^ (uint32_t)(v >> 32); }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// substr(view, start, count) in characters (UTF-8): the result is an inline view when it fits in 12 bytes, otherwise it
// points into the source string (a result longer than 12 bytes implies a long source, whose bytes the batch keeps alive)
__device__ __forceinline__ uint8_t view_byte(const ulonglong2& v, uint32_t i) {
  if ((uint32_t)v.x <= 12) return i < 4 ? (uint8_t)(v.x >> (32 + 8 * i)) : (uint8_t)(v.y >> (8 * (i - 4)));
  return reinterpret_cast<const uint8_t*>(v.y)[i];
}
__device__ __forceinline__ ulonglong2 view_substr(const ulonglong2& v, long long start, long long count) {
  const uint32_t n = (uint32_t)v.x;
  const long long c0 = start - 1;
  uint32_t b0 = n, b1 = n;
  long long ci = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if ((view_byte(v, i) & 0xC0) != 0x80) {          // first byte of a character
      if (ci == c0) b0 = i;
      if (count >= 0 && ci == c0 + count) { b1 = i; break; }
      ++ci;
    }
  }
  if (b0 > b1) b0 = b1;
  const uint32_t rl = b1 - b0;
  ulonglong2 r; r.x = rl; r.y = 0;
  if (rl <= 12) {
    for (uint32_t k = 0; k < rl; ++k) {
      const unsigned long long b = view_byte(v, b0 + k);
      if (k < 4) r.x |= b << (32 + 8 * k); else r.y |= b << (8 * (k - 4));
    }
  } else {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(v.y) + b0;
    for (uint32_t k = 0; k < 4; ++k) r.x |= (unsigned long long)p[k] << (32 + 8 * k);
    r.y = reinterpret_cast<unsigned long long>(p);
  }
  return r;
}

}  // namespace sg
