// pipeline.cu -- the fused tile-pipeline kernel (see vm.h) and its helper kernels.
//
// Replaces, in one pass over HBM, the DataFusion operator chain
//   FilterExec -> ProjectionExec -> AggregateExec(Partial|Single|Final*)      
// that Sail drives per 8192-row batch (reference call sites: crates/sail-execution/src/
// job_runner.rs:64 `execute_stream`, crates/sail-physical-plan/src/streaming/filter.rs:104-116,
// crates/sail-plan/src/function/aggregate.rs:51-72,302-353,678-710).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "dev_ops.cuh"
#include "dev_util.cuh"
#include "kernels.hpp"
#include "vm.h"

namespace sg {

struct Smem {
  uint64_t full[2];        // TMA "tile landed" barriers, one per stage
  int32_t tile[2];
  int32_t hot_n;           // groups in the CTA-local hot dictionary
  int32_t elect;
  uint32_t cache_count;    // BUILD: occupied entries of the CTA chain cache
  int32_t defer[2];        // AGG with a bounded table: "stop taking tiles" flag, double buffered across iterations
  unsigned long long tile_base;   // COMPACT: exclusive prefix of this tile
  uint32_t warp_sums[33];
};
constexpr int SMEM_HDR = 256;

struct TileCtx {
  uint8_t* arena;
  uint32_t stage_off;      // added to stage-relative (bit 31) offsets
  int nrows;               // valid rows of this tile
  int64_t row0;
};
// offsets arrive pre-resolved per stage (KernelArgs): nothing to compute on the device
__device__ __forceinline__ uint32_t eff(const TileCtx&, uint32_t off) { return off; }

// ================================================================================================
// tile VM
// ================================================================================================
template <typename T> struct ImmOf;
template <> struct ImmOf<int32_t> { static __device__ __forceinline__ int32_t get(const VmInst& I) { return (int32_t)I.imm0; } };
template <> struct ImmOf<int64_t> { static __device__ __forceinline__ int64_t get(const VmInst& I) { return (int64_t)I.imm0; } };
template <> struct ImmOf<double> { static __device__ __forceinline__ double get(const VmInst& I) { return __longlong_as_double((long long)I.imm0); } };
template <> struct ImmOf<i128> { static __device__ __forceinline__ i128 get(const VmInst& I) { return (i128)(((u128)I.imm1 << 64) | I.imm0); } };
template <> struct ImmOf<uint8_t> { static __device__ __forceinline__ uint8_t get(const VmInst& I) { return (uint8_t)I.imm0; } };


// Operands of all RPT rows are loaded first, then computed, then stored: shared-memory loads of
// different rows may alias the stores as far as the compiler knows, so interleaving them would
// serialise the rows; batching exposes RPT independent dependency chains per instruction.
template <int RPT, typename T, bool IA, bool IB>
__device__ __forceinline__ void vm_load2(const VmInst& I, const TileCtx& c, T (&a)[RPT], T (&b)[RPT]) {
  const T imm = ImmOf<T>::get(I);
  const uint8_t* pa = c.arena + eff(c, I.a) + threadIdx.x * I.sa;
  const uint8_t* pb = c.arena + eff(c, I.b) + threadIdx.x * I.sb;
  const int sa = I.sa * NT, sb = I.sb * NT;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    a[k] = IA ? imm : lds<T>(pa + k * sa);
    b[k] = IB ? imm : lds<T>(pb + k * sb);
  }
}
template <int RPT, typename T, typename F, bool IA, bool IB>
__device__ __forceinline__ void vm_bin_i(const VmInst& I, const TileCtx& c) {
  T a[RPT], b[RPT];
  vm_load2<RPT, T, IA, IB>(I, c, a, b);
  uint8_t* pd = c.arena + eff(c, I.dst) + threadIdx.x * (int)sizeof(T);
#pragma unroll
  for (int k = 0; k < RPT; ++k) sts<T>(pd + k * NT * (int)sizeof(T), F::template f<T>(a[k], b[k]));
}
template <int RPT, typename T, typename F>
__device__ __forceinline__ void vm_bin(const VmInst& I, const TileCtx& c) {
  if (I.flags & F_IMM_B) vm_bin_i<RPT, T, F, false, true>(I, c);
  else if (I.flags & F_IMM_A) vm_bin_i<RPT, T, F, true, false>(I, c);
  else vm_bin_i<RPT, T, F, false, false>(I, c);
}
template <int RPT, typename T, typename F, bool IA, bool IB>
__device__ __forceinline__ void vm_cmp_i(const VmInst& I, const TileCtx& c) {
  T a[RPT], b[RPT];
  vm_load2<RPT, T, IA, IB>(I, c, a, b);
  uint8_t* pd = c.arena + eff(c, I.dst) + threadIdx.x;
#pragma unroll
  for (int k = 0; k < RPT; ++k) pd[k * NT] = F::template f<T>(a[k], b[k]) ? 1 : 0;
}
template <int RPT, typename T, typename F>
__device__ __forceinline__ void vm_cmp(const VmInst& I, const TileCtx& c) {
  if (I.flags & F_IMM_B) vm_cmp_i<RPT, T, F, false, true>(I, c);
  else if (I.flags & F_IMM_A) vm_cmp_i<RPT, T, F, true, false>(I, c);
  else vm_cmp_i<RPT, T, F, false, false>(I, c);
}
template <int RPT, typename F>
__device__ __forceinline__ void vm_bin_kind(const VmInst& I, int kind, const TileCtx& c) {
  switch (kind) {
    case K_I32: vm_bin<RPT, int32_t, F>(I, c); break;
    case K_I64: vm_bin<RPT, int64_t, F>(I, c); break;
    case K_F64: vm_bin<RPT, double, F>(I, c); break;
    case K_I128: vm_bin<RPT, i128, F>(I, c); break;
    default: break;
  }
}
template <int RPT, typename F>
__device__ __forceinline__ void vm_cmp_kind(const VmInst& I, int kind, const TileCtx& c) {
  switch (kind) {
    case K_B: vm_cmp<RPT, uint8_t, F>(I, c); break;
    case K_I32: vm_cmp<RPT, int32_t, F>(I, c); break;
    case K_I64: vm_cmp<RPT, int64_t, F>(I, c); break;
    case K_F64: vm_cmp<RPT, double, F>(I, c); break;
    case K_I128: vm_cmp<RPT, i128, F>(I, c); break;
    default: break;
  }
}

// views: equality only (ordering comparisons on strings are rejected by the compiler)
template <int RPT>
__device__ __noinline__ void vm_view_eq(const VmInst& I, const TileCtx& c, bool negate) {
  const uint8_t* pa = c.arena + eff(c, I.a);
  const uint8_t* pb = c.arena + eff(c, I.b);
  uint8_t* pd = c.arena + eff(c, I.dst);
  ulonglong2 imm; imm.x = I.imm0; imm.y = I.imm1;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    ulonglong2 a = (I.flags & F_IMM_A) ? imm : *reinterpret_cast<const ulonglong2*>(pa + r * I.sa);
    ulonglong2 b = (I.flags & F_IMM_B) ? imm : *reinterpret_cast<const ulonglong2*>(pb + r * I.sb);
    bool eq = view_equal(a, b);
    pd[r] = (eq != negate) ? 1 : 0;
  }
}

template <typename T> __device__ __forceinline__ T trunc_div(T a, T b) { return a / b; }

template <int RPT, typename T, bool REM>
__device__ __noinline__ void vm_div(const VmInst& I, const TileCtx& c, uint32_t* err) {
  const T imm = ImmOf<T>::get(I);
  const uint8_t* pa = c.arena + eff(c, I.a);
  const uint8_t* pb = c.arena + eff(c, I.b);
  const uint8_t* pg = I.c == NO_SLOT ? nullptr : c.arena + eff(c, I.c);
  uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    T a = (I.flags & F_IMM_A) ? imm : lds<T>(pa + r * I.sa);
    T b = (I.flags & F_IMM_B) ? imm : lds<T>(pb + r * I.sb);
    bool live = r < c.nrows && (pg == nullptr || pg[r]);
    T q = 0;
    if (b == 0) {
      if (live) atomicOr(err, ERR_DIV_ZERO);
    } else {
      q = REM ? (T)(a % b) : (T)(a / b);
    }
    sts<T>(pd + r * (int)sizeof(T), q);
  }
}
template <int RPT>
__device__ __noinline__ void vm_div_f64(const VmInst& I, const TileCtx& c, bool rem) {
  const double imm = ImmOf<double>::get(I);
  const uint8_t* pa = c.arena + eff(c, I.a);
  const uint8_t* pb = c.arena + eff(c, I.b);
  uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    double a = (I.flags & F_IMM_A) ? imm : lds<double>(pa + r * I.sa);
    double b = (I.flags & F_IMM_B) ? imm : lds<double>(pb + r * I.sb);
    sts<double>(pd + r * 8, rem ? fmod(a, b) : a / b);
  }
}

// decimal rescale down: round half away from zero (arrow `rescale_decimal`)
template <int RPT, typename T>
__device__ __noinline__ void vm_divround(const VmInst& I, const TileCtx& c) {
  const T d = ImmOf<T>::get(I);
  const uint8_t* pa = c.arena + eff(c, I.a);
  uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    T a = lds<T>(pa + r * I.sa);
    T q = a / d, rem = a % d;
    T twice = rem < 0 ? -rem * 2 : rem * 2;
    if (twice >= d) q += (a < 0 ? -1 : 1);
    sts<T>(pd + r * (int)sizeof(T), q);
  }
}


template <int RPT>
__device__ __noinline__ void vm_cvt(const VmInst& I, int dkind, const TileCtx& c) {
  const uint8_t* pa = c.arena + eff(c, I.a);
  uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    const uint8_t* p = pa + r * I.sa;
    i128 iv = 0; double fv = 0.0; bool isf = false;
    switch (I.aux) {
      case SRC_I8: iv = lds<int8_t>(p); break;
      case SRC_I16: iv = lds<int16_t>(p); break;
      case SRC_U8: iv = lds<uint8_t>(p); break;
      case SRC_U16: iv = lds<uint16_t>(p); break;
      case SRC_U32: iv = lds<uint32_t>(p); break;
      case SRC_I32: iv = lds<int32_t>(p); break;
      case SRC_I64: iv = lds<int64_t>(p); break;
      case SRC_I128: iv = lds<i128>(p); break;
      case SRC_B: iv = lds<uint8_t>(p) ? 1 : 0; break;
      case SRC_F32: fv = lds<float>(p); isf = true; break;
      case SRC_F64: fv = lds<double>(p); isf = true; break;
    }
    switch (dkind) {
      case K_I32: sts<int32_t>(pd + r * 4, isf ? (int32_t)fv : (int32_t)iv); break;
      case K_I64: sts<int64_t>(pd + r * 8, isf ? (int64_t)fv : (int64_t)iv); break;
      case K_I128: sts<i128>(pd + r * 16, isf ? (i128)(int64_t)fv : iv); break;
      case K_F64: sts<double>(pd + r * 8, isf ? fv : (I.aux == SRC_I128 ? (double)iv : (double)(int64_t)iv)); break;
      case K_B: pd[r] = isf ? (fv != 0.0) : (iv != 0); break;
    }
  }
}

template <int RPT, typename T>
__device__ __forceinline__ void vm_select(const VmInst& I, const TileCtx& c) {
  const T imm = ImmOf<T>::get(I);
  const uint8_t* pa = c.arena + eff(c, I.a);
  const uint8_t* pb = c.arena + eff(c, I.b);
  const uint8_t* pc = c.arena + eff(c, I.c);
  uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    T a = (I.flags & F_IMM_A) ? imm : lds<T>(pa + r * I.sa);
    T b = (I.flags & F_IMM_B) ? imm : lds<T>(pb + r * I.sb);
    sts<T>(pd + r * (int)sizeof(T), pc[r] ? a : b);
  }
}
template <int RPT>
__device__ __noinline__ void vm_select_v16(const VmInst& I, const TileCtx& c) {
  const uint8_t* pa = c.arena + eff(c, I.a);
  const uint8_t* pb = c.arena + eff(c, I.b);
  const uint8_t* pc = c.arena + eff(c, I.c);
  uint8_t* pd = c.arena + eff(c, I.dst);
  ulonglong2 imm; imm.x = I.imm0; imm.y = I.imm1;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    ulonglong2 a = (I.flags & F_IMM_A) ? imm : *reinterpret_cast<const ulonglong2*>(pa + r * I.sa);
    ulonglong2 b = (I.flags & F_IMM_B) ? imm : *reinterpret_cast<const ulonglong2*>(pb + r * I.sb);
    *reinterpret_cast<ulonglong2*>(pd + r * 16) = pc[r] ? a : b;
  }
}



template <int RPT>
__device__ __noinline__ void vm_like(const VmInst& I, const TileCtx& c) {
  const uint8_t* pa = c.arena + eff(c, I.a);
  uint8_t* pd = c.arena + eff(c, I.dst);
  const uint8_t* pat = reinterpret_cast<const uint8_t*>(I.imm1);
  const uint32_t plen = (uint32_t)I.imm0;
  const int cls = I.aux & 0xFF;
  const bool negate = (I.aux >> 8) & 1;
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    bool hit = false;
    if (r < c.nrows) {
      const uint8_t* vp = pa + r * I.sa;
      ulonglong2 v = *reinterpret_cast<const ulonglong2*>(vp);
      hit = like_match(view_ptr(v, vp), (uint32_t)v.x, pat, plen, cls);
    }
    pd[r] = (hit != negate) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// join probe (OP_PROBE): open-addressing lookup of the probe key in the build table
// ------------------------------------------------------------------------------------------------

// Packs the key columns of row r into 8-byte words; returns the combined hash.  *has_null is set
// when any key is NULL.  Views are kept as their 16 raw bytes (view_equal/view_hash resolve long
// strings through the absolute pointer stored by the importer).
template <int NKMAX>
__device__ __forceinline__ uint64_t pack_key(const KeyDesc* keys, int n_keys, int has_null_word, const TileCtx& c, int r,
                                             KeyRegs& k, bool* has_null) {
  uint64_t h = 0x243F6A8885A308D3ull;
  uint64_t nullmask = 0;
  int w = has_null_word ? 1 : 0;
#pragma unroll
  for (int i = 0; i < NKMAX; ++i) {
    if (i < n_keys) {
      const KeyDesc& d = keys[i];
      const uint8_t* p = c.arena + eff(c, d.slot) + r * d.stride;
      bool isnull = d.valid_slot != NO_SLOT && c.arena[eff(c, d.valid_slot) + r] == 0;
      if (isnull) nullmask |= 1ull << i;
      if (d.width == 16) {
        ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
        if (isnull) { v.x = 0; v.y = 0; }
        k.w[w] = v.x; k.w[w + 1] = v.y;
        h = mix64(h ^ (d.is_view ? view_hash(v) : mix64(v.x ^ mix64(v.y))));
        w += 2;
      } else {
        uint64_t v = isnull ? 0 : load_key_word(p, d.width);
        k.w[w] = v;
        h = mix64(h ^ v);
        w += 1;
      }
    }
  }
  if (has_null_word) { k.w[0] = nullmask; h = mix64(h ^ nullmask); }
  *has_null = nullmask != 0;
  return h;
}

__device__ __forceinline__ bool key_words_equal(const KeyDesc* keys, int n_keys, int has_null_word, const uint64_t* a,
                                                const KeyRegs& b) {
  int w = 0;
  if (has_null_word) { if (a[0] != b.w[0]) return false; w = 1; }
  for (int i = 0; i < n_keys; ++i) {
    if (keys[i].width == 16) {
      if (keys[i].is_view) {
        ulonglong2 x, y; x.x = a[w]; x.y = a[w + 1]; y.x = b.w[w]; y.y = b.w[w + 1];
        if (!view_equal(x, y)) return false;
      } else if (a[w] != b.w[w] || a[w + 1] != b.w[w + 1]) return false;
      w += 2;
    } else {
      if (a[w] != b.w[w]) return false;
      w += 1;
    }
  }
  return true;
}

// join table slot: { u64 tag (0 = empty; hash | 1), i64 row }
// General lookup: any number / type of keys; the rows of a thread are resolved one after the other.
template <int RPT>
__device__ __noinline__ void vm_probe_general(const ProbeParams& P, const TileCtx& c, uint32_t mask_slot) {
  uint8_t* pm = c.arena + eff(c, P.match_slot);
  uint8_t* prow = c.arena + eff(c, P.rowid_slot);
  const uint8_t* pact = mask_slot == NO_SLOT ? nullptr : c.arena + eff(c, mask_slot);
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    int64_t row = -1;
    bool live = r < c.nrows && (pact == nullptr || pact[r]);
    if (live) {
      KeyRegs key; bool has_null;
      uint64_t h = pack_key<MAX_KEYS>(P.keys, P.n_keys, 0, c, r, key, &has_null);
      if (!has_null) {   // NullEqualsNothing
        uint64_t tag = h | 1ull;
        uint64_t idx = (h >> 1) & P.capacity_mask;
        for (;;) {
          const ulonglong2 s = *reinterpret_cast<const ulonglong2*>(P.table + idx * 16);
          if (s.x == 0) break;
          if (s.x == tag) {
            // verify against the build-side key columns
            const int64_t cand = (int64_t)s.y - 1;                 // slot head = row + 1
            bool eq = true;
            int w = 0;
            for (int i = 0; i < P.n_keys && eq; ++i) {
              const KeyDesc& d = P.keys[i];
              const uint8_t* bp = P.build_keys[i] + cand * P.build_stride[i];
              if (d.width == 16) {
                ulonglong2 bv = *reinterpret_cast<const ulonglong2*>(bp);
                if (d.is_view) { ulonglong2 pv; pv.x = key.w[w]; pv.y = key.w[w + 1]; eq = view_equal(bv, pv); }
                else eq = bv.x == key.w[w] && bv.y == key.w[w + 1];
                w += 2;
              } else {
                eq = load_key_word(bp, d.width) == key.w[w];
                w += 1;
              }
            }
            if (eq) { row = cand; break; }
          }
          idx = (idx + 1) & P.capacity_mask;
        }
      }
    }
    if (row >= 0 && P.visited) P.visited[row] = 1;
    pm[r] = row >= 0 ? 1 : 0;
    sts<int64_t>(prow + r * 8, row);
  }
}

// One key of at most 8 bytes (the usual integer / date join key).  The lookup is latency bound (table slot, then the
// build key of the candidate: two dependent random loads), so the RPT rows of a thread advance in lock step: all
// first-slot loads are issued together, then all candidate key loads; only rows that collide continue probing.
// Hash as pack_key(): mix64(seed ^ key word).
template <int RPT>
__device__ __noinline__ void vm_probe_narrow(const ProbeParams& P, const TileCtx& c, uint32_t mask_slot) {
  uint8_t* pm = c.arena + eff(c, P.match_slot);
  uint8_t* prow = c.arena + eff(c, P.rowid_slot);
  const uint8_t* pact = mask_slot == NO_SLOT ? nullptr : c.arena + eff(c, mask_slot);
  const KeyDesc d = P.keys[0];
  const uint8_t* pk = c.arena + eff(c, d.slot);
  const uint8_t* pv = d.valid_slot == NO_SLOT ? nullptr : c.arena + eff(c, d.valid_slot);
  const uint8_t* bcol = P.build_keys[0];
  const int bstride = P.build_stride[0];
  uint64_t key[RPT], tag[RPT], bkey[RPT];
  ulonglong2 s[RPT];
  bool act[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    act[k] = r < c.nrows && (pact == nullptr || pact[r]) && (pv == nullptr || pv[r]);     // NULL keys match nothing
    key[k] = load_key_word(pk + r * d.stride, d.width);
    const uint64_t h = mix64(0x243F6A8885A308D3ull ^ key[k]);
    tag[k] = h | 1ull;
    s[k].x = 0; s[k].y = 0;
    if (act[k]) s[k] = *reinterpret_cast<const ulonglong2*>(P.table + ((h >> 1) & P.capacity_mask) * 16);
  }
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    bkey[k] = 0;
    if (s[k].x == tag[k]) bkey[k] = load_key_word(bcol + ((int64_t)s[k].y - 1) * bstride, d.width);      // slot head = row + 1
  }
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    int64_t row = -1;
    if (s[k].x != 0) {
      if (s[k].x == tag[k] && bkey[k] == key[k]) row = (int64_t)s[k].y - 1;
      else {                                                  // collision: ordinary linear probe from the next slot
        uint64_t idx = (((tag[k] >> 1) & P.capacity_mask) + 1) & P.capacity_mask;
        for (;;) {
          const ulonglong2 cur = *reinterpret_cast<const ulonglong2*>(P.table + idx * 16);
          if (cur.x == 0) break;
          if (cur.x == tag[k] && load_key_word(bcol + ((int64_t)cur.y - 1) * bstride, d.width) == key[k]) { row = (int64_t)cur.y - 1; break; }
          idx = (idx + 1) & P.capacity_mask;
        }
      }
    }
    if (row >= 0 && P.visited) P.visited[row] = 1;
    pm[r] = row >= 0 ? 1 : 0;
    sts<int64_t>(prow + r * 8, row);
  }
}

template <int RPT>
__device__ __forceinline__ void vm_probe(const ProbeParams& P, const TileCtx& c, uint32_t mask_slot) {
  if (P.n_keys == 1 && P.keys[0].width != 16) vm_probe_narrow<RPT>(P, c, mask_slot);
  else vm_probe_general<RPT>(P, c, mask_slot);
}

template <int RPT>
__device__ __noinline__ void vm_gather(const VmInst& I, const TileCtx& c) {
  const uint8_t* prow = c.arena + eff(c, I.a);
  uint8_t* pd = c.arena + eff(c, I.dst);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(I.imm1);
  const int w = I.aux;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    int64_t row = lds<int64_t>(prow + r * 8);
    if (w == 16) {
      ulonglong2 v; v.x = 0; v.y = 0;
      if (row >= 0) v = *reinterpret_cast<const ulonglong2*>(src + row * 16);
      *reinterpret_cast<ulonglong2*>(pd + r * 16) = v;
    } else if (w == 8) {
      sts<uint64_t>(pd + r * 8, row >= 0 ? *reinterpret_cast<const uint64_t*>(src + row * 8) : 0ull);
    } else if (w == 4) {
      sts<uint32_t>(pd + r * 4, row >= 0 ? *reinterpret_cast<const uint32_t*>(src + row * 4) : 0u);
    } else {   // bytes (validity / booleans stored one byte per row on the build side)
      pd[r] = row >= 0 ? src[row] : 0;
    }
  }
}

template <int RPT>
__device__ __forceinline__ void vm_exec(const VmInst* prog, int n_inst, const TileCtx& c, const PipelineParams& P,
                                        const PipelineAux* aux) {
  for (int pc = 0; pc < n_inst; ++pc) {
    const VmInst& I = prog[pc];      // stays in shared memory: fields are read with uniform LDS
    const int base = I.op & 0xFF, kind = I.op >> 8;
    switch (base) {
      case OP_UNPACK_BITS: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          pd[r] = (pa[r >> 3] >> (r & 7)) & 1;
        }
        break;
      }
      case OP_CONST: {
        uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          switch (kind) {
            case K_B: pd[r] = (uint8_t)I.imm0; break;
            case K_I32: sts<int32_t>(pd + r * 4, (int32_t)I.imm0); break;
            case K_I64: case K_F64: sts<uint64_t>(pd + r * 8, I.imm0); break;
            default: { ulonglong2 v; v.x = I.imm0; v.y = I.imm1; *reinterpret_cast<ulonglong2*>(pd + r * 16) = v; }
          }
        }
        break;
      }
      case OP_MOV: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
        const int w = kind_width(kind);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          if (w == 16) *reinterpret_cast<ulonglong2*>(pd + r * 16) = *reinterpret_cast<const ulonglong2*>(pa + r * I.sa);
          else if (w == 8) sts<uint64_t>(pd + r * 8, lds<uint64_t>(pa + r * I.sa));
          else if (w == 4) sts<uint32_t>(pd + r * 4, lds<uint32_t>(pa + r * I.sa));
          else pd[r] = pa[r * I.sa];
        }
        break;
      }
      case OP_CVT: vm_cvt<RPT>(I, kind, c); break;
      case OP_ADD: vm_bin_kind<RPT, OpAdd>(I, kind, c); break;
      case OP_SUB: vm_bin_kind<RPT, OpSub>(I, kind, c); break;
      case OP_MUL: vm_bin_kind<RPT, OpMul>(I, kind, c); break;
      case OP_DIV:
      case OP_REM:
        if (kind == K_F64) vm_div_f64<RPT>(I, c, base == OP_REM);
        else if (kind == K_I32) { if (base == OP_DIV) vm_div<RPT, int32_t, false>(I, c, P.error_flag); else vm_div<RPT, int32_t, true>(I, c, P.error_flag); }
        else if (kind == K_I64) { if (base == OP_DIV) vm_div<RPT, int64_t, false>(I, c, P.error_flag); else vm_div<RPT, int64_t, true>(I, c, P.error_flag); }
        else { if (base == OP_DIV) vm_div<RPT, i128, false>(I, c, P.error_flag); else vm_div<RPT, i128, true>(I, c, P.error_flag); }
        break;
      case OP_NEG: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          switch (kind) {
            case K_I32: sts<int32_t>(pd + r * 4, (int32_t)(0u - (uint32_t)lds<int32_t>(pa + r * I.sa))); break;
            case K_I64: sts<int64_t>(pd + r * 8, (int64_t)(0ull - (uint64_t)lds<int64_t>(pa + r * I.sa))); break;
            case K_F64: sts<double>(pd + r * 8, -lds<double>(pa + r * I.sa)); break;
            default: sts<i128>(pd + r * 16, (i128)((u128)0 - (u128)lds<i128>(pa + r * I.sa)));
          }
        }
        break;
      }
      case OP_MULW: {
        int64_t a[RPT], b[RPT];
        if (I.flags & F_IMM_B) vm_load2<RPT, int64_t, false, true>(I, c, a, b);
        else if (I.flags & F_IMM_A) vm_load2<RPT, int64_t, true, false>(I, c, a, b);
        else vm_load2<RPT, int64_t, false, false>(I, c, a, b);
        uint8_t* pd = c.arena + eff(c, I.dst) + threadIdx.x * 16;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          ulonglong2 w;
          w.x = (unsigned long long)a[k] * (unsigned long long)b[k];
          w.y = (unsigned long long)__mul64hi((long long)a[k], (long long)b[k]);
          *reinterpret_cast<ulonglong2*>(pd + k * NT * 16) = w;
        }
        break;
      }
      case OP_MUL128_64: {
        const int64_t imm = (int64_t)I.imm0;
        const uint8_t* pa = c.arena + eff(c, I.a) + threadIdx.x * I.sa;
        const uint8_t* pb = c.arena + eff(c, I.b) + threadIdx.x * I.sb;
        uint8_t* pd = c.arena + eff(c, I.dst) + threadIdx.x * 16;
        ulonglong2 a[RPT]; int64_t b[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          a[k] = *reinterpret_cast<const ulonglong2*>(pa + k * NT * I.sa);
          b[k] = (I.flags & F_IMM_B) ? imm : lds<int64_t>(pb + k * NT * I.sb);
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          // (ahi:alo) * sext(b)  mod 2^128
          const unsigned long long ub = (unsigned long long)b[k];
          ulonglong2 w;
          w.x = a[k].x * ub;
          w.y = __umul64hi(a[k].x, ub) + a[k].y * ub - (b[k] < 0 ? a[k].x : 0ull);
          *reinterpret_cast<ulonglong2*>(pd + k * NT * 16) = w;
        }
        break;
      }
      case OP_DIVROUND:
        if (kind == K_I64) vm_divround<RPT, int64_t>(I, c); else vm_divround<RPT, i128>(I, c);
        break;
      case OP_EQ: if (kind == K_V16) vm_view_eq<RPT>(I, c, false); else vm_cmp_kind<RPT, CmpEq>(I, kind, c); break;
      case OP_NE: if (kind == K_V16) vm_view_eq<RPT>(I, c, true); else vm_cmp_kind<RPT, CmpNe>(I, kind, c); break;
      case OP_LT: vm_cmp_kind<RPT, CmpLt>(I, kind, c); break;
      case OP_LE: vm_cmp_kind<RPT, CmpLe>(I, kind, c); break;
      case OP_GT: vm_cmp_kind<RPT, CmpGt>(I, kind, c); break;
      case OP_GE: vm_cmp_kind<RPT, CmpGe>(I, kind, c); break;
      case OP_AND: case OP_OR: case OP_ANDNOT: case OP_NOT: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        const uint8_t* pb = c.arena + eff(c, I.b);
        uint8_t* pd = c.arena + eff(c, I.dst);
        const uint8_t imm = (uint8_t)I.imm0;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          uint8_t a = (I.flags & F_IMM_A) ? imm : pa[r];
          uint8_t b = (base == OP_NOT) ? 0 : ((I.flags & F_IMM_B) ? imm : pb[r]);
          uint8_t v = base == OP_AND ? (a & b) : base == OP_OR ? (a | b) : base == OP_ANDNOT ? (a & (b ^ 1)) : (a ^ 1);
          pd[r] = v & 1;
        }
        break;
      }
      case OP_SELECT:
        switch (kind) {
          case K_B: vm_select<RPT, uint8_t>(I, c); break;
          case K_I32: vm_select<RPT, int32_t>(I, c); break;
          case K_I64: vm_select<RPT, int64_t>(I, c); break;
          case K_F64: vm_select<RPT, double>(I, c); break;
          case K_I128: vm_select<RPT, i128>(I, c); break;
          default: vm_select_v16<RPT>(I, c);
        }
        break;
      case OP_STR_EQ_LONG: {
        // literal longer than 12 bytes: imm0 = len | prefix << 32, imm1 = device pointer to the bytes
        ulonglong2 lit; lit.x = I.imm0; lit.y = I.imm1;
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          ulonglong2 a = *reinterpret_cast<const ulonglong2*>(pa + r * I.sa);
          bool eq = r < c.nrows && view_equal(a, lit);
          pd[r] = (eq != (bool)(I.aux & 1)) ? 1 : 0;
        }
        break;
      }
      case OP_STR_LIKE: vm_like<RPT>(I, c); break;
      case OP_DATE_PART: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          int y, m, d;
          civil_from_days(lds<int32_t>(pa + r * I.sa), y, m, d);
          sts<int32_t>(pd + r * 4, I.aux == 0 ? y : I.aux == 1 ? m : d);
        }
        break;
      }
      case OP_SUBSTR: {
        const uint8_t* pa = c.arena + eff(c, I.a);
        uint8_t* pd = c.arena + eff(c, I.dst);
        for (int k = 0; k < RPT; ++k) {
          const int r = threadIdx.x + k * NT;
          ulonglong2 v; v.x = 0; v.y = 0;
          if (r < c.nrows) v = view_substr(*reinterpret_cast<const ulonglong2*>(pa + r * I.sa), (long long)I.imm0, (long long)I.imm1);
          *reinterpret_cast<ulonglong2*>(pd + r * 16) = v;
        }
        break;
      }
      case OP_PROBE: vm_probe<RPT>(aux->probe[I.aux], c, I.c); break;
      case OP_GATHER: vm_gather<RPT>(I, c); break;
      default: break;
    }
  }
}

// ================================================================================================
// tile loading
// ================================================================================================
__device__ __forceinline__ uint32_t tile_bytes_of(const InputCol& in, int tile_rows) {
  return in.width ? (uint32_t)in.width * tile_rows : (uint32_t)(tile_rows >> 3);
}

// cooperative copy (partial / unaligned tiles): zero-fills rows >= nrows
__device__ __forceinline__ void load_tile_generic(const PipelineParams& P, uint8_t* arena, int64_t row0, int nrows) {
  for (int i = 0; i < P.n_inputs; ++i) {
    const InputCol& in = P.in[i];
    uint8_t* dst = arena + in.slot;
    const uint32_t total = tile_bytes_of(in, P.tile_rows);
    const uint32_t valid = in.width ? (uint32_t)in.width * nrows : (uint32_t)((nrows + 7) >> 3);
    const uint8_t* src = in.data + (in.width ? (int64_t)in.width * row0 : (row0 >> 3));
    if (in.tma_ok && (valid & 15u) == 0) {
      for (uint32_t o = threadIdx.x * 16; o < total; o += NT * 16) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (o < valid) v = *reinterpret_cast<const uint4*>(src + o);
        *reinterpret_cast<uint4*>(dst + o) = v;
      }
    } else {
      for (uint32_t o = threadIdx.x; o < total; o += NT) dst[o] = o < valid ? src[o] : 0;
    }
  }
}

// ================================================================================================
// aggregation: global table
// ================================================================================================


// direct-key protocol (AggParams::direct_key): the key word itself is the slot's lock
__device__ __forceinline__ uint64_t* agg_find_or_insert_direct(const AggParams& A, unsigned long long k, uint64_t h, uint32_t* err) {
  if (k == DIRECT_EMPTY_KEY) {                       // the sentinel as a real key: the entry behind the table
    uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + (A.capacity_mask + 1) * A.entry_words;
    if (*reinterpret_cast<volatile unsigned long long*>(e) == 0ull && atomicCAS(reinterpret_cast<unsigned long long*>(e), 0ull, h | 1ull) == 0ull) atomicAdd(A.n_groups, 1ull);
    return e;
  }
  uint64_t idx = h & A.capacity_mask;
  for (uint64_t probes = 0; probes <= A.capacity_mask; ++probes) {
    uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + idx * A.entry_words;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(e + 2);
    if (cur == DIRECT_EMPTY_KEY) {
      cur = atomicCAS(reinterpret_cast<unsigned long long*>(e + 2), DIRECT_EMPTY_KEY, k);
      if (cur == DIRECT_EMPTY_KEY) {
        e[0] = h | 1ull;
        const unsigned m = __activemask();           // one counter update for the lanes that insert in the same step
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(A.n_groups, (unsigned long long)__popc(m));
        return e;
      }
    }
    if (cur == k) return e;
    idx = (idx + 1) & A.capacity_mask;
  }
  atomicOr(err, ERR_TABLE_FULL);
  return nullptr;
}

// state word: 0 = empty, else (tag30 << 2) | {1 = being written, 2 = ready}.  Carrying the tag in the
// state lets a thread skip a slot that is being written for a different key without waiting on it.
__device__ __forceinline__ uint64_t* agg_find_or_insert(const AggParams& A, const KeyRegs& key, uint64_t h, uint32_t* err) {
  if (A.direct_key) return agg_find_or_insert_direct(A, key.w[0], h, err);
  const uint32_t tag = (uint32_t)(h >> 34) << 2;
  uint64_t idx = h & A.capacity_mask;
  uint64_t probes = 0;
  uint32_t spins = 0;
  while (probes <= A.capacity_mask) {
    uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + idx * A.entry_words;
    uint32_t* st = A.state + idx;
    uint32_t s = ld_acquire_u32(st);
    if (s == ST_EMPTY) {
      s = atomicCAS(st, ST_EMPTY, tag | ST_LOCKED);
      if (s == ST_EMPTY) {
        e[0] = h;
        e[1] = 0;                                                   // seen
        for (int w = 0; w < A.key_words; ++w) e[2 + w] = key.w[w];
        for (int j = 0; j < A.n_accs; ++j)
          for (int w = 0; w < acc_words_of(A.accs[j].op); ++w)
            e[2 + A.key_words + A.accs[j].word + w] = acc_identity(A.accs[j].op, w);
        st_release_u32(st, tag | ST_READY);                         // release: the entry words above are visible first
        {
          // occupied-slot list: one counter for the whole table, so the increment is aggregated over the lanes
          // that insert in the same step (same-address atomics serialise in one L2 slice)
          const unsigned m = __activemask();
          const unsigned lane = threadIdx.x & 31;
          const int lead = __ffs(m) - 1;
          unsigned long long base = 0;
          if ((int)lane == lead) base = atomicAdd(A.n_groups, (unsigned long long)__popc(m));
          base = __shfl_sync(m, base, lead);
          A.occ[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)idx;
        }
        return e;
      }
    }
    if ((s & ~3u) == tag) {
      if ((s & 3u) == ST_LOCKED) {          // same tag, still being published: look again (bounded)
        if (++spins > (1u << 24)) { atomicOr(err, ERR_TABLE_FULL); return nullptr; }
        __nanosleep(32);
        continue;
      }
      if (e[0] == h && key_words_equal(A.keys, A.n_keys, A.has_null_word, e + 2, key)) return e;
    }
    idx = (idx + 1) & A.capacity_mask;
    ++probes;
  }
  atomicOr(err, ERR_TABLE_FULL);
  return nullptr;
}

// Warp-cooperative front end: lanes of one warp that carry the same key hash elect a leader
// (__match_any_sync); only leaders touch the table, followers receive the entry pointer by shuffle.
// Removes same-warp contention on a slot that is being published.  All 32 lanes must call.
__device__ __forceinline__ uint64_t* agg_find_or_insert_warp(const AggParams& A, const KeyRegs& key, uint64_t h, bool need, uint32_t* err) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long probe = need ? h : (0xFFFFFFFF00000000ull | lane);   // idle lanes match nobody useful
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, probe);
  const int leader = __ffs(peers) - 1;
  uint64_t* e = nullptr;
  if (need && (int)lane == leader) e = agg_find_or_insert(A, key, h, err);
  unsigned long long p = __shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(e), leader);
  uint64_t* got = reinterpret_cast<uint64_t*>(p);
  // same hash but different key (64-bit collision inside one warp): fall back to an own lookup
  if (need && (int)lane != leader && got && !key_words_equal(A.keys, A.n_keys, A.has_null_word, got + 2, key))
    got = agg_find_or_insert(A, key, h, err);
  return need ? got : nullptr;
}

// hash of an already packed key (same value pack_key() returns for the row it was packed from)
__device__ __forceinline__ uint64_t hash_packed_key(const AggParams& A, const KeyRegs& key) {
  uint64_t h = 0x243F6A8885A308D3ull;
  int w = A.has_null_word ? 1 : 0;
  for (int i = 0; i < A.n_keys; ++i) {
    if (A.keys[i].width == 16) {
      ulonglong2 v; v.x = key.w[w]; v.y = key.w[w + 1];
      h = mix64(h ^ (A.keys[i].is_view ? view_hash(v) : mix64(v.x ^ mix64(v.y))));
      w += 2;
    } else { h = mix64(h ^ key.w[w]); w += 1; }
  }
  if (A.has_null_word) h = mix64(h ^ key.w[0]);
  return h;
}


__device__ __forceinline__ AccVal load_acc_value(const AccDesc& d, const TileCtx& c, int r) {
  AccVal v; v.i = 0; v.f = 0.0; v.valid = true;
  if (d.valid_slot != NO_SLOT) v.valid = c.arena[eff(c, d.valid_slot) + r] != 0;
  if (d.value_slot == NO_SLOT) return v;
  const uint8_t* p = c.arena + eff(c, d.value_slot) + r * d.stride;
  switch (d.vkind) {
    case K_I32: v.i = lds<int32_t>(p); break;
    case K_I64: v.i = lds<int64_t>(p); break;
    case K_I128: v.i = lds<i128>(p); break;
    case K_F64: v.f = lds<double>(p); break;
    case K_B: v.i = *p; break;
    default: break;
  }
  return v;
}


// apply one accumulator update to a GLOBAL table entry (atomics)
__device__ __forceinline__ void acc_global(uint64_t* e, const AggParams& A, int j, const AccVal& v) {
  const AccDesc& d = A.accs[j];
  uint64_t* w = e + 2 + A.key_words + d.word;
  if (d.op == ACC_COUNT) { if (v.valid) atomicAdd(reinterpret_cast<unsigned long long*>(w), 1ull); return; }
  if (!v.valid) return;
  switch (d.op) {
    case ACC_SUM_I64: atomicAdd(reinterpret_cast<unsigned long long*>(w), (unsigned long long)(int64_t)v.i); break;
    case ACC_SUM_I128: atomic_add_i128(w, v.i); break;
    case ACC_SUM_F64: atomicAdd(reinterpret_cast<double*>(w), v.f); break;
    case ACC_MIN_I32: case ACC_MIN_I64: atomicMin(reinterpret_cast<long long*>(w), (long long)(int64_t)v.i); break;
    case ACC_MAX_I32: case ACC_MAX_I64: atomicMax(reinterpret_cast<long long*>(w), (long long)(int64_t)v.i); break;
    case ACC_MIN_I128: atomic_minmax_i128(w, v.i, true); break;
    case ACC_MAX_I128: atomic_minmax_i128(w, v.i, false); break;
    case ACC_MIN_F64: atomic_minmax_f64(w, v.f, true); break;
    case ACC_MAX_F64: atomic_minmax_f64(w, v.f, false); break;
    default: break;
  }
  if (d.track_seen) {
    unsigned long long bit = 1ull << j;
    if (!(*reinterpret_cast<volatile unsigned long long*>(e + 1) & bit)) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), bit);
  }
}


// Hot path for low-cardinality grouping (TPC-H Q1: 4 groups).  Scratch in shared memory, at
// arena + A.hot_smem_off:
//   u64 keys[G][key_words]     CTA-local dictionary of the first G distinct keys seen
//   u64 fps[G]                 key fingerprints (cheap multiply-add mix) for the lookup
//   u64 entry[G]               global table entries (resolved lazily / at flush)
//   u64 wacc[(warp*G + g) * (1 + 2*n_accs)]   per-WARP accumulators: [seen][acc0 lo,hi][acc1 lo,hi]...
// Every row finds its group id by fingerprint; then, accumulator by accumulator and group by
// group, the warp reduces its rows with shuffles and lane 0 folds the warp total into wacc.  No
// atomics, no per-thread state, ~G*n_accs*30 instructions per warp-tile regardless of tile size.
constexpr int NWARPS = NT / 32;
struct HotView {
  uint32_t* fp32; uint64_t* keys; uint64_t* fps; uint64_t* entry; uint64_t* wacc; int G; int aw;
};
__device__ __forceinline__ HotView hot_view(const AggParams& A, uint8_t* arena) {
  HotView h; h.G = A.hot_groups; h.aw = 1 + 2 * A.n_accs;
  uint8_t* p = arena + A.hot_smem_off;
  h.fp32 = reinterpret_cast<uint32_t*>(p); p += 32;                    // 8 x u32 fingerprints (register path: one LDS.128)
  h.keys = reinterpret_cast<uint64_t*>(p); p += (size_t)h.G * HOT_KEY_WORDS * 8;
  h.fps = reinterpret_cast<uint64_t*>(p); p += (size_t)h.G * 8;
  h.entry = reinterpret_cast<uint64_t*>(p); p += (size_t)h.G * 8;
  h.wacc = reinterpret_cast<uint64_t*>(p);
  return h;
}

// packed key of row r as 8-byte words held in registers (static indexing) + its fingerprint
__device__ __forceinline__ uint64_t pack_words(const AggParams& A, const TileCtx& c, int r, uint64_t (&kw)[HOT_KEY_WORDS]) {
  uint64_t nullmask = 0;
  uint64_t fp = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int w = 0; w < HOT_KEY_WORDS; ++w) {
    kw[w] = 0;
    if (w < A.key_words && !(A.has_null_word && w == 0)) {
      const KeyWord& d = A.kwords[w];
      const uint8_t* p = c.arena + d.slot + r * d.stride + d.byte_off;
      uint64_t v = d.width == 8 ? lds<uint64_t>(p) : d.width == 4 ? (uint64_t)lds<uint32_t>(p) : (uint64_t)*p;
      if (d.valid_slot != NO_SLOT && c.arena[d.valid_slot + r] == 0) { v = 0; nullmask |= 1ull << d.key_index; }
      kw[w] = v;
      fp = ((fp << 9) | (fp >> 55)) ^ v;        // only a fast reject: every candidate is verified word by word
    }
  }
  if (A.has_null_word) { kw[0] = nullmask; fp = ((fp << 9) | (fp >> 55)) ^ nullmask; }
  return fp;
}

// dictionary entries are stored padded to HOT_KEY_WORDS words, so the verify is 4 unconditional compares
__device__ __forceinline__ int hot_lookup(const AggParams& A, const HotView& H, int hot_n, const uint64_t (&kw)[HOT_KEY_WORDS], uint64_t fp) {
  for (int g = 0; g < hot_n; ++g) {
    if (H.fps[g] != fp) continue;
    const uint64_t* hk = H.keys + g * HOT_KEY_WORDS;
    if (hk[0] == kw[0] && hk[1] == kw[1] && hk[2] == kw[2] && hk[3] == kw[3]) return g;
  }
  return -1;
}

__device__ __noinline__ uint64_t* hot_entry(const PipelineParams& P, const AggParams& A, const HotView& H, int g) {
  uint64_t* e = reinterpret_cast<uint64_t*>(H.entry[g]);
  if (e) return e;
  KeyRegs key;
  for (int w = 0; w < MAX_KEY_WORDS; ++w) key.w[w] = (w < A.key_words && w < HOT_KEY_WORDS) ? H.keys[g * HOT_KEY_WORDS + w] : 0;
  e = agg_find_or_insert(A, key, hash_packed_key(A, key), P.error_flag);
  H.entry[g] = reinterpret_cast<uint64_t>(e);     // benign race: every writer stores the same pointer
  return e;
}



// rows whose group is not in the CTA-local dictionary: global table, one warp-cooperative lookup per row slot

template <int RPT>
__device__ __noinline__ void agg_cold_rows(const PipelineParams& P, const AggParams& A, const TileCtx& c, const int (&gid)[RPT], const bool (&live)[RPT]) {
  // the table walk below is a chain of dependent global accesses per row: start the first state word and entry line of
  // every row of this thread on their way first
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    if (live[k] && gid[k] < 0) {
      KeyRegs key; bool hn;
      const uint64_t h = pack_key<MAX_KEYS>(A.keys, A.n_keys, A.has_null_word, c, threadIdx.x + k * NT, key, &hn);
      const uint64_t idx = h & A.capacity_mask;
      prefetch_l2(A.state + idx);
      prefetch_l2(reinterpret_cast<const uint64_t*>(A.table) + idx * A.entry_words);
    }
  }
  for (int k = 0; k < RPT; ++k) {
    const bool cold = live[k] && gid[k] < 0;
    if (__any_sync(0xFFFFFFFFu, cold)) {
      const int r = threadIdx.x + k * NT;
      KeyRegs key; bool hn; uint64_t h = 0;
      if (cold) h = pack_key<MAX_KEYS>(A.keys, A.n_keys, A.has_null_word, c, r, key, &hn);
      uint64_t* e = agg_find_or_insert_warp(A, key, h, cold, P.error_flag);
      if (cold && e) {
        for (int j = 0; j < A.n_accs; ++j) acc_global(e, A, j, load_acc_value(A.accs[j], c, r));
      }
    }
  }
}

// high-cardinality variant (AggParams::cold_only): no dictionary, every live row goes to the global table
template <int RPT>
__device__ __forceinline__ void sink_agg_cold(const PipelineParams& P, const AggParams& A, const TileCtx& c) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + P.mask_slot;
  int gid[RPT];
  bool live[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    live[k] = r < c.nrows && (pact == nullptr || pact[r]);
    gid[k] = -1;
  }
  agg_cold_rows<RPT>(P, A, c, gid, live);
}

template <int RPT>
__device__ __forceinline__ void sink_agg(const PipelineParams& P, const AggParams& A, const TileCtx& c, Smem* sm) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + P.mask_slot;
  HotView H = hot_view(A, c.arena);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int gid[RPT];
  bool live[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    live[k] = r < c.nrows && (pact == nullptr || pact[r]);
    gid[k] = -1;
  }
  if (A.hot_groups > 0) {
    // phase A: look every live row up in the CTA-local dictionary
    bool miss = false;
    const int hot_n0 = sm->hot_n;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (live[k]) {
        uint64_t kw[HOT_KEY_WORDS];
        const uint64_t fp = pack_words(A, c, threadIdx.x + k * NT, kw);
        gid[k] = hot_lookup(A, H, hot_n0, kw, fp);
        miss |= gid[k] < 0;
      }
    }
    // phase B: grow the dictionary one group per round while there is room (rare after warm-up)
    while (__syncthreads_or(miss && sm->hot_n < A.hot_groups)) {
      if (threadIdx.x == 0) sm->elect = NT;
      __syncthreads();
      // (the dictionary size is read once, before the barrier: the winner appends at that index.  Reading it again next to
      // `elect` let the compiler fetch both words with one 64-bit load in EVERY thread, which racecheck rightly reports against
      // the winner's store -- harmless, the losers never use the value, but there is no reason to keep it)
      const int hot_cur = sm->hot_n;
      const bool want = miss && hot_cur < A.hot_groups;
      if (want) atomicMin(&sm->elect, (int)threadIdx.x);
      __syncthreads();
      if (want && sm->elect == (int)threadIdx.x) {
        for (int k = 0; k < RPT; ++k) {
          if (live[k] && gid[k] < 0) {
            uint64_t kw[HOT_KEY_WORDS];
            const uint64_t fp = pack_words(A, c, threadIdx.x + k * NT, kw);
            const int g = hot_cur;
            for (int w = 0; w < HOT_KEY_WORDS; ++w) H.keys[g * HOT_KEY_WORDS + w] = kw[w];
            H.fps[g] = fp;
            H.entry[g] = 0;
            sm->hot_n = g + 1;
            break;
          }
        }
      }
      __syncthreads();
      miss = false;
      const int hot_n1 = sm->hot_n;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        if (live[k] && gid[k] < 0) {
          uint64_t kw[HOT_KEY_WORDS];
          const uint64_t fp = pack_words(A, c, threadIdx.x + k * NT, kw);
          gid[k] = hot_lookup(A, H, hot_n1, kw, fp);
          miss |= gid[k] < 0;
        }
      }
    }
    // phase C: accumulator by accumulator, group by group: thread-local partial -> warp reduce -> lane 0
    const int hot_n = sm->hot_n;
    bool anyhot = false;
#pragma unroll
    for (int k = 0; k < RPT; ++k) anyhot |= live[k] && gid[k] >= 0;
    if (__any_sync(0xFFFFFFFFu, anyhot)) {
      for (int j = 0; j < A.n_accs; ++j) {
        const AccDesc& d = A.accs[j];
        i128 vi[RPT]; double vf[RPT]; bool ok[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          vi[k] = 0; vf[k] = 0.0; ok[k] = false;
          if (live[k] && gid[k] >= 0) {
            const int r = threadIdx.x + k * NT;
            ok[k] = d.valid_slot == NO_SLOT || c.arena[d.valid_slot + r] != 0;
            if (d.value_slot != NO_SLOT) {
              const uint8_t* p = c.arena + d.value_slot + r * d.stride;
              switch (d.vkind) {
                case K_I32: vi[k] = lds<int32_t>(p); break;
                case K_I64: vi[k] = lds<int64_t>(p); break;
                case K_I128: vi[k] = lds<i128>(p); break;
                case K_F64: vf[k] = lds<double>(p); break;
                default: vi[k] = *p;
              }
            }
            if (d.op == ACC_SUM_I128 && ok[k] && !fits55(vi[k])) {   // rare: exact value straight to the table
              uint64_t* e = hot_entry(P, A, H, gid[k]);
              if (e) { atomic_add_i128(e + 2 + A.key_words + d.word, vi[k]); if (d.track_seen) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), 1ull << j); }
              ok[k] = false;
            }
          }
        }
        for (int g = 0; g < hot_n; ++g) {
          uint64_t* wa = H.wacc + ((size_t)(warp * H.G + g)) * H.aw;
          uint64_t* slot = wa + 1 + 2 * j;
          bool any = false;
#pragma unroll
          for (int k = 0; k < RPT; ++k) any |= ok[k] && gid[k] == g;
          const unsigned members = __ballot_sync(0xFFFFFFFFu, any);
          if (members == 0) continue;
          switch (d.op) {
            case ACC_COUNT: {
              int cnt = 0;
#pragma unroll
              for (int k = 0; k < RPT; ++k) cnt += (ok[k] && gid[k] == g) ? 1 : 0;
              cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
              if (lane == 0) slot[0] += (uint64_t)cnt;
              break;
            }
            case ACC_SUM_I64: case ACC_SUM_I128: {
              int64_t part = 0;
#pragma unroll
              for (int k = 0; k < RPT; ++k) part += (ok[k] && gid[k] == g) ? (int64_t)vi[k] : 0;
              part = warp_sum_i64(part);
              if (lane == 0) {
                if (d.op == ACC_SUM_I64) slot[0] += (uint64_t)part;
                else { const uint64_t lo = slot[0] + (uint64_t)part; slot[1] += (uint64_t)(part >> 63) + (lo < slot[0] ? 1ull : 0ull); slot[0] = lo; }
              }
              break;
            }
            case ACC_SUM_F64: {
              double part = 0.0;
#pragma unroll
              for (int k = 0; k < RPT; ++k) part += (ok[k] && gid[k] == g) ? vf[k] : 0.0;
              part = warp_sum_f64(part);
              if (lane == 0) slot[0] = (uint64_t)__double_as_longlong(__longlong_as_double((long long)slot[0]) + part);
              break;
            }
            default: {   // min / max: serial over member lanes through lane 0 (rarely hot)
              uint64_t w0 = acc_identity(d.op, 0), w1 = acc_identity(d.op, 1);
#pragma unroll
              for (int k = 0; k < RPT; ++k) {
                if (ok[k] && gid[k] == g) {
                  const bool isf = d.op == ACC_MIN_F64 || d.op == ACC_MAX_F64;
                  const uint64_t v0 = isf ? (uint64_t)__double_as_longlong(vf[k]) : (uint64_t)(u128)vi[k];
                  const uint64_t v1 = isf ? 0 : (uint64_t)((u128)vi[k] >> 64);
                  acc_combine_words(d.op, w0, w1, v0, v1);
                }
              }
#pragma unroll
              for (int dlt = 16; dlt; dlt >>= 1) {
                const uint64_t o0 = __shfl_xor_sync(0xFFFFFFFFu, w0, dlt), o1 = __shfl_xor_sync(0xFFFFFFFFu, w1, dlt);
                acc_combine_words(d.op, w0, w1, o0, o1);
              }
              if (lane == 0) { uint64_t a0 = slot[0], a1 = slot[1]; acc_combine_words(d.op, a0, a1, w0, w1); slot[0] = a0; slot[1] = a1; }
            }
          }
          if (d.track_seen && lane == 0) wa[0] |= 1ull << j;
        }
      }
    }
  }
  {
    bool anycold = false;
#pragma unroll
    for (int k = 0; k < RPT; ++k) anycold |= live[k] && gid[k] < 0;
    if (__any_sync(0xFFFFFFFFu, anycold)) agg_cold_rows<RPT>(P, A, c, gid, live);   // warp-level: no CTA barrier
  }
}

// ---- integer fast path: register-resident partials -------------------------------------------------
// racc[g][j]: this thread's running 64-bit partial of accumulator j for hot group g.  Values are
// admitted only below 2^55 in magnitude and the partials are spilled to the CTA accumulators (shared
// memory atomics, once per REG_FLUSH rows) so they can never overflow.
constexpr int REG_FLUSH = 224;    // 224 + RPT values below 2^55 cannot overflow 64 bits
struct RegAcc { int64_t v[REG_GROUPS][REG_ACCS]; int rows; };

__device__ __forceinline__ void reg_flush(const AggParams& A, const HotView& H, RegAcc& R, int hot_n) {
  // CTA accumulators live in warp 0's wacc blocks: slot (g, j) = {lo, hi}.  The 32 lanes of a warp are summed with
  // shuffles first, so only one lane per warp touches the shared accumulators (8-way instead of 256-way contention).
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < REG_GROUPS; ++g) {
    if (g < hot_n) {
      uint64_t* wa = H.wacc + (size_t)g * H.aw;
#pragma unroll
      for (int j = 0; j < REG_ACCS; ++j) {
        if (j < A.n_accs) {
          // |partial| < 2^63 per lane and the warp total may exceed 64 bits: reduce as 128-bit (lo, carry-aware hi)
          const int64_t part = R.v[g][j];
          unsigned long long lo = (unsigned long long)part;
          long long hi = part >> 63;
#pragma unroll
          for (int d = 16; d; d >>= 1) {
            const unsigned long long olo = __shfl_xor_sync(0xFFFFFFFFu, lo, d);
            const long long ohi = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
            const unsigned long long s = lo + olo;
            hi += ohi + (s < lo ? 1 : 0);
            lo = s;
          }
          if (lane == 0 && (lo | (unsigned long long)hi)) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(wa + 1 + 2 * j);
            const unsigned long long old = atomicAdd(dst, lo);
            const unsigned long long carry = (old + lo) < old ? 1ull : 0ull;
            const unsigned long long h2 = (unsigned long long)hi + carry;
            if (h2) atomicAdd(dst + 1, h2);
          }
          R.v[g][j] = 0;
        }
      }
    }
  }
  R.rows = 0;
}

// register path key handling: every key word is a plain 8-byte load (views, int64, decimals) unless the plan has
// narrow or nullable keys, in which case the general packer runs; fingerprints are 32 bits so that the four
// dictionary fingerprints arrive in one LDS.128
__device__ __forceinline__ uint32_t reg_pack(const AggParams& A, const TileCtx& c, int r, uint64_t (&kw)[HOT_KEY_WORDS]) {
  if (A.kw_simple) {
#pragma unroll
    for (int w = 0; w < HOT_KEY_WORDS; ++w)
      kw[w] = w < A.key_words ? lds<uint64_t>(c.arena + A.kwords[w].slot + r * A.kwords[w].stride + A.kwords[w].byte_off) : 0ull;
  } else {
    pack_words(A, c, r, kw);
  }
  uint32_t fp = fold32(kw[0]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[1]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[2]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[3]);
  return fp;
}
__device__ __forceinline__ bool reg_verify(const HotView& H, int g, const uint64_t (&kw)[HOT_KEY_WORDS]) {
  const ulonglong2* hk = reinterpret_cast<const ulonglong2*>(H.keys + g * HOT_KEY_WORDS);
  const ulonglong2 a = hk[0], b = hk[1];
  return a.x == kw[0] && a.y == kw[1] && b.x == kw[2] && b.y == kw[3];
}
__device__ __forceinline__ int reg_lookup(const HotView& H, int hot_n, const uint4& f4, const uint64_t (&kw)[HOT_KEY_WORDS], uint32_t fp) {
  if (hot_n > 0 && f4.x == fp && reg_verify(H, 0, kw)) return 0;
  if (hot_n > 1 && f4.y == fp && reg_verify(H, 1, kw)) return 1;
  if (hot_n > 2 && f4.z == fp && reg_verify(H, 2, kw)) return 2;
  if (hot_n > 3 && f4.w == fp && reg_verify(H, 3, kw)) return 3;
  return -1;
}

template <int RPT>
__device__ __forceinline__ void sink_agg_reg(const PipelineParams& P, const AggParams& A, const TileCtx& c, Smem* sm, RegAcc& R) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + P.mask_slot;
  HotView H = hot_view(A, c.arena);
  int gid[RPT];
  bool live[RPT];
  bool miss = false;
  const int hot_n0 = sm->hot_n;
  const uint4 f4 = *reinterpret_cast<const uint4*>(H.fp32);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    live[k] = r < c.nrows && (pact == nullptr || pact[r]);
    gid[k] = -1;
    if (live[k]) {
      uint64_t kw[HOT_KEY_WORDS];
      const uint32_t fp = reg_pack(A, c, r, kw);
      gid[k] = reg_lookup(H, hot_n0, f4, kw, fp);
      miss |= gid[k] < 0;
    }
  }
  // dictionary growth (only while fewer than REG_GROUPS groups are known): same protocol as sink_agg
  if (hot_n0 < REG_GROUPS) {
    while (__syncthreads_or(miss && sm->hot_n < REG_GROUPS)) {
      if (threadIdx.x == 0) sm->elect = NT;
      __syncthreads();
      const bool want = miss && sm->hot_n < REG_GROUPS;
      if (want) atomicMin(&sm->elect, (int)threadIdx.x);
      __syncthreads();
      if (want && sm->elect == (int)threadIdx.x) {
        for (int k = 0; k < RPT; ++k) {
          if (live[k] && gid[k] < 0) {
            uint64_t kw[HOT_KEY_WORDS];
            const uint32_t fp = reg_pack(A, c, threadIdx.x + k * NT, kw);
            const int g = sm->hot_n;
            for (int w = 0; w < HOT_KEY_WORDS; ++w) H.keys[g * HOT_KEY_WORDS + w] = kw[w];
            H.fp32[g] = fp;
            H.entry[g] = 0;
            sm->hot_n = g + 1;
            break;
          }
        }
      }
      __syncthreads();
      miss = false;
      const int hot_n1 = sm->hot_n;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        if (live[k] && gid[k] < 0) {
          uint64_t kw[HOT_KEY_WORDS];
          const uint32_t fp = reg_pack(A, c, threadIdx.x + k * NT, kw);
          const uint4 g4 = *reinterpret_cast<const uint4*>(H.fp32);
          gid[k] = reg_lookup(H, hot_n1, g4, kw, fp);
          miss |= gid[k] < 0;
        }
      }
    }
  }
  // accumulate: values first (RPT x n_accs loads), then predicated adds into the register partials
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    const bool hot = live[k] && gid[k] >= 0;
    if (hot) {
      int64_t val[REG_ACCS];
#pragma unroll
      for (int j = 0; j < REG_ACCS; ++j) {
        val[j] = 0;
        if (j < A.n_accs) {
          const AggParams::RegLoad& L = A.rload[j];          // constant bank, static index
          if (L.mode == 0) val[j] = 1;
          else {
            const uint8_t* p = c.arena + L.slot + r * L.stride;
            if (L.mode == 3) { val[j] = lds<int64_t>(p); continue; }         // statically below 2^55: no check
            i128 x;
            if (L.mode == 1) x = (i128)lds<int64_t>(p); else x = lds<i128>(p);
            if (fits55(x)) val[j] = (int64_t)x;
            else {                                            // rare: exact value straight to the table entry
              uint64_t* e = hot_entry(P, A, H, gid[k]);
              if (e) atomic_add_i128(e + 2 + A.key_words + A.accs[j].word, x);
            }
          }
        }
      }
      switch (gid[k]) {            // only this row's group is touched (divergent, but 4x fewer adds than predication)
        case 0:
#pragma unroll
          for (int j = 0; j < REG_ACCS; ++j) R.v[0][j] += val[j];
          break;
        case 1:
#pragma unroll
          for (int j = 0; j < REG_ACCS; ++j) R.v[1][j] += val[j];
          break;
        case 2:
#pragma unroll
          for (int j = 0; j < REG_ACCS; ++j) R.v[2][j] += val[j];
          break;
        default:
#pragma unroll
          for (int j = 0; j < REG_ACCS; ++j) R.v[3][j] += val[j];
      }
    }
  }
  R.rows += RPT;
  if (R.rows >= REG_FLUSH) reg_flush(A, H, R, sm->hot_n);
  {
    bool anycold = false;
#pragma unroll
    for (int k = 0; k < RPT; ++k) anycold |= live[k] && gid[k] < 0;
    if (__any_sync(0xFFFFFFFFu, anycold)) agg_cold_rows<RPT>(P, A, c, gid, live);   // warp-level: no CTA barrier
  }
}

__device__ __forceinline__ void hot_init(const AggParams& A, uint8_t* arena) {
  if (A.hot_groups <= 0) return;
  HotView H = hot_view(A, arena);
  for (int i = threadIdx.x; i < NWARPS * H.G; i += NT) {
    uint64_t* wa = H.wacc + (size_t)i * H.aw;
    wa[0] = 0;
    for (int j = 0; j < A.n_accs; ++j) { wa[1 + 2 * j] = acc_identity(A.accs[j].op, 0); wa[2 + 2 * j] = acc_identity(A.accs[j].op, 1); }
  }
}

// end of kernel: fold the per-warp accumulators of every hot group into the global table
__device__ __forceinline__ void hot_flush(const PipelineParams& P, const AggParams& A, uint8_t* arena, Smem* sm) {
  if (A.hot_groups <= 0) return;
  __syncthreads();
  HotView H = hot_view(A, arena);
  const int n = sm->hot_n;
  for (int g = threadIdx.x; g < n; g += NT) hot_entry(P, A, H, g);
  __syncthreads();
  const int per = A.n_accs + 1;   // accumulators + the seen word
  for (int p = threadIdx.x; p < n * per; p += NT) {
    const int g = p / per, j = p % per;
    uint64_t* e = reinterpret_cast<uint64_t*>(H.entry[g]);
    if (!e) continue;
    if (j == A.n_accs) {
      uint64_t seen = 0;
      for (int w = 0; w < NWARPS; ++w) seen |= H.wacc[((size_t)(w * H.G + g)) * H.aw];
      if (seen) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), (unsigned long long)seen);
      continue;
    }
    const AccDesc& d = A.accs[j];
    uint64_t w0 = acc_identity(d.op, 0), w1 = acc_identity(d.op, 1);
    if (d.op == ACC_SUM_I128) { w0 = 0; w1 = 0; }
    for (int w = 0; w < NWARPS; ++w) {
      const uint64_t* slot = H.wacc + ((size_t)(w * H.G + g)) * H.aw + 1 + 2 * j;
      acc_combine_words(d.op, w0, w1, slot[0], slot[1]);
    }
    uint64_t* dst = e + 2 + A.key_words + d.word;
    switch (d.op) {
      case ACC_SUM_I64: case ACC_COUNT: if (w0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)w0); break;
      case ACC_SUM_I128: atomic_add_i128(dst, (i128)(((u128)w1 << 64) | w0)); break;
      case ACC_SUM_F64: atomicAdd(reinterpret_cast<double*>(dst), __longlong_as_double((long long)w0)); break;
      case ACC_MIN_I32: case ACC_MIN_I64: atomicMin(reinterpret_cast<long long*>(dst), (long long)w0); break;
      case ACC_MAX_I32: case ACC_MAX_I64: atomicMax(reinterpret_cast<long long*>(dst), (long long)w0); break;
      case ACC_MIN_I128: atomic_minmax_i128(dst, (i128)(((u128)w1 << 64) | w0), true); break;
      case ACC_MAX_I128: atomic_minmax_i128(dst, (i128)(((u128)w1 << 64) | w0), false); break;
      case ACC_MIN_F64: atomic_minmax_f64(dst, __longlong_as_double((long long)w0), true); break;
      case ACC_MAX_F64: atomic_minmax_f64(dst, __longlong_as_double((long long)w0), false); break;
      default: break;
    }
  }
}

// ================================================================================================
// store / compact sinks
// ================================================================================================
__device__ __forceinline__ void store_value(const OutputCol& o, const TileCtx& c, int r, int64_t pos) {
  const uint8_t* p = c.arena + eff(c, o.slot) + r * o.stride;
  uint8_t* dst = o.data + pos * o.width;
  // (slot stride, output width): 16->16 copy, 8->16 sign-extend (narrow decimal), 8->8, 4->4, 4->1/2 truncate, 8->4
  if (o.width == 16) {
    if (o.stride >= 16) *reinterpret_cast<ulonglong2*>(dst) = *reinterpret_cast<const ulonglong2*>(p);
    else { int64_t v = lds<int64_t>(p); ulonglong2 w; w.x = (unsigned long long)v; w.y = (unsigned long long)(v >> 63); *reinterpret_cast<ulonglong2*>(dst) = w; }
  } else if (o.width == 8) {
    *reinterpret_cast<uint64_t*>(dst) = lds<uint64_t>(p);
  } else if (o.width == 4) {
    *reinterpret_cast<uint32_t*>(dst) = lds<uint32_t>(p);
  } else if (o.width == 2) {
    *reinterpret_cast<uint16_t*>(dst) = (uint16_t)lds<uint32_t>(p);
  } else {
    *dst = *p;
  }
}

template <int RPT>
__device__ __forceinline__ void sink_store(const PipelineParams& P, const TileCtx& c) {
  const int lane = threadIdx.x & 31;
  for (int j = 0; j < P.n_out; ++j) {
    const OutputCol& o = P.out[j];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int r = threadIdx.x + k * NT;
      const bool in = r < c.nrows;
      if (o.width) { if (in) store_value(o, c, r, c.row0 + r); }
      else {   // boolean column -> bitmap word per 32 rows
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, in && c.arena[eff(c, o.slot) + r]);
        if (lane == 0 && (r - lane) < c.nrows) reinterpret_cast<uint32_t*>(o.data)[(c.row0 + r) >> 5] = bits;
      }
      if (o.valid_slot != NO_SLOT) {
        const uint32_t vb = __ballot_sync(0xFFFFFFFFu, in && c.arena[eff(c, o.valid_slot) + r]);
        if (lane == 0 && (r - lane) < c.nrows) reinterpret_cast<uint32_t*>(o.valid_bytes)[(c.row0 + r) >> 5] = vb;
      }
    }
  }
}

// look-back word: bits 63..62 = 0 invalid / 1 tile aggregate / 2 inclusive prefix
template <int RPT>
__device__ __forceinline__ void sink_compact(const PipelineParams& P, const TileCtx& c, Smem* sm, int tile) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + eff(c, P.mask_slot);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool keep[RPT]; uint32_t before[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    keep[k] = r < c.nrows && (pact == nullptr || pact[r]);
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, keep[k]);
    before[k] = __popc(b & ((1u << lane) - 1));
    if (lane == 0) sm->warp_sums[k * (NT / 32) + warp] = __popc(b);
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive scan of the RPT * 8 (<= 32) warp totals, in row order
    const int n = RPT * (NT / 32);
    uint32_t v = lane < n ? sm->warp_sums[lane] : 0, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane < n) sm->warp_sums[lane] = incl - v;
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    // decoupled look-back, one warp wide: 32 predecessor tiles are inspected per step; the walk stops at the
    // closest tile that already published an inclusive prefix (flag 2) and adds the aggregates (flag 1) in between
    unsigned long long excl = 0;
    if (P.tile_offsets) {
      excl = P.tile_offsets[tile];             // two-pass filter: the mask pass already fixed every tile's position
    } else if (tile > 0) {
      if (lane == 0) st_release_u64(P.tile_status + tile, (1ull << 62) | total);
      int hi = tile - 1;                       // newest tile not yet accounted for
      uint32_t spins = 0;
      for (;;) {
        const int t = hi - lane;
        unsigned long long sw = t >= 0 ? ld_acquire_u64(P.tile_status + t) : (2ull << 62);   // before tile 0: prefix 0
        const unsigned flag = (unsigned)(sw >> 62);
        const unsigned ready = __ballot_sync(0xFFFFFFFFu, flag != 0);
        const unsigned prefix = __ballot_sync(0xFFFFFFFFu, flag == 2);
        // usable window: lanes 0..first_prefix (inclusive) if all of them are ready
        const int first_prefix = prefix ? __ffs(prefix) - 1 : 32;
        const unsigned need = first_prefix >= 31 ? 0xFFFFFFFFu : ((2u << first_prefix) - 1);
        if ((ready & need) != need) { if (++spins > (1u << 24)) __trap(); continue; }      // somebody in the window is not published yet
        unsigned long long v = (lane <= first_prefix) ? (sw & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        excl += v;
        if (first_prefix < 32) break;          // reached a published prefix
        hi -= 32;
      }
    }
    if (lane == 0) {
      sm->tile_base = excl;
      if (!P.tile_offsets) {
        st_release_u64(P.tile_status + tile, (2ull << 62) | (excl + total));
        if ((int64_t)(tile + 1) * P.tile_rows >= P.n_rows) *P.out_count = excl + total;
      }
    }
  }
  __syncthreads();
  const unsigned long long base = sm->tile_base;
  for (int j = 0; j < P.n_out; ++j) {
    const OutputCol& o = P.out[j];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (!keep[k]) continue;
      const int r = threadIdx.x + k * NT;
      const int64_t pos = (int64_t)(base + sm->warp_sums[k * (NT / 32) + warp] + before[k]);
      if (o.width) store_value(o, c, r, pos);
      else o.data[pos] = c.arena[eff(c, o.slot) + r];                  // boolean column as bytes (packed later)
      if (o.valid_slot != NO_SLOT) o.valid_bytes[pos] = c.arena[eff(c, o.valid_slot) + r];
    }
  }
}

// ================================================================================================
// join build sink / partition sink
// ================================================================================================
// does build row `cand` carry the key packed in `key`?
__device__ __forceinline__ bool build_row_has_key(const KeyDesc* keys, const uint8_t* const* cols, const uint8_t* strides, int n_keys,
                                                  int64_t cand, const KeyRegs& key) {
  int w = 0;
  for (int i = 0; i < n_keys; ++i) {
    const KeyDesc& d = keys[i];
    const uint8_t* bp = cols[i] + cand * strides[i];
    if (d.width == 16) {
      ulonglong2 bv = *reinterpret_cast<const ulonglong2*>(bp);
      if (d.is_view) { ulonglong2 pv; pv.x = key.w[w]; pv.y = key.w[w + 1]; if (!view_equal(bv, pv)) return false; }
      else if (bv.x != key.w[w] || bv.y != key.w[w + 1]) return false;
      w += 2;
    } else {
      if (load_key_word(bp, d.width) != key.w[w]) return false;
      w += 1;
    }
  }
  return true;
}

// Join table slot = { u64 tag (0 = empty; hash | 1), u64 head (row + 1 of the newest row with this key; 0 = none yet) }.
// The tag is claimed once with a 64-bit CAS and never changes; rows are pushed on the slot's stack with ONE atomic
// exchange of `head` (wait-free: no retry loop, so a build side with a handful of distinct keys does not collapse into
// CAS retries), and next[] links them.  The probe runs in a later launch and reads slots with plain loads.
__device__ __forceinline__ void chain_push(const BuildParams& B, uint8_t* slot, long long first, long long last) {
  const unsigned long long prev = atomicExch(reinterpret_cast<unsigned long long*>(slot + 8), (unsigned long long)(first + 1));
  B.next[last] = (long long)prev - 1;                            // -1 ends the chain
  if ((prev != 0 || first != last) && *reinterpret_cast<volatile uint32_t*>(B.dup_flag) == 0) atomicOr(B.dup_flag, 1u);
}
// Finds the slot of `key`, claiming an empty one if the key is new.  A claimer publishes its pre-linked chain first..last
// at once (*pushed = true) so that later arrivals can verify their key against a row of the chain; for an existing key
// the slot is returned and the caller decides where to push (CTA chain cache or the slot itself).  nullptr: table full.
__device__ __forceinline__ uint8_t* build_find_or_claim(const BuildParams& B, const KeyRegs& key, uint64_t h, long long first, long long last, bool* pushed) {
  const unsigned long long tag = h | 1ull;
  uint64_t idx = (h >> 1) & B.capacity_mask;
  *pushed = false;
  for (uint64_t probes = 0; probes <= B.capacity_mask; ++probes) {
    uint8_t* slot = B.table + idx * 16;
    unsigned long long t = ld_volatile_u64(slot);
    if (t == 0ull) {
      t = atomicCAS(reinterpret_cast<unsigned long long*>(slot), 0ull, tag);
      if (t == 0ull) { chain_push(B, slot, first, last); *pushed = true; return slot; }      // claimed an empty slot
    }
    if (t == tag) {
      // same tag: compare with a row that is already on the chain (the claimer publishes its rows right after the CAS)
      unsigned long long head = ld_volatile_u64(slot + 8);
      for (uint32_t spins = 0; head == 0ull; ++spins) {
        if (spins > (1u << 24)) __trap();
        __nanosleep(20);
        head = ld_volatile_u64(slot + 8);
      }
      if (build_row_has_key(B.keys, B.key_cols, B.key_stride, B.n_keys, (int64_t)head - 1, key)) return slot;
    }
    idx = (idx + 1) & B.capacity_mask;
  }
  return nullptr;
}

// CTA chain cache (shared memory, lives for the whole kernel): chains for keys that already own a slot are collected
// per CTA and pushed on the slot once -- at the end of the kernel, or when the cache runs full.  A build side with a
// handful of distinct keys (a fact-like intermediate joined to a dimension: TPC-H Q5 / Q7 shapes) otherwise funnels
// millions of atomic exchanges into a few addresses (measured 37 ns each: 560 ms for an 18 M-row build).
struct ChainCacheEntry { unsigned long long slot, first; long long last; };
constexpr int CHAIN_CACHE_ENTRIES = 512;
__device__ __forceinline__ bool chain_cache_push(const BuildParams& B, ChainCacheEntry* cache, uint32_t* count, uint8_t* slot, long long first, long long last) {
  const unsigned long long sp = reinterpret_cast<unsigned long long>(slot);
  uint32_t idx = (uint32_t)mix64(sp) & (CHAIN_CACHE_ENTRIES - 1);
  for (int probe = 0; probe < 8; ++probe) {
    ChainCacheEntry* e = cache + idx;
    const unsigned long long s = atomicCAS(&e->slot, 0ull, sp);
    if (s == 0ull) atomicAdd(count, 1u);
    if (s == 0ull || s == sp) {
      const unsigned long long prev = atomicExch(&e->first, (unsigned long long)(first + 1));
      if (prev == 0ull) e->last = last;                  // this chain is the bottom of the cached stack
      else B.next[last] = (long long)prev - 1;
      return true;
    }
    idx = (idx + 1) & (CHAIN_CACHE_ENTRIES - 1);
  }
  return false;
}
// all threads of the CTA; ends with a barrier
__device__ __forceinline__ void chain_cache_flush(const BuildParams& B, ChainCacheEntry* cache, uint32_t* count) {
  __syncthreads();
  for (int i = threadIdx.x; i < CHAIN_CACHE_ENTRIES; i += NT) {
    ChainCacheEntry* e = cache + i;
    if (e->slot) {
      chain_push(B, reinterpret_cast<uint8_t*>(e->slot), (long long)e->first - 1, e->last);
      e->slot = 0ull; e->first = 0ull;
    }
  }
  if (threadIdx.x == 0) *count = 0;
  __syncthreads();
}

// HashJoinExec build: one slot per DISTINCT key; rows with an equal key are pushed on the slot's chain (next[]).
// Lanes of a warp that carry the same key are linked to each other first (__match_any_sync); the lowest lane of each
// group then either claims the key's slot and publishes the group's chain, or -- the key exists already -- parks the
// chain in the CTA chain cache.
template <int RPT>
__device__ __forceinline__ void sink_build(const PipelineParams& P, const BuildParams& B, const TileCtx& c, Smem* sm) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + eff(c, P.mask_slot);
  const int lane = threadIdx.x & 31;
  ChainCacheEntry* cache = reinterpret_cast<ChainCacheEntry*>(c.arena + B.smem_off);
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    const bool live = r < c.nrows && (pact == nullptr || pact[r]);
    KeyRegs key; bool has_null = false; uint64_t h = 0;
    const long long row = B.row_base + c.row0 + r;
    if (live) { h = pack_key<MAX_KEYS>(B.keys, B.n_keys, 0, c, r, key, &has_null); B.next[row] = -1; }
    const bool ins = live && !has_null;                         // NULL keys never match (NullEqualsNothing)
    const unsigned long long tag = h | 1ull;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, ins ? tag : (0xFFFFFFFFFFFFFF00ull | (unsigned)lane) & ~1ull);
    const int leader = __ffs(peers) - 1;
    const long long leader_row = __shfl_sync(0xFFFFFFFFu, row, leader);
    // members of a group: the leader and the peers whose key really equals the leader's (equal hash is not enough)
    const bool follower = ins && lane != leader;
    const bool same = follower && build_row_has_key(B.keys, B.key_cols, B.key_stride, B.n_keys, (int64_t)leader_row, key);
    const bool member = ins && (lane == leader || same);
    const unsigned cm = __ballot_sync(0xFFFFFFFFu, member) & peers;
    long long last_row = row;
    if (member) {
      const unsigned above = cm & ~((2u << lane) - 1);          // members in higher lanes: link in lane order
      const int nxt = above ? __ffs(above) - 1 : -1;
      const long long next_row = __shfl_sync(cm, row, nxt >= 0 ? nxt : lane);
      if (nxt >= 0) B.next[row] = next_row;
      last_row = __shfl_sync(cm, row, 31 - __clz(cm));
    }
    if (member && lane == leader) {
      bool pushed;
      uint8_t* slot = build_find_or_claim(B, key, h, row, last_row, &pushed);
      if (slot && !pushed && !chain_cache_push(B, cache, &sm->cache_count, slot, row, last_row)) chain_push(B, slot, row, last_row);
    } else if (follower && !same) {                             // same hash, different key (64-bit collision)
      bool pushed;
      uint8_t* slot = build_find_or_claim(B, key, h, row, row, &pushed);
      if (slot && !pushed) chain_push(B, slot, row, row);
    }
  }
}

// hash of the partition key columns of row r exactly as oracle/ops.py::hash_partition_ids: h = mix64(h ^ colhash)
__device__ __forceinline__ uint32_t partition_of(const PartitionParams& Q, const TileCtx& c, int r) {
  uint64_t h = 0;
  for (int i = 0; i < Q.n_keys; ++i) {
    const KeyDesc& d = Q.keys[i];
    const uint8_t* p = c.arena + eff(c, d.slot) + r * d.stride;
    uint64_t ch;
    const bool isnull = d.valid_slot != NO_SLOT && c.arena[eff(c, d.valid_slot) + r] == 0;
    if (isnull) ch = 0x6E756C6C6E756C6Cull;
    else if (d.width == 16) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
      if (d.is_view) {
        // length-seeded chain over 8-byte little-endian words of the string bytes
        const uint32_t len = (uint32_t)v.x;
        const uint8_t* s = view_ptr(v, p);
        uint64_t x = len;
        for (uint32_t o = 0; o < len; o += 8) {
          uint64_t w = 0;
          for (uint32_t b = 0; b < 8 && o + b < len; ++b) w |= (uint64_t)s[o + b] << (8 * b);
          x = mix64(x ^ w);
        }
        ch = mix64(x);
      } else ch = mix64(v.x ^ mix64(v.y));
    } else if (d.width == 8) ch = mix64(lds<uint64_t>(p));
    else if (d.width == 4) ch = mix64((uint64_t)(int64_t)lds<int32_t>(p));
    else ch = mix64((uint64_t)*p);
    h = mix64(h ^ ch);
  }
  return (uint32_t)(h % (uint64_t)Q.n_parts);
}

// RepartitionExec Hash: per tile, rows are counted per partition in shared memory (warp-aggregated:
// __match_any_sync elects one lane per distinct partition id in the warp), the CTA reserves one contiguous range per
// non-empty partition with a single global atomic, and every row scatters to base[pid] + its rank.
// pass 0 only accumulates the global histogram.
template <int RPT>
__device__ __forceinline__ void sink_partition(const PipelineParams& P, const PartitionParams& Q, const TileCtx& c) {
  const uint8_t* pact = P.mask_slot == NO_SLOT ? nullptr : c.arena + eff(c, P.mask_slot);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(c.arena + Q.smem_off);
  unsigned long long* base = reinterpret_cast<unsigned long long*>(c.arena + Q.smem_off + (((size_t)Q.n_parts * 4 + 7) & ~(size_t)7));
  const int lane = threadIdx.x & 31;
  for (int p = threadIdx.x; p < Q.n_parts; p += NT) cnt[p] = 0;
  __syncthreads();
  uint32_t pid[RPT], rank[RPT];
  bool live[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int r = threadIdx.x + k * NT;
    live[k] = r < c.nrows && (pact == nullptr || pact[r]);
    pid[k] = live[k] ? partition_of(Q, c, r) : 0xFFFFFFFFu - lane;       // idle lanes match nobody
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, pid[k]);
    const int leader = __ffs(peers) - 1;
    uint32_t first = 0;
    if (live[k] && lane == leader) first = atomicAdd(&cnt[pid[k]], (uint32_t)__popc(peers));
    first = __shfl_sync(0xFFFFFFFFu, first, leader);
    rank[k] = first + __popc(peers & ((1u << lane) - 1));
  }
  __syncthreads();
  for (int p = threadIdx.x; p < Q.n_parts; p += NT) {
    const uint32_t n = cnt[p];
    if (n) {
      const unsigned long long at = atomicAdd(Q.part_counts + p, (unsigned long long)n);
      if (Q.pass) base[p] = (unsigned long long)Q.part_offsets[p] + at;
    }
  }
  if (Q.pass == 0) { __syncthreads(); return; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    if (!live[k]) continue;
    const int r = threadIdx.x + k * NT;
    const unsigned long long pos = base[pid[k]] + rank[k];
    for (int j = 0; j < P.n_out; ++j) {
      const OutputCol& o = P.out[j];
      if (o.width) store_value(o, c, r, (int64_t)pos);
      else o.data[pos] = c.arena[eff(c, o.slot) + r];
      if (o.valid_slot != NO_SLOT) o.valid_bytes[pos] = c.arena[eff(c, o.valid_slot) + r];
    }
  }
}

// ================================================================================================
// the kernel
// ================================================================================================
// KC (kernel class) splits the instantiations by sink family: the aggregation kernels carry the register accumulators
// and the bounded-table protocol, the others the ordered compaction / join / partition sinks -- each compiles to less
// code and keeps its registers for what it runs.
enum : int { KC_AGG = 0, KC_OTHER = 1, KC_AGG_COLD = 2 };   // KC_AGG_COLD: global table only (no dictionary, no register accumulators)
template <int RPT, int MINB, int KC>
__global__ void __launch_bounds__(NT, MINB) pipeline_kernel(const __grid_constant__ KernelArgs K, int n_stages) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem* sm = reinterpret_cast<Smem*>(smem_raw);
  uint8_t* arena = smem_raw + SMEM_HDR;
  const PipelineParams& P0 = K.P[0];
  const int tile_rows = RPT * NT;
  const int64_t n_tiles = (P0.n_rows + tile_rows - 1) / tile_rows;
  const bool dynamic = KC == KC_OTHER && P0.sink == SINK_COMPACT && P0.tile_offsets == nullptr;
  // static order walks positions blockIdx.x, +gridDim.x, ...: tile numbers themselves or entries of an explicit list
  const int64_t n_pos = P0.tile_list ? P0.n_list : n_tiles;
  auto tile_of = [&](int64_t pos) -> int64_t { return P0.tile_list ? (int64_t)P0.tile_list[pos] : pos; };
  // bounded aggregation table (AggParams::group_limit): once the table is nearly full this CTA stops taking tiles
  const bool guarded = KC != KC_OTHER && K.aux[0].agg.deferred != nullptr;
  auto groups_now = [&]() -> unsigned long long { return *reinterpret_cast<volatile unsigned long long*>(K.aux[0].agg.n_groups); };
  auto table_full = [&]() -> int { return groups_now() > K.aux[0].agg.group_limit ? 1 : 0; };
  auto defer_rest = [&](int64_t from) {                   // uniform across the CTA
    if (from >= n_pos) return;
    const AggParams& A = K.aux[0].agg;
    const int64_t cnt = (n_pos - from + gridDim.x - 1) / gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) sm->tile_base = atomicAdd(A.n_deferred, (unsigned long long)cnt);
    __syncthreads();
    const unsigned long long base = sm->tile_base;
    for (int64_t j = threadIdx.x; j < cnt; j += NT) A.deferred[base + j] = (uint32_t)tile_of(from + j * gridDim.x);
  };

  if (threadIdx.x == 0) {
    mbar_init(&sm->full[0], 1);
    mbar_init(&sm->full[1], 1);
    fence_barrier_init();
    sm->hot_n = 0;
    sm->tile[0] = dynamic ? (int)atomicAdd(P0.ticket, 1u) : (int)blockIdx.x;
    sm->defer[0] = guarded ? table_full() : 0;
  }
  if (KC == KC_AGG) hot_init(K.aux[0].agg, arena);
  if (KC == KC_OTHER && P0.sink == SINK_BUILD) {
    unsigned long long* z = reinterpret_cast<unsigned long long*>(arena + K.aux[0].build.smem_off);
    for (int i = threadIdx.x; i < CHAIN_CACHE_ENTRIES * 3; i += NT) z[i] = 0ull;
    if (threadIdx.x == 0) sm->cache_count = 0;
  }
  __syncthreads();

  uint32_t tma_bytes = 0;
  for (int i = 0; i < P0.n_inputs; ++i) tma_bytes += tile_bytes_of(P0.in[i], tile_rows);

  // returns true when the tile was copied cooperatively (generic proxy): the caller then needs a CTA barrier before
  // anybody reads it; TMA tiles are ordered by their mbarrier instead
  auto issue = [&](int64_t tile, int stage) -> bool {      // uniform across the CTA
    const PipelineParams& P = K.P[stage];
    const int64_t row0 = tile * tile_rows;
    const int nrows = (int)min((int64_t)tile_rows, P.n_rows - row0);
    if (P.use_tma && nrows == tile_rows) {
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&sm->full[stage], tma_bytes);
        for (int i = 0; i < P.n_inputs; ++i) {
          const InputCol& in = P.in[i];
          const uint32_t bytes = tile_bytes_of(in, tile_rows);
          const uint8_t* src = in.data + (in.width ? (int64_t)in.width * row0 : (row0 >> 3));
          tma_load_1d(arena + in.slot, src, bytes, &sm->full[stage]);
        }
      }
      return false;
    }
    load_tile_generic(P, arena, row0, nrows);
    return true;
  };

  RegAcc R;
#pragma unroll
  for (int g = 0; g < REG_GROUPS; ++g)
#pragma unroll
    for (int j = 0; j < REG_ACCS; ++j) R.v[g][j] = 0;
  R.rows = 0;
  // `cur` / `nxt` are tile numbers in ticket mode and positions (see n_pos) in static mode
  int64_t cur = sm->tile[0];
  const int64_t n_end = dynamic ? n_tiles : n_pos;
  uint32_t parity[2] = {0, 0};
  int it = 0;
  bool defer = sm->defer[0] != 0;
  if (defer) { defer_rest(cur); cur = n_end; }
  if (cur < n_end && issue(dynamic ? cur : tile_of(cur), 0)) __syncthreads();
  for (;; ++it) {
    if (cur >= n_end) break;
    const int s = (n_stages == 2) ? (it & 1) : 0;
    const PipelineParams& P = K.P[s];
    const PipelineAux* aux = &K.aux[s];
    int64_t nxt;
    if (dynamic) {       // ticket order (decoupled look-back needs tiles to start in order): published through shared memory
      if (threadIdx.x == 0) sm->tile[(it + 1) & 1] = (int)atomicAdd(P.ticket, 1u);
      __syncthreads();
      nxt = sm->tile[(it + 1) & 1];
    } else {
      nxt = cur + gridDim.x;                                  // static stride: no barrier needed
    }
    // the group count is read at the top of the tile and only looked at before barrier (B): its latency is hidden
    unsigned long long g_now = 0;
    if (guarded && threadIdx.x == 0) g_now = groups_now();
    const bool prefetched = n_stages == 2 && nxt < n_end && !defer;
    if (prefetched) issue(dynamic ? nxt : tile_of(nxt), s ^ 1);   // prefetch while this tile is computed (consumed after barrier B)
    const int64_t cur_tile = dynamic ? cur : tile_of(cur);
    TileCtx c;
    c.arena = arena; c.stage_off = 0; c.row0 = cur_tile * tile_rows;
    c.nrows = (int)min((int64_t)tile_rows, P.n_rows - c.row0);
    if (P.use_tma && c.nrows == tile_rows) { mbar_wait(&sm->full[s], parity[s]); parity[s] ^= 1; }

    vm_exec<RPT>(K.prog[s], P.n_inst, c, P, aux);
    if constexpr (KC == KC_AGG_COLD) {
      sink_agg_cold<RPT>(P, aux->agg, c);
    } else if constexpr (KC == KC_AGG) {
      if (aux->agg.reg_path) sink_agg_reg<RPT>(P, aux->agg, c, sm, R);
      else sink_agg<RPT>(P, aux->agg, c, sm);
    } else {
      switch (P.sink) {
        case SINK_STORE: sink_store<RPT>(P, c); break;
        case SINK_COMPACT: sink_compact<RPT>(P, c, sm, (int)cur_tile); break;
        case SINK_BUILD: sink_build<RPT>(P, aux->build, c, sm); break;
        case SINK_PARTITION: sink_partition<RPT>(P, aux->part, c); break;
        default: break;
      }
    }
    if (guarded && threadIdx.x == 0) sm->defer[(it + 1) & 1] = g_now > K.aux[0].agg.group_limit ? 1 : 0;
    if (KC == KC_OTHER && P.sink == SINK_BUILD && threadIdx.x == 0) sm->defer[(it + 1) & 1] = sm->cache_count > CHAIN_CACHE_ENTRIES / 2 ? 1 : 0;
    __syncthreads();                                          // (B) stage s and scratch are free again
    if (KC == KC_OTHER && P.sink == SINK_BUILD && sm->defer[(it + 1) & 1]) {      // crowded cache: push everything, start over
      chain_cache_flush(aux->build, reinterpret_cast<ChainCacheEntry*>(arena + aux->build.smem_off), &sm->cache_count);
    }
    if (guarded) {
      const bool was = defer;
      defer = sm->defer[(it + 1) & 1] != 0;
      // a tile that is already on its way (prefetched) is still processed; everything after it is handed back
      if (n_stages == 2 ? (was && !prefetched) : defer) { defer_rest(nxt); break; }
    }
    if (n_stages == 1 && nxt < n_end && issue(dynamic ? nxt : tile_of(nxt), 0)) __syncthreads();
    cur = nxt;
  }
  if (KC == KC_OTHER && P0.sink == SINK_BUILD)
    chain_cache_flush(K.aux[0].build, reinterpret_cast<ChainCacheEntry*>(arena + K.aux[0].build.smem_off), &sm->cache_count);
  if constexpr (KC == KC_AGG) {
    if (K.aux[0].agg.reg_path) { HotView H = hot_view(K.aux[0].agg, arena); reg_flush(K.aux[0].agg, H, R, sm->hot_n); }
    hot_flush(P0, K.aux[0].agg, arena, sm);
  }
}

// ================================================================================================
// helper kernels
// ================================================================================================
// direct-key tables: every entry starts as [0, 0, DIRECT_EMPTY_KEY, accumulator identities]
__global__ void agg_init_direct_kernel(AggParams A, uint64_t n_entries) {
  const uint64_t words = n_entries * A.entry_words;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t w = (uint32_t)(i % A.entry_words);
    uint64_t v = 0;
    if (w == 2) v = DIRECT_EMPTY_KEY;
    else if (w >= 3) {
      const int aw = (int)w - 3;
      for (int j = 0; j < A.n_accs; ++j) {
        const int nw = acc_words_of(A.accs[j].op);
        if (aw >= A.accs[j].word && aw < A.accs[j].word + nw) v = acc_identity(A.accs[j].op, aw - A.accs[j].word);
      }
    }
    reinterpret_cast<uint64_t*>(A.table)[i] = v;
  }
}
// direct-key tables: the list of occupied entries (any order), counted into *counter
__global__ void agg_build_occ_kernel(AggParams A, unsigned long long* counter) {
  const uint64_t n_entries = A.capacity_mask + 2;
  const uint64_t rounded = (n_entries + 31) & ~(uint64_t)31;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < rounded; i += (uint64_t)gridDim.x * blockDim.x) {
    bool occ = false;
    if (i < n_entries) {
      const uint64_t* e = reinterpret_cast<const uint64_t*>(A.table) + i * A.entry_words;
      occ = i <= A.capacity_mask ? e[2] != DIRECT_EMPTY_KEY : e[0] != 0ull;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, occ);
    if (m) {
      const unsigned lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (occ) A.occ[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)i;
    }
  }
}

// grow: move every READY entry of `old` into the (larger, empty) table of `A`
__global__ void agg_rehash_kernel(AggParams A, const uint8_t* old_table, const uint32_t* old_occ, uint64_t old_groups, uint32_t* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_groups; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(old_table) + (uint64_t)old_occ[i] * A.entry_words;
    KeyRegs key;
    for (int w = 0; w < A.key_words; ++w) key.w[w] = src[2 + w];
    uint64_t* e = agg_find_or_insert(A, key, hash_packed_key(A, key), err);
    if (!e) continue;
    e[1] = src[1];
    for (int w = 0; w < A.acc_words; ++w) e[2 + A.key_words + w] = src[2 + A.key_words + w];
  }
}

__global__ void agg_migrate_kernel(AggParams A, AggMigrateMap M, const uint8_t* old_table, const uint32_t* old_occ, uint64_t old_groups, uint32_t* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_groups; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(old_table) + (uint64_t)old_occ[i] * M.old_entry_words;
    KeyRegs key;
    for (int w = 0; w < MAX_KEY_WORDS; ++w) key.w[w] = 0;
    for (int w = 0; w < M.old_key_words; ++w) key.w[w + M.key_shift] = src[2 + w];
    uint64_t* e = agg_find_or_insert(A, key, hash_packed_key(A, key), err);
    if (!e) continue;
    uint64_t seen = 0;
    for (int j = 0; j < A.n_accs; ++j) {
      for (int w = 0; w < acc_words_of(A.accs[j].op); ++w) e[2 + A.key_words + A.accs[j].word + w] = src[2 + M.old_key_words + M.acc_src_word[j] + w];
      if (M.seen_src[j] == -1) seen |= 1ull << j;
      else if (M.seen_src[j] >= 0) seen |= ((src[1] >> M.seen_src[j]) & 1ull) << j;
    }
    e[1] = seen;
  }
}

__global__ void agg_extract_kernel(AggParams A, AggExtractParams X, uint64_t n_groups, uint32_t* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_groups; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t* e = reinterpret_cast<const uint64_t*>(A.table) + (uint64_t)A.occ[i] * A.entry_words;
    const unsigned long long pos = i;
    const uint64_t seen = e[1];
    const uint64_t* kw = e + 2;
    const uint64_t* aw = e + 2 + A.key_words;
    for (int cidx = 0; cidx < X.n_cols; ++cidx) {
      const AggOutCol& o = X.cols[cidx];
      uint8_t* dst = o.data + pos * o.width;
      bool valid = true;
      if (o.kind == 0) {
        if (A.has_null_word) valid = !((kw[0] >> o.a) & 1);
        const uint64_t* s = kw + o.key_word;
        if (o.width == 16) {
          reinterpret_cast<uint64_t*>(dst)[0] = s[0];
          reinterpret_cast<uint64_t*>(dst)[1] = o.src_words == 2 ? s[1] : (uint64_t)((int64_t)s[0] >> 63);
        }
        else if (o.width == 8) *reinterpret_cast<uint64_t*>(dst) = s[0];
        else if (o.width == 4) *reinterpret_cast<uint32_t*>(dst) = (uint32_t)s[0];
        else if (o.width == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)s[0];      // Int16 / UInt16 group keys
        else *dst = (uint8_t)s[0];
      } else if (o.kind == 1) {
        const AccDesc& d = A.accs[o.a];
        if (d.track_seen) valid = (seen >> o.a) & 1;
        const uint64_t* s = aw + d.word;
        if (o.width == 16) {
          uint64_t lo = s[0], hi = o.src_words == 2 ? s[1] : (uint64_t)((int64_t)s[0] >> 63);
          reinterpret_cast<uint64_t*>(dst)[0] = valid ? lo : 0; reinterpret_cast<uint64_t*>(dst)[1] = valid ? hi : 0;
        } else if (o.width == 8) *reinterpret_cast<uint64_t*>(dst) = valid ? s[0] : 0;
        else if (o.width == 4) *reinterpret_cast<uint32_t*>(dst) = valid ? (uint32_t)s[0] : 0;
        else if (o.width == 2) *reinterpret_cast<uint16_t*>(dst) = valid ? (uint16_t)s[0] : 0;
        else *dst = valid ? (uint8_t)s[0] : 0;
      } else {
        const AccDesc& ds = A.accs[o.a];
        const uint64_t cnt = aw[A.accs[o.b].word];
        valid = cnt != 0 && (!ds.track_seen || ((seen >> o.a) & 1));
        if (o.is_float) {
          double sum = __longlong_as_double((long long)aw[ds.word]);
          *reinterpret_cast<double*>(dst) = valid ? sum / (double)cnt : 0.0;
        } else {
          // DecimalAverager: (sum * 10^(s_out - s_in)) / count, truncating toward zero
          i128 sum = acc_words_of(ds.op) == 2 ? (i128)(((u128)aw[ds.word + 1] << 64) | aw[ds.word]) : (i128)(int64_t)aw[ds.word];
          i128 mul = (i128)(((u128)o.scale_mul_hi << 64) | o.scale_mul_lo);
          i128 q = 0;
          if (valid) q = (sum * mul) / (i128)cnt;
          reinterpret_cast<uint64_t*>(dst)[0] = (uint64_t)(u128)q; reinterpret_cast<uint64_t*>(dst)[1] = (uint64_t)((u128)q >> 64);
        }
      }
      if (o.nullable && o.valid_bytes) o.valid_bytes[pos] = valid ? 1 : 0;
    }
  }
}

// bytes (one per row) -> Arrow bitmap
__global__ void pack_bytes_kernel(const uint8_t* __restrict__ bytes, uint32_t* __restrict__ bits, int64_t n, unsigned long long* null_count) {
  const int lane = threadIdx.x & 31;
  const int64_t n_round = (n + 31) & ~31ll;
  unsigned long long nulls = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    const bool v = i < n && bytes[i] != 0;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, v);
    if (lane == 0) { bits[i >> 5] = b; }
    if (i < n && !v) ++nulls;
  }
  if (null_count && nulls) atomicAdd(null_count, nulls);
}

// Arrow bitmap -> bytes (used when a nullable / boolean column must be row-addressable, e.g. join payloads)
__global__ void unpack_bits_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ bytes, int64_t n, int64_t bit_offset) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i + bit_offset;
    bytes[i] = (bits[b >> 3] >> (b & 7)) & 1;
  }
}

// Utf8View import: turn (buffer_index, offset) of long views into absolute device pointers.
__global__ void resolve_views_kernel(ulonglong2* views, int64_t n, const uint64_t* __restrict__ buffer_bases) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    ulonglong2 v = views[i];
    if ((uint32_t)v.x > 12) {
      const uint32_t buf = (uint32_t)v.y, off = (uint32_t)(v.y >> 32);
      v.y = buffer_bases[buf] + off;
      views[i] = v;
    }
  }
}
// Utf8 (offsets + bytes) -> resolved views
__global__ void utf8_to_views_kernel(const int32_t* __restrict__ offsets, const uint8_t* __restrict__ bytes, ulonglong2* views, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t o = offsets[i];
    const uint32_t len = (uint32_t)(offsets[i + 1] - o);
    const uint8_t* s = bytes + o;
    ulonglong2 v; v.x = len; v.y = 0;
    if (len <= 12) {
      uint64_t a = 0, b = 0;
      for (uint32_t k = 0; k < len && k < 4; ++k) a |= (uint64_t)s[k] << (8 * k);
      for (uint32_t k = 4; k < len; ++k) b |= (uint64_t)s[k] << (8 * (k - 4));
      v.x |= a << 32; v.y = b;
    } else {
      uint64_t a = 0;
      for (uint32_t k = 0; k < 4; ++k) a |= (uint64_t)s[k] << (8 * k);
      v.x |= a << 32; v.y = reinterpret_cast<uint64_t>(s);
    }
    views[i] = v;
  }
}
// Export: lengths of long strings (0 for inline) -> exclusive scan on host side is avoided by a
// single-block scan for small outputs or the two-pass below for large ones.
__global__ void view_long_lengths_kernel(const ulonglong2* __restrict__ views, int64_t n, uint32_t* __restrict__ lens, int all) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t len = (uint32_t)views[i].x;
    lens[i] = (all || len > 12) ? len : 0;
  }
}
// copy long-string bytes into a compact heap at offs[i] and rewrite views to (buffer 0, offset)
__global__ void views_to_arrow_kernel(ulonglong2* views, int64_t n, const uint64_t* __restrict__ offs, uint8_t* heap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    ulonglong2 v = views[i];
    const uint32_t len = (uint32_t)v.x;
    if (len > 12) {
      const uint8_t* s = reinterpret_cast<const uint8_t*>(v.y);
      uint8_t* d = heap + offs[i];
      for (uint32_t k = 0; k < len; ++k) d[k] = s[k];
      v.y = (uint64_t)0 | (offs[i] << 32);
      views[i] = v;
    }
  }
}
// views -> Utf8 (offsets int32 + bytes)
__global__ void views_to_utf8_kernel(const ulonglong2* __restrict__ views, int64_t n, const uint64_t* __restrict__ offs, int32_t* out_offsets, uint8_t* heap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* vp = reinterpret_cast<const uint8_t*>(views + i);
    ulonglong2 v = views[i];
    const uint32_t len = (uint32_t)v.x;
    const uint8_t* s = len <= 12 ? vp + 4 : reinterpret_cast<const uint8_t*>(v.y);
    uint8_t* d = heap + offs[i];
    for (uint32_t k = 0; k < len; ++k) d[k] = s[k];
    out_offsets[i] = (int32_t)offs[i];
    if (i == n - 1) out_offsets[n] = (int32_t)(offs[i] + len);
  }
}

// exclusive scan of u32 -> u64, single kernel, one CTA walking the array (outputs are small/medium;
// large inputs use chunked partial sums: pass 1 per-block totals, pass 2 applies block offsets)
__global__ void scan_block_totals_kernel(const uint32_t* __restrict__ in, int64_t n, uint64_t* __restrict__ block_totals, int64_t per_block) {
  __shared__ unsigned long long ws[32];
  const int64_t b0 = blockIdx.x * per_block, b1 = min(n, b0 + per_block);
  unsigned long long s = 0;
  for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) s += in[i];
  for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w]; block_totals[blockIdx.x] = t; }
}
__global__ void scan_apply_kernel(const uint32_t* __restrict__ in, int64_t n, const uint64_t* __restrict__ block_offsets, uint64_t* __restrict__ out, int64_t per_block) {
  // one warp per block-chunk walks sequentially in 32-wide steps (chunks are sized so this is cheap)
  const int64_t b0 = blockIdx.x * per_block, b1 = min(n, b0 + per_block);
  const int lane = threadIdx.x;
  unsigned long long run = block_offsets[blockIdx.x];
  for (int64_t i = b0; i < b1; i += 32) {
    const int64_t idx = i + lane;
    const unsigned long long v = idx < b1 ? in[idx] : 0;
    unsigned long long incl = v;
    for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (idx < b1) out[idx] = run + incl - v;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

// ================================================================================================
// host-callable launchers (C++ linkage inside libsailgpu)
// ================================================================================================
typedef void (*PipelineFn)(const KernelArgs, int);
// register budget variants: MINB = resident CTAs per SM the compiler must allow (2 -> 128 regs, 3 -> 80, 4 -> 64)
// instantiated variants: dictionary aggregation always runs 2 CTAs/SM (128 registers for the accumulators), the
// global-table aggregation 4 CTAs/SM, the streaming sinks whatever shared memory allows (2..4)
static PipelineFn pick_kernel(int rpt, int minb, int sink, bool cold) {
  if (sink == SINK_AGG && cold) return rpt == 1 ? pipeline_kernel<1, 4, KC_AGG_COLD> : rpt == 2 ? pipeline_kernel<2, 4, KC_AGG_COLD> : pipeline_kernel<4, 2, KC_AGG_COLD>;
  if (sink == SINK_AGG) return rpt == 1 ? pipeline_kernel<1, 2, KC_AGG> : rpt == 2 ? pipeline_kernel<2, 2, KC_AGG> : pipeline_kernel<4, 2, KC_AGG>;
  if (minb >= 4 && rpt != 4) return rpt == 1 ? pipeline_kernel<1, 4, KC_OTHER> : pipeline_kernel<2, 4, KC_OTHER>;
  if (minb >= 3) return rpt == 1 ? pipeline_kernel<1, 3, KC_OTHER> : rpt == 2 ? pipeline_kernel<2, 3, KC_OTHER> : pipeline_kernel<4, 3, KC_OTHER>;
  return rpt == 1 ? pipeline_kernel<1, 2, KC_OTHER> : rpt == 2 ? pipeline_kernel<2, 2, KC_OTHER> : pipeline_kernel<4, 2, KC_OTHER>;
}
cudaError_t launch_pipeline(const KernelArgs& K, int rpt, int n_stages, size_t smem_bytes, int grid, int minb, cudaStream_t stream) {
  PipelineFn k = pick_kernel(rpt, minb, K.P[0].sink, K.aux[0].agg.cold_only != 0);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  k<<<grid, NT, smem_bytes, stream>>>(K, n_stages);
  return cudaGetLastError();
}

int pipeline_max_ctas_per_sm(int rpt, int minb, size_t smem_bytes, int sink, bool cold) {
  PipelineFn k = pick_kernel(rpt, minb, sink, cold);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, NT, smem_bytes) != cudaSuccess) return 0;
  return n;
}

// two-pass filter: rows kept per tile = popcount of the tile's slice of the packed mask (one warp per tile)
__global__ void tile_popcount_kernel(const uint32_t* __restrict__ bits, int64_t n_rows, int tile_rows, int64_t n_tiles, uint32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int words_per_tile = tile_rows >> 5;
  const int64_t n_words = (n_rows + 31) >> 5;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    uint32_t c = 0;
    for (int w = lane; w < words_per_tile; w += 32) {
      const int64_t g = t * words_per_tile + w;
      if (g < n_words) {
        uint32_t x = bits[g];
        const int64_t rem = n_rows - (g << 5);
        if (rem < 32) x &= (1u << rem) - 1;        // bits past the last row are not rows
        c += __popc(x);
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if (lane == 0) counts[t] = c;
  }
}
cudaError_t launch_tile_popcount(const uint32_t* bits, int64_t n_rows, int tile_rows, int64_t n_tiles, uint32_t* counts, cudaStream_t s) {
  if (n_tiles == 0) return cudaSuccess;
  const int grid = (int)std::min<int64_t>((n_tiles * 32 + 255) / 256, 148 * 8);
  tile_popcount_kernel<<<grid, 256, 0, s>>>(bits, n_rows, tile_rows, n_tiles, counts);
  return cudaGetLastError();
}

cudaError_t launch_agg_rehash(const AggParams& A, const uint8_t* old_table, const uint32_t* old_occ, uint64_t old_groups, uint32_t* err, cudaStream_t s) {
  if (old_groups == 0) return cudaSuccess;
  int grid = (int)std::min<uint64_t>((old_groups + 255) / 256, 148 * 8);
  agg_rehash_kernel<<<grid, 256, 0, s>>>(A, old_table, old_occ, old_groups, err);
  return cudaGetLastError();
}
cudaError_t launch_agg_migrate(const AggParams& A_new, const AggMigrateMap& M, const uint8_t* old_table, const uint32_t* old_occ, uint64_t old_groups, uint32_t* err, cudaStream_t s) {
  if (old_groups == 0) return cudaSuccess;
  int grid = (int)std::min<uint64_t>((old_groups + 255) / 256, 148 * 8);
  agg_migrate_kernel<<<grid, 256, 0, s>>>(A_new, M, old_table, old_occ, old_groups, err);
  return cudaGetLastError();
}
cudaError_t launch_agg_init_direct(const AggParams& A, cudaStream_t s) {
  const uint64_t n_entries = A.capacity_mask + 2;
  const uint64_t words = n_entries * A.entry_words;
  agg_init_direct_kernel<<<(int)std::min<uint64_t>((words + 255) / 256, 148 * 16), 256, 0, s>>>(A, n_entries);
  return cudaGetLastError();
}
cudaError_t launch_agg_build_occ(const AggParams& A, unsigned long long* counter, cudaStream_t s) {
  const uint64_t n_entries = A.capacity_mask + 2;
  agg_build_occ_kernel<<<(int)std::min<uint64_t>((n_entries + 255) / 256, 148 * 16), 256, 0, s>>>(A, counter);
  return cudaGetLastError();
}
cudaError_t launch_agg_extract(const AggParams& A, const AggExtractParams& X, uint64_t n_groups, uint32_t* err, cudaStream_t s) {
  if (n_groups == 0) return cudaSuccess;
  int grid = (int)std::min<uint64_t>((n_groups + 255) / 256, 148 * 8);
  agg_extract_kernel<<<grid, 256, 0, s>>>(A, X, n_groups, err);
  return cudaGetLastError();
}
__global__ void u32_to_bytes_kernel(const uint32_t* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] ? 1 : 0;
}
cudaError_t launch_u32_to_bytes(const uint32_t* in, uint8_t* out, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  u32_to_bytes_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}
cudaError_t launch_pack_bytes(const uint8_t* bytes, uint32_t* bits, int64_t n, unsigned long long* null_count, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  pack_bytes_kernel<<<grid, 256, 0, s>>>(bytes, bits, n, null_count);
  return cudaGetLastError();
}
cudaError_t launch_unpack_bits(const uint8_t* bits, uint8_t* bytes, int64_t n, int64_t bit_offset, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  unpack_bits_kernel<<<grid, 256, 0, s>>>(bits, bytes, n, bit_offset);
  return cudaGetLastError();
}
cudaError_t launch_resolve_views(void* views, int64_t n, const uint64_t* bases, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  resolve_views_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<ulonglong2*>(views), n, bases);
  return cudaGetLastError();
}
cudaError_t launch_utf8_to_views(const int32_t* offsets, const uint8_t* bytes, void* views, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  utf8_to_views_kernel<<<grid, 256, 0, s>>>(offsets, bytes, reinterpret_cast<ulonglong2*>(views), n);
  return cudaGetLastError();
}
__global__ void scan_totals_kernel(uint64_t* t, int64_t nb) {
  const int lane = threadIdx.x; unsigned long long run = 0;
  for (int64_t i = 0; i < nb; i += 32) {
    const int64_t idx = i + lane; const unsigned long long v = idx < nb ? t[idx] : 0; unsigned long long incl = v;
    for (int d = 1; d < 32; d <<= 1) { unsigned long long x = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += x; }
    if (idx < nb) t[idx] = run + incl - v;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  if (lane == 0) t[nb] = run;     // grand total after the offsets
}
// exclusive scan of u32 lengths into u64 offsets; block_scratch[nblocks] receives the grand total
cudaError_t launch_exclusive_scan_u32(const uint32_t* in, int64_t n, uint64_t* out, uint64_t* block_scratch /* >= 1025 u64 */, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int64_t nblocks = std::min<int64_t>(1024, (n + 4095) / 4096);
  const int64_t per_block = (n + nblocks - 1) / nblocks;
  scan_block_totals_kernel<<<(int)nblocks, 256, 0, s>>>(in, n, block_scratch, per_block);
  scan_totals_kernel<<<1, 32, 0, s>>>(block_scratch, nblocks);
  scan_apply_kernel<<<(int)nblocks, 32, 0, s>>>(in, n, block_scratch, out, per_block);
  return cudaGetLastError();
}
cudaError_t launch_view_lengths(const void* views, int64_t n, uint32_t* lens, int all, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  view_long_lengths_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const ulonglong2*>(views), n, lens, all);
  return cudaGetLastError();
}
cudaError_t launch_views_to_arrow(void* views, int64_t n, const uint64_t* offs, uint8_t* heap, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  views_to_arrow_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<ulonglong2*>(views), n, offs, heap);
  return cudaGetLastError();
}
cudaError_t launch_views_to_utf8(const void* views, int64_t n, const uint64_t* offs, int32_t* out_offsets, uint8_t* heap, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  views_to_utf8_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const ulonglong2*>(views), n, offs, out_offsets, heap);
  return cudaGetLastError();
}

}  // namespace sg
