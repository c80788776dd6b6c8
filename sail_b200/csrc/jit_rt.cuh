// jit_rt.cuh -- device runtime of the SPECIALISED pipeline kernels.
//
// pipeline.cu interprets a fused Filter -> Projection -> Aggregate chain (a tile VM over shared-memory slots, sinks
// driven by descriptors).  For inputs that are worth a second of compilation the host generates, per distinct pipeline,
// a struct `G` (jit.cu) and compiles `jit_main<G>` with NVRTC for sm_100a: expressions become straight-line register
// code (no slots, no dispatch), every descriptor a compile-time constant, and the tile loop becomes an mbarrier ring of
// TMA stages without a CTA-wide barrier per tile (warps drift independently; thread 0 re-arms a stage as soon as all
// eight warps have released it).  The kernel argument block (KernelArgs) is the interpreter's: the generated code reads
// only pointers and sizes from it, so the host-side launch path is shared.
//
// Same operator semantics as pipeline.cu (reference: DataFusion FilterExec / ProjectionExec / AggregateExec as driven by
// crates/sail-execution/src/job_runner.rs:64); the two kernels share the group-table layout and can serve one operator.
#pragma once
#include "dev_ops.cuh"

namespace sg {

template <int I> struct IC { static constexpr int value = I; __device__ constexpr operator int() const { return I; } };
template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) { f(IC<I>{}); static_for<I + 1, N>(f); }
}

constexpr int JIT_MAX_STAGES = 4;
constexpr int JIT_NWARPS = NT / 32;
struct JitSmem {
  uint64_t full[JIT_MAX_STAGES];       // "tile landed" (TMA transaction barrier, 1 arrival)
  uint64_t empty[JIT_MAX_STAGES];      // "stage released" (one arrival per warp)
  long long tile_no[JIT_MAX_STAGES];   // tile number of the stage (bit 62: partial tile, copied cooperatively); -1 = end
  unsigned long long tile_base;        // COMPACT: exclusive prefix of the tile
  uint32_t warp_sums[33];
  uint32_t dict_n;                     // groups in the CTA dictionary (release/acquire)
  uint32_t dict_lock;
};
constexpr int JIT_HDR = 256;
static_assert(sizeof(JitSmem) <= JIT_HDR, "JitSmem header");     // jit.hpp: JIT_HDR_BYTES
constexpr long long JIT_PARTIAL = 1ll << 62;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// ---- expression helpers (one call per generated statement) --------------------------------------
__device__ __forceinline__ i128 mk128(uint64_t lo, uint64_t hi) { return (i128)(((u128)hi << 64) | lo); }
__device__ __forceinline__ ulonglong2 mkv16(uint64_t lo, uint64_t hi) { ulonglong2 v; v.x = lo; v.y = hi; return v; }
__device__ __forceinline__ i128 jit_mulw(int64_t a, int64_t b) {
  return mk128((unsigned long long)a * (unsigned long long)b, (unsigned long long)__mul64hi((long long)a, (long long)b));
}
__device__ __forceinline__ i128 jit_mul128_64(i128 a, int64_t b) {
  const unsigned long long alo = (unsigned long long)(u128)a, ahi = (unsigned long long)((u128)a >> 64), ub = (unsigned long long)b;
  return mk128(alo * ub, __umul64hi(alo, ub) + ahi * ub - (b < 0 ? alo : 0ull));
}
template <typename T> __device__ __forceinline__ T jit_divround(T a, T d) {
  T q = a / d, rem = a % d;
  T twice = rem < 0 ? -rem * 2 : rem * 2;
  if (twice >= d) q += (a < 0 ? -1 : 1);
  return q;
}
template <typename T, bool REM> __device__ __forceinline__ T jit_div(T a, T b, bool live, uint32_t* err) {
  if (b == 0) { if (live) atomicOr(err, ERR_DIV_ZERO); return (T)0; }
  return REM ? (T)(a % b) : (T)(a / b);
}
template <typename D, typename S> __device__ __forceinline__ D jit_cvt(S v) {
  // mirrors the interpreter's OP_CVT: integers widen through i128, floats go through int64 towards integers
  constexpr bool sf = sizeof(S) == 8 && (S)0.5 != (S)0, sf32 = sizeof(S) == 4 && (S)0.5 != (S)0;
  constexpr bool df = sizeof(D) == 8 && (D)0.5 != (D)0;
  if constexpr (sf || sf32) {
    const double fv = (double)v;
    if constexpr (df) return fv;
    else if constexpr (sizeof(D) == 16) return (D)(int64_t)fv;
    else if constexpr (sizeof(D) == 1) return (D)(fv != 0.0);
    else return (D)fv;
  } else {
    if constexpr (df) { if constexpr (sizeof(S) == 16) return (double)v; else return (double)(int64_t)v; }
    else if constexpr (sizeof(D) == 1) return (D)(v != 0);
    else return (D)v;
  }
}
__device__ __forceinline__ int32_t jit_date_part(int32_t days, int part) {
  int y, m, d;
  civil_from_days(days, y, m, d);
  return part == 0 ? y : part == 1 ? m : d;
}
// LIKE over a view held in registers: inline strings are matched from a local copy
__device__ __noinline__ bool jit_like(ulonglong2 v, const uint8_t* pat, uint32_t plen, int cls) {
  const uint32_t len = (uint32_t)v.x;
  if (len > 12) return like_match(reinterpret_cast<const uint8_t*>(v.y), len, pat, plen, cls);
  uint8_t tmp[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmp[i] = (uint8_t)(v.x >> (32 + 8 * i));
#pragma unroll
  for (int i = 0; i < 8; ++i) tmp[4 + i] = (uint8_t)(v.y >> (8 * i));
  return like_match(tmp, len, pat, plen, cls);
}
__device__ __forceinline__ uint64_t i128_lo(i128 v) { return (uint64_t)(u128)v; }
__device__ __forceinline__ uint64_t i128_hi(i128 v) { return (uint64_t)((u128)v >> 64); }

// ================================================================================================
// hash-join probe with one key of at most 8 bytes (pipeline.cu::vm_probe_narrow; table slot = {hash | 1, build row + 1})
// ================================================================================================
__device__ __forceinline__ int64_t jit_probe_narrow(const ProbeParams& P, uint64_t key, int width, bool act) {
  if (!act) return -1;
  const uint64_t h = mix64(0x243F6A8885A308D3ull ^ key);
  const uint64_t tag = h | 1ull;
  const uint8_t* bcol = P.build_keys[0];
  const int bstride = P.build_stride[0];
  uint64_t idx = (h >> 1) & P.capacity_mask;
  int64_t row = -1;
  for (;;) {
    const ulonglong2 cur = *reinterpret_cast<const ulonglong2*>(P.table + idx * 16);
    if (cur.x == 0) break;
    if (cur.x == tag && load_key_word(bcol + ((int64_t)cur.y - 1) * bstride, width) == key) { row = (int64_t)cur.y - 1; break; }
    idx = (idx + 1) & P.capacity_mask;
  }
  if (row >= 0 && P.visited) P.visited[row] = 1;
  return row;
}
template <class T>
__device__ __forceinline__ T jit_gather(uint64_t base, int64_t row) {
  if (row < 0) return T{};
  return *reinterpret_cast<const T*>(base + (uint64_t)row * sizeof(T));
}

// ================================================================================================
// global group table (layout and protocol of pipeline.cu::agg_find_or_insert, constants from G)
// ================================================================================================
template <class G>
__device__ __forceinline__ uint64_t* jit_find_or_insert(const AggParams& A, const uint64_t (&kw)[MAX_KEY_WORDS], uint64_t h, uint32_t* err) {
  if constexpr (G::KEY_WORDS == 1) {
    if (A.direct_key) {          // direct-key protocol (vm.h): one CAS on the key word, no fence, no state word
      const unsigned long long k = kw[0];
      if (k == DIRECT_EMPTY_KEY) {
        uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + (A.capacity_mask + 1) * G::ENTRY_WORDS;
        if (*reinterpret_cast<volatile unsigned long long*>(e) == 0ull && atomicCAS(reinterpret_cast<unsigned long long*>(e), 0ull, h | 1ull) == 0ull) atomicAdd(A.n_groups, 1ull);
        return e;
      }
      uint64_t idx = h & A.capacity_mask;
      for (uint64_t probes = 0; probes <= A.capacity_mask; ++probes) {
        uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + idx * G::ENTRY_WORDS;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(e + 2);
        if (cur == DIRECT_EMPTY_KEY) {
          cur = atomicCAS(reinterpret_cast<unsigned long long*>(e + 2), DIRECT_EMPTY_KEY, k);
          if (cur == DIRECT_EMPTY_KEY) {
            e[0] = h | 1ull;
            const unsigned m = __activemask();
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
  }
  const uint32_t tag = (uint32_t)(h >> 34) << 2;
  uint64_t idx = h & A.capacity_mask;
  uint64_t probes = 0;
  uint32_t spins = 0;
  while (probes <= A.capacity_mask) {
    uint64_t* e = reinterpret_cast<uint64_t*>(A.table) + idx * G::ENTRY_WORDS;
    uint32_t* st = A.state + idx;
    uint32_t s = ld_acquire_u32(st);
    if (s == ST_EMPTY) {
      s = atomicCAS(st, ST_EMPTY, tag | ST_LOCKED);
      if (s == ST_EMPTY) {
        e[0] = h;
        e[1] = 0;
#pragma unroll
        for (int w = 0; w < G::KEY_WORDS; ++w) e[2 + w] = kw[w];
        static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
#pragma unroll
          for (int w = 0; w < acc_words_of(G::acc_op(J)); ++w) e[2 + G::KEY_WORDS + G::acc_word(J) + w] = acc_identity(G::acc_op(J), w);
        });
        st_release_u32(st, tag | ST_READY);
        {
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
      if ((s & 3u) == ST_LOCKED) {
        if (++spins > (1u << 24)) { atomicOr(err, ERR_TABLE_FULL); return nullptr; }
        __nanosleep(32);
        continue;
      }
      if (e[0] == h && G::keys_equal(e + 2, kw)) return e;
    }
    idx = (idx + 1) & A.capacity_mask;
    ++probes;
  }
  atomicOr(err, ERR_TABLE_FULL);
  return nullptr;
}

template <class G>
__device__ __forceinline__ uint64_t* jit_find_or_insert_warp(const AggParams& A, const uint64_t (&kw)[MAX_KEY_WORDS], uint64_t h, bool need, uint32_t* err) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long probe = need ? h : (0xFFFFFFFF00000000ull | lane);
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, probe);
  const int leader = __ffs(peers) - 1;
  uint64_t* e = nullptr;
  if (need && (int)lane == leader) e = jit_find_or_insert<G>(A, kw, h, err);
  unsigned long long p = __shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(e), leader);
  uint64_t* got = reinterpret_cast<uint64_t*>(p);
  if (need && (int)lane != leader && got && !G::keys_equal(got + 2, kw)) got = jit_find_or_insert<G>(A, kw, h, err);
  return need ? got : nullptr;
}

// one accumulator update on a global table entry
template <class G, int J>
__device__ __forceinline__ void jit_acc_global(uint64_t* e, const AccVal& v) {
  constexpr int op = G::acc_op(J);
  uint64_t* w = e + 2 + G::KEY_WORDS + G::acc_word(J);
  if constexpr (op == ACC_COUNT) { if (v.valid) atomicAdd(reinterpret_cast<unsigned long long*>(w), 1ull); return; }
  if (!v.valid) return;
  if constexpr (op == ACC_SUM_I64) atomicAdd(reinterpret_cast<unsigned long long*>(w), (unsigned long long)(int64_t)v.i);
  else if constexpr (op == ACC_SUM_I128) atomic_add_i128(w, v.i);
  else if constexpr (op == ACC_SUM_F64) atomicAdd(reinterpret_cast<double*>(w), v.f);
  else if constexpr (op == ACC_MIN_I32 || op == ACC_MIN_I64) atomicMin(reinterpret_cast<long long*>(w), (long long)(int64_t)v.i);
  else if constexpr (op == ACC_MAX_I32 || op == ACC_MAX_I64) atomicMax(reinterpret_cast<long long*>(w), (long long)(int64_t)v.i);
  else if constexpr (op == ACC_MIN_I128) atomic_minmax_i128(w, v.i, true);
  else if constexpr (op == ACC_MAX_I128) atomic_minmax_i128(w, v.i, false);
  else if constexpr (op == ACC_MIN_F64) atomic_minmax_f64(w, v.f, true);
  else if constexpr (op == ACC_MAX_F64) atomic_minmax_f64(w, v.f, false);
  if constexpr (G::acc_seen(J) != 0) {
    const unsigned long long bit = 1ull << J;
    if (!(*reinterpret_cast<volatile unsigned long long*>(e + 1) & bit)) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), bit);
  }
}

// ================================================================================================
// CTA dictionary of hot groups (shared memory).  Entries are immutable once published; dict_n is
// released after the entry is written, so readers need no lock.  Growth takes a CTA-wide spin lock,
// one lane per warp at a time.
// ================================================================================================
struct JitHot { uint32_t* fp32; uint64_t* keys; uint64_t* entry; uint64_t* wacc; };
template <class G> __device__ __forceinline__ JitHot jit_hot(uint8_t* scratch) {
  JitHot h;
  uint8_t* p = scratch;
  h.fp32 = reinterpret_cast<uint32_t*>(p); p += 32;
  h.keys = reinterpret_cast<uint64_t*>(p); p += (size_t)G::HOT_G * HOT_KEY_WORDS * 8;
  h.entry = reinterpret_cast<uint64_t*>(p); p += (size_t)G::HOT_G * 8;
  h.wacc = reinterpret_cast<uint64_t*>(p);
  return h;
}
template <class G> constexpr int jit_aw() { return 1 + 2 * G::N_ACCS; }     // per (warp, group): seen word + {lo, hi} per accumulator

__device__ __forceinline__ uint32_t jit_fp(const uint64_t (&kw)[MAX_KEY_WORDS]) {
  uint32_t fp = fold32(kw[0]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[1]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[2]);
  fp = __funnelshift_l(fp, fp, 7) ^ fold32(kw[3]);
  return fp;
}
__device__ __forceinline__ bool jit_dict_verify(const JitHot& H, int g, const uint64_t (&kw)[MAX_KEY_WORDS]) {
  const ulonglong2* hk = reinterpret_cast<const ulonglong2*>(H.keys + g * HOT_KEY_WORDS);
  const ulonglong2 a = hk[0], b = hk[1];
  return ((a.x ^ kw[0]) | (a.y ^ kw[1]) | (b.x ^ kw[2]) | (b.y ^ kw[3])) == 0ull;
}
template <int CAP>
__device__ __forceinline__ int jit_dict_lookup(const JitHot& H, int n, const uint64_t (&kw)[MAX_KEY_WORDS], uint32_t fp) {
  const uint4 f0 = *reinterpret_cast<const uint4*>(H.fp32);
  if (n > 0 && f0.x == fp && jit_dict_verify(H, 0, kw)) return 0;
  if (n > 1 && f0.y == fp && jit_dict_verify(H, 1, kw)) return 1;
  if (n > 2 && f0.z == fp && jit_dict_verify(H, 2, kw)) return 2;
  if (n > 3 && f0.w == fp && jit_dict_verify(H, 3, kw)) return 3;
  if constexpr (CAP > 4) {
    const uint4 f1 = *reinterpret_cast<const uint4*>(H.fp32 + 4);
    if (n > 4 && f1.x == fp && jit_dict_verify(H, 4, kw)) return 4;
    if (n > 5 && f1.y == fp && jit_dict_verify(H, 5, kw)) return 5;
    if (n > 6 && f1.z == fp && jit_dict_verify(H, 6, kw)) return 6;
    if (n > 7 && f1.w == fp && jit_dict_verify(H, 7, kw)) return 7;
  }
  return -1;
}
// warp-collective: lanes with `want` find their key in the dictionary or append it while there is room (CAP entries);
// returns the group id or -1 (dictionary full)
template <int CAP>
__device__ __noinline__ int jit_dict_add(JitSmem* sm, const JitHot& H, bool want, const uint64_t (&kw)[MAX_KEY_WORDS], uint32_t fp) {
  const int lane = threadIdx.x & 31;
  int g = -1;
  bool gave_up = false;
  for (;;) {
    const bool need = want && g < 0 && !gave_up;
    const unsigned pend = __ballot_sync(0xFFFFFFFFu, need);
    if (!pend) break;
    const int leader = __ffs(pend) - 1;
    if (lane == leader) {
      while (atomicCAS(&sm->dict_lock, 0u, 1u) != 0u) __nanosleep(20);
      const int n = (int)lds_acquire_u32(&sm->dict_n);
      g = jit_dict_lookup<CAP>(H, n, kw, fp);
      if (g < 0) {
        if (n < CAP) {
#pragma unroll
          for (int w = 0; w < HOT_KEY_WORDS; ++w) H.keys[n * HOT_KEY_WORDS + w] = kw[w];
          H.fp32[n] = fp;
          H.entry[n] = 0;
          sts_release_u32(&sm->dict_n, (uint32_t)(n + 1));
          g = n;
        } else gave_up = true;
      }
      __threadfence_block();
      atomicExch(&sm->dict_lock, 0u);
    }
    __syncwarp();
    if (need && lane != leader) {
      const int n = (int)lds_acquire_u32(&sm->dict_n);
      g = jit_dict_lookup<CAP>(H, n, kw, fp);
      if (g < 0 && n >= CAP) gave_up = true;
    }
  }
  return g;
}

template <class G>
__device__ __noinline__ uint64_t* jit_hot_entry(const KernelArgs& K, const JitHot& H, int g) {
  uint64_t* e = reinterpret_cast<uint64_t*>(*reinterpret_cast<volatile uint64_t*>(H.entry + g));
  if (e) return e;
  uint64_t kw[MAX_KEY_WORDS];
#pragma unroll
  for (int w = 0; w < MAX_KEY_WORDS; ++w) kw[w] = (w < G::KEY_WORDS && w < HOT_KEY_WORDS) ? H.keys[g * HOT_KEY_WORDS + w] : 0ull;
  e = jit_find_or_insert<G>(K.aux[0].agg, kw, G::key_hash(kw), K.P[0].error_flag);
  H.entry[g] = reinterpret_cast<uint64_t>(e);       // benign race: every writer stores the same pointer
  return e;
}

// warp-level pre-aggregation of one accumulator (jit_cold_rows): `prepare` turns a row's value into a partial state (count(*) of
// one row = 1), `merge` folds another lane's partial state in, `jit_acc_global_n` applies a partial state to the table entry
template <class G, int J>
__device__ __forceinline__ void jit_acc_prepare(AccVal& v) {
  constexpr int op = G::acc_op(J);
  if constexpr (op == ACC_COUNT) { v.i = v.valid ? 1 : 0; }
  else if constexpr (op == ACC_SUM_I64 || op == ACC_SUM_I128) { if (!v.valid) v.i = 0; }
  else if constexpr (op == ACC_SUM_F64) { if (!v.valid) v.f = 0.0; }
}
template <class G, int J>
__device__ __forceinline__ void jit_acc_merge(AccVal& a, const AccVal& b) {
  constexpr int op = G::acc_op(J);
  if constexpr (op == ACC_COUNT || op == ACC_SUM_I64 || op == ACC_SUM_I128) { a.i += b.i; a.valid = a.valid || b.valid; }
  else if constexpr (op == ACC_SUM_F64) { a.f += b.f; a.valid = a.valid || b.valid; }
  else if constexpr (op == ACC_MIN_I32 || op == ACC_MIN_I64 || op == ACC_MIN_I128) { if (b.valid && (!a.valid || b.i < a.i)) { a.i = b.i; a.valid = true; } }
  else if constexpr (op == ACC_MAX_I32 || op == ACC_MAX_I64 || op == ACC_MAX_I128) { if (b.valid && (!a.valid || b.i > a.i)) { a.i = b.i; a.valid = true; } }
  else if constexpr (op == ACC_MIN_F64) { if (b.valid && (!a.valid || b.f < a.f)) { a.f = b.f; a.valid = true; } }
  else if constexpr (op == ACC_MAX_F64) { if (b.valid && (!a.valid || b.f > a.f)) { a.f = b.f; a.valid = true; } }
}
template <class G, int J>
__device__ __forceinline__ void jit_acc_global_n(uint64_t* e, const AccVal& v) {
  constexpr int op = G::acc_op(J);
  if constexpr (op == ACC_COUNT) {
    if (v.valid && (int64_t)v.i != 0) atomicAdd(reinterpret_cast<unsigned long long*>(e + 2 + G::KEY_WORDS + G::acc_word(J)), (unsigned long long)(int64_t)v.i);
  } else jit_acc_global<G, J>(e, v);
}

// rows whose group is not in the dictionary (or every row of the high-cardinality variant): global table
template <class G>
__device__ __forceinline__ void jit_cold_rows(const KernelArgs& K, const typename G::Row (&rows)[G::RPT], const int (&gid)[G::RPT]) {
  const AggParams& A = K.aux[0].agg;
  uint64_t hh[G::RPT];
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    hh[k] = 0;
    if (rows[k].live && gid[k] < 0) {
      uint64_t kw[MAX_KEY_WORDS];
      G::key_words(rows[k], kw);
      hh[k] = G::key_hash(kw);
      const uint64_t idx = hh[k] & A.capacity_mask;
      prefetch_l2(A.state + idx);
      prefetch_l2(reinterpret_cast<const uint64_t*>(A.table) + idx * G::ENTRY_WORDS);
    }
  }
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    const bool cold = rows[k].live && gid[k] < 0;
    if (__any_sync(0xFFFFFFFFu, cold)) {
      uint64_t kw[MAX_KEY_WORDS];
      G::key_words(rows[k], kw);
      uint64_t* e = jit_find_or_insert_warp<G>(A, kw, hh[k], cold, K.P[0].error_flag);
      // Rows of one group that sit in the same warp (clustered inputs: the 1-7 lineitems of an order are neighbours) are combined
      // in registers first: one set of accumulator atomics per (warp, group) instead of per row.  The atomics are what bounds this
      // path -- every accumulator adds ~1.5 ms per 60 M rows on top of the 3.5 ms of the inserts (profiles/README.md).
      const bool upd = cold && e != nullptr;
      const unsigned lane = threadIdx.x & 31;
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, upd ? reinterpret_cast<unsigned long long>(e) : (0xFFFFFFFF00000000ull | lane));
      const int leader = __ffs(peers) - 1;
      const bool solo = peers == (1u << lane);
      if (__all_sync(0xFFFFFFFFu, solo)) {
        if (upd) static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value; jit_acc_global<G, J>(e, G::template acc<J>(rows[k])); });
      } else {
        static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
          AccVal v = G::template acc<J>(rows[k]);
          if (!upd) v.valid = false;
          jit_acc_prepare<G, J>(v);
          for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1) {      // same trip count for every lane of a peer group
            const int src = __ffs(rest) - 1;
            AccVal o;
            o.i = mk128(__shfl_sync(peers, i128_lo(v.i), src), __shfl_sync(peers, i128_hi(v.i), src));
            o.f = __shfl_sync(peers, v.f, src);
            o.valid = __shfl_sync(peers, (int)v.valid, src) != 0;
            if ((int)lane == leader) jit_acc_merge<G, J>(v, o);
          }
          if (upd && (int)lane == leader) jit_acc_global_n<G, J>(e, v);
        });
      }
    }
  }
}

// per-thread register partials of the integer fast path (tier 2)
template <class G> struct JitAggRegs { int64_t v[REG_GROUPS][G::N_ACCS > 0 ? G::N_ACCS : 1]; int rows; };

// Warp-collective (full-mask shuffles): nothing in here may depend on the dictionary size, which other warps change
// asynchronously -- every register group is flushed, groups that do not exist yet hold zeros.
template <class G>
__device__ __forceinline__ void jit_reg_flush(const JitHot& H, JitAggRegs<G>& R) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < REG_GROUPS; ++g) {
    {
      uint64_t* wa = H.wacc + (size_t)g * jit_aw<G>();
#pragma unroll
      for (int j = 0; j < G::N_ACCS; ++j) {
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
  R.rows = 0;
}

constexpr int JIT_REG_FLUSH = 224;

// tier 2: counts / integer / decimal sums, first REG_GROUPS groups in registers
template <class G>
__device__ __forceinline__ void jit_agg_reg_tile(const KernelArgs& K, const typename G::Row (&rows)[G::RPT], JitSmem* sm, uint8_t* scratch, JitAggRegs<G>& R) {
  const JitHot H = jit_hot<G>(scratch);
  int gid[G::RPT];
  uint32_t fpv[G::RPT];
  bool miss = false;
  // the lanes of a warp leave the mbarrier wait one by one and other warps grow the dictionary meanwhile: lane 0's reading is
  // broadcast so that the warp collectives below are executed by all lanes or by none
  const int n0 = __shfl_sync(0xFFFFFFFFu, (int)lds_acquire_u32(&sm->dict_n), 0);
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    gid[k] = -1; fpv[k] = 0;
    if (rows[k].live) {
      uint64_t kw[MAX_KEY_WORDS];
      G::key_words(rows[k], kw);
      fpv[k] = jit_fp(kw);
      gid[k] = jit_dict_lookup<REG_GROUPS>(H, n0, kw, fpv[k]);
      miss |= gid[k] < 0;
    }
  }
  if (n0 < REG_GROUPS && __any_sync(0xFFFFFFFFu, miss)) {
#pragma unroll
    for (int k = 0; k < G::RPT; ++k) {
      const bool want = rows[k].live && gid[k] < 0;
      if (__any_sync(0xFFFFFFFFu, want)) {
        uint64_t kw[MAX_KEY_WORDS];
        G::key_words(rows[k], kw);
        const int g = jit_dict_add<REG_GROUPS>(sm, H, want, kw, fpv[k]);
        if (want) gid[k] = g;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    if (rows[k].live && gid[k] >= 0) {
      int64_t val[G::N_ACCS > 0 ? G::N_ACCS : 1];
      static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
        constexpr int mode = G::reg_mode(J);
        if constexpr (mode == 0) val[J] = 1;
        else {
          const AccVal a = G::template acc<J>(rows[k]);
          if constexpr (mode == 3) val[J] = (int64_t)a.i;
          else {
            if (fits55(a.i)) val[J] = (int64_t)a.i;
            else {                                               // rare: exact value straight to the table entry
              val[J] = 0;
              uint64_t* e = jit_hot_entry<G>(K, H, gid[k]);
              if (e) atomic_add_i128(e + 2 + G::KEY_WORDS + G::acc_word(J), a.i);
            }
          }
        }
      });
      switch (gid[k]) {
        case 0:
#pragma unroll
          for (int j = 0; j < G::N_ACCS; ++j) R.v[0][j] += val[j];
          break;
        case 1:
#pragma unroll
          for (int j = 0; j < G::N_ACCS; ++j) R.v[1][j] += val[j];
          break;
        case 2:
#pragma unroll
          for (int j = 0; j < G::N_ACCS; ++j) R.v[2][j] += val[j];
          break;
        default:
#pragma unroll
          for (int j = 0; j < G::N_ACCS; ++j) R.v[3][j] += val[j];
      }
    }
  }
  R.rows += G::RPT;
  if (R.rows >= JIT_REG_FLUSH) jit_reg_flush<G>(H, R);
  bool anycold = false;
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) anycold |= rows[k].live && gid[k] < 0;
  if (__any_sync(0xFFFFFFFFu, anycold)) jit_cold_rows<G>(K, rows, gid);
}

// tier 1: any accumulator mix, up to HOT_G groups, per-warp accumulators in shared memory (warp-shuffle reductions)
template <class G>
__device__ __forceinline__ void jit_agg_dict_tile(const KernelArgs& K, const typename G::Row (&rows)[G::RPT], JitSmem* sm, uint8_t* scratch) {
  const JitHot H = jit_hot<G>(scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int gid[G::RPT];
  uint32_t fpv[G::RPT];
  bool miss = false;
  const int n0 = __shfl_sync(0xFFFFFFFFu, (int)lds_acquire_u32(&sm->dict_n), 0);      // warp-uniform (see jit_agg_reg_tile)
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    gid[k] = -1; fpv[k] = 0;
    if (rows[k].live) {
      uint64_t kw[MAX_KEY_WORDS];
      G::key_words(rows[k], kw);
      fpv[k] = jit_fp(kw);
      gid[k] = jit_dict_lookup<G::HOT_G>(H, n0, kw, fpv[k]);
      miss |= gid[k] < 0;
    }
  }
  if (n0 < G::HOT_G && __any_sync(0xFFFFFFFFu, miss)) {
#pragma unroll
    for (int k = 0; k < G::RPT; ++k) {
      const bool want = rows[k].live && gid[k] < 0;
      if (__any_sync(0xFFFFFFFFu, want)) {
        uint64_t kw[MAX_KEY_WORDS];
        G::key_words(rows[k], kw);
        const int g = jit_dict_add<G::HOT_G>(sm, H, want, kw, fpv[k]);
        if (want) gid[k] = g;
      }
    }
  }
  const int hot_n = __shfl_sync(0xFFFFFFFFu, (int)lds_acquire_u32(&sm->dict_n), 0);      // every gid of this warp is below it
  bool anyhot = false;
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) anyhot |= rows[k].live && gid[k] >= 0;
  if (__any_sync(0xFFFFFFFFu, anyhot)) {
    static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
      constexpr int op = G::acc_op(J);
      AccVal av[G::RPT];
      bool ok[G::RPT];
#pragma unroll
      for (int k = 0; k < G::RPT; ++k) {
        ok[k] = false;
        av[k].i = 0; av[k].f = 0.0; av[k].valid = false;
        if (rows[k].live && gid[k] >= 0) {
          av[k] = G::template acc<J>(rows[k]);
          ok[k] = av[k].valid;
          if constexpr (op == ACC_SUM_I128) {
            if (ok[k] && !fits55(av[k].i)) {
              uint64_t* e = jit_hot_entry<G>(K, H, gid[k]);
              if (e) { atomic_add_i128(e + 2 + G::KEY_WORDS + G::acc_word(J), av[k].i); if (G::acc_seen(J)) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), 1ull << J); }
              ok[k] = false;
            }
          }
        }
      }
      for (int g = 0; g < hot_n; ++g) {
        uint64_t* wa = H.wacc + ((size_t)(warp * G::HOT_G + g)) * jit_aw<G>();
        uint64_t* slot = wa + 1 + 2 * J;
        bool any = false;
#pragma unroll
        for (int k = 0; k < G::RPT; ++k) any |= ok[k] && gid[k] == g;
        if (__ballot_sync(0xFFFFFFFFu, any) == 0) continue;
        if constexpr (op == ACC_COUNT) {
          int cnt = 0;
#pragma unroll
          for (int k = 0; k < G::RPT; ++k) cnt += (ok[k] && gid[k] == g) ? 1 : 0;
          cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
          if (lane == 0) slot[0] += (uint64_t)cnt;
        } else if constexpr (op == ACC_SUM_I64 || op == ACC_SUM_I128) {
          int64_t part = 0;
#pragma unroll
          for (int k = 0; k < G::RPT; ++k) part += (ok[k] && gid[k] == g) ? (int64_t)av[k].i : 0;
          part = warp_sum_i64(part);
          if (lane == 0) {
            if constexpr (op == ACC_SUM_I64) slot[0] += (uint64_t)part;
            else { const uint64_t lo = slot[0] + (uint64_t)part; slot[1] += (uint64_t)(part >> 63) + (lo < slot[0] ? 1ull : 0ull); slot[0] = lo; }
          }
        } else if constexpr (op == ACC_SUM_F64) {
          double part = 0.0;
#pragma unroll
          for (int k = 0; k < G::RPT; ++k) part += (ok[k] && gid[k] == g) ? av[k].f : 0.0;
          part = warp_sum_f64(part);
          if (lane == 0) slot[0] = (uint64_t)__double_as_longlong(__longlong_as_double((long long)slot[0]) + part);
        } else {
          uint64_t w0 = acc_identity(op, 0), w1 = acc_identity(op, 1);
#pragma unroll
          for (int k = 0; k < G::RPT; ++k) {
            if (ok[k] && gid[k] == g) {
              constexpr bool isf = op == ACC_MIN_F64 || op == ACC_MAX_F64;
              const uint64_t v0 = isf ? (uint64_t)__double_as_longlong(av[k].f) : i128_lo(av[k].i);
              const uint64_t v1 = isf ? 0 : i128_hi(av[k].i);
              acc_combine_words(op, w0, w1, v0, v1);
            }
          }
#pragma unroll
          for (int dlt = 16; dlt; dlt >>= 1) {
            const uint64_t o0 = __shfl_xor_sync(0xFFFFFFFFu, w0, dlt), o1 = __shfl_xor_sync(0xFFFFFFFFu, w1, dlt);
            acc_combine_words(op, w0, w1, o0, o1);
          }
          if (lane == 0) { uint64_t a0 = slot[0], a1 = slot[1]; acc_combine_words(op, a0, a1, w0, w1); slot[0] = a0; slot[1] = a1; }
        }
        if (G::acc_seen(J) && lane == 0) wa[0] |= 1ull << J;
      }
    });
  }
  bool anycold = false;
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) anycold |= rows[k].live && gid[k] < 0;
  if (__any_sync(0xFFFFFFFFu, anycold)) jit_cold_rows<G>(K, rows, gid);
}

template <class G> __device__ __forceinline__ void jit_hot_init(uint8_t* scratch) {
  if constexpr (G::AGG_TIER > 0) {
    const JitHot H = jit_hot<G>(scratch);
    for (int i = threadIdx.x; i < JIT_NWARPS * G::HOT_G; i += NT) {
      uint64_t* wa = H.wacc + (size_t)i * jit_aw<G>();
      wa[0] = 0;
      static_for<0, G::N_ACCS>([&](auto Jc) { constexpr int J = decltype(Jc)::value; wa[1 + 2 * J] = acc_identity(G::acc_op(J), 0); wa[2 + 2 * J] = acc_identity(G::acc_op(J), 1); });
    }
  }
}

// end of kernel (after a CTA barrier): fold the per-warp accumulators of every hot group into the global table
template <class G> __device__ __forceinline__ void jit_hot_flush(const KernelArgs& K, JitSmem* sm, uint8_t* scratch) {
  if constexpr (G::AGG_TIER > 0) {
    const JitHot H = jit_hot<G>(scratch);
    const int n = (int)lds_acquire_u32(&sm->dict_n);
    for (int g = threadIdx.x; g < n; g += NT) jit_hot_entry<G>(K, H, g);
    __syncthreads();
    constexpr int per = G::N_ACCS + 1;
    for (int p = threadIdx.x; p < n * per; p += NT) {
      const int g = p / per, j = p % per;
      uint64_t* e = reinterpret_cast<uint64_t*>(H.entry[g]);
      if (!e) continue;
      if (j == G::N_ACCS) {
        uint64_t seen = 0;
        for (int w = 0; w < JIT_NWARPS; ++w) seen |= H.wacc[((size_t)(w * G::HOT_G + g)) * jit_aw<G>()];
        if (seen) atomicOr(reinterpret_cast<unsigned long long*>(e + 1), (unsigned long long)seen);
        continue;
      }
      const int op = G::acc_op(j);
      uint64_t w0 = acc_identity(op, 0), w1 = acc_identity(op, 1);
      if (op == ACC_SUM_I128) { w0 = 0; w1 = 0; }
      for (int w = 0; w < JIT_NWARPS; ++w) {
        const uint64_t* slot = H.wacc + ((size_t)(w * G::HOT_G + g)) * jit_aw<G>() + 1 + 2 * j;
        acc_combine_words(op, w0, w1, slot[0], slot[1]);
      }
      uint64_t* dst = e + 2 + G::KEY_WORDS + G::acc_word(j);
      switch (op) {
        case ACC_SUM_I64: case ACC_COUNT: if (w0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)w0); break;
        case ACC_SUM_I128: atomic_add_i128(dst, mk128(w0, w1)); break;
        case ACC_SUM_F64: atomicAdd(reinterpret_cast<double*>(dst), __longlong_as_double((long long)w0)); break;
        case ACC_MIN_I32: case ACC_MIN_I64: atomicMin(reinterpret_cast<long long*>(dst), (long long)w0); break;
        case ACC_MAX_I32: case ACC_MAX_I64: atomicMax(reinterpret_cast<long long*>(dst), (long long)w0); break;
        case ACC_MIN_I128: atomic_minmax_i128(dst, mk128(w0, w1), true); break;
        case ACC_MAX_I128: atomic_minmax_i128(dst, mk128(w0, w1), false); break;
        case ACC_MIN_F64: atomic_minmax_f64(dst, __longlong_as_double((long long)w0), true); break;
        case ACC_MAX_F64: atomic_minmax_f64(dst, __longlong_as_double((long long)w0), false); break;
        default: break;
      }
    }
  }
}

// ================================================================================================
// store / compact sinks
// ================================================================================================
template <class G>
__device__ __forceinline__ void jit_store_tile(const KernelArgs& K, const typename G::Row (&rows)[G::RPT], int64_t row0, int nrows) {
  const PipelineParams& P = K.P[0];
  const int lane = threadIdx.x & 31;
  static_for<0, G::N_OUT>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
#pragma unroll
    for (int k = 0; k < G::RPT; ++k) {
      const int r = threadIdx.x + k * NT;
      const bool in = r < nrows;
      if constexpr (G::out_width(J) != 0) { if (in) G::template store<J>(rows[k], P.out[J].data, row0 + r); }
      else {
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, in && G::template out_bool<J>(rows[k]));
        if (lane == 0 && (r - lane) < nrows) reinterpret_cast<uint32_t*>(P.out[J].data)[(row0 + r) >> 5] = bits;
      }
      if constexpr (G::out_nullable(J) != 0) {
        const uint32_t vb = __ballot_sync(0xFFFFFFFFu, in && G::template out_valid<J>(rows[k]));
        if (lane == 0 && (r - lane) < nrows) reinterpret_cast<uint32_t*>(P.out[J].valid_bytes)[(row0 + r) >> 5] = vb;
      }
    }
  });
}

// order-preserving compaction; all warps of the CTA work on the same tile (two CTA barriers)
template <class G>
__device__ __forceinline__ void jit_compact_tile(const KernelArgs& K, const typename G::Row (&rows)[G::RPT], JitSmem* sm, int64_t tile) {
  const PipelineParams& P = K.P[0];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool keep[G::RPT]; uint32_t before[G::RPT];
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) {
    keep[k] = rows[k].live;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, keep[k]);
    before[k] = __popc(b & ((1u << lane) - 1));
    if (lane == 0) sm->warp_sums[k * JIT_NWARPS + warp] = __popc(b);
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int n = G::RPT * JIT_NWARPS;
    uint32_t v = lane < n ? sm->warp_sums[lane] : 0, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane < n) sm->warp_sums[lane] = incl - v;
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long excl = 0;
    if (P.tile_offsets) {
      excl = P.tile_offsets[tile];
    } else if (tile > 0) {
      if (lane == 0) st_release_u64(P.tile_status + tile, (1ull << 62) | total);
      int64_t hi = tile - 1;
      uint32_t spins = 0;
      for (;;) {
        const int64_t t = hi - lane;
        unsigned long long sw = t >= 0 ? ld_acquire_u64(P.tile_status + t) : (2ull << 62);
        const unsigned flag = (unsigned)(sw >> 62);
        const unsigned ready = __ballot_sync(0xFFFFFFFFu, flag != 0);
        const unsigned prefix = __ballot_sync(0xFFFFFFFFu, flag == 2);
        const int first_prefix = prefix ? __ffs(prefix) - 1 : 32;
        const unsigned need = first_prefix >= 31 ? 0xFFFFFFFFu : ((2u << first_prefix) - 1);
        if ((ready & need) != need) { if (++spins > (1u << 24)) __trap(); continue; }
        unsigned long long vv = (lane <= first_prefix) ? (sw & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) vv += __shfl_xor_sync(0xFFFFFFFFu, vv, d);
        excl += vv;
        if (first_prefix < 32) break;
        hi -= 32;
      }
    }
    if (lane == 0) {
      sm->tile_base = excl;
      if (!P.tile_offsets) {
        st_release_u64(P.tile_status + tile, (2ull << 62) | (excl + total));
        if ((tile + 1) * (int64_t)G::TILE >= P.n_rows) *P.out_count = excl + total;
      }
    }
  }
  __syncthreads();
  const unsigned long long base = sm->tile_base;
  uint32_t wbase[G::RPT];
#pragma unroll
  for (int k = 0; k < G::RPT; ++k) wbase[k] = sm->warp_sums[k * JIT_NWARPS + warp];
  static_for<0, G::N_OUT>([&](auto Jc) { constexpr int J = decltype(Jc)::value;
#pragma unroll
    for (int k = 0; k < G::RPT; ++k) {
      if (!keep[k]) continue;
      const int64_t pos = (int64_t)(base + wbase[k] + before[k]);
      if constexpr (G::out_width(J) != 0) G::template store<J>(rows[k], P.out[J].data, pos);
      else P.out[J].data[pos] = G::template out_bool<J>(rows[k]) ? 1 : 0;
      if constexpr (G::out_nullable(J) != 0) P.out[J].valid_bytes[pos] = G::template out_valid<J>(rows[k]) ? 1 : 0;
    }
  });
  __syncthreads();      // warp_sums / tile_base are reused by the next tile
}

// ================================================================================================
// the kernel body
// ================================================================================================
template <class G>
__device__ __forceinline__ void jit_main(const KernelArgs& K) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  JitSmem* sm = reinterpret_cast<JitSmem*>(smem_raw);
  uint8_t* scratch = smem_raw + JIT_HDR;
  uint8_t* ring = scratch + G::SCRATCH_BYTES;
  const PipelineParams& P = K.P[0];
  constexpr int S = G::STAGES;
  constexpr int TILE = G::TILE;
  static_assert(S >= 2 && S <= JIT_MAX_STAGES, "stage ring");
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n_tiles = (P.n_rows + TILE - 1) / TILE;
  const bool dynamic = G::SINK == SINK_COMPACT && P.tile_offsets == nullptr;
  const int64_t n_pos = P.tile_list ? P.n_list : n_tiles;
  const bool guarded = G::SINK == SINK_AGG && K.aux[0].agg.deferred != nullptr;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], JIT_NWARPS); }
    fence_barrier_init();
    sm->dict_n = 0; sm->dict_lock = 0;
  }
  if constexpr (G::SINK == SINK_AGG) jit_hot_init<G>(scratch);
  __syncthreads();

  // ---- producer (thread 0): hands tiles to stages ------------------------------------------------
  int64_t next_pos = blockIdx.x;
  bool stopped = false;
  auto produce = [&](int j) {
    const int s = j % S;
    long long tn = -1;
    if (!stopped) {
      int64_t pos;
      if (dynamic) pos = (int64_t)atomicAdd(P.ticket, 1u);
      else { pos = next_pos; next_pos += gridDim.x; }
      if (pos >= (dynamic ? n_tiles : n_pos)) stopped = true;
      else if (guarded && *reinterpret_cast<volatile unsigned long long*>(K.aux[0].agg.n_groups) > K.aux[0].agg.group_limit) {
        // bounded table: hand the tiles this CTA still owns back to the host
        const AggParams& A = K.aux[0].agg;
        const int64_t cnt = (n_pos - pos + gridDim.x - 1) / gridDim.x;
        const unsigned long long base = atomicAdd(A.n_deferred, (unsigned long long)cnt);
        for (int64_t q = 0; q < cnt; ++q) {
          const int64_t pp = pos + q * gridDim.x;
          A.deferred[base + q] = P.tile_list ? P.tile_list[pp] : (uint32_t)pp;
        }
        stopped = true;
      } else tn = (dynamic || !P.tile_list) ? pos : (long long)P.tile_list[pos];
    }
    if (tn >= 0) {
      const int64_t row0 = tn * TILE;
      if (P.n_rows - row0 >= TILE) {
        sm->tile_no[s] = tn;
        fence_proxy_async();
        mbar_expect_tx(&sm->full[s], G::TX_BYTES);
        G::issue(ring + (size_t)s * G::STAGE_BYTES, K, row0, &sm->full[s]);
      } else {
        sm->tile_no[s] = tn | JIT_PARTIAL;
        mbar_arrive(&sm->full[s]);
      }
    } else {
      sm->tile_no[s] = -1;
      mbar_arrive(&sm->full[s]);
    }
  };
  if (tid == 0)
    for (int j = 0; j < S - 1; ++j) produce(j);

  JitAggRegs<G> R;
  if constexpr (G::SINK == SINK_AGG && G::AGG_TIER == 2) {
#pragma unroll
    for (int g = 0; g < REG_GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < G::N_ACCS; ++j) R.v[g][j] = 0;
    R.rows = 0;
  }

  for (int it = 0;; ++it) {
    const int s = it % S;
    if (tid == 0) {
      if (it >= 1) mbar_wait(&sm->empty[(it - 1) % S], (uint32_t)(((it - 1) / S) & 1));
      produce(it + S - 1);
    }
    __syncwarp();
    mbar_wait(&sm->full[s], (uint32_t)((it / S) & 1));
    __syncwarp();
    const long long tn = sm->tile_no[s];
    if (tn < 0) break;
    const int64_t tile = tn & ~JIT_PARTIAL;
    const uint8_t* stg = ring + (size_t)s * G::STAGE_BYTES;
    const int64_t row0 = tile * TILE;
    const int nrows = (int)min((int64_t)TILE, P.n_rows - row0);
    if (tn & JIT_PARTIAL) {       // the last tile of the batch: cooperative copy, zero-filled past the end
      __syncthreads();
      G::copy_partial(ring + (size_t)s * G::STAGE_BYTES, K, row0, nrows);
      __syncthreads();
    }
    typename G::Row rows[G::RPT];
#pragma unroll
    for (int k = 0; k < G::RPT; ++k) {
      const int r = tid + k * NT;
      G::eval(stg, r, r < nrows, K, rows[k]);
    }
    if constexpr (G::SINK == SINK_AGG) {
      if constexpr (G::AGG_TIER == 2) jit_agg_reg_tile<G>(K, rows, sm, scratch, R);
      else if constexpr (G::AGG_TIER == 1) jit_agg_dict_tile<G>(K, rows, sm, scratch);
      else {
        int gid[G::RPT];
#pragma unroll
        for (int k = 0; k < G::RPT; ++k) gid[k] = -1;
        jit_cold_rows<G>(K, rows, gid);
      }
    } else if constexpr (G::SINK == SINK_STORE) {
      jit_store_tile<G>(K, rows, row0, nrows);
    } else if constexpr (G::SINK == SINK_COMPACT) {
      jit_compact_tile<G>(K, rows, sm, tile);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm->empty[s]);
  }
  if constexpr (G::SINK == SINK_AGG) {
    if constexpr (G::AGG_TIER == 2) { const JitHot H = jit_hot<G>(scratch); jit_reg_flush<G>(H, R); }
    __syncthreads();
    jit_hot_flush<G>(K, sm, scratch);
  }
}

// cooperative copy of one column's partial tile (rows >= nrows zero-filled); called by all NT threads
__device__ __forceinline__ void jit_copy_col(uint8_t* dst, const uint8_t* src, uint32_t total, uint32_t valid, bool aligned16) {
  if (aligned16 && (valid & 15u) == 0) {
    for (uint32_t o = threadIdx.x * 16; o < total; o += NT * 16) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (o < valid) v = *reinterpret_cast<const uint4*>(src + o);
      *reinterpret_cast<uint4*>(dst + o) = v;
    }
  } else {
    for (uint32_t o = threadIdx.x; o < total; o += NT) dst[o] = o < valid ? src[o] : 0;
  }
}

}  // namespace sg
