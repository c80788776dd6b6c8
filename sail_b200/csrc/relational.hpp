// relational.hpp -- parameter blocks and launchers of relational.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vm.h"

namespace sg {

struct RawKeyCol {
  const uint8_t* data;
  const uint8_t* validity_bits;   // Arrow bitmap or null
  int32_t width;                  // bytes compared / hashed: 1, 4, 8, 16
  int32_t is_view;
  int32_t stride;                 // bytes between rows (16 for a <=18-digit decimal keyed on its low word)
  int32_t pad;
};

struct JoinMultiParams {
  int64_t n_probe;
  int32_t n_keys;
  int32_t pass;                   // 0 count, 1 emit
  int32_t emit_unmatched_probe;   // right-outer: a probe row without match yields (-1, row)
  int32_t pad;
  RawKeyCol probe_keys[MAX_KEYS];
  RawKeyCol build_keys[MAX_KEYS];
  const uint8_t* table;
  uint64_t capacity_mask;
  uint32_t* counts;
  const uint64_t* offsets;
  int64_t* out_build;
  int64_t* out_probe;
  uint8_t* visited;
  const int64_t* next;            // duplicate-key chains built by the build sink
};

enum SortKind : int32_t { SORT_INT = 0, SORT_UINT = 1, SORT_F64 = 2, SORT_BOOL = 3, SORT_VIEW = 4 };
struct SortKeyCol {
  const uint8_t* data;
  const uint8_t* validity_bits;
  int32_t kind, width;
  int32_t asc, nulls_first;
  int32_t out_off;                // offset of this key inside the encoded row
  int32_t enc_bytes;              // value bytes after the null byte
  int32_t str_len;                // SORT_VIEW: padded string bytes (max length in the column)
  int32_t pad;
};
struct SortEncodeParams {
  int64_t n;
  int32_t n_keys, key_bytes;
  uint8_t* keys;
  uint32_t* bits;                 // [2 * key_bytes], zeroed: per byte position the OR of the bytes, then the OR of their complements
  SortKeyCol cols[8];
};
struct RadixScratch {
  uint32_t *idx_a, *idx_b;        // [n]
  uint64_t *kw_a, *kw_b;          // [n]
  uint32_t* hist;                 // [256 * ceil(n / 2048)]
  uint64_t* offs;                 // [256 * ceil(n / 2048)]
  uint64_t* scan_scratch;         // [1026]
};

cudaError_t launch_join_multi(const JoinMultiParams& P, cudaStream_t s);
cudaError_t launch_gather_bits(const uint8_t* bits, uint8_t* out, const int64_t* idx, int64_t n, int dflt, cudaStream_t s);
cudaError_t launch_max_view_len(const void* views, int64_t n, unsigned int* out, cudaStream_t s);
cudaError_t launch_sort_encode(const SortEncodeParams& P, cudaStream_t s);
cudaError_t radix_sort_indices(const uint8_t* keys, int key_bytes, int64_t n, const RadixScratch& S, const uint32_t* bits, cudaStream_t s, int* launches);
cudaError_t launch_topk_hist(const uint8_t* keys, int key_bytes, int64_t n, int used, uint64_t prefix, int digit_bits, uint32_t* hist /* [2048], zeroed */, cudaStream_t s);
cudaError_t launch_topk_compact(const uint8_t* keys, int key_bytes, int64_t n, int used, uint64_t threshold, int64_t* out, unsigned long long* counter, cudaStream_t s);
cudaError_t launch_merge_rank(const uint8_t* keys, int key_bytes, const int64_t* run_off, int n_runs, int64_t n, int64_t* perm, cudaStream_t s);
cudaError_t launch_key_hash(const uint8_t* keys, int key_bytes, int64_t n, uint8_t* out8, cudaStream_t s);
cudaError_t launch_group_heads(const uint8_t* keys, int key_bytes, const uint32_t* idx, int64_t n, uint32_t* heads, const uint8_t* hashes,
                               unsigned long long* collisions, cudaStream_t s);
cudaError_t launch_assign_groups(const uint32_t* idx, const uint32_t* heads, const uint64_t* before, int64_t n, int64_t* gid_of_row, int64_t* rep, cudaStream_t s);
cudaError_t launch_iota(int64_t* out, int64_t n, cudaStream_t s);
cudaError_t launch_iota_stride(int64_t* out, int64_t first, int64_t stride, int64_t n, cudaStream_t s);
cudaError_t launch_widen_u32(const uint32_t* in, int64_t* out, int64_t n, cudaStream_t s);
// many small device-to-device copies in one launch (packed exchange messages)
struct CopySeg { const uint8_t* src; uint8_t* dst; unsigned long long bytes; };
cudaError_t launch_multi_copy_raw(const CopySeg* dev_segs, int n, cudaStream_t s);
cudaError_t launch_rebase_views(void* views, int64_t n, uint64_t heap_base, cudaStream_t s);

}  // namespace sg
