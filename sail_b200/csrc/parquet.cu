// parquet.cu -- Parquet column chunks -> Arrow columns in HBM.
//
// The step before the operator path: Sail's scans are DataFusion `DataSourceExec(ParquetSource)` nodes
// (crates/sail-data-source/src/listing/planner.rs:47, adapted at crates/sail-execution/src/task_runner/core.rs:115-133)
// that decode pages to Arrow on CPU cores; an operator path at TB/s is then fed at the pace of that decode and of PCIe
// moving 16-byte decimals.  Here the column chunk crosses PCIe AS STORED (dictionary indices, 7-byte decimals, ...) and is
// decoded on the device.  The host only walks what is inherently sequential and tiny: Thrift page headers and the run
// headers of the RLE / bit-packed hybrid streams (a few bytes per run); every value is produced by a GPU thread.
//
// Covered: data pages V1 / V2, dictionary pages, PLAIN and RLE_DICTIONARY / PLAIN_DICTIONARY values, RLE definition
// levels of flat optional columns (max level 1), physical types INT32 / INT64 / DOUBLE / FLOAT / FIXED_LEN_BYTE_ARRAY /
// BYTE_ARRAY, uncompressed pages.  Compressed pages, nested columns and the DELTA_* encodings are refused
// (SAILGPU_ERR_UNSUPPORTED): the caller keeps the CPU reader for those files.
#include <cstring>

#include "device.hpp"
#include "h2d.hpp"
#include "kernels.hpp"

namespace sg {

namespace {

// ---- Thrift compact protocol (reader for the page header structs) ---------------------------------
struct TReader {
  const uint8_t* p; const uint8_t* end;
  int16_t last = 0;
  uint64_t varint() {
    uint64_t v = 0; int sh = 0;
    for (;;) {
      SG_CHECK(p < end && sh < 64, SAILGPU_ERR_INVALID, "parquet: truncated page header");
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << sh;
      if (!(b & 0x80)) return v;
      sh += 7;
    }
  }
  int64_t zigzag() { const uint64_t v = varint(); return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }
  // next field of the current struct: false at STOP
  bool field(int* id, int* type) {
    SG_CHECK(p < end, SAILGPU_ERR_INVALID, "parquet: truncated page header");
    const uint8_t b = *p++;
    if (b == 0) return false;
    *type = b & 0x0F;
    const int delta = b >> 4;
    last = delta ? (int16_t)(last + delta) : (int16_t)zigzag();
    *id = last;
    return true;
  }
  void skip(int type) {
    switch (type) {
      case 1: case 2: break;                       // booleans live in the field header
      case 3: ++p; break;
      case 4: case 5: case 6: (void)zigzag(); break;
      case 7: p += 8; break;
      case 8: { const uint64_t n = varint(); p += n; break; }
      case 9: case 10: {
        const uint8_t h = *p++;
        uint64_t n = h >> 4; const int et = h & 0x0F;
        if (n == 15) n = varint();
        for (uint64_t i = 0; i < n; ++i) skip(et);
        break;
      }
      case 12: { const int16_t save = last; last = 0; int id, t; while (field(&id, &t)) skip(t); last = save; break; }
      default: fail(SAILGPU_ERR_UNSUPPORTED, "parquet: page header field type " + std::to_string(type));
    }
    SG_CHECK(p <= end, SAILGPU_ERR_INVALID, "parquet: truncated page header");
  }
};

struct PageHeader {
  int type = -1, uncompressed = 0, compressed = 0;
  int num_values = 0, encoding = -1, def_encoding = 3;
  int num_nulls = -1, def_bytes = 0, rep_bytes = 0; bool v2_compressed = true;
};
enum { PAGE_DATA = 0, PAGE_DICT = 2, PAGE_DATA_V2 = 3 };
enum { ENC_PLAIN = 0, ENC_PLAIN_DICT = 2, ENC_RLE = 3, ENC_RLE_DICT = 8 };
enum { PT_BOOLEAN = 0, PT_INT32 = 1, PT_INT64 = 2, PT_INT96 = 3, PT_FLOAT = 4, PT_DOUBLE = 5, PT_BYTE_ARRAY = 6, PT_FLBA = 7 };

PageHeader read_page_header(TReader& r) {
  PageHeader h;
  r.last = 0;
  int id, t;
  while (r.field(&id, &t)) {
    if (id == 1 && t == 5) h.type = (int)r.zigzag();
    else if (id == 2 && t == 5) h.uncompressed = (int)r.zigzag();
    else if (id == 3 && t == 5) h.compressed = (int)r.zigzag();
    else if ((id == 5 || id == 7 || id == 8) && t == 12) {
      const int16_t save = r.last; r.last = 0;
      int fid, ft;
      while (r.field(&fid, &ft)) {
        if (id == 5) {
          if (fid == 1 && ft == 5) h.num_values = (int)r.zigzag();
          else if (fid == 2 && ft == 5) h.encoding = (int)r.zigzag();
          else if (fid == 3 && ft == 5) h.def_encoding = (int)r.zigzag();
          else r.skip(ft);
        } else if (id == 7) {
          if (fid == 1 && ft == 5) h.num_values = (int)r.zigzag();
          else if (fid == 2 && ft == 5) h.encoding = (int)r.zigzag();
          else r.skip(ft);
        } else {
          if (fid == 1 && ft == 5) h.num_values = (int)r.zigzag();
          else if (fid == 2 && ft == 5) h.num_nulls = (int)r.zigzag();
          else if (fid == 4 && ft == 5) h.encoding = (int)r.zigzag();
          else if (fid == 5 && ft == 5) h.def_bytes = (int)r.zigzag();
          else if (fid == 6 && ft == 5) h.rep_bytes = (int)r.zigzag();
          else if (fid == 7 && (ft == 1 || ft == 2)) h.v2_compressed = ft == 1;
          else r.skip(ft);
        }
      }
      r.last = save;
    } else r.skip(t);
  }
  return h;
}

// ---- RLE / bit-packed hybrid: the host reads the run headers, the device expands the runs ---------
struct Run {
  int64_t out_start;      // first output element
  int64_t src_bit;        // bit-packed: bit offset of the run's first value inside the chunk
  uint32_t count;
  uint32_t value;         // RLE: the repeated value
  uint32_t packed;        // 1: bit-packed
  uint32_t bit_width;
};

// walks a hybrid stream of `n_values` values; appends runs (clipped to n_values); returns the number of values equal to `count_value`
int64_t scan_hybrid(const uint8_t* chunk, const uint8_t* p, const uint8_t* end, int bit_width, int64_t n_values, int64_t out_start, std::vector<Run>* runs,
                    uint32_t count_value) {
  int64_t produced = 0, matches = 0;
  TReader r{p, end};
  const int vbytes = (bit_width + 7) / 8;
  while (produced < n_values) {
    SG_CHECK(r.p < end, SAILGPU_ERR_INVALID, "parquet: RLE stream ends before its page does");
    const uint64_t h = r.varint();
    if (h & 1) {
      const uint64_t groups = h >> 1;
      const int64_t cnt = std::min<int64_t>((int64_t)groups * 8, n_values - produced);
      Run run{out_start + produced, (int64_t)(r.p - chunk) * 8, (uint32_t)cnt, 0, 1, (uint32_t)bit_width};
      runs->push_back(run);
      if (bit_width == 1) {                      // definition levels: count the set bits (how many values the page carries)
        for (int64_t i = 0; i < cnt; ++i) matches += ((r.p[i >> 3] >> (i & 7)) & 1u) == count_value;
      }
      r.p += groups * (uint64_t)bit_width;
      SG_CHECK(r.p <= end, SAILGPU_ERR_INVALID, "parquet: bit-packed run overruns its page");
      produced += cnt;
    } else {
      const int64_t cnt = std::min<int64_t>((int64_t)(h >> 1), n_values - produced);
      uint32_t v = 0;
      SG_CHECK(r.p + vbytes <= end, SAILGPU_ERR_INVALID, "parquet: RLE run overruns its page");
      for (int b = 0; b < vbytes; ++b) v |= (uint32_t)r.p[b] << (8 * b);
      r.p += vbytes;
      if (cnt > 0) runs->push_back(Run{out_start + produced, 0, (uint32_t)cnt, v, 0, (uint32_t)bit_width});
      if (v == count_value) matches += cnt;
      produced += cnt;
    }
  }
  return matches;
}

__global__ void expand_runs_kernel(const uint8_t* __restrict__ chunk, const Run* __restrict__ runs, const int64_t* __restrict__ run_start, int n_runs, int64_t n,
                                   uint32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_runs - 1;                 // last run whose start <= i
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (run_start[mid] <= i) lo = mid; else hi = mid - 1; }
    const Run r = runs[lo];
    uint32_t v = r.value;
    if (r.packed) {
      const int64_t bit = r.src_bit + (i - r.out_start) * (int64_t)r.bit_width;
      uint64_t w = 0;
      const uint8_t* p = chunk + (bit >> 3);
#pragma unroll
      for (int b = 0; b < 5; ++b) w |= (uint64_t)p[b] << (8 * b);            // 32 value bits + 7 bits of offset fit in 5 bytes (buffers are padded)
      v = (uint32_t)((w >> (bit & 7)) & ((r.bit_width >= 32) ? 0xFFFFFFFFull : ((1ull << r.bit_width) - 1)));
    }
    out[i] = v;
  }
}

// ---- values -> Arrow ------------------------------------------------------------------------------------
struct Segment { int64_t dense_start; int64_t byte_off; int64_t is_dict; };   // one data page: dense index of its first value, where PLAIN values start, or dictionary-encoded
struct DecodeParams {
  const uint8_t* chunk;
  const uint32_t* valid;            // per row (1 = value present) or null
  const uint64_t* vpos;             // per row: dense value index (exclusive scan of valid) or null (= row)
  const uint32_t* dict_idx;         // per dense value: dictionary index, or null (PLAIN)
  const uint8_t* dict_vals;         // dictionary in Arrow layout (out_width bytes per entry)
  const Segment* segs; int n_segs;  // PLAIN pages
  const uint64_t* str_off;          // PLAIN BYTE_ARRAY: byte offset of every dense value's length prefix inside the chunk
  int64_t n_rows;
  int physical, type_length, out_width;   // out_width: 4, 8, 16 (Decimal128) or 16 with is_view
  int is_view;
  uint32_t dict_size;
  uint32_t* error;
};
__device__ __forceinline__ ulonglong2 make_view(const uint8_t* s, uint32_t len) {
  ulonglong2 v; v.x = len; v.y = 0;
  if (len <= 12) {
    for (uint32_t k = 0; k < len; ++k) { const unsigned long long b = s[k]; if (k < 4) v.x |= b << (32 + 8 * k); else v.y |= b << (8 * (k - 4)); }
  } else {
    for (uint32_t k = 0; k < 4; ++k) v.x |= (unsigned long long)s[k] << (32 + 8 * k);
    v.y = reinterpret_cast<unsigned long long>(s);
  }
  return v;
}
__device__ __forceinline__ void store_plain(const DecodeParams& D, const uint8_t* src, uint8_t* dst) {
  if (D.is_view) {
    uint32_t len; memcpy(&len, src, 4);
    *reinterpret_cast<ulonglong2*>(dst) = make_view(src + 4, len);
  } else if (D.physical == PT_FLBA) {           // big-endian two's complement, type_length bytes -> Decimal128
    unsigned __int128 v = (src[0] & 0x80) ? ~(unsigned __int128)0 : 0;
    for (int b = 0; b < D.type_length; ++b) v = (v << 8) | src[b];
    ulonglong2 w; w.x = (unsigned long long)v; w.y = (unsigned long long)(v >> 64);
    *reinterpret_cast<ulonglong2*>(dst) = w;
  } else if (D.physical == PT_INT32 || D.physical == PT_FLOAT) {
    int32_t x; memcpy(&x, src, 4);
    if (D.out_width == 16) { ulonglong2 w; w.x = (unsigned long long)(long long)x; w.y = (unsigned long long)((long long)x >> 63); *reinterpret_cast<ulonglong2*>(dst) = w; }
    else if (D.out_width == 8) *reinterpret_cast<long long*>(dst) = x;
    else *reinterpret_cast<int32_t*>(dst) = x;
  } else {                                      // INT64 / DOUBLE
    long long x; memcpy(&x, src, 8);
    if (D.out_width == 16) { ulonglong2 w; w.x = (unsigned long long)x; w.y = (unsigned long long)(x >> 63); *reinterpret_cast<ulonglong2*>(dst) = w; }
    else *reinterpret_cast<long long*>(dst) = x;
  }
}
__global__ void decode_values_kernel(DecodeParams D, uint8_t* __restrict__ out) {
  const int vw = D.physical == PT_FLBA ? D.type_length : (D.physical == PT_INT32 || D.physical == PT_FLOAT) ? 4 : 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < D.n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t* dst = out + i * D.out_width;
    if (D.valid && !D.valid[i]) {
      if (D.out_width == 16) { ulonglong2 z; z.x = 0; z.y = 0; *reinterpret_cast<ulonglong2*>(dst) = z; }
      else if (D.out_width == 8) *reinterpret_cast<long long*>(dst) = 0;
      else *reinterpret_cast<int32_t*>(dst) = 0;
      continue;
    }
    const int64_t v = D.vpos ? (int64_t)D.vpos[i] : i;
    int lo = 0, hi = D.n_segs - 1;               // the page this value came from
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (D.segs[mid].dense_start <= v) lo = mid; else hi = mid - 1; }
    if (D.segs[lo].is_dict) {
      const uint32_t k = D.dict_idx[v];
      if (k >= D.dict_size) { atomicOr(D.error, ERR_UNSUPPORTED); continue; }
      const uint8_t* s = D.dict_vals + (size_t)k * D.out_width;
      if (D.out_width == 16) *reinterpret_cast<ulonglong2*>(dst) = *reinterpret_cast<const ulonglong2*>(s);
      else if (D.out_width == 8) *reinterpret_cast<long long*>(dst) = *reinterpret_cast<const long long*>(s);
      else *reinterpret_cast<int32_t*>(dst) = *reinterpret_cast<const int32_t*>(s);
    } else if (D.is_view) {
      store_plain(D, D.chunk + D.str_off[v], dst);
    } else {
      store_plain(D, D.chunk + D.segs[lo].byte_off + (v - D.segs[lo].dense_start) * vw, dst);
    }
  }
}

BufPtr upload_vec(Ctx* ctx, const void* p, size_t bytes) {
  BufPtr b = dev_alloc(ctx, bytes);
  if (bytes) SG_CUDA(cudaMemcpyAsync(b->ptr, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return b;
}

}  // namespace

struct ParquetColumnDesc {      // mirrors sailgpu_parquet_column (include/sailgpu.h)
  const uint8_t* chunk; uint64_t chunk_len;
  int32_t physical_type, type_length, max_def_level, codec;
  int64_t num_values;
};

// Everything the host has to find out about a column chunk before the device can decode it: where the pages are, the run
// headers of their level / index streams, how many values each page really carries.  Pure host code (no device touched).
struct ColumnPlan {
  std::vector<Run> level_runs, index_runs;
  std::vector<Segment> segs;
  std::vector<uint64_t> str_off;
  int64_t dense = 0, dict_count = 0, dict_len = 0, n_pages = 0;
  const uint8_t* dict_bytes = nullptr;
  bool any_dict_page = false, any_plain_page = false, is_str = false;
  int out_width = 0;
};

ColumnPlan plan_parquet_column(const Field& f, const ParquetColumnDesc& c, int64_t n_rows) {
  ColumnPlan P;

  SG_CHECK(c.codec == 0, SAILGPU_ERR_UNSUPPORTED, "parquet: compressed pages are not decoded on the GPU path yet (column '" + f.name + "')");
  SG_CHECK(c.max_def_level == 0 || c.max_def_level == 1, SAILGPU_ERR_UNSUPPORTED, "parquet: nested / repeated columns are not supported (column '" + f.name + "')");
  SG_CHECK(c.num_values == n_rows, SAILGPU_ERR_INVALID, "parquet: column '" + f.name + "' has " + std::to_string(c.num_values) + " values for " + std::to_string(n_rows) + " rows");
  const int pt = c.physical_type;
  const bool is_str = f.type.is_string();
  P.is_str = is_str;
  SG_CHECK(pt == PT_INT32 || pt == PT_INT64 || pt == PT_DOUBLE || pt == PT_FLBA || pt == PT_BYTE_ARRAY, SAILGPU_ERR_UNSUPPORTED,
           "parquet: physical type " + std::to_string(pt) + " (column '" + f.name + "')");
  SG_CHECK((pt == PT_BYTE_ARRAY) == is_str, SAILGPU_ERR_UNSUPPORTED, "parquet: BYTE_ARRAY columns decode to strings only (column '" + f.name + "')");
  SG_CHECK(f.type.id != TypeId::Utf8, SAILGPU_ERR_UNSUPPORTED, "parquet: strings decode to Utf8View (what Sail reads Parquet strings as: application.yaml:375-381)");
  const int out_width = is_str ? 16 : f.type.arrow_width();
  P.out_width = out_width;
  SG_CHECK(out_width == 4 || out_width == 8 || out_width == 16, SAILGPU_ERR_UNSUPPORTED, "parquet: target type " + f.type.str());
  if (pt == PT_FLBA) SG_CHECK(f.type.is_decimal() && c.type_length >= 1 && c.type_length <= 16, SAILGPU_ERR_UNSUPPORTED, "parquet: FIXED_LEN_BYTE_ARRAY decodes to Decimal128 only");
  if (pt == PT_DOUBLE) SG_CHECK(f.type.id == TypeId::Float64, SAILGPU_ERR_UNSUPPORTED, "parquet: DOUBLE decodes to Float64");
  if (pt == PT_INT32) SG_CHECK(out_width == 4 || f.type.is_decimal(), SAILGPU_ERR_UNSUPPORTED, "parquet: INT32 decodes to 32-bit types or Decimal128");
  if (pt == PT_INT64) SG_CHECK(out_width == 8 || f.type.is_decimal(), SAILGPU_ERR_UNSUPPORTED, "parquet: INT64 decodes to 64-bit types or Decimal128");

  // ---- host: page headers and run headers -----------------------------------------------------------------------
  const uint8_t* base = c.chunk; const uint8_t* end = c.chunk + c.chunk_len;
  std::vector<Run>& level_runs = P.level_runs; std::vector<Run>& index_runs = P.index_runs;
  std::vector<Segment>& segs = P.segs;
  std::vector<uint64_t>& str_off = P.str_off;
  int64_t rows_done = 0; int64_t& dense_done = P.dense;
  bool have_dict = false; bool& any_dict_page = P.any_dict_page; bool& any_plain_page = P.any_plain_page;
  int64_t& dict_count = P.dict_count; const uint8_t*& dict_bytes = P.dict_bytes; int64_t& dict_len = P.dict_len;
  TReader r{base, end};
  while (r.p < end && rows_done < n_rows) {
    const PageHeader h = read_page_header(r);
    SG_CHECK(h.compressed == h.uncompressed, SAILGPU_ERR_UNSUPPORTED, "parquet: compressed page in column '" + f.name + "'");
    const uint8_t* body = r.p; const uint8_t* body_end = body + h.compressed;
    SG_CHECK(body_end <= end, SAILGPU_ERR_INVALID, "parquet: page overruns its column chunk");
    r.p = body_end;
    if (h.type == PAGE_DICT) {
      SG_CHECK(h.encoding == ENC_PLAIN || h.encoding == ENC_PLAIN_DICT, SAILGPU_ERR_UNSUPPORTED, "parquet: dictionary page encoding " + std::to_string(h.encoding));
      have_dict = true; dict_count = h.num_values; dict_bytes = body; dict_len = h.compressed;
      continue;
    }
    if (h.type != PAGE_DATA && h.type != PAGE_DATA_V2) continue;       // index pages etc.
    P.n_pages++;
    const int64_t nv = h.num_values;
    const uint8_t* vals = body;
    int64_t non_null = nv;
    if (h.type == PAGE_DATA_V2) {
      SG_CHECK(h.rep_bytes == 0, SAILGPU_ERR_UNSUPPORTED, "parquet: repetition levels");
      if (c.max_def_level > 0) non_null = scan_hybrid(base, body, body + h.def_bytes, 1, nv, rows_done, &level_runs, 1);
      vals = body + h.def_bytes;
    } else if (c.max_def_level > 0) {
      SG_CHECK(h.def_encoding == ENC_RLE, SAILGPU_ERR_UNSUPPORTED, "parquet: definition levels must be RLE encoded");
      uint32_t len; memcpy(&len, body, 4);
      SG_CHECK(body + 4 + len <= body_end, SAILGPU_ERR_INVALID, "parquet: definition levels overrun their page");
      non_null = scan_hybrid(base, body + 4, body + 4 + len, 1, nv, rows_done, &level_runs, 1);
      vals = body + 4 + len;
    }
    if (h.encoding == ENC_RLE_DICT || h.encoding == ENC_PLAIN_DICT) {
      SG_CHECK(have_dict, SAILGPU_ERR_INVALID, "parquet: dictionary-encoded page without a dictionary page");
      any_dict_page = true;
      segs.push_back(Segment{dense_done, 0, 1});
      SG_CHECK(vals < body_end || non_null == 0, SAILGPU_ERR_INVALID, "parquet: empty dictionary-index stream");
      if (non_null > 0) {
        const int bw = vals[0];
        SG_CHECK(bw <= 32, SAILGPU_ERR_INVALID, "parquet: dictionary index width " + std::to_string(bw));
        if (bw == 0) index_runs.push_back(Run{dense_done, 0, (uint32_t)non_null, 0, 0, 0});
        else scan_hybrid(base, vals + 1, body_end, bw, non_null, dense_done, &index_runs, 0xFFFFFFFFu);
      }
    } else if (h.encoding == ENC_PLAIN) {
      any_plain_page = true;
      segs.push_back(Segment{dense_done, (int64_t)(vals - base), 0});
      if (is_str) {
        // offsets are indexed by dense value number: values of earlier DICTIONARY pages get a zero (never read) -- filled
        // only now, so that an all-dictionary chunk keeps no per-value host state at all
        str_off.resize((size_t)dense_done, 0);
        const uint8_t* q = vals;
        for (int64_t i = 0; i < non_null; ++i) {
          SG_CHECK(q + 4 <= body_end, SAILGPU_ERR_INVALID, "parquet: BYTE_ARRAY value overruns its page");
          uint32_t len; memcpy(&len, q, 4);
          str_off.push_back((uint64_t)(q - base));
          q += 4 + (size_t)len;
        }
        SG_CHECK(q <= body_end, SAILGPU_ERR_INVALID, "parquet: BYTE_ARRAY value overruns its page");
      }
    } else fail(SAILGPU_ERR_UNSUPPORTED, "parquet: value encoding " + std::to_string(h.encoding) + " (column '" + f.name + "')");
    rows_done += nv; dense_done += non_null;
  }
  SG_CHECK(rows_done == n_rows, SAILGPU_ERR_INVALID, "parquet: pages of column '" + f.name + "' hold " + std::to_string(rows_done) + " values, expected " + std::to_string(n_rows));
  // (a writer that outgrows its dictionary falls back to PLAIN pages mid-chunk: every page carries its own kind)

  return P;
}

DevColumn decode_parquet_column(Ctx* ctx, const Field& f, const ParquetColumnDesc& c, int64_t n_rows) {
  ColumnPlan P = plan_parquet_column(f, c, n_rows);
  const uint8_t* base = c.chunk;
  const int pt = c.physical_type;
  const bool is_str = P.is_str;
  const int out_width = P.out_width;
  std::vector<Run>& level_runs = P.level_runs; std::vector<Run>& index_runs = P.index_runs;
  std::vector<Segment>& segs = P.segs;
  std::vector<uint64_t>& str_off = P.str_off;
  const int64_t dense_done = P.dense, dict_count = P.dict_count, dict_len = P.dict_len;
  const uint8_t* dict_bytes = P.dict_bytes;
  const bool any_dict_page = P.any_dict_page, any_plain_page = P.any_plain_page;
  // ---- device ---------------------------------------------------------------------------------------------------------
  DevColumn col; col.type = f.type; col.length = n_rows;
  BufPtr dchunk = dev_alloc(ctx, (size_t)c.chunk_len + 64);
  {
    HostStager st(ctx);
    st.add_raw(dchunk->ptr, c.chunk, (size_t)c.chunk_len);
    st.flush();
  }
  const uint8_t* dbase = static_cast<const uint8_t*>(dchunk->ptr);
  BufPtr err = dev_alloc_zero(ctx, 8);
  auto expand = [&](const std::vector<Run>& runs, int64_t n) -> BufPtr {
    BufPtr out = dev_alloc(ctx, (size_t)n * 4 + 16);
    if (n == 0) return out;
    std::vector<int64_t> starts(runs.size());
    for (size_t i = 0; i < runs.size(); ++i) starts[i] = runs[i].out_start;
    BufPtr druns = upload_vec(ctx, runs.data(), runs.size() * sizeof(Run)), dstarts = upload_vec(ctx, starts.data(), starts.size() * 8);
    expand_runs_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(dbase, static_cast<const Run*>(druns->ptr), static_cast<const int64_t*>(dstarts->ptr),
                                                                                                (int)runs.size(), n, static_cast<uint32_t*>(out->ptr));
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(ctx->stream));      // `starts` / `runs` are host vectors
    return out;
  };
  DecodeParams D; memset(&D, 0, sizeof(D));
  D.chunk = dbase; D.n_rows = n_rows; D.physical = pt; D.type_length = c.type_length; D.out_width = out_width; D.is_view = is_str ? 1 : 0;
  D.error = static_cast<uint32_t*>(err->ptr);
  BufPtr valid, vpos, scratch, didx, ddict, dsegs, dstr;
  if (c.max_def_level > 0) {
    valid = expand(level_runs, n_rows);
    vpos = dev_alloc(ctx, (size_t)(n_rows + 1) * 8);
    scratch = dev_alloc(ctx, 1026 * 8);
    SG_CUDA(launch_exclusive_scan_u32(static_cast<const uint32_t*>(valid->ptr), n_rows, static_cast<uint64_t*>(vpos->ptr), static_cast<uint64_t*>(scratch->ptr), ctx->stream));
    D.valid = static_cast<const uint32_t*>(valid->ptr); D.vpos = static_cast<const uint64_t*>(vpos->ptr);
  }
  if (any_dict_page) {
    // the dictionary itself is a PLAIN page: decode it into the Arrow layout with the same kernel
    DecodeParams Q; memset(&Q, 0, sizeof(Q));
    Q.chunk = dbase; Q.n_rows = dict_count; Q.physical = pt; Q.type_length = c.type_length; Q.out_width = out_width; Q.is_view = D.is_view; Q.error = D.error;
    std::vector<Segment> ds{Segment{0, (int64_t)(dict_bytes - base), 0}};
    std::vector<uint64_t> doff;
    BufPtr qsegs, qoff;
    if (is_str) {
      const uint8_t* q = dict_bytes;
      for (int64_t i = 0; i < dict_count; ++i) {
        SG_CHECK(q + 4 <= dict_bytes + dict_len, SAILGPU_ERR_INVALID, "parquet: dictionary value overruns its page");
        uint32_t len; memcpy(&len, q, 4);
        doff.push_back((uint64_t)(q - base));
        q += 4 + (size_t)len;
      }
      qoff = upload_vec(ctx, doff.data(), doff.size() * 8);
      Q.str_off = static_cast<const uint64_t*>(qoff->ptr);
    }
    qsegs = upload_vec(ctx, ds.data(), sizeof(Segment));
    Q.segs = static_cast<const Segment*>(qsegs->ptr); Q.n_segs = 1;
    ddict = dev_alloc(ctx, (size_t)std::max<int64_t>(dict_count, 1) * out_width);
    if (dict_count) decode_values_kernel<<<(int)std::min<int64_t>((dict_count + 255) / 256, 148 * 4), 256, 0, ctx->stream>>>(Q, static_cast<uint8_t*>(ddict->ptr));
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    didx = expand(index_runs, dense_done);
    D.dict_idx = static_cast<const uint32_t*>(didx->ptr); D.dict_vals = static_cast<const uint8_t*>(ddict->ptr); D.dict_size = (uint32_t)dict_count;
  }
  if (is_str && any_plain_page) {
    str_off.resize((size_t)dense_done, 0);
    dstr = upload_vec(ctx, str_off.data(), str_off.size() * 8);
    D.str_off = static_cast<const uint64_t*>(dstr->ptr);
  }
  if (segs.empty()) segs.push_back(Segment{0, 0, 0});
  dsegs = upload_vec(ctx, segs.data(), segs.size() * sizeof(Segment));
  D.segs = static_cast<const Segment*>(dsegs->ptr); D.n_segs = (int)segs.size();
  col.data = dev_alloc(ctx, (size_t)n_rows * out_width);
  if (n_rows) decode_values_kernel<<<(int)std::min<int64_t>((n_rows + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(D, static_cast<uint8_t*>(col.data->ptr));
  SG_CUDA(cudaGetLastError());
  if (is_str) col.heaps = {dchunk};                 // long views point into the chunk bytes
  col.null_count = 0;
  if (c.max_def_level > 0 && dense_done < n_rows) {
    // validity bitmap from the expanded levels (u32 0/1 -> bytes -> bits)
    BufPtr bytes = dev_alloc(ctx, (size_t)n_rows + 4);
    SG_CUDA(launch_u32_to_bytes(static_cast<const uint32_t*>(valid->ptr), static_cast<uint8_t*>(bytes->ptr), n_rows, ctx->stream));
    col.validity = dev_alloc_zero(ctx, (size_t)((n_rows + 31) / 32 * 4));
    SG_CUDA(launch_pack_bytes(static_cast<const uint8_t*>(bytes->ptr), static_cast<uint32_t*>(col.validity->ptr), n_rows, nullptr, ctx->stream));
    col.null_count = n_rows - dense_done;
  }
  uint32_t e = 0;
  SG_CUDA(cudaMemcpyAsync(&e, err->ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(cudaStreamSynchronize(ctx->stream));       // host vectors above; error flag
  SG_CHECK(e == 0, SAILGPU_ERR_INVALID, "parquet: dictionary index out of range in column '" + f.name + "'");
  return col;
}

// host-only summary of a column plan (tests/test_parquet_plan.py pins the page / run walking against pyarrow's own metadata)
std::string parquet_plan_summary(const Field& f, const ParquetColumnDesc& c, int64_t n_rows) {
  ColumnPlan P = plan_parquet_column(f, c, n_rows);
  int64_t level_vals = 0, index_vals = 0;
  for (auto& r : P.level_runs) level_vals += r.count;
  for (auto& r : P.index_runs) index_vals += r.count;
  char b[512];
  snprintf(b, sizeof b, "{\"pages\":%lld,\"dense\":%lld,\"dict_count\":%lld,\"level_values\":%lld,\"index_values\":%lld,\"level_runs\":%zu,\"index_runs\":%zu,"
                        "\"plain_strings\":%zu,\"dict_pages\":%d,\"plain_pages\":%d}",
           (long long)P.n_pages, (long long)P.dense, (long long)P.dict_count, (long long)level_vals, (long long)index_vals, P.level_runs.size(), P.index_runs.size(),
           P.str_off.size(), P.any_dict_page ? 1 : 0, P.any_plain_page ? 1 : 0);
  return b;
}

}  // namespace sg
