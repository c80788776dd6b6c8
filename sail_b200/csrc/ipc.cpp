// ipc.cpp -- result sink: a host Arrow batch as an Arrow IPC stream (host-only C++, no CUDA).
//
// Sail hands every result batch to the Spark Connect client as one self-contained IPC stream
// (crates/sail-spark-connect/src/executor.rs:320-330 `to_arrow_batch`: StreamWriter::try_new(schema) + write(batch) +
// finish()).  This is that framing for the batches libsailgpu pulls to the host: Schema message, one RecordBatch message,
// end-of-stream marker.  The flatbuffer metadata (format/Message.fbs, format/Schema.fbs of the Arrow columnar format,
// MetadataVersion V5) is written by hand -- forward, parent before child, offsets patched -- there is no flatbuffers
// dependency.  Types: the ones the operators produce (include/sailgpu.h "Types T") plus Binary / LargeUtf8 / BinaryView.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sailgpu.h"

namespace {

struct IpcError { int code; std::string msg; };
[[noreturn]] void fail(int code, const std::string& m) { throw IpcError{code, m}; }

// ---- a forward flatbuffer writer ---------------------------------------------------------------------------------------
// Offsets in a flatbuffer (uoffset_t) point forward from the field that holds them, so a parent may be written before its
// children if the field is patched once the child's position is known.  A table is [soffset to vtable][fields]; its vtable
// ([u16 vtable bytes][u16 table bytes][u16 field offset per slot]) is written immediately before it.
struct Fb {
  std::vector<uint8_t> b;
  size_t size() const { return b.size(); }
  void pad_to(size_t a) { while (b.size() % a) b.push_back(0); }
  template <class T> void put(T v) { const size_t n = b.size(); b.resize(n + sizeof(T)); memcpy(&b[n], &v, sizeof(T)); }
  template <class T> void set(size_t at, T v) { memcpy(&b[at], &v, sizeof(T)); }
  void point(size_t field_at, size_t target) { set<uint32_t>(field_at, (uint32_t)(target - field_at)); }

  struct FieldDef { int slot; int size; uint64_t value; };   // size 1/2/4/8 scalar (value = bits) or 0 = offset (patched later)
  // writes vtable + table; returns the table position and, per definition, the position of its field
  size_t table(const std::vector<FieldDef>& defs, std::vector<size_t>* field_at) {
    int n_slots = 0;
    for (auto& d : defs) n_slots = d.slot + 1 > n_slots ? d.slot + 1 : n_slots;
    // inline layout: soffset (4 bytes) then fields, largest first, each naturally aligned relative to the table start
    std::vector<int> order(defs.size());
    for (size_t i = 0; i < defs.size(); ++i) order[i] = (int)i;
    auto bytes = [&](int i) { return defs[(size_t)i].size == 0 ? 4 : defs[(size_t)i].size; };
    for (size_t i = 0; i < order.size(); ++i)
      for (size_t j = i + 1; j < order.size(); ++j)
        if (bytes(order[j]) > bytes(order[i])) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    std::vector<int> off(defs.size());
    int cur = 4;
    for (int i : order) { const int w = bytes(i); cur = (cur + w - 1) / w * w; off[(size_t)i] = cur; cur += w; }
    const int table_bytes = cur;
    const int vt_bytes = 4 + 2 * n_slots;
    // the table must start 8-aligned (it may hold 8-byte scalars); the vtable sits right before it
    while ((b.size() + (size_t)vt_bytes) % 8) b.push_back(0);
    const size_t vt_at = b.size();
    put<uint16_t>((uint16_t)vt_bytes);
    put<uint16_t>((uint16_t)table_bytes);
    std::vector<uint16_t> slots((size_t)n_slots, 0);
    for (size_t i = 0; i < defs.size(); ++i) slots[(size_t)defs[i].slot] = (uint16_t)off[i];
    for (uint16_t s : slots) put<uint16_t>(s);
    const size_t t_at = b.size();
    b.resize(t_at + (size_t)table_bytes, 0);
    set<int32_t>(t_at, (int32_t)(t_at - vt_at));
    if (field_at) field_at->assign(defs.size(), 0);
    for (size_t i = 0; i < defs.size(); ++i) {
      const size_t at = t_at + (size_t)off[i];
      switch (defs[i].size) {
        case 1: set<uint8_t>(at, (uint8_t)defs[i].value); break;
        case 2: set<uint16_t>(at, (uint16_t)defs[i].value); break;
        case 4: set<uint32_t>(at, (uint32_t)defs[i].value); break;
        case 8: set<uint64_t>(at, defs[i].value); break;
        default: break;
      }
      if (field_at) (*field_at)[i] = at;
    }
    return t_at;
  }
  size_t string(const std::string& s) {
    pad_to(4);
    const size_t at = b.size();
    put<uint32_t>((uint32_t)s.size());
    b.insert(b.end(), s.begin(), s.end());
    b.push_back(0);
    return at;
  }
  // vector of n uoffsets (patched by the caller); returns the position of element 0
  size_t offset_vector(size_t n, size_t* vec_at) {
    pad_to(4);
    *vec_at = b.size();
    put<uint32_t>((uint32_t)n);
    const size_t first = b.size();
    b.resize(first + 4 * n, 0);
    return first;
  }
  // vector of 8-byte-aligned elements (int64 or structs of int64): elements 8-aligned, the count right before them
  size_t vector64(const std::vector<int64_t>& words, size_t n_elems) {
    while ((b.size() + 4) % 8) b.push_back(0);
    const size_t at = b.size();
    put<uint32_t>((uint32_t)n_elems);
    for (int64_t w : words) put<int64_t>(w);
    return at;
  }
};

// ---- Arrow C Data Interface -> flatbuffer type ---------------------------------------------------------------------------
enum TypeTag : uint8_t { T_Int = 2, T_FloatingPoint = 3, T_Binary = 4, T_Utf8 = 5, T_Bool = 6, T_Decimal = 7, T_Date = 8, T_LargeBinary = 19, T_LargeUtf8 = 20,
                         T_BinaryView = 23, T_Utf8View = 24 };
enum Layout { L_FIXED, L_BOOL, L_VARLEN32, L_VARLEN64, L_VIEW };
struct ColType { TypeTag tag; Layout layout; int width = 0; int bits = 0; bool is_signed = false; int precision = 0, scale = 0; };

ColType parse_format(const char* f) {
  const std::string s = f ? f : "";
  ColType t{};
  auto integer = [&](int bits, bool sg) { t.tag = T_Int; t.layout = L_FIXED; t.width = bits / 8; t.bits = bits; t.is_signed = sg; return t; };
  if (s == "b") { t.tag = T_Bool; t.layout = L_BOOL; return t; }
  if (s == "c") return integer(8, true);
  if (s == "C") return integer(8, false);
  if (s == "s") return integer(16, true);
  if (s == "S") return integer(16, false);
  if (s == "i") return integer(32, true);
  if (s == "I") return integer(32, false);
  if (s == "l") return integer(64, true);
  if (s == "L") return integer(64, false);
  if (s == "f") { t.tag = T_FloatingPoint; t.layout = L_FIXED; t.width = 4; t.bits = 32; return t; }
  if (s == "g") { t.tag = T_FloatingPoint; t.layout = L_FIXED; t.width = 8; t.bits = 64; return t; }
  if (s == "tdD") { t.tag = T_Date; t.layout = L_FIXED; t.width = 4; return t; }
  if (s == "u") { t.tag = T_Utf8; t.layout = L_VARLEN32; return t; }
  if (s == "z") { t.tag = T_Binary; t.layout = L_VARLEN32; return t; }
  if (s == "U") { t.tag = T_LargeUtf8; t.layout = L_VARLEN64; return t; }
  if (s == "Z") { t.tag = T_LargeBinary; t.layout = L_VARLEN64; return t; }
  if (s == "vu") { t.tag = T_Utf8View; t.layout = L_VIEW; t.width = 16; return t; }
  if (s == "vz") { t.tag = T_BinaryView; t.layout = L_VIEW; t.width = 16; return t; }
  if (s.rfind("d:", 0) == 0) {
    int p = 0, sc = 0, bw = 128;
    const int n = sscanf(s.c_str() + 2, "%d,%d,%d", &p, &sc, &bw);
    if (n >= 2 && bw == 128) { t.tag = T_Decimal; t.layout = L_FIXED; t.width = 16; t.precision = p; t.scale = sc; return t; }
  }
  fail(SAILGPU_ERR_UNSUPPORTED, "ipc: column format '" + s + "' has no IPC encoding here");
}

size_t write_type(Fb& fb, const ColType& t) {
  switch (t.tag) {
    case T_Int: return fb.table({{0, 4, (uint64_t)t.bits}, {1, 1, t.is_signed ? 1u : 0u}}, nullptr);
    case T_FloatingPoint: return fb.table({{0, 2, t.bits == 32 ? 1u : 2u}}, nullptr);                 // Precision: SINGLE = 1, DOUBLE = 2
    case T_Decimal: return fb.table({{0, 4, (uint64_t)t.precision}, {1, 4, (uint64_t)t.scale}, {2, 4, 128u}}, nullptr);
    case T_Date: return fb.table({{0, 2, 0u}}, nullptr);                                              // DateUnit DAY = 0 (the default is MILLISECOND)
    default: return fb.table({}, nullptr);                                                            // Utf8, Binary, Bool, views: no members
  }
}

// Field { name:0, nullable:1, type_type:2, type:3, dictionary:4, children:5 } -- written after its parent vector, patched in
void write_field(Fb& fb, size_t slot_at, const ArrowSchema* f, const ColType& t) {
  std::vector<size_t> at;
  const size_t field = fb.table({{0, 0, 0}, {1, 1, (f->flags & ARROW_FLAG_NULLABLE) ? 1u : 0u}, {2, 1, (uint64_t)t.tag}, {3, 0, 0}, {5, 0, 0}}, &at);
  fb.point(slot_at, field);
  fb.point(at[0], fb.string(f->name ? f->name : ""));
  fb.point(at[3], write_type(fb, t));
  size_t vec;
  fb.offset_vector(0, &vec);                    // readers insist on a children vector, empty for flat types
  fb.point(at[4], vec);
}

// wraps a finished flatbuffer as an encapsulated message: continuation marker, metadata length (padded so that the body starts
// 8-aligned), metadata
void append_message(std::vector<uint8_t>& out, Fb& meta) {
  meta.pad_to(8);
  const uint32_t marker = 0xFFFFFFFFu;
  const int32_t len = (int32_t)meta.size();
  const size_t n = out.size();
  out.resize(n + 8 + meta.size());
  memcpy(&out[n], &marker, 4);
  memcpy(&out[n + 4], &len, 4);
  memcpy(&out[n + 8], meta.b.data(), meta.size());
}

// Message { version:0 (V5 = 4), header_type:1, header:2, bodyLength:3 }; the root uoffset comes first
struct MessageHead { Fb fb; size_t header_field = 0; };
MessageHead begin_message(uint8_t header_type, int64_t body_length) {
  MessageHead m;
  m.fb.put<uint32_t>(0);                        // root offset, patched
  std::vector<size_t> at;
  const size_t msg = m.fb.table({{0, 2, 4u}, {1, 1, header_type}, {2, 0, 0}, {3, 8, (uint64_t)body_length}}, &at);
  m.fb.point(0, msg);
  m.header_field = at[2];
  return m;
}

void check_schema(const ArrowSchema* s) {
  if (!s || !s->format || std::string(s->format) != "+s") fail(SAILGPU_ERR_INVALID, "ipc: the schema must be a struct of columns (format \"+s\")");
  for (int64_t i = 0; i < s->n_children; ++i)
    if (s->children[i]->dictionary || s->children[i]->n_children) fail(SAILGPU_ERR_UNSUPPORTED, "ipc: nested and dictionary columns are not produced by the operators");
}

void write_schema_message(std::vector<uint8_t>& out, const ArrowSchema* s, const std::vector<ColType>& types) {
  MessageHead m = begin_message(/*Schema*/ 1, 0);
  std::vector<size_t> at;
  const size_t schema = m.fb.table({{0, 2, 0u}, {1, 0, 0}}, &at);          // endianness Little, fields
  m.fb.point(m.header_field, schema);
  size_t vec;
  const size_t first = m.fb.offset_vector((size_t)s->n_children, &vec);
  m.fb.point(at[1], vec);
  for (int64_t i = 0; i < s->n_children; ++i) write_field(m.fb, first + 4 * (size_t)i, s->children[i], types[(size_t)i]);
  append_message(out, m.fb);
}

inline int64_t pad8(int64_t n) { return (n + 7) & ~(int64_t)7; }

struct BodyPiece { const uint8_t* src; int64_t len; int bit_offset; int64_t bits; };   // bits > 0: a bitmap to re-pack from bit_offset

void copy_bits(uint8_t* dst, const uint8_t* src, int64_t first, int64_t n) {
  if ((first & 7) == 0) { memcpy(dst, src + (first >> 3), (size_t)((n + 7) >> 3)); if (n & 7) dst[(n - 1) >> 3] &= (uint8_t)((1u << (n & 7)) - 1); return; }
  memset(dst, 0, (size_t)((n + 7) >> 3));
  for (int64_t i = 0; i < n; ++i) { const int64_t j = first + i; if ((src[j >> 3] >> (j & 7)) & 1) dst[i >> 3] |= (uint8_t)(1u << (i & 7)); }
}

int64_t count_nulls(const ArrowArray* a) {
  if (a->null_count >= 0) return a->null_count;
  const uint8_t* v = a->n_buffers > 0 ? static_cast<const uint8_t*>(a->buffers[0]) : nullptr;
  if (!v) return 0;
  int64_t valid = 0, i = 0;
  const int64_t n = a->length, o = a->offset;
  for (; i < n && ((o + i) & 7); ++i) valid += (v[(o + i) >> 3] >> ((o + i) & 7)) & 1;          // up to a byte boundary
  for (; i + 64 <= n; i += 64) { uint64_t w; memcpy(&w, v + ((o + i) >> 3), 8); valid += __builtin_popcountll(w); }
  for (; i < n; ++i) valid += (v[(o + i) >> 3] >> ((o + i) & 7)) & 1;
  return n - valid;
}

void write_batch_message(std::vector<uint8_t>& out, const ArrowSchema* s, const ArrowArray* batch, const std::vector<ColType>& types) {
  if (batch->n_children != s->n_children) fail(SAILGPU_ERR_INVALID, "ipc: the batch does not have the schema's number of columns");
  if (batch->offset != 0) fail(SAILGPU_ERR_UNSUPPORTED, "ipc: a sliced struct array");
  std::vector<int64_t> nodes, buffers, variadic;
  std::vector<BodyPiece> pieces;
  std::vector<std::vector<uint8_t>> owned;       // rebased offsets buffers of sliced variable-length columns
  int64_t body = 0;
  auto add = [&](const uint8_t* src, int64_t len, int bit_offset = 0, int64_t bits = 0) {
    buffers.push_back(body); buffers.push_back(len);
    pieces.push_back({src, len, bit_offset, bits});
    body += pad8(len);
  };
  for (int64_t c = 0; c < s->n_children; ++c) {
    const ArrowArray* a = batch->children[c];
    const ColType& t = types[(size_t)c];
    if (a->length != batch->length) fail(SAILGPU_ERR_INVALID, "ipc: column length differs from the batch length");
    const int64_t n = a->length, o = a->offset;
    const int64_t nulls = count_nulls(a);
    nodes.push_back(n); nodes.push_back(nulls);
    const uint8_t* valid = a->n_buffers > 0 ? static_cast<const uint8_t*>(a->buffers[0]) : nullptr;
    if (nulls == 0 || !valid) add(nullptr, 0);
    else add(valid + (o >> 3), (n + 7) >> 3, (int)(o & 7), n);
    const uint8_t* b1 = a->n_buffers > 1 ? static_cast<const uint8_t*>(a->buffers[1]) : nullptr;
    switch (t.layout) {
      case L_FIXED: add(n ? b1 + o * t.width : nullptr, n * t.width); break;
      case L_BOOL: add(n ? b1 + (o >> 3) : nullptr, (n + 7) >> 3, (int)(o & 7), n); break;
      case L_VIEW: {
        add(n ? b1 + o * 16 : nullptr, n * 16);
        const int64_t n_var = a->n_buffers - 3;                       // [validity, views, data..., sizes]
        if (n_var < 0) fail(SAILGPU_ERR_INVALID, "ipc: a view column needs its buffer-sizes buffer");
        const int64_t* sizes = static_cast<const int64_t*>(a->buffers[a->n_buffers - 1]);
        for (int64_t k = 0; k < n_var; ++k) add(static_cast<const uint8_t*>(a->buffers[2 + k]), sizes[k]);
        variadic.push_back(n_var);
        break;
      }
      case L_VARLEN32: case L_VARLEN64: {
        const int w = t.layout == L_VARLEN32 ? 4 : 8;
        const uint8_t* data = a->n_buffers > 2 ? static_cast<const uint8_t*>(a->buffers[2]) : nullptr;
        auto off_at = [&](int64_t i) -> int64_t { if (w == 4) { int32_t v; memcpy(&v, b1 + (o + i) * 4, 4); return v; } int64_t v; memcpy(&v, b1 + (o + i) * 8, 8); return v; };
        if (n == 0 || !b1) {                                          // an empty column still carries one offset
          owned.emplace_back((size_t)w, 0);
          add(owned.back().data(), w); add(nullptr, 0);
          break;
        }
        const int64_t first = off_at(0), last = off_at(n);
        if (first == 0) add(b1 + o * w, (n + 1) * w);
        else {                                                        // rebase so that the data buffer starts at the first value
          owned.emplace_back((size_t)((n + 1) * w));
          for (int64_t i = 0; i <= n; ++i) { const int64_t v = off_at(i) - first; if (w == 4) { const int32_t x = (int32_t)v; memcpy(&owned.back()[(size_t)i * 4], &x, 4); } else memcpy(&owned.back()[(size_t)i * 8], &v, 8); }
          add(owned.back().data(), (n + 1) * w);
        }
        add(data ? data + first : nullptr, last - first);
        break;
      }
    }
  }
  MessageHead m = begin_message(/*RecordBatch*/ 3, body);
  std::vector<size_t> at;
  std::vector<Fb::FieldDef> defs = {{0, 8, (uint64_t)batch->length}, {1, 0, 0}, {2, 0, 0}};
  if (!variadic.empty()) defs.push_back({4, 0, 0});
  const size_t rb = m.fb.table(defs, &at);
  m.fb.point(m.header_field, rb);
  m.fb.point(at[1], m.fb.vector64(nodes, nodes.size() / 2));
  m.fb.point(at[2], m.fb.vector64(buffers, buffers.size() / 2));
  if (!variadic.empty()) m.fb.point(at[3], m.fb.vector64(variadic, variadic.size()));
  append_message(out, m.fb);
  const size_t body_at = out.size();
  out.resize(body_at + (size_t)body, 0);
  for (size_t i = 0; i < pieces.size(); ++i) {
    const BodyPiece& p = pieces[i];
    if (p.len == 0) continue;
    uint8_t* dst = &out[body_at + (size_t)buffers[2 * i]];
    if (p.bits) copy_bits(dst, p.src, p.bit_offset, p.bits);
    else memcpy(dst, p.src, (size_t)p.len);
  }
}

thread_local std::string g_ipc_error;

}  // namespace

extern "C" {

#define SG_IPC_API __attribute__((visibility("default")))

// One self-contained IPC stream -- Schema message, a RecordBatch message for `batch` (omitted when batch is NULL), end-of-stream
// marker -- in a malloc'ed buffer the caller returns with sailgpu_ipc_free.  Host arrays only; touches no device.
SG_IPC_API int32_t sailgpu_ipc_stream(const struct ArrowSchema* schema, const struct ArrowArray* batch, uint8_t** data, size_t* len) {
  if (!data || !len) return SAILGPU_ERR_INVALID;
  *data = nullptr; *len = 0;
  try {
    check_schema(schema);
    std::vector<ColType> types;
    for (int64_t i = 0; i < schema->n_children; ++i) types.push_back(parse_format(schema->children[i]->format));
    std::vector<uint8_t> out;
    write_schema_message(out, schema, types);
    if (batch) write_batch_message(out, schema, batch, types);
    const uint32_t eos[2] = {0xFFFFFFFFu, 0u};
    const size_t n = out.size();
    out.resize(n + 8);
    memcpy(&out[n], eos, 8);
    uint8_t* p = static_cast<uint8_t*>(malloc(out.size()));
    if (!p) { g_ipc_error = "ipc: out of host memory"; return SAILGPU_ERR_CUDA; }
    memcpy(p, out.data(), out.size());
    *data = p; *len = out.size();
    return SAILGPU_OK;
  } catch (const IpcError& e) {
    g_ipc_error = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    g_ipc_error = std::string("ipc: ") + e.what();
    return SAILGPU_ERR_INVALID;
  }
}

SG_IPC_API const char* sailgpu_ipc_last_error(void) { return g_ipc_error.c_str(); }

SG_IPC_API void sailgpu_ipc_free(uint8_t* data) { free(data); }

}  // extern "C"
