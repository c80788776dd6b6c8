// h2d.hpp -- host (pageable) Arrow buffers -> HBM: the ingest side of sailgpu_op_push.
//
// A DataFusion RecordBatch lives in pageable memory; cudaMemcpy from it is staged by the driver one piece at a time
// (~10 GB/s) and ships Arrow's fixed widths (16 bytes for a Decimal128(15,2) that needs 2, 16 for a one-character
// view).  Here a pool of host threads walks the columns in 256 K-row pieces: each piece is range-checked, PACKED into
// pinned staging memory (frame-of-reference integers of 1/2/4/8 bytes for Int32/Int64/Date32 and Decimal128 whose high
// word is a sign extension; [length byte | bytes] for inline string views; raw otherwise), copied to the device on the
// context's copy stream and expanded back into the Arrow layout by a small kernel.  Packing overlaps the copies of the
// previous pieces (two staging slots per thread) and the compute stream only waits for the last copy through an event.
// The producer's buffers are no longer needed once flush() returns (everything was read into staging memory).
#pragma once
#include <condition_variable>
#include <thread>

#include "device.hpp"

namespace sg {

struct PackPool;      // lives in the context (created on first host batch)

enum class HostCol : int { Raw = 0, Dec128, Int64, Int32, View16 };

class HostStager {
 public:
  explicit HostStager(Ctx* c) : ctx(c) {}
  // dst: device buffer of n * out_width bytes in Arrow layout; src: host values buffer (n elements of the Arrow width)
  void add(HostCol kind, void* dst, const void* src, int64_t n, int width);
  void add_raw(void* dst, const void* src, size_t bytes) { add(HostCol::Raw, dst, src, (int64_t)bytes, 1); }
  // packs + copies everything queued; on return the host sources are no longer referenced and ctx->stream waits for the data
  void flush();
  bool empty() const { return items.empty(); }

  struct Item { HostCol kind; uint8_t* dst; const uint8_t* src; int64_t n; int width; };
 private:
  Ctx* ctx;
  std::vector<Item> items;
};

void destroy_pack_pool(Ctx* ctx);

}  // namespace sg
