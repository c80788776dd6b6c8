// ops_more.cu -- HashJoinExec, SortExec, RepartitionExec and the NCCL exchange (filled in below).
#include "engine.hpp"

namespace sg {
std::unique_ptr<Op> make_join_op(Ctx*, const Json&, const std::vector<Schema>&) { fail(SAILGPU_ERR_UNSUPPORTED, "hash_join: not built yet"); }
std::unique_ptr<Op> make_sort_op(Ctx*, const Json&, const std::vector<Schema>&) { fail(SAILGPU_ERR_UNSUPPORTED, "sort: not built yet"); }
std::unique_ptr<Op> make_repartition_op(Ctx*, const Json&, const std::vector<Schema>&) { fail(SAILGPU_ERR_UNSUPPORTED, "repartition: not built yet"); }
}  // namespace sg

extern "C" {
SAILGPU_API int32_t sailgpu_comm_unique_id(uint8_t*) { return SAILGPU_ERR_UNSUPPORTED; }
SAILGPU_API int32_t sailgpu_ctx_comm_init(sailgpu_ctx*, const uint8_t*, int32_t, int32_t) { return SAILGPU_ERR_UNSUPPORTED; }
SAILGPU_API int32_t sailgpu_exchange(sailgpu_ctx*, const struct ArrowSchema*, struct ArrowDeviceArray*, int32_t, struct ArrowDeviceArray*) { return SAILGPU_ERR_UNSUPPORTED; }
}
