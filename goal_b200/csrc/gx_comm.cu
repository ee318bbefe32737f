// gx_comm.cu -- interface exchange between mesh parts (placeholder; filled in below).
#include "gx_internal.h"

namespace gx {
void comm_destroy(gx_ctx*) {}
int comm_setup_lists(gx_ctx* c, const gx_desc* d) {
  if (d->n_ranks > 1 && d->n_peers > 0) { c->err = "multi-part contexts are not implemented yet"; return GX_ERR_UNSUPPORTED; }
  return GX_OK;
}
}  // namespace gx

extern "C" {
int gx_comm_init(gx_ctx* ctx, const void*, size_t) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
int gx_nccl_unique_id(void*, size_t*) { return GX_ERR_UNSUPPORTED; }
int gx_reduce_interfaces(gx_ctx* ctx, int) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
int gx_allreduce_sum(gx_ctx* ctx, double*, int) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
int gx_interface_bytes(gx_ctx* ctx, int, int, int64_t*, int64_t*) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
int gx_pack_interface(gx_ctx* ctx, int, int, void**) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
int gx_unpack_add_interface(gx_ctx* ctx, int, int, const void*) { if (ctx) ctx->err = "not implemented"; return GX_ERR_UNSUPPORTED; }
}
