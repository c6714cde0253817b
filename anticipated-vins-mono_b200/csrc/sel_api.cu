// sel_api.cu -- placeholder until the selector kernels land (next commit).
#include "ctx.h"
using namespace bvio;
extern "C" {
void bvio_sel_ctx_destroy(bvio_ctx*) {}
int bvio_select(bvio_ctx* ctx, const bvio_select_in*, int32_t*, double*, bvio_select_summary*) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
int bvio_nccl_unique_id(void*) { return BVIO_ERR_UNSUPPORTED; }
int bvio_comm_init(bvio_ctx* ctx, const void*, int32_t, int32_t) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
int bvio_select_sharded(bvio_ctx* ctx, const bvio_select_in*, int32_t*, double*, bvio_select_summary*) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
int bvio_select_upload(bvio_ctx* ctx, const bvio_select_in*, bvio_selprob**) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
int bvio_select_run(bvio_ctx* ctx, bvio_selprob*) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
int bvio_select_fetch(bvio_ctx* ctx, bvio_selprob*, int32_t*, double*, bvio_select_summary*) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
void bvio_select_free(bvio_ctx*, bvio_selprob*) {}
int bvio_debug_build_delta(bvio_ctx* ctx, const bvio_select_in*, double*, int32_t*, double*) { return fail(ctx, BVIO_ERR_UNSUPPORTED, "selector not built"); }
}
