// placeholder until the tcgen05 kernel lands
#include "common.cuh"
extern "C" {
int pg_tc_linear_fwd(pg_ctx* ctx, const void*, const void*, const void*, void*, int, int64_t, int64_t, int64_t, int) { PG_CHECK_CTX(ctx); pg_set_error("tc path not built"); return PG_ERR_UNSUPPORTED; }
int pg_tc_linear_bwd(pg_ctx* ctx, const void*, const void*, const void*, int, void*, void*, void*, const void*, int64_t, int64_t, int64_t) { PG_CHECK_CTX(ctx); pg_set_error("tc path not built"); return PG_ERR_UNSUPPORTED; }
int pg_tc_gemm(pg_ctx* ctx, int, const void*, const void*, void*, int, int64_t, int64_t, int64_t) { PG_CHECK_CTX(ctx); pg_set_error("tc path not built"); return PG_ERR_UNSUPPORTED; }
}
