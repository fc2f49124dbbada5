// net.cu -- placeholder until the network kernels land (next commit).
#include "engine.cuh"
using namespace szb;
namespace szb {
int net_evaluate_batch(szb_ctx* ctx, int, int) { return fail(ctx, SZB_ERR_STATE, "network weights not loaded"); }
void net_destroy(szb_ctx*) {}
}
extern "C" {
int szb_net_load(szb_ctx* ctx, int32_t, const char* const*, const float* const*, const int64_t*) { return fail(ctx, SZB_ERR_UNSUPPORTED, "not built yet"); }
int szb_net_forward(szb_ctx* ctx, int32_t, const uint64_t*, int32_t, float*, float*) { return fail(ctx, SZB_ERR_STATE, "network weights not loaded"); }
int szb_net_forward_logits(szb_ctx* ctx, int32_t, const uint64_t*, int32_t, float*, float*) { return fail(ctx, SZB_ERR_STATE, "network weights not loaded"); }
}
