// Second translation unit of libdiffphore_sm100.so: dp_conv_fused2 (conv_fused2.cuh).  Compiled separately so that the two
// generations of the fused convolution kernel build in parallel; everything it shares with dp_abi.cu comes in through headers
// with internal linkage (anonymous namespace), the error helpers are dp_abi.cu's.
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"

namespace {
#include "edge_mlp.cuh"
#include "edge_mlp_tc.cuh"
#include "tp_tables.cuh"
#include "conv_fused.cuh"
#include "conv_fused2.cuh"
}  // namespace

#define NEED(cond, msg)                \
    do {                               \
        if (!(cond)) {                 \
            dp_set_error("%s", msg);   \
            return DP_ERR_ARG;         \
        }                              \
    } while (0)

static long long* g_cf2_dbg = nullptr;
/* profiling aid (not part of the public header): per-phase clock stamps of dp_conv_fused2 (layers 0 and 3) */
extern "C" int dp_debug_set_cf2_probe(long long* buf) { g_cf2_dbg = buf; return DP_OK; }

extern "C" int dp_conv_fused2(int32_t layer, const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                              const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const void* w1img,
                              float inv_w1scale, const void* w2img, float inv_wscale, const float* node_in, const int32_t* gather_idx,
                              const float* sh, int32_t sh_stride, const int32_t* seg_ptr, const int32_t* tile_node,
                              const int32_t* n_tiles_dev, int32_t n_tiles_cap, const float* oscale, const float* oshift, float* out,
                              const float* residual, int32_t res_dim, int32_t mode, void* stream) {
    NEED(mode != 1 || residual != nullptr, "dp_conv_fused2: mode 1 needs a residual");
    NEED(emb && tb && idxB && tc && idxC && w1img && w2img && node_in && sh && seg_ptr && tile_node && out,
         "dp_conv_fused2: null argument");
    NEED(strideB % 2 == 0 && strideC % 2 == 0, "dp_conv_fused2: node rows must be 8-byte aligned");
    NEED((reinterpret_cast<uintptr_t>(node_in) & 15) == 0, "dp_conv_fused2: node_in must be 16-byte aligned");
    ConvFusedArgs a;
    a.emb = emb; a.tb = tb; a.idxB = idxB; a.strideB = strideB; a.tc = tc; a.idxC = idxC; a.idxC2 = idxC2; a.strideC = strideC;
    a.w1img = w1img; a.inv_w1scale = inv_w1scale;
    a.w2img = w2img; a.inv_wscale = inv_wscale; a.node_in = node_in; a.gather_idx = gather_idx; a.perm = perm;
    a.sh = sh; a.sh_stride = sh_stride; a.seg_ptr = seg_ptr; a.tile_node = tile_node; a.n_tiles_dev = n_tiles_dev;
    a.n_tiles = n_tiles_cap; a.oscale = oscale; a.oshift = oshift; a.out = out; a.residual = residual; a.res_dim = res_dim;
    a.mode = mode;
    a.dbg = g_cf2_dbg;
    cudaStream_t st = (cudaStream_t)stream;
    switch (layer) {
        case DP_TP_L0: return conv_fused2_launch<TpL0>(a, st);
        case DP_TP_L1: return conv_fused2_launch<TpL1>(a, st);
        case DP_TP_L2: return conv_fused2_launch<TpL2>(a, st);
        case DP_TP_L3: return conv_fused2_launch<TpL3>(a, st);
        case DP_TP_TOR: return conv_fused2_launch<TpTor>(a, st);
    }
    dp_set_error("dp_conv_fused2: unsupported layer %d", layer);
    return DP_ERR_ARG;
}
