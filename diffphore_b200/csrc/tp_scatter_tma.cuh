// dp_tp_scatter, TMA-pipelined variant: same math and contract as tp_scatter.cuh; the per-edge weight rows are
// staged in shared memory by 1-D bulk copies (cp.async.bulk, one instruction per edge row, completion on an
// mbarrier) through a per-warp ring of STAGES rows, so STAGES-1 rows per warp are always in flight independent of
// register pressure (the register-streamed kernel is latency bound: ~13 resident warps x <= 1 row each).
// Persistent grid: every warp owns a contiguous, edge-balanced range of output nodes => a contiguous byte range of
// w; node boundaries only trigger the (shuffle-reduce + BatchNorm + residual) epilogue.
#pragma once
#include "tp_scatter.cuh"
#include "edge_mlp_tc.cuh"      // mbarrier / bulk-copy helpers

template <class Cfg, int WARPS, int STAGES>
struct TpTmaSmem {
    static constexpr int W_BYTES = Cfg::W * 4;
    static constexpr int RING = STAGES * W_BYTES;                          // per warp
    static constexpr int ZB = (Cfg::ZROWS + 2) * 16;
    static constexpr int XB = 128 * 4;
    static constexpr int BAR = ((STAGES * 8 + 127) / 128) * 128;
    static constexpr int PER_WARP = RING + ZB + XB + BAR;            // multiple of 16 B (bulk-copy destination alignment)
    static constexpr int TOTAL = WARPS * PER_WARP + 128;
};

__device__ __forceinline__ int tp_lower_bound(const int* __restrict__ a, int n, int v) {   // first i in [0,n] with a[i] >= v
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <class Cfg, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
tp_scatter_tma_kernel(const float* __restrict__ node_in, const int* __restrict__ gather_idx, const int* __restrict__ perm,
                      const float* __restrict__ sh, int sh_stride, const float* __restrict__ w,
                      const int* __restrict__ seg_ptr, const float* __restrict__ oscale, const float* __restrict__ oshift,
                      float* __restrict__ out, const float* __restrict__ residual, int res_dim, int mode, int n_out) {
    using S = TpTmaSmem<Cfg, WARPS, STAGES>;
    extern __shared__ __align__(128) uint8_t tp_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = tp_smem_raw + warp * S::PER_WARP;
    float* ring = reinterpret_cast<float*>(base);
    float (*zb)[4] = reinterpret_cast<float (*)[4]>(base + S::RING);
    float* xrow = reinterpret_cast<float*>(base + S::RING + S::ZB);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + S::RING + S::ZB + S::XB);

    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    const int E = seg_ptr[n_out];
    const int e_lo = (int)((long long)E * gw / nw), e_hi = (int)((long long)E * (gw + 1) / nw);
    const int n_begin = tp_lower_bound(seg_ptr, n_out, e_lo);
    const int n_end = (gw == nw - 1) ? n_out : tp_lower_bound(seg_ptr, n_out, e_hi);
    if (n_begin >= n_end) return;
    const int eb = seg_ptr[n_begin], ee = seg_ptr[n_end];

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) tc_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < STAGES && eb + s < ee; ++s) {
            tc_mbar_expect_tx(&full[s], S::W_BYTES);
            tc_bulk_load(ring + s * Cfg::W, w + (size_t)(eb + s) * Cfg::W, S::W_BYTES, &full[s]);
        }
    }
    float acc[Cfg::NO][2][3];
#pragma unroll
    for (int o = 0; o < Cfg::NO; ++o)
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[o][v][k] = 0.f;

    // software prefetch (one edge ahead) of the gathered node row and the edge's spherical harmonics
    constexpr int XR = (Cfg::D_IN + 31) / 32;
    float xpre[XR], shpre[Cfg::SH_USED];
    auto prefetch = [&](int e) {
        const int src = gather_idx ? gather_idx[e] : e;
        const int ce = perm ? perm[e] : e;
#pragma unroll
        for (int i = 0; i < XR; ++i) {
            const int c = lane + 32 * i;
            xpre[i] = (c < Cfg::D_IN) ? __ldg(node_in + (size_t)src * Cfg::D_IN + c) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < Cfg::SH_USED; ++i) shpre[i] = __ldg(sh + (size_t)ce * sh_stride + i);
    };
    if (eb < ee) prefetch(eb);

    int e = eb;
    for (int n = n_begin; n < n_end; ++n) {
        const int e1 = seg_ptr[n + 1];
        const int deg = e1 - seg_ptr[n];
        for (; e < e1; ++e) {
            const int j = e - eb, st = j % STAGES;
            float shv[Cfg::SH_USED];
#pragma unroll
            for (int i = 0; i < Cfg::SH_USED; ++i) shv[i] = shpre[i];
            __syncwarp();                                   // previous edge's readers of xrow / zb are done
#pragma unroll
            for (int i = 0; i < XR; ++i) xrow[lane + 32 * i] = xpre[i];
            if (e + 1 < ee) prefetch(e + 1);
            __syncwarp();
            tp_compute_z<Cfg, 0, true>(xrow, shv, zb, lane);
            __syncwarp();
            tc_mbar_wait(&full[st], (uint32_t)((j / STAGES) & 1));
            tp_accumulate<Cfg, 0, true>(ring + st * Cfg::W, zb, acc, lane);
            __syncwarp();                                   // all lanes finished reading this ring slot
            if (lane == 0 && e + STAGES < ee) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_mbar_expect_tx(&full[st], S::W_BYTES);
                tc_bulk_load(ring + st * Cfg::W, w + (size_t)(e + STAGES) * Cfg::W, S::W_BYTES, &full[st]);
            }
        }
        const float inv_deg = 1.0f / (float)(deg > 0 ? deg : 1);
        tp_epilogue<Cfg, 0>(acc, inv_deg, oscale, oshift, out + (size_t)n * Cfg::D_OUT,
                            residual ? residual + (size_t)n * res_dim : nullptr, res_dim, mode, lane);
#pragma unroll
        for (int o = 0; o < Cfg::NO; ++o)
#pragma unroll
            for (int v = 0; v < 2; ++v)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[o][v][k] = 0.f;
    }
}

template <class Cfg, int WARPS, int STAGES>
static int tp_scatter_tma_launch(const float* node_in, const int* gather_idx, const int* perm, const float* sh, int sh_stride,
                                 const float* w, const int* seg_ptr, const float* oscale, const float* oshift, float* out,
                                 const float* residual, int res_dim, int mode, int n_out, cudaStream_t st) {
    if (n_out <= 0) return DP_OK;
    using S = TpTmaSmem<Cfg, WARPS, STAGES>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
    static_assert((Cfg::W * 4) % 16 == 0, "bulk copies need 16-byte multiples");
    static bool attr_set = false;
    static int n_sm = 0;
    auto kern = tp_scatter_tma_kernel<Cfg, WARPS, STAGES>;
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    int grid = n_sm > 0 ? n_sm : 148;
    const int max_useful = (n_out + WARPS - 1) / WARPS;
    if (grid > max_useful) grid = max_useful;
    kern<<<grid, WARPS * 32, S::TOTAL, st>>>(node_in, gather_idx, perm, sh, sh_stride, w, seg_ptr, oscale, oshift, out, residual,
                                             res_dim, mode, n_out);
    return dp_check_launch("tp_scatter_tma");
}
