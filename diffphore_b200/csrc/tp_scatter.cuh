// dp_tp_scatter: irrep tensor product + Clebsch-Gordan contraction + edge->node scatter-mean + eval BatchNorm.
//
// Replaces, per TensorProductConvLayer.forward (score_model_phore.py:134-149):
//     tp  = e3nn FullyConnectedTensorProduct(node_attr[edge_dst], edge_sh, w)      (K6, 'uvw', per-edge weights)
//     out = torch_scatter.scatter(tp, edge_src, reduce='mean')                     (K7)
//     out = e3nn.nn.BatchNorm(out)   (eval)                                        (K8)
// plus the encoder's residual accumulation (score_model_phore.py:703-710) folded into the epilogue.
//
// Layout: edges are grouped by OUTPUT node (CSR seg_ptr); one warp owns one output node and streams the
// per-edge weight rows w[e, 0:W] (fp32, contiguous, written by dp_edge_mlp) exactly once from HBM with
// 8-byte loads covering 240 contiguous bytes per warp instruction; this kernel is HBM-bound on that stream
// (algorithmic bytes = E*(4W + 4 D_in + 4 SH + 8) + n_out*4*D_out).
// Per path (U x V weight block, row-major) a lane owns a pair of output channels (v, v+1) and one of 60/V
// interleaved row groups; the per-lane partial sums over rows AND over all edges of the node stay in
// registers and are combined across row groups with warp shuffles once per node (segmented reduction without
// atomics).  Z[u,k] = sum_ij C_ijk x[u,i] sh[j] is staged per edge in shared memory.
#pragma once
#include "common.cuh"
#include "tp_tables.cuh"

#define TP_WARPS 4

template <class Cfg, int P, bool X_IN_SMEM = false>
__device__ __forceinline__ void tp_compute_z(const float* __restrict__ x, const float* sh, float (*zb)[4], int lane) {
    if constexpr (P < Cfg::NP) {
        constexpr TpPath p = Cfg::paths[P];
        constexpr int LO = Cfg::outs[p.oi].lo;
        constexpr int D1 = 2 * p.l1 + 1;
        if (lane < p.U) {
            float xv[3], z[3];
#pragma unroll
            for (int i = 0; i < D1; ++i) xv[i] = X_IN_SMEM ? x[p.in_off + lane * D1 + i] : __ldg(x + p.in_off + lane * D1 + i);
            dp_cg<p.l1, p.l2, LO>(xv, sh + p.sh_off, z);
#pragma unroll
            for (int k = 0; k < 2 * LO + 1; ++k) zb[p.zoff + lane][k] = z[k];
        }
        tp_compute_z<Cfg, P + 1, X_IN_SMEM>(x, sh, zb, lane);
    }
}

template <class Cfg, int P, bool W_IN_SMEM = false>
__device__ __forceinline__ void tp_accumulate(const float* __restrict__ wrow, const float (*zb)[4],
                                              float (&acc)[Cfg::NO][2][3], int lane) {
    if constexpr (P < Cfg::NP) {
        constexpr TpPath p = Cfg::paths[P];
        constexpr int V = Cfg::outs[p.oi].V;
        constexpr int K = 2 * Cfg::outs[p.oi].lo + 1;
        constexpr int LPR = V / 2;            // lanes per weight row (each lane: 2 consecutive output channels)
        constexpr int RPI = 30 / LPR;         // rows per warp iteration (30 active lanes)
        constexpr int NIT = (p.U + RPI - 1) / RPI;
        const int g = lane / LPR, j = lane % LPR;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int row = it * RPI + g;
            const bool ok = (lane < 30) && (row < p.U);
            const int r = ok ? row : 0;
            float2 wv = make_float2(0.f, 0.f);
            if (ok) {
                if constexpr (W_IN_SMEM) wv = *(reinterpret_cast<const float2*>(wrow + p.w_off + r * V) + j);
                else wv = __ldg(reinterpret_cast<const float2*>(wrow + p.w_off + r * V) + j);
            }
            if constexpr (K == 1) {
                const float z0 = zb[p.zoff + r][0];
                acc[p.oi][0][0] = fmaf(wv.x, z0, acc[p.oi][0][0]);
                acc[p.oi][1][0] = fmaf(wv.y, z0, acc[p.oi][1][0]);
            } else {
                const float4 z = *reinterpret_cast<const float4*>(&zb[p.zoff + r][0]);
                acc[p.oi][0][0] = fmaf(wv.x, z.x, acc[p.oi][0][0]);
                acc[p.oi][0][1] = fmaf(wv.x, z.y, acc[p.oi][0][1]);
                acc[p.oi][0][2] = fmaf(wv.x, z.z, acc[p.oi][0][2]);
                acc[p.oi][1][0] = fmaf(wv.y, z.x, acc[p.oi][1][0]);
                acc[p.oi][1][1] = fmaf(wv.y, z.y, acc[p.oi][1][1]);
                acc[p.oi][1][2] = fmaf(wv.y, z.z, acc[p.oi][1][2]);
            }
        }
        tp_accumulate<Cfg, P + 1, W_IN_SMEM>(wrow, zb, acc, lane);
    }
}

template <int LPR>
__device__ __forceinline__ float tp_group_reduce(float a) {
    const unsigned m = 0xffffffffu;
    if constexpr (LPR == 10) {            // 3 row groups at lane offsets 0,10,20
        return a + __shfl_down_sync(m, a, 10) + __shfl_down_sync(m, a, 20);
    } else if constexpr (LPR == 5) {      // 6 row groups at lane offsets 0,5,...,25
        float b = a + __shfl_down_sync(m, a, 15);
        return b + __shfl_down_sync(m, b, 5) + __shfl_down_sync(m, b, 10);
    } else {                              // LPR == 1: 30 row groups, lanes 30/31 hold zeros
        static_assert(LPR == 1, "unsupported channel count");
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(m, a, o);
        return a;
    }
}

template <class Cfg, int O>
__device__ __forceinline__ void tp_epilogue(float (&acc)[Cfg::NO][2][3], float inv_deg, const float* __restrict__ oscale,
                                            const float* __restrict__ oshift, float* __restrict__ orow,
                                            const float* __restrict__ res, int res_dim, int mode, int lane) {
    if constexpr (O < Cfg::NO) {
        constexpr TpOut o = Cfg::outs[O];
        constexpr int K = 2 * o.lo + 1, LPR = o.V / 2;
#pragma unroll
        for (int vv = 0; vv < 2; ++vv)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float s = tp_group_reduce<LPR>(acc[O][vv][k]);
                if (lane < LPR) {
                    const int d = o.off + (2 * lane + vv) * K + k;
                    float val = s * inv_deg * oscale[d] + oshift[d];
                    if (mode == 1) val += (d < res_dim) ? res[d] : 0.0f;
                    else if (mode == 2) val += orow[d];
                    orow[d] = val;
                }
            }
        tp_epilogue<Cfg, O + 1>(acc, inv_deg, oscale, oshift, orow, res, res_dim, mode, lane);
    }
}

// mode 0: out = conv ; mode 1: out = pad(residual) + conv ; mode 2: out += conv
template <class Cfg>
__global__ void __launch_bounds__(TP_WARPS * 32)
tp_scatter_kernel(const float* __restrict__ node_in, const int* __restrict__ gather_idx, const int* __restrict__ perm,
                  const float* __restrict__ sh, int sh_stride, const float* __restrict__ w,
                  const int* __restrict__ seg_ptr, const float* __restrict__ oscale, const float* __restrict__ oshift,
                  float* __restrict__ out, const float* __restrict__ residual, int res_dim, int mode, int n_out) {
    __shared__ __align__(16) float zbuf[TP_WARPS][Cfg::ZROWS + 2][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * TP_WARPS + warp;
    if (n >= n_out) return;
    float (*zb)[4] = zbuf[warp];
    float acc[Cfg::NO][2][3];
#pragma unroll
    for (int o = 0; o < Cfg::NO; ++o)
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[o][v][k] = 0.f;
    const int e0 = seg_ptr[n], e1 = seg_ptr[n + 1];
    for (int e = e0; e < e1; ++e) {
        const int src = gather_idx ? gather_idx[e] : e;
        const int ce = perm ? perm[e] : e;
        float shv[Cfg::SH_USED];
#pragma unroll
        for (int i = 0; i < Cfg::SH_USED; ++i) shv[i] = __ldg(sh + (size_t)ce * sh_stride + i);
        __syncwarp();
        tp_compute_z<Cfg, 0, false>(node_in + (size_t)src * Cfg::D_IN, shv, zb, lane);
        __syncwarp();
        tp_accumulate<Cfg, 0, false>(w + (size_t)e * Cfg::W, zb, acc, lane);
    }
    const int deg = e1 - e0;
    const float inv_deg = 1.0f / (float)(deg > 0 ? deg : 1);
    tp_epilogue<Cfg, 0>(acc, inv_deg, oscale, oshift, out + (size_t)n * Cfg::D_OUT,
                        residual ? residual + (size_t)n * res_dim : nullptr, res_dim, mode, lane);
}

template <class Cfg>
static int tp_scatter_launch(const float* node_in, const int* gather_idx, const int* perm, const float* sh, int sh_stride,
                             const float* w, const int* seg_ptr, const float* oscale, const float* oshift, float* out,
                             const float* residual, int res_dim, int mode, int n_out, cudaStream_t st) {
    if (n_out <= 0) return DP_OK;
    dim3 grid((n_out + TP_WARPS - 1) / TP_WARPS);
    tp_scatter_kernel<Cfg><<<grid, TP_WARPS * 32, 0, st>>>(node_in, gather_idx, perm, sh, sh_stride, w, seg_ptr, oscale,
                                                          oshift, out, residual, res_dim, mode, n_out);
    return dp_check_launch("tp_scatter");
}
