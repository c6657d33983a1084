// Per-step graph construction and edge/node feature kernels (everything of the score model that is not a
// TensorProductConvLayer).  One CTA per ligand-pharmacophore graph; positions staged in shared memory.
//
//   lig_count / lig_fill    radius_graph + bond edges + GaussianSmearing + SH + lig_edge_embedding
//                           (score_model_phore.py:715-739, 657; torch_cluster 1.6.0 radius_graph semantics)
//   pp_setup / pp_step      build_phore_conv_graph + phore_edge_embedding (smp:742-756, 663): geometry is static
//   cross_setup / cross_step _build_phoretype_cross_conv_graph + cross_edge_embedding (smp:759-895, 673)
//   node_embed              AtomEncoder lig/phore + boarder_analyze/boarder_embedding (smp:64-73, 649-655, 898-935)
//   center_step / score_head build_center_conv_graph + tr/rot heads (smp:335-352, 381-406)
//   tor_count / tor_fill / tor_head   build_bond_conv_graph + FullTensorProduct + tor head (smp:361-377, 409-437)
#pragma once
#include "common.cuh"
#include "tp_tables.cuh"

// per-step constant block (floats), filled on the host once per noise level
#define SC_SEMB 0
#define SC_LIG_NODE 20
#define SC_PH_NODE 40
#define SC_LIG_EDGE 60
#define SC_PP_EDGE 80
#define SC_CROSS_EDGE 100
#define SC_CENTER 120
#define SC_TR 140
#define SC_ROT 160
#define SC_INV_TR_SIGMA 180
#define SC_SO3_NORM 181
#define SC_SQRT_TORUS 182
#define SC_TR_A 183
#define SC_TR_B 184
#define SC_ROT_A 185
#define SC_ROT_B 186
#define SC_TOR_A 187
#define SC_TOR_B 188
#define SC_SIZE 256

__device__ __forceinline__ float dp_dist2(float ax, float ay, float az, float bx, float by, float bz) {
    // un-fused (mul, add) evaluation like the reference's ((a-b)**2).sum(-1)
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// 20 -> 20 second layer shared by all edge-embedding MLPs: out = w3 . relu(h) + b3
__device__ __forceinline__ void dp_mlp20_out(const float* h, const float* __restrict__ w3, const float* __restrict__ b3,
                                             float* __restrict__ out) {
#pragma unroll 4
    for (int o = 0; o < 20; ++o) {
        float s = b3[o];
#pragma unroll
        for (int c = 0; c < 20; ++c) s = fmaf(fmaxf(h[c], 0.f), w3[o * 20 + c], s);
        out[o] = s;
    }
}

// The 20 -> 20 (-> 20) edge-embedding MLPs of the per-edge kernels with the weights TRANSPOSED in shared memory (wt[c * 20 + o]) and
// packed fp32 FMAs: out pairs (2j, 2j + 1) accumulate x[c] * w[o][c] over c = 0 .. NIN-1 in the same order as the scalar loops they
// replace (every FFMA2 lane is an IEEE fma: identical results), with one LDS.128 per four weights instead of one LDS per FMA and all
// accumulators in registers (the scalar version kept h[20] in local memory and ran ~4000 instructions per edge).
template <int NIN>
__device__ __forceinline__ void dp_dense20(const float (&x)[NIN], const float* __restrict__ wt, float2 (&acc)[10]) {
#pragma unroll
    for (int c = 0; c < NIN; ++c) {
        if ((c & 1) == 0) asm volatile("" ::: "memory");     // keeps ptxas from hoisting all 5 * NIN weight loads (it spills otherwise)
        const float2 xc = make_float2(x[c], x[c]);
        const float4* w4 = reinterpret_cast<const float4*>(wt + c * 20);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float4 w = w4[q];
            acc[2 * q] = __ffma2_rn(xc, make_float2(w.x, w.y), acc[2 * q]);
            acc[2 * q + 1] = __ffma2_rn(xc, make_float2(w.z, w.w), acc[2 * q + 1]);
        }
    }
}
// second layer: out = w3 . relu(h) + b3, written as five float4 (rows of the embedding arrays are 80 bytes: 16-byte aligned)
__device__ __forceinline__ void dp_mlp20_out_t(const float2 (&h2)[10], const float* __restrict__ w3t, const float* __restrict__ b3,
                                               float* __restrict__ out) {
    float r[20];
#pragma unroll
    for (int j = 0; j < 10; ++j) { r[2 * j] = fmaxf(h2[j].x, 0.f); r[2 * j + 1] = fmaxf(h2[j].y, 0.f); }
    float2 acc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = make_float2(b3[2 * j], b3[2 * j + 1]);
    dp_dense20<20>(r, w3t, acc);
#pragma unroll
    for (int q = 0; q < 5; ++q)
        reinterpret_cast<float4*>(out)[q] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
}

// ---------------------------------------------------------------------------------------------------------------
// ligand-ligand graph
// ---------------------------------------------------------------------------------------------------------------
#define LG_THREADS 128
#define LG_MAXN 512

// pass 1: per-centre cap thresholds + per-out-node degrees
__global__ void __launch_bounds__(LG_THREADS)
lig_count_kernel(const float* __restrict__ pos, const int* __restrict__ lig_ptr, const int* __restrict__ bond_ptr,
                 int* __restrict__ thr, int* __restrict__ deg, int* __restrict__ gcount) {
    __shared__ float sp[LG_MAXN * 3];
    __shared__ int sthr[LG_MAXN];
    __shared__ int stot;
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0;
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) sp[i] = pos[(size_t)a0 * 3 + i];
    if (threadIdx.x == 0) stot = 0;
    __syncthreads();
    const float r2 = c_dp.lig_radius * c_dp.lig_radius;
    const int cap = c_dp.max_neighbors + 1;     // radius(x, x, max_num_neighbors + 1) incl. the self pair
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int cnt = 0, t = n;
        for (int k = 0; k < n; ++k) {
            if (dp_dist2(sp[i * 3], sp[i * 3 + 1], sp[i * 3 + 2], sp[k * 3], sp[k * 3 + 1], sp[k * 3 + 2]) < r2) {
                if (++cnt == cap) { t = k; break; }
            }
        }
        sthr[i] = t;
        thr[a0 + i] = t;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        int d = bond_ptr[a0 + j + 1] - bond_ptr[a0 + j];
        for (int i = 0; i < n; ++i)
            if (i != j && j <= sthr[i] &&
                dp_dist2(sp[i * 3], sp[i * 3 + 1], sp[i * 3 + 2], sp[j * 3], sp[j * 3 + 1], sp[j * 3 + 2]) < r2) ++d;
        deg[a0 + j] = d;
        atomicAdd(&stot, d);
    }
    __syncthreads();
    if (threadIdx.x == 0) gcount[g] = stot;
}

// exclusive scan of per-graph counts (single CTA); total -> gstart[n] and *total_out
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ cnt, int* __restrict__ start, int n,
                                                    int* __restrict__ total_out) {
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (n + 1023) / 1024;
    const int b = t * per, e = min(b + per, n);
    int s = 0;
    for (int i = b; i < e; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < 1024; ++i) { int v = part[i]; part[i] = run; run += v; }
        start[n] = run;
        if (total_out) *total_out = run;
    }
    __syncthreads();
    int run = part[t];
    for (int i = b; i < e; ++i) { start[i] = run; run += cnt[i]; }
}

// pass 2: emit edges (CSR by out node) and their features
__global__ void __launch_bounds__(LG_THREADS)
lig_fill_kernel(const float* __restrict__ pos, const int* __restrict__ lig_ptr, const int* __restrict__ bond_ptr,
                const int* __restrict__ bond_dst, const int* __restrict__ bond_type, const int* __restrict__ thr,
                const int* __restrict__ deg, const int* __restrict__ gstart, int n_graphs, DpSmallWeights sw,
                const float* __restrict__ sc, int* __restrict__ seg_ptr, int* __restrict__ e_src, int* __restrict__ e_dst,
                float* __restrict__ e_emb, float* __restrict__ e_sh) {
    __shared__ float sp[LG_MAXN * 3];
    __shared__ int sthr[LG_MAXN], soff[LG_MAXN + 1];
    __shared__ __align__(16) float w_rbf[20 * 20], w3[20 * 20];       // transposed: [c][o]
    __shared__ float w_bond[20 * 4], b3[20], cst[20];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0;
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) sp[i] = pos[(size_t)a0 * 3 + i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sthr[i] = thr[a0 + i];
    for (int i = threadIdx.x; i < 400; i += blockDim.x) {
        const int c = i / 20, o = i % 20;
        w_rbf[i] = sw.lig_edge.w0[o * 44 + 24 + c];
        w3[i] = sw.lig_edge.w3[o * 20 + c];
    }
    for (int i = threadIdx.x; i < 80; i += blockDim.x) w_bond[i] = sw.lig_edge.w0[(i / 4) * 44 + (i % 4)];
    for (int i = threadIdx.x; i < 20; i += blockDim.x) { b3[i] = sw.lig_edge.b3[i]; cst[i] = sc[SC_LIG_EDGE + i]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = gstart[g];
        for (int j = 0; j < n; ++j) { soff[j] = run; run += deg[a0 + j]; }
        soff[n] = run;
    }
    __syncthreads();
    // gridDim.y CTAs share a graph (small jobs: 40 graphs would occupy 40 of 148 SMs): each takes a contiguous range of out nodes
    const int chunk = (n + (int)gridDim.y - 1) / (int)gridDim.y;
    const int j_lo = min((int)blockIdx.y * chunk, n), j_hi = min(j_lo + chunk, n);
    for (int j = j_lo + threadIdx.x; j < j_hi; j += blockDim.x) seg_ptr[a0 + j] = soff[j];
    if (g == n_graphs - 1 && blockIdx.y == 0 && threadIdx.x == 0) seg_ptr[a0 + n] = soff[n];
    const float r2 = c_dp.lig_radius * c_dp.lig_radius;
    // topology: WARP per out node (a thread per node left 3/4 of the CTA waiting at the barrier below: 19 % of the kernel's
    // samples): the lanes test 32 candidate sources at once and a ballot keeps the edges in index order
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
        const unsigned lt = (1u << lane) - 1u;
        for (int j = j_lo + warp; j < j_hi; j += nwarp) {
            int o = soff[j];
            const int b0 = bond_ptr[a0 + j], b1 = bond_ptr[a0 + j + 1];
            for (int b = b0 + lane; b < b1; b += 32) {
                const int ob = o + (b - b0);
                e_src[ob] = a0 + j; e_dst[ob] = bond_dst[b];
                e_sh[(size_t)ob * DP_SH] = (float)bond_type[b];  // stash the bond type, overwritten below
            }
            o += b1 - b0;
            const float jx = sp[j * 3], jy = sp[j * 3 + 1], jz = sp[j * 3 + 2];
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const bool hit = i < n && i != j && j <= sthr[i] && dp_dist2(sp[i * 3], sp[i * 3 + 1], sp[i * 3 + 2], jx, jy, jz) < r2;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const int oe = o + __popc(m & lt);
                    e_src[oe] = a0 + j; e_dst[oe] = a0 + i;
                    e_sh[(size_t)oe * DP_SH] = -1.0f;
                }
                o += __popc(m);
            }
        }
    }
    __syncthreads();
    // features: thread per edge
    for (int e = soff[j_lo] + threadIdx.x; e < soff[j_hi]; e += blockDim.x) {
        const int s = e_src[e] - a0, d = e_dst[e] - a0;
        const int bt = (int)e_sh[(size_t)e * DP_SH];
        const float vx = sp[d * 3] - sp[s * 3], vy = sp[d * 3 + 1] - sp[s * 3 + 1], vz = sp[d * 3 + 2] - sp[s * 3 + 2];
        float rbf[20], sh[9];
        dp_rbf20(sqrtf(vx * vx + vy * vy + vz * vz), DP_RBF_LIG, rbf);
        float2 h2[10];
#pragma unroll
        for (int j = 0; j < 10; ++j)
            h2[j] = make_float2(cst[2 * j] + (bt >= 0 ? w_bond[(2 * j) * 4 + bt] : 0.f), cst[2 * j + 1] + (bt >= 0 ? w_bond[(2 * j + 1) * 4 + bt] : 0.f));
        dp_dense20<20>(rbf, w_rbf, h2);
        dp_mlp20_out_t(h2, w3, b3, e_emb + (size_t)e * 20);
        dp_sh9(vx, vy, vz, sh);
#pragma unroll
        for (int i = 0; i < 9; ++i) e_sh[(size_t)e * DP_SH + i] = sh[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// phore-phore edges: static geometry -> precomputed first-layer partial + SH; per step only the sigma term moves
// ---------------------------------------------------------------------------------------------------------------
__global__ void pp_setup_kernel(const float* __restrict__ ppos, const int* __restrict__ src, const int* __restrict__ dst,
                                int n_edges, DpSmallWeights sw, float* __restrict__ pp_h, float* __restrict__ pp_sh) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int s = src[e], d = dst[e];
    const float vx = ppos[d * 3] - ppos[s * 3], vy = ppos[d * 3 + 1] - ppos[s * 3 + 1], vz = ppos[d * 3 + 2] - ppos[s * 3 + 2];
    float rbf[20], sh[9];
    dp_rbf20(sqrtf(vx * vx + vy * vy + vz * vz), DP_RBF_PHORE, rbf);
    for (int o = 0; o < 20; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 20; ++c) acc = fmaf(rbf[c], sw.pp_edge.w0[o * 40 + 20 + c], acc);
        pp_h[(size_t)e * 20 + o] = acc;
    }
    dp_sh9(vx, vy, vz, sh);
    for (int i = 0; i < 9; ++i) pp_sh[(size_t)e * DP_SH + i] = sh[i];
}

__global__ void pp_step_kernel(const float* __restrict__ pp_h, int n_edges, DpSmallWeights sw, const float* __restrict__ sc,
                               float* __restrict__ pp_emb) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    float h[20];
#pragma unroll
    for (int o = 0; o < 20; ++o) h[o] = pp_h[(size_t)e * 20 + o] + sc[SC_PP_EDGE + o];
    dp_mlp20_out(h, sw.pp_edge.w3, sw.pp_edge.b3, pp_emb + (size_t)e * 20);
}

// ---------------------------------------------------------------------------------------------------------------
// cross (ligand x pharmacophore) edges: complete bipartite per graph, edge id = cross_ptr[g] + a_local*P + p_local
// ---------------------------------------------------------------------------------------------------------------
// static part: agreement / phoretype_attr dependent terms
__global__ void cross_setup_kernel(const int* __restrict__ cross_lig, const int* __restrict__ cross_ph, int n_edges,
                                   const float* __restrict__ phorefp, const float* __restrict__ phoretype,
                                   DpSmallWeights sw, float* __restrict__ cross_h, float* __restrict__ cross_fm) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int a = cross_lig[e], p = cross_ph[e];
    float attr[33];
    const bool is_ex = phoretype[p * 11 + 10] == 1.0f;
    for (int t = 0; t < 11; ++t) {
        const float pt = phoretype[p * 11 + t], fp = phorefp[a * 11 + t];
        attr[t] = is_ex ? 0.f : pt * fp;
        attr[11 + t] = pt;
        attr[22 + t] = fp;
    }
    for (int o = 0; o < 20; ++o) {
        float acc = 0.f;
        for (int c = 0; c < 33; ++c) acc = fmaf(attr[c], sw.cross_edge.w0[o * 73 + 40 + c], acc);
        cross_h[(size_t)e * 20 + o] = acc;
    }
    float fm = sw.pmt.b3[0];
    for (int o = 0; o < 11; ++o) {
        float acc = sw.pmt.b0[o];
        for (int c = 0; c < 33; ++c) acc = fmaf(attr[c], sw.pmt.w0[o * 33 + c], acc);
        fm = fmaf(fmaxf(acc, 0.f), sw.pmt.w3[o], fm);
    }
    cross_fm[e] = dp_softplus(fm);
}

#define CR_THREADS 128
__global__ void __launch_bounds__(CR_THREADS)
cross_step_kernel(const float* __restrict__ lpos, const float* __restrict__ lnorm, const float* __restrict__ ppos,
                  const float* __restrict__ pnorm, const int* __restrict__ lig_ptr, const int* __restrict__ ph_ptr,
                  const int* __restrict__ cross_ptr, const float* __restrict__ phorefp, const float* __restrict__ phoretype,
                  const float* __restrict__ nangle1, const float* __restrict__ nangle2,
                  const float* __restrict__ cross_h, const float* __restrict__ cross_fm, DpSmallWeights sw,
                  const float* __restrict__ sc, float* __restrict__ tw_scratch, float* __restrict__ cross_emb,
                  float* __restrict__ cross_sh, float* __restrict__ cross_nsh) {
    extern __shared__ float sden[];        // per-atom softmax denominators
    __shared__ __align__(16) float w_rbf[20 * 20], w3[20 * 20];       // transposed: [c][o]
    __shared__ float b3[20], cst[20], cd_w0[10 * 20], cd_b0[10], cd_w3[10];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0, p0 = ph_ptr[g], P = ph_ptr[g + 1] - p0;
    const int c0 = cross_ptr[g];
    // gridDim.y CTAs share a graph: each takes a contiguous range of ligand atoms with all their P edges
    const int chunk = (n + (int)gridDim.y - 1) / (int)gridDim.y;
    const int al_lo = min((int)blockIdx.y * chunk, n), al_hi = min(al_lo + chunk, n);
    const int k_lo = al_lo * P, ne = al_hi * P;
    for (int i = threadIdx.x; i < 400; i += blockDim.x) {
        const int c = i / 20, o = i % 20;
        w_rbf[i] = sw.cross_edge.w0[o * 73 + 20 + c];
        w3[i] = sw.cross_edge.w3[o * 20 + c];
    }
    for (int i = threadIdx.x; i < 200; i += blockDim.x) cd_w0[i] = sw.cdt.w0[i];
    for (int i = threadIdx.x; i < 20; i += blockDim.x) { b3[i] = sw.cross_edge.b3[i]; cst[i] = sc[SC_CROSS_EDGE + i]; }
    for (int i = threadIdx.x; i < 10; i += blockDim.x) { cd_b0[i] = sw.cdt.b0[i]; cd_w3[i] = sw.cdt.w3[i]; }
    __syncthreads();
    // pass 1: distance features, edge embedding, total_weight
    for (int k = k_lo + threadIdx.x; k < ne; k += blockDim.x) {
        const int a = a0 + k / P, p = p0 + k % P, e = c0 + k;
        const float vx = ppos[p * 3] - lpos[a * 3], vy = ppos[p * 3 + 1] - lpos[a * 3 + 1], vz = ppos[p * 3 + 2] - lpos[a * 3 + 2];
        float rbf[20];
        dp_rbf20(sqrtf(vx * vx + vy * vy + vz * vz), DP_RBF_CROSS, rbf);
        float2 h2[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float4 ch = reinterpret_cast<const float4*>(cross_h + (size_t)e * 20)[q];
            h2[2 * q] = make_float2(cst[4 * q] + ch.x, cst[4 * q + 1] + ch.y);
            h2[2 * q + 1] = make_float2(cst[4 * q + 2] + ch.z, cst[4 * q + 3] + ch.w);
        }
        dp_dense20<20>(rbf, w_rbf, h2);
        dp_mlp20_out_t(h2, w3, b3, cross_emb + (size_t)e * 20);
        float dsum = sw.cdt.b3[0];
#pragma unroll 2
        for (int o = 0; o < 10; ++o) {
            float acc = cd_b0[o];
#pragma unroll
            for (int c = 0; c < 20; ++c) acc = fmaf(rbf[c], cd_w0[o * 20 + c], acc);
            dsum = fmaf(fmaxf(acc, 0.f), cd_w3[o], dsum);
        }
        tw_scratch[e] = cross_fm[e] * dp_softplus(dsum) * c_dp.scaler;       // smp:801-811
    }
    __syncthreads();
    // per-atom denominator of the 'phore' atom weight (smp:839; no max-subtraction, like the reference)
    for (int a = al_lo + threadIdx.x; a < al_hi; a += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += expf(tw_scratch[c0 + a * P + p]);
        sden[a] = s;
    }
    __syncthreads();
    // pass 2: gated edge vector SH + angle-matched normal SH
    for (int k = k_lo + threadIdx.x; k < ne; k += blockDim.x) {
        const int al = k / P, a = a0 + al, p = p0 + k % P, e = c0 + k;
        const float tw = tw_scratch[e];
        float dir = sw.pdt.b3[0];
        for (int o = 0; o < 11; ++o) dir = fmaf(dp_leaky(fmaf(tw, sw.pdt.w0[o], sw.pdt.b0[o])), sw.pdt.w3[o], dir);
        dir = dp_leaky(dir);
        const float sgn = dir < 0.f ? -1.f : 1.f;                              // pow(-1, (x < 0))
        const float aw = expf(tw) / sden[al];                                   // multiple=False -> total_weight = atom_weight
        float vx = (ppos[p * 3] - lpos[a * 3]) * sgn * aw, vy = (ppos[p * 3 + 1] - lpos[a * 3 + 1]) * sgn * aw,
              vz = (ppos[p * 3 + 2] - lpos[a * 3 + 2]) * sgn * aw;
        float sh[9];
        dp_sh9(vx, vy, vz, sh);
#pragma unroll
        for (int i = 0; i < 9; ++i) cross_sh[(size_t)e * DP_SH + i] = sh[i];
        // angle_match (smp:874-889)
        const bool is_ex = phoretype[p * 11 + 10] == 1.0f;
        float lx = 0.f, ly = 0.f, lz = 0.f, sag = 0.f, a1 = 0.f, a2 = 0.f;
        if (!is_ex) {
            for (int t = 0; t < 11; ++t) {
                const float ag = phoretype[p * 11 + t] * phorefp[a * 11 + t];
                lx = fmaf(ag, lnorm[a * 33 + t * 3], lx);
                ly = fmaf(ag, lnorm[a * 33 + t * 3 + 1], ly);
                lz = fmaf(ag, lnorm[a * 33 + t * 3 + 2], lz);
                sag += ag;
                a1 = fmaf(ag, nangle1[a * 11 + t], a1);
                a2 = fmaf(ag, nangle2[a * 11 + t], a2);
            }
        }
        const float px = pnorm[p * 3], py = pnorm[p * 3 + 1], pz = pnorm[p * 3 + 2];
        float cx = ly * pz - lz * py, cy = lz * px - lx * pz, cz = lx * py - ly * px;
        if (!c_dp.no_clamp) { cx = fmaxf(cx, 1e-12f); cy = fmaxf(cy, 1e-12f); cz = fmaxf(cz, 1e-12f); }   // H10
        cx *= sag; cy *= sag; cz *= sag;
        const float cn = fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);
        cx /= cn; cy /= cn; cz /= cn;
        const float ln = sqrtf(lx * lx + ly * ly + lz * lz), pn = sqrtf(px * px + py * py + pz * pz);
        const float mx = lx * pn - ln * px, my = ly * pn - ln * py, mz = lz * pn - ln * pz;
        const float qx = lx * pn + ln * px, qy = ly * pn + ln * py, qz = lz * pn + ln * pz;
        const float cur = 2.0f * atan2f(sqrtf(mx * mx + my * my + mz * mz), sqrtf(qx * qx + qy * qy + qz * qz));
        const float d1 = cur - a1, d2 = cur - a2;
        const float nr = (fabsf(d2) < fabsf(d1)) ? d2 : d1;
        dp_sh9(cx * nr, cy * nr, cz * nr, sh);
#pragma unroll
        for (int i = 0; i < 9; ++i) cross_nsh[(size_t)e * DP_SH + i] = sh[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// node embeddings
// ---------------------------------------------------------------------------------------------------------------
__global__ void node_embed_kernel(const float* __restrict__ lpos, const float* __restrict__ ppos,
                                  const int* __restrict__ lig_batch, const int* __restrict__ ph_ptr,
                                  const float* __restrict__ phoretype, const float* __restrict__ lig_static,
                                  const float* __restrict__ ph_static, int n_lig, int n_ph, DpSmallWeights sw,
                                  const float* __restrict__ sc, float* __restrict__ lig_h0, float* __restrict__ ph_h0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lig) {
        const int g = lig_batch[i];
        float dmin = 3.0e38f;
        bool any = false;
        for (int p = ph_ptr[g]; p < ph_ptr[g + 1]; ++p) {
            if (phoretype[p * 11 + 10] == 1.0f) {
                const float dx = lpos[i * 3] - ppos[p * 3], dy = lpos[i * 3 + 1] - ppos[p * 3 + 1], dz = lpos[i * 3 + 2] - ppos[p * 3 + 2];
                dmin = fminf(dmin, sqrtf(dx * dx + dy * dy + dz * dz));
                any = true;
            }
        }
        if (!any) dmin = 1e9f;                                  // smp:912 (H8)
        int flag[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) flag[c] = dmin <= c_dp.clash_cutoff[c] ? 1 : 0;
        for (int o = 0; o < 20; ++o) {
            float v = lig_static[(size_t)i * 20 + o] + sc[SC_LIG_NODE + o];
            float b = 0.f;
#pragma unroll
            for (int c = 0; c < 5; ++c) b += sw.boarder_tables[(c * 2 + flag[c]) * 20 + o];
            b += fmaf(dmin, sw.boarder_w[o], sw.boarder_b[o]);
            lig_h0[(size_t)i * 20 + o] = v + b;
        }
    } else if (i < n_lig + n_ph) {
        const int p = i - n_lig;
        for (int o = 0; o < 20; ++o) ph_h0[(size_t)p * 20 + o] = ph_static[(size_t)p * 20 + o] + sc[SC_PH_NODE + o];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// centre graph + tr / rot heads
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
center_step_kernel(const float* __restrict__ lpos, const int* __restrict__ lig_ptr, DpSmallWeights sw,
                   const float* __restrict__ sc, float* __restrict__ c_emb, float* __restrict__ c_sh) {
    __shared__ float cen[3];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0;
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += lpos[(size_t)(a0 + i) * 3 + threadIdx.x];      // index_add_ order
        cen[threadIdx.x] = s / (float)n;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int a = a0 + i;
        const float vx = lpos[a * 3] - cen[0], vy = lpos[a * 3 + 1] - cen[1], vz = lpos[a * 3 + 2] - cen[2];
        float rbf[20], h[20], sh[9];
        dp_rbf20(sqrtf(vx * vx + vy * vy + vz * vz), DP_RBF_CENTER, rbf);
        for (int o = 0; o < 20; ++o) {
            float acc = sc[SC_CENTER + o];
#pragma unroll
            for (int c = 0; c < 20; ++c) acc = fmaf(rbf[c], sw.center_edge.w0[o * 40 + c], acc);
            h[o] = acc;
        }
        dp_mlp20_out(h, sw.center_edge.w3, sw.center_edge.b3, c_emb + (size_t)a * 20);
        dp_sh9(vx, vy, vz, sh);
#pragma unroll
        for (int k = 0; k < 9; ++k) c_sh[(size_t)a * DP_SH + k] = sh[k];
    }
}

__global__ void score_head_kernel(const float* __restrict__ gpred, int n_graphs, DpSmallWeights sw,
                                  const float* __restrict__ sc, float* __restrict__ tr_out, float* __restrict__ rot_out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_graphs) return;
    const float* o = gpred + (size_t)g * 12;
    float tr[3] = {o[0] + o[6], o[1] + o[7], o[2] + o[8]};
    float rot[3] = {o[3] + o[9], o[4] + o[10], o[5] + o[11]};
    const float tn = sqrtf(tr[0] * tr[0] + tr[1] * tr[1] + tr[2] * tr[2]);
    const float rn = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
    float ts = sw.tr_final.b3[0], rs = sw.rot_final.b3[0];
    for (int h = 0; h < 20; ++h) {
        ts = fmaf(fmaxf(fmaf(tn, sw.tr_final.w0[h * 21], sc[SC_TR + h]), 0.f), sw.tr_final.w3[h], ts);
        rs = fmaf(fmaxf(fmaf(rn, sw.rot_final.w0[h * 21], sc[SC_ROT + h]), 0.f), sw.rot_final.w3[h], rs);
    }
    for (int k = 0; k < 3; ++k) {
        tr_out[g * 3 + k] = tr[k] / tn * ts * sc[SC_INV_TR_SIGMA];
        rot_out[g * 3 + k] = rot[k] / rn * rs * sc[SC_SO3_NORM];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// torsion graph: rotatable-bond centres -> atoms within 5 A (cap 32, lowest index first)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tor_count_kernel(const float* __restrict__ lpos, const int* __restrict__ lig_ptr, const int* __restrict__ rot_ptr,
                 const int* __restrict__ rot_u, const int* __restrict__ rot_v, int* __restrict__ deg, int* __restrict__ gcount) {
    __shared__ int stot;
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0, r0 = rot_ptr[g], nr = rot_ptr[g + 1] - r0;
    if (threadIdx.x == 0) stot = 0;
    __syncthreads();
    const float r2 = c_dp.lig_radius * c_dp.lig_radius;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int u = rot_u[r0 + r], v = rot_v[r0 + r];
        const float cx = (lpos[u * 3] + lpos[v * 3]) / 2, cy = (lpos[u * 3 + 1] + lpos[v * 3 + 1]) / 2, cz = (lpos[u * 3 + 2] + lpos[v * 3 + 2]) / 2;
        int d = 0;
        for (int i = 0; i < n && d < c_dp.max_neighbors; ++i)
            if (dp_dist2(cx, cy, cz, lpos[(a0 + i) * 3], lpos[(a0 + i) * 3 + 1], lpos[(a0 + i) * 3 + 2]) < r2) ++d;
        deg[r0 + r] = d;
        atomicAdd(&stot, d);
    }
    __syncthreads();
    if (threadIdx.x == 0) gcount[g] = stot;
}

__global__ void __launch_bounds__(128)
tor_fill_kernel(const float* __restrict__ lpos, const int* __restrict__ lig_ptr, const int* __restrict__ rot_ptr,
                const int* __restrict__ rot_u, const int* __restrict__ rot_v, const int* __restrict__ deg,
                const int* __restrict__ gstart, int n_graphs, DpSmallWeights sw, int* __restrict__ seg_ptr,
                int* __restrict__ e_atom, int* __restrict__ e_u, int* __restrict__ e_v, float* __restrict__ e_emb,
                float* __restrict__ e_sh) {
    __shared__ int soff[LG_MAXN + 1];
    __shared__ __align__(16) float sw0[400], sw3[400];               // transposed: [c][o]
    __shared__ float sb0[20], sb3[20];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0, r0 = rot_ptr[g], nr = rot_ptr[g + 1] - r0;
    for (int i = threadIdx.x; i < 400; i += blockDim.x) {
        const int c = i / 20, o = i % 20;
        sw0[i] = sw.final_edge.w0[o * 20 + c];
        sw3[i] = sw.final_edge.w3[o * 20 + c];
    }
    for (int i = threadIdx.x; i < 20; i += blockDim.x) { sb0[i] = sw.final_edge.b0[i]; sb3[i] = sw.final_edge.b3[i]; }
    if (threadIdx.x == 0) {
        int run = gstart[g];
        for (int r = 0; r < nr; ++r) { soff[r] = run; run += deg[r0 + r]; }
        soff[nr] = run;
    }
    __syncthreads();
    // gridDim.y CTAs share a graph: each takes a contiguous range of rotatable bonds
    const int chunk = (nr + (int)gridDim.y - 1) / (int)gridDim.y;
    const int r_lo = min((int)blockIdx.y * chunk, nr), r_hi = min(r_lo + chunk, nr);
    for (int r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) seg_ptr[r0 + r] = soff[r];
    if (g == n_graphs - 1 && blockIdx.y == 0 && threadIdx.x == 0) seg_ptr[r0 + nr] = soff[nr];
    const float r2 = c_dp.lig_radius * c_dp.lig_radius;
    // topology: thread per rotatable bond (torch_cluster.radius: atoms within 5 A of the bond centre, lowest indices first)
    // (WARP per bond: 32 candidate atoms per step, ballot for the index order and the neighbour cap; a thread per bond kept ~7 of
    //  128 threads busy on a serial scan while the rest waited at the barrier below: 23 % of the kernel's samples)
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
        const unsigned lt = (1u << lane) - 1u;
        for (int r = r_lo + warp; r < r_hi; r += nwarp) {
            const int u = rot_u[r0 + r], v = rot_v[r0 + r];
            const float cx = (lpos[u * 3] + lpos[v * 3]) / 2, cy = (lpos[u * 3 + 1] + lpos[v * 3 + 1]) / 2, cz = (lpos[u * 3 + 2] + lpos[v * 3 + 2]) / 2;
            const int o = soff[r];
            int d = 0;
            for (int i0 = 0; i0 < n && d < c_dp.max_neighbors; i0 += 32) {
                const int i = i0 + lane, a = a0 + i;
                const bool hit = i < n && dp_dist2(cx, cy, cz, lpos[a * 3], lpos[a * 3 + 1], lpos[a * 3 + 2]) < r2;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                const int rank = d + __popc(m & lt);
                if (hit && rank < c_dp.max_neighbors) { e_atom[o + rank] = a; e_u[o + rank] = u; e_v[o + rank] = v; }
                d += __popc(m);
            }
        }
    }
    __syncthreads();
    // features: thread per edge (the 20 -> 20 -> 20 embedding MLP dominates; all 128 threads busy instead of one per bond)
    for (int e = soff[r_lo] + threadIdx.x; e < soff[r_hi]; e += blockDim.x) {
        const int a = e_atom[e], u = e_u[e], v = e_v[e];
        const float cx = (lpos[u * 3] + lpos[v * 3]) / 2, cy = (lpos[u * 3 + 1] + lpos[v * 3 + 1]) / 2, cz = (lpos[u * 3 + 2] + lpos[v * 3 + 2]) / 2;
        float b2[9];
        dp_sh9(lpos[v * 3] - lpos[u * 3], lpos[v * 3 + 1] - lpos[u * 3 + 1], lpos[v * 3 + 2] - lpos[u * 3 + 2], b2);   // Y2 = b2[4..8]
        const float vx = lpos[a * 3] - cx, vy = lpos[a * 3 + 1] - cy, vz = lpos[a * 3 + 2] - cz;
        float rbf[20], sh[9], o7[8];
        dp_rbf20(sqrtf(vx * vx + vy * vy + vz * vz), DP_RBF_LIG, rbf);
        float2 h2[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) h2[j] = make_float2(sb0[2 * j], sb0[2 * j + 1]);
        dp_dense20<20>(rbf, sw0, h2);
        dp_mlp20_out_t(h2, sw3, sb3, e_emb + (size_t)e * 20);
        dp_sh9(vx, vy, vz, sh);
        dp_fulltp7(sh, b2 + 4, o7);
        o7[7] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) e_sh[(size_t)e * 8 + k] = o7[k];
    }
}

__global__ void tor_head_kernel(const float* __restrict__ tor_feat, int n_rot, DpSmallWeights sw, const float* __restrict__ sc,
                                float* __restrict__ tor_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rot) return;
    float x[40];
#pragma unroll
    for (int c = 0; c < 40; ++c) x[c] = tor_feat[(size_t)r * 40 + c];
    float s = 0.f;
    for (int h = 0; h < 20; ++h) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 40; ++c) acc = fmaf(x[c], sw.tor_w0[h * 40 + c], acc);
        s = fmaf(tanhf(acc), sw.tor_w3[h], s);
    }
    tor_out[r] = s * sc[SC_SQRT_TORUS];
}
