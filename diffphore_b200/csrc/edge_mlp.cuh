// dp_edge_mlp: per-edge tensor-product weights  w[e, :] = Linear2(ReLU(Linear1(edge_attr_e)))   (K5)
//
// Replaces TensorProductConvLayer.fc (score_model_phore.py:125-130,137) together with the torch.cat that
// assembles edge_attr = [edge_embedding | node_a[:ns] | node_b[:ns]] (score_model_phore.py:678,682,692,695,337,368):
// the three 20-wide parts are gathered straight from their sources, never materialised in HBM.
//
// FP32 FFMA tile kernel (parity-safe): one CTA = 128 edges x all W output columns, swept in 128-column chunks of
// the transposed second-layer weights (W2T[k, n], bias as row k = hid) that are double-buffered in shared memory
// with cp.async; 8x8 register tile per thread; hidden activations kept k-major in shared memory.
#pragma once
#include "common.cuh"

#define EM_TM 128
#define EM_TN 128
#define EM_THREADS 256
#define EM_KMAX 61            // hid + 1 (bias row)

struct EdgeMlpArgs {
    const float* emb;         // [*, 20] edge embedding, row = perm ? perm[e] : e
    const int* perm;
    const float* tb;          // part B rows (first 20 features of a node tensor), row stride strideB
    const int* idxB;
    int strideB;
    const float* tc;          // part C (nullptr when in_dim == 40)
    const int* idxC;
    const int* idxC2;         // optional second row added to part C (tor_bond_conv: node[b0] + node[b1])
    int strideC;
    const float* w1;          // [hid, in_dim]  (nn.Linear weight)
    const float* b1;          // [hid]
    const float* w2t;         // [hid + 1, W]   (transposed nn.Linear weight, last row = bias)
    int in_dim, hid, W;
    const int* n_edges_dev;   // optional device-side edge count (dynamic graphs); else n_edges
    int n_edges;
    float* out;               // [E, W]
};

__device__ __forceinline__ void em_cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void em_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void em_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void em_load_chunk(float* bs, const float* __restrict__ w2t, int K1, int W, int n0, int tid) {
    // bs[k][0..127] <- w2t[k*W + n0 + 0..127], zero-filled beyond W
    const int nvec = K1 * (EM_TN / 4);
    for (int i = tid; i < nvec; i += EM_THREADS) {
        const int k = i / (EM_TN / 4), c = (i % (EM_TN / 4)) * 4;
        const bool ok = (n0 + c) < W;
        const float* src = w2t + (size_t)k * W + (ok ? n0 + c : 0);
        em_cp_async16(bs + k * EM_TN + c, src, ok);
    }
}

__global__ void __launch_bounds__(EM_THREADS, 2) edge_mlp_kernel(EdgeMlpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int K1 = a.hid + 1;
    float* hT = smem;                               // [K1][128]
    float* bs0 = hT + EM_KMAX * EM_TM;              // [K1][128] x 2
    float* bs1 = bs0 + EM_KMAX * EM_TN;
    float* w1s = bs1 + EM_KMAX * EM_TN;             // [hid][in_dim]
    float* attr = bs0;                              // [128][in_dim+1], aliased onto the B buffers before the GEMM
    const int tid = threadIdx.x;
    const int E = a.n_edges_dev ? *a.n_edges_dev : a.n_edges;
    const int e0 = blockIdx.x * EM_TM;
    if (e0 >= E) return;
    const int in_dim = a.in_dim, hid = a.hid, AS = in_dim + 1;

    for (int i = tid; i < hid * in_dim; i += EM_THREADS) w1s[i] = a.w1[i];
    // ---- gather edge_attr rows (float2 granularity: node rows are only 8-byte aligned for D = 50)
    const int parts = in_dim / 20, per_edge = parts * 10;
    for (int i = tid; i < EM_TM * per_edge; i += EM_THREADS) {
        const int m = i / per_edge, q = i % per_edge, part = q / 10, c = (q % 10) * 2;
        const int e = min(e0 + m, E - 1);
        float2 v;
        if (part == 0) {
            const int r = a.perm ? a.perm[e] : e;
            v = *reinterpret_cast<const float2*>(a.emb + (size_t)r * 20 + c);
        } else if (part == 1) {
            v = *reinterpret_cast<const float2*>(a.tb + (size_t)a.idxB[e] * a.strideB + c);
        } else {
            v = *reinterpret_cast<const float2*>(a.tc + (size_t)a.idxC[e] * a.strideC + c);
            if (a.idxC2) {
                float2 v2 = *reinterpret_cast<const float2*>(a.tc + (size_t)a.idxC2[e] * a.strideC + c);
                v.x += v2.x; v.y += v2.y;
            }
        }
        attr[m * AS + part * 20 + c] = v.x;
        attr[m * AS + part * 20 + c + 1] = v.y;
    }
    __syncthreads();
    // ---- first layer + ReLU -> hT[k][m]
    {
        const int m = tid & (EM_TM - 1), h0 = tid >> 7;        // 2 threads per edge
        float x[60];
#pragma unroll
        for (int c = 0; c < 60; ++c) x[c] = (c < in_dim) ? attr[m * AS + c] : 0.f;
        for (int h = h0; h < hid; h += 2) {
            float s = a.b1[h];
            const float* wr = w1s + h * in_dim;
#pragma unroll
            for (int c = 0; c < 60; ++c)
                if (c < in_dim) s = fmaf(x[c], wr[c], s);
            hT[h * EM_TM + m] = fmaxf(s, 0.f);
        }
        if (h0 == 0) hT[hid * EM_TM + m] = 1.0f;
    }
    __syncthreads();
    // ---- second layer: [128 x K1] x [K1 x W]
    const int W = a.W, nchunks = (W + EM_TN - 1) / EM_TN;
    em_load_chunk(bs0, a.w2t, K1, W, 0, tid);
    em_cp_commit();
    const int tx = tid & 15, ty = tid >> 4;
    for (int ch = 0; ch < nchunks; ++ch) {
        float* bs = (ch & 1) ? bs1 : bs0;
        if (ch + 1 < nchunks) {
            em_load_chunk((ch & 1) ? bs0 : bs1, a.w2t, K1, W, (ch + 1) * EM_TN, tid);
            em_cp_commit();
            em_cp_wait<1>();
        } else {
            em_cp_wait<0>();
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int k = 0; k < K1; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(hT + k * EM_TM + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(hT + k * EM_TM + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + k * EM_TN + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + k * EM_TN + 64 + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        const int n0 = ch * EM_TN;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            const int e = e0 + m;
            if (e < E) {
                float* orow = a.out + (size_t)e * W + n0;
                const int c0 = tx * 4, c1 = 64 + tx * 4;
                if (n0 + c0 < W) *reinterpret_cast<float4*>(orow + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                if (n0 + c1 < W) *reinterpret_cast<float4*>(orow + c1) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
            }
        }
        __syncthreads();
    }
}

static const int EM_SMEM_BYTES = (3 * EM_KMAX * 128 + 60 * 60) * (int)sizeof(float);

static int edge_mlp_launch(const EdgeMlpArgs& a, cudaStream_t st) {
    if (a.n_edges <= 0) return DP_OK;
    if ((a.in_dim != 40 && a.in_dim != 60) || a.hid > 60 || (a.W % 4) != 0) {
        dp_set_error("dp_edge_mlp: unsupported shape in=%d hid=%d W=%d", a.in_dim, a.hid, a.W);
        return DP_ERR_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(edge_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_SMEM_BYTES);
        attr_set = true;
    }
    dim3 grid((a.n_edges + EM_TM - 1) / EM_TM);
    edge_mlp_kernel<<<grid, EM_THREADS, EM_SMEM_BYTES, st>>>(a);
    return dp_check_launch("edge_mlp");
}
