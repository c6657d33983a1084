// dp_edge_mlp (tensor-core path): the same per-edge weight MLP as edge_mlp.cuh, with the second (97 % of FLOPs)
// layer  w[128 edges, W] = h[128, 64] . W2aug[64, W]  on the 5th-gen tensor cores:
//   tcgen05.mma.cta_group::1.kind::tf32, M=128, N=128, K=8, accumulators in TMEM (2 stages x 128 columns),
//   operands in shared memory in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B),
//   B (second-layer weights, bias folded in as K row 60) streamed per 128-column chunk by cp.async.bulk (TMA, 1-D)
//   into a 2-stage ring, mbarrier full/empty pipeline, one elected thread issues the MMAs, 4 warps drain TMEM
//   with tcgen05.ld (thread = edge row) and store to HBM.
// fp32 parity: both operands are split x = hi + lo with hi = tf32(x) (round-to-nearest) and three MMAs
// hi*hi + hi*lo + lo*hi are accumulated in fp32 ("3xTF32"), relative error ~2^-21 per product.
// With 3 x 8 MMAs per chunk the tensor pipe needs ~1.6k clk per 128x128 chunk while the 64 KB of output need
// ~4k clk of this SM's share of HBM write bandwidth: the kernel is HBM-write-bound by design.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "edge_mlp.cuh"

#define TC_WORKERS 256              // warps 0-7: two threads per edge row (row = tid & 127, half = tid >> 7)
#define TC_THREADS 288              // + warp 8: TMEM owner, TMA + MMA issuer
#define TC_BN 128                   // columns per chunk
#define TC_K 64                     // padded hidden (+bias) dimension
#define TC_OPER_BYTES (128 * TC_K * 4)          // one 128 x 64 fp32 operand tile = 32 KB
// per-warp 32x32 fp32 transpose tiles (XOR-swizzled 16-B chunks) alias the first-layer weights, which are dead by then
#define TC_SMEM_BYTES (2 * TC_OPER_BYTES + 2 * 2 * TC_OPER_BYTES + 8 * 32 * 32 * 4 + 64 * 4 + 256)

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (-> CUDA error reported through the ABI) instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = tc_smem(bar);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tc_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem(dst)),
                 "l"(src), "r"(bytes), "r"(tc_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar)) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = bytes between 8-row groups, LBO = bytes
// between the two 16-B K chunks an MMA (K = 8 tf32) consumes.  (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float tc_tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ long long* g_tc_dbg = nullptr;     // profiling aid: per-phase clock64() stamps of CTA 200 (tools/tc_phase_probe.py)
#define TC_STAMP(i) do { if (g_tc_dbg && blockIdx.x == 200 && (tid & 127) == 0) g_tc_dbg[(tid >> 7) * 16 + (i)] = clock64(); } while (0)


// ---------------------------------------------------------------------------------------------------------------
// pass 1: hidden activations h = ReLU(W1 . attr + b1) for every edge, row-major [edge][64] (k = 60 carries the
// constant 1 that multiplies the bias row of W2aug, k = 61..63 zero).  256 B per edge (3-10 % of the per-edge weight stream); keeps the gather and
// the CUDA-core layer out of the tensor-core kernel, whose shared memory is full and cannot overlap them.
// ---------------------------------------------------------------------------------------------------------------
#define EH_THREADS 128
#define EH_TILE 64                  // edges per CTA iteration
#define EH_LD 68                    // padded row length of the transposed attribute tile
__global__ void __launch_bounds__(EH_THREADS, 6) edge_hidden_kernel(EdgeMlpArgs a, float* __restrict__ himg) {
    // register-tiled SGEMM  h[64 edges, 64] = attr[64, 60] . W1T[60, 64]: a thread owns 4 edges x 8 outputs, per k it
    // reads 4 attributes (one float4 of the k-major tile) and 8 weights (two broadcast float4) for 32 FMAs.  Small tiles
    // (32 KB of shared memory, <= 80 registers) keep 6-7 CTAs resident per SM so that the two dependent gather round
    // trips (index -> row) of one CTA overlap the FMA phase of the others.
    __shared__ __align__(16) float w1t[60 * 64];          // [k][o], o >= 60 zero
    __shared__ __align__(16) float attrT[60 * EH_LD];     // [c][m]
    const int tid = threadIdx.x;
    const int E = a.n_edges_dev ? *a.n_edges_dev : a.n_edges;
    const int ntiles = (E + EH_TILE - 1) / EH_TILE;
    if ((int)blockIdx.x >= ntiles) return;
    for (int i = tid; i < 60 * 64; i += EH_THREADS) {
        const int k = i >> 6, o = i & 63;
        w1t[i] = o < 60 ? a.w1[o * 60 + k] : 0.f;
    }
    const int mg = tid & 15, og = tid >> 4, m0 = 4 * mg, o0 = 8 * og;
    float bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias[j] = (o0 + j < 60) ? a.b1[o0 + j] : 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {          // persistent: W1 is staged once per CTA
        const int e0 = tile * EH_TILE;
        __syncthreads();                                                     // previous tile fully consumed
        {
            constexpr int NT = EH_TILE * 30 / EH_THREADS;                    // 15 float2 items per thread, coalesced by part
            int ridx[NT], ridx2[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int i = tid + EH_THREADS * t, m = i / 30, part = (i % 30) / 10;
                const int e = min(e0 + m, E - 1);
                ridx[t] = part == 0 ? (a.perm ? a.perm[e] : e) : (part == 1 ? a.idxB[e] : a.idxC[e]);
                ridx2[t] = (part == 2 && a.idxC2) ? a.idxC2[e] : -1;
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int i = tid + EH_THREADS * t, m = i / 30, q = i % 30, part = q / 10, c = (q % 10) * 2;
                const float* src = part == 0 ? a.emb + (size_t)ridx[t] * 20
                                             : (part == 1 ? a.tb + (size_t)ridx[t] * a.strideB : a.tc + (size_t)ridx[t] * a.strideC);
                float2 v = *reinterpret_cast<const float2*>(src + c);
                if (ridx2[t] >= 0) {
                    const float2 v2 = *reinterpret_cast<const float2*>(a.tc + (size_t)ridx2[t] * a.strideC + c);
                    v.x += v2.x; v.y += v2.y;
                }
                attrT[(2 * q) * EH_LD + m] = v.x;                            // part*20 + c == 2*q
                attrT[(2 * q + 1) * EH_LD + m] = v.y;
            }
        }
        __syncthreads();
        float h[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) h[i][j] = bias[j];
#pragma unroll 4
        for (int k = 0; k < 60; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(attrT + k * EH_LD + m0);
            const float4 b0 = *reinterpret_cast<const float4*>(w1t + k * 64 + o0);
            const float4 b1v = *reinterpret_cast<const float4*>(w1t + k * 64 + o0 + 4);
            const float am[4] = {av.x, av.y, av.z, av.w};
            const float bm[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) h[i][j] = fmaf(am[i], bm[j], h[i][j]);
        }
        // ReLU, constant-1 column (k = 60) and zero padding; row-major [edge][64] (256 B per edge)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = o0 + j;
                r[j] = o < 60 ? fmaxf(h[i][j], 0.f) : (o == 60 ? 1.0f : 0.f);
            }
            float* dst = himg + ((size_t)e0 + m0 + i) * TC_K + o0;
            *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(r[4], r[5], r[6], r[7]);
        }
    }
}

struct EdgeMlpTcArgs {
    EdgeMlpArgs base;        // w2t unused here
    const void* w2img;       // [ceil(W/64)][2 (hi, lo)][8 k-chunks][8 row groups][8 rows][8] fp16 of W2aug * wscale, zero padded
    float inv_wscale;        // 1 / wscale (power of two)
    float* himg;             // scratch: [E rounded up to 128][64] hidden activations (edge_hidden_kernel)
};

// ---------------------------------------------------------------------------------------------------------------
// pass 2: w[256 edges, W] = h . W2aug on tcgen05.  One CTA owns TWO 128-edge tiles.
//  * fp32 parity through a 2-way FP16 split with exact power-of-two scaling: every row of h is scaled so that its
//    largest entry lies in [2^12, 2^13) and W2aug by one per-layer power of two, then x = hi + lo with hi = fp16(x),
//    lo = fp16(x - hi) (22 significant bits, the same as a tf32 hi/lo split) and the three products hi*hi + hi*lo +
//    lo*hi are accumulated in fp32 by kind::f16 MMAs; the scales are undone exactly in the epilogue.  Compared with
//    3xTF32 this halves the MMA time (K = 16 per instruction at the same instruction cost) and the operand bytes.
//  * A operands live in TENSOR MEMORY (tcgen05.st, lane = edge row, 2 tiles x (32 hi + 32 lo) columns of packed
//    half2), B (64-column half-chunks, hi|lo = 16 KB) streams through a 6-stage TMA ring: each weight byte fetched
//    from L2 feeds 256 rows.
//  * Per half-chunk: 2 x 12 MMAs (M=128, N=64, K=16) into 4 accumulator slots (2 stages x 2 tiles x 64 columns); 8
//    epilogue warps drain one 32x32 block each per (half-chunk, tile): tcgen05.ld -> rescale -> 128B-swizzled smem
//    tile -> TMA tensor store.
// ---------------------------------------------------------------------------------------------------------------
#define TC2_BN 64
#define TC2_STAGES 6
#define TC2_B_STAGE_BYTES (2 * TC2_BN * TC_K * 2)                 // hi + lo of a 64 x 64 fp16 block = 16 KB
#define TC2_SMEM_BYTES (TC2_STAGES * TC2_B_STAGE_BYTES + 8 * 32 * 32 * 4 + 256 * 4 + 256)

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ uint32_t tc_pack_half2(__half lo16, __half hi16) {
    return (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16);
}

__global__ void __launch_bounds__(TC_THREADS, 1) edge_mlp_tc_kernel(EdgeMlpTcArgs args, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    const EdgeMlpArgs& a = args.base;
    uint8_t* b_st = tc_smem_raw;                                      // TC2_STAGES x (hi 8 KB | lo 8 KB), fp16
    float* stg_all = reinterpret_cast<float*>(b_st + TC2_STAGES * TC2_B_STAGE_BYTES);     // 8 x [32][32] swizzled tiles
    float* row_scale = stg_all + 8 * 32 * 32;                         // [256] 1 / (row scale * weight scale)
    uint64_t* bars = reinterpret_cast<uint64_t*>(row_scale + 256);
    uint64_t *b_full = bars, *b_empty = bars + TC2_STAGES, *t_full = bars + 2 * TC2_STAGES /*[acc stage][tile]*/,
             *t_empty = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = a.n_edges_dev ? *a.n_edges_dev : a.n_edges;
    const int e0 = blockIdx.x * 256;
    if (e0 >= E) return;
    const int ntile = (E - e0 > 128) ? 2 : 1;
    const int W = a.W, nhc = (W + TC2_BN - 1) / TC2_BN;

    TC_STAMP(0);
    if (tid == 0) {
        for (int i = 0; i < TC2_STAGES; ++i) { tc_mbar_init(&b_full[i], 1); tc_mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { tc_mbar_init(&t_full[i], 1); tc_mbar_init(&t_empty[i], TC_WORKERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    TC_STAMP(1);

    if (tid == TC_WORKERS) {                                          // start streaming the weights right away
        for (int st = 0; st < TC2_STAGES - 1 && st < nhc; ++st) {
            tc_mbar_expect_tx(&b_full[st], TC2_B_STAGE_BYTES);
            tc_bulk_load(b_st + st * TC2_B_STAGE_BYTES, reinterpret_cast<const uint8_t*>(args.w2img) + (size_t)st * TC2_B_STAGE_BYTES,
                         TC2_B_STAGE_BYTES, &b_full[st]);
        }
    }
    if (warp < 8) {
        // A operands -> tensor memory: thread (tile = tid / 128, row = tid % 128) loads its edge's 64 hidden activations,
        // scales the row by a power of two, splits into fp16 hi + lo and stores both as 32 packed TMEM columns of its lane.
        const int tile = tid >> 7, row = tid & 127, wq = warp & 3;
        if (tile < ntile) {
            const int e = min(e0 + tile * 128 + row, E - 1);
            const float4* hp = reinterpret_cast<const float4*>(args.himg + (size_t)e * TC_K);
            float h[64];
            float m = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 v = __ldg(hp + q);
                h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
            // s = 2^(12 - floor(log2 m)): m * s in [2^12, 2^13); m >= 1 always (the bias column holds 1.0)
            const int ex = (int)((__float_as_uint(m) >> 23) & 0xFF) - 127;
            const float sc = __uint_as_float((uint32_t)(127 + 12 - ex) << 23);
            const float isc = __uint_as_float((uint32_t)(127 - 12 + ex) << 23);
            row_scale[tid] = isc * args.inv_wscale;
            uint32_t hi_p[32], lo_p[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float x0 = h[2 * c] * sc, x1 = h[2 * c + 1] * sc;
                const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                hi_p[c] = tc_pack_half2(h0, h1);
                lo_p[c] = tc_pack_half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
            }
            const uint32_t a_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + 256u + (uint32_t)(tile * 64);
            tc_tmem_st32(a_addr, hi_p);                               // hi: columns [0, 32) of the tile's operand block
            tc_tmem_st32(a_addr + 32u, lo_p);                         // lo: columns [32, 64)
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        TC_STAMP(3);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_STAMP(4);

    if (tid == TC_WORKERS) {
        // ================= TMA producer + MMA issuer (single thread) =================
        // instruction descriptor: D=F32 (c_format 1), A=B=F16 (format 0), K-major, N=64, M=128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(TC2_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int hc = 0; hc < nhc; ++hc) {
            const int s = hc % TC2_STAGES, u = hc / TC2_STAGES, as = hc & 1, au = hc >> 1;
            const int nx = hc + TC2_STAGES - 1;                       // keep TC2_STAGES-1 weight loads in flight
            if (nx < nhc) {
                const int s1 = nx % TC2_STAGES, u1 = nx / TC2_STAGES;
                tc_mbar_wait(&b_empty[s1], (u1 & 1) ^ 1);
                tc_mbar_expect_tx(&b_full[s1], TC2_B_STAGE_BYTES);
                tc_bulk_load(b_st + s1 * TC2_B_STAGE_BYTES, reinterpret_cast<const uint8_t*>(args.w2img) + (size_t)nx * TC2_B_STAGE_BYTES,
                             TC2_B_STAGE_BYTES, &b_full[s1]);
            }
            tc_mbar_wait(&b_full[s], u & 1);
            const uint32_t b_hi_s = tc_smem(b_st + s * TC2_B_STAGE_BYTES), b_lo_s = b_hi_s + TC2_B_STAGE_BYTES / 2;
            for (int i = 0; i < ntile; ++i) {
                tc_mbar_wait(&t_empty[as * 2 + i], (au & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi_t = tmem_base + 256u + (uint32_t)(i * 64), a_lo_t = a_hi_t + 32u;
                const uint32_t d = tmem_base + (uint32_t)((as * 2 + i) * TC2_BN);
#pragma unroll
                for (int combo = 0; combo < 3; ++combo) {
                    const uint32_t at = combo == 2 ? a_lo_t : a_hi_t;
                    const uint32_t bs = combo == 1 ? b_lo_s : b_hi_s;
#pragma unroll
                    for (int ks = 0; ks < TC_K / 16; ++ks) {
                        // fp16 K-major no-swizzle: core matrix = 8 rows x 8 halfs; K chunk stride 8 x 128 B, 2 chunks per MMA
                        const uint64_t bd = tc_smem_desc(bs + ks * 2 * 1024, 1024, 128);
                        tc_mma_f16_ts(d, at + (uint32_t)(ks * 8), bd, idesc, (combo | ks) ? 1u : 0u);
                    }
                }
                tc_commit(&t_full[as * 2 + i]);   // accumulator slot complete
            }
            tc_commit(&b_empty[s]);               // weight stage reusable once all MMAs above have read it
        }
    } else if (warp < 8) {
        // ================= epilogue: TMEM -> registers -> swizzled smem tile -> TMA tensor store =================
        // warp w drains the 32x32 block (rows 32*(w%4).., columns 32*(w/4)..) of every (half-chunk, tile) accumulator.
        float* stg = stg_all + warp * 32 * 32;
        const int wq = warp & 3, ch = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        const float rs[2] = {row_scale[wq * 32 + lane], row_scale[128 + wq * 32 + lane]};
        for (int hc = 0; hc < nhc; ++hc) {
            const int as = hc & 1, au = hc >> 1;
            for (int i = 0; i < ntile; ++i) {
                tc_mbar_wait(&t_full[as * 2 + i], au & 1);
                if (hc < 2 && i == 0) TC_STAMP(5 + 2 * hc);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float v[32];
                tc_tmem_ld32(lane_base + (uint32_t)((as * 2 + i) * TC2_BN + ch * 32), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tc_mbar_arrive(&t_empty[as * 2 + i]);                // accumulator slot is free again (data in registers)
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store finished reading the tile
                }
                __syncwarp();
                const float r = rs[i];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(v[4 * j] * r, v[4 * j + 1] * r, v[4 * j + 2] * r, v[4 * j + 3] * r);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                const int n0 = hc * TC2_BN + ch * 32;
                if (lane == 0 && n0 < W) {
                    // 32 x 32 fp32 box, 128-byte swizzle (== the XOR pattern above); rows/columns beyond the tensor are clipped
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&out_map),
                                 "r"(n0), "r"(e0 + i * 128 + wq * 32), "r"(tc_smem(stg))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (hc < 2 && i == 0) TC_STAMP(6 + 2 * hc);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        TC_STAMP(11);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_STAMP(12);
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

static int edge_mlp_tc_launch(const EdgeMlpTcArgs& t, cudaStream_t st) {
    const EdgeMlpArgs& a = t.base;
    if (a.n_edges <= 0) return DP_OK;
    if (a.in_dim != 60 || a.hid != 60 || (a.W % 4) != 0 || t.w2img == nullptr || a.tc == nullptr || t.himg == nullptr) {
        dp_set_error("dp_edge_mlp_tc: unsupported shape in=%d hid=%d W=%d", a.in_dim, a.hid, a.W);
        return DP_ERR_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES);
        attr_set = true;
    }
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            dp_set_error("dp_edge_mlp_tc: cuTensorMapEncodeTiled is not available");
            return DP_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)a.W, (cuuint64_t)a.n_edges};
    const cuuint64_t gstride[1] = {(cuuint64_t)a.W * 4};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dp_set_error("dp_edge_mlp_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return DP_ERR_CUDA;
    }
    dim3 grid((a.n_edges + 127) / 128);
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    edge_hidden_kernel<<<dim3(min((a.n_edges + EH_TILE - 1) / EH_TILE, n_sm * 6)), EH_THREADS, 0, st>>>(a, t.himg);
    edge_mlp_tc_kernel<<<dim3((a.n_edges + 255) / 256), TC_THREADS, TC2_SMEM_BYTES, st>>>(t, map);
    return dp_check_launch("edge_mlp_tc");
}
