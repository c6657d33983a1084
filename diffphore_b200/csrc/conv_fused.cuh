// dp_conv_fused: one TensorProductConvLayer (score_model_phore.py:125-149 plus the torch.cat of its edge attributes,
// :678-695, 368) in one kernel, WITHOUT materialising per-edge hidden activations or tensor-product weights in HBM.
//
//   h[e]        = ReLU(W1 . [emb | node_B | node_C](e) + b1)   tcgen05 tensor cores (K5, layer 1)
//   w[e, 0:W]   = [h(e) | 1] . W2aug                            tcgen05 tensor cores, accumulators in tensor memory (K5, layer 2)
//   tp[e]       = FCTP(node_in[gather[e]], sh[e], w[e])          CUDA cores (FFMA2), thread = edge, w read straight from TMEM (K6)
//   out[n]      = BatchNorm(mean_{e in seg(n)} tp[e]) (+ residual)      shared-memory segmented reduction (K7, K8)
//
// The unfused pipeline (dp_edge_mlp_tc -> dp_tp_scatter) writes and re-reads 4*W bytes per edge (2.4-8.8 KB): both
// of its kernels are HBM-bound on that stream.  Here the weights only ever exist as a [128 edges x 100 columns] fp32
// tile in tensor memory, which turns the convolution into a tensor-pipe-bound kernel ("a K5->K6-fused
// kernel never materialises weights: report against the FLOP roofline").
//
// Work decomposition (one persistent CTA per SM, 352 threads, all 512 TMEM columns)
//   * Edges are grouped by output node (CSR seg_ptr).  The unit of work is a PAIR TILE: a run of whole nodes with
//     <= 256 edges (tile_node[], built by dp_build_tiles or on the host), so the edge->node reduction never leaves the
//     CTA: no atomics, the sum over a node's edges is sequential in edge order => bit-identical results for any batch
//     composition.  Its first 128 edges are MMA tile 0, the rest MMA tile 1 (a node may straddle the two: the reduction
//     runs over the pair's 256 staged rows), so the node alignment costs half a node per 256 rows instead of per 128
//     and nodes of up to 256 edges are supported (a 79-point pharmacophore: 3 atoms x 79 edges fill 237 of 256 rows).
//   * Every weight chunk fetched from L2 by TMA feeds the two M=128 MMA groups of the pair.
//     Warps 0-3 own the 128 TMEM lanes of tile 0, warps 4-7 those of tile 1 (thread = edge for the whole pair); warps
//     8 / 9 issue the MMAs of tile 0 / 1 (converged warp, elect.sync-predicated asm: bare UTCHMMA in SASS), warp 10
//     is the TMA producer.  Item 0 of a pair is the hidden layer (N = 64, weights = one 16 KB ring stage), items
//     1..W/100 are the weight chunks.
//   * fp32 parity: exactly-scaled 2-way FP16 split of both operands (row * 2^s into [2^12, 2^13), hi = fp16(x),
//     lo = fp16(x - hi); hi*hi + hi*lo + lo*hi accumulated in fp32).  A hi lives in tensor memory (tcgen05.st, TS-form
//     MMA), A lo in shared memory (SS-form MMA), B streams through the TMA ring (hi | lo per chunk).
//   * The per-path weight blocks [U x V] are all multiples of 100 columns, so a chunk = 100 consecutive columns of the
//     e3nn weight layout = 5 rows (V = 20) or 10 rows (V = 10) of one path; MMA N = 112 (100 + zero padding).
//     TMEM: 4 accumulator slots x 112 columns (slots t, t + 2 belong to tile t: every barrier has one producer warp
//     and one consumer group), A hi operands in columns 448..511.
//   * Per pair: attributes (prefetched one pair ahead) -> layer-1 operand -> hidden MMA -> ReLU -> layer-2 operand ->
//     node rows gathered under the first chunk MMAs -> chunk loop -> rows staged as float4 -> per-(node, quad)
//     sequential sums, mean, BatchNorm scale/shift, residual.
#pragma once
#include <type_traits>
#include "edge_mlp_tc.cuh"
#include "tp_tables.cuh"

#define CF_WORKERS 256
#define CF_THREADS 384                                // + warps 8, 9: MMA issuers of tile 0 / tile 1, warp 10: TMA producer, warp 11: idle
// Register split (setmaxnreg works per warpgroup, hence the 12th warp): the kernel starts with 168 registers per thread; the
// auxiliary warpgroup keeps CF_REG_AUX and hands the rest to the two worker warpgroups (256 x 216 + 128 x 72 = 64512 <= 65536).  At
// 168 the worker loop spilled ~50 accumulator registers per pair tile, and a spill is an L2 round trip here (L1 is ~20 KB next to
// 227 KB of shared memory and thrashed by the gathers).
#define CF_REG_WORK 216
#define CF_REG_AUX 72
#define CF_CHUNK 100                                  // weight columns per chunk
#define CF_N 112                                      // MMA N (chunk padded to a multiple of 16)
#define CF_B_HALF (CF_N * TC_K * 2)                   // one fp16 operand image (hi or lo) of a chunk: 14336 B
#define CF_B_STAGE (2 * CF_B_HALF)                    // hi | lo
#define CF_SLOTS 4                                    // accumulator slots (TMEM), rotating over the (chunk, tile) items
#define CF_SLOT_COLS 112
#define CF_A_COL 448                                  // A hi operands: tile t -> TMEM columns 448 + 32 t
#define CF_W1_BYTES (2 * 64 * TC_K * 2)               // hi | lo image of the first-layer weights (N = 64): 16 KB, one ring stage
#define CF_ALO_TILE (128 * TC_K * 2)                  // A lo operand of one tile in shared memory (K-major core matrices): 16 KB

// profiling aid (tools/conv_fused_probe.py --stamps): clock64() stamps of CTA 0, second pair; role 0 = MMA issuer,
// 1 = warp 0, 2 = warp 4; [role][item (< 64)][3]
#define CF_STAMP(role, idx, k) do { if (PROBE && blockIdx.x == 0 && probe_on && (idx) < 64) a_dbg[((role) * 64 + (idx)) * 3 + (k)] = clock64(); } while (0)

// EXPERIMENTAL weight-column layout (dp_conv_fused_flat; not the default, not yet validated on hardware): the W columns of the
// e3nn weight layout are cut into consecutive 112-column chunks regardless of the path boundaries ("flat"), so that only the
// last chunk carries zero padding: ceil(W / 112) chunks instead of W / 100 chunks of 112 (2200 columns: 20 instead of 22 MMA
// groups, 9 % less issued tensor-pipe work).  The per-column work is generated from compile-time column -> (path, row, channel)
// maps, so the accumulation order (paths, then weight rows, in e3nn order) and every product are those of the path-aligned layout.
template <class Base>
struct CfFlat : Base {
    static constexpr bool FLAT = true;
};
// ... and with the last chunk's MMA N trimmed to its valid columns rounded up to 16 (1600 columns: 14 x 112 + 32 instead of 15 x 112)
template <class Base>
struct CfFlatTrim : Base {
    static constexpr bool FLAT = true;
    static constexpr bool TRIM = true;
};
template <class Cfg, class = void>
struct cf_is_trim { static constexpr bool value = false; };
template <class Cfg>
struct cf_is_trim<Cfg, std::enable_if_t<Cfg::TRIM>> { static constexpr bool value = true; };
template <class Cfg, class = void>
struct cf_is_flat { static constexpr bool value = false; };
template <class Cfg>
struct cf_is_flat<Cfg, std::enable_if_t<Cfg::FLAT>> { static constexpr bool value = true; };

struct ConvFusedArgs {
    // first MLP layer: attr(e) = [emb[perm[e]] | tb[idxB[e], 0:20] | tc[idxC[e], 0:20] (+ tc[idxC2[e], 0:20])]
    const float* emb;           // [E_canonical, 20] edge embedding
    const float* tb;            // node table of part B, row stride strideB floats
    const int* idxB;
    int strideB;
    const float* tc;            // node table of part C
    const int* idxC;
    const int* idxC2;           // optional second row of tc, added (tor_bond_conv) or nullptr
    int strideC;
    const void* w1img;          // [hi|lo][8 k-chunks][8 row groups][8 rows][8] fp16 of W1aug * w1scale (bias in k = 60)
    float inv_w1scale;
    const void* w2img;          // [W/100][hi|lo][8 k-chunks][14 row groups][8 rows][8] fp16 of W2aug * wscale
    float inv_wscale;
    const float* node_in;       // [n_in, D_IN]
    const int* gather_idx;      // [E] row of node_in per edge (nullptr: identity)
    const int* perm;            // [E] row of sh per edge (nullptr: identity)
    const float* sh;
    int sh_stride;
    const int* seg_ptr;         // [n_out + 1]
    const int* tile_node;       // [n_tiles + 1] first output node of every pair tile (<= 256 edges); tile_node[n_tiles] = n_out
    const int* n_tiles_dev;     // device tile count (dynamic graphs) or nullptr
    int n_tiles;
    const float* oscale;
    const float* oshift;
    float* out;
    const float* residual;
    int res_dim, mode;
    long long* dbg;             // profiling aid (PROBE instantiation only)
};

template <class Cfg>
struct ConvFusedSmem {
    // Gathered node rows: stride = D_IN rounded so that rows stay 16-byte (8-byte for D_IN = 50) aligned for one vector store per
    // lane, with an odd number of vectors per row: the thread-per-row scalar reads of the main loop are then at most 4-way bank
    // conflicted (a few hundred LDS per pair on an otherwise idle LSU), while the gather needs 4x fewer store instructions
    // than an odd scalar stride (measured: the gather, not the reads, was exposed).
    static constexpr int XVEC = (Cfg::D_IN % 4 == 0) ? 4 : 2;
    static constexpr int XS = (((Cfg::D_IN / XVEC) | 1)) * XVEC;
    static constexpr int OQ = ((Cfg::D_OUT + 3) / 4) | 1;                // float4 per staged row, odd: conflict-free STS.128 / LDS.128
    static constexpr int OS = 4 * OQ;
    static constexpr int ROW_FLOATS = XS > OS ? XS : OS;
    static constexpr int X_BYTES = ((256 * ROW_FLOATS * 4 + 127) / 128) * 128;
    static constexpr int TAIL = 2176;                                    // node_seg[257+], oscale/oshift, barriers, tmem slot
    static constexpr int AVAIL = 227 * 1024 - X_BYTES - TAIL - 2 * CF_ALO_TILE;
#ifndef CF_MAX_STAGES
#define CF_MAX_STAGES 6
#endif
    static constexpr int STAGES = AVAIL / CF_B_STAGE > CF_MAX_STAGES ? CF_MAX_STAGES : AVAIL / CF_B_STAGE;
    static constexpr int TOTAL = STAGES * CF_B_STAGE + 2 * CF_ALO_TILE + X_BYTES + TAIL;
    static_assert(STAGES >= 3, "weight ring too small");
};

// TMEM -> registers, 16 (or 4) consecutive columns of this thread's lane
__device__ __forceinline__ void cf_tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void cf_tmem_ld4(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
// wait for the outstanding tcgen05.ld; the registers are threaded through the asm so that no consumer can be
// scheduled above the wait
__device__ __forceinline__ void cf_wait_ld16(float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// Converged-warp issue: every lane of the issuing warp executes these, one elected lane issues.  In uniform control flow
// ptxas emits bare UTCHMMA / UTCBAR instructions; under `if (lane == 0)` it wraps each one in an ELECT / BRA.U.ANY loop
// (~58 clk per MMA, measured), which starves the tensor pipe at N = 112 (56 clk per MMA).
__device__ __forceinline__ void cf_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void cf_mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void cf_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(tc_smem(bar))
        : "memory");
}
// One warp-wide round trip for up to three barriers: lane 0 -> (b0, p0), lane 1 -> (b1, p1), lane 2 -> (b2, p2), the other
// lanes shadow lane 0.  A barrier try_wait costs ~100 clk even when its phase is complete; the issuing warp's critical
// path pays that once per weight chunk instead of three times.
__device__ __forceinline__ void cf_wait3(uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1, uint64_t* b2, uint32_t p2, int lane) {
    const uint32_t addr = tc_smem(lane == 1 ? b1 : (lane == 2 ? b2 : b0));
    const uint32_t parity = lane == 1 ? p1 : (lane == 2 ? p2 : p0);
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
        if (__all_sync(0xffffffffu, ok)) return;
#ifdef CF_WAIT_BACKOFF
        if (it >= 4) __nanosleep(32);
#endif
    }
    __trap();
}
__device__ __forceinline__ void cf_bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Accumulator layout: the per-edge output row is kept as float2 pairs of ADJACENT OUTPUT CHANNELS (v, v + 1) of one component k,
// acc2[off/2 + k * V/2 + v/2], so that one FFMA2 (sm_100 packed fp32 FMA, __ffma2_rn: two IEEE fmas) consumes two adjacent
// weight columns.  cf_acc_slot maps an e3nn output index d = off + v * K + k to 2 * pair + half.
template <class Cfg>
__host__ __device__ constexpr int cf_acc_slot(int d) {
    for (int oi = 0; oi < Cfg::NO; ++oi) {
        const TpOut t = Cfg::outs[oi];
        const int K = 2 * t.lo + 1;
        if (d >= t.off && d < t.off + t.V * K) {
            const int r = d - t.off, v = r / K, k = r % K;
            return 2 * (t.off / 2 + k * (t.V / 2) + v / 2) + (v & 1);
        }
    }
    return 0;
}

template <class Cfg, int D>
__device__ __forceinline__ float cf_acc_get(const float2 (&acc)[Cfg::D_OUT / 2]) {
    if constexpr (D < Cfg::D_OUT) {
        constexpr int sl = cf_acc_slot<Cfg>(D);                            // compile-time evaluation
        return (sl & 1) ? acc[sl >> 1].y : acc[sl >> 1].x;
    } else {
        return 0.f;
    }
}
// stage this thread's output row (e3nn order) as float4 quads
template <class Cfg, int Q, int NQ>
__device__ __forceinline__ void cf_stage_row(const float2 (&acc)[Cfg::D_OUT / 2], float* dst) {
    if constexpr (Q < NQ) {
        *reinterpret_cast<float4*>(dst + 4 * Q) = make_float4(cf_acc_get<Cfg, 4 * Q>(acc), cf_acc_get<Cfg, 4 * Q + 1>(acc),
                                                              cf_acc_get<Cfg, 4 * Q + 2>(acc), cf_acc_get<Cfg, 4 * Q + 3>(acc));
        cf_stage_row<Cfg, Q + 1, NQ>(acc, dst);
    }
}

// One 100-column chunk of path P (rows u0 .. u0 + 100/V - 1), this thread's edge.
template <class Cfg, int P>
__device__ __forceinline__ void cf_chunk(uint32_t tslot, const float* __restrict__ xrow, int u0, const float* shv,
                                         float2 (&acc)[Cfg::D_OUT / 2]) {
    constexpr TpPath p = Cfg::paths[P];
    constexpr TpOut o = Cfg::outs[p.oi];
    constexpr int V = o.V, K = 2 * o.lo + 1, D1 = 2 * p.l1 + 1;
    static_assert(CF_CHUNK % V == 0 && V % 2 == 0 && o.off % 2 == 0, "chunk must hold whole weight rows of channel pairs");
    float wv[2][16];
    float2 zz[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    cf_tmem_ld16(tslot, wv[0]);
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        cf_wait_ld16(wv[q & 1]);
        if (q + 1 < 6) cf_tmem_ld16(tslot + 16 * (q + 1), wv[(q + 1) & 1]);
        else if (q + 1 == 6) cf_tmem_ld4(tslot + 96, wv[0]);
#pragma unroll
        for (int j = 0; j < (q < 6 ? 16 : 4); j += 2) {
            const int col = 16 * q + j, r = col / V, v = col % V;
            if (v == 0) {                                                  // next weight row: Z[u, :] = CG(x[u, :], sh)
                float xv[3], z[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < D1; ++i) xv[i] = xrow[p.in_off + (u0 + r) * D1 + i];
                dp_cg<p.l1, p.l2, o.lo>(xv, shv + p.sh_off, z);
#pragma unroll
                for (int k = 0; k < K; ++k) zz[k] = make_float2(z[k], z[k]);
            }
            const float2 w2 = make_float2(wv[q & 1][j], wv[q & 1][j + 1]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float2& a2 = acc[o.off / 2 + k * (V / 2) + v / 2];
                a2 = __ffma2_rn(w2, zz[k], a2);
            }
        }
    }
}

template <class Cfg, int P, int END, bool PROBE>
__device__ __forceinline__ void cf_paths(uint32_t tmem_lane_base, uint64_t* t_full, uint64_t* t_empty, uint32_t& item, int tile,
                                         int ntile, bool active, const float* __restrict__ xrow, const float* shv,
                                         float2 (&acc)[Cfg::D_OUT / 2], int lane, bool probe_on, uint32_t item0, long long* a_dbg) {
    if constexpr (P < END) {
        constexpr TpPath p = Cfg::paths[P];
        constexpr int V = Cfg::outs[p.oi].V;
        constexpr int NCH = p.U * V / CF_CHUNK, R = CF_CHUNK / V;
        static_assert(p.U * V % CF_CHUNK == 0 && p.w_off % CF_CHUNK == 0, "path blocks must be multiples of the chunk");
#pragma unroll 1
        for (int j = 0; j < NCH; ++j) {
            if (active) {
                const uint32_t it = item, slot = 2 * (it & 1) + (uint32_t)tile, use = it >> 1;   // item: chunks drained by this thread so far
                const bool st_on = probe_on && lane == 0 && (threadIdx.x >> 5 & 3) == 0;
                if (st_on) CF_STAMP(1 + tile, it - item0, 0);
                if (lane == 0) tc_mbar_wait(&t_full[slot], use & 1);       // one poller per warp keeps the barrier unit quiet
                __syncwarp();
                if (st_on) CF_STAMP(1 + tile, it - item0, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                cf_chunk<Cfg, P>(tmem_lane_base + slot * CF_SLOT_COLS, xrow, j * R, shv, acc);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&t_empty[slot]);
                if (st_on) CF_STAMP(1 + tile, it - item0, 2);
            }
            if (active) ++item;
        }
        cf_paths<Cfg, P + 1, END, PROBE>(tmem_lane_base, t_full, t_empty, item, tile, ntile, active, xrow, shv, acc, lane, probe_on, item0, a_dbg);
    }
}

// ---- flat layout (CfFlat<Cfg>): chunk C holds weight columns [112 C, 112 C + 112) of the e3nn layout ----
template <class Cfg>
__host__ __device__ constexpr int cf_path_of(int g) {                    // path that owns weight column g
    for (int pi = 0; pi < Cfg::NP; ++pi) {
        const TpPath p = Cfg::paths[pi];
        if (g >= p.w_off && g < p.w_off + p.U * Cfg::outs[p.oi].V) return pi;
    }
    return 0;
}
template <class Cfg, int C>
struct CfFlatChunk {
    static constexpr int G0 = C * CF_N;
    static constexpr int NV = (Cfg::W - G0) < CF_N ? (Cfg::W - G0) : CF_N;            // valid columns of this chunk (even)
};
// columns COL, COL + 1 of chunk C (one FFMA2 per output component), then the rest of the chunk
template <class Cfg, int C, int COL>
__device__ __forceinline__ void cf_flat_cols(uint32_t tslot, float (&wv)[2][16], float2 (&zz)[3], const float* __restrict__ xrow,
                                             const float* shv, float2 (&acc)[Cfg::D_OUT / 2]) {
    using CH = CfFlatChunk<Cfg, C>;
    if constexpr (COL < CH::NV) {
        constexpr int q = COL / 16, j = COL % 16;
        if constexpr (j == 0) {
            cf_wait_ld16(wv[q & 1]);
            if constexpr (16 * (q + 1) < CH::NV) cf_tmem_ld16(tslot + 16 * (q + 1), wv[(q + 1) & 1]);
        }
        constexpr int g = CH::G0 + COL;
        constexpr TpPath p = Cfg::paths[cf_path_of<Cfg>(g)];
        constexpr TpOut o = Cfg::outs[p.oi];
        constexpr int V = o.V, K = 2 * o.lo + 1, D1 = 2 * p.l1 + 1, r = (g - p.w_off) / V, v = (g - p.w_off) % V;
        static_assert(V % 2 == 0 && o.off % 2 == 0 && p.w_off % 2 == 0 && CF_N % 2 == 0, "channel pairs must not straddle rows or chunks");
        if constexpr (v == 0 || COL == 0) {                                // next weight row (or a row continued from the previous chunk)
            float xv[3], z[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < D1; ++i) xv[i] = xrow[p.in_off + r * D1 + i];
            dp_cg<p.l1, p.l2, o.lo>(xv, shv + p.sh_off, z);
#pragma unroll
            for (int k = 0; k < K; ++k) zz[k] = make_float2(z[k], z[k]);
        }
        const float2 w2 = make_float2(wv[q & 1][j], wv[q & 1][j + 1]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float2& a2 = acc[o.off / 2 + k * (V / 2) + v / 2];
            a2 = __ffma2_rn(w2, zz[k], a2);
        }
        cf_flat_cols<Cfg, C, COL + 2>(tslot, wv, zz, xrow, shv, acc);
    }
}
template <class Cfg, int C, int END>
__device__ __forceinline__ void cf_flat_chunks(uint32_t tmem_lane_base, uint64_t* t_full, uint64_t* t_empty, uint32_t& item, int tile,
                                               const float* __restrict__ xrow, const float* shv, float2 (&acc)[Cfg::D_OUT / 2], int lane) {
    if constexpr (C < END) {
        const uint32_t it = item, slot = 2 * (it & 1) + (uint32_t)tile, use = it >> 1;
        if (lane == 0) tc_mbar_wait(&t_full[slot], use & 1);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const uint32_t tslot = tmem_lane_base + slot * CF_SLOT_COLS;
            float wv[2][16];
            float2 zz[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            cf_tmem_ld16(tslot, wv[0]);
            cf_flat_cols<Cfg, C, 0>(tslot, wv, zz, xrow, shv, acc);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&t_empty[slot]);
        ++item;
        cf_flat_chunks<Cfg, C + 1, END>(tmem_lane_base, t_full, t_empty, item, tile, xrow, shv, acc, lane);
    }
}

template <class Cfg, bool PROBE>
__global__ void __launch_bounds__(CF_THREADS, 1) conv_fused_kernel(ConvFusedArgs a) {
    long long* const a_dbg = a.dbg;
    using S = ConvFusedSmem<Cfg>;
    constexpr bool FLAT = cf_is_flat<Cfg>::value;
    constexpr int NCH = FLAT ? (Cfg::W + CF_N - 1) / CF_N : Cfg::W / CF_CHUNK;
    static_assert(FLAT || Cfg::W % CF_CHUNK == 0, "W must be a multiple of the chunk");
    constexpr bool TRIM = cf_is_trim<Cfg>::value;                        // flat layout only: last chunk with a narrower MMA
    constexpr int N_LAST = TRIM ? ((Cfg::W - (NCH - 1) * CF_N + 15) / 16) * 16 : CF_N;
    extern __shared__ __align__(1024) uint8_t cf_smem_raw[];
    uint8_t* b_st = cf_smem_raw;                                                    // STAGES x (hi | lo)
    uint8_t* alo = b_st + S::STAGES * CF_B_STAGE;                                   // 2 tiles x [k/8][m/8][m%8][k%8] fp16 (A lo)
    float* xs = reinterpret_cast<float*>(alo + 2 * CF_ALO_TILE);                    // [256][XS]; later [256][OS] staging
    int* node_seg = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(xs) + S::X_BYTES);          // [<= 257] seg_ptr of the pair's nodes
    float* osc = reinterpret_cast<float*>(node_seg + 260);                                        // oscale | oshift
    float* osh = osc + Cfg::D_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(osc + 2 * 104);
    uint64_t *b_full = bars, *b_empty = bars + S::STAGES, *t_full = bars + 2 * S::STAGES, *t_empty = t_full + CF_SLOTS,
             *a_ready = t_empty + CF_SLOTS;   // [0]: layer-1 operands in place, [1]: layer-2 operands; one phase per pair each
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // REDUX results live in uniform registers: everything derived from them (trip counts, the early exit, the MMA operands)
    // is provably warp-uniform.  With the tile count or the TMEM base in vector registers ptxas treats the issuing warp as
    // possibly divergent and brackets every UTCHMMA with ELECT / R2UR / VOTEU sequences that cost as much as the MMA itself.
    const int n_tiles = __reduce_max_sync(0xffffffffu, a.n_tiles_dev ? *a.n_tiles_dev : a.n_tiles);
    const int n_pairs = n_tiles;                                      // one pair tile (2 x 128 MMA rows) per item of work
    if ((int)blockIdx.x >= n_pairs) return;
    const int my_pairs = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < S::STAGES; ++i) { tc_mbar_init(&b_full[i], 1); tc_mbar_init(&b_empty[i], 2); }
        for (int i = 0; i < CF_SLOTS; ++i) { tc_mbar_init(&t_full[i], 1); tc_mbar_init(&t_empty[i], 4); }
        tc_mbar_init(&a_ready[0], CF_WORKERS / 32);
        tc_mbar_init(&a_ready[1], CF_WORKERS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Cfg::D_OUT; i += CF_THREADS) { osc[i] = a.oscale[i]; osh[i] = a.oshift[i]; }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CF_REG_AUX));
      if (warp == 11) {
        // (idle: completes the auxiliary warpgroup)
      } else if (warp == 10) {
        // ================= TMA producer (one thread): the weight chunks cycle through the ring, pair after pair =================
        if (lane == 0) {
            const uint32_t total_chunks = (uint32_t)my_pairs * (NCH + 1);
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w2img);
            for (uint32_t g = 0; g < total_chunks; ++g) {
                const uint32_t s = g % S::STAGES, u = g / S::STAGES, c = g % (NCH + 1);   // c = 0: first-layer weights
                tc_mbar_wait(&b_empty[s], (u & 1) ^ 1);
                if (c == 0) {
                    tc_mbar_expect_tx(&b_full[s], CF_W1_BYTES);
                    tc_bulk_load(b_st + s * CF_B_STAGE, a.w1img, CF_W1_BYTES, &b_full[s]);
                } else {
                    tc_mbar_expect_tx(&b_full[s], CF_B_STAGE);
                    tc_bulk_load(b_st + s * CF_B_STAGE, wsrc + (size_t)(c - 1) * CF_B_STAGE, CF_B_STAGE, &b_full[s]);
                }
            }
        }
      } else {
        // ================= MMA issuers: warp 8 -> tile 0 (accumulator slots 0, 2), warp 9 -> tile 1 (slots 1, 3).
        // One issuing warp needs ~55 clk of instructions per UTCHMMA, about the 56 clk an M=128, N=112, K=16 MMA occupies
        // the tensor pipe, so every barrier round trip of a single issuer would starve the pipe; two issuers hide each
        // other's waits.  Each (slot, barrier) pair has exactly one producer and one consumer group, so every parity wait
        // stays within one phase of its barrier. =================
        {
            const int t = __reduce_max_sync(0xffffffffu, warp - 8);         // warp-uniform by construction: keep it in a uniform register
            // instruction descriptor: D = F32, A = B = F16, K-major, N = 112, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(CF_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t b_base = tc_smem(b_st);
            const uint32_t a_hi_t = tmem_base + CF_A_COL + (uint32_t)(t * 32);
            const uint32_t a_lo_s = tc_smem(alo) + (uint32_t)(t * CF_ALO_TILE);
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // first layer: N = 64
            const uint32_t idesc_last = (1u << 4) | ((uint32_t)(N_LAST >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t g = 0, use = 0;                                        // use: items issued by this warp so far
            for (int pi = 0; pi < my_pairs; ++pi) {
                constexpr bool mine = true;                             // (both MMA tiles of a pair tile always run; rows beyond its edges are zero)
                const bool probe_on = pi == 1 && lane == 0 && t == 0;
                int pidx = 0;
                // barriers of one item: its weights landed, this tile's accumulator slot is drained
                auto wait_chunk = [&](uint32_t gg, uint32_t uu) {
                    uint64_t* te = &t_empty[2 * (uu & 1) + t];
                    const uint32_t tp = ((uu >> 1) & 1) ^ 1;
                    cf_wait3(&b_full[gg % S::STAGES], (gg / S::STAGES) & 1, mine ? te : &b_full[gg % S::STAGES],
                             mine ? tp : (gg / S::STAGES) & 1, &b_full[gg % S::STAGES], (gg / S::STAGES) & 1, lane);
                };
                // ---- item 0 of the pair: hidden layer  D1[128, 64] = attr . W1aug  (vote-terminated waits keep the warp converged)
                cf_wait3(&a_ready[0], (uint32_t)(pi & 1), &a_ready[0], (uint32_t)(pi & 1), &a_ready[0], (uint32_t)(pi & 1), lane);   // layer-1 A operands
                wait_chunk(g, use);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    const uint32_t s = g % S::STAGES;
                    if (mine) {
                        const uint32_t b_hi_s = b_base + s * CF_B_STAGE, b_lo_s = b_hi_s + CF_W1_BYTES / 2;
                        const uint32_t slot = 2 * (use & 1) + (uint32_t)t;
                        const uint32_t d = tmem_base + slot * CF_SLOT_COLS;
#pragma unroll
                        for (int combo = 0; combo < 3; ++combo) {
                            const uint32_t bs = combo == 1 ? b_lo_s : b_hi_s;
#pragma unroll
                            for (int ks = 0; ks < TC_K / 16; ++ks) {
                                const uint64_t bd = tc_smem_desc(bs + ks * 2 * 1024, 1024, 128);      // 8 row groups per K chunk
                                if (combo < 2) cf_mma_f16_ts(d, a_hi_t + (uint32_t)(ks * 8), bd, idesc1, (combo | ks) ? 1u : 0u);
                                else cf_mma_f16_ss(d, tc_smem_desc(a_lo_s + ks * 2 * 2048, 2048, 128), bd, idesc1, 1u);
                            }
                        }
                        cf_commit(&t_full[slot]);
                        cf_commit(&b_empty[s]);
                        ++use;
                    } else {
                        if (lane == 0) tc_mbar_arrive(&b_empty[s]);
                        __syncwarp();
                    }
                    ++g;
                }
                // ---- items 1 .. NCH: weight chunks
                cf_wait3(&a_ready[1], (uint32_t)(pi & 1), &a_ready[1], (uint32_t)(pi & 1), &a_ready[1], (uint32_t)(pi & 1), lane);   // layer-2 A operands
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < NCH; ++c, ++g) {
                    const uint32_t s = g % S::STAGES;
                    // Wait at the top of the item (weights landed, slot drained).  A wait inside the item (before its last
                    // MMAs) measured slower: with a 3-stage weight ring the NEXT chunk's weights are still in flight then,
                    // and the item's own tail MMAs would be held back behind that round trip.
                    CF_STAMP(0, pidx, 0);
                    wait_chunk(g, use);
                    if (!mine) {                                            // single-tile pair: only release the weight stage
                        if (lane == 0) tc_mbar_arrive(&b_empty[s]);
                        __syncwarp();
                        continue;
                    }
                    CF_STAMP(0, pidx, 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_hi_s = b_base + s * CF_B_STAGE, b_lo_s = b_hi_s + CF_B_HALF;
                    const uint32_t slot = 2 * (use & 1) + (uint32_t)t;
                    const uint32_t d = tmem_base + slot * CF_SLOT_COLS;
                    const uint32_t idc = (TRIM && c == NCH - 1) ? idesc_last : idesc;
#pragma unroll
                    for (int combo = 0; combo < 3; ++combo) {              // hi*hi + hi*lo (A from TMEM) + lo*hi (A from smem)
                        const uint32_t bs = combo == 1 ? b_lo_s : b_hi_s;
#pragma unroll
                        for (int ks = 0; ks < TC_K / 16; ++ks) {
                            // fp16 K-major no-swizzle: core matrix = 8 rows x 8 halfs (128 B); B: 14 row groups per K chunk,
                            // A lo: 16 row groups per K chunk
                            const uint64_t bd = tc_smem_desc(bs + ks * 2 * (CF_N * 16), CF_N * 16, 128);
                            if (combo < 2) cf_mma_f16_ts(d, a_hi_t + (uint32_t)(ks * 8), bd, idc, (combo | ks) ? 1u : 0u);
                            else cf_mma_f16_ss(d, tc_smem_desc(a_lo_s + ks * 2 * 2048, 2048, 128), bd, idc, 1u);
                        }
                    }
                    cf_commit(&t_full[slot]);
                    cf_commit(&b_empty[s]);
                    CF_STAMP(0, pidx, 2);
                    ++pidx;
                    ++use;
                }
            }
        }
      }
    } else {
        // ================= workers: thread = edge =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CF_REG_WORK));
        const int tile = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        float* xrow = xs + (size_t)tid * S::XS;
        uint32_t item = 0;
        // indices of this thread's edge in pair `pr`: node range of the pair, first edge / edge count of its tile, gathered
        // node row, SH row
        struct Idx { int n_lo, n_hi, eb, ne, src, ce, ib, ic, ic2; };
        auto fetch = [&](int pr) {
            Idx x;
            x.n_lo = a.tile_node[pr]; x.n_hi = a.tile_node[pr + 1];
            x.src = -1; x.ce = 0; x.ib = 0; x.ic = 0; x.ic2 = -1;
            {
                const int e0 = a.seg_ptr[x.n_lo], ne_pair = a.seg_ptr[x.n_hi] - e0;
                x.eb = e0 + 128 * tile;                                  // MMA tile 0: edges 0..127 of the pair tile, tile 1: the rest
                x.ne = min(max(ne_pair - 128 * tile, 0), 128);
                if (row < x.ne) {
                    x.src = a.gather_idx ? a.gather_idx[x.eb + row] : x.eb + row;
                    x.ce = a.perm ? a.perm[x.eb + row] : x.eb + row;
                    x.ib = a.idxB[x.eb + row];
                    x.ic = a.idxC[x.eb + row];
                    if (a.idxC2) x.ic2 = a.idxC2[x.eb + row];
                }
            }
            return x;
        };
        // edge attributes [emb | node B | node C (+ C2) | 1 | 0 0 0] of this thread's edge (zeros for rows beyond the tile)
        auto load_attr = [&](const Idx& x, float (&at)[64]) {
            if (x.src >= 0) {
                const float4* pe = reinterpret_cast<const float4*>(a.emb + (size_t)x.ce * 20);
                const float2* pb = reinterpret_cast<const float2*>(a.tb + (size_t)x.ib * a.strideB);
                const float2* pc = reinterpret_cast<const float2*>(a.tc + (size_t)x.ic * a.strideC);
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const float4 v = __ldg(pe + q);
                    at[4 * q] = v.x; at[4 * q + 1] = v.y; at[4 * q + 2] = v.z; at[4 * q + 3] = v.w;
                }
                if (((a.strideB | a.strideC) & 3) == 0) {                     // 16-byte aligned node rows: half the load instructions
                    const float4* pb4 = reinterpret_cast<const float4*>(pb);
                    const float4* pc4 = reinterpret_cast<const float4*>(pc);
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const float4 vb = __ldg(pb4 + q), vc = __ldg(pc4 + q);
                        at[20 + 4 * q] = vb.x; at[21 + 4 * q] = vb.y; at[22 + 4 * q] = vb.z; at[23 + 4 * q] = vb.w;
                        at[40 + 4 * q] = vc.x; at[41 + 4 * q] = vc.y; at[42 + 4 * q] = vc.z; at[43 + 4 * q] = vc.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 10; ++q) {
                        const float2 vb = __ldg(pb + q), vc = __ldg(pc + q);
                        at[20 + 2 * q] = vb.x; at[21 + 2 * q] = vb.y;
                        at[40 + 2 * q] = vc.x; at[41 + 2 * q] = vc.y;
                    }
                }
                if (x.ic2 >= 0) {
                    const float2* pc2 = reinterpret_cast<const float2*>(a.tc + (size_t)x.ic2 * a.strideC);
#pragma unroll
                    for (int q = 0; q < 10; ++q) {
                        const float2 vc = __ldg(pc2 + q);
                        at[40 + 2 * q] += vc.x; at[41 + 2 * q] += vc.y;
                    }
                }
                at[60] = 1.0f; at[61] = 0.f; at[62] = 0.f; at[63] = 0.f;       // k = 60 multiplies the bias row of W1aug
            } else {
#pragma unroll
                for (int q = 0; q < 64; ++q) at[q] = 0.f;
            }
        };
        Idx ix = fetch((int)blockIdx.x);
        float at[64];                                                        // loaded one pair ahead (before the previous reduce)
        load_attr(ix, at);
        for (int pi = 0; pi < my_pairs; ++pi) {
            const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
            constexpr int ntile = 2;
            constexpr bool active = true;
            const bool valid = row < ix.ne;
            const bool probe_on = pi == 1 && tid == 0;
            CF_STAMP(1, 50, 0);
            float shv[Cfg::SH_USED];
            // exactly-scaled FP16 split of one operand row (64 values): row * 2^s with max in [2^12, 2^13), hi -> TMEM, lo -> smem
            // (K-major core matrices [k/8][m/8][m%8][k%8]; 8 consecutive rows write 128 contiguous bytes).  Returns 2^-s.
            auto put_operand = [&](const float (&v)[64]) -> float {
                float m = 1.0f;                                             // >= 1: column 60 holds the constant 1.0
#pragma unroll
                for (int q = 0; q < 64; ++q) m = fmaxf(m, fabsf(v[q]));
                const int ex = (int)((__float_as_uint(m) >> 23) & 0xFF) - 127;
                const float sc = __uint_as_float((uint32_t)(127 + 12 - ex) << 23);
                uint32_t hi_p[32], lo_p[32];
                const float2 sc2 = make_float2(sc, sc), neg1 = make_float2(-1.f, -1.f);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float2 x = __fmul2_rn(make_float2(v[2 * c], v[2 * c + 1]), sc2);    // packed fp32 ops (sm_100)
                    const __half2 hh = __floats2half2_rn(x.x, x.y);         // packed conversions (F2FP), 2 values per instruction
                    const float2 lo = __ffma2_rn(__half22float2(hh), neg1, x);                  // x - hi, exact
                    const __half2 ll = __floats2half2_rn(lo.x, lo.y);
                    hi_p[c] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo_p[c] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                tc_tmem_st32(lane_base + CF_A_COL + (uint32_t)(tile * 32), hi_p);
                uint8_t* lo_dst = alo + tile * CF_ALO_TILE + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
                for (int kc = 0; kc < 8; ++kc)
                    *reinterpret_cast<uint4*>(lo_dst + kc * 2048) = make_uint4(lo_p[4 * kc], lo_p[4 * kc + 1], lo_p[4 * kc + 2], lo_p[4 * kc + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                return __uint_as_float((uint32_t)(127 - 12 + ex) << 23);
            };
            // ---- prologue 1: edge attributes [emb | node B | node C] -> layer-1 A operand ----
            float rs1 = 0.f;
            if (active) rs1 = put_operand(at) * a.inv_w1scale;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&a_ready[0]);                    // layer-1 operands in place
            CF_STAMP(1, 50, 1);
            // ---- prologue 2 (under the hidden-layer MMA): seg_ptr of the pair's nodes -> smem for the epilogue, SH -> registers ----
            for (int i = tid; i <= ix.n_hi - ix.n_lo; i += CF_WORKERS) node_seg[i] = a.seg_ptr[ix.n_lo + i];
#pragma unroll
            for (int i = 0; i < Cfg::SH_USED; ++i) shv[i] = (active && valid) ? __ldg(a.sh + (size_t)ix.ce * a.sh_stride + i) : 0.f;
            // ---- prologue 3: hidden activations h = ReLU(D1) from tensor memory -> layer-2 A operand ----
            float rs = 0.f;
            if (active) {
                const uint32_t slot = 2 * (item & 1) + (uint32_t)tile, use = item >> 1;
                if (lane == 0) tc_mbar_wait(&t_full[slot], use & 1);
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float h[64];
#pragma unroll
                for (int q = 0; q < 4; ++q) cf_tmem_ld16(lane_base + slot * CF_SLOT_COLS + 16 * q, h + 16 * q);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 4; ++q) cf_wait_ld16(h + 16 * q);      // (register dependency only)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&t_empty[slot]);
                ++item;
#pragma unroll
                for (int q = 0; q < 60; ++q) h[q] = fmaxf(h[q] * rs1, 0.f);
                h[60] = valid ? 1.0f : 0.f; h[61] = 0.f; h[62] = 0.f; h[63] = 0.f;   // k = 60 multiplies the bias row of W2aug
                rs = put_operand(h) * a.inv_wscale;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&a_ready[1]);                    // layer-2 operands in place
            // the power-of-two operand scales are undone exactly by scaling the spherical harmonics (Z is linear in them)
#pragma unroll
            for (int i = 0; i < Cfg::SH_USED; ++i) shv[i] *= rs;
            // ---- prologue 4 (under the first chunk's MMAs): gathered node rows -> smem, warp-cooperative (lanes walk a row:
            //      coalesced in global memory, conflict-free in the odd-stride smem rows), 32 rows = 4 x 32 loads in flight ----
            if (active) {
                // lanes read VEC consecutive floats of a row: one 16- or 8-byte load and one vector store per row and lane
                constexpr int VEC = S::XVEC;
                constexpr int NL = Cfg::D_IN / VEC;                         // active lanes per row (<= 32)
                static_assert(Cfg::D_IN % VEC == 0 && NL <= 32, "row gather layout");
                float v[32][VEC];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int sr = __shfl_sync(0xffffffffu, ix.src, j);
                    const float* srow = a.node_in + (size_t)(sr < 0 ? 0 : sr) * Cfg::D_IN + VEC * lane;
                    if (sr >= 0 && lane < NL) {
                        if constexpr (VEC == 4) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(srow));
                            v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
                        } else {
                            const float2 t = __ldg(reinterpret_cast<const float2*>(srow));
                            v[j][0] = t.x; v[j][1] = t.y;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) v[j][i] = 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float* dst = xs + (size_t)(tid - lane + j) * S::XS + VEC * lane;
                    if (lane < NL) {
                        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
                        else *reinterpret_cast<float2*>(dst) = make_float2(v[j][0], v[j][1]);
                    }
                }
            }
            __syncwarp();
            // ---- main loop: weights from TMEM, contraction on CUDA cores ----
            float2 acc[Cfg::D_OUT / 2];
#pragma unroll
            for (int d = 0; d < Cfg::D_OUT / 2; ++d) acc[d] = make_float2(0.f, 0.f);
            // The next pair's indices are a chain of three dependent global loads (tile -> segment -> edge level).  Inside the
            // chunk loop its stalls are free: the workers drain a chunk in about half the time the tensor pipe needs for it.
            const uint32_t item0 = item;
            Idx nx = ix;
            if constexpr (FLAT) {
                constexpr int CS = NCH < 4 ? NCH : 4;                    // (final_conv has two chunks)
                cf_flat_chunks<Cfg, 0, CS>(lane_base, t_full, t_empty, item, tile, xrow, shv, acc, lane);
                if (pi + 1 < my_pairs) nx = fetch(pair + (int)gridDim.x);
                cf_flat_chunks<Cfg, CS, NCH>(lane_base, t_full, t_empty, item, tile, xrow, shv, acc, lane);
            } else {
                cf_paths<Cfg, 0, 1, PROBE>(lane_base, t_full, t_empty, item, tile, ntile, active, xrow, shv, acc, lane, pi == 1, item0, a_dbg);
                if (pi + 1 < my_pairs) nx = fetch(pair + (int)gridDim.x);
                cf_paths<Cfg, 1, Cfg::NP, PROBE>(lane_base, t_full, t_empty, item, tile, ntile, active, xrow, shv, acc, lane, pi == 1, item0, a_dbg);
            }
            CF_STAMP(1, 50, 2);
            // ---- epilogue: per-edge results -> smem (aliases the node rows), segmented mean + BatchNorm + residual ----
            cf_bar_workers();                                              // every worker is done with its node row
            CF_STAMP(1, 52, 0);
            float* stg = xs;
            if (active) {
                cf_stage_row<Cfg, 0, S::OQ>(acc, stg + (size_t)tid * S::OS);
            }
            CF_STAMP(1, 52, 1);
            cf_bar_workers();
            CF_STAMP(1, 52, 2);
            load_attr(nx, at);             // next pair's attributes (unconditional: `at` must not stay live across the main loop)
            CF_STAMP(1, 51, 0);
            {
                const int e_lo = node_seg[0];
                if (node_seg[ix.n_hi - ix.n_lo] - e_lo > 256) __trap();      // tile builder contract violated
                const int items = (ix.n_hi - ix.n_lo) * S::OQ;
                // residual / running-sum operand of item `it` (loads issued one iteration ahead of their use)
                auto load_add = [&](int it, float (&add)[4]) {
                    const int nl = it / S::OQ, q = it % S::OQ, node = ix.n_lo + nl;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = 4 * q + j;
                        add[j] = 0.f;
                        if (it < items && d < Cfg::D_OUT) {
                            if (a.mode == 1) add[j] = (d < a.res_dim) ? __ldg(a.residual + (size_t)node * a.res_dim + d) : 0.0f;
                            else if (a.mode == 2 && Cfg::D_OUT % 4 != 0) add[j] = a.out[(size_t)node * Cfg::D_OUT + d];
                        }
                    }
                    if constexpr (Cfg::D_OUT % 4 == 0) {                     // 16-byte aligned output rows: one load per quad
                        if (a.mode == 2 && it < items && 4 * q < Cfg::D_OUT) {        // (the last staged quad may be padding)
                            const float4 t = *reinterpret_cast<const float4*>(a.out + (size_t)node * Cfg::D_OUT + 4 * q);
                            add[0] = t.x; add[1] = t.y; add[2] = t.z; add[3] = t.w;
                        }
                    }
                };
                // Batches of 4 items per thread: their residual / running-sum loads are all in flight before the first sum is
                // consumed (one exposed HBM round trip per batch; a one-ahead prefetch still exposed one per item).
#pragma unroll 1
                for (int it0 = tid; it0 < items; it0 += 4 * CF_WORKERS) {
                    float add[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) load_add(it0 + u * CF_WORKERS, add[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int it = it0 + u * CF_WORKERS;
                        if (it >= items) break;
                        const int nl = it / S::OQ, q = it % S::OQ, node = ix.n_lo + nl;
                        const int s0 = node_seg[nl], s1 = node_seg[nl + 1];
                        float* orow = a.out + (size_t)node * Cfg::D_OUT;
                        const int r0 = s0 - e_lo;                            // staged rows = the pair tile's edges in order
                        const float* sp = stg + (size_t)r0 * S::OS + 4 * q;
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                        for (int r = 0; r < s1 - s0; ++r) {                  // sequential in edge order: composition-invariant
                            const float4 v = *reinterpret_cast<const float4*>(sp + (size_t)r * S::OS);
                            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                        }
                        const int deg = s1 - s0;
                        const float inv_deg = 1.0f / (float)(deg > 0 ? deg : 1);
                        const float sv[4] = {s.x, s.y, s.z, s.w};
                        float ov[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int d = 4 * q + j;
                            ov[j] = d < Cfg::D_OUT ? sv[j] * inv_deg * osc[d] + osh[d] + add[u][j] : 0.f;
                        }
                        if constexpr (Cfg::D_OUT % 4 == 0) {
                            if (4 * q < Cfg::D_OUT) *reinterpret_cast<float4*>(orow + 4 * q) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (4 * q + j < Cfg::D_OUT) orow[4 * q + j] = ov[j];
                        }
                    }
                }
            }
            CF_STAMP(1, 51, 1);
            cf_bar_workers();                                              // staging is free before the next pair's rows land
            CF_STAMP(1, 51, 2);
            ix = nx;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

template <class Cfg>
static int conv_fused_launch(const ConvFusedArgs& a, cudaStream_t st) {
    if (a.n_tiles <= 0) return DP_OK;
    using S = ConvFusedSmem<Cfg>;
    static bool attr_set = false;
    static int n_sm = 0;
    auto kern = conv_fused_kernel<Cfg, false>;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
            return dp_check_launch("dp_conv_fused(attr)");
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const int n_pairs = a.n_tiles;
    const int grid = n_pairs < n_sm ? n_pairs : n_sm;
    if constexpr (Cfg::W == 2200 && !cf_is_flat<Cfg>::value) {
        if (a.dbg) {                                                       // profiling aid (tools/conv_fused_probe.py --stamps)
            auto pk = conv_fused_kernel<Cfg, true>;
            cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
            pk<<<grid, CF_THREADS, S::TOTAL, st>>>(a);
            return dp_check_launch("dp_conv_fused(probe)");
        }
    }
    kern<<<grid, CF_THREADS, S::TOTAL, st>>>(a);
    return dp_check_launch("dp_conv_fused");
}

// ---------------------------------------------------------------------------------------------------------------
// Tile builder for dynamic graphs: greedy node-aligned packing (<= 256 edges and <= 256 nodes per pair tile), restarted at every GROUP of
// graphs (node_ptr holds the first node of every group; the engine groups 8 consecutive graphs) so that groups can be
// processed in parallel, two passes around the exclusive scan.  Tiles may span graphs: a row of
// the M = 128 MMA operand is one edge and rows are independent, so the per-node sums do not depend on the tiling
// (bit-identical, tested); restarting at every graph instead left the last tile of every graph mostly empty
// (cfg2: 5.2 -> 4.8 ligand-ligand M=128 operands, 3.8 -> 3.4 torsion operands per graph).  The node cap bounds node_seg[] in
// shared memory (zero-degree nodes ride along with their neighbours).
// ---------------------------------------------------------------------------------------------------------------
// One CTA per group.  The greedy rule is sequential over NODES (~100 clk per node on one thread or one warp: 50-70 us per build
// for the 1024-node groups of a 40-graph job, 10 % of its step), but the tile that STARTS at node i is a function of i alone: it ends
// before nxt(i) = the first j > i with seg[j + 1] - seg[i] > 256 or j - i == 256.  All nxt(i) are found in parallel (binary search
// in the monotone seg_ptr, L1-resident), then one thread hops from tile start to tile start through shared memory: sequential
// over TILES only.  Groups beyond TILE_WALK_CAP nodes fall back to the node-by-node walk.
#define TILE_WALK_THREADS 128
#define TILE_WALK_CAP 4096
template <class F>
__device__ __forceinline__ void tile_walk(const int* __restrict__ seg_ptr, int n0, int n1, int* nxt /*shared[TILE_WALK_CAP]*/, F&& on_tile) {
    const int cnt = n1 - n0;
    if (cnt <= TILE_WALK_CAP) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int s_i = seg_ptr[n0 + i];
            int lo = i + 1, hi = min(i + 256, cnt);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (seg_ptr[n0 + mid + 1] - s_i > 256) hi = mid; else lo = mid + 1;
            }
            nxt[i] = lo;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int t = 0; t < cnt; t = nxt[t]) on_tile(n0 + t);
    } else if (threadIdx.x == 0) {
        int fill = 0, nodes = 0;
        bool open = false;
        for (int i = n0; i < n1; ++i) {
            const int d = seg_ptr[i + 1] - seg_ptr[i];
            if (!open || fill + d > 256 || nodes == 256) { on_tile(i); fill = 0; nodes = 0; open = true; }
            fill += d;
            ++nodes;
        }
    }
}
__global__ void __launch_bounds__(TILE_WALK_THREADS)
tile_count_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ node_ptr, int n_groups, int* __restrict__ cnt) {
    __shared__ int nxt[TILE_WALK_CAP];
    const int g = blockIdx.x;
    int tiles = 0;
    tile_walk(seg_ptr, node_ptr[g], node_ptr[g + 1], nxt, [&](int) { ++tiles; });
    if (threadIdx.x == 0) cnt[g] = tiles;
}
__global__ void __launch_bounds__(TILE_WALK_THREADS)
tile_fill_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ node_ptr, int n_groups,
                 const int* __restrict__ start, int* __restrict__ tile_node) {
    __shared__ int nxt[TILE_WALK_CAP];
    const int g = blockIdx.x;
    const int n1 = node_ptr[g + 1];
    int t = start[g];
    tile_walk(seg_ptr, node_ptr[g], n1, nxt, [&](int n) { tile_node[t++] = n; });
    if (threadIdx.x == 0 && g == n_groups - 1) tile_node[start[n_groups]] = n1;     // sentinel: one past the last node
}
