// dp_conv_fused2: second generation of the fused TensorProductConvLayer kernel (conv_fused.cuh has the maths and the first
// generation).  Same products, same accumulation order, same results bit for bit; what changes is WHO does the per-pair
// prologue and WHEN, so that the tensor pipe no longer idles ~12 k clk per pair tile (profiles/README.md, round 1):
//
//   * 16 warps.  Warps 0-7: workers (thread = edge; drain the weight chunks from tensor memory, Clebsch-Gordan + channel
//     mixing on FFMA2, segmented mean).  Warps 8-11: PREP warpgroup - builds the layer-1 operand (edge attributes -> exactly
//     scaled fp16 hi/lo) and, after the hidden-layer MMA, the layer-2 operand (ReLU -> hi/lo) of the NEXT pair tile while
//     the workers and the tensor pipe are still busy with the current one.  Warps 12/13: MMA issuers of tile 0/1, warp 14:
//     TMA producer of the weight ring, warp 15: copies the window of node rows the next pair gathers from into shared memory.
//     setmaxnreg moves registers from warps 8-15 (88) to the workers (168).
//   * Both A operands are double buffered: hi in tensor memory (2 buffers x 2 tiles x 32 columns), lo in shared memory
//     (2 x 2 x 16 KB).  The hidden-layer MMA of pair p+1 is issued in the middle of pair p's chunk stream (after chunk HPOS),
//     its result is drained by the prep warps, and the first weight chunks of pair p+1 are issued while the workers are
//     still in pair p's epilogue.
//   * TMEM budget: 4 accumulator slots x 96 columns + 128 columns of A hi = 512, hence MMA N = 96: the W weight columns are
//     cut into consecutive 96-column chunks regardless of path boundaries ("flat"), the last chunk's MMA is trimmed to its
//     valid columns rounded up to 16 (600 -> 608, 1100 -> 1104, 1600, 2200 -> 2208 issued columns).
//   * Shared-memory budget: the per-edge gathered node rows (256 x D_in floats, 100 KB at layer 3) are replaced by the WINDOW
//     of node rows [min src, max src] of the pair tile (graphs are contiguous in memory: 8-64 rows for the cfg2 shapes; a
//     window that does not fit falls back to reading the rows from global memory through L1), and the per-edge output rows
//     are staged and reduced in PARTS of 5-7 float4 quads.  That pays for the second A-lo buffer and a 4th ring stage.
#pragma once
#include "conv_fused.cuh"

#define CF2_THREADS 512
#define CF2_N 96                                      // MMA N = weight columns per chunk
#define CF2_B_HALF (CF2_N * TC_K * 2)                 // one fp16 operand image (hi or lo) of a chunk: 12288 B
#define CF2_B_STAGE (2 * CF2_B_HALF)                  // hi | lo
#define CF2_SLOT_COLS 96
#define CF2_A_COL 384                                 // A hi operands: buffer b, tile t -> TMEM columns 384 + 64 b + 32 t
#define CF2_REG_WORK 176
#define CF2_REG_AUX 80

template <class Cfg>
struct ConvFused2Smem {
    static constexpr int XVEC = (Cfg::D_IN % 4 == 0) ? 4 : 2;
    static constexpr int XS = (((Cfg::D_IN / XVEC) | 1)) * XVEC;          // row stride of the node-row window (floats)
    static constexpr int XCAP_RAW = 32768 / (XS * 4);
    static constexpr int XCAP = XCAP_RAW > 256 ? 256 : XCAP_RAW;          // rows the window can hold
    static constexpr int X_BYTES = ((XCAP * XS * 4 + 127) / 128) * 128;
    static constexpr int OQ = (Cfg::D_OUT + 3) / 4;                        // float4 quads of an output row
    static constexpr int PQ = ((OQ + 6) / 7 < (OQ + 4) / 5) ? 7 : 5;       // quads per staged part (odd: conflict-free STS.128 / LDS.128), fewest parts
                                                                           // (two parts of 13 quads with a 3-stage ring measured the same)
    static constexpr int NPART = (OQ + PQ - 1) / PQ;
    static constexpr int PS = 4 * PQ;
    static constexpr int STG_BYTES = 256 * PS * 4;
    static constexpr int TAIL = 6656;                                      // row scales, node_seg (per tile), straddler sums, oscale/oshift, barriers
    static constexpr int AVAIL = 227 * 1024 - 4 * CF_ALO_TILE - X_BYTES - STG_BYTES - TAIL;
#ifndef CF2_MAX_STAGES
#define CF2_MAX_STAGES 6
#endif
    static constexpr int STAGES = AVAIL / CF2_B_STAGE > CF2_MAX_STAGES ? CF2_MAX_STAGES : AVAIL / CF2_B_STAGE;
    static constexpr int TOTAL = STAGES * CF2_B_STAGE + 4 * CF_ALO_TILE + X_BYTES + STG_BYTES + TAIL;
    static_assert(STAGES >= 3, "weight ring too small");
};

__device__ __forceinline__ void cf2_flag_set(int* p, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(tc_smem(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void cf2_flag_wait(int* p, int v) {            // bounded like the mbarrier waits: a protocol bug traps
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        int cur;
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(cur) : "r"(tc_smem(p)) : "memory");
        if (cur >= v) return;
    }
    __trap();
}
// Waits of the auxiliary warps (prep, TMA producer, window copier, and the issuers once a wait has lasted a few polls): with 8
// auxiliary warps next to the 8 workers a hot try_wait loop takes a measurable share of the issue slots of its SM sub-partition
// (two spinning warps beside two worker warps), so they back off with nanosleep; the waits are tens of thousands of clocks long
// (prep, copier) or have a whole ring stage / accumulator slot of slack (producer, issuers).
__device__ __forceinline__ void cf2_mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
    const uint32_t addr = tc_smem(bar);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(ns);
    }
    __trap();
}
__device__ __forceinline__ void cf2_wait3(uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1, uint64_t* b2, uint32_t p2, int lane) {
    const uint32_t addr = tc_smem(lane == 1 ? b1 : (lane == 2 ? b2 : b0));
    const uint32_t parity = lane == 1 ? p1 : (lane == 2 ? p2 : p0);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 27); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (__all_sync(0xffffffffu, ok)) return;
        if (it >= 4) __nanosleep(32);
    }
    __trap();
}
// named barrier of one MMA tile's 128 worker threads (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void cf_bar_tile(int tile) { asm volatile("bar.sync %0, 128;" ::"r"(tile + 1) : "memory"); }

template <class Cfg, int C>
struct Cf2Chunk {
    static constexpr int G0 = C * CF2_N;
    static constexpr int NV = (Cfg::W - G0) < CF2_N ? (Cfg::W - G0) : CF2_N;          // valid columns of this chunk (even)
};
// TMEM -> registers, 8 consecutive columns of this thread's lane (two of these in flight: 16 registers instead of the 32 of the
// x16 double buffer - the worker loop is at the register limit, and a spill there is an L2 round trip: L1 is 20 KB and thrashed)
__device__ __forceinline__ void cf2_tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void cf2_wait_ld8(float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// columns COL, COL + 1 of chunk C (one FFMA2 per output component), then the rest of the chunk
template <class Cfg, int C, int COL>
__device__ __forceinline__ void cf2_cols(uint32_t tslot, float (&wv)[3][8], float2 (&zz)[3], const float* __restrict__ xrow,
                                         const float* shv, float2 (&acc)[Cfg::D_OUT / 2]) {
    using CH = Cf2Chunk<Cfg, C>;
    if constexpr (COL < CH::NV) {
        constexpr int q = COL / 8, j = COL % 8;
        if constexpr (j == 0) {                                            // groups q + 1, q + 2 stay in flight
            cf2_wait_ld8(wv[q % 3]);
            if constexpr (8 * (q + 2) < CH::NV) cf2_tmem_ld8(tslot + 8 * (q + 2), wv[(q + 2) % 3]);
        }
        constexpr int g = CH::G0 + COL;
        constexpr TpPath p = Cfg::paths[cf_path_of<Cfg>(g)];
        constexpr TpOut o = Cfg::outs[p.oi];
        constexpr int V = o.V, K = 2 * o.lo + 1, D1 = 2 * p.l1 + 1, r = (g - p.w_off) / V, v = (g - p.w_off) % V;
        static_assert(V % 2 == 0 && o.off % 2 == 0 && p.w_off % 2 == 0 && CF2_N % 2 == 0, "channel pairs must not straddle rows or chunks");
        if constexpr (v == 0 || COL == 0) {                                // next weight row (or a row continued from the previous chunk)
            float xv[3], z[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < D1; ++i) xv[i] = xrow[p.in_off + r * D1 + i];
            dp_cg<p.l1, p.l2, o.lo>(xv, shv + p.sh_off, z);
#pragma unroll
            for (int k = 0; k < K; ++k) zz[k] = make_float2(z[k], z[k]);
        }
        const float2 w2 = make_float2(wv[q % 3][j], wv[q % 3][j + 1]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float2& a2 = acc[o.off / 2 + k * (V / 2) + v / 2];
            a2 = __ffma2_rn(w2, zz[k], a2);
        }
        cf2_cols<Cfg, C, COL + 2>(tslot, wv, zz, xrow, shv, acc);
    }
}
// The hidden-layer items share the accumulator slots with the chunks but complete on their own barriers (h_full), so that every
// waiter sees EVERY phase of the barrier it polls (a parity wait that skips a phase returns early): t_full[slot] counts the chunk
// items of the slot only - use0 / use1 = chunks drained so far from the slots of even / odd items.
template <class Cfg, int C, int END, int HPOS, bool PROBE>
__device__ __forceinline__ void cf2_chunks(uint32_t tmem_lane_base, uint64_t* t_full, uint64_t* t_empty, uint32_t item_base, uint32_t hn,
                                           uint32_t& use0, uint32_t& use1, int tile, const float* __restrict__ xrow, const float* shv,
                                           float2 (&acc)[Cfg::D_OUT / 2], int lane, long long* dbg, int pi) {
    if constexpr (C < END) {
        const bool st_on = PROBE && blockIdx.x == 0 && lane == 0 && (threadIdx.x >> 5 & 3) == 0 && (pi == 2 || pi == 3);
        long long* sp = dbg + (((1 + tile) * 2 + (pi - 2)) * 40 + C) * 3;
        if (st_on) sp[0] = clock64();
        const uint32_t it = item_base + C + (C >= HPOS ? hn : 0u), slot = 2 * (it & 1) + (uint32_t)tile;
        const uint32_t use = (it & 1) ? use1 : use0;
        if (it & 1) ++use1; else ++use0;
        if (lane == 0) tc_mbar_wait(&t_full[slot], use & 1);
        __syncwarp();
        if (st_on) sp[1] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const uint32_t tslot = tmem_lane_base + slot * CF2_SLOT_COLS;
            float wv[3][8];
            float2 zz[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            cf2_tmem_ld8(tslot, wv[0]);
            if constexpr (Cf2Chunk<Cfg, C>::NV > 8) cf2_tmem_ld8(tslot + 8, wv[1]);
            cf2_cols<Cfg, C, 0>(tslot, wv, zz, xrow, shv, acc);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&t_empty[slot]);
        if (st_on) sp[2] = clock64();
        cf2_chunks<Cfg, C + 1, END, HPOS, PROBE>(tmem_lane_base, t_full, t_empty, item_base, hn, use0, use1, tile, xrow, shv, acc, lane, dbg, pi);
    }
}

// stage quads [J PQ, J PQ + PQ) of this thread's output row
template <class Cfg, int J, int Q>
__device__ __forceinline__ void cf2_stage_part(const float2 (&acc)[Cfg::D_OUT / 2], float* dst) {
    using S = ConvFused2Smem<Cfg>;
    if constexpr (Q < S::PQ && J * S::PQ + Q < S::OQ) {
        constexpr int G = J * S::PQ + Q;
        *reinterpret_cast<float4*>(dst + 4 * Q) = make_float4(cf_acc_get<Cfg, 4 * G>(acc), cf_acc_get<Cfg, 4 * G + 1>(acc),
                                                              cf_acc_get<Cfg, 4 * G + 2>(acc), cf_acc_get<Cfg, 4 * G + 3>(acc));
        cf2_stage_part<Cfg, J, Q + 1>(acc, dst);
    }
}

// exactly-scaled FP16 split of one operand row (64 values): row * 2^s with max in [2^12, 2^13), hi -> TMEM (8 columns per
// tcgen05.st), lo -> smem (K-major core matrices [k/8][m/8][m%8][k%8]).  Returns 2^-s.  Same arithmetic as conv_fused.cuh.
__device__ __forceinline__ void cf2_tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float cf2_put_operand(const float (&v)[64], uint32_t hi_taddr, uint8_t* lo_dst) {
    float m = 1.0f;                                             // >= 1: column 60 holds the constant 1.0
#pragma unroll
    for (int q = 0; q < 64; ++q) m = fmaxf(m, fabsf(v[q]));
    const int ex = (int)((__float_as_uint(m) >> 23) & 0xFF) - 127;
    const float sc = __uint_as_float((uint32_t)(127 + 12 - ex) << 23);
    const float2 sc2 = make_float2(sc, sc), neg1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {                               // 16 values: 8 packed hi registers, two 16-byte lo stores
        uint32_t hi_p[8], lo_p[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float2 x = __fmul2_rn(make_float2(v[16 * g + 2 * c], v[16 * g + 2 * c + 1]), sc2);
            const __half2 hh = __floats2half2_rn(x.x, x.y);
            const float2 lo = __ffma2_rn(__half22float2(hh), neg1, x);                  // x - hi, exact
            const __half2 ll = __floats2half2_rn(lo.x, lo.y);
            hi_p[c] = *reinterpret_cast<const uint32_t*>(&hh);
            lo_p[c] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        cf2_tmem_st8(hi_taddr + 8 * g, hi_p);
        *reinterpret_cast<uint4*>(lo_dst + (2 * g) * 2048) = make_uint4(lo_p[0], lo_p[1], lo_p[2], lo_p[3]);
        *reinterpret_cast<uint4*>(lo_dst + (2 * g + 1) * 2048) = make_uint4(lo_p[4], lo_p[5], lo_p[6], lo_p[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    return __uint_as_float((uint32_t)(127 - 12 + ex) << 23);
}

// profiling aid (tools/cf2_phases.py): clock64() stamps of CTA 0, pairs 2 and 3: [role (5)][pair (2)][idx (40)][3];
// role 0 = MMA issuer of tile 0, 1 / 2 = worker warp 0 / 4, 3 = prep warp 8, 4 = window copier
#define CF2_STAMP(role, pr, idx, k) do { if (PROBE && blockIdx.x == 0 && lane == 0 && ((pr) == 2 || (pr) == 3)) \
    a.dbg[((((role) * 2 + ((pr) - 2)) * 40) + (idx)) * 3 + (k)] = clock64(); } while (0)

template <class Cfg, bool PROBE>
__global__ void __launch_bounds__(CF2_THREADS, 1) conv_fused2_kernel(ConvFusedArgs a) {
    using S = ConvFused2Smem<Cfg>;
    constexpr int NCH = (Cfg::W + CF2_N - 1) / CF2_N;
    constexpr int N_LAST = ((Cfg::W - (NCH - 1) * CF2_N + 15) / 16) * 16;
    constexpr int HPOS = NCH > 5 ? NCH - 4 : 1;                         // the next pair's hidden-layer item goes before chunk HPOS
    static_assert(NCH >= 2 && HPOS < NCH, "chunk stream too short");
    extern __shared__ __align__(1024) uint8_t cf2_smem_raw[];
    uint8_t* b_st = cf2_smem_raw;                                                   // STAGES x (hi | lo)
    uint8_t* alo = b_st + S::STAGES * CF2_B_STAGE;                                  // [buffer][tile] x 16 KB (A lo)
    float* xwin = reinterpret_cast<float*>(alo + 4 * CF_ALO_TILE);                  // [XCAP][XS] window of node rows
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xwin) + S::X_BYTES);      // [256][PS] staged output parts
    float* rsv = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stg) + S::STG_BYTES);    // [buffer][256] row scales
    int* node_seg = reinterpret_cast<int*>(rsv + 512);                                         // [tile][<= 257 (264)] seg_ptr of the pair's nodes
    int* nsplit = node_seg + 2 * 264;                                                          // [tile] nodes that start in MMA tile 0
    float* strad_buf = reinterpret_cast<float*>(nsplit + 4);                                   // [part (4)][quad (16)][4] partial sums of the straddling node
    float* osc = strad_buf + 4 * 16 * 4;
    float* osh = osc + 104;
    uint64_t* bars = reinterpret_cast<uint64_t*>(osh + 104);
    uint64_t *b_full = bars, *b_empty = bars + 6, *t_full = bars + 12, *t_empty = bars + 16,
             *a1_ready = bars + 20,      // [buffer][tile]: layer-1 operand in place (prep -> issuer)
             *a2_ready = bars + 24,      // [buffer][tile]: layer-2 operand + row scale in place (prep -> issuer, workers)
             *a_free = bars + 28,        // [buffer][tile]: the pair's last chunk MMA has completed (issuer -> prep)
             *x_ready = bars + 32, *x_free = bars + 33,
             *h_full = bars + 34,        // [tile]: hidden-layer accumulator of the next pair complete (issuer -> prep), one phase per pair
             *spare_bar = bars + 36;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);
    // parts reduced by tile 0 so far (its partial sums of the straddling node are in smem).  A counter, not an mbarrier: tile 0 is not
    // throttled by tile 1, and a parity wait breaks as soon as the producer runs two phases ahead.
    int* strad_cnt = reinterpret_cast<int*>(tmem_slot + 2);
    volatile int* win_lo_s = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = __reduce_max_sync(0xffffffffu, a.n_tiles_dev ? *a.n_tiles_dev : a.n_tiles);
    if ((int)blockIdx.x >= n_tiles) return;
    const int my_pairs = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < S::STAGES; ++i) { tc_mbar_init(&b_full[i], 1); tc_mbar_init(&b_empty[i], 2); }
        for (int i = 0; i < 4; ++i) {
            tc_mbar_init(&t_full[i], 1); tc_mbar_init(&t_empty[i], 4);
            tc_mbar_init(&a1_ready[i], 4); tc_mbar_init(&a2_ready[i], 4); tc_mbar_init(&a_free[i], 1);
        }
        tc_mbar_init(x_ready, 1);
        tc_mbar_init(spare_bar, 1);
        *strad_cnt = 0;
        tc_mbar_init(&h_full[0], 1);
        tc_mbar_init(&h_full[1], 1);
        tc_mbar_init(x_free, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Cfg::D_OUT; i += CF2_THREADS) { osc[i] = a.oscale[i]; osh[i] = a.oshift[i]; }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CF2_REG_AUX));
        if (warp == 14) {
            // ================= TMA producer (one thread): the weight stream in the order the issuers consume it =================
            if (lane == 0) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w2img);
                const uint32_t total = (uint32_t)my_pairs * (NCH + 1);          // every pair: its hidden-layer item + NCH chunks
                for (uint32_t g = 0; g < total; ++g) {
                    const uint32_t s = g % S::STAGES, u = g / S::STAGES;
                    // stream: H(0) | pair p: C(0..HPOS-1), H(p+1), C(HPOS..NCH-1) | last pair: C(0..NCH-1)
                    int c;                                                          // chunk index, or -1 for first-layer weights
                    if (g == 0) c = -1;
                    else {
                        const uint32_t r = g - 1, p = r / (NCH + 1), j = r % (NCH + 1);
                        const bool has_next = (int)p + 1 < my_pairs;
                        c = !has_next ? (int)j : ((int)j < HPOS ? (int)j : ((int)j == HPOS ? -1 : (int)j - 1));
                    }
                    cf2_mbar_wait_relaxed(&b_empty[s], (u & 1) ^ 1, 64);
                    if (c < 0) {
                        tc_mbar_expect_tx(&b_full[s], CF_W1_BYTES);
                        tc_bulk_load(b_st + s * CF2_B_STAGE, a.w1img, CF_W1_BYTES, &b_full[s]);
                    } else {
                        tc_mbar_expect_tx(&b_full[s], CF2_B_STAGE);
                        tc_bulk_load(b_st + s * CF2_B_STAGE, wsrc + (size_t)c * CF2_B_STAGE, CF2_B_STAGE, &b_full[s]);
                    }
                }
            }
        } else if (warp == 15) {
            // ================= node-row window of every pair -> shared memory =================
            for (int p = 0; p < my_pairs; ++p) {
                if (p > 0) CF2_STAMP(4, p - 1, 0, 2);
                const int pr = (int)blockIdx.x + p * (int)gridDim.x;
                const int n_lo = a.tile_node[pr], n_hi = a.tile_node[pr + 1];
                const int e0 = a.seg_ptr[n_lo], ne_pair = a.seg_ptr[n_hi] - e0;
                CF2_STAMP(4, p, 0, 0);
                int lo = 0x7fffffff, hi = -1;
                for (int e = lane; e < ne_pair; e += 32) {
                    const int s = a.gather_idx ? a.gather_idx[e0 + e] : e0 + e;
                    lo = min(lo, s); hi = max(hi, s);
                }
                lo = __reduce_min_sync(0xffffffffu, lo);
                hi = __reduce_max_sync(0xffffffffu, hi);
                const int rows = hi >= lo ? hi - lo + 1 : 0;
                if (p >= 1) {
                    if (lane == 0) cf2_mbar_wait_relaxed(x_free, (uint32_t)((p - 1) & 1), 200);   // the workers are done with the previous window
                    __syncwarp();
                }
                CF2_STAMP(4, p, 0, 1);
                const bool fits = rows <= S::XCAP;
                if (lane == 0) *win_lo_s = fits ? (rows ? lo : 0) : -1;
                if constexpr (S::XVEC == 4) {
                    // 16-byte aligned rows: one bulk copy (TMA) per row into the padded window
                    if (fits && rows > 0) {
                        if (lane == 0) tc_mbar_expect_tx(x_ready, (uint32_t)rows * Cfg::D_IN * 4);
                        __syncwarp();
                        for (int r = lane; r < rows; r += 32)
                            tc_bulk_load(xwin + (size_t)r * S::XS, a.node_in + (size_t)(lo + r) * Cfg::D_IN, Cfg::D_IN * 4, x_ready);
                    } else {
                        __syncwarp();
                        if (lane == 0) tc_mbar_arrive(x_ready);
                    }
                } else {
                    if (fits) {
                        constexpr int NV2 = Cfg::D_IN / 2;
                        const float2* src = reinterpret_cast<const float2*>(a.node_in + (size_t)lo * Cfg::D_IN);
                        for (int i0 = 0; i0 < rows * NV2; i0 += 32 * 8) {
                            float2 v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int i = i0 + u * 32 + lane;
                                v[u] = i < rows * NV2 ? __ldg(src + i) : make_float2(0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int i = i0 + u * 32 + lane;
                                if (i < rows * NV2) *reinterpret_cast<float2*>(xwin + (size_t)(i / NV2) * S::XS + 2 * (i % NV2)) = v[u];
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(x_ready);
                }
            }
        } else if (warp >= 12) {
            // ================= MMA issuers: warp 12 -> tile 0 (accumulator slots 0, 2), warp 13 -> tile 1 (slots 1, 3) =================
            const int t = __reduce_max_sync(0xffffffffu, warp - 12);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(CF2_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // first layer: N = 64
            const uint32_t idesc_last = (1u << 4) | ((uint32_t)(N_LAST >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t b_base = tc_smem(b_st);
            uint32_t g = 0, it = 0;                                          // ring position, items issued by this warp
            // one item: wait for its weights and its accumulator slot, 12 MMAs (hi*hi + hi*lo from TMEM A, lo*hi from smem A)
#define CF2_ISSUE(ABUF, NCOLS, IDESC, BHALF, FULLBAR)                                                                               \
            do {                                                                                                              \
                const uint32_t s_ = g % S::STAGES;                                                                            \
                if (t == 0) CF2_STAMP(0, st_pair, st_idx, 0);                                                                 \
                cf2_wait3(&b_full[s_], (g / S::STAGES) & 1, &t_empty[2 * (it & 1) + t], ((it >> 1) & 1) ^ 1, &b_full[s_],      \
                         (g / S::STAGES) & 1, lane);                                                                          \
                if (t == 0) CF2_STAMP(0, st_pair, st_idx, 1);                                                                 \
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                               \
                const uint32_t b_hi_s = b_base + s_ * CF2_B_STAGE, b_lo_s = b_hi_s + (BHALF);                                 \
                const uint32_t slot_ = 2 * (it & 1) + (uint32_t)t;                                                            \
                const uint32_t d_ = tmem_base + slot_ * CF2_SLOT_COLS;                                                        \
                const uint32_t a_hi_t = tmem_base + CF2_A_COL + (uint32_t)((ABUF) * 64 + t * 32);                             \
                const uint32_t a_lo_s = tc_smem(alo) + (uint32_t)(((ABUF) * 2 + t) * CF_ALO_TILE);                            \
                _Pragma("unroll") for (int combo = 0; combo < 3; ++combo) {                                                   \
                    const uint32_t bs = combo == 1 ? b_lo_s : b_hi_s;                                                         \
                    _Pragma("unroll") for (int ks = 0; ks < TC_K / 16; ++ks) {                                                \
                        const uint64_t bd = tc_smem_desc(bs + ks * 2 * ((NCOLS) * 16), (NCOLS) * 16, 128);                    \
                        if (combo < 2) cf_mma_f16_ts(d_, a_hi_t + (uint32_t)(ks * 8), bd, (IDESC), (combo | ks) ? 1u : 0u);   \
                        else cf_mma_f16_ss(d_, tc_smem_desc(a_lo_s + ks * 2 * 2048, 2048, 128), bd, (IDESC), 1u);             \
                    }                                                                                                         \
                }                                                                                                             \
                cf_commit(FULLBAR);                                                                                           \
                cf_commit(&b_empty[s_]);                                                                                      \
                if (t == 0) CF2_STAMP(0, st_pair, st_idx, 2);                                                                 \
                ++st_idx;                                                                                                     \
                ++g; ++it;                                                                                                    \
            } while (0)
            int st_pair = -1, st_idx = 0;                                    // (profiling stamps)
            // hidden layer of pair 0
            cf2_wait3(&a1_ready[t], 0u, &a1_ready[t], 0u, &a1_ready[t], 0u, lane);
            CF2_ISSUE(0, 64, idesc1, CF_W1_BYTES / 2, &h_full[t]);
            for (int p = 0; p < my_pairs; ++p) {
                const int b = __reduce_max_sync(0xffffffffu, p & 1);
                const uint32_t ph = (uint32_t)((p >> 1) & 1);
                st_pair = p; st_idx = 0;
                cf2_wait3(&a2_ready[2 * b + t], ph, &a2_ready[2 * b + t], ph, &a2_ready[2 * b + t], ph, lane);   // layer-2 operand
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < NCH; ++c) {
                    if (c == HPOS && p + 1 < my_pairs) {                    // hidden layer of the next pair (its operand is in the other buffer)
                        const uint32_t ph1 = (uint32_t)(((p + 1) >> 1) & 1);
                        cf2_wait3(&a1_ready[2 * (b ^ 1) + t], ph1, &a1_ready[2 * (b ^ 1) + t], ph1, &a1_ready[2 * (b ^ 1) + t], ph1, lane);
                        CF2_ISSUE(b ^ 1, 64, idesc1, CF_W1_BYTES / 2, &h_full[t]);
                    }
                    const uint32_t idc = (c == NCH - 1) ? idesc_last : idesc;
                    CF2_ISSUE(b, CF2_N, idc, CF2_B_HALF, &t_full[slot_]);
                }
                cf_commit(&a_free[2 * b + t]);                              // all MMAs reading this pair's operands have completed
            }
#undef CF2_ISSUE
        } else {
            // ================= prep warpgroup (warps 8-11): operands of the pair tiles, one pair ahead of the workers =================
            const int wq = warp & 3, row = wq * 32 + lane;
            const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
            float rs1[2] = {0.f, 0.f};
            bool valid[2] = {false, false};
            // ---- layer-1 operands of pair q: edge attributes [emb | node B | node C (+ C2) | 1 | 0 0 0] of both tiles
            auto layer1 = [&](int q) {
                const int b = q & 1, pr = (int)blockIdx.x + q * (int)gridDim.x;
                const int n_lo = a.tile_node[pr], n_hi = a.tile_node[pr + 1];
                const int e0 = a.seg_ptr[n_lo], ne_pair = a.seg_ptr[n_hi] - e0;
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    const int eb = e0 + 128 * t, ne = min(max(ne_pair - 128 * t, 0), 128);
                    valid[t] = row < ne;
                    if (wq == 0) CF2_STAMP(3, q, t, 0);
                    float at[64];
                    if (valid[t]) {
                        const int e = eb + row;
                        const int ce = a.perm ? a.perm[e] : e, ib = a.idxB[e], ic = a.idxC[e];
                        const int ic2 = a.idxC2 ? a.idxC2[e] : -1;
                        const float4* pe = reinterpret_cast<const float4*>(a.emb + (size_t)ce * 20);
                        const float2* pb = reinterpret_cast<const float2*>(a.tb + (size_t)ib * a.strideB);
                        const float2* pc = reinterpret_cast<const float2*>(a.tc + (size_t)ic * a.strideC);
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const float4 v = __ldg(pe + k);
                            at[4 * k] = v.x; at[4 * k + 1] = v.y; at[4 * k + 2] = v.z; at[4 * k + 3] = v.w;
                        }
                        if (((a.strideB | a.strideC) & 3) == 0) {
                            const float4* pb4 = reinterpret_cast<const float4*>(pb);
                            const float4* pc4 = reinterpret_cast<const float4*>(pc);
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                const float4 vb = __ldg(pb4 + k), vc = __ldg(pc4 + k);
                                at[20 + 4 * k] = vb.x; at[21 + 4 * k] = vb.y; at[22 + 4 * k] = vb.z; at[23 + 4 * k] = vb.w;
                                at[40 + 4 * k] = vc.x; at[41 + 4 * k] = vc.y; at[42 + 4 * k] = vc.z; at[43 + 4 * k] = vc.w;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 10; ++k) {
                                const float2 vb = __ldg(pb + k), vc = __ldg(pc + k);
                                at[20 + 2 * k] = vb.x; at[21 + 2 * k] = vb.y;
                                at[40 + 2 * k] = vc.x; at[41 + 2 * k] = vc.y;
                            }
                        }
                        if (ic2 >= 0) {
                            const float2* pc2 = reinterpret_cast<const float2*>(a.tc + (size_t)ic2 * a.strideC);
#pragma unroll
                            for (int k = 0; k < 10; ++k) {
                                const float2 vc = __ldg(pc2 + k);
                                at[40 + 2 * k] += vc.x; at[41 + 2 * k] += vc.y;
                            }
                        }
                        at[60] = 1.0f; at[61] = 0.f; at[62] = 0.f; at[63] = 0.f;       // k = 60 multiplies the bias row of W1aug
                    } else {
#pragma unroll
                        for (int k = 0; k < 64; ++k) at[k] = 0.f;
                    }
                    if (wq == 0) CF2_STAMP(3, q, t, 1);
                    if (q >= 2) {                                           // buffer b, tile t: pair q - 2 must be through the tensor pipe
                        if (lane == 0) cf2_mbar_wait_relaxed(&a_free[2 * b + t], (uint32_t)(((q - 2) >> 1) & 1), 100);
                        __syncwarp();
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    uint8_t* lo_dst = alo + (2 * b + t) * CF_ALO_TILE + (row >> 3) * 128 + (row & 7) * 16;
                    rs1[t] = cf2_put_operand(at, lane_base + CF2_A_COL + (uint32_t)(b * 64 + t * 32), lo_dst) * a.inv_w1scale;
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(&a1_ready[2 * b + t]);
                    if (wq == 0) CF2_STAMP(3, q, t, 2);
                }
            };
            // ---- layer-2 operands of pair q: h = ReLU(D1) from its hidden-layer accumulator (rs1 / valid: the preceding layer1(q))
            auto layer2 = [&](int q) {
                const int b = q & 1;
                const uint32_t it_h = q == 0 ? 0u : (uint32_t)((q - 1) * NCH + q + HPOS);
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    const uint32_t slot = 2 * (it_h & 1) + (uint32_t)t;
                    if (wq == 0) CF2_STAMP(3, q, 2 + t, 0);
                    if (lane == 0) cf2_mbar_wait_relaxed(&h_full[t], (uint32_t)(q & 1), 100);
                    __syncwarp();
                    if (wq == 0) CF2_STAMP(3, q, 2 + t, 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    float h[64];
#pragma unroll
                    for (int k = 0; k < 4; ++k) cf_tmem_ld16(lane_base + slot * CF2_SLOT_COLS + 16 * k, h + 16 * k);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 4; ++k) cf_wait_ld16(h + 16 * k);      // (register dependency only)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(&t_empty[slot]);
                    const float r1 = rs1[t];
#pragma unroll
                    for (int k = 0; k < 60; ++k) h[k] = fmaxf(h[k] * r1, 0.f);
                    h[60] = valid[t] ? 1.0f : 0.f; h[61] = 0.f; h[62] = 0.f; h[63] = 0.f;   // k = 60 multiplies the bias row of W2aug
                    uint8_t* lo_dst = alo + (2 * b + t) * CF_ALO_TILE + (row >> 3) * 128 + (row & 7) * 16;
                    const float rs = cf2_put_operand(h, lane_base + CF2_A_COL + (uint32_t)(b * 64 + t * 32), lo_dst) * a.inv_wscale;
                    rsv[b * 256 + t * 128 + row] = rs;
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(&a2_ready[2 * b + t]);
                    if (wq == 0) CF2_STAMP(3, q, 2 + t, 2);
                }
            };
            for (int q = 0; q < my_pairs; ++q) { layer1(q); layer2(q); }
        }
    } else {
        // ================= workers: thread = edge =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CF2_REG_WORK));
        const int tile = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        struct Idx { int n_lo, n_hi, e0, ne, src, ce; };
        auto fetch = [&](int pr) {
            Idx x;
            x.n_lo = a.tile_node[pr]; x.n_hi = a.tile_node[pr + 1];
            x.src = -1; x.ce = 0;
            const int e0 = a.seg_ptr[x.n_lo], ne_pair = a.seg_ptr[x.n_hi] - e0;
            x.e0 = e0;
            const int eb = e0 + 128 * tile;                              // MMA tile 0: edges 0..127 of the pair tile, tile 1: the rest
            x.ne = min(max(ne_pair - 128 * tile, 0), 128);
            if (row < x.ne) {
                x.src = a.gather_idx ? a.gather_idx[eb + row] : eb + row;
                x.ce = a.perm ? a.perm[eb + row] : eb + row;
            }
            return x;
        };
        Idx ix = fetch((int)blockIdx.x);
        uint32_t use0 = 0, use1 = 0;                                     // chunk items drained from the even / odd slots
        uint32_t kpart = 0;                                              // parts reduced so far (phase of strad_ready)
        const int ltid = tid & 127;                                      // thread within the tile's worker group
        int* nseg = node_seg + tile * 264;                               // this tile's copy of the pair's seg_ptr
        float shn[Cfg::SH_USED];                                         // spherical harmonics of the coming pair's edge
#pragma unroll
        for (int i = 0; i < Cfg::SH_USED; ++i) shn[i] = row < ix.ne ? __ldg(a.sh + (size_t)ix.ce * a.sh_stride + i) : 0.f;
        for (int pi = 0; pi < my_pairs; ++pi) {
            const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
            const int b = pi & 1;
            const bool valid = row < ix.ne;
            const uint32_t hn = pi + 1 < my_pairs ? 1u : 0u;
            const int nnodes = ix.n_hi - ix.n_lo;
            if (wq == 0) CF2_STAMP(1 + tile, pi, 30, 0);
            // seg_ptr of the pair's nodes -> this tile's smem copy, and the number of nodes that start in MMA tile 0 (rows < 128);
            // the last barrier of the previous epilogue protects the reuse, the first barrier of this one publishes it
            for (int i = ltid; i <= nnodes; i += 128) {
                const int s_i = a.seg_ptr[ix.n_lo + i];
                nseg[i] = s_i;
                if (i < nnodes && s_i - ix.e0 < 128 && (i == nnodes - 1 || a.seg_ptr[ix.n_lo + i + 1] - ix.e0 >= 128)) nsplit[tile] = i + 1;
            }
            // layer-2 operand of this pair in place -> row scale; the power-of-two operand scales are undone exactly by scaling
            // the spherical harmonics (Z is linear in them)
            if (lane == 0) tc_mbar_wait(&a2_ready[2 * b + tile], (uint32_t)((pi >> 1) & 1));
            __syncwarp();
            float shv[Cfg::SH_USED];
            {
                const float rs = rsv[b * 256 + tile * 128 + row];
#pragma unroll
                for (int i = 0; i < Cfg::SH_USED; ++i) shv[i] = shn[i] * rs;
            }
            if (wq == 0) CF2_STAMP(1 + tile, pi, 30, 1);
            if (lane == 0) tc_mbar_wait(x_ready, (uint32_t)(pi & 1));
            __syncwarp();
            if (wq == 0) CF2_STAMP(1 + tile, pi, 30, 2);
            const int w_lo = *win_lo_s;
            const float* xrow = w_lo >= 0 ? xwin + (size_t)((valid ? ix.src : w_lo) - w_lo) * S::XS
                                          : a.node_in + (size_t)(valid ? ix.src : 0) * Cfg::D_IN;
            float2 acc[Cfg::D_OUT / 2];
#pragma unroll
            for (int d = 0; d < Cfg::D_OUT / 2; ++d) acc[d] = make_float2(0.f, 0.f);
            const uint32_t item_base = (uint32_t)(pi * NCH + pi + 1);
            Idx nx = ix;
            constexpr int CSPLIT = NCH > 4 ? 4 : NCH;
            cf2_chunks<Cfg, 0, CSPLIT, HPOS, PROBE>(lane_base, t_full, t_empty, item_base, hn, use0, use1, tile, xrow, shv, acc, lane, a.dbg, pi);
            if (hn) nx = fetch(pair + (int)gridDim.x);                    // three dependent global loads, hidden in the chunk loop
            cf2_chunks<Cfg, CSPLIT, NCH, HPOS, PROBE>(lane_base, t_full, t_empty, item_base, hn, use0, use1, tile, xrow, shv, acc, lane, a.dbg, pi);
            __syncwarp();
            if constexpr (PROBE) {                                        // pin the accumulators here: the stamps must not see sunk FFMA2 work
#pragma unroll
                for (int d = 0; d < Cfg::D_OUT / 2; ++d) asm volatile("" : "+f"(acc[d].x), "+f"(acc[d].y));
            }
            if (wq == 0) CF2_STAMP(1 + tile, pi, 31, 0);
            if (lane == 0) tc_mbar_arrive(x_free);                        // this warp is done with the node-row window
            // the coming pair's spherical harmonics: in flight across the epilogue
#pragma unroll
            for (int i = 0; i < Cfg::SH_USED; ++i) shn[i] = (hn && row < nx.ne) ? __ldg(a.sh + (size_t)nx.ce * a.sh_stride + i) : 0.f;
            // ---- epilogue, TILE-LOCAL: each MMA tile's 128 threads stage their rows part by part and reduce the nodes of their
            // rows on their own (named barrier of 128 threads); the one node that may straddle row 128 is folded by tile 0 up to
            // row 127, handed over through smem (strad_ready) and continued by tile 1 - the same left-to-right sum as one loop.
            cf_bar_tile(tile);                                            // nseg / nsplit of this pair are complete
            const int e_lo = nseg[0], nA = nsplit[tile];
            if (nseg[nnodes] - e_lo > 256) __trap();                      // tile builder contract violated
            const bool strad = nseg[nA] - e_lo > 128;                     // node nA - 1 continues in MMA tile 1
            if (wq == 0) CF2_STAMP(1 + tile, pi, 32, 0);
            const int n_first = tile == 0 ? 0 : (strad ? nA - 1 : nA), n_cnt = tile == 0 ? nA : nnodes - n_first;
            auto load_add = [&](auto part_c, int it, float (&ad)[4]) {
                constexpr int part = decltype(part_c)::value;
                constexpr int pqj = S::PQ < S::OQ - part * S::PQ ? S::PQ : S::OQ - part * S::PQ;        // compile-time divisor
                const int nl = n_first + it / pqj, qd = part * S::PQ + it % pqj, node = ix.n_lo + nl;
                const bool on = it < n_cnt * pqj && !(tile == 0 && strad && nl == nA - 1);     // (tile 0 only hands the straddler's partial sum over)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = 4 * qd + j;
                    ad[j] = 0.f;
                    if (on && d < Cfg::D_OUT) {
                        if (a.mode == 1) ad[j] = (d < a.res_dim) ? __ldg(a.residual + (size_t)node * a.res_dim + d) : 0.0f;
                        else if (a.mode == 2 && Cfg::D_OUT % 4 != 0) ad[j] = a.out[(size_t)node * Cfg::D_OUT + d];
                    }
                }
                if constexpr (Cfg::D_OUT % 4 == 0) {                     // 16-byte aligned output rows: one load per quad
                    if (a.mode == 2 && on && 4 * qd < Cfg::D_OUT) {
                        const float4 tt = *reinterpret_cast<const float4*>(a.out + (size_t)node * Cfg::D_OUT + 4 * qd);
                        ad[0] = tt.x; ad[1] = tt.y; ad[2] = tt.z; ad[3] = tt.w;
                    }
                }
            };
            // residual / running-sum operand of this thread's first item of a part: loaded one part ahead (4 + 4 registers)
            float add_cur[4], add_nxt[4] = {0.f, 0.f, 0.f, 0.f};
            load_add(std::integral_constant<int, 0>{}, ltid, add_cur);
            if (wq == 0) CF2_STAMP(1 + tile, pi, 32, 1);
            auto reduce_part = [&](auto part_c, const float (&ad0)[4]) {
                constexpr int part = decltype(part_c)::value;
                constexpr int pqj = S::PQ < S::OQ - part * S::PQ ? S::PQ : S::OQ - part * S::PQ;
                const int items = n_cnt * pqj;
#pragma unroll 1
                for (int it = ltid; it < items; it += 128) {
                    float ad[4];
                    if (it == ltid) { ad[0] = ad0[0]; ad[1] = ad0[1]; ad[2] = ad0[2]; ad[3] = ad0[3]; }
                    else load_add(part_c, it, ad);
                    const int nl = n_first + it / pqj, ql = it % pqj, qd = part * S::PQ + ql, node = ix.n_lo + nl;
                    const int s0 = nseg[nl] - e_lo, s1 = nseg[nl + 1] - e_lo;
                    const bool is_strad = strad && nl == nA - 1;
                    const int r0 = tile == 0 ? s0 : max(s0, 128), r1 = tile == 0 ? min(s1, 128) : s1;       // this tile's rows of the node
                    float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (tile == 1 && is_strad) sm = *reinterpret_cast<const float4*>(strad_buf + (part * 16 + ql) * 4);
                    const float* sp = stg + (size_t)r0 * S::PS + 4 * ql;                                // staged rows = the pair tile's edges in order
#pragma unroll 4
                    for (int r = 0; r < r1 - r0; ++r) {                  // sequential in edge order: composition-invariant
                        const float4 v = *reinterpret_cast<const float4*>(sp + (size_t)r * S::PS);
                        sm.x += v.x; sm.y += v.y; sm.z += v.z; sm.w += v.w;
                    }
                    if (tile == 0 && is_strad) {                         // partial sum over rows < 128: tile 1 continues it
                        *reinterpret_cast<float4*>(strad_buf + (part * 16 + ql) * 4) = sm;
                        continue;
                    }
                    const int deg = s1 - s0;
                    const float inv_deg = 1.0f / (float)(deg > 0 ? deg : 1);
                    const float sv[4] = {sm.x, sm.y, sm.z, sm.w};
                    float ov[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = 4 * qd + j;
                        ov[j] = d < Cfg::D_OUT ? sv[j] * inv_deg * osc[d] + osh[d] + ad[j] : 0.f;
                    }
                    float* orow = a.out + (size_t)node * Cfg::D_OUT;
                    if constexpr (Cfg::D_OUT % 4 == 0) {
                        if (4 * qd < Cfg::D_OUT) *reinterpret_cast<float4*>(orow + 4 * qd) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (4 * qd + j < Cfg::D_OUT) orow[4 * qd + j] = ov[j];
                    }
                }
            };
            float* my_stg = stg + (size_t)tid * S::PS;
#define CF2_PART(J)                                                                                                         \
            if constexpr (S::NPART > (J)) {                                                                                 \
                cf2_stage_part<Cfg, (J), 0>(acc, my_stg);                                                                   \
                if ((J) == 0 && wq == 0) CF2_STAMP(1 + tile, pi, 32, 2);                                                    \
                cf_bar_tile(tile);                                        /* this tile's rows of the part are staged */     \
                if ((J) == 0 && wq == 0) CF2_STAMP(1 + tile, pi, 33, 0);                                                    \
                if (tile == 1) {                                          /* tile 0's partial sum of the straddling node */ \
                    if (lane == 0) cf2_flag_wait(strad_cnt, (int)kpart + 1);                                                \
                    __syncwarp();                                                                                           \
                }                                                                                                           \
                if ((J) == 0 && wq == 0) CF2_STAMP(1 + tile, pi, 33, 1);                                                    \
                if constexpr ((J) + 1 < S::NPART) load_add(std::integral_constant<int, ((J) + 1 < S::NPART ? (J) + 1 : 0)>{}, ltid, add_nxt); \
                reduce_part(std::integral_constant<int, (J)>{}, add_cur);                                                   \
                add_cur[0] = add_nxt[0]; add_cur[1] = add_nxt[1]; add_cur[2] = add_nxt[2]; add_cur[3] = add_nxt[3];         \
                if ((J) == 0 && wq == 0) CF2_STAMP(1 + tile, pi, 33, 2);                                                    \
                cf_bar_tile(tile);                                        /* staging free again, partial sums written */    \
                if (tile == 0 && ltid == 0) cf2_flag_set(strad_cnt, (int)kpart + 1);                                        \
                ++kpart;                                                                                                    \
                if ((J) == 0 && wq == 0) CF2_STAMP(1 + tile, pi, 31, 1);                                                    \
            }
            CF2_PART(0) CF2_PART(1) CF2_PART(2) CF2_PART(3)
#undef CF2_PART
            static_assert(S::NPART <= 4, "epilogue parts");
            if (wq == 0) CF2_STAMP(1 + tile, pi, 31, 2);
            ix = nx;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

template <class Cfg>
static int conv_fused2_launch(const ConvFusedArgs& a, cudaStream_t st) {
    if (a.n_tiles <= 0) return DP_OK;
    using S = ConvFused2Smem<Cfg>;
    static bool attr_set = false;
    static int n_sm = 0;
    auto kern = conv_fused2_kernel<Cfg, false>;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
            return dp_check_launch("dp_conv_fused2(attr)");
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
    if constexpr (Cfg::W == 2200 || Cfg::W == 600) {
        if (a.dbg) {                                                       // profiling aid (tools/cf2_phases.py)
            auto pk = conv_fused2_kernel<Cfg, true>;
            cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
            pk<<<grid, CF2_THREADS, S::TOTAL, st>>>(a);
            return dp_check_launch("dp_conv_fused2(probe)");
        }
    }
    kern<<<grid, CF2_THREADS, S::TOTAL, st>>>(a);
    return dp_check_launch("dp_conv_fused2");
}
