// C ABI of libdiffphore_sm100.so — see include/diffphore_b200.h for the contract and reference citations.
// Single translation unit: all kernels are included here so that constant memory and templates link trivially.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tp_scatter.cuh"
#include "edge_mlp.cuh"
#include "edge_mlp_tc.cuh"
#include "tp_scatter_tma.cuh"
#include "conv_fused.cuh"
#include "graph_kernels.cuh"
#include "conformer.cuh"

static thread_local char g_err[512] = "";

void dp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int dp_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        dp_set_error("%s: %s", what, cudaGetErrorString(e));
        return DP_ERR_CUDA;
    }
    return DP_OK;
}

#define ST(s) ((cudaStream_t)(s))
#define NEED(cond, msg)                \
    do {                               \
        if (!(cond)) {                 \
            dp_set_error("%s", msg);   \
            return DP_ERR_ARG;         \
        }                              \
    } while (0)

extern "C" {

const char* dp_last_error(void) { return g_err; }
int dp_version(void) { return 100; }

int dp_set_constants(const DpConstants* c) {
    NEED(c != nullptr, "dp_set_constants: null");
    cudaError_t e = cudaMemcpyToSymbol(c_dp, c, sizeof(DpConstants));
    if (e != cudaSuccess) { dp_set_error("dp_set_constants: %s", cudaGetErrorString(e)); return DP_ERR_CUDA; }
    return DP_OK;
}

// CTAs per graph for the per-graph kernels that can share a graph (lig_fill, cross_step, tor_fill): 1 once the graphs alone fill the
// GPU twice over, up to 8 for small jobs
static int dp_graph_split(int n_graphs) {
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    const int s = (2 * n_sm) / (n_graphs > 0 ? n_graphs : 1);
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}

int dp_lig_graph(const float* pos, const int32_t* lig_ptr, const int32_t* bond_ptr, const int32_t* bond_dst,
                 const int32_t* bond_type, int32_t n_graphs, int32_t n_lig, int32_t max_atoms,
                 const DpSmallWeights* sw, const float* sc, int32_t* thr, int32_t* deg, int32_t* gcount,
                 int32_t* gstart, int32_t* seg_ptr, int32_t* e_src, int32_t* e_dst, float* e_emb, float* e_sh,
                 int32_t* n_edges_out, void* stream) {
    NEED(max_atoms <= LG_MAXN, "dp_lig_graph: more than 512 atoms in one ligand");
    if (n_graphs <= 0) return DP_OK;
    lig_count_kernel<<<n_graphs, LG_THREADS, 0, ST(stream)>>>(pos, lig_ptr, bond_ptr, thr, deg, gcount);
    scan_kernel<<<1, 1024, 0, ST(stream)>>>(gcount, gstart, n_graphs, n_edges_out);
    lig_fill_kernel<<<dim3(n_graphs, dp_graph_split(n_graphs)), LG_THREADS, 0, ST(stream)>>>(pos, lig_ptr, bond_ptr, bond_dst, bond_type, thr, deg, gstart,
                                                            n_graphs, *sw, sc, seg_ptr, e_src, e_dst, e_emb, e_sh);
    return dp_check_launch("dp_lig_graph");
}

int dp_pp_setup(const float* ppos, const int32_t* src, const int32_t* dst, int32_t n_edges, const DpSmallWeights* sw,
                float* pp_h, float* pp_sh, void* stream) {
    if (n_edges <= 0) return DP_OK;
    pp_setup_kernel<<<(n_edges + 127) / 128, 128, 0, ST(stream)>>>(ppos, src, dst, n_edges, *sw, pp_h, pp_sh);
    return dp_check_launch("dp_pp_setup");
}

int dp_pp_step(const float* pp_h, int32_t n_edges, const DpSmallWeights* sw, const float* sc, float* pp_emb, void* stream) {
    if (n_edges <= 0) return DP_OK;
    pp_step_kernel<<<(n_edges + 127) / 128, 128, 0, ST(stream)>>>(pp_h, n_edges, *sw, sc, pp_emb);
    return dp_check_launch("dp_pp_step");
}

int dp_cross_setup(const int32_t* cross_lig, const int32_t* cross_ph, int32_t n_edges, const float* phorefp,
                   const float* phoretype, const DpSmallWeights* sw, float* cross_h, float* cross_fm, void* stream) {
    if (n_edges <= 0) return DP_OK;
    cross_setup_kernel<<<(n_edges + 127) / 128, 128, 0, ST(stream)>>>(cross_lig, cross_ph, n_edges, phorefp, phoretype, *sw,
                                                                     cross_h, cross_fm);
    return dp_check_launch("dp_cross_setup");
}

int dp_cross_step(const float* lpos, const float* lnorm, const float* ppos, const float* pnorm, const int32_t* lig_ptr,
                  const int32_t* ph_ptr, const int32_t* cross_ptr, int32_t n_graphs, int32_t max_atoms,
                  const float* phorefp, const float* phoretype, const float* nangle1, const float* nangle2,
                  const float* cross_h, const float* cross_fm, const DpSmallWeights* sw, const float* sc,
                  float* tw_scratch, float* cross_emb, float* cross_sh, float* cross_nsh, void* stream) {
    if (n_graphs <= 0) return DP_OK;
    NEED(max_atoms * sizeof(float) <= 40000, "dp_cross_step: ligand too large");
    cross_step_kernel<<<dim3(n_graphs, dp_graph_split(n_graphs)), CR_THREADS, max_atoms * sizeof(float), ST(stream)>>>(
        lpos, lnorm, ppos, pnorm, lig_ptr, ph_ptr, cross_ptr, phorefp, phoretype, nangle1, nangle2, cross_h, cross_fm, *sw,
        sc, tw_scratch, cross_emb, cross_sh, cross_nsh);
    return dp_check_launch("dp_cross_step");
}

int dp_node_embed(const float* lpos, const float* ppos, const int32_t* lig_batch, const int32_t* ph_ptr,
                  const float* phoretype, const float* lig_static, const float* ph_static, int32_t n_lig, int32_t n_ph,
                  const DpSmallWeights* sw, const float* sc, float* lig_h0, float* ph_h0, void* stream) {
    const int n = n_lig + n_ph;
    if (n <= 0) return DP_OK;
    node_embed_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(lpos, ppos, lig_batch, ph_ptr, phoretype, lig_static, ph_static,
                                                              n_lig, n_ph, *sw, sc, lig_h0, ph_h0);
    return dp_check_launch("dp_node_embed");
}

int dp_edge_mlp(const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const float* w1,
                const float* b1, const float* w2t, int32_t in_dim, int32_t hid, int32_t W, const int32_t* n_edges_dev,
                int32_t n_edges_cap, float* w_out, void* stream) {
    EdgeMlpArgs a;
    a.emb = emb; a.perm = perm; a.tb = tb; a.idxB = idxB; a.strideB = strideB; a.tc = tc; a.idxC = idxC; a.idxC2 = idxC2;
    a.strideC = strideC; a.w1 = w1; a.b1 = b1; a.w2t = w2t; a.in_dim = in_dim; a.hid = hid; a.W = W;
    a.n_edges_dev = n_edges_dev; a.n_edges = n_edges_cap; a.out = w_out;
    NEED(in_dim == 40 || tc != nullptr, "dp_edge_mlp: part C missing");
    return edge_mlp_launch(a, ST(stream));
}

int dp_edge_mlp_tc(const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                   const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const float* w1,
                   const float* b1, const void* w2img, float inv_wscale, int32_t in_dim, int32_t hid, int32_t W,
                   const int32_t* n_edges_dev, int32_t n_edges_cap, float* h_scratch, float* w_out, void* stream) {
    EdgeMlpTcArgs t;
    EdgeMlpArgs& a = t.base;
    a.emb = emb; a.perm = perm; a.tb = tb; a.idxB = idxB; a.strideB = strideB; a.tc = tc; a.idxC = idxC; a.idxC2 = idxC2;
    a.strideC = strideC; a.w1 = w1; a.b1 = b1; a.w2t = nullptr; a.in_dim = in_dim; a.hid = hid; a.W = W;
    a.n_edges_dev = n_edges_dev; a.n_edges = n_edges_cap; a.out = w_out;
    t.w2img = w2img;
    t.inv_wscale = inv_wscale;
    t.himg = h_scratch;
    return edge_mlp_tc_launch(t, ST(stream));
}

int dp_build_tiles(const int32_t* seg_ptr, const int32_t* node_ptr, int32_t n_graphs, int32_t* cnt, int32_t* start,
                   int32_t* tile_node, int32_t* n_tiles_out, void* stream) {
    if (n_graphs <= 0) return DP_OK;
    tile_count_kernel<<<n_graphs, TILE_WALK_THREADS, 0, ST(stream)>>>(seg_ptr, node_ptr, n_graphs, cnt);
    scan_kernel<<<1, 1024, 0, ST(stream)>>>(cnt, start, n_graphs, n_tiles_out);
    tile_fill_kernel<<<n_graphs, TILE_WALK_THREADS, 0, ST(stream)>>>(seg_ptr, node_ptr, n_graphs, start, tile_node);
    return dp_check_launch("dp_build_tiles");
}

static long long* g_cf_dbg_host = nullptr;
/* profiling aid (not part of the public header): per-phase clock stamps of dp_conv_fused (layer 3) */
int dp_debug_set_cf_probe(long long* buf) { g_cf_dbg_host = buf; return DP_OK; }

static int conv_fused_dispatch(bool flat, int32_t layer, const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                  const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const void* w1img,
                  float inv_w1scale, const void* w2img, float inv_wscale, const float* node_in, const int32_t* gather_idx,
                  const float* sh, int32_t sh_stride, const int32_t* seg_ptr, const int32_t* tile_node,
                  const int32_t* n_tiles_dev, int32_t n_tiles_cap, const float* oscale, const float* oshift, float* out,
                  const float* residual, int32_t res_dim, int32_t mode, void* stream) {
    NEED(mode != 1 || residual != nullptr, "dp_conv_fused: mode 1 needs a residual");
    NEED(emb && tb && idxB && tc && idxC && w1img && w2img && node_in && sh && seg_ptr && tile_node && out,
         "dp_conv_fused: null argument");
    NEED(strideB % 2 == 0 && strideC % 2 == 0, "dp_conv_fused: node rows must be 8-byte aligned");
    ConvFusedArgs a;
    a.emb = emb; a.tb = tb; a.idxB = idxB; a.strideB = strideB; a.tc = tc; a.idxC = idxC; a.idxC2 = idxC2; a.strideC = strideC;
    a.w1img = w1img; a.inv_w1scale = inv_w1scale;
    a.w2img = w2img; a.inv_wscale = inv_wscale; a.node_in = node_in; a.gather_idx = gather_idx; a.perm = perm;
    a.sh = sh; a.sh_stride = sh_stride; a.seg_ptr = seg_ptr; a.tile_node = tile_node; a.n_tiles_dev = n_tiles_dev;
    a.n_tiles = n_tiles_cap; a.oscale = oscale; a.oshift = oshift; a.out = out; a.residual = residual; a.res_dim = res_dim;
    a.mode = mode;
    a.dbg = g_cf_dbg_host;
    if (flat && (mode & 16)) {                                            /* flat layout, last chunk's MMA trimmed (experimental) */
        a.dbg = nullptr;
        a.mode = mode & 15;
        switch (layer) {
            case DP_TP_L0: return conv_fused_launch<CfFlatTrim<TpL0>>(a, ST(stream));
            case DP_TP_L1: return conv_fused_launch<CfFlatTrim<TpL1>>(a, ST(stream));
            case DP_TP_L2: return conv_fused_launch<CfFlatTrim<TpL2>>(a, ST(stream));
            case DP_TP_L3: return conv_fused_launch<CfFlatTrim<TpL3>>(a, ST(stream));
            case DP_TP_TOR: return conv_fused_launch<CfFlatTrim<TpTor>>(a, ST(stream));
            case DP_TP_FINAL: return conv_fused_launch<CfFlatTrim<TpFinal>>(a, ST(stream));   /* fc 40 -> 40 -> 200 zero padded to 60 / 60 */
        }
        dp_set_error("dp_conv_fused_flat: unsupported layer %d", layer);
        return DP_ERR_ARG;
    }
    if (flat) {
        a.dbg = nullptr;
        switch (layer) {
            case DP_TP_L0: return conv_fused_launch<CfFlat<TpL0>>(a, ST(stream));
            case DP_TP_L1: return conv_fused_launch<CfFlat<TpL1>>(a, ST(stream));
            case DP_TP_L2: return conv_fused_launch<CfFlat<TpL2>>(a, ST(stream));
            case DP_TP_L3: return conv_fused_launch<CfFlat<TpL3>>(a, ST(stream));
            case DP_TP_TOR: return conv_fused_launch<CfFlat<TpTor>>(a, ST(stream));
        }
        dp_set_error("dp_conv_fused_flat: unsupported layer %d", layer);
        return DP_ERR_ARG;
    }
    switch (layer) {
        case DP_TP_L0: return conv_fused_launch<TpL0>(a, ST(stream));
        case DP_TP_L1: return conv_fused_launch<TpL1>(a, ST(stream));
        case DP_TP_L2: return conv_fused_launch<TpL2>(a, ST(stream));
        case DP_TP_L3: return conv_fused_launch<TpL3>(a, ST(stream));
        case DP_TP_TOR: return conv_fused_launch<TpTor>(a, ST(stream));
    }
    dp_set_error("dp_conv_fused: unsupported layer %d", layer);
    return DP_ERR_ARG;
}

int dp_conv_fused(int32_t layer, const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                  const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const void* w1img,
                  float inv_w1scale, const void* w2img, float inv_wscale, const float* node_in, const int32_t* gather_idx,
                  const float* sh, int32_t sh_stride, const int32_t* seg_ptr, const int32_t* tile_node,
                  const int32_t* n_tiles_dev, int32_t n_tiles_cap, const float* oscale, const float* oshift, float* out,
                  const float* residual, int32_t res_dim, int32_t mode, void* stream) {
    return conv_fused_dispatch(false, layer, emb, perm, tb, idxB, strideB, tc, idxC, idxC2, strideC, w1img, inv_w1scale, w2img,
                               inv_wscale, node_in, gather_idx, sh, sh_stride, seg_ptr, tile_node, n_tiles_dev, n_tiles_cap, oscale,
                               oshift, out, residual, res_dim, mode, stream);
}
/* EXPERIMENTAL: same contract, second-layer weights in the flat 112-column layout (engine._make_w2imgflat) */
int dp_conv_fused_flat(int32_t layer, const float* emb, const int32_t* perm, const float* tb, const int32_t* idxB, int32_t strideB,
                       const float* tc, const int32_t* idxC, const int32_t* idxC2, int32_t strideC, const void* w1img,
                       float inv_w1scale, const void* w2img, float inv_wscale, const float* node_in, const int32_t* gather_idx,
                       const float* sh, int32_t sh_stride, const int32_t* seg_ptr, const int32_t* tile_node,
                       const int32_t* n_tiles_dev, int32_t n_tiles_cap, const float* oscale, const float* oshift, float* out,
                       const float* residual, int32_t res_dim, int32_t mode, void* stream) {
    return conv_fused_dispatch(true, layer, emb, perm, tb, idxB, strideB, tc, idxC, idxC2, strideC, w1img, inv_w1scale, w2img,
                               inv_wscale, node_in, gather_idx, sh, sh_stride, seg_ptr, tile_node, n_tiles_dev, n_tiles_cap, oscale,
                               oshift, out, residual, res_dim, mode, stream);
}

/* profiling aid (not part of the public header): pass 1 of dp_edge_mlp_tc alone */
int dp_debug_edge_hidden(const float* emb, const float* tb, const int32_t* idxB, int32_t strideB, const float* tc,
                         const int32_t* idxC, int32_t strideC, const float* w1, const float* b1, int32_t n_edges,
                         float* h_scratch, void* stream) {
    EdgeMlpArgs a;
    a.emb = emb; a.perm = nullptr; a.tb = tb; a.idxB = idxB; a.strideB = strideB; a.tc = tc; a.idxC = idxC; a.idxC2 = nullptr;
    a.strideC = strideC; a.w1 = w1; a.b1 = b1; a.w2t = nullptr; a.in_dim = 60; a.hid = 60; a.W = 0;
    a.n_edges_dev = nullptr; a.n_edges = n_edges; a.out = nullptr;
    edge_hidden_kernel<<<min((n_edges + EH_TILE - 1) / EH_TILE, 148 * 6), EH_THREADS, 0, ST(stream)>>>(a, h_scratch);
    return dp_check_launch("edge_hidden");
}

/* profiling aid (not part of the public header): route per-phase clock stamps of dp_edge_mlp_tc to a device buffer */
int dp_debug_set_tc_probe(long long* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_tc_dbg, &buf, sizeof(buf));
    return e == cudaSuccess ? DP_OK : DP_ERR_CUDA;
}

int dp_tp_scatter(int32_t layer, const float* node_in, const int32_t* gather_idx, const int32_t* perm, const float* sh,
                  int32_t sh_stride, const float* w, const int32_t* seg_ptr, const float* oscale, const float* oshift,
                  float* out, const float* residual, int32_t res_dim, int32_t mode, int32_t n_out, void* stream) {
    NEED(mode != 1 || residual != nullptr, "dp_tp_scatter: mode 1 needs a residual");
    static int use_tma = -1;                      // DIFFPHORE_TP_SCATTER=reg selects the register-streamed kernel
    if (use_tma < 0) {
        const char* e = getenv("DIFFPHORE_TP_SCATTER");
        use_tma = (e && e[0] == 'r') ? 0 : 1;
    }
#define TP_CASE(ID, CFG, WARPS, STAGES)                                                                               \
    case ID:                                                                                                          \
        if (use_tma)                                                                                                  \
            return tp_scatter_tma_launch<CFG, WARPS, STAGES>(node_in, gather_idx, perm, sh, sh_stride, w, seg_ptr,    \
                                                             oscale, oshift, out, residual, res_dim, mode, n_out,     \
                                                             ST(stream));                                             \
        return tp_scatter_launch<CFG>(node_in, gather_idx, perm, sh, sh_stride, w, seg_ptr, oscale, oshift, out,      \
                                      residual, res_dim, mode, n_out, ST(stream));
    switch (layer) {
        TP_CASE(DP_TP_L0, TpL0, 8, 8)
        TP_CASE(DP_TP_L1, TpL1, 8, 4)
        TP_CASE(DP_TP_L2, TpL2, 8, 3)
        TP_CASE(DP_TP_L3, TpL3, 6, 3)
        TP_CASE(DP_TP_FINAL, TpFinal, 8, 8)
        TP_CASE(DP_TP_TOR, TpTor, 8, 3)
    }
    dp_set_error("dp_tp_scatter: unknown layer %d", layer);
    return DP_ERR_ARG;
}

int dp_center_step(const float* lpos, const int32_t* lig_ptr, int32_t n_graphs, const DpSmallWeights* sw, const float* sc,
                   float* c_emb, float* c_sh, void* stream) {
    if (n_graphs <= 0) return DP_OK;
    center_step_kernel<<<n_graphs, 128, 0, ST(stream)>>>(lpos, lig_ptr, *sw, sc, c_emb, c_sh);
    return dp_check_launch("dp_center_step");
}

int dp_score_head(const float* gpred, int32_t n_graphs, const DpSmallWeights* sw, const float* sc, float* tr, float* rot,
                  void* stream) {
    if (n_graphs <= 0) return DP_OK;
    score_head_kernel<<<(n_graphs + 127) / 128, 128, 0, ST(stream)>>>(gpred, n_graphs, *sw, sc, tr, rot);
    return dp_check_launch("dp_score_head");
}

int dp_tor_graph(const float* lpos, const int32_t* lig_ptr, const int32_t* rot_ptr, const int32_t* rot_u,
                 const int32_t* rot_v, int32_t n_graphs, int32_t n_rot, const DpSmallWeights* sw, int32_t* deg,
                 int32_t* gcount, int32_t* gstart, int32_t* seg_ptr, int32_t* e_atom, int32_t* e_u, int32_t* e_v,
                 float* e_emb, float* e_sh, int32_t* n_edges_out, void* stream) {
    if (n_graphs <= 0 || n_rot <= 0) return DP_OK;
    tor_count_kernel<<<n_graphs, 128, 0, ST(stream)>>>(lpos, lig_ptr, rot_ptr, rot_u, rot_v, deg, gcount);
    scan_kernel<<<1, 1024, 0, ST(stream)>>>(gcount, gstart, n_graphs, n_edges_out);
    tor_fill_kernel<<<dim3(n_graphs, dp_graph_split(n_graphs)), 128, 0, ST(stream)>>>(lpos, lig_ptr, rot_ptr, rot_u, rot_v, deg, gstart, n_graphs, *sw,
                                                     seg_ptr, e_atom, e_u, e_v, e_emb, e_sh);
    return dp_check_launch("dp_tor_graph");
}

int dp_tor_head(const float* tor_feat, int32_t n_rot, const DpSmallWeights* sw, const float* sc, float* tor, void* stream) {
    if (n_rot <= 0) return DP_OK;
    tor_head_kernel<<<(n_rot + 127) / 128, 128, 0, ST(stream)>>>(tor_feat, n_rot, *sw, sc, tor);
    return dp_check_launch("dp_tor_head");
}

int dp_conformer_update(float* pos, float* norm, const int32_t* lig_ptr, const int32_t* rot_ptr, const int32_t* rot_u,
                        const int32_t* rot_v, const uint8_t* mask, const int64_t* mask_off, int32_t n_graphs,
                        int32_t max_atoms, int32_t max_rot, const float* tr_score, const float* rot_score,
                        const float* tor_score, const float* tr_z, const float* rot_z, const float* tor_z,
                        const float* sc, int32_t no_torsion, void* stream) {
    if (n_graphs <= 0) return DP_OK;
    const size_t smem = (size_t)(39 * max_atoms + max_rot) * sizeof(float) + 2 * (size_t)max_atoms + 16;
    NEED(smem <= 200 * 1024, "dp_conformer_update: ligand too large for shared memory");
    if (smem > 48 * 1024) cudaFuncSetAttribute(conformer_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conformer_update_kernel<<<n_graphs, CU_THREADS, smem, ST(stream)>>>(
        pos, norm, lig_ptr, rot_ptr, rot_u, rot_v, mask, (const long long*)mask_off, tr_score, rot_score, tor_score, tr_z,
        rot_z, tor_z, sc, no_torsion);
    return dp_check_launch("dp_conformer_update");
}

int dp_randomize_position(float* pos, float* norm, const int32_t* lig_ptr, const int32_t* rot_ptr, const int32_t* rot_u,
                          const int32_t* rot_v, const uint8_t* mask, const int64_t* mask_off, int32_t n_graphs,
                          int32_t max_atoms, int32_t max_rot, const float* tor_init, const float* rot_init,
                          const float* tr_init, int32_t no_torsion, void* stream) {
    if (n_graphs <= 0) return DP_OK;
    NEED(rot_init != nullptr, "dp_randomize_position: rot_init missing");
    const size_t smem = (size_t)(36 * max_atoms + max_rot) * sizeof(float) + 2 * (size_t)max_atoms + 16;
    NEED(smem <= 200 * 1024, "dp_randomize_position: ligand too large for shared memory");
    if (smem > 48 * 1024) cudaFuncSetAttribute(randomize_position_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    randomize_position_kernel<<<n_graphs, CU_THREADS, smem, ST(stream)>>>(pos, norm, lig_ptr, rot_ptr, rot_u, rot_v, mask,
                                                                         (const long long*)mask_off, tor_init, rot_init,
                                                                         tr_init, no_torsion);
    return dp_check_launch("dp_randomize_position");
}

}  // extern "C"
