// Shared helpers for the DiffPhore denoising kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/diffphore_b200.h"

#define DP_OK 0
#define DP_ERR_ARG 1
#define DP_ERR_CUDA 2

void dp_set_error(const char* fmt, ...);
int dp_check_launch(const char* what);

#define DP_NS 20            // scalar channels (ns)
#define DP_NV 10            // vector channels (nv)
#define DP_RBF 20           // radial basis size
#define DP_SH 9             // lmax=2 spherical harmonics

// e3nn real spherical harmonics, lmax = 2, 'component' normalisation, of normalize(v) (F.normalize eps 1e-12).
// Order: Y0 | Y1 (x,y,z) | Y2 (xz, xy, y^2-(x^2+z^2)/2, yz, (z^2-x^2)/2)   [e3nn 0.5.1, smp:737]
__device__ __forceinline__ void dp_sh9(float vx, float vy, float vz, float* sh) {
    float n = sqrtf(vx * vx + vy * vy + vz * vz);
    float inv = 1.0f / fmaxf(n, 1e-12f);
    float x = vx * inv, y = vy * inv, z = vz * inv;
    const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f, s15 = 3.872983346207417f;
    sh[0] = 1.0f;
    sh[1] = s3 * x; sh[2] = s3 * y; sh[3] = s3 * z;
    sh[4] = s15 * x * z;
    sh[5] = s15 * x * y;
    sh[6] = s5 * (y * y - 0.5f * (x * x + z * z));
    sh[7] = s15 * y * z;
    sh[8] = (s15 * 0.5f) * (z * z - x * x);
}

// GaussianSmearing (smp:978-1015): exp(coeff * (d - mu_k)^2).  The four expansions' offsets (the checkpoint's
// `*_distance_expansion.offset` buffers) and coefficients live in constant memory (dp_set_constants).
#define DP_RBF_LIG 0
#define DP_RBF_PHORE 1
#define DP_RBF_CROSS 2
#define DP_RBF_CENTER 3
static __constant__ DpConstants c_dp;   // used by dp_abi.cu only (static: conv_fused2.cu includes this header too)

__device__ __forceinline__ void dp_rbf20(float d, int which, float* out) {
    const float coeff = c_dp.rbf_coeff[which];
#pragma unroll
    for (int k = 0; k < DP_RBF; ++k) {
        float t = d - c_dp.rbf_mu[which][k];
        out[k] = expf(coeff * t * t);
    }
}

__device__ __forceinline__ float dp_softplus(float x) {     // torch.nn.Softplus(beta=1, threshold=20)
    return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float dp_leaky(float x) { return x > 0.0f ? x : 0.01f * x; }
