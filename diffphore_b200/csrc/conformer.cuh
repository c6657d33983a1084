// dp_conformer_update: the per-sample rigid-body + torsion update driven by the predicted scores.
//
// Replaces, per sample and per step (one CTA per sample):
//   perturbation assembly                     src/utils/sampling.py:223-246   (g^2 dt score + g sqrt(dt) z)
//   modify_conformer                          src/utils/diffusion_utils.py:23-79
//   axis_angle_to_matrix (via quaternion)     src/utils/geometry.py:38-85
//   modify_conformer_torsion_angles           src/utils/torsion.py:64-109     (sequential over rotatable bonds)
//   rigid_transform_Kabsch_3D_torch           src/utils/geometry.py:88-136    (3x3 SVD -> Jacobi, reflection-safe)
// and dp_randomize_position: src/utils/sampling.py:16-63 with injected draws.
//
// Bug-for-bug items: the 11 "norm points" per atom are addressed through the reference's raw reshape
// norm.reshape(-1, N, 3): flat vector slot f = a*11 + t rides on atom (f mod N).  Rotation matrices
// of the torsion moves are built and applied in fp64 and rounded to fp32 like scipy/numpy do there (H9).
#pragma once
#include "common.cuh"

#define CU_THREADS 128

__device__ __forceinline__ void cu_rodrigues(double ax, double ay, double az, double* R) {
    // scipy Rotation.from_rotvec(v).as_matrix()
    const double th = sqrt(ax * ax + ay * ay + az * az);
    double s, c1;                      // s = sin(th)/th, c1 = (1-cos(th))/th^2
    if (th < 1e-3) {
        const double t2 = th * th;
        s = 1.0 - t2 / 6.0 + t2 * t2 / 120.0;
        c1 = 0.5 - t2 / 24.0 + t2 * t2 / 720.0;
    } else {
        s = sin(th) / th;
        c1 = (1.0 - cos(th)) / (th * th);
    }
    R[0] = 1.0 - c1 * (ay * ay + az * az); R[1] = c1 * ax * ay - s * az;        R[2] = c1 * ax * az + s * ay;
    R[3] = c1 * ax * ay + s * az;        R[4] = 1.0 - c1 * (ax * ax + az * az); R[5] = c1 * ay * az - s * ax;
    R[6] = c1 * ax * az - s * ay;        R[7] = c1 * ay * az + s * ax;        R[8] = 1.0 - c1 * (ax * ax + ay * ay);
}

__device__ __forceinline__ void cu_axis_angle_to_matrix(float ax, float ay, float az, float* R) {
    // geometry.py:38-85 (fp32, small-angle branch below 1e-6)
    const float ang = sqrtf(ax * ax + ay * ay + az * az);
    const float half = 0.5f * ang;
    const float sh = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
    const float r = cosf(half), i = ax * sh, j = ay * sh, k = az * sh;
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);     R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);     R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);     R[7] = two_s * (j * k + i * r);     R[8] = 1 - two_s * (i * i + j * j);
}

// Jacobi eigen-decomposition of a symmetric 3x3 (double); V columns = eigenvectors
__device__ void cu_jacobi3(double* A, double* V) {
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        // converged once the off-diagonal part is below 1e-18 of the diagonal: further rotations have |s| < 1e-18 and no longer
        // change a double (the old test, off < 1e-300, ran ~10 sweeps where 5-6 do all the work; this routine is on the
        // critical path of a CTA, on one thread)
        const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (off <= 1e-36 * (A[0] * A[0] + A[4] * A[4] + A[8] * A[8])) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[p * 3 + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
}

// Kabsch rotation for H = Am Bm^T (geometry.py:121-131): R = V diag(1,1,det) U^T with H = U S V^T, built from
// right-handed triads so that the reflection case is handled implicitly.
__device__ void cu_kabsch_rotation(const double* H, double* R) {
    double M[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = H[0 * 3 + i] * H[0 * 3 + j] + H[1 * 3 + i] * H[1 * 3 + j] + H[2 * 3 + i] * H[2 * 3 + j];
    cu_jacobi3(M, V);
    int i0 = 0, i1 = 1, i2 = 2;
    double e0 = M[0], e1 = M[4], e2 = M[8];
    if (e0 < e1) { double t = e0; e0 = e1; e1 = t; int ti = i0; i0 = i1; i1 = ti; }
    if (e0 < e2) { double t = e0; e0 = e2; e2 = t; int ti = i0; i0 = i2; i2 = ti; }
    if (e1 < e2) { double t = e1; e1 = e2; e2 = t; int ti = i1; i1 = i2; i2 = ti; }
    double v0[3] = {V[0 * 3 + i0], V[1 * 3 + i0], V[2 * 3 + i0]}, v1[3] = {V[0 * 3 + i1], V[1 * 3 + i1], V[2 * 3 + i1]};
    double u0[3], u1[3];
    for (int i = 0; i < 3; ++i) {
        u0[i] = H[i * 3] * v0[0] + H[i * 3 + 1] * v0[1] + H[i * 3 + 2] * v0[2];
        u1[i] = H[i * 3] * v1[0] + H[i * 3 + 1] * v1[1] + H[i * 3 + 2] * v1[2];
    }
    double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
    for (int i = 0; i < 3; ++i) u0[i] /= n0;
    double d = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
    for (int i = 0; i < 3; ++i) u1[i] -= d * u0[i];
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = v0[i] * u0[j] + v1[i] * u1[j] + v2[i] * u2[j];
}

// Sequential torsion moves on the 12N points held in shared memory (pts[0..N) = atoms, pts[N..12N) = norm points,
// norm point f rides on atom f % N).  theta[r] == 0 skips the bond (torsion.py:83-84).
// One barrier per bond: every thread builds the bond's rotation matrix itself (same instructions in every lane; the bond's end
// points u and v are not moved by their own bond - u is on the fixed side, v is the pivot and maps onto itself exactly), the
// mask row of the NEXT bond is fetched from global memory under the current bond's arithmetic and handed over through
// shared memory (smask: 2 x n bytes), and the atom a point rides on is tracked incrementally instead of f % n per point.
__device__ void cu_apply_torsions(float* pts, int n, int nr, const int* __restrict__ rot_u, const int* __restrict__ rot_v,
                                  const unsigned char* __restrict__ mask, const float* theta, int a0, unsigned char* smask) {
    const int step = (int)blockDim.x % n, a_first = (int)threadIdx.x % n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) smask[i] = mask[i];
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
        unsigned char* cur = smask + (r & 1) * n;
        unsigned char* nxt = smask + ((r + 1) & 1) * n;
        // (a thread owns at most a few atoms' worth of mask bytes: n <= blockDim.x in practice, the loop covers the rest)
        unsigned char mk = 0;
        const bool pf = r + 1 < nr && (int)threadIdx.x < n;
        if (pf) mk = mask[(size_t)(r + 1) * n + threadIdx.x];
        const float th = theta[r];
        if (th != 0.0f) {                          // uniform across the CTA
            const int u = rot_u[r] - a0, v = rot_v[r] - a0;
            // rot_vec = (pos[u]-pos[v]) * theta / |pos[u]-pos[v]| in fp32 (numpy), then scipy in fp64
            const float pvx = pts[v * 3], pvy = pts[v * 3 + 1], pvz = pts[v * 3 + 2];
            const float dx = pts[u * 3] - pvx, dy = pts[u * 3 + 1] - pvy, dz = pts[u * 3 + 2] - pvz;
            const float nn = sqrtf(dx * dx + dy * dy + dz * dz);
            double Rt[9];
            cu_rodrigues((double)(dx * th / nn), (double)(dy * th / nn), (double)(dz * th / nn), Rt);
            int a = a_first;
            for (int f = threadIdx.x; f < 12 * n; f += blockDim.x) {
                if (cur[a]) {
                    const double x = (double)(pts[f * 3] - pvx), y = (double)(pts[f * 3 + 1] - pvy), z = (double)(pts[f * 3 + 2] - pvz);
                    pts[f * 3] = (float)(x * Rt[0] + y * Rt[1] + z * Rt[2] + (double)pvx);
                    pts[f * 3 + 1] = (float)(x * Rt[3] + y * Rt[4] + z * Rt[5] + (double)pvy);
                    pts[f * 3 + 2] = (float)(x * Rt[6] + y * Rt[7] + z * Rt[8] + (double)pvz);
                }
                a += step;
                if (a >= n) a -= n;
            }
        }
        if (pf) nxt[threadIdx.x] = mk;
        for (int i = threadIdx.x + blockDim.x; i < n && r + 1 < nr; i += blockDim.x) nxt[i] = mask[(size_t)(r + 1) * n + i];
        __syncthreads();
    }
}

__device__ __forceinline__ void cu_load_points(float* pts, const float* __restrict__ pos, const float* __restrict__ norm,
                                               int a0, int n) {
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) pts[i] = pos[(size_t)a0 * 3 + i];
    __syncthreads();
    const int step = (int)blockDim.x % n;
    int a = (int)threadIdx.x % n;
    for (int f = threadIdx.x; f < 11 * n; f += blockDim.x) {          // lig_norm = norm.reshape(-1,N,3) + pos[None]
#pragma unroll
        for (int k = 0; k < 3; ++k) pts[(n + f) * 3 + k] = norm[(size_t)a0 * 33 + (size_t)f * 3 + k] + pts[a * 3 + k];
        a += step;
        if (a >= n) a -= n;
    }
    __syncthreads();
}
// norm = points - the atom they ride on, back to global memory
__device__ __forceinline__ void cu_store_norm(const float* pts, float* __restrict__ norm, int a0, int n) {
    const int step = (int)blockDim.x % n;
    int a = (int)threadIdx.x % n;
    for (int f = threadIdx.x; f < 11 * n; f += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) norm[(size_t)a0 * 33 + (size_t)f * 3 + k] = pts[(n + f) * 3 + k] - pts[a * 3 + k];
        a += step;
        if (a >= n) a -= n;
    }
}

// the centroids of two point sets at once (threads 0-2 and 32-34: two warps, one barrier)
__device__ __forceinline__ void cu_mean3x2(const float* pa, const float* pb, int n, float* oa, float* ob) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w < 2 && l < 3) {
        const float* p = w ? pb : pa;
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += p[i * 3 + l];
        (w ? ob : oa)[l] = s / (float)n;
    }
    __syncthreads();
}
__device__ __forceinline__ void cu_mean3(const float* pts, int n, float* out /*shared[3]*/) {
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += pts[i * 3 + threadIdx.x];
        out[threadIdx.x] = s / (float)n;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(CU_THREADS)
conformer_update_kernel(float* __restrict__ pos, float* __restrict__ norm, const int* __restrict__ lig_ptr,
                        const int* __restrict__ rot_ptr, const int* __restrict__ rot_u, const int* __restrict__ rot_v,
                        const unsigned char* __restrict__ mask, const long long* __restrict__ mask_off,
                        const float* __restrict__ tr_score, const float* __restrict__ rot_score,
                        const float* __restrict__ tor_score, const float* __restrict__ tr_z, const float* __restrict__ rot_z,
                        const float* __restrict__ tor_z, const float* __restrict__ sc, int no_torsion) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0, r0 = rot_ptr[g], nr = rot_ptr[g + 1] - r0;
    float* pts = smem;                       // [12n][3] current (flexible) points
    float* rigid = pts + 36 * n;             // [n][3] rigid-body positions (Kabsch target)
    float* theta = rigid + 3 * n;            // [nr]
    unsigned char* smask = reinterpret_cast<unsigned char*>(theta + nr);      // [2][n] mask rows of the bond in flight / the next one
    __shared__ float cen[3], cA[3], cB[3], Rm[9], tr[3], Kr[9], Kt[3];
    __shared__ double Hs[9];
    cu_load_points(pts, pos, norm, a0, n);
    cu_mean3(pts, n, cen);
    if (threadIdx.x == 0) {
        float ru[3];
        for (int k = 0; k < 3; ++k) {
            tr[k] = sc[SC_TR_A] * tr_score[g * 3 + k] + (tr_z ? sc[SC_TR_B] * tr_z[g * 3 + k] : 0.f);
            ru[k] = rot_score[g * 3 + k] * sc[SC_ROT_A] + (rot_z ? sc[SC_ROT_B] * rot_z[g * 3 + k] : 0.f);
        }
        cu_axis_angle_to_matrix(ru[0], ru[1], ru[2], Rm);
    }
    for (int r = threadIdx.x; r < nr; r += blockDim.x)
        theta[r] = sc[SC_TOR_A] * tor_score[r0 + r] + (tor_z ? sc[SC_TOR_B] * tor_z[r0 + r] : 0.f);
    __syncthreads();
    // rigid move of atoms and norm points: (p - c) @ R^T + tr + c
    for (int f = threadIdx.x; f < 12 * n; f += blockDim.x) {
        const float x = pts[f * 3] - cen[0], y = pts[f * 3 + 1] - cen[1], z = pts[f * 3 + 2] - cen[2];
        const float nx = x * Rm[0] + y * Rm[1] + z * Rm[2] + tr[0] + cen[0];
        const float ny = x * Rm[3] + y * Rm[4] + z * Rm[5] + tr[1] + cen[1];
        const float nz = x * Rm[6] + y * Rm[7] + z * Rm[8] + tr[2] + cen[2];
        pts[f * 3] = nx; pts[f * 3 + 1] = ny; pts[f * 3 + 2] = nz;
        if (f < n) { rigid[f * 3] = nx; rigid[f * 3 + 1] = ny; rigid[f * 3 + 2] = nz; }
    }
    __syncthreads();
    if (!no_torsion && nr > 0) {
        cu_apply_torsions(pts, n, nr, rot_u + r0, rot_v + r0, mask + mask_off[g], theta, a0, smask);
        // Kabsch: align flexible (A) onto rigid (B)
        cu_mean3x2(pts, rigid, n, cA, cB);
        if (threadIdx.x < 9) {
            const int i = threadIdx.x / 3, j = threadIdx.x % 3;
            float s = 0.f;                                      // H = Am @ Bm^T (fp32 like torch)
            for (int k = 0; k < n; ++k) s = fmaf(pts[k * 3 + i] - cA[i], rigid[k * 3 + j] - cB[j], s);
            Hs[threadIdx.x] = (double)s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double R[9];
            cu_kabsch_rotation(Hs, R);
            for (int i = 0; i < 9; ++i) Kr[i] = (float)R[i];
            for (int i = 0; i < 3; ++i) Kt[i] = -(Kr[i * 3] * cA[0] + Kr[i * 3 + 1] * cA[1] + Kr[i * 3 + 2] * cA[2]) + cB[i];
        }
        __syncthreads();
        for (int f = threadIdx.x; f < 12 * n; f += blockDim.x) {
            const float x = pts[f * 3], y = pts[f * 3 + 1], z = pts[f * 3 + 2];
            pts[f * 3] = x * Kr[0] + y * Kr[1] + z * Kr[2] + Kt[0];
            pts[f * 3 + 1] = x * Kr[3] + y * Kr[4] + z * Kr[5] + Kt[1];
            pts[f * 3 + 2] = x * Kr[6] + y * Kr[7] + z * Kr[8] + Kt[2];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) pos[(size_t)a0 * 3 + i] = pts[i];
    cu_store_norm(pts, norm, a0, n);
}

// randomize_position (sampling.py:16-63): random torsions on the input pose, then centre, rotate, translate.
__global__ void __launch_bounds__(CU_THREADS)
randomize_position_kernel(float* __restrict__ pos, float* __restrict__ norm, const int* __restrict__ lig_ptr,
                          const int* __restrict__ rot_ptr, const int* __restrict__ rot_u, const int* __restrict__ rot_v,
                          const unsigned char* __restrict__ mask, const long long* __restrict__ mask_off,
                          const float* __restrict__ tor_init, const float* __restrict__ rot_init,
                          const float* __restrict__ tr_init, int no_torsion) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x, a0 = lig_ptr[g], n = lig_ptr[g + 1] - a0, r0 = rot_ptr[g], nr = rot_ptr[g + 1] - r0;
    float* pts = smem;
    float* theta = pts + 36 * n;
    unsigned char* smask = reinterpret_cast<unsigned char*>(theta + nr);
    __shared__ float cen[3];
    cu_load_points(pts, pos, norm, a0, n);
    for (int r = threadIdx.x; r < nr; r += blockDim.x) theta[r] = tor_init ? tor_init[r0 + r] : 0.f;
    __syncthreads();
    if (!no_torsion && nr > 0) cu_apply_torsions(pts, n, nr, rot_u + r0, rot_v + r0, mask + mask_off[g], theta, a0, smask);
    // the reference stores norm = points (absolute!) after the torsion pass and only later subtracts: replicate:
    //   pos, norm(abs points) = modify_conformer_torsion_angles(...);  pos' = (pos - c) @ R^T
    //   norm' = (norm_abs.reshape(N,33)... - c) @ R^T - pos'     (sampling.py:50-54)
    cu_mean3(pts, n, cen);
    const float* R = rot_init + (size_t)g * 9;
    for (int f = threadIdx.x; f < 12 * n; f += blockDim.x) {
        const float x = pts[f * 3] - cen[0], y = pts[f * 3 + 1] - cen[1], z = pts[f * 3 + 2] - cen[2];
        pts[f * 3] = x * R[0] + y * R[1] + z * R[2];
        pts[f * 3 + 1] = x * R[3] + y * R[4] + z * R[5];
        pts[f * 3 + 2] = x * R[6] + y * R[7] + z * R[8];
    }
    __syncthreads();
    cu_store_norm(pts, norm, a0, n);
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x)
        pos[(size_t)a0 * 3 + i] = pts[i] + (tr_init ? tr_init[g * 3 + (i % 3)] : 0.f);
}
