// vio_kernels.cuh — hand-written sm_100a kernels of the LM hot path.
// Reference functions restated (A15 = 15-vio-backend, A17 = 17-vins-initialization/vins-mono under
// /root/reference/workspace/assignments):
//   k_pose_prep            VertexPose params -> R,t                (A15/backend/edge_reprojection.cc:23-29,68-70)
//   k_linearize_lm         EdgeReprojection::ComputeResidual/ComputeJacobians + MakeHessian + Schur
//                          (A15/backend/edge_reprojection.cc:20-111, A15/backend/problem.cc:280-337,353-399;
//                           A17/src/backend/edge.cc:50-74, A17/src/backend/problem.cc:303-389,406-437)
//   k_se3prior             EdgeSE3Prior                            (A15/backend/edge_prior.cpp:39-80)
//   k_chi2_lm              Σ Chi2 / RobustChi2                     (A15/backend/problem.cc:457-462,501-507)
//   k_backsub              landmark back-substitution              (A15/backend/problem.cc:407-421)
//   k_update_*             UpdateStates / RollbackStates           (A15/backend/problem.cc:425-450, A15/backend/vertex_pose.cc:7-16)
#pragma once
#include "vio_dev.h"
#include "vio_math.cuh"

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// FP64 accumulate into global memory: RED.E.ADD.F64 on the device.  The host branch exists only so
// tests/host_emul.cu can run the same per-landmark bodies on the CPU as a debugging/unit-test aid.
// Warp aggregation: in the per-landmark kernels consecutive landmarks usually share their host and observers, so all 32
// lanes of a warp tend to add to the SAME element; when they do, one shuffle reduction + one RED replaces 32 REDs.
VIO_HD void vio_add(double *p, double x) {
#ifdef __CUDA_ARCH__
    const unsigned active = __activemask();
    if (active == 0xffffffffu) {
        const unsigned long long a0 = __shfl_sync(0xffffffffu, (unsigned long long)p, 0);
        if (__all_sync(0xffffffffu, a0 == (unsigned long long)p)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(p, x);
            return;
        }
    }
    atomicAdd(p, x);
#else
    *p += x;
#endif
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- warp reduce-scatter by recursive halving (47 shuffles for a 48-vector instead of 240 for 48 butterflies) ----
template <int N>
__device__ __forceinline__ void rs_step(double *v, bool up, int mask) {
    constexpr int H = N / 2;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double send = up ? v[i] : v[i + H];
        const double keep = up ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}
// warp sum of a 48-vector, scattered: afterwards lane holds element `base` in v[0] and (if !(lane&1)) base+1 in v[1]
__device__ __forceinline__ int reduce_scatter48(double *v, int lane) {
    int base = 0;
    rs_step<48>(v, lane & 16, 16); base += (lane & 16) ? 24 : 0;
    rs_step<24>(v, lane & 8, 8);   base += (lane & 8) ? 12 : 0;
    rs_step<12>(v, lane & 4, 4);   base += (lane & 4) ? 6 : 0;
    rs_step<6>(v, lane & 2, 2);    base += (lane & 2) ? 3 : 0;
    v[3] = 0.0;
    rs_step<4>(v, lane & 1, 1);    base += (lane & 1) ? 2 : 0;
    return base;
}
__device__ __forceinline__ int reduce_scatter32(double *v, int lane) {
    int base = 0;
    rs_step<32>(v, lane & 16, 16); base += (lane & 16) ? 16 : 0;
    rs_step<16>(v, lane & 8, 8);   base += (lane & 8) ? 8 : 0;
    rs_step<8>(v, lane & 4, 4);    base += (lane & 4) ? 4 : 0;
    rs_step<4>(v, lane & 2, 2);    base += (lane & 2) ? 2 : 0;
    rs_step<2>(v, lane & 1, 1);    base += (lane & 1) ? 1 : 0;
    return base;  // v[0] holds element `base`
}


// block-level deterministic sum -> partial[blockIdx.x]   (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ void block_sum_to(double v, double *partial) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? sm[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) partial[blockIdx.x] = t;
    }
    __syncthreads();
}

// out[idx] (+)= Σ partial[0..n)  in fixed order (single block)
__global__ void k_sum_partials(const double *partial, int n, double *out, int accumulate) {
    __shared__ double sm[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t += partial[i];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = accumulate ? (*out + sm[0]) : sm[0];
}

// four fixed-order sums in one launch (blockIdx.x selects the array): the LM scalars of a trial step
struct Sum4 {
    const double *partial[4];
    double *out[4];
    int n[4];  // 0: skip
};
__global__ void k_sum_partials4(Sum4 a) {
    __shared__ double sm[256];
    const int w = blockIdx.x;
    if (a.n[w] <= 0) return;
    double t = 0.0;
    for (int i = threadIdx.x; i < a.n[w]; i += blockDim.x) t += a.partial[w][i];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.out[w] = sm[0];
}

__global__ void k_max_partials(const double *partial, int n, double *out) {
    __shared__ double sm[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t = fmax(t, partial[i]);
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

// ------------------------------------------------------------------------------------------------
// pose preparation: q -> R once per state (instead of 2x per edge in the reference)
// ------------------------------------------------------------------------------------------------
VIO_HD void pose_prep(const DevView &v, int i) {
    const double *p = v.pose + 7 * (size_t)i;
    double q[4] = {p[3], p[4], p[5], p[6]};
    double R[9];
    quat_to_R(q, R);
    double *o = v.poseRT + 16 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = R[k];
    o[9] = p[0];
    o[10] = p[1];
    o[11] = p[2];
    // R^T t is not needed; pad
    o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 0;
}

__global__ void k_pose_prep(DevView v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.C) pose_prep(v, i);
}

// ------------------------------------------------------------------------------------------------
// reduced-system addressing
// ------------------------------------------------------------------------------------------------
// Pointer to element (0,0) of the 6x6 block (pose a, pose b) with ordering offset(a) <= offset(b).
VIO_HD double *s_block(const DevView &v, int a, int b, int &ld) {
    if (v.storage == 1) {
        ld = v.Pper;
        return v.S + (size_t)v.pose_off[a] * v.Pper + (v.pose_off[b] % v.Pper);
    }
    ld = 6;
    const int ra = v.pose_blk[a], cb = v.pose_blk[b];
    int lo = v.bsr_rowptr[ra], hi = v.bsr_rowptr[ra + 1] - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (v.bsr_col[mid] < cb) lo = mid + 1; else hi = mid;
    }
    return v.S + 36 * (size_t)lo;
}

#ifdef __CUDA_ARCH__
// Warp-uniform 6x6 block add: when all 32 lanes add a block to the SAME place (consecutive landmarks of the per-landmark
// kernels share their observers), the 36 values are summed over the warp by one reduce-scatter (47 shuffles) and every
// lane issues at most two REDs, instead of 36 shuffle-reduced adds (180 shuffles) or 36 x 32 plain REDs.
// Y[e] is the value for element e = 6 r + c of the block at p (leading dimension ld); upper_only skips r > c.
__device__ __forceinline__ bool warp_block_add(double *p, int ld, const double Y[36], bool upper_only) {
    if (__activemask() != 0xffffffffu) return false;
    const unsigned long long p0 = __shfl_sync(0xffffffffu, (unsigned long long)p, 0);
    const int u0 = __shfl_sync(0xffffffffu, (int)upper_only, 0);
    if (!__all_sync(0xffffffffu, p0 == (unsigned long long)p && u0 == (int)upper_only)) return false;
    double v[48];
#pragma unroll
    for (int e = 0; e < 36; ++e) v[e] = Y[e];
#pragma unroll
    for (int e = 36; e < 48; ++e) v[e] = 0.0;
    const int lane = threadIdx.x & 31;
    const int base = reduce_scatter48(v, lane);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = base + q;
        if (q == 1 && (lane & 1)) break;
        if (e < 36) {
            const int r = e / 6, c = e % 6;
            if (!upper_only || c >= r) atomicAdd(p + (size_t)r * ld + c, v[q]);
        }
    }
    return true;
}
#endif

// S(block a,b) += sgn * X (6x6, row-major X) where X is the (a,b) block; handles orientation so that
// only the upper block-triangle (and upper element-triangle of diagonal blocks) is touched.
VIO_HD void s_add_block(const DevView &v, int a, int b, const double X[36], double sgn) {
    int ld;
    double Y[36];
    double *p;
    bool upper = false;
    if (a == b) {
        // X + X^T contribution when both ends of an edge hit the same vertex
        p = s_block(v, a, a, ld);
        upper = true;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) Y[6 * r + c] = sgn * (X[6 * r + c] + X[6 * c + r]);
    } else if (v.pose_off[a] < v.pose_off[b]) {
        p = s_block(v, a, b, ld);
#pragma unroll
        for (int e = 0; e < 36; ++e) Y[e] = sgn * X[e];
    } else {
        p = s_block(v, b, a, ld);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) Y[6 * r + c] = sgn * X[6 * c + r];
    }
#ifdef __CUDA_ARCH__
    if (warp_block_add(p, ld, Y, upper)) return;
#endif
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = upper ? r : 0; c < 6; ++c) vio_add(p + (size_t)r * ld + c, Y[6 * r + c]);
}
// diagonal block (a,a) += sgn * X, X symmetric: only the upper element-triangle is stored
VIO_HD void s_add_diag(const DevView &v, int a, const double X[36], double sgn) {
    int ld;
    double *p = s_block(v, a, a, ld);
    double Y[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) Y[e] = sgn * X[e];
#ifdef __CUDA_ARCH__
    if (warp_block_add(p, ld, Y, true)) return;
#endif
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) vio_add(p + (size_t)r * ld + c, Y[6 * r + c]);
}

// ------------------------------------------------------------------------------------------------
// per-edge reprojection chain (shared by linearise and chi2)
// ------------------------------------------------------------------------------------------------
struct EdgeLin {
    double r[2];
    double B[6];  // 2x3: reduce * Ric^T * Rj^T ; J_lambda = B g, J_i = B [I G], J_j = B [-I N]
    double N[9];  // Rj * hat(p_bj)
};

VIO_HD void reproj_residual(const double Ric[9], const double tic[3], const double *RTj,
                                                const double pw[3], double pjx, double pjy, double pcj[3],
                                                double pbj[3], double r[2]) {
    const double d[3] = {pw[0] - RTj[9], pw[1] - RTj[10], pw[2] - RTj[11]};
    mat3t_mul_vec(RTj, d, pbj);
    const double e[3] = {pbj[0] - tic[0], pbj[1] - tic[1], pbj[2] - tic[2]};
    mat3t_mul_vec(Ric, e, pcj);
    const double iz = vio_rcp(pcj[2]);  // one reciprocal (1 ulp) instead of the reference's two divisions (edge_reprojection.cc:38)
    r[0] = pcj[0] * iz - pjx;
    r[1] = pcj[1] * iz - pjy;
}

// robust weights (A17/src/backend/edge.cc:39-74) for information = c*I2:
//   e2 = c r.r ; rho = loss(e2); W = c*(rho1 I + [rho1+2rho2 e2>0] 2 rho2 c r r^T); b uses drho*c
VIO_HD void robust_weights(int loss, double delta, double c, const double r[2], double &rho0,
                                               double &drho, double W[3]) {
    const double e2 = c * (r[0] * r[0] + r[1] * r[1]);
    if (loss == 0) {
        rho0 = e2; drho = 1.0;
        W[0] = c; W[1] = 0.0; W[2] = c;
        return;
    }
    double rho[3];
    loss_compute(loss, delta, e2, rho);
    rho0 = rho[0];
    drho = rho[1];
    double a = rho[1], k = 0.0;
    if (rho[1] + 2.0 * rho[2] * e2 > 0.0) k = 2.0 * rho[2] * c;  // weight_err = sqrt(c) r
    W[0] = c * (a + k * r[0] * r[0]);
    W[1] = c * (k * r[0] * r[1]);
    W[2] = c * (a + k * r[1] * r[1]);
}

// ------------------------------------------------------------------------------------------------
// v0 linearise + accumulate + Schur: one thread per landmark, global FP64 atomics into the
// upper block-triangle of the reduced system.  (The grouped shared-memory version replaces this
// for large scenes; this generic kernel stays as the fallback for irregular graphs.)
// ------------------------------------------------------------------------------------------------
template <bool WITH_SCHUR>
VIO_HD void linearize_landmark(const DevView &v, int l) {
    const int h = v.lm_host[l];
    const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
    if (e0 == e1) {  // landmark without edges: Hmm block is 0 (reference would divide by zero)
        v.Hll[l] = 0.0; v.bl[l] = 0.0;
        for (int k = 0; k < 6; ++k) v.wh[6 * (size_t)l + k] = 0.0;
        return;
    }
    const double lam = v.invdep[l];
    const double pts_i[3] = {v.lm_pix[l], v.lm_piy[l], v.lm_piz[l]};
    const double *RTh = v.poseRT + 16 * (size_t)h;
    const bool hfix = v.pose_fixed[h] != 0;

    // host chain: p_ci = pts_i / lambda ; p_bi = Ric p_ci + tic ; p_w = Ri p_bi + Pi
    const double pci[3] = {pts_i[0] / lam, pts_i[1] / lam, pts_i[2] / lam};
    double pbi[3], pw[3], tmp[3];
    mat3_mul_vec(v.Ric, pci, pbi);
    pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
    mat3_mul_vec(RTh, pbi, pw);
    pw[0] += RTh[9]; pw[1] += RTh[10]; pw[2] += RTh[11];
    // g = Ri Ric pts_i * (-1/lambda^2)
    double g[3];
    mat3_mul_vec(v.Ric, pts_i, tmp);
    mat3_mul_vec(RTh, tmp, g);
    // a fixed landmark has no Jacobian block (MakeHessian skips fixed vertices): J_lambda = B g = 0
    const bool lfix = v.lm_fixed && v.lm_fixed[l];
    const double il2 = lfix ? 0.0 : -1.0 / (lam * lam);
    g[0] *= il2; g[1] *= il2; g[2] *= il2;
    // G = -Ri hat(p_bi)
    double G[9];
    mat3_mul_hat(RTh, pbi, G);
#pragma unroll
    for (int k = 0; k < 9; ++k) G[k] = -G[k];

    double Ms[6] = {0, 0, 0, 0, 0, 0};  // Σ M_e, symmetric 3x3: 00 01 02 11 12 22
    double ms[3] = {0, 0, 0};           // Σ drho c B^T r
    // free extrinsic vertex (v17 4-vertex edge, A17/src/backend/edge_reprojection.cc:97-103): its H_lp row Σ J_ex^T W J_lambda
    const int xe = v.ext_pose;
    double wes[6] = {0, 0, 0, 0, 0, 0};

    for (int e = e0; e < e1; ++e) {
        const int j = v.e_pose_j[e];
        const double *RTj = v.poseRT + 16 * (size_t)j;
        double pcj[3], pbj[3], r[2];
        reproj_residual(v.Ric, v.tic, RTj, pw, v.e_pjx[e], v.e_pjy[e], pcj, pbj, r);
        const double iz = 1.0 / pcj[2];
        const double red[6] = {iz, 0.0, -pcj[0] * iz * iz, 0.0, iz, -pcj[1] * iz * iz};
        // A = Ric^T Rj^T  ;  B = red * A
        double A[9];
        {
            // (Rj Ric)^T
            double RjRic[9];
            mat3_mul(RTj, v.Ric, RjRic);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) A[3 * a + b] = RjRic[3 * b + a];
        }
        double B[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            B[c] = red[0] * A[c] + red[2] * A[6 + c];
            B[3 + c] = red[4] * A[3 + c] + red[5] * A[6 + c];
        }
        double rho0, drho, W[3];
        robust_weights(v.rp_loss, v.rp_delta, v.rp_info, r, rho0, drho, W);
        // WB = W B (2x3) ; M = B^T W B (sym 3x3) ; m = drho c B^T r
        double WB[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            WB[c] = W[0] * B[c] + W[1] * B[3 + c];
            WB[3 + c] = W[1] * B[c] + W[2] * B[3 + c];
        }
        double M[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[3 * a + b] = B[a] * WB[b] + B[3 + a] * WB[3 + b];
        const double dc = drho * v.rp_info;
        const double m[3] = {dc * (B[0] * r[0] + B[3] * r[1]), dc * (B[1] * r[0] + B[4] * r[1]),
                             dc * (B[2] * r[0] + B[5] * r[1])};
        Ms[0] += M[0]; Ms[1] += M[1]; Ms[2] += M[2]; Ms[3] += M[4]; Ms[4] += M[5]; Ms[5] += M[8];
        ms[0] += m[0]; ms[1] += m[1]; ms[2] += m[2];

        if (xe >= 0) {
            // J_ex = red * [ Ric^T (Rj^T Ri - I) | -T skew(p_ci) + skew(T p_ci) + skew(Ric^T (Rj^T (Ri tic + Pi - Pj) - tic)) ],
            // T = Ric^T Rj^T Ri Ric
            double RjtRi[9], M1[9], Lm[9], T1[9], Tm[9], S1[9], TS[9], tv[3], wv[3], uv[3], qv[3];
            mat3t_mul(RTj, RTh, RjtRi);
#pragma unroll
            for (int k = 0; k < 9; ++k) M1[k] = RjtRi[k] - ((k % 4 == 0) ? 1.0 : 0.0);
            mat3t_mul(v.Ric, M1, Lm);
            mat3t_mul(v.Ric, RjtRi, T1);
            mat3_mul(T1, v.Ric, Tm);
            mat3_mul_hat(Tm, pci, TS);                       // T skew(p_ci)
            mat3_mul_vec(Tm, pci, tv);                       // T p_ci
            mat3_mul_vec(RTh, v.tic, wv);
            wv[0] += RTh[9] - RTj[9]; wv[1] += RTh[10] - RTj[10]; wv[2] += RTh[11] - RTj[11];
            mat3t_mul_vec(RTj, wv, uv);
            uv[0] -= v.tic[0]; uv[1] -= v.tic[1]; uv[2] -= v.tic[2];
            mat3t_mul_vec(v.Ric, uv, qv);
            const double h1[9] = {0.0, -tv[2], tv[1], tv[2], 0.0, -tv[0], -tv[1], tv[0], 0.0};
            const double h2[9] = {0.0, -qv[2], qv[1], qv[2], 0.0, -qv[0], -qv[1], qv[0], 0.0};
#pragma unroll
            for (int k = 0; k < 9; ++k) S1[k] = -TS[k] + h1[k] + h2[k];
            double Je[12], WJe[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Je[c] = red[0] * Lm[c] + red[2] * Lm[6 + c];
                Je[6 + c] = red[4] * Lm[3 + c] + red[5] * Lm[6 + c];
                Je[3 + c] = red[0] * S1[c] + red[2] * S1[6 + c];
                Je[9 + c] = red[4] * S1[3 + c] + red[5] * S1[6 + c];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) { WJe[c] = W[0] * Je[c] + W[1] * Je[6 + c]; WJe[6 + c] = W[1] * Je[c] + W[2] * Je[6 + c]; }
            double Xe[36];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c) Xe[6 * a + c] = Je[a] * WJe[c] + Je[6 + a] * WJe[6 + c];
            s_add_diag(v, xe, Xe, 1.0);
            double *hd = v.hdiag + v.pose_off[xe], *be = v.bp + v.pose_off[xe];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                vio_add(hd + k, Xe[7 * k]);
                vio_add(be + k, -dc * (Je[k] * r[0] + Je[6 + k] * r[1]));
            }
            // H_lp row of the extrinsic vertex: J_ex^T W J_lambda, J_lambda = B g
            const double jl0 = B[0] * g[0] + B[1] * g[1] + B[2] * g[2], jl1 = B[3] * g[0] + B[4] * g[1] + B[5] * g[2];
#pragma unroll
            for (int k = 0; k < 6; ++k) wes[k] += WJe[k] * jl0 + WJe[6 + k] * jl1;
            // (ext, host) = J_ex^T W B [I G] ; (ext, j) = J_ex^T W B [-I N]
            double JeWB[18];  // 6x3 = J_ex^T (W B)
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) JeWB[3 * a + c] = Je[a] * WB[c] + Je[6 + a] * WB[3 + c];
            if (!hfix) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        Xe[6 * a + c] = JeWB[3 * a + c];
                        Xe[6 * a + 3 + c] = JeWB[3 * a] * G[c] + JeWB[3 * a + 1] * G[3 + c] + JeWB[3 * a + 2] * G[6 + c];
                    }
                s_add_block(v, xe, h, Xe, 1.0);
            }
            if (!v.pose_fixed[j]) {
                double Nn[9];
                mat3_mul_hat(RTj, pbj, Nn);
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        Xe[6 * a + c] = -JeWB[3 * a + c];
                        Xe[6 * a + 3 + c] = JeWB[3 * a] * Nn[c] + JeWB[3 * a + 1] * Nn[3 + c] + JeWB[3 * a + 2] * Nn[6 + c];
                    }
                s_add_block(v, xe, j, Xe, 1.0);
            }
        }

        double *wj = v.wo + 6 * (size_t)e;
        if (v.pose_fixed[j]) {
#pragma unroll
            for (int k = 0; k < 6; ++k) wj[k] = 0.0;
            continue;
        }
        // N = Rj hat(p_bj);  J_j = B [-I N]
        double N[9], MN[9], NMN[9];
        mat3_mul_hat(RTj, pbj, N);
        mat3_mul(M, N, MN);
        mat3t_mul(N, MN, NMN);
        // (j,j) += [[M, -MN],[-MN^T, N^T M N]]
        double X[36];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                X[6 * a + b] = M[3 * a + b];
                X[6 * a + 3 + b] = -MN[3 * a + b];
                X[6 * (3 + a) + b] = -MN[3 * b + a];
                X[6 * (3 + a) + 3 + b] = NMN[3 * a + b];
            }
        s_add_diag(v, j, X, 1.0);
        {
            double *hd = v.hdiag + v.pose_off[j];
#pragma unroll
            for (int k = 0; k < 6; ++k) vio_add(hd + k, X[7 * k]);
        }
        // b_j -= [-I; N^T] m
        double Ntm[3];
        mat3t_mul_vec(N, m, Ntm);
        {
            double *bj = v.bp + v.pose_off[j];
            vio_add(bj + 0, m[0]); vio_add(bj + 1, m[1]); vio_add(bj + 2, m[2]);
            vio_add(bj + 3, -Ntm[0]); vio_add(bj + 4, -Ntm[1]); vio_add(bj + 5, -Ntm[2]);
        }
        // w_j = J_j^T W J_lambda = [-I; N^T] M g
        double Mg[3], NtMg[3];
        mat3_mul_vec(M, g, Mg);
        mat3t_mul_vec(N, Mg, NtMg);
        wj[0] = -Mg[0]; wj[1] = -Mg[1]; wj[2] = -Mg[2];
        wj[3] = NtMg[0]; wj[4] = NtMg[1]; wj[5] = NtMg[2];
        if (!hfix) {
            // (h,j) = [I; G^T] M [-I N] = [[-M, MN],[-G^T M, G^T M N]]
            double GtM[9], GtMN[9];
            mat3t_mul(G, M, GtM);
            mat3t_mul(G, MN, GtMN);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    X[6 * a + b] = -M[3 * a + b];
                    X[6 * a + 3 + b] = MN[3 * a + b];
                    X[6 * (3 + a) + b] = -GtM[3 * a + b];
                    X[6 * (3 + a) + 3 + b] = GtMN[3 * a + b];
                }
            s_add_block(v, h, j, X, 1.0);
        }
    }
    // landmark block and host blocks from the summed M
    const double Msf[9] = {Ms[0], Ms[1], Ms[2], Ms[1], Ms[3], Ms[4], Ms[2], Ms[4], Ms[5]};
    double Mg[3];
    mat3_mul_vec(Msf, g, Mg);
    const double Hll = g[0] * Mg[0] + g[1] * Mg[1] + g[2] * Mg[2];
    const double bl = -(g[0] * ms[0] + g[1] * ms[1] + g[2] * ms[2]);
    v.Hll[l] = Hll;
    v.bl[l] = bl;
    double wh[6] = {0, 0, 0, 0, 0, 0};
    if (!hfix) {
        double GtMg[3];
        mat3t_mul_vec(G, Mg, GtMg);
        wh[0] = Mg[0]; wh[1] = Mg[1]; wh[2] = Mg[2];
        wh[3] = GtMg[0]; wh[4] = GtMg[1]; wh[5] = GtMg[2];
        double MG[9], GtMG[9], X[36];
        mat3_mul(Msf, G, MG);
        mat3t_mul(G, MG, GtMG);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                X[6 * a + b] = Msf[3 * a + b];
                X[6 * a + 3 + b] = MG[3 * a + b];
                X[6 * (3 + a) + b] = MG[3 * b + a];
                X[6 * (3 + a) + 3 + b] = GtMG[3 * a + b];
            }
        s_add_diag(v, h, X, 1.0);
        double *hd = v.hdiag + v.pose_off[h];
#pragma unroll
        for (int k = 0; k < 6; ++k) vio_add(hd + k, X[7 * k]);
        double Gtm[3];
        mat3t_mul_vec(G, ms, Gtm);
        double *bh = v.bp + v.pose_off[h];
        vio_add(bh + 0, -ms[0]); vio_add(bh + 1, -ms[1]); vio_add(bh + 2, -ms[2]);
        vio_add(bh + 3, -Gtm[0]); vio_add(bh + 4, -Gtm[1]); vio_add(bh + 5, -Gtm[2]);
    }
    double *whp = v.wh + 6 * (size_t)l;
#pragma unroll
    for (int k = 0; k < 6; ++k) whp[k] = wh[k];

    double *wep = nullptr;
    if (xe >= 0) {
        wep = v.we + 6 * (size_t)l;
#pragma unroll
        for (int k = 0; k < 6; ++k) wep[k] = wes[k];
    }
    if (!WITH_SCHUR || lfix) return;  // a fixed landmark is a constant: nothing to eliminate
    // Schur complement: S -= Hpl Hll^-1 Hlp ; bS -= Hpl Hll^-1 bl     (landmark diagonal is never damped)
    // vertex list: host (a = -1), observers (0 .. n_obs-1) and, when it is being estimated, the extrinsic vertex (a = n_obs)
    const double inv = 1.0 / Hll;
    const int n_obs = e1 - e0;
    const int n = n_obs + (xe >= 0 ? 1 : 0);
    for (int a = -1; a < n; ++a) {
        const int pa = a < 0 ? h : (a < n_obs ? v.e_pose_j[e0 + a] : xe);
        if (v.pose_fixed[pa]) continue;
        double wa[6];
        const double *wap = a < 0 ? whp : (a < n_obs ? v.wo + 6 * (size_t)(e0 + a) : wep);
#pragma unroll
        for (int k = 0; k < 6; ++k) wa[k] = wap[k] * inv;
        double *bc = v.bcorr + v.pose_off[pa];
#pragma unroll
        for (int k = 0; k < 6; ++k) vio_add(bc + k, wa[k] * bl);
        for (int b = a; b < n; ++b) {
            const int pb = b < 0 ? h : (b < n_obs ? v.e_pose_j[e0 + b] : xe);
            if (v.pose_fixed[pb]) continue;
            const double *wbp = b < 0 ? whp : (b < n_obs ? v.wo + 6 * (size_t)(e0 + b) : wep);
            double X[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) X[6 * r + c] = wa[r] * wbp[c];
            if (a == b) s_add_diag(v, pa, X, -1.0);
            else s_add_block(v, pa, pb, X, -1.0);
        }
    }
}

template <bool WITH_SCHUR>
__global__ void __launch_bounds__(128) k_linearize_lm(DevView v) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < v.L) linearize_landmark<WITH_SCHUR>(v, l);
}

// ------------------------------------------------------------------------------------------------
// EdgeSE3Prior: r = [log(Rp^-1 Ri); Pi - Pp], J = [[0, JrInv(r_R)],[I, 0]]
// ------------------------------------------------------------------------------------------------
struct Se3PriorView {
    int n;
    const int *pose;
    const double *p, *q, *info;
};

VIO_HD void se3prior_residual(const double *pose7, const double *pp, const double *qp, double r[6]) {
    // SO3(Qi), SO3(Qp): both normalised by the Sophus constructor
    double qi[4] = {pose7[3], pose7[4], pose7[5], pose7[6]};
    double qpn[4] = {qp[0], qp[1], qp[2], qp[3]};
    double ni = sqrt(qi[0] * qi[0] + qi[1] * qi[1] + qi[2] * qi[2] + qi[3] * qi[3]);
    double np = sqrt(qpn[0] * qpn[0] + qpn[1] * qpn[1] + qpn[2] * qpn[2] + qpn[3] * qpn[3]);
    for (int k = 0; k < 4; ++k) { qi[k] /= ni; qpn[k] /= np; }
    const double qpc[4] = {-qpn[0], -qpn[1], -qpn[2], qpn[3]};  // SO3::inverse = conjugate
    double qr[4];
    quat_mul(qpc, qi, qr);
    const double nr = sqrt(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
    for (int k = 0; k < 4; ++k) qr[k] /= nr;  // operator*= normalises
    so3_log(qr, r);
    r[3] = pose7[0] - pp[0];
    r[4] = pose7[1] - pp[1];
    r[5] = pose7[2] - pp[2];
}

VIO_HD void se3prior_edge(const DevView &v, const Se3PriorView &s, int i) {
    const int a = s.pose[i];
    if (v.pose_fixed[a]) return;
    double r[6];
    se3prior_residual(v.pose + 7 * (size_t)a, s.p + 3 * i, s.q + 4 * i, r);
    double Jr[9];
    so3_jr_inv(r, Jr);
    double J[36];
    for (int k = 0; k < 36; ++k) J[k] = 0.0;
    for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 3; ++c) J[6 * rr + 3 + c] = Jr[3 * rr + c];
    J[6 * 3 + 0] = 1.0; J[6 * 4 + 1] = 1.0; J[6 * 5 + 2] = 1.0;
    const double *Om = s.info + 36 * (size_t)i;
    // JtW = J^T Om ; H = JtW J ; b -= JtW r
    double JtW[36], X[36];
    for (int rr = 0; rr < 6; ++rr)
        for (int c = 0; c < 6; ++c) {
            double t = 0;
            for (int k = 0; k < 6; ++k) t += J[6 * k + rr] * Om[6 * k + c];
            JtW[6 * rr + c] = t;
        }
    for (int rr = 0; rr < 6; ++rr)
        for (int c = 0; c < 6; ++c) {
            double t = 0;
            for (int k = 0; k < 6; ++k) t += JtW[6 * rr + k] * J[6 * k + c];
            X[6 * rr + c] = t;
        }
    s_add_diag(v, a, X, 1.0);
    double *hd = v.hdiag + v.pose_off[a];
    double *bp = v.bp + v.pose_off[a];
    for (int rr = 0; rr < 6; ++rr) {
        vio_add(hd + rr, X[7 * rr]);
        double t = 0;
        for (int k = 0; k < 6; ++k) t += JtW[6 * rr + k] * r[k];
        vio_add(bp + rr, -t);
    }
}

__global__ void k_se3prior(DevView v, Se3PriorView s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < s.n) se3prior_edge(v, s, i);
}

__global__ void k_se3prior_chi2(DevView v, Se3PriorView s, double *out /* accumulates */) {
    // few edges: a single thread keeps the summation order fixed
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double chi = 0.0;
    for (int i = 0; i < s.n; ++i) {
        double r[6];
        se3prior_residual(v.pose + 7 * (size_t)s.pose[i], s.p + 3 * i, s.q + 4 * i, r);
        const double *Om = s.info + 36 * (size_t)i;
        for (int a = 0; a < 6; ++a) {
            double t = 0;
            for (int b = 0; b < 6; ++b) t += Om[6 * a + b] * r[b];
            chi += r[a] * t;
        }
    }
    *out += chi;
}

// ------------------------------------------------------------------------------------------------
// chi2 pass: Σ rho(c r.r) over reprojection edges; one thread per landmark (host chain shared)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_chi2_lm(DevView v, double *partial) {
    double chi = 0.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < v.L; l += gridDim.x * blockDim.x) {
        const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
        if (e0 == e1) continue;
        const int h = v.lm_host[l];
        const double lam = v.invdep[l];
        const double *RTh = v.poseRT + 16 * (size_t)h;
        const double pci[3] = {v.lm_pix[l] / lam, v.lm_piy[l] / lam, v.lm_piz[l] / lam};
        double pbi[3], pw[3];
        mat3_mul_vec(v.Ric, pci, pbi);
        pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
        mat3_mul_vec(RTh, pbi, pw);
        pw[0] += RTh[9]; pw[1] += RTh[10]; pw[2] += RTh[11];
        // the next edge's observer index and observation are requested while the current residual is evaluated
        int nj = v.e_pose_j[e0];
        double n_pjx = v.e_pjx[e0], n_pjy = v.e_pjy[e0];
        for (int e = e0; e < e1; ++e) {
            const double *RTj = v.poseRT + 16 * (size_t)nj;
            const double pjx = n_pjx, pjy = n_pjy;
            if (e + 1 < e1) { nj = v.e_pose_j[e + 1]; n_pjx = v.e_pjx[e + 1]; n_pjy = v.e_pjy[e + 1]; }
            double pcj[3], pbj[3], r[2];
            reproj_residual(v.Ric, v.tic, RTj, pw, pjx, pjy, pcj, pbj, r);
            const double e2 = v.rp_info * (r[0] * r[0] + r[1] * r[1]);
            if (v.rp_loss == 0) chi += e2;
            else {
                double rho[3];
                loss_compute(v.rp_loss, v.rp_delta, e2, rho);
                chi += rho[0];
            }
        }
    }
    block_sum_to(chi, partial);
}

// The same sum for SMALL graphs (sliding windows, TestMonoBA): with a few hundred landmarks the kernel above is one latency
// chain of ~20 edges per thread on a handful of warps.  Here a warp takes LPW landmarks (lane <-> landmark for the host chain
// p_w) and walks their contiguous run of edges lane <-> edge; an edge finds its landmark by a binary search over the lanes'
// first-edge indices and fetches p_w by shuffle.  (On large scenes this mapping loses: the observers' R, t become a many-way
// gather instead of a broadcast - 320 us against 75 us at config 5.)
__global__ void __launch_bounds__(256) k_chi2_lm_small(DevView v, double *partial) {
    double chi = 0.0;
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
    const int LPW = v.L >= 8 * nwg ? 8 : 2;
    for (int l0 = wg * LPW; l0 < v.L; l0 += nwg * LPW) {
        const int nl = min(LPW, v.L - l0), l = l0 + min(lane, nl - 1);
        const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
        const int E0 = __shfl_sync(0xffffffffu, e0, 0), E1 = __shfl_sync(0xffffffffu, e1, nl - 1);
        double pw[3];
        {
            const int h = v.lm_host[l];
            const double lam = v.invdep[l];
            const double *RTh = v.poseRT + 16 * (size_t)h;
            const double pci[3] = {v.lm_pix[l] / lam, v.lm_piy[l] / lam, v.lm_piz[l] / lam};
            double pbi[3];
            mat3_mul_vec(v.Ric, pci, pbi);
            pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
            mat3_mul_vec(RTh, pbi, pw);
            pw[0] += RTh[9]; pw[1] += RTh[10]; pw[2] += RTh[11];
        }
        for (int eb = E0; eb < E1; eb += 32) {
            const int e = eb + lane;
            const bool ev = e < E1;
            int r = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cand = r + step;
                const int ce0 = __shfl_sync(0xffffffffu, e0, cand & 31);
                if (cand < nl && ce0 <= e) r = cand;
            }
            const double pwe[3] = {__shfl_sync(0xffffffffu, pw[0], r), __shfl_sync(0xffffffffu, pw[1], r), __shfl_sync(0xffffffffu, pw[2], r)};
            if (!ev) continue;
            const double *RTj = v.poseRT + 16 * (size_t)v.e_pose_j[e];
            double pcj[3], pbj[3], rr[2];
            reproj_residual(v.Ric, v.tic, RTj, pwe, v.e_pjx[e], v.e_pjy[e], pcj, pbj, rr);
            const double e2 = v.rp_info * (rr[0] * rr[0] + rr[1] * rr[1]);
            if (v.rp_loss == 0) chi += e2;
            else {
                double rho[3];
                loss_compute(v.rp_loss, v.rp_delta, e2, rho);
                chi += rho[0];
            }
        }
    }
    block_sum_to(chi, partial);
}

// ------------------------------------------------------------------------------------------------
// finalize reduced system after (all-reduced) accumulation: mirror the upper triangle, bS = bp - bcorr
// ------------------------------------------------------------------------------------------------
__global__ void k_mirror_dense(double *S, int P) {  // gridDim.z = number of stacked P x P problems
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    S += (size_t)blockIdx.z * P * P;
    if (r < P && c < P && r > c) S[(size_t)r * P + c] = S[(size_t)c * P + r];
}
__global__ void k_mirror_bsr(DevView v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= v.nnzb * 36) return;
    const int id = (int)(t / 36), k = (int)(t % 36), r = k / 6, c = k % 6;
    const int tr = v.bsr_tr[id];
    if (tr == id) {  // diagonal block
        if (r > c) v.S[36 * (size_t)id + k] = v.S[36 * (size_t)id + 6 * c + r];
    } else if (tr < id) {  // lower block = transpose of its upper partner
        v.S[36 * (size_t)id + k] = v.S[36 * (size_t)tr + 6 * c + r];
    }
}
__global__ void k_finalize_b(DevView v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.P) v.bS[i] = v.bp[i] - v.bcorr[i];
}

// max |diag(Hessian_)| : pose part from hdiag, landmark part from Hll
__global__ void k_maxdiag(DevView v, double *partial) {
    double m = 0.0;
    const int n = v.P + v.L;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        m = fmax(m, fabs(i < v.P ? v.hdiag[i] : v.Hll[i - v.P]));
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    m = warp_max(m);
    if (lane == 0) sm[wid] = m;
    __syncthreads();
    if (wid == 0) {
        double t = lane < ((blockDim.x + 31) >> 5) ? sm[lane] : 0.0;
        t = warp_max(t);
        if (lane == 0) partial[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// back-substitution and the LM scalars of IsGoodStepInLM
// ------------------------------------------------------------------------------------------------
// dxl = Hll^-1 (bl - Hlp dxp); also partial sums of  dx_l (lambda dx_l + b_l)  and dx_l^2
// lam_p (here and in the other kernels that take the damping): when not null, lambda is read from device memory - the LM body
// is replayed as a CUDA graph whose kernel parameters are frozen while lambda changes every iteration
// Mapping: a warp takes 32 consecutive landmarks; their observer rows of H_lp are one contiguous run of `wo` (edges are stored
// landmark by landmark), read lane <-> edge (48 contiguous bytes per lane, three 16-byte loads: fully coalesced) and summed per
// landmark by a segmented warp scan in a fixed order; the host row, b_l and H_ll are read lane <-> landmark.
#define VIO_BACKSUB_THREADS 512
__global__ void __launch_bounds__(VIO_BACKSUB_THREADS) k_backsub(DevView v, double lambda, double *partial_scale, double *partial_n2,
                                                                 const double *lam_p = nullptr) {
    __shared__ double s_obs[VIO_BACKSUB_THREADS];
    if (lam_p) lambda = *lam_p;
    double sc = 0.0, n2 = 0.0;
    const int lane = threadIdx.x & 31;
    double *obs = s_obs + (threadIdx.x & ~31);
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
    // landmarks per warp: 32 on large scenes; small graphs (a sliding window has ~1000 landmarks) spread over more warps, or a
    // handful of warps would walk all the edge rounds one after the other
    const int LPW = v.L >= 32 * nwg ? 32 : (v.L >= 8 * nwg ? 8 : 2);
    for (int l0 = wg * LPW; l0 < v.L; l0 += nwg * LPW) {
        const int nl = min(LPW, v.L - l0), l = l0 + min(lane, nl - 1);
        const bool mine = lane < nl;
        const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
        const int E0 = __shfl_sync(0xffffffffu, e0, 0), E1 = __shfl_sync(0xffffffffu, e1, nl - 1);
        // per-landmark operands first: their loads are in flight while the edge rounds run
        const double bl = v.bl[l], hll = v.Hll[l];
        double th = 0.0;
        {
            const double2 *wh = reinterpret_cast<const double2 *>(v.wh + 6 * (size_t)l);
            const double *dh = v.dxp + v.pose_off[v.lm_host[l]];
            const double2 w0 = wh[0], w1 = wh[1], w2 = wh[2];
            th = w0.x * dh[0] + w0.y * dh[1] + w1.x * dh[2] + w1.y * dh[3] + w2.x * dh[4] + w2.y * dh[5];
        }
        obs[lane] = 0.0;
        __syncwarp();
        // software pipeline: the row and the observer index of the NEXT round are requested before this round is reduced
        double2 n0 = make_double2(0.0, 0.0), n1 = n0, n2_ = n0;
        int nj = 0;
        if (E0 + lane < E1) {
            const double2 *w = reinterpret_cast<const double2 *>(v.wo + 6 * (size_t)(E0 + lane));
            n0 = w[0]; n1 = w[1]; n2_ = w[2];
            nj = v.e_pose_j[E0 + lane];
        }
        for (int eb = E0; eb < E1; eb += 32) {
            const int e = eb + lane;
            const bool ev = e < E1;
            const double2 w0 = n0, w1 = n1, w2 = n2_;
            const int j = nj;
            if (e + 32 < E1) {
                const double2 *w = reinterpret_cast<const double2 *>(v.wo + 6 * (size_t)(e + 32));
                n0 = w[0]; n1 = w[1]; n2_ = w[2];
                nj = v.e_pose_j[e + 32];
            }
            double part = 0.0;
            if (ev) {
                const double *dj = v.dxp + v.pose_off[j];
                part = w0.x * dj[0] + w0.y * dj[1] + w1.x * dj[2] + w1.y * dj[3] + w2.x * dj[4] + w2.y * dj[5];
            }
            // landmark of this edge, relative to l0: the last r with eptr[l0 + r] <= e (binary search over the lanes' e0)
            int r = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cand = r + step;
                const int ce0 = __shfl_sync(0xffffffffu, e0, cand & 31);
                if (cand < nl && ce0 <= e) r = cand;
            }
            if (!ev) r = -1;
            // segmented inclusive scan: afterwards the last lane of a run of equal r holds the run's sum
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, part, off);
                const int ur = __shfl_up_sync(0xffffffffu, r, off);
                if (lane >= off && ur == r) part += up;
            }
            const int rn = __shfl_down_sync(0xffffffffu, r, 1);
            if (ev && (lane == 31 || rn != r)) obs[r] += part;  // one writer per landmark and round
            __syncwarp();
        }
        double t = bl - th - obs[lane];
        __syncwarp();
        if (v.ext_pose >= 0) {  // free extrinsic vertex
            const double *w = v.we + 6 * (size_t)l;
            const double *dx = v.dxp + v.pose_off[v.ext_pose];
#pragma unroll
            for (int k = 0; k < 6; ++k) t -= w[k] * dx[k];
        }
        if (!mine) continue;
        if (e0 == e1) { v.dxl[l] = 0.0; continue; }
        const double d = (v.lm_fixed && v.lm_fixed[l]) ? 0.0 : t / hll;  // fixed landmark: constant
        v.dxl[l] = d;
        sc += d * (lambda * d + bl);
        n2 += d * d;
    }
    block_sum_to(sc, partial_scale);
    block_sum_to(n2, partial_n2);
}

// own: nullptr, or (multi-GPU, un-reduced b_p) 1 for the rows this rank owns - the quadratic terms are counted there only
__global__ void __launch_bounds__(256) k_pose_scale(DevView v, double lambda, const uint8_t *__restrict__ own, double *partial_scale,
                                                    double *partial_n2, const double *lam_p = nullptr) {
    if (lam_p) lambda = *lam_p;
    double sc = 0.0, n2 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.P; i += gridDim.x * blockDim.x) {
        const double d = v.dxp[i];
        const double mine = (own == nullptr || own[i]) ? 1.0 : 0.0;
        sc += d * (mine * lambda * d + v.bp[i]);
        n2 += mine * d * d;
    }
    block_sum_to(sc, partial_scale);
    block_sum_to(n2, partial_n2);
}

// UpdateStates: backup + Plus.  sign = +1 (update) ; v15 rollback calls it again with sign = -1, no backup.
VIO_HD void update_pose(const DevView &v, int i, double sign, int backup) {
    if (v.act && !v.act[i / v.Cper]) return;  // lock-step batch: this problem is not taking a step
    double *p = v.pose + 7 * (size_t)i;
    if (backup) {
        double *b = v.pose_bak + 7 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 7; ++k) b[k] = p[k];
    }
    const double *d = v.dxp + v.pose_off[i];
    p[0] += sign * d[0]; p[1] += sign * d[1]; p[2] += sign * d[2];
    const double w[3] = {sign * d[3], sign * d[4], sign * d[5]};
    double dq[4], q[4] = {p[3], p[4], p[5], p[6]}, qn[4];
    so3_exp(w, dq);
    quat_mul(q, dq, qn);  // right multiplication; the reference's q.normalized() result is discarded
    p[3] = qn[0]; p[4] = qn[1]; p[5] = qn[2]; p[6] = qn[3];
}
__global__ void k_update_pose(DevView v, double sign, int backup) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.C) update_pose(v, i, sign, backup);
}
__global__ void k_update_sb(DevView v, double sign, int backup) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.NSB) return;
    if (v.act && !v.act[i / v.NSBper]) return;
    double *p = v.sb + 9 * (size_t)i;
    const double *d = v.dxp + v.sb_off[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (backup) v.sb_bak[9 * (size_t)i + k] = p[k];
        p[k] += sign * d[k];
    }
}
__global__ void k_update_lm(DevView v, double sign, int backup) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= v.L) return;
    if (v.act && !v.act[v.lm_prob[l]]) return;
    if (backup) v.invdep_bak[l] = v.invdep[l];
    v.invdep[l] += sign * v.dxl[l];
}
// UpdateStates of all three vertex classes in one launch (thread i: pose i, speed-bias i, landmark i)
__global__ void k_update_all(DevView v, double sign, int backup) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.C) update_pose(v, i, sign, backup);
    if (i < v.NSB && !(v.act && !v.act[i / v.NSBper])) {
        double *p = v.sb + 9 * (size_t)i;
        const double *d = v.dxp + v.sb_off[i];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (backup) v.sb_bak[9 * (size_t)i + k] = p[k];
            p[k] += sign * d[k];
        }
    }
    if (i < v.L && !(v.act && !v.act[v.lm_prob[i]])) {
        if (backup) v.invdep_bak[i] = v.invdep[i];
        v.invdep[i] += sign * v.dxl[i];
    }
}
__global__ void k_restore(DevView v) {  // v.act (batch): restore only the problems whose step was rejected
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)v.C * 7 && (!v.act || v.act[(t / 7) / v.Cper])) v.pose[t] = v.pose_bak[t];
    if (t < (long long)v.NSB * 9 && (!v.act || v.act[(t / 9) / v.NSBper])) v.sb[t] = v.sb_bak[t];
    if (t < v.L && (!v.act || v.act[v.lm_prob[t]])) v.invdep[t] = v.invdep_bak[t];
}
