// vio_marg.cuh — Problem::Marginalize on the device.
// Reference: A17/src/backend/problem.cc:617-795 (A17 = /root/reference/workspace/assignments/17-vins-initialization/vins-mono):
//   1. edges connected to the frame being marginalised are re-linearised into a dense H_marg / b_marg over
//      [pose-class block (pose_dim) | their landmarks]; NO vertex is treated as fixed here (:651-682), so the
//      4-vertex EdgeReprojection also contributes its extrinsic Jacobian (A17/src/backend/edge_reprojection.cc:97-103);
//   2. the landmarks are Schur-eliminated (:690-712), the old prior is added (:714-719);
//   3. the marginalised vertices are permuted to the bottom-right (:724-748) and eliminated with an
//      eigen-decomposition pseudo-inverse, eps = 1e-8 (:750-768);
//   4. the new prior is re-factored: J = sqrt(S) V^T, Jt_prior_inv = sqrt(S^-1) V^T, err = -Jt_prior_inv b,
//      H_prior = J^T J with |entries| <= 1e-9 zeroed (:770-784).
// The symmetric eigen-decompositions (Eigen::SelfAdjointEigenSolver upstream) are a one-sided Jacobi iteration in one
// CTA: columns of U = A V are orthogonalised pairwise (round-robin ordering, one warp per column pair), eigenvalue
// lambda_p = v_p . u_p, sorted ascending like Eigen.  Eigenvectors are defined up to sign, so Jt_prior_inv / err_prior
// agree with the reference up to a per-row sign; H_prior, b_prior, |err_prior| and Jt^T Jt are sign-free.
#pragma once
#include "vio_dev.h"
#include "vio_kernels.cuh"
#include "vio_imu.cuh"

struct MargEdgeView {
    int n;
    const int *edge;   // packed edge index
    const int *lm;     // local landmark index of the edge
    const int *mslot;  // index of that landmark among the marginalised landmarks
    int ext_off;       // ordering offset of the extrinsic vertex, -1 if extrinsics are constants (v15 style)
};

// one thread per connected reprojection edge: H_marg += J^T W J (upper element triangle), b_marg -= drho c J^T r
__global__ void k_marg_reproj(DevView v, MargEdgeView m, double *H, double *b, int n_tot, int P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m.n) return;
    const int e = m.edge[t], l = m.lm[t];
    const int h = v.lm_host[l], j = v.e_pose_j[e];
    const double *RTh = v.poseRT + 16 * (size_t)h, *RTj = v.poseRT + 16 * (size_t)j;
    const double lam = v.invdep[l];
    const double pts_i[3] = {v.lm_pix[l], v.lm_piy[l], v.lm_piz[l]};
    const double pci[3] = {pts_i[0] / lam, pts_i[1] / lam, pts_i[2] / lam};
    double pbi[3], pw[3];
    mat3_mul_vec(v.Ric, pci, pbi);
    for (int k = 0; k < 3; ++k) pbi[k] += v.tic[k];
    mat3_mul_vec(RTh, pbi, pw);
    for (int k = 0; k < 3; ++k) pw[k] += RTh[9 + k];
    double pcj[3], pbj[3], r[2];
    reproj_residual(v.Ric, v.tic, RTj, pw, v.e_pjx[e], v.e_pjy[e], pcj, pbj, r);
    const double iz = 1.0 / pcj[2];
    const double red[6] = {iz, 0.0, -pcj[0] * iz * iz, 0.0, iz, -pcj[1] * iz * iz};
    // A = Ric^T Rj^T
    double RjRic[9], A[9], RicT[9];
    mat3_mul(RTj, v.Ric, RjRic);
    for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) { A[3 * a + c] = RjRic[3 * c + a]; RicT[3 * a + c] = v.Ric[3 * c + a]; }
    // 3 x 19 "jaco" [lambda | pose_i | pose_j | ext] before the 2x3 reduce
    double Jf[3 * 19];
    for (int k = 0; k < 57; ++k) Jf[k] = 0.0;
    double ARi[9], Hh[9], T1[9];
    mat3_mul(A, RTh, ARi);
    // lambda: A Ri Ric pts_i * (-1/lambda^2)
    {
        double Rp[3], col[3];
        mat3_mul_vec(v.Ric, pts_i, Rp);
        mat3_mul_vec(ARi, Rp, col);
        const double il2 = (v.lm_fixed && v.lm_fixed[l]) ? 0.0 : -1.0 / (lam * lam);  // fixed landmark: no Jacobian block
        for (int a = 0; a < 3; ++a) Jf[19 * a + 0] = col[a] * il2;
    }
    // pose_i: [A, A Ri * -hat(p_bi)]
    {
        const double hx[9] = {0, pbi[2], -pbi[1], -pbi[2], 0, pbi[0], pbi[1], -pbi[0], 0};  // -hat(p_bi)
        mat3_mul(ARi, hx, T1);
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) { Jf[19 * a + 1 + c] = A[3 * a + c]; Jf[19 * a + 4 + c] = T1[3 * a + c]; }
    }
    // pose_j: [-A, Ric^T hat(p_bj)]
    {
        const double hx[9] = {0, -pbj[2], pbj[1], pbj[2], 0, -pbj[0], -pbj[1], pbj[0], 0};
        mat3_mul(RicT, hx, Hh);
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) { Jf[19 * a + 7 + c] = -A[3 * a + c]; Jf[19 * a + 10 + c] = Hh[3 * a + c]; }
    }
    // ext: [Ric^T (Rj^T Ri - I), -tmp_r skew(p_ci) + skew(tmp_r p_ci) + skew(Ric^T (Rj^T (Ri tic + Pi - Pj) - tic))]
    if (m.ext_off >= 0) {
        double tmp_r[9], tp[3], q1[3], q2[3], q3[3];
        mat3_mul(ARi, v.Ric, tmp_r);
        const double sk1[9] = {0, -pci[2], pci[1], pci[2], 0, -pci[0], -pci[1], pci[0], 0};
        double M1[9];
        mat3_mul(tmp_r, sk1, M1);
        mat3_mul_vec(tmp_r, pci, tp);
        mat3_mul_vec(RTh, v.tic, q1);
        for (int k = 0; k < 3; ++k) q1[k] += RTh[9 + k] - RTj[9 + k];
        mat3t_mul_vec(RTj, q1, q2);
        for (int k = 0; k < 3; ++k) q2[k] -= v.tic[k];
        mat3_mul_vec(RicT, q2, q3);
        const double s2[9] = {0, -tp[2], tp[1], tp[2], 0, -tp[0], -tp[1], tp[0], 0};
        const double s3[9] = {0, -q3[2], q3[1], q3[2], 0, -q3[0], -q3[1], q3[0], 0};
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) {
                Jf[19 * a + 13 + c] = ARi[3 * a + c] - RicT[3 * a + c];
                Jf[19 * a + 16 + c] = -M1[3 * a + c] + s2[3 * a + c] + s3[3 * a + c];
            }
    }
    // J = reduce * Jf (2 x 19)
    double J[38];
    for (int c = 0; c < 19; ++c) {
        J[c] = red[0] * Jf[c] + red[2] * Jf[38 + c];
        J[19 + c] = red[4] * Jf[19 + c] + red[5] * Jf[38 + c];
    }
    double rho0, drho, W[3];
    robust_weights(v.rp_loss, v.rp_delta, v.rp_info, r, rho0, drho, W);
    int gidx[19];
    gidx[0] = P + m.mslot[t];
    for (int k = 0; k < 6; ++k) {
        gidx[1 + k] = v.pose_off[h] + k;
        gidx[7 + k] = v.pose_off[j] + k;
        gidx[13 + k] = m.ext_off >= 0 ? m.ext_off + k : -1;
    }
    const double dc = drho * v.rp_info;
    for (int a = 0; a < 19; ++a) {
        const int ga = gidx[a];
        if (ga < 0) continue;
        const double wa0 = W[0] * J[a] + W[1] * J[19 + a], wa1 = W[1] * J[a] + W[2] * J[19 + a];
        for (int c = 0; c < 19; ++c) {
            const int gc = gidx[c];
            if (gc < 0 || ga > gc) continue;
            vio_add(H + (size_t)ga * n_tot + gc, wa0 * J[c] + wa1 * J[19 + c]);
        }
        vio_add(b + ga, -dc * (J[a] * r[0] + J[19 + a] * r[1]));
    }
}

// IMU edges connected to the marginalised frame: same evaluation as k_imu_linearize, no fixed-vertex masking
__global__ void __launch_bounds__(256) k_marg_imu(ImuView s, DevView v, const int *edge_list, double *H, double *b, int n_tot) {
    __shared__ double J[450], OJ[450], r[15], Or[15];
    __shared__ int gidx[30];
    const int e = edge_list[blockIdx.x];
    if (threadIdx.x == 0) {
        imu_edge_eval(s, v, e, r, J);
        const int offs[4] = {v.pose_off[s.pose_i[e]], v.sb_off[s.sb_i[e]], v.pose_off[s.pose_j[e]], v.sb_off[s.sb_j[e]]};
        const int dims[4] = {6, 9, 6, 9};
        int c = 0;
        for (int k = 0; k < 4; ++k)
            for (int d = 0; d < dims[k]; ++d) gidx[c++] = offs[k] + d;
    }
    __syncthreads();
    const double *Om = s.info + 225 * (size_t)e;
    for (int t = threadIdx.x; t < 450; t += blockDim.x) {
        const int a = t / 30, c = t % 30;
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += Om[15 * a + k] * J[30 * k + c];
        OJ[t] = acc;
    }
    if (threadIdx.x < 15) {
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += Om[15 * threadIdx.x + k] * r[k];
        Or[threadIdx.x] = acc;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 900; t += blockDim.x) {
        const int a = t / 30, c = t % 30;
        const int ga = gidx[a], gc = gidx[c];
        if (ga > gc) continue;
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += J[30 * k + a] * OJ[30 * k + c];
        atomicAdd(H + (size_t)ga * n_tot + gc, acc);
    }
    if (threadIdx.x < 30) {
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += J[30 * k + threadIdx.x] * Or[k];
        atomicAdd(b + gidx[threadIdx.x], -acc);
    }
}

// Hpp = H[0:P,0:P] - Hpm Hmm^-1 Hmp (+ prior), bpp likewise; Hmm is diagonal (inverse-depth landmarks)
__global__ void k_marg_schur(const double *H, const double *b, int n_tot, int P, int Mm, const double *Hprior,
                             const double *bprior, double *Hpp, double *bpp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * (P + 1)) return;
    const int r = t / (P + 1), c = t % (P + 1);
    double acc = 0.0;
    for (int l = 0; l < Mm; ++l) {
        const double hrl = H[(size_t)r * n_tot + P + l];
        if (hrl == 0.0) continue;
        const double inv = 1.0 / H[(size_t)(P + l) * n_tot + P + l];
        acc += hrl * inv * (c < P ? H[(size_t)(P + l) * n_tot + c] : b[P + l]);
    }
    if (c < P) Hpp[(size_t)r * P + c] = H[(size_t)r * n_tot + c] - acc + (Hprior ? Hprior[(size_t)r * P + c] : 0.0);
    else bpp[r] = b[r] - acc + (bprior ? bprior[r] : 0.0);
}

// gather with the reference's row/column permutation (marginalised vertices last)
__global__ void k_marg_permute(const double *Hpp, const double *bpp, const int *perm, int P, double *Hq, double *bq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * (P + 1)) return;
    const int r = t / (P + 1), c = t % (P + 1);
    if (c < P) Hq[(size_t)r * P + c] = Hpp[(size_t)perm[r] * P + perm[c]];
    else bq[r] = bpp[perm[r]];
}

// ---- one-sided Jacobi eigen-decomposition of a symmetric n x n matrix, one CTA ---------------------------------------
// A: row-major n x n (element (r,c) at A[r*lda + c]); U, V: column-major n x n workspaces; lam: n eigenvalues ascending;
// order: column of V holding the k-th smallest eigenvalue.
__global__ void __launch_bounds__(1024) k_jacobi_eigh(const double *A, int lda, int n, double *U, double *V, double *lam, int *order) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    __shared__ int s_rot;
    for (int i = tid; i < n * n; i += nt) {
        const int c = i / n, r = i % n;
        U[i] = 0.5 * (A[(size_t)r * lda + c] + A[(size_t)c * lda + r]);
        V[i] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int m = (n + 1) & ~1;  // round-robin needs an even number of players; player m-1 may be a dummy
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < m - 1; ++round) {
            for (int pr = warp; pr < m / 2; pr += nw) {
                // circle method: player m-1 fixed, others rotate
                int a = pr == 0 ? m - 1 : (round + pr) % (m - 1);
                int bq = (round + (m - 1) - pr) % (m - 1);
                if (pr == 0) bq = round % (m - 1);
                int p = min(a, bq), q = max(a, bq);
                if (q >= n || p == q) continue;
                double *up = U + (size_t)p * n, *uq = U + (size_t)q * n;
                double al = 0, be = 0, ga = 0;
                for (int i = lane; i < n; i += 32) {
                    const double x = up[i], y = uq[i];
                    al += x * x; be += y * y; ga += x * y;
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
                if (lane == 0) s_rot = 1;
                const double zeta = (be - al) / (2.0 * ga);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                double *vp = V + (size_t)p * n, *vq = V + (size_t)q * n;
                for (int i = lane; i < n; i += 32) {
                    const double x = up[i], y = uq[i];
                    up[i] = cs * x - sn * y; uq[i] = sn * x + cs * y;
                    const double vx = vp[i], vy = vq[i];
                    vp[i] = cs * vx - sn * vy; vq[i] = sn * vx + cs * vy;
                }
            }
            __syncthreads();
        }
        const int any = s_rot;
        __syncthreads();
        if (!any) break;
    }
    // eigenvalues: lambda_p = v_p . u_p  (u_p = A v_p)
    for (int p = warp; p < n; p += nw) {
        double d = 0;
        for (int i = lane; i < n; i += 32) d += V[(size_t)p * n + i] * U[(size_t)p * n + i];
        d = warp_sum(d);
        if (lane == 0) lam[n + p] = d;  // unsorted copy in the second half of lam
    }
    __syncthreads();
    // ascending rank sort (ties broken by index)
    for (int p = tid; p < n; p += nt) {
        const double lp = lam[n + p];
        int rank = 0;
        for (int q = 0; q < n; ++q) {
            const double lq = lam[n + q];
            rank += (lq < lp) || (lq == lp && q < p);
        }
        lam[rank] = lp;
        order[rank] = p;
    }
}

// Schur elimination of the marginalised block with the eigen pseudo-inverse (eps), then the prior re-factorisation inputs.
//   Hq: P x P permuted, n2 = kept dim, m2 = marginalised dim; (Vm, lamm, ordm) = eigh of Amm
__global__ void k_marg_eliminate(const double *Hq, const double *bq, int P, int n2, int m2, const double *Vm, const double *lamm,
                                 const int *ordm, double eps, double *Ainv /* m2 x m2 */, double *Hn /* n2 x n2 */, double *bn) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // Amm_inv = V diag(lam > eps ? 1/lam : 0) V^T
    for (int t = tid; t < m2 * m2; t += nt) {
        const int r = t / m2, c = t % m2;
        double acc = 0.0;
        for (int k = 0; k < m2; ++k) {
            const double l = lamm[k];
            if (l > eps) acc += Vm[(size_t)ordm[k] * m2 + r] * (1.0 / l) * Vm[(size_t)ordm[k] * m2 + c];
        }
        Ainv[t] = acc;
    }
    __syncthreads();
    // tempB = Arm Amm_inv ; Hn = Arr - tempB Amr ; bn = brr - tempB bmm      (grid-stride over one CTA: sizes are tiny)
    for (int t = tid; t < n2 * (n2 + 1); t += nt) {
        const int r = t / (n2 + 1), c = t % (n2 + 1);
        double acc = 0.0;
        for (int k = 0; k < m2; ++k) {
            double tb = 0.0;
            for (int q = 0; q < m2; ++q) tb += Hq[(size_t)r * P + n2 + q] * Ainv[q * m2 + k];
            acc += tb * (c < n2 ? Hq[(size_t)(n2 + k) * P + c] : bq[n2 + k]);
        }
        if (c < n2) Hn[(size_t)r * n2 + c] = Hq[(size_t)r * P + c] - acc;
        else bn[r] = bq[r] - acc;
    }
}

// Jt_inv = diag(sqrt(S_inv)) V^T ; err = -Jt_inv b ; H = V diag(S) V^T with |.| <= 1e-9 zeroed    (S = lam > eps ? lam : 0)
__global__ void k_marg_refactor(const double *V, const double *lam, const int *ord, int n, double eps, const double *bn,
                                double *Jt, double *err, double *Hout) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < n * n; t += nt) {
        const int r = t / n, c = t % n;
        const double l = lam[r];
        Jt[t] = l > eps ? sqrt(1.0 / l) * V[(size_t)ord[r] * n + c] : 0.0;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
            const double lk = lam[k];
            if (lk > eps) acc += V[(size_t)ord[k] * n + r] * lk * V[(size_t)ord[k] * n + c];
        }
        Hout[t] = fabs(acc) > 1e-9 ? acc : 0.0;
    }
    __syncthreads();
    for (int r = tid; r < n; r += nt) {
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc += Jt[(size_t)r * n + c] * bn[c];
        err[r] = -acc;
    }
}
