// vio_xyz.cuh — VertexPointXYZ landmarks with EdgeReprojectionXYZ observations
// (A15/backend/vertex_point_xyz.h:12-18, A15/backend/edge_reprojection.cc:113-163; the same factor in
// A17/src/backend/edge_reprojection.cc:130-180).  2-vertex factor [X_w, T_i]:
//     p_b = Q_i^-1 (X_w - P_i),  p_c = q_ic^-1 (p_b - t_ic),  r = p_c.xy / p_c.z - obs.xy
//     J_X = reduce R_ic^T R_i^T (2x3),   J_T = reduce [ -R_ic^T R_i^T | R_ic^T hat(p_b) ] (2x6)
// One thread per point (the per-landmark pattern of k_linearize_lm: register accumulation of the 3x3 landmark block,
// FP64 atomics into the reduced system); the landmark block is 3x3, so the Schur complement uses its closed-form inverse
// (the reference: Hmm.block(idx, idx, 3, 3).inverse(), A15/backend/problem.cc:383-388).  The bodies are VIO_HD so
// tests/host_emul.cu can run them on the CPU.
#pragma once
#include "vio_dev.h"
#include "vio_kernels.cuh"

struct XyzEdge {
    double r[2];
    double JX[6];   // 2x3
    double JT[12];  // 2x6
};

VIO_HD void xyz_residual(const DevView &v, const double X[3], int e, double pb[3], double pc[3], double r[2]) {
    const double *RT = v.poseRT + 16 * (size_t)v.ex_pose[e];
    const double d[3] = {X[0] - RT[9], X[1] - RT[10], X[2] - RT[11]};
    mat3t_mul_vec(RT, d, pb);
    const double q[3] = {pb[0] - v.tic[0], pb[1] - v.tic[1], pb[2] - v.tic[2]};
    mat3t_mul_vec(v.Ric, q, pc);
    const double z = pc[2];
    r[0] = pc[0] / z - v.ex_ox[e];
    r[1] = pc[1] / z - v.ex_oy[e];
}

VIO_HD void xyz_edge(const DevView &v, const double X[3], int e, XyzEdge &o) {
    double pb[3], pc[3];
    xyz_residual(v, X, e, pb, pc, o.r);
    const double *RT = v.poseRT + 16 * (size_t)v.ex_pose[e];
    const double iz = 1.0 / pc[2];
    const double red[6] = {iz, 0.0, -pc[0] * iz * iz, 0.0, iz, -pc[1] * iz * iz};
    double A[9], RjRic[9];  // A = Ric^T Rj^T = (Rj Ric)^T
    mat3_mul(RT, v.Ric, RjRic);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) A[3 * a + b] = RjRic[3 * b + a];
    // H = Ric^T hat(p_b)
    const double hat[9] = {0.0, -pb[2], pb[1], pb[2], 0.0, -pb[0], -pb[1], pb[0], 0.0};
    double Hm[9];
    mat3t_mul(v.Ric, hat, Hm);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.JX[c] = red[0] * A[c] + red[2] * A[6 + c];
        o.JX[3 + c] = red[4] * A[3 + c] + red[5] * A[6 + c];
        o.JT[c] = -o.JX[c];
        o.JT[6 + c] = -o.JX[3 + c];
        o.JT[3 + c] = red[0] * Hm[c] + red[2] * Hm[6 + c];
        o.JT[9 + c] = red[4] * Hm[3 + c] + red[5] * Hm[6 + c];
    }
}

// closed-form inverse of a symmetric 3x3 (xx xy xz yy yz zz) -> full row-major 3x3
VIO_HD void inv3_sym(const double s[6], double Ai[9]) {
    const double a = s[0], b = s[1], c = s[2], d = s[3], e = s[4], f = s[5];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    const double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = c01 * id; Ai[2] = c02 * id;
    Ai[3] = Ai[1]; Ai[4] = (a * f - c * c) * id; Ai[5] = (b * c - a * e) * id;
    Ai[6] = Ai[2]; Ai[7] = Ai[5]; Ai[8] = (a * d - b * b) * id;
}

template <bool WITH_SCHUR>
VIO_HD void linearize_point(const DevView &v, int l) {
    const int e0 = v.px_eptr[l], e1 = v.px_eptr[l + 1];
    const double X[3] = {v.pt[3 * (size_t)l], v.pt[3 * (size_t)l + 1], v.pt[3 * (size_t)l + 2]};
    double Hs[6] = {0, 0, 0, 0, 0, 0}, bs[3] = {0, 0, 0};
    const bool pfix = v.pt_fixed && v.pt_fixed[l];
    for (int e = e0; e < e1; ++e) {
        XyzEdge E;
        xyz_edge(v, X, e, E);
        if (pfix) {  // fixed point: no Jacobian block (MakeHessian skips fixed vertices)
#pragma unroll
            for (int c = 0; c < 6; ++c) E.JX[c] = 0.0;
        }
        double rho0, drho, W[3];
        robust_weights(v.rp_loss, v.rp_delta, v.rp_info, E.r, rho0, drho, W);
        const double dc = drho * v.rp_info;
        // W JX (2x3), W JT (2x6)
        double WX[6], WT[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) { WX[c] = W[0] * E.JX[c] + W[1] * E.JX[3 + c]; WX[3 + c] = W[1] * E.JX[c] + W[2] * E.JX[3 + c]; }
#pragma unroll
        for (int c = 0; c < 6; ++c) { WT[c] = W[0] * E.JT[c] + W[1] * E.JT[6 + c]; WT[6 + c] = W[1] * E.JT[c] + W[2] * E.JT[6 + c]; }
        Hs[0] += E.JX[0] * WX[0] + E.JX[3] * WX[3]; Hs[1] += E.JX[0] * WX[1] + E.JX[3] * WX[4]; Hs[2] += E.JX[0] * WX[2] + E.JX[3] * WX[5];
        Hs[3] += E.JX[1] * WX[1] + E.JX[4] * WX[4]; Hs[4] += E.JX[1] * WX[2] + E.JX[4] * WX[5]; Hs[5] += E.JX[2] * WX[2] + E.JX[5] * WX[5];
#pragma unroll
        for (int c = 0; c < 3; ++c) bs[c] -= dc * (E.JX[c] * E.r[0] + E.JX[3 + c] * E.r[1]);
        double *w = v.wx + 18 * (size_t)e;
        const int j = v.ex_pose[e];
        if (v.pose_fixed[j]) {
#pragma unroll
            for (int k = 0; k < 18; ++k) w[k] = 0.0;
            continue;
        }
        // H_lp row block (3x6) = JX^T W JT
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) w[6 * a + c] = E.JX[a] * WT[c] + E.JX[3 + a] * WT[6 + c];
        // (j,j) += JT^T W JT ; b_j -= drho c JT^T r
        double Xb[36];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) Xb[6 * a + c] = E.JT[a] * WT[c] + E.JT[6 + a] * WT[6 + c];
        s_add_diag(v, j, Xb, 1.0);
        double *hd = v.hdiag + v.pose_off[j], *bj = v.bp + v.pose_off[j];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vio_add(hd + k, Xb[7 * k]);
            vio_add(bj + k, -dc * (E.JT[k] * E.r[0] + E.JT[6 + k] * E.r[1]));
        }
    }
    double *Ho = v.Hxx + 6 * (size_t)l, *bo = v.bx + 3 * (size_t)l;
#pragma unroll
    for (int k = 0; k < 6; ++k) Ho[k] = Hs[k];
    bo[0] = bs[0]; bo[1] = bs[1]; bo[2] = bs[2];
    if (!WITH_SCHUR || e0 == e1 || pfix) return;
    // Schur complement of the 3x3 landmark block: S -= Hpl Hll^-1 Hlp ; bS -= Hpl Hll^-1 bl
    double Hi[9];
    inv3_sym(Hs, Hi);
    const double hb[3] = {Hi[0] * bs[0] + Hi[1] * bs[1] + Hi[2] * bs[2], Hi[3] * bs[0] + Hi[4] * bs[1] + Hi[5] * bs[2],
                          Hi[6] * bs[0] + Hi[7] * bs[1] + Hi[8] * bs[2]};
    for (int a = e0; a < e1; ++a) {
        const int pa = v.ex_pose[a];
        if (v.pose_fixed[pa]) continue;
        const double *wa = v.wx + 18 * (size_t)a;
        // T = Hpl_a Hll^-1 = wa^T Hi  (6x3)
        double T[18];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) T[3 * r + c] = wa[r] * Hi[c] + wa[6 + r] * Hi[3 + c] + wa[12 + r] * Hi[6 + c];
        double *bc = v.bcorr + v.pose_off[pa];
#pragma unroll
        for (int r = 0; r < 6; ++r) vio_add(bc + r, wa[r] * hb[0] + wa[6 + r] * hb[1] + wa[12 + r] * hb[2]);
        for (int b = a; b < e1; ++b) {
            const int pb = v.ex_pose[b];
            if (v.pose_fixed[pb]) continue;
            const double *wb = v.wx + 18 * (size_t)b;
            double Xb[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) Xb[6 * r + c] = T[3 * r] * wb[c] + T[3 * r + 1] * wb[6 + c] + T[3 * r + 2] * wb[12 + c];
            if (a == b) s_add_diag(v, pa, Xb, -1.0);
            else s_add_block(v, pa, pb, Xb, -1.0);
        }
    }
}

template <bool WITH_SCHUR>
__global__ void __launch_bounds__(128) k_linearize_xyz(DevView v) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < v.Lx) linearize_point<WITH_SCHUR>(v, l);
}

// chi2 partials over the points' edges
VIO_HD double chi2_point(const DevView &v, int l) {
    const double X[3] = {v.pt[3 * (size_t)l], v.pt[3 * (size_t)l + 1], v.pt[3 * (size_t)l + 2]};
    double chi = 0.0;
    for (int e = v.px_eptr[l]; e < v.px_eptr[l + 1]; ++e) {
        double pb[3], pc[3], r[2];
        xyz_residual(v, X, e, pb, pc, r);
        const double e2 = v.rp_info * (r[0] * r[0] + r[1] * r[1]);
        if (v.rp_loss == 0) chi += e2;
        else { double rho[3]; loss_compute(v.rp_loss, v.rp_delta, e2, rho); chi += rho[0]; }
    }
    return chi;
}
__global__ void __launch_bounds__(256) k_chi2_xyz(DevView v, double *partial) {
    double chi = 0.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < v.Lx; l += gridDim.x * blockDim.x) chi += chi2_point(v, l);
    block_sum_to(chi, partial);
}

// back-substitution dx_l = Hll^-1 (b_l - sum_j Hlp_j dx_j) and the LM scalars dx_l.(lambda dx_l + b_l), |dx_l|^2
VIO_HD void backsub_point(const DevView &v, int l, double lambda, double &sc, double &n2) {
    const int e0 = v.px_eptr[l], e1 = v.px_eptr[l + 1];
    double *dx = v.dxx + 3 * (size_t)l;
    if (e0 == e1 || (v.pt_fixed && v.pt_fixed[l])) { dx[0] = dx[1] = dx[2] = 0.0; return; }  // no edges, or a fixed point
    const double *b = v.bx + 3 * (size_t)l;
    double t[3] = {b[0], b[1], b[2]};
    for (int e = e0; e < e1; ++e) {
        const double *w = v.wx + 18 * (size_t)e;
        const double *dj = v.dxp + v.pose_off[v.ex_pose[e]];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) t[a] -= w[6 * a + c] * dj[c];
    }
    double Hi[9];
    inv3_sym(v.Hxx + 6 * (size_t)l, Hi);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double d = Hi[3 * a] * t[0] + Hi[3 * a + 1] * t[1] + Hi[3 * a + 2] * t[2];
        dx[a] = d;
        sc += d * (lambda * d + b[a]);
        n2 += d * d;
    }
}
__global__ void __launch_bounds__(256) k_backsub_xyz(DevView v, double lambda, double *part_scale, double *part_n2, const double *lam_p = nullptr) {
    if (lam_p) lambda = *lam_p;
    double sc = 0.0, n2 = 0.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < v.Lx; l += gridDim.x * blockDim.x) backsub_point(v, l, lambda, sc, n2);
    block_sum_to(sc, part_scale);
    block_sum_to(n2, part_n2);
}

// UpdateStates on the points (Vertex::Plus: x += delta) / v15 rollback Plus(-delta) / v17 restore
__global__ void k_update_xyz(DevView v, double sign, int backup) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3LL * v.Lx) return;
    if (backup) v.pt_bak[t] = v.pt[t];
    v.pt[t] += sign * v.dxx[t];
}
__global__ void k_restore_xyz(DevView v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3LL * v.Lx) v.pt[t] = v.pt_bak[t];
}
// max |diag| over the points' 3x3 blocks -> partial[blockIdx.x]
__global__ void __launch_bounds__(256) k_maxdiag_xyz(DevView v, double *partial) {
    __shared__ double sm[32];
    double m = 0.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < v.Lx; l += gridDim.x * blockDim.x) {
        const double *h = v.Hxx + 6 * (size_t)l;
        m = fmax(m, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    m = warp_max(m);
    if (lane == 0) sm[wid] = m;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        t = warp_max(t);
        if (lane == 0) partial[blockIdx.x] = t;
    }
}
