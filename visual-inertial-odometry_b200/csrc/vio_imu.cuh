// vio_imu.cuh — v17 pose-only factors: EdgeImu on IntegrationBase constants, and the dense
// marginalisation prior.  Reference (A17 = /root/reference/workspace/assignments/17-vins-initialization/vins-mono):
//   residual            IntegrationBase::evaluate           A17/include/factor/integration_base.h:160-186
//   Jacobians           EdgeImu::ComputeJacobians           A17/src/backend/edge_imu.cc:38-157
//   information         covariance.inverse()                A17/src/backend/edge_imu.cc:35
//   helpers             Utility::deltaQ/Qleft/Qright/skew   A17/include/utility/utility.h:11-64
//   prior in H, b       Problem::MakeHessian                A17/src/backend/problem.cc:365-384
//   prior update        Problem::UpdateStates               A17/src/backend/problem.cc:465-474
#pragma once
#include "vio_host.h"
#include "vio_dev.h"
#include "vio_kernels.cuh"
#include "../../include/vio_b200.h"

struct ImuBuffers {
    int n = 0;
    DBuf<int> pose_i, sb_i, pose_j, sb_j;
    DBuf<double> sum_dt, dp, dq, dv, lba, lbg, jac, cov, info;
    DBuf<int> blk_off, blk_dim;
    DBuf<uint8_t> blk_fixed;
    DBuf<uint8_t> row_fixed;
};

struct ImuView {
    int n;
    const int *pose_i, *sb_i, *pose_j, *sb_j;
    const double *sum_dt, *dp, *dq, *dv, *lba, *lbg, *jac, *info;
    double G[3];
};

// 15x15 inverse by Gauss-Jordan with partial pivoting (one thread per edge, once per graph)
__global__ void k_imu_info(int n, const double *cov, double *info) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double A[15][30];
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) {
            A[r][c] = cov[225 * (size_t)e + 15 * r + c];
            A[r][15 + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int k = 0; k < 15; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
        for (int r = k + 1; r < 15; ++r)
            if (fabs(A[r][k]) > best) { best = fabs(A[r][k]); piv = r; }
        if (piv != k)
            for (int c = 0; c < 30; ++c) { double t = A[k][c]; A[k][c] = A[piv][c]; A[piv][c] = t; }
        const double d = 1.0 / A[k][k];
        for (int c = 0; c < 30; ++c) A[k][c] *= d;
        for (int r = 0; r < 15; ++r) {
            if (r == k) continue;
            const double f = A[r][k];
            if (f != 0.0)
                for (int c = 0; c < 30; ++c) A[r][c] -= f * A[k][c];
        }
    }
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) info[225 * (size_t)e + 15 * r + c] = A[r][15 + c];
}

__device__ __forceinline__ void skew3(const double v[3], double S[9]) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
// bottom-right 3x3 of Qleft(q): w I + skew(vec)
__device__ __forceinline__ void qleft33(const double q[4], double M[9]) {
    skew3(q, M);
    M[0] += q[3]; M[4] += q[3]; M[8] += q[3];
}

// residual (15) and, if J != nullptr, the 15x30 Jacobian [pose_i(6) sb_i(9) pose_j(6) sb_j(9)] row-major
__device__ void imu_edge_eval(const ImuView &s, const DevView &v, int e, double r[15], double *J) {
    const double *pi = v.pose + 7 * (size_t)s.pose_i[e], *pj = v.pose + 7 * (size_t)s.pose_j[e];
    const double *si = v.sb + 9 * (size_t)s.sb_i[e], *sj = v.sb + 9 * (size_t)s.sb_j[e];
    const double Qi[4] = {pi[3], pi[4], pi[5], pi[6]}, Qj[4] = {pj[3], pj[4], pj[5], pj[6]};
    const double dt = s.sum_dt[e];
    const double *jac = s.jac + 225 * (size_t)e;
    double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            dp_dba[3 * a + b] = jac[15 * (0 + a) + 9 + b];
            dp_dbg[3 * a + b] = jac[15 * (0 + a) + 12 + b];
            dq_dbg[3 * a + b] = jac[15 * (3 + a) + 12 + b];
            dv_dba[3 * a + b] = jac[15 * (6 + a) + 9 + b];
            dv_dbg[3 * a + b] = jac[15 * (6 + a) + 12 + b];
        }
    const double dba[3] = {si[3] - s.lba[3 * e], si[4] - s.lba[3 * e + 1], si[5] - s.lba[3 * e + 2]};
    const double dbg[3] = {si[6] - s.lbg[3 * e], si[7] - s.lbg[3 * e + 1], si[8] - s.lbg[3 * e + 2]};
    double th[3];
    mat3_mul_vec(dq_dbg, dbg, th);
    const double dQ[4] = {th[0] / 2.0, th[1] / 2.0, th[2] / 2.0, 1.0};  // Utility::deltaQ (not normalised)
    const double *dq = s.dq + 4 * (size_t)e;
    double cdq[4];
    quat_mul(dq, dQ, cdq);
    double t1[3], t2[3], cdv[3], cdp[3];
    mat3_mul_vec(dv_dba, dba, t1); mat3_mul_vec(dv_dbg, dbg, t2);
    for (int k = 0; k < 3; ++k) cdv[k] = s.dv[3 * e + k] + t1[k] + t2[k];
    mat3_mul_vec(dp_dba, dba, t1); mat3_mul_vec(dp_dbg, dbg, t2);
    for (int k = 0; k < 3; ++k) cdp[k] = s.dp[3 * e + k] + t1[k] + t2[k];
    double Qii[4], Rii[9];
    quat_inv(Qi, Qii);
    quat_to_R(Qii, Rii);
    double a1[3], a2[3], u1[3], u2[3];
    for (int k = 0; k < 3; ++k) {
        a1[k] = 0.5 * s.G[k] * dt * dt + pj[k] - pi[k] - si[k] * dt;
        a2[k] = s.G[k] * dt + sj[k] - si[k];
    }
    mat3_mul_vec(Rii, a1, u1);
    mat3_mul_vec(Rii, a2, u2);
    double cdqi[4], QiiQj[4], qr[4];
    quat_inv(cdq, cdqi);
    quat_mul(Qii, Qj, QiiQj);
    quat_mul(cdqi, QiiQj, qr);
    for (int k = 0; k < 3; ++k) {
        r[k] = u1[k] - cdp[k];
        r[3 + k] = 2.0 * qr[k];
        r[6 + k] = u2[k] - cdv[k];
        r[9 + k] = sj[3 + k] - si[3 + k];
        r[12 + k] = sj[6 + k] - si[6 + k];
    }
    if (!J) return;
    for (int k = 0; k < 450; ++k) J[k] = 0.0;
    auto put = [&](int r0, int c0, const double *M, double sgn) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) J[30 * (r0 + a) + c0 + b] = sgn * M[3 * a + b];
    };
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double S1[9], S2[9];
    skew3(u1, S1);
    skew3(u2, S2);
    // pose_i: cols 0..5
    put(0, 0, Rii, -1.0);
    put(0, 3, S1, 1.0);
    {
        double Qji[4], QjiQi[4], L[9], Rr[9], T[9];
        quat_inv(Qj, Qji);
        quat_mul(Qji, Qi, QjiQi);
        // (Qleft(q) Qright(p)).bottomRight = -v_q v_p^T + (w_q I + skew v_q)(w_p I - skew v_p)
        qleft33(QjiQi, L);
        skew3(cdq, Rr);
        for (int k = 0; k < 9; ++k) Rr[k] = -Rr[k];
        Rr[0] += cdq[3]; Rr[4] += cdq[3]; Rr[8] += cdq[3];
        mat3_mul(L, Rr, T);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) T[3 * a + b] -= QjiQi[a] * cdq[b];
        put(3, 3, T, -1.0);
        // sb_i rotation/bg block: -Qleft(Qj^-1 Qi delta_q).bottomRight * dq_dbg   (uncorrected delta_q)
        double q3[4], L3[9], T3[9];
        quat_mul(QjiQi, dq, q3);
        qleft33(q3, L3);
        mat3_mul(L3, dq_dbg, T3);
        put(3, 6 + 6, T3, -1.0);
    }
    put(6, 3, S2, 1.0);
    // sb_i: cols 6..14  [v ba bg]
    {
        double Rdt[9];
        for (int k = 0; k < 9; ++k) Rdt[k] = Rii[k] * dt;
        put(0, 6 + 0, Rdt, -1.0);
        put(0, 6 + 3, dp_dba, -1.0);
        put(0, 6 + 6, dp_dbg, -1.0);
        put(6, 6 + 0, Rii, -1.0);
        put(6, 6 + 3, dv_dba, -1.0);
        put(6, 6 + 6, dv_dbg, -1.0);
        put(9, 6 + 3, I3, -1.0);
        put(12, 6 + 6, I3, -1.0);
    }
    // pose_j: cols 15..20
    put(0, 15, Rii, 1.0);
    {
        double L[9];
        qleft33(qr, L);  // qr = corrected_delta_q^-1 Qi^-1 Qj
        put(3, 15 + 3, L, 1.0);
    }
    // sb_j: cols 21..29
    put(6, 21 + 0, Rii, 1.0);
    put(9, 21 + 3, I3, 1.0);
    put(12, 21 + 6, I3, 1.0);
}

// one CTA per IMU edge: thread 0 evaluates r, J into shared memory; all threads form J^T Om J and J^T Om r
__global__ void __launch_bounds__(256) k_imu_linearize(ImuView s, DevView v) {
    __shared__ double J[450], OJ[450], r[15], Or[15];
    __shared__ int gidx[30];
    const int e = blockIdx.x;
    if (threadIdx.x == 0) {
        imu_edge_eval(s, v, e, r, J);
        const int offs[4] = {v.pose_off[s.pose_i[e]], v.sb_off[s.sb_i[e]], v.pose_off[s.pose_j[e]], v.sb_off[s.sb_j[e]]};
        const bool fx[4] = {v.pose_fixed[s.pose_i[e]] != 0, v.sb_fixed[s.sb_i[e]] != 0, v.pose_fixed[s.pose_j[e]] != 0,
                            v.sb_fixed[s.sb_j[e]] != 0};
        const int dims[4] = {6, 9, 6, 9};
        int c = 0;
        for (int k = 0; k < 4; ++k)
            for (int d = 0; d < dims[k]; ++d) gidx[c++] = fx[k] ? -1 : offs[k] + d;
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
        if (ga < 0 || gc < 0 || ga > gc) continue;
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += J[30 * k + a] * OJ[30 * k + c];
        atomicAdd(v.S + (size_t)ga * v.Pper + (gc % v.Pper), acc);
        if (ga == gc) atomicAdd(v.hdiag + ga, acc);
    }
    if (threadIdx.x < 30 && gidx[threadIdx.x] >= 0) {
        double acc = 0.0;
        for (int k = 0; k < 15; ++k) acc += J[30 * k + threadIdx.x] * Or[k];
        atomicAdd(v.bp + gidx[threadIdx.x], -acc);
    }
}

// one warp per IMU edge (lane 0 evaluates the residual, the lanes share the 15x15 quadratic form); per-edge values
// are summed by thread 0 in edge order, so the result does not depend on scheduling
__global__ void __launch_bounds__(1024) k_imu_chi2(ImuView s, DevView v, double *out /* accumulates */) {
    __shared__ double r_s[32][15];
    __shared__ double chi_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double total = 0.0;
    for (int e0 = 0; e0 < s.n; e0 += nw) {
        const int e = e0 + warp;
        if (e < s.n) {
            if (lane == 0) imu_edge_eval(s, v, e, r_s[warp], nullptr);
            __syncwarp();
            const double *Om = s.info + 225 * (size_t)e;
            double part = 0.0;
            for (int t = lane; t < 225; t += 32) part += r_s[warp][t / 15] * Om[t] * r_s[warp][t % 15];
            part = warp_sum(part);
            if (lane == 0) chi_s[warp] = part;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < nw && e0 + k < s.n; ++k) total += chi_s[k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += total;
}

inline ImuView imu_view(const ImuBuffers &b, const double G[3]) {
    ImuView s;
    s.n = b.n; s.pose_i = b.pose_i.p; s.sb_i = b.sb_i.p; s.pose_j = b.pose_j.p; s.sb_j = b.sb_j.p;
    s.sum_dt = b.sum_dt.p; s.dp = b.dp.p; s.dq = b.dq.p; s.dv = b.dv.p; s.lba = b.lba.p; s.lbg = b.lbg.p;
    s.jac = b.jac.p; s.info = b.info.p;
    s.G[0] = G[0]; s.G[1] = G[1]; s.G[2] = G[2];
    return s;
}

inline int imu_upload(ImuBuffers &b, const vio_graph *g, cudaStream_t st) {
    const int n = g->n_imu;
    b.n = n;
    for (int i = 0; i < n; ++i) {
        if (g->imu_pose_i[i] < 0 || g->imu_pose_i[i] >= g->n_pose || g->imu_pose_j[i] < 0 || g->imu_pose_j[i] >= g->n_pose ||
            g->imu_sb_i[i] < 0 || g->imu_sb_i[i] >= g->n_speedbias || g->imu_sb_j[i] < 0 || g->imu_sb_j[i] >= g->n_speedbias)
            return VIO_ERR_INVALID;
    }
    bool ok = true;
    ok &= upload(b.pose_i, g->imu_pose_i, (size_t)n, st) == cudaSuccess;
    ok &= upload(b.sb_i, g->imu_sb_i, (size_t)n, st) == cudaSuccess;
    ok &= upload(b.pose_j, g->imu_pose_j, (size_t)n, st) == cudaSuccess;
    ok &= upload(b.sb_j, g->imu_sb_j, (size_t)n, st) == cudaSuccess;
    ok &= upload(b.sum_dt, g->imu_sum_dt, (size_t)n, st) == cudaSuccess;
    ok &= upload(b.dp, g->imu_delta_p, 3 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.dq, g->imu_delta_q, 4 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.dv, g->imu_delta_v, 3 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.lba, g->imu_lin_ba, 3 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.lbg, g->imu_lin_bg, 3 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.jac, g->imu_jacobian, 225 * (size_t)n, st) == cudaSuccess;
    ok &= upload(b.cov, g->imu_covariance, 225 * (size_t)n, st) == cudaSuccess;
    ok &= b.info.alloc(225 * (size_t)n) == cudaSuccess;
    if (!ok) return VIO_ERR_CUDA;
    k_imu_info<<<(n + 31) / 32, 32, 0, st>>>(n, b.cov.p, b.info.p);
    return VIO_OK;
}

inline void imu_linearize(const ImuBuffers &b, const DevView &v, const double G[3], cudaStream_t st) {
    k_imu_linearize<<<b.n, 256, 0, st>>>(imu_view(b, G), v);
}
inline void imu_chi2(const ImuBuffers &b, const DevView &v, const double G[3], double *out, cudaStream_t st) {
    k_imu_chi2<<<1, 32 * (b.n < 32 ? (b.n > 0 ? b.n : 1) : 32), 0, st>>>(imu_view(b, G), v, out);
}

// ---- dense marginalisation prior ------------------------------------------------------------------
// Hessian_.topLeft += H_prior with rows/cols of fixed pose-class vertices zeroed; b_.head += b_prior (masked)
__global__ void k_add_dense_prior(DevView v, const double *Hp, const double *bp, const uint8_t *row_fixed) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)v.P * v.Pper;  // tall: P (total rows) x Pper
    if (t >= n) return;
    const int r = (int)(t / v.Pper), cl = (int)(t % v.Pper);
    const int c = (r / v.Pper) * v.Pper + cl;  // global index of the column inside the row's problem
    if (r > c || row_fixed[r] || row_fixed[c]) return;
    const double h = Hp[t];
    v.S[t] += h;  // exclusive element ownership within this kernel; edge kernels ran earlier on the stream
    if (r == c) {
        v.hdiag[r] += h;
        v.bp[r] += bp[r];
    }
}

// UpdateStates prior part: backup, b_prior -= H_prior dx_p, err_prior = -Jt_prior_inv b_prior.head(err_dim)
__global__ void __launch_bounds__(1024) k_prior_update(const double *Hp, double *bp, double *bp_bak, double *err,
                                                       double *err_bak, const double *Jt, const double *dx, int P,
                                                       int err_dim) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < P; i += blockDim.x) bp_bak[i] = bp[i];
    for (int i = threadIdx.x; i < err_dim; i += blockDim.x) err_bak[i] = err[i];
    __syncthreads();
    for (int i = warp; i < P; i += nw) {
        double t = 0.0;
        for (int j = lane; j < P; j += 32) t += Hp[(size_t)i * P + j] * dx[j];
        t = warp_sum(t);
        if (lane == 0) bp[i] -= t;
    }
    __syncthreads();
    for (int i = warp; i < err_dim; i += nw) {
        double t = 0.0;
        for (int j = lane; j < err_dim; j += 32) t += Jt[(size_t)i * err_dim + j] * bp[j];
        t = warp_sum(t);
        if (lane == 0) err[i] = -t;
    }
}

// out += ||v||_2   (err_prior_.norm() — norm, not squared: A17/src/backend/problem.cc:505-506)
__global__ void k_vec_norm_add(const double *x, int n, double *out) {
    __shared__ double sm[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t += x[i] * x[i];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += sqrt(sm[0]);
}
