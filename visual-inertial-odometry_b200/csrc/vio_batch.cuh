// vio_batch.cuh — lock-step batch of structurally identical sliding windows (BASELINE config 3).
// All problems are concatenated into ONE packed graph (poses, landmarks, edges, IMU edges, priors); the per-edge
// kernels (k_linearize_grouped, k_imu_linearize, k_add_dense_prior, updates) run once over the whole batch, the
// reduced systems are `batch` stacked Pper x Pper blocks solved by one CTA each, and the reductions that feed the LM
// control (chi2, scale, |dx|^2, max diagonal) are done per problem by one CTA per problem in a fixed order.
// The LM control itself is the v17 loop of A17/src/backend/problem.cc:169-250 run per problem on the host.
#pragma once
#include "vio_dev.h"
#include "vio_kernels.cuh"
#include "vio_imu.cuh"
#include "vio_solvers.cuh"

// per-problem outputs: out[8*k + ...] = {chi_reproj, chi_other, scale_lm, n2_lm, scale_pose, n2_pose, maxdiag, -}
__global__ void __launch_bounds__(256) k_chi2_batch(DevView v, const int *lm_rng, double *out) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    double chi = 0.0;
    for (int l = lm_rng[k] + threadIdx.x; l < lm_rng[k + 1]; l += blockDim.x) {
        const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
        if (e0 == e1) continue;
        const double lam = v.invdep[l];
        const double *RTh = v.poseRT + 16 * (size_t)v.lm_host[l];
        const double pci[3] = {v.lm_pix[l] / lam, v.lm_piy[l] / lam, v.lm_piz[l] / lam};
        double pbi[3], pw[3];
        mat3_mul_vec(v.Ric, pci, pbi);
        pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
        mat3_mul_vec(RTh, pbi, pw);
        pw[0] += RTh[9]; pw[1] += RTh[10]; pw[2] += RTh[11];
        for (int e = e0; e < e1; ++e) {
            double pcj[3], pbj[3], r[2];
            reproj_residual(v.Ric, v.tic, v.poseRT + 16 * (size_t)v.e_pose_j[e], pw, v.e_pjx[e], v.e_pjy[e], pcj, pbj, r);
            const double e2 = v.rp_info * (r[0] * r[0] + r[1] * r[1]);
            if (v.rp_loss == 0) chi += e2;
            else { double rho[3]; loss_compute(v.rp_loss, v.rp_delta, e2, rho); chi += rho[0]; }
        }
    }
    const double tot = cta_sum(chi, red);
    if (threadIdx.x == 0) out[8 * k + 0] = tot;
}

__global__ void __launch_bounds__(320) k_other_chi2_batch(ImuView s, DevView v, const int *imu_rng, const double *err, int err_dim,
                                                          double *out) {
    __shared__ double r_s[10][15];
    __shared__ double chi_s[10];
    __shared__ double red[32];
    const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double total = 0.0;
    for (int e0 = imu_rng[k]; e0 < imu_rng[k + 1]; e0 += nw) {
        const int e = e0 + warp;
        if (e < imu_rng[k + 1]) {
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
            for (int q = 0; q < nw && e0 + q < imu_rng[k + 1]; ++q) total += chi_s[q];
        __syncthreads();
    }
    // + || err_prior ||   (norm, not squared)
    double nn = 0.0;
    if (err_dim > 0) {
        double t = 0.0;
        for (int i = threadIdx.x; i < err_dim; i += blockDim.x) { const double x = err[(size_t)k * err_dim + i]; t += x * x; }
        nn = cta_sum(t, red);
    }
    if (threadIdx.x == 0) out[8 * k + 1] = total + (err_dim > 0 ? sqrt(nn) : 0.0);
}

// back-substitution of problem k's landmarks + its LM scalars (lambda per problem)
__global__ void __launch_bounds__(256) k_backsub_batch(DevView v, const int *lm_rng, const double *lambdas, double *out) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    const double lambda = lambdas[k];
    double sc = 0.0, n2 = 0.0;
    for (int l = lm_rng[k] + threadIdx.x; l < lm_rng[k + 1]; l += blockDim.x) {
        const int e0 = v.lm_eptr[l], e1 = v.lm_eptr[l + 1];
        if (e0 == e1) { v.dxl[l] = 0.0; continue; }
        const double bl = v.bl[l];
        double t = bl;
        const double *wh = v.wh + 6 * (size_t)l;
        const double *dh = v.dxp + v.pose_off[v.lm_host[l]];
        for (int q = 0; q < 6; ++q) t -= wh[q] * dh[q];
        for (int e = e0; e < e1; ++e) {
            const double *w = v.wo + 6 * (size_t)e;
            const double *dj = v.dxp + v.pose_off[v.e_pose_j[e]];
            for (int q = 0; q < 6; ++q) t -= w[q] * dj[q];
        }
        const double d = (v.lm_fixed && v.lm_fixed[l]) ? 0.0 : t / v.Hll[l];  // fixed landmark: constant
        v.dxl[l] = d;
        sc += d * (lambda * d + bl);
        n2 += d * d;
    }
    const double a = cta_sum(sc, red), b = cta_sum(n2, red);
    double sp = 0.0, np = 0.0;
    for (int i = k * v.Pper + threadIdx.x; i < (k + 1) * v.Pper; i += blockDim.x) {
        const double d = v.dxp[i];
        sp += d * (lambda * d + v.bp[i]);
        np += d * d;
    }
    const double c = cta_sum(sp, red), dd = cta_sum(np, red);
    if (threadIdx.x == 0) { out[8 * k + 2] = a; out[8 * k + 3] = b; out[8 * k + 4] = c; out[8 * k + 5] = dd; }
}

__global__ void __launch_bounds__(256) k_maxdiag_batch(DevView v, const int *lm_rng, double *out) {
    __shared__ double sm[32];
    const int k = blockIdx.x;
    double m = 0.0;
    for (int i = k * v.Pper + threadIdx.x; i < (k + 1) * v.Pper; i += blockDim.x) m = fmax(m, fabs(v.hdiag[i]));
    for (int l = lm_rng[k] + threadIdx.x; l < lm_rng[k + 1]; l += blockDim.x) m = fmax(m, fabs(v.Hll[l]));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    m = warp_max(m);
    if (lane == 0) sm[wid] = m;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        t = warp_max(t);
        if (lane == 0) out[8 * k + 6] = t;
    }
}

// one CTA per problem: (S_k + lambda_k I) dx_k = bS_k, packed lower triangle in shared memory
__global__ void __launch_bounds__(512) k_chol_batch(const double *__restrict__ S, const double *__restrict__ b, const double *lambdas,
                                                     const uint8_t *act, int P, double *__restrict__ x) {
    extern __shared__ double Lm[];
    const int k = blockIdx.x;
    if (act && !act[k]) return;
    S += (size_t)k * P * P; b += (size_t)k * P; x += (size_t)k * P;
    const double lambda = lambdas[k];
    double *y = Lm + (size_t)P * (P + 1) / 2;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int i = warp; i < P; i += nw)
        for (int j = lane; j <= i; j += 32) Lm[tri_idx(i, j)] = S[(size_t)i * P + j] + (i == j ? lambda : 0.0);
    for (int i = tid; i < P; i += nt) y[i] = b[i];
    __syncthreads();
    for (int c = 0; c < P; ++c) {
        const double d = Lm[tri_idx(c, c)];
        const double dkk = sqrt(d);
        __syncthreads();
        if (tid == 0) Lm[tri_idx(c, c)] = dkk;
        for (int i = c + 1 + tid; i < P; i += nt) Lm[tri_idx(i, c)] /= dkk;
        __syncthreads();
        for (int i = c + 1 + warp; i < P; i += nw) {
            const double lik = Lm[tri_idx(i, c)];
            double *row = Lm + tri_idx(i, 0);
            for (int j = c + 1 + lane; j <= i; j += 32) row[j] -= lik * Lm[tri_idx(j, c)];
        }
        __syncthreads();
    }
    for (int c = 0; c < P; ++c) {
        const double yk = y[c] / Lm[tri_idx(c, c)];
        __syncthreads();
        if (tid == 0) y[c] = yk;
        for (int i = c + 1 + tid; i < P; i += nt) y[i] -= Lm[tri_idx(i, c)] * yk;
        __syncthreads();
    }
    for (int c = P - 1; c >= 0; --c) {
        const double xk = y[c] / Lm[tri_idx(c, c)];
        __syncthreads();
        if (tid == 0) y[c] = xk;
        for (int i = tid; i < c; i += nt) y[i] -= Lm[tri_idx(c, i)] * xk;
        __syncthreads();
    }
    for (int i = tid; i < P; i += nt) x[i] = y[i];
}

// UpdateStates / RollbackStates prior part per problem (A17/src/backend/problem.cc:465-474, 488-492)
__global__ void __launch_bounds__(512) k_prior_update_batch(const double *Hp, double *bp, double *bp_bak, double *err, double *err_bak,
                                                             const double *Jt, const double *dx, const uint8_t *act, int P, int err_dim) {
    const int k = blockIdx.x;
    if (act && !act[k]) return;
    Hp += (size_t)k * P * P; bp += (size_t)k * P; bp_bak += (size_t)k * P; dx += (size_t)k * P;
    err += (size_t)k * err_dim; err_bak += (size_t)k * err_dim; Jt += (size_t)k * err_dim * err_dim;
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
__global__ void k_prior_restore_batch(double *bp, const double *bp_bak, double *err, const double *err_bak, const uint8_t *act, int P,
                                      int err_dim) {
    const int k = blockIdx.x;
    if (!act[k]) return;
    for (int i = threadIdx.x; i < P; i += blockDim.x) bp[(size_t)k * P + i] = bp_bak[(size_t)k * P + i];
    for (int i = threadIdx.x; i < err_dim; i += blockDim.x) err[(size_t)k * err_dim + i] = err_bak[(size_t)k * err_dim + i];
}
