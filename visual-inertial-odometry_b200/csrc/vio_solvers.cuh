// vio_solvers.cuh — reduced camera system solvers (FP64).
//   k_dense_chol_solve   (S + lambda I) x = b by Cholesky; replaces S.ldlt().solve  (A17/src/backend/problem.cc:434-440)
//   k_ref_pcg            Problem::PCGSolver restated verbatim, including the missing first x update
//                        (A15/backend/problem.cc:530-560)
//   k_bpcg_*             6x6 block-Jacobi PCG on the block-sparse reduced system (large BA)
#pragma once
#include "vio_dev.h"
#include "vio_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// dense Cholesky, one CTA.  A (P*P workspace) holds the lower triangle of S + lambda I.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_dense_chol_solve(const double *__restrict__ S, const double *__restrict__ b,
                                                            double lambda, int P, double *__restrict__ A,
                                                            double *__restrict__ x, int *info) {
    extern __shared__ double colk[];  // P doubles: current column
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (size_t idx = tid; idx < (size_t)P * P; idx += nt) {
        const int r = (int)(idx / P), c = (int)(idx % P);
        if (c <= r) A[idx] = S[idx] + (r == c ? lambda : 0.0);
    }
    if (tid == 0) *info = 0;
    __syncthreads();
    for (int k = 0; k < P; ++k) {
        const double d = A[(size_t)k * P + k];
        if (tid == 0 && !(d > 0.0)) *info = k + 1;
        const double dkk = sqrt(d);
        for (int i = k + tid; i < P; i += nt) {
            const double vv = (i == k) ? dkk : A[(size_t)i * P + k] / dkk;
            colk[i] = vv;
        }
        __syncthreads();
        for (int i = k + tid; i < P; i += nt) A[(size_t)i * P + k] = colk[i];
        for (int i = k + 1 + warp; i < P; i += nw) {
            const double aik = colk[i];
            double *row = A + (size_t)i * P;
            for (int j = k + 1 + lane; j <= i; j += 32) row[j] -= aik * colk[j];
        }
        __syncthreads();
    }
    // forward substitution L y = b (y in x)
    for (int i = tid; i < P; i += nt) x[i] = b[i];
    __syncthreads();
    for (int k = 0; k < P; ++k) {
        const double xk = x[k] / A[(size_t)k * P + k];
        __syncthreads();
        if (tid == 0) x[k] = xk;
        for (int i = k + 1 + tid; i < P; i += nt) x[i] -= A[(size_t)i * P + k] * xk;
        __syncthreads();
    }
    // backward substitution L^T z = y
    for (int k = P - 1; k >= 0; --k) {
        const double xk = x[k] / A[(size_t)k * P + k];
        __syncthreads();
        if (tid == 0) x[k] = xk;
        for (int i = tid; i < k; i += nt) x[i] -= A[(size_t)k * P + i] * xk;
        __syncthreads();
    }
}

// Same factorisation with the packed lower triangle resident in shared memory (P(P+1)/2 + P doubles <= 220 KB, i.e.
// P <= 234: the sliding-window sizes 120..171).  ~25x faster than the global-memory version at P = 171.
__device__ __forceinline__ int tri_idx(int i, int j) { return (i * (i + 1)) / 2 + j; }

__global__ void __launch_bounds__(512) k_dense_chol_smem(const double *__restrict__ S, const double *__restrict__ b,
                                                          double lambda, int P, double *__restrict__ x, int *info) {
    extern __shared__ double Lm[];  // packed lower triangle, then y[P]
    double *y = Lm + (size_t)P * (P + 1) / 2;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int i = warp; i < P; i += nw)
        for (int j = lane; j <= i; j += 32) Lm[tri_idx(i, j)] = S[(size_t)i * P + j] + (i == j ? lambda : 0.0);
    for (int i = tid; i < P; i += nt) y[i] = b[i];
    if (tid == 0) *info = 0;
    __syncthreads();
    for (int k = 0; k < P; ++k) {
        const double d = Lm[tri_idx(k, k)];
        const double dkk = sqrt(d);
        __syncthreads();  // everybody has read the pivot before it is overwritten
        if (tid == 0) {
            Lm[tri_idx(k, k)] = dkk;
            if (!(d > 0.0)) *info = k + 1;
        }
        for (int i = k + 1 + tid; i < P; i += nt) Lm[tri_idx(i, k)] /= dkk;
        __syncthreads();
        for (int i = k + 1 + warp; i < P; i += nw) {
            const double lik = Lm[tri_idx(i, k)];
            double *row = Lm + tri_idx(i, 0);
            for (int j = k + 1 + lane; j <= i; j += 32) row[j] -= lik * Lm[tri_idx(j, k)];
        }
        __syncthreads();
    }
    // L y = b
    for (int k = 0; k < P; ++k) {
        const double yk = y[k] / Lm[tri_idx(k, k)];
        __syncthreads();
        if (tid == 0) y[k] = yk;
        for (int i = k + 1 + tid; i < P; i += nt) y[i] -= Lm[tri_idx(i, k)] * yk;
        __syncthreads();
    }
    // L^T x = y
    for (int k = P - 1; k >= 0; --k) {
        const double xk = y[k] / Lm[tri_idx(k, k)];
        __syncthreads();
        if (tid == 0) y[k] = xk;
        for (int i = tid; i < k; i += nt) y[i] -= Lm[tri_idx(k, i)] * xk;
        __syncthreads();
    }
    for (int i = tid; i < P; i += nt) x[i] = y[i];
}

// ------------------------------------------------------------------------------------------------
// reference PCG (Jacobi preconditioner) on the dense reduced system, one CTA.
// Vectors live in dynamic shared memory: x, r, p, w, minv (5*P doubles).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double cta_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];  // every thread sums in the same order
    return t;
}

__global__ void __launch_bounds__(1024) k_ref_pcg(const double *__restrict__ S, const double *__restrict__ b,
                                                   double lambda, int P, int max_iter, double *__restrict__ xout,
                                                   int *iters_out) {
    extern __shared__ double sm[];
    double *x = sm, *r = sm + P, *p = sm + 2 * P, *w = sm + 3 * P, *minv = sm + 4 * P;
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    auto matvec = [&]() {  // w = (S + lambda I) p
        for (int i = warp; i < P; i += nw) {
            const double *row = S + (size_t)i * P;
            double t = 0.0;
            for (int j = lane; j < P; j += 32) t += row[j] * p[j];
            t = warp_sum(t);
            if (lane == 0) w[i] = t + lambda * p[i];
        }
        __syncthreads();
    };
    double l_r0z0 = 0.0, l_pw = 0.0, l_r0n = 0.0;
    for (int i = tid; i < P; i += nt) {
        x[i] = 0.0;
        const double mi = 1.0 / (S[(size_t)i * P + i] + lambda);
        minv[i] = mi;
        const double ri = b[i];
        r[i] = ri;
        const double zi = mi * ri;
        p[i] = zi;
        l_r0z0 += ri * zi;
        l_r0n += ri * ri;
    }
    __syncthreads();
    double r0z0 = cta_sum(l_r0z0, red);
    const double thr = 1e-6 * sqrt(cta_sum(l_r0n, red));
    matvec();
    for (int i = tid; i < P; i += nt) l_pw += p[i] * w[i];
    double alpha = r0z0 / cta_sum(l_pw, red);
    double l_rn = 0.0;
    for (int i = tid; i < P; i += nt) {
        r[i] -= alpha * w[i];  // r1 = r0 - alpha w ; NOTE: x is NOT updated here (reference defect)
        l_rn += r[i] * r[i];
    }
    double rn = sqrt(cta_sum(l_rn, red));
    int it = 0;
    while (rn > thr && it < max_iter) {
        ++it;
        double l_r1z1 = 0.0;
        for (int i = tid; i < P; i += nt) l_r1z1 += r[i] * (minv[i] * r[i]);
        const double r1z1 = cta_sum(l_r1z1, red);
        const double beta = r1z1 / r0z0;
        r0z0 = r1z1;
        for (int i = tid; i < P; i += nt) p[i] = beta * p[i] + minv[i] * r[i];
        __syncthreads();
        matvec();
        double l2 = 0.0;
        for (int i = tid; i < P; i += nt) l2 += p[i] * w[i];
        alpha = r1z1 / cta_sum(l2, red);
        double l3 = 0.0;
        for (int i = tid; i < P; i += nt) {
            x[i] += alpha * p[i];
            r[i] -= alpha * w[i];
            l3 += r[i] * r[i];
        }
        rn = sqrt(cta_sum(l3, red));
    }
    __syncthreads();
    for (int i = tid; i < P; i += nt) xout[i] = x[i];
    if (tid == 0) *iters_out = it;
}

// ------------------------------------------------------------------------------------------------
// block-Jacobi PCG on BSR (6x6 blocks).  Three kernels per iteration, no host round trip:
// every block re-sums the previous kernel's partials in a fixed order, so all blocks see bitwise
// identical scalars and the run is deterministic.
//   scal[0]=rz (parity 0) scal[1]=rz (parity 1) scal[2]=bnorm2 scal[3]=done scal[4]=iterations scal[5]=rr
// ------------------------------------------------------------------------------------------------
#define BPCG_MAXPART 1024

__device__ __forceinline__ double sum_partials_all(const double *part, int n, double *red) {
    // all threads of the block return the same, order-fixed sum
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t += part[i];
    return cta_sum(t, red);
}

__device__ __forceinline__ void inv6_spd(const double *A /*36 row-major*/, double *Ai) {
    // Cholesky based inverse of a 6x6 SPD block
    double L[36];
    for (int k = 0; k < 36; ++k) L[k] = 0.0;
    for (int j = 0; j < 6; ++j) {
        double d = A[7 * j];
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
        d = sqrt(d);
        L[7 * j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[6 * i + j];
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = s / d;
        }
    }
    // invert L (lower) -> Li
    double Li[36];
    for (int k = 0; k < 36; ++k) Li[k] = 0.0;
    for (int j = 0; j < 6; ++j) {
        Li[7 * j] = 1.0 / L[7 * j];
        for (int i = j + 1; i < 6; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s -= L[6 * i + k] * Li[6 * k + j];
            Li[6 * i + j] = s / L[7 * i];
        }
    }
    // Ai = Li^T Li
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
            for (int k = (r > c ? r : c); k < 6; ++k) s += Li[6 * k + r] * Li[6 * k + c];
            Ai[6 * r + c] = s;
        }
}

struct BpcgView {
    int nb;  // block rows
    const int *rowptr, *col, *diag;
    const double *val, *b;
    double *minv, *x, *r, *z, *p, *w;
    double *part_a, *part_b;  // BPCG_MAXPART each
    double *scal;
    double lambda, tol;
};

__global__ void __launch_bounds__(256) k_bpcg_init(BpcgView s) {
    __shared__ double red[32];
    double l_rz = 0.0, l_bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.nb; i += gridDim.x * blockDim.x) {
        double A[36], Ai[36];
        const double *d = s.val + 36 * (size_t)s.diag[i];
        for (int k = 0; k < 36; ++k) A[k] = d[k];
        for (int k = 0; k < 6; ++k) A[7 * k] += s.lambda;
        inv6_spd(A, Ai);
        double *mo = s.minv + 36 * (size_t)i;
        for (int k = 0; k < 36; ++k) mo[k] = Ai[k];
        double rb[6];
        for (int k = 0; k < 6; ++k) rb[k] = s.b[6 * (size_t)i + k];
        for (int r = 0; r < 6; ++r) {
            double z = 0.0;
            for (int c = 0; c < 6; ++c) z += Ai[6 * r + c] * rb[c];
            const size_t o = 6 * (size_t)i + r;
            s.x[o] = 0.0; s.r[o] = rb[r]; s.z[o] = z; s.p[o] = z;
            l_rz += rb[r] * z;
            l_bb += rb[r] * rb[r];
        }
    }
    const double a = cta_sum(l_rz, red), bb = cta_sum(l_bb, red);
    if (threadIdx.x == 0) { s.part_a[blockIdx.x] = a; s.part_b[blockIdx.x] = bb; }
}
__global__ void k_bpcg_init2(BpcgView s, int nparts) {
    __shared__ double red[32];
    const double rz = sum_partials_all(s.part_a, nparts, red);
    const double bb = sum_partials_all(s.part_b, nparts, red);
    if (threadIdx.x == 0) {
        s.scal[0] = rz; s.scal[1] = rz; s.scal[2] = bb; s.scal[4] = 0.0; s.scal[5] = bb;
        s.scal[3] = (bb == 0.0) ? 1.0 : 0.0;
    }
}

// w = (S + lambda I) p ; partial p.w     (6 threads per block row)
__global__ void __launch_bounds__(192) k_bpcg_spmv(BpcgView s) {
    __shared__ double red[32];
    if (s.scal[3] != 0.0) return;
    double pw = 0.0;
    const int n = 6 * s.nb;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int i = t / 6, rr = t % 6;
        double acc = 0.0;
        for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k) {
            const double *a = s.val + 36 * (size_t)k + 6 * rr;
            const double *pp = s.p + 6 * (size_t)s.col[k];
            acc += a[0] * pp[0] + a[1] * pp[1] + a[2] * pp[2] + a[3] * pp[3] + a[4] * pp[4] + a[5] * pp[5];
        }
        const double pi = s.p[t];
        acc += s.lambda * pi;
        s.w[t] = acc;
        pw += pi * acc;
    }
    const double a = cta_sum(pw, red);
    if (threadIdx.x == 0) s.part_a[blockIdx.x] = a;
}

// alpha = rz / p.w ; x += alpha p ; r -= alpha w ; z = Minv r ; partial r.z, r.r   (thread per block row)
__global__ void __launch_bounds__(128) k_bpcg_update(BpcgView s, int nparts_in, int par) {
    __shared__ double red[32];
    if (s.scal[3] != 0.0) return;
    const double pw = sum_partials_all(s.part_a, nparts_in, red);
    const double alpha = s.scal[par] / pw;
    double l_rz = 0.0, l_rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.nb; i += gridDim.x * blockDim.x) {
        double rn[6];
        for (int k = 0; k < 6; ++k) {
            const size_t o = 6 * (size_t)i + k;
            s.x[o] += alpha * s.p[o];
            rn[k] = s.r[o] - alpha * s.w[o];
            s.r[o] = rn[k];
            l_rr += rn[k] * rn[k];
        }
        const double *mi = s.minv + 36 * (size_t)i;
        for (int r = 0; r < 6; ++r) {
            double z = 0.0;
            for (int c = 0; c < 6; ++c) z += mi[6 * r + c] * rn[c];
            s.z[6 * (size_t)i + r] = z;
            l_rz += rn[r] * z;
        }
    }
    const double a = cta_sum(l_rz, red), b = cta_sum(l_rr, red);
    if (threadIdx.x == 0) { s.part_b[blockIdx.x] = a; s.part_a[BPCG_MAXPART + blockIdx.x] = b; }
}

// beta = rz' / rz ; p = z + beta p ; convergence test ; bookkeeping
__global__ void __launch_bounds__(256) k_bpcg_dir(BpcgView s, int nparts_in, int par, int max_iter) {
    __shared__ double red[32];
    if (s.scal[3] != 0.0) return;
    const double rz_new = sum_partials_all(s.part_b, nparts_in, red);
    const double rr = sum_partials_all(s.part_a + BPCG_MAXPART, nparts_in, red);
    const double beta = rz_new / s.scal[par];
    const bool done = !(sqrt(rr) > s.tol * sqrt(s.scal[2])) || (s.scal[4] + 1.0 >= (double)max_iter);
    if (!done) {
        const int n = 6 * s.nb;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            s.p[i] = s.z[i] + beta * s.p[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        s.scal[par ^ 1] = rz_new;
        s.scal[5] = rr;
        s.scal[6] = s.scal[4] + 1.0;  // staged; committed by k_bpcg_commit to avoid intra-kernel races
        s.scal[7] = done ? 1.0 : 0.0;
    }
}
__global__ void k_bpcg_commit(BpcgView s) {
    if (s.scal[3] != 0.0) return;
    s.scal[4] = s.scal[6];
    s.scal[3] = s.scal[7];
}

// ------------------------------------------------------------------------------------------------
// Persistent block-Jacobi PCG: ONE cooperative launch per reduced solve.  Every CTA owns a contiguous range
// of block rows; three software grid barriers per iteration (after S p, after the r/z update, after the
// direction update).  All CTAs re-sum the per-CTA partial dot products in the same fixed order, so alpha,
// beta and the convergence decision are bitwise identical everywhere (and across ranks).  The reduced
// system (60 MB at config 5) stays L2-resident between iterations.
//   bar[0] = arrival counter, bar[1] = generation   (zeroed by the host before the launch)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *vgen = bar + 1;
        const unsigned gen = *vgen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*vgen == gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ double sum_partials_cg(const double *part, int n, double *red) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t += __ldcg(part + i);
    return cta_sum(t, red);
}


// ------------------------------------------------------------------------------------------------
// Two-level preconditioner for the block PCG:  M^-1 = blockdiag(S + lambda I)^-1  +  Z Ac^-1 Z^T.
// The coarse space is built from AGGREGATES of consecutive pose blocks (each CTA of the persistent PCG kernel splits its
// own block rows into `apc` aggregates, so restriction and prolongation are CTA-local).  On every aggregate Z holds the
// seven infinitesimal similarity motions of the world frame restricted to the aggregate's cameras (translation, rotation
// and scale about the aggregate's centroid - the gauge freedoms of bundle adjustment, i.e. the near-null space of S):
//     d t_i = tau + omega x (p_i - c) + sigma (p_i - c),     d theta_i = R_i^T omega     (Plus is q <- q * Exp(d theta))
// Ac = Z^T (S + lambda I) Z is the (7 na)^2 Galerkin coarse matrix, inverted explicitly once per trial step.
// Block-Jacobi alone leaves the long-wavelength drift modes of a camera CHAIN (~25k iterations at 10k cameras); the
// coarse space removes them.  Everything is summed in a fixed order: ranks that solve the same reduced system
// redundantly stay bitwise equal.
// ------------------------------------------------------------------------------------------------
#define CZ_KD 7  // coarse degrees of freedom per aggregate
struct CoarseView {
    int apc;              // aggregates per PCG CTA; 0 = plain block-Jacobi
    int ma;               // block rows per aggregate (last aggregate of a CTA may hold fewer)
    int nc;               // coarse dimension = CZ_KD * grid * apc
    const double *Ainv;   // nc x nc
    const double *Z;      // [nb][6][CZ_KD]
    double *rc;           // [nc] restricted residual
};

// Z_i of every pose block: one CTA per aggregate (blocks agg_ptr[a] .. agg_ptr[a+1])
__global__ void __launch_bounds__(64) k_coarse_basis(const double *__restrict__ pose, const uint8_t *__restrict__ pose_fixed,
                                                      const int *__restrict__ blk_pose, const int *__restrict__ agg_ptr,
                                                      double *__restrict__ Z) {
    __shared__ double c_s[3];
    const int a = blockIdx.x, i0 = agg_ptr[a], i1 = agg_ptr[a + 1];
    if (threadIdx.x == 0) {
        double c[3] = {0, 0, 0};
        int n = 0;
        for (int i = i0; i < i1; ++i) {
            const int pi = blk_pose[i];
            if (pose_fixed[pi]) continue;
            c[0] += pose[7 * (size_t)pi]; c[1] += pose[7 * (size_t)pi + 1]; c[2] += pose[7 * (size_t)pi + 2];
            ++n;
        }
        const double inv = n > 0 ? 1.0 / n : 0.0;
        c_s[0] = c[0] * inv; c_s[1] = c[1] * inv; c_s[2] = c[2] * inv;
    }
    __syncthreads();
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int pi = blk_pose[i];
        double *z = Z + 6 * CZ_KD * (size_t)i;
        for (int k = 0; k < 6 * CZ_KD; ++k) z[k] = 0.0;
        if (pose_fixed[pi]) continue;
        const double *pp = pose + 7 * (size_t)pi;
        const double d[3] = {pp[0] - c_s[0], pp[1] - c_s[1], pp[2] - c_s[2]};
        double R[9];
        quat_to_R(pp + 3, R);
        // rows 0..2: [ I | -[d]x | d ]
        z[0 * CZ_KD + 0] = 1.0; z[1 * CZ_KD + 1] = 1.0; z[2 * CZ_KD + 2] = 1.0;
        z[0 * CZ_KD + 4] = d[2];  z[0 * CZ_KD + 5] = -d[1];
        z[1 * CZ_KD + 3] = -d[2]; z[1 * CZ_KD + 5] = d[0];
        z[2 * CZ_KD + 3] = d[1];  z[2 * CZ_KD + 4] = -d[0];
        z[0 * CZ_KD + 6] = d[0]; z[1 * CZ_KD + 6] = d[1]; z[2 * CZ_KD + 6] = d[2];
        // rows 3..5: [ 0 | R^T | 0 ]
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) z[(3 + r) * CZ_KD + 3 + c] = R[3 * c + r];
    }
}

// Galerkin assembly, one CTA per coarse block (a, b): Ac_ab = sum over the fine blocks (i, j) listed by the host (fixed
// order) of Z_i^T S_ij Z_j, plus lambda * sum_i Z_i^T Z_i on the diagonal blocks.  784 threads = 49 outputs x 16 lanes.
__global__ void __launch_bounds__(784) k_coarse_assemble(const double *__restrict__ val, const int *__restrict__ bsr_col,
                                                          const int *__restrict__ cb_ptr, const int *__restrict__ cb_fine,
                                                          const int *__restrict__ cb_frow, const int *__restrict__ cb_row,
                                                          const int *__restrict__ cb_col, const int *__restrict__ agg_ptr,
                                                          const double *__restrict__ Z, double lambda, int nc, double *__restrict__ Ac) {
    __shared__ double part[16][CZ_KD * CZ_KD];
    const int cb = blockIdx.x, e = threadIdx.x % (CZ_KD * CZ_KD), g = threadIdx.x / (CZ_KD * CZ_KD);
    const int r = e / CZ_KD, c = e % CZ_KD;
    double acc = 0.0;
    for (int q = cb_ptr[cb] + g; q < cb_ptr[cb + 1]; q += 16) {
        const int k = cb_fine[q], i = cb_frow[q], j = bsr_col[k];
        const double *V = val + 36 * (size_t)k, *Zi = Z + 6 * CZ_KD * (size_t)i, *Zj = Z + 6 * CZ_KD * (size_t)j;
        double t = 0.0;
#pragma unroll
        for (int x = 0; x < 6; ++x) {
            double u = 0.0;
#pragma unroll
            for (int y = 0; y < 6; ++y) u += __ldg(V + 6 * x + y) * __ldg(Zj + y * CZ_KD + c);
            t += __ldg(Zi + x * CZ_KD + r) * u;
        }
        acc += t;
    }
    part[g][e] = acc;
    __syncthreads();
    if (g == 0) {
        const int a = cb_row[cb], b = cb_col[cb];
        double t = 0.0;
        for (int q = 0; q < 16; ++q) t += part[q][e];
        if (a == b) {
            double zz = 0.0;
            for (int i = agg_ptr[a]; i < agg_ptr[a + 1]; ++i) {
                const double *Zi = Z + 6 * CZ_KD * (size_t)i;
                for (int x = 0; x < 6; ++x) zz += Zi[x * CZ_KD + r] * Zi[x * CZ_KD + c];
            }
            t += lambda * zz;
            if (r == c && !(t > 0.0)) t = 1.0;  // aggregate without free cameras: identity
        }
        Ac[(size_t)(CZ_KD * a + r) * nc + CZ_KD * b + c] = t;
    }
}

// In-place BLOCK Gauss-Jordan inverse of the SPD coarse matrix (7x7 pivot blocks, no pivoting: every pivot block is a
// Schur complement of an SPD matrix), cooperative launch.  CTA c keeps its `bpc` block rows (7*bpc scalar rows, all nc
// columns) in shared memory.  Step k: the owner inverts the pivot block, scales its block row and publishes it (7 x nc
// doubles, double buffered); one grid barrier; every CTA eliminates block column k from its rows, each thread taking
// whole columns (7 pivot values from L2, reused for all own rows).  na barriers instead of nc.
__device__ inline void inv7_spd(const double *A /* 7x7 row-major, ld */, int ld, double *Ai /* 7x7 dense */) {
    double L[49], Li[49];
    for (int i = 0; i < 49; ++i) { L[i] = 0.0; Li[i] = 0.0; }
    for (int j = 0; j < 7; ++j) {
        double d = A[j * ld + j];
        for (int k = 0; k < j; ++k) d -= L[7 * j + k] * L[7 * j + k];
        d = sqrt(d);
        L[7 * j + j] = d;
        for (int i = j + 1; i < 7; ++i) {
            double t = A[i * ld + j];
            for (int k = 0; k < j; ++k) t -= L[7 * i + k] * L[7 * j + k];
            L[7 * i + j] = t / d;
        }
    }
    for (int j = 0; j < 7; ++j) {  // Li = L^-1 (lower)
        Li[7 * j + j] = 1.0 / L[7 * j + j];
        for (int i = j + 1; i < 7; ++i) {
            double t = 0.0;
            for (int k = j; k < i; ++k) t -= L[7 * i + k] * Li[7 * k + j];
            Li[7 * i + j] = t / L[7 * i + i];
        }
    }
    for (int r = 0; r < 7; ++r)
        for (int c = 0; c < 7; ++c) {
            double t = 0.0;
            for (int k = (r > c ? r : c); k < 7; ++k) t += Li[7 * k + r] * Li[7 * k + c];
            Ai[7 * r + c] = t;
        }
}

// 7x7 SPD inverse by Gauss-Jordan in shared memory, all threads of the CTA call it (49 of them work)
__device__ __forceinline__ void inv7_cta(const double *A, int ld, double *M /* shared [49] out */) {
    const int tid = threadIdx.x, r = tid / CZ_KD, c = tid % CZ_KD;
    if (tid < CZ_KD * CZ_KD) M[tid] = A[(size_t)r * ld + c];
    __syncthreads();
    for (int k = 0; k < CZ_KD; ++k) {
        double d = 0.0, f = 0.0, pk = 0.0, cur = 0.0;
        if (tid < CZ_KD * CZ_KD) { d = 1.0 / M[CZ_KD * k + k]; f = M[CZ_KD * r + k]; pk = M[CZ_KD * k + c]; cur = M[tid]; }
        __syncthreads();
        if (tid < CZ_KD * CZ_KD) {
            double v;
            if (r == k) v = (c == k) ? d : pk * d;
            else v = (c == k) ? -f * d : cur - f * (pk * d);
            M[tid] = v;
        }
        __syncthreads();
    }
}

#define CZ_INV_THREADS 512
#define CZ_NCOL 4
// Flags instead of grid barriers: pivot block row k is published into its OWN slot of `pivbuf` ([na][7][nc], never
// reused within a launch, so there is no write-after-read hazard) and flag[k] = epoch is set with release semantics;
// consumers spin on flag[k].  The critical path is the owner chain (eliminate block row k+1 with pivot k, invert the
// 7x7 pivot, scale, publish); everybody else runs behind it without any all-to-all synchronisation.  Cooperative launch
// (co-residency) makes the spinning safe.
__device__ __forceinline__ void flag_wait(const unsigned *flag, unsigned epoch) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        } while (v != epoch);
    }
    __syncthreads();
}
__device__ __forceinline__ void flag_set(unsigned *flag, unsigned epoch) {
    __syncthreads();  // every thread's stores are done ...
    if (threadIdx.x == 0) {
        __threadfence();  // ... and visible device-wide before the flag
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
}

__global__ void __launch_bounds__(CZ_INV_THREADS, 1) k_coarse_invert(double *A, int nc, int bpc, double *pivbuf /* [na][7][nc] */,
                                                                     unsigned *flags /* [na] */, unsigned epoch,
                                                                     unsigned long long *prof /* [8] or nullptr */) {
    extern __shared__ double csm[];
    double *rows = csm;  // [7*bpc][nc]
    long long pt0 = 0;
#define GJ_MARK(slot_) do { if (prof && tid == 0) { const long long n_ = clock64(); atomicAdd(prof + (slot_), (unsigned long long)(n_ - pt0)); pt0 = n_; } } while (0)
    __shared__ __align__(16) double F[4 * CZ_KD][CZ_KD + 1];  // own rows' block column k before the update (bpc <= 4)
    __shared__ double Pm[CZ_KD * CZ_KD];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int na = nc / CZ_KD;
    // cyclic ownership: coarse block row k lives in CTA (k mod gridDim.x), local slot k / gridDim.x, so consecutive pivots
    // always sit on different CTAs and a CTA's own elimination work never delays the pivot it has to publish next
    const int nblk = gridDim.x, cta = blockIdx.x;
    const int nslot = cta < na ? (na - cta + nblk - 1) / nblk : 0;  // <= bpc
    const int nr = CZ_KD * nslot;
    (void)bpc;
    for (int t = tid; t < nr * nc; t += nt) {
        const int i = t / nc, j = t - i * nc;
        rows[t] = A[(size_t)(CZ_KD * (cta + (i / CZ_KD) * nblk) + i % CZ_KD) * nc + j];
    }
    __syncthreads();
    // pivot block row k <- [P A_k,: with P = A_kk^-1 | P inside the pivot block], written to shared memory and published
    auto publish_pivot = [&](int k) {
        double *pub = pivbuf + (size_t)k * CZ_KD * nc;
        const int kc = CZ_KD * k;
        double *rk = rows + (size_t)(CZ_KD * (k / nblk)) * nc;
        inv7_cta(rk + kc, nc, Pm);
        GJ_MARK(2);
        for (int j = tid; j < nc; j += nt) {
            double v[CZ_KD], o[CZ_KD];
#pragma unroll
            for (int m = 0; m < CZ_KD; ++m) v[m] = rk[(size_t)m * nc + j];
            const bool inpiv = j >= kc && j < kc + CZ_KD;
#pragma unroll
            for (int r = 0; r < CZ_KD; ++r) {
                double t = 0.0;
#pragma unroll
                for (int m = 0; m < CZ_KD; ++m) t += Pm[CZ_KD * r + m] * v[m];
                o[r] = inpiv ? Pm[CZ_KD * r + (j - kc)] : t;
            }
#pragma unroll
            for (int r = 0; r < CZ_KD; ++r) { rk[(size_t)r * nc + j] = o[r]; __stcg(pub + (size_t)r * nc + j, o[r]); }
        }
        __syncthreads();
        GJ_MARK(3);
        flag_set(flags + k, epoch);
    };
    // elimination of block column k from own scalar rows [i_lo, i_hi) (the pivot block row itself is skipped).  Each
    // thread takes CZ_NCOL columns at a time: the 7 multipliers of a row (three 16-byte shared loads + one) are reused
    // for all of them, which keeps the shared-memory traffic below the FP64 work.
    auto eliminate = [&](int k, int i_lo, int i_hi) {
        const double *pub = pivbuf + (size_t)k * CZ_KD * nc;
        const int kc = CZ_KD * k;
        for (int jb = 0; jb < nc; jb += CZ_NCOL * nt) {
            double pv[CZ_NCOL][CZ_KD];
            int jj[CZ_NCOL];
#pragma unroll
            for (int c = 0; c < CZ_NCOL; ++c) {
                jj[c] = jb + c * nt + tid;
#pragma unroll
                for (int m = 0; m < CZ_KD; ++m) pv[c][m] = jj[c] < nc ? __ldcg(pub + (size_t)m * nc + jj[c]) : 0.0;
            }
            for (int i = i_lo; i < i_hi; ++i) {
                if (cta + (i / CZ_KD) * nblk == k) continue;
                const double2 f01 = *reinterpret_cast<const double2 *>(&F[i][0]), f23 = *reinterpret_cast<const double2 *>(&F[i][2]),
                              f45 = *reinterpret_cast<const double2 *>(&F[i][4]);
                const double f6 = F[i][6];
#pragma unroll
                for (int c = 0; c < CZ_NCOL; ++c) {
                    if (jj[c] >= nc) continue;
                    const double t = f01.x * pv[c][0] + f01.y * pv[c][1] + f23.x * pv[c][2] + f23.y * pv[c][3] + f45.x * pv[c][4] +
                                     f45.y * pv[c][5] + f6 * pv[c][6];
                    double *e = rows + (size_t)i * nc + jj[c];
                    const bool inpiv = jj[c] >= kc && jj[c] < kc + CZ_KD;
                    *e = inpiv ? -t : *e - t;  // inside the pivot block pv holds P:  A_ik <- -F P
                }
            }
        }
    };
    if (prof && tid == 0) pt0 = clock64();
    if (cta == 0 && nslot > 0) publish_pivot(0);
    for (int k = 0; k < na; ++k) {
        const int kc = CZ_KD * k;
        const bool own_k = (k % nblk) == cta;
        const bool own_next = (k + 1 < na) && ((k + 1) % nblk) == cta;
        if (own_next && prof && tid == 0) pt0 = clock64();
        if (!own_k) flag_wait(flags + k, epoch);  // pivot block row k is published (the owner has it already)
        if (own_next) GJ_MARK(0);
        for (int t = tid; t < nr * CZ_KD; t += nt) F[t / CZ_KD][t % CZ_KD] = rows[(size_t)(t / CZ_KD) * nc + kc + t % CZ_KD];
        __syncthreads();
        if (own_next) {
            // the owner chain: bring block row k+1 up to date first and publish the next pivot, then the other rows
            const int lo = CZ_KD * ((k + 1) / nblk);
            eliminate(k, lo, lo + CZ_KD);
            __syncthreads();
            GJ_MARK(1);
            publish_pivot(k + 1);
            GJ_MARK(4);
            eliminate(k, 0, lo);
            eliminate(k, lo + CZ_KD, nr);
        } else {
            eliminate(k, 0, nr);
        }
        __syncthreads();
    }
    for (int t = tid; t < nr * nc; t += nt) {
        const int i = t / nc, j = t - i * nc;
        A[(size_t)(CZ_KD * (cta + (i / CZ_KD) * nblk) + i % CZ_KD) * nc + j] = rows[t];
    }
}

#define BPCG_P_THREADS 1024
#define BPCG_P_GROUP 8  // lanes cooperating on one scalar row of S p

// per-CTA tables built once per graph by the host (do_solve_step): the block columns a CTA's rows touch
// ("window") and, for every stored block, its index inside the owning CTA's window.
struct BpcgTables {
    int br;                 // block rows per CTA
    int win_max;            // largest window (block columns)
    const int *cta_colptr;  // [grid+1]
    const int *cta_cols;    // window block columns, per CTA
    const int *lcol;        // [nnzb] local window index of block k
};

// fixed-order sum of n partials by warp 0, broadcast to the CTA (identical bits on every CTA and rank)
__device__ __forceinline__ double sum_partials_w0(const double *part, int n, double *bc) {
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int i = threadIdx.x; i < n; i += 32) t += __ldcg(part + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) *bc = t;
    }
    __syncthreads();
    return *bc;
}

__global__ void __launch_bounds__(BPCG_P_THREADS, 1) k_bpcg_persistent(BpcgView s, BpcgTables tb, int max_iter, unsigned *bar,
                                                                       double *pbuf2, int n_init_parts, CoarseView cv) {
    extern __shared__ double dsm[];
    __shared__ double red[32];
    __shared__ double bc[2];
    const int nblk = gridDim.x, tid = threadIdx.x, nt = blockDim.x;
    const int br = tb.br;
    const int i0 = min(s.nb, blockIdx.x * br), i1 = min(s.nb, i0 + br);
    const int r0 = 6 * i0, nrow = 6 * (i1 - i0);
    const int c0 = tb.cta_colptr[blockIdx.x], nwin = tb.cta_colptr[blockIdx.x + 1] - c0;
    double *pw = dsm;                           // [win_max*6]  direction vector on the CTA's column window
    double *w_s = pw + 6 * (size_t)tb.win_max;  // [br*6] each: S p, r, z, p, x of the CTA's own rows
    double *r_s = w_s + 6 * (size_t)br;
    double *z_s = r_s + 6 * (size_t)br;
    double *p_s = z_s + 6 * (size_t)br;
    double *x_s = p_s + 6 * (size_t)br;
    double *mi_s = x_s + 6 * (size_t)br;        // [br*36] block-Jacobi preconditioner of the own rows
    double *part_a = s.part_a, *part_b = s.part_b, *part_c = s.part_a + BPCG_MAXPART;
    double *pb[2] = {s.p, pbuf2};  // direction vector in HBM/L2, double buffered (neighbours read the previous one)
    // init (Minv, r = b, z = p = Minv r, partial r.z and b.b) was done by k_bpcg_init with n_init_parts CTAs
    for (int t = tid; t < nrow; t += nt) {
        r_s[t] = s.r[r0 + t];
        z_s[t] = s.z[r0 + t];
        p_s[t] = 0.0;
        x_s[t] = 0.0;
    }
    for (int t = tid; t < 36 * (i1 - i0); t += nt) mi_s[t] = s.minv[36 * (size_t)i0 + t];
    double rz = sum_partials_w0(part_a, n_init_parts, bc);
    const double bb = sum_partials_w0(part_b, n_init_parts, bc + 1);
    double rr = bb, beta = 0.0;
    int it = 0, cur = 0;
    const double thr = s.tol * sqrt(bb);
    grid_barrier(bar, nblk);  // init partials consumed everywhere before the buffers are reused
    // ---- coarse correction z += Z Ac^-1 Z^T r on the CTA's own rows; see CoarseView.  Every CTA keeps the whole restricted
    // residual rc = Z^T r in shared memory and updates it by the CG recurrence rc -= alpha Z^T (S p), whose pieces are
    // exchanged on the iteration's first barrier: the coarse term costs no barrier of its own.
    __shared__ double zc_w[32][28];  // [warp][coarse row of this CTA] partial sums
    __shared__ double zc_f[32];
    const int crow_n = CZ_KD * cv.apc;  // coarse rows owned by this CTA (<= 32)
    double *Z_s = mi_s + 36 * (size_t)br;  // [br][6][CZ_KD] (only allocated when cv.apc > 0)
    if (cv.apc > 0)
        for (int t = tid; t < 6 * CZ_KD * (i1 - i0); t += nt) Z_s[t] = cv.Z[6 * CZ_KD * (size_t)i0 + t];
    double *rc_s = Z_s + 6 * CZ_KD * (size_t)br;  // [nc] the whole restricted residual Z^T r, kept by every CTA
    // own aggregates' part of Z^T v (v = r at the start, S p inside the iteration) -> global exchange buffer
    auto restrict_publish = [&](const double *v_s) {
        const int nbl = i1 - i0;
        // 8 lanes per coarse row (block rows strided over them), combined by a fixed-order shuffle tree
        if (tid < 8 * crow_n) {  // crow_n <= 28: at most 7 warps, every warp fully populated or tail-guarded below
            const int cr = tid >> 3, g8 = tid & 7;
            const int al = cr / CZ_KD, m = cr % CZ_KD;
            const int b0 = min(nbl, al * cv.ma), b1 = min(nbl, b0 + cv.ma);
            double t = 0.0;
            for (int ib = b0 + g8; ib < b1; ib += 8) {
                const double *zi = Z_s + 6 * CZ_KD * ib + m, *vb = v_s + 6 * ib;
                t += zi[0] * vb[0] + zi[CZ_KD] * vb[1] + zi[2 * CZ_KD] * vb[2] + zi[3 * CZ_KD] * vb[3] + zi[4 * CZ_KD] * vb[4] +
                     zi[5 * CZ_KD] * vb[5];
            }
            const unsigned grp = 0xffu << ((tid & 31) & ~7);
            t += __shfl_xor_sync(grp, t, 1);
            t += __shfl_xor_sync(grp, t, 2);
            t += __shfl_xor_sync(grp, t, 4);
            if (g8 == 0) __stcg(cv.rc + (size_t)crow_n * blockIdx.x + cr, t);
        }
    };
    // z_s += Z (Ainv[own rows, :] rc_s): every thread takes whole columns (one Ainv element per own coarse row, all loads
    // independent), then a fixed-order reduction over lanes and warps
    auto coarse_apply = [&]() {
        const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
        for (int al = 0; al < cv.apc; ++al) {  // one aggregate (CZ_KD coarse rows) at a time: 7 accumulators in registers
            double acc[CZ_KD];
#pragma unroll
            for (int q = 0; q < CZ_KD; ++q) acc[q] = 0.0;
            const double *Ar = cv.Ainv + (size_t)(crow_n * blockIdx.x + CZ_KD * al) * cv.nc;
            for (int j = tid; j < cv.nc; j += nt) {
                const double rj = rc_s[j];
#pragma unroll
                for (int q = 0; q < CZ_KD; ++q) acc[q] += __ldcs(Ar + (size_t)q * cv.nc + j) * rj;  // evict-first: keep S in L2
            }
#pragma unroll
            for (int q = 0; q < CZ_KD; ++q) {
                double t = acc[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) zc_w[warp][CZ_KD * al + q] = t;
            }
        }
        __syncthreads();
        // combine the per-warp partials: warp cr sums the nwarp (<= 32) partials of coarse row cr in a fixed shuffle order
        if (warp < crow_n) {
            double t = lane < nwarp ? zc_w[lane][warp] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) zc_f[warp] = t;
        }
        __syncthreads();
        for (int t = tid; t < nrow; t += nt) {
            const int ib = t / 6, x = t % 6;
            const double *zi = Z_s + 6 * CZ_KD * ib + x * CZ_KD, *zc = zc_f + CZ_KD * min(cv.apc - 1, ib / cv.ma);
            double u = 0.0;
#pragma unroll
            for (int m = 0; m < CZ_KD; ++m) u += zi[m] * zc[m];
            z_s[t] += u;
        }
        __syncthreads();
    };
    bool use_coarse = cv.apc > 0;
    if (use_coarse && bb > 0.0) {
        // the init kernel applied block-Jacobi only: rc = Z^T r, add the coarse term to z and recompute r.z
        __syncthreads();
        restrict_publish(r_s);
        grid_barrier(bar, nblk);
        for (int j = tid; j < cv.nc; j += nt) rc_s[j] = __ldcg(cv.rc + j);
        __syncthreads();
        coarse_apply();
        double l = 0.0;
        for (int t = tid; t < nrow; t += nt) { l += r_s[t] * z_s[t]; s.z[r0 + t] = z_s[t]; }
        const double a = cta_sum(l, red);
        if (tid == 0) part_b[blockIdx.x] = a;
        grid_barrier(bar, nblk);
        rz = sum_partials_w0(part_b, nblk, bc);
        grid_barrier(bar, nblk);  // part_b and the exchange buffer are rewritten in the first iteration
        if (!(rz > 0.0) || !(rz < 1e300)) {
            // the coarse inverse broke down (a pivot block lost definiteness to rounding): this solve falls back to
            // block-Jacobi.  rz is bitwise the same on every CTA, so the decision is uniform.
            use_coarse = false;
            double l2 = 0.0;
            for (int t = tid; t < nrow; t += nt) {
                const int ib = t / 6, rw = t % 6;
                const double *mi = mi_s + 36 * (size_t)ib + 6 * rw;
                const double *rb = r_s + 6 * (size_t)ib;
                const double z = mi[0] * rb[0] + mi[1] * rb[1] + mi[2] * rb[2] + mi[3] * rb[3] + mi[4] * rb[4] + mi[5] * rb[5];
                z_s[t] = z;
                s.z[r0 + t] = z;
                l2 += rb[rw] * z;
            }
            const double a2 = cta_sum(l2, red);
            if (tid == 0) part_b[blockIdx.x] = a2;
            grid_barrier(bar, nblk);
            rz = sum_partials_w0(part_b, nblk, bc);
            grid_barrier(bar, nblk);
        }
    }
    const int G = BPCG_P_GROUP;
    const int sub = tid % G;
    long long tp0 = 0, c_spmv = 0, c_bar1 = 0, c_upd = 0, c_bar2 = 0;
#define PCG_MARK(acc_) do { if (blockIdx.x == 0 && tid == 0) { const long long n_ = clock64(); acc_ += n_ - tp0; tp0 = n_; } } while (0)
    if (blockIdx.x == 0 && tid == 0) tp0 = clock64();
    if (bb > 0.0) {
        while (it < max_iter) {
            const double *pold = pb[cur];
            double *pnew = pb[cur ^ 1];
            // ---- gather p_new = z + beta p_old on the column window (beta = 0 in the first iteration)
            for (int t = tid; t < 6 * nwin; t += nt) {
                const int j = tb.cta_cols[c0 + t / 6], c = t % 6;
                pw[t] = __ldcg(s.z + 6 * (size_t)j + c) + beta * __ldcg(pold + 6 * (size_t)j + c);
            }
            for (int t = tid; t < nrow; t += nt) {
                const double pi = z_s[t] + beta * p_s[t];
                p_s[t] = pi;
                pnew[r0 + t] = pi;
            }
            __syncthreads();
            // ---- w = (S + lambda I) p on own rows: G lanes per scalar row, blocks strided over the lanes
            for (int t0 = 0; t0 < nrow; t0 += nt / G) {
                const int t = t0 + tid / G;
                double acc = 0.0;
                if (t < nrow) {
                    const int i = i0 + t / 6, rw = t % 6;
                    const int k1 = s.rowptr[i + 1];
                    int k = s.rowptr[i] + sub;
                    // up to three blocks per lane in flight (rows of <= 24 blocks are covered in one batch)
                    for (; k < k1; k += 3 * G) {
                        const int ka = k, kb = k + G, kc = k + 2 * G;
                        const bool hb = kb < k1, hc = kc < k1;
                        const double2 *A = reinterpret_cast<const double2 *>(s.val + 36 * (size_t)ka + 6 * rw);
                        const double2 *B = reinterpret_cast<const double2 *>(s.val + 36 * (size_t)(hb ? kb : ka) + 6 * rw);
                        const double2 *Cc = reinterpret_cast<const double2 *>(s.val + 36 * (size_t)(hc ? kc : ka) + 6 * rw);
                        const double2 a0 = __ldg(A), a1 = __ldg(A + 1), a2 = __ldg(A + 2);
                        const double2 b0 = __ldg(B), b1 = __ldg(B + 1), b2 = __ldg(B + 2);
                        const double2 d0 = __ldg(Cc), d1 = __ldg(Cc + 1), d2 = __ldg(Cc + 2);
                        const int la = __ldg(tb.lcol + ka), lb = __ldg(tb.lcol + (hb ? kb : ka)), lc = __ldg(tb.lcol + (hc ? kc : ka));
                        const double *pa = pw + 6 * la, *pbb = pw + 6 * lb, *pc = pw + 6 * lc;
                        acc += a0.x * pa[0] + a0.y * pa[1] + a1.x * pa[2] + a1.y * pa[3] + a2.x * pa[4] + a2.y * pa[5];
                        if (hb) acc += b0.x * pbb[0] + b0.y * pbb[1] + b1.x * pbb[2] + b1.y * pbb[3] + b2.x * pbb[4] + b2.y * pbb[5];
                        if (hc) acc += d0.x * pc[0] + d0.y * pc[1] + d1.x * pc[2] + d1.y * pc[3] + d2.x * pc[4] + d2.y * pc[5];
                    }
                }
#pragma unroll
                for (int m = 1; m < G; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                if (t < nrow && sub == 0) w_s[t] = acc + s.lambda * p_s[t];
            }
            __syncthreads();
            double l_pw = 0.0;
            for (int t = tid; t < nrow; t += nt) l_pw += p_s[t] * w_s[t];
            {
                const double a = cta_sum(l_pw, red);
                if (tid == 0) part_a[blockIdx.x] = a;
            }
            if (use_coarse) restrict_publish(w_s);  // Z^T (S p) rides on the same barrier: rc -= alpha Z^T S p below
            PCG_MARK(c_spmv);
            grid_barrier(bar, nblk);
            const double alpha = rz / sum_partials_w0(part_a, nblk, bc);
            PCG_MARK(c_bar1);
            if (use_coarse)
                for (int j = tid; j < cv.nc; j += nt) rc_s[j] -= alpha * __ldcg(cv.rc + j);
            // ---- x += alpha p ; r -= alpha w ; z = Minv r ; partial r.z, r.r   (own rows, all in shared memory)
            for (int t = tid; t < nrow; t += nt) {
                x_s[t] += alpha * p_s[t];
                r_s[t] -= alpha * w_s[t];
            }
            __syncthreads();
            double l_rz2 = 0.0, l_rr = 0.0;
            if (!use_coarse) {
                for (int t = tid; t < nrow; t += nt) {
                    const int ib = t / 6, rw = t % 6;
                    const double *mi = mi_s + 36 * (size_t)ib + 6 * rw;
                    const double *rb = r_s + 6 * (size_t)ib;
                    const double z = mi[0] * rb[0] + mi[1] * rb[1] + mi[2] * rb[2] + mi[3] * rb[3] + mi[4] * rb[4] + mi[5] * rb[5];
                    z_s[t] = z;
                    s.z[r0 + t] = z;
                    l_rz2 += rb[rw] * z;
                    l_rr += rb[rw] * rb[rw];
                }
            } else {
                for (int t = tid; t < nrow; t += nt) {
                    const int ib = t / 6, rw = t % 6;
                    const double *mi = mi_s + 36 * (size_t)ib + 6 * rw;
                    const double *rb = r_s + 6 * (size_t)ib;
                    z_s[t] = mi[0] * rb[0] + mi[1] * rb[1] + mi[2] * rb[2] + mi[3] * rb[3] + mi[4] * rb[4] + mi[5] * rb[5];
                }
                __syncthreads();
                coarse_apply();
                for (int t = tid; t < nrow; t += nt) {
                    const double z = z_s[t], r = r_s[t];
                    s.z[r0 + t] = z;
                    l_rz2 += r * z;
                    l_rr += r * r;
                }
            }
            {
                const double a = cta_sum(l_rz2, red), b = cta_sum(l_rr, red);
                if (tid == 0) { part_b[blockIdx.x] = a; part_c[blockIdx.x] = b; }
            }
            PCG_MARK(c_upd);
            grid_barrier(bar, nblk);
            const double rz_new = sum_partials_w0(part_b, nblk, bc);
            rr = sum_partials_w0(part_c, nblk, bc + 1);
            PCG_MARK(c_bar2);
            ++it;
            cur ^= 1;
            if (!(sqrt(rr) > thr)) break;  // identical on every CTA
            beta = rz_new / rz;
            rz = rz_new;
        }
    }
    for (int t = tid; t < nrow; t += nt) s.x[r0 + t] = x_s[t];
    if (blockIdx.x == 0 && tid == 0) {
        s.scal[4] = (double)it;
        s.scal[5] = rr;
        s.scal[3] = 1.0;
        s.scal[2] = bb;
        s.scal[8] = (double)c_spmv; s.scal[9] = (double)c_bar1; s.scal[10] = (double)c_upd; s.scal[11] = (double)c_bar2;
    }
#undef PCG_MARK
}

// ------------------------------------------------------------------------------------------------
// GENERIC_PROBLEM lane: dense H = J^T W J, b = -J^T Wb r from host-evaluated factors
// ------------------------------------------------------------------------------------------------
struct DenseSysView {
    int n, rows, dmax;
    const double *J, *r, *W, *Wb;
    const int *row_edge0, *row_dim;
};
// one thread per (a, c) of H (and per a of b): fixed summation order over the stacked rows
__global__ void k_dense_accumulate(DenseSysView s, double *H, double *b) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = s.n;
    if (t >= n * (n + 1)) return;
    const int a = t / (n + 1), c = t % (n + 1);
    double acc = 0.0;
    for (int i = 0; i < s.rows; ++i) {
        const double ja = s.J[(size_t)i * n + a];
        if (ja == 0.0) continue;
        const int e0 = s.row_edge0[i], d = s.row_dim[i];
        double wj = 0.0;
        if (c < n) {
            for (int k = 0; k < d; ++k) wj += s.W[(size_t)i * s.dmax + k] * s.J[(size_t)(e0 + k) * n + c];
        } else {
            for (int k = 0; k < d; ++k) wj += s.Wb[(size_t)i * s.dmax + k] * s.r[e0 + k];
        }
        acc += ja * wj;
    }
    if (c < n) H[(size_t)a * n + c] = acc;
    else b[a] = -acc;
}
__global__ void k_dense_chi2(int rows, int dmax, const double *r, const int *row_edge0, const int *row_dim,
                             const double *info, const int *loss_kind, const double *loss_delta, double *out) {
    // single CTA; per-edge e2 = r^T Info r accumulated by the edge's first row, fixed order
    __shared__ double sm[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        if (row_edge0[i] != i) continue;
        const int d = row_dim[i];
        double e2 = 0.0;
        for (int a = 0; a < d; ++a) {
            double w = 0.0;
            for (int k = 0; k < d; ++k) w += info[(size_t)(i + a) * dmax + k] * r[i + k];
            e2 += r[i + a] * w;
        }
        double rho[3];
        loss_compute(loss_kind ? loss_kind[i] : 0, loss_delta ? loss_delta[i] : 1.0, e2, rho);
        t += (loss_kind && loss_kind[i] != 0) ? rho[0] : e2;
    }
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
__global__ void k_dense_scale(const double *dx, const double *b, int n, double lambda, double *out2) {
    __shared__ double s1[256], s2[256];
    double sc = 0.0, n2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sc += dx[i] * (lambda * dx[i] + b[i]); n2 += dx[i] * dx[i]; }
    s1[threadIdx.x] = sc; s2[threadIdx.x] = n2;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if (threadIdx.x < st) { s1[threadIdx.x] += s1[threadIdx.x + st]; s2[threadIdx.x] += s2[threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = s1[0]; out2[1] = s2[0]; }
}
__global__ void k_absmax_diag(const double *H, int n, double *out) {
    __shared__ double sm[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(H[(size_t)i * n + i]));
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
