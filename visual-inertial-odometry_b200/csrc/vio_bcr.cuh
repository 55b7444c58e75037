// vio_bcr.cuh — device side of the block cyclic reduction solver (plan: vio_bcr.h).
//   k_bcr_load     block-sparse S (BSR) + lambda -> dense node tiles D_i / couplings E_i, b -> node vectors
//   k_bcr_run      persistent kernel: CTAs take the plan's items in order from an atomic counter, wait for the items
//                  they depend on (per-item flags, acquire/release), and do the item's dense M x M work in shared memory
//   k_bcr_finish   node vectors -> dx_p ; isolated pose blocks solved as 6x6 systems
// Exact replacement of S.ldlt().solve (A17/src/backend/problem.cc:434-440) on a camera chain / ring.
// FP64 everywhere.  All products are of the form C = A^T B with row-major tiles, so both operands are read along rows
// (16-byte shared loads, broadcast across the lanes that share a tile row); the inverse Cholesky factor is kept
// transposed (U = L^-T) for the same reason.
#pragma once
#include "vio_dev.h"
#include "vio_bcr.h"

#define BCR_THREADS 256

struct BcrView {
    int n, M, n_items;
    int nbuf;            // 7 or 5 operand tiles in shared memory (see k_bcr_run)
    const BcrItem *items;
    double *pool;        // [n_slots][M*M]
    double *bv, *xv;     // [n][M]
    unsigned *flags;     // [n_items] = epoch when the item is complete
    unsigned *counter;   // work queue head (zeroed before the launch)
    unsigned epoch;
    int *info;           // != 0: a pivot was not positive
    unsigned long long *prof;  // optional [8] cycle counters (VIO_B200_PROFILE=1): dependency wait, updates + couplings, Cholesky,
                               // W products + stores, back-substitution items, kept-node items, #eliminations, #items
};

// ---- loader ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bcr_load(const double *__restrict__ val, const long long *__restrict__ dst, long long nnzb,
                                                  const double *__restrict__ b, const int *__restrict__ blk_node,
                                                  const int *__restrict__ blk_loc, const int *__restrict__ node_size, int nb, int n, int M,
                                                  double lambda, double *__restrict__ pool, double *__restrict__ bv) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long t = t0; t < nnzb * 36; t += stride) {
        const long long k = t / 36;
        const int e = (int)(t - 36 * k);
        const long long d = dst[k];
        if (d < 0) continue;
        const bool dg = (d & BCR_DST_DIAG) != 0;
        pool[(d & ~BCR_DST_DIAG) + (long long)(e / 6) * M + e % 6] = val[t] + ((dg && e % 7 == 0) ? lambda : 0.0);
    }
    for (long long t = t0; t < (long long)n * M; t += stride) {  // identity padding of ragged nodes
        const int a = (int)(t / M), q = (int)(t % M);
        if (q >= 6 * node_size[a]) pool[(long long)a * M * M + (long long)q * M + q] = 1.0;
    }
    for (long long t = t0; t < 6LL * nb; t += stride) {
        const int i = (int)(t / 6), c = (int)(t % 6);
        if (blk_node[i] >= 0) bv[(long long)blk_node[i] * M + 6 * blk_loc[i] + c] = b[t];
    }
}

// ---- finish: gather x, solve the isolated 6x6 blocks ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bcr_finish(const double *__restrict__ xv, const int *__restrict__ blk_node, const int *__restrict__ blk_loc,
                                                    int nb, int M, const double *__restrict__ val, const int *__restrict__ diag,
                                                    const double *__restrict__ b, double lambda, double *__restrict__ x, int *info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    if (blk_node[i] >= 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) x[6 * (size_t)i + c] = xv[(size_t)blk_node[i] * M + 6 * blk_loc[i] + c];
        return;
    }
    double A[36], y[6];
    const double *d = val + 36 * (size_t)diag[i];
#pragma unroll
    for (int e = 0; e < 36; ++e) A[e] = d[e] + (e % 7 == 0 ? lambda : 0.0);
#pragma unroll
    for (int c = 0; c < 6; ++c) y[c] = b[6 * (size_t)i + c];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        if (!(A[7 * c] > 0.0)) { *info = -(i + 1); A[7 * c] = 1.0; }
        const double inv = 1.0 / A[7 * c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[6 * r + c] * inv;
#pragma unroll
            for (int k = c; k < 6; ++k) A[6 * r + k] -= f * A[6 * c + k];
            y[r] -= f * y[c];
        }
    }
#pragma unroll
    for (int c = 5; c >= 0; --c) {
        double a = y[c];
#pragma unroll
        for (int k = c + 1; k < 6; ++k) a -= A[6 * c + k] * y[k];
        y[c] = a / A[7 * c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) x[6 * (size_t)i + c] = y[c];
}

// ---- tile helpers (all threads of the CTA) ---------------------------------------------------------------------------
__device__ __forceinline__ void bcr_load_tile(double *dst, const double *src, int M, bool transpose) {
    const int MM = M * M;
    if (!transpose) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (int t = threadIdx.x; t < (MM >> 1); t += blockDim.x) d2[t] = __ldcg(s2 + t);
    } else {
        for (int t = threadIdx.x; t < MM; t += blockDim.x) {
            const int r = t / M, c = t - r * M;
            dst[c * M + r] = __ldcg(src + t);
        }
    }
}
__device__ __forceinline__ void bcr_store_tile(double *dst, const double *src, int M) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (int t = threadIdx.x; t < ((M * M) >> 1); t += blockDim.x) __stcg(d2 + t, s2[t]);
}

// C = A^T B over 4x4 register tiles.  A thread's tile is NOT a contiguous 4x4 patch: it owns rows {2ti, 2ti+1, h+2ti,
// h+2ti+1} and columns {2tj, 2tj+1, h+2tj, h+2tj+1} (h = M/2), so the 16-byte operand loads of consecutive threads are
// contiguous in shared memory (conflict free; a contiguous 4-wide patch would put consecutive threads 32 bytes apart,
// a 2-way bank conflict per quarter warp that makes the products shared-memory bound instead of FP64 bound).
// epi(i, j, v0, v1) receives elements (i, j) and (i, j+1) of the product (j even: 16-byte aligned).  TRI: A is upper triangular (A[r][i] = 0 for r > i).
template <bool TRI, class Epi>
__device__ __forceinline__ void bcr_tn(const double *__restrict__ A, const double *__restrict__ B, int M, Epi epi) {
    const int T = M >> 2, h = M >> 1;
    for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
        const int ti = t / T, tj = t - ti * T;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        const int rend = TRI ? min(M, h + 2 * ti + 2) : M;
        const double *ap = A + 2 * ti, *bp = B + 2 * tj;
#pragma unroll 4
        for (int r = 0; r < rend; ++r, ap += M, bp += M) {
            const double2 a01 = *reinterpret_cast<const double2 *>(ap), a23 = *reinterpret_cast<const double2 *>(ap + h);
            const double2 b01 = *reinterpret_cast<const double2 *>(bp), b23 = *reinterpret_cast<const double2 *>(bp + h);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bw[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bw[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; b += 2) epi(2 * ti + (a & 1) + (a >> 1) * h, 2 * tj + (b >> 1) * h, acc[a][b], acc[a][b + 1]);
    }
}

// fused pair sharing the A operand:  C1 = A^T A (symmetric update) and C2 = A^T B
template <class Epi1, class Epi2>
__device__ __forceinline__ void bcr_tn_pair(const double *__restrict__ A, const double *__restrict__ B, int M, Epi1 epi1, Epi2 epi2) {
    const int T = M >> 2, h = M >> 1;
    for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
        const int ti = t / T, tj = t - ti * T;
        double c1[4][4], c2[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { c1[a][b] = 0.0; c2[a][b] = 0.0; }
        const double *ap = A + 2 * ti, *aq = A + 2 * tj, *bp = B + 2 * tj;
#pragma unroll 2
        for (int r = 0; r < M; ++r, ap += M, aq += M, bp += M) {
            const double2 a01 = *reinterpret_cast<const double2 *>(ap), a23 = *reinterpret_cast<const double2 *>(ap + h);
            const double2 q01 = *reinterpret_cast<const double2 *>(aq), q23 = *reinterpret_cast<const double2 *>(aq + h);
            const double2 b01 = *reinterpret_cast<const double2 *>(bp), b23 = *reinterpret_cast<const double2 *>(bp + h);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, qv[4] = {q01.x, q01.y, q23.x, q23.y}, bw[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { c1[a][b] += av[a] * qv[b]; c2[a][b] += av[a] * bw[b]; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; b += 2) {
                const int i = 2 * ti + (a & 1) + (a >> 1) * h, j = 2 * tj + (b >> 1) * h;
                epi1(i, j, c1[a][b], c1[a][b + 1]);
                epi2(i, j, c2[a][b], c2[a][b + 1]);
            }
    }
}

// v[i] -= sum_r A[r][i] * y[r]   (A: shared tile, y: shared vector).  Four threads per element (r = part mod 4), combined
// through `scratch` ([4][M]); contains two barriers - every thread of the CTA must call it.
__device__ __forceinline__ void bcr_gemv_t_sub(const double *A, const double *y, double *v, double *scratch, int M) {
    for (int t = threadIdx.x; t < 4 * M; t += blockDim.x) {
        const int part = t / M, i = t - part * M;
        double a = 0.0;
        for (int r = part; r < M; r += 4) a += A[r * M + i] * y[r];
        scratch[part * M + i] = a;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < M; t += blockDim.x) v[t] -= (scratch[t] + scratch[M + t]) + (scratch[2 * M + t] + scratch[3 * M + t]);
    __syncthreads();
}

// D (shared, symmetric positive definite, destroyed) -> U = L^-T (shared) with D = L L^T, by forward elimination on
// [D | I]: column j scales row j by 1/sqrt(d_jj) and subtracts it from the rows below; the identity part, kept
// transposed, turns into L^-T.  Blocked by PANELS of 4 columns with one-panel LOOK-AHEAD:
//   * warp 0 owns the critical path: it brings the NEXT panel's 4 rows up to date in registers (rank-4 update with the
//     current panel), factorises them there (lanes = columns, pivots and multipliers by shuffle) and publishes the
//     scaled rows V'[k][0..3], the reciprocal pivots and the panel's 4x4 triangle for the following iteration;
//   * warps 1.. apply the current panel to everything else: rank-4 update of the trailing rows of D, and for the
//     rows of U first the panel's 4x4 triangular transform of columns j0..j0+3, then the same rank-4 update.
// One barrier per panel (M/4 panels); V and the panel scalars are double buffered.
#define BCR_PANEL 4
#define BCR_MAXPASS 3  // ceil(BCR_MAX_M / 32)
static_assert(BCR_MAX_M <= 32 * BCR_MAXPASS, "panel rows are held by one warp");

// factorise 4 panel rows held in registers (r[q][sp] = row j0+q, column j0 + lane + 32 sp); sc[0..3] = 1/sqrt(pivot),
// sc[4..9] = v_a[j0+q] for a < q in the order (0,1) (0,2) (1,2) (0,3) (1,3) (2,3)
template <int NPASS>
__device__ __forceinline__ void bcr_panel_factor(double (&r)[BCR_PANEL][NPASS], int j0, int lane, int M, double *__restrict__ Vn,
                                                 double *__restrict__ sc, int *info) {
    double pq[BCR_PANEL];
#pragma unroll
    for (int q = 0; q < BCR_PANEL; ++q) {
        const double d = __shfl_sync(0xffffffffu, r[q][0], q);
        if (lane == 0 && !(d > 0.0)) *info = j0 + q + 1;
        const double p = rsqrt(d > 0.0 ? d : 1.0);
        pq[q] = p;
#pragma unroll
        for (int sp = 0; sp < NPASS; ++sp) r[q][sp] *= p;
        if (lane < q) r[q][0] = 0.0;  // left of the diagonal: already eliminated
#pragma unroll
        for (int i = q + 1; i < BCR_PANEL; ++i) {
            const double f = __shfl_sync(0xffffffffu, r[q][0], i);  // v_q[j0 + i]
#pragma unroll
            for (int sp = 0; sp < NPASS; ++sp) r[i][sp] -= f * r[q][sp];
        }
    }
#pragma unroll
    for (int sp = 0; sp < NPASS; ++sp) {
        const int k = j0 + lane + 32 * sp;
        if (k < M) {
            *reinterpret_cast<double2 *>(Vn + 4 * k) = make_double2(r[0][sp], r[1][sp]);
            *reinterpret_cast<double2 *>(Vn + 4 * k + 2) = make_double2(r[2][sp], r[3][sp]);
        }
    }
    if (lane == 0) { sc[0] = pq[0]; sc[1] = pq[1]; sc[2] = pq[2]; sc[3] = pq[3]; }
    if (lane == 1) sc[4] = r[0][0];
    if (lane == 2) { sc[5] = r[0][0]; sc[6] = r[1][0]; }
    if (lane == 3) { sc[7] = r[0][0]; sc[8] = r[1][0]; sc[9] = r[2][0]; }
}

template <int NPASS>  // ceil(M / 32)
__device__ __forceinline__ void bcr_chol_inv(double *__restrict__ D, double *__restrict__ U, double *__restrict__ V /* [2][M][4] */,
                                             double *__restrict__ sc /* [2][16] */, int M, int *info, unsigned long long *prof) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    long long cw0 = 0, cw1 = 0, cb0 = 0, cb1 = 0;
    for (int t = tid; t < M * M; t += nt) U[t] = (t / M == t % M) ? 1.0 : 0.0;
    if (warp == 0) {
        double r[BCR_PANEL][NPASS];
#pragma unroll
        for (int q = 0; q < BCR_PANEL; ++q)
#pragma unroll
            for (int sp = 0; sp < NPASS; ++sp) {
                const int k = lane + 32 * sp;
                r[q][sp] = k < M ? D[q * M + k] : 0.0;
            }
        bcr_panel_factor<NPASS>(r, 0, lane, M, V, sc, info);
    }
    __syncthreads();
    const int np = M / BCR_PANEL;
    for (int pnl = 0; pnl < np; ++pnl) {
        const int j0 = BCR_PANEL * pnl, k0 = j0 + BCR_PANEL, cur = pnl & 1;
        const long long tA = prof ? clock64() : 0;
        const double *Vc = V + cur * 4 * M, *scc = sc + cur * 16;
        double vk[NPASS][BCR_PANEL];
#pragma unroll
        for (int sp = 0; sp < NPASS; ++sp) {
            const int k = k0 + lane + 32 * sp;
            if (k < M) {
                const double2 a = *reinterpret_cast<const double2 *>(Vc + 4 * k), b2 = *reinterpret_cast<const double2 *>(Vc + 4 * k + 2);
                vk[sp][0] = a.x; vk[sp][1] = a.y; vk[sp][2] = b2.x; vk[sp][3] = b2.y;
            } else {
                vk[sp][0] = vk[sp][1] = vk[sp][2] = vk[sp][3] = 0.0;
            }
        }
        if (warp == 0) {
            if (k0 < M) {
                double r[BCR_PANEL][NPASS];
#pragma unroll
                for (int q = 0; q < BCR_PANEL; ++q) {
                    const double *fp = Vc + 4 * (k0 + q);
                    const double2 f01 = *reinterpret_cast<const double2 *>(fp), f23 = *reinterpret_cast<const double2 *>(fp + 2);
#pragma unroll
                    for (int sp = 0; sp < NPASS; ++sp) {
                        const int k = k0 + lane + 32 * sp;
                        r[q][sp] = k < M ? D[(k0 + q) * M + k] - (f01.x * vk[sp][0] + f01.y * vk[sp][1] + f23.x * vk[sp][2] + f23.y * vk[sp][3]) : 0.0;
                    }
                }
                bcr_panel_factor<NPASS>(r, k0, lane, M, V + (cur ^ 1) * 4 * M, sc + (cur ^ 1) * 16, info);
            }
        } else {
            const int nD = max(0, M - k0 - BCR_PANEL);  // trailing rows of D below the next panel
            const int nwk = nw - 1, w1 = warp - 1;
            int kk[NPASS];
            bool kv[NPASS];
#pragma unroll
            for (int sp = 0; sp < NPASS; ++sp) { kk[sp] = k0 + lane + 32 * sp; kv[sp] = kk[sp] < M; if (!kv[sp]) kk[sp] = M - 1; }
            // (1) trailing rows of D, two at a time (loads grouped before the stores: the rows are independent but live in
            //     the same array, the compiler would otherwise serialise them)
            for (int rr = w1; rr < nD; rr += 2 * nwk) {
                const int ia = k0 + BCR_PANEL + rr, ib = ia + nwk;
                const bool hb = rr + nwk < nD;
                double *ra = D + ia * M, *rb = D + (hb ? ib : ia) * M;
                const double2 fa01 = *reinterpret_cast<const double2 *>(Vc + 4 * ia), fa23 = *reinterpret_cast<const double2 *>(Vc + 4 * ia + 2);
                const double2 fb01 = *reinterpret_cast<const double2 *>(Vc + 4 * (hb ? ib : ia)), fb23 = *reinterpret_cast<const double2 *>(Vc + 4 * (hb ? ib : ia) + 2);
                double va[NPASS], vb[NPASS];
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) { va[sp] = ra[kk[sp]]; vb[sp] = rb[kk[sp]]; }
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) {
                    va[sp] -= fa01.x * vk[sp][0] + fa01.y * vk[sp][1] + fa23.x * vk[sp][2] + fa23.y * vk[sp][3];
                    vb[sp] -= fb01.x * vk[sp][0] + fb01.y * vk[sp][1] + fb23.x * vk[sp][2] + fb23.y * vk[sp][3];
                }
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) {
                    if (kv[sp]) ra[kk[sp]] = va[sp];
                    if (kv[sp] && hb) rb[kk[sp]] = vb[sp];
                }
            }
            // (2) rows 0 .. j0+3 of U: the panel's 4x4 triangular transform of columns j0..j0+3, then the same rank-4 update
            const double p0 = scc[0], p1 = scc[1], p2 = scc[2], p3 = scc[3];
            const double s01 = scc[4], s02 = scc[5], s12 = scc[6], s03 = scc[7], s13 = scc[8], s23 = scc[9];
            for (int rr = w1; rr < k0; rr += 2 * nwk) {
                const bool hb = rr + nwk < k0;
                double *ra = U + rr * M, *rb = U + (hb ? rr + nwk : rr) * M;
                const double2 ua01 = *reinterpret_cast<const double2 *>(ra + j0), ua23 = *reinterpret_cast<const double2 *>(ra + j0 + 2);
                const double2 ub01 = *reinterpret_cast<const double2 *>(rb + j0), ub23 = *reinterpret_cast<const double2 *>(rb + j0 + 2);
                double va[NPASS], vb[NPASS];
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) { va[sp] = ra[kk[sp]]; vb[sp] = rb[kk[sp]]; }
                const double a0 = ua01.x * p0, b0 = ub01.x * p0;
                const double a1 = (ua01.y - s01 * a0) * p1, b1 = (ub01.y - s01 * b0) * p1;
                const double a2 = (ua23.x - s02 * a0 - s12 * a1) * p2, b2 = (ub23.x - s02 * b0 - s12 * b1) * p2;
                const double a3 = (ua23.y - s03 * a0 - s13 * a1 - s23 * a2) * p3, b3 = (ub23.y - s03 * b0 - s13 * b1 - s23 * b2) * p3;
                __syncwarp();  // every lane has read columns j0..j0+3 before lane 0 overwrites them
                if (lane == 0) {
                    *reinterpret_cast<double2 *>(ra + j0) = make_double2(a0, a1);
                    *reinterpret_cast<double2 *>(ra + j0 + 2) = make_double2(a2, a3);
                    if (hb) {
                        *reinterpret_cast<double2 *>(rb + j0) = make_double2(b0, b1);
                        *reinterpret_cast<double2 *>(rb + j0 + 2) = make_double2(b2, b3);
                    }
                }
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) {
                    va[sp] -= a0 * vk[sp][0] + a1 * vk[sp][1] + a2 * vk[sp][2] + a3 * vk[sp][3];
                    vb[sp] -= b0 * vk[sp][0] + b1 * vk[sp][1] + b2 * vk[sp][2] + b3 * vk[sp][3];
                }
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) {
                    if (kv[sp]) ra[kk[sp]] = va[sp];
                    if (kv[sp] && hb) rb[kk[sp]] = vb[sp];
                }
            }
        }
        const long long tB = prof ? clock64() : 0;
        __syncthreads();
        if (prof) {
            const long long tC = clock64();
            if (tid == 0) { cw0 += tB - tA; cb0 += tC - tB; }
            if (tid == 32) { cw1 += tB - tA; cb1 += tC - tB; }
        }
    }
    if (prof && tid == 0) { atomicAdd(prof + 8, (unsigned long long)cw0); atomicAdd(prof + 9, (unsigned long long)cb0); }
    if (prof && tid == 32) { atomicAdd(prof + 10, (unsigned long long)cw1); atomicAdd(prof + 11, (unsigned long long)cb1); }
}

__device__ __forceinline__ void bcr_wait(const unsigned *flag, unsigned epoch) {
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    } while (v != epoch);
}

// ---- 1-D bulk asynchronous copies global -> shared (TMA, cp.async.bulk) completing on an mbarrier ----------------------
// An item's operand tiles are contiguous M*M*8-byte slabs of the pool: ONE thread issues all of them up front, the copy
// engine fills shared memory while the CTA does nothing else (no registers, no per-thread round trips), and everybody
// waits on the barrier's phase once.
__device__ __forceinline__ unsigned bcr_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bcr_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bcr_saddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bcr_mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bcr_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bcr_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bcr_saddr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bcr_saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void bcr_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BCR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BCR_DONE;\n"
        "bra BCR_WAIT;\n"
        "BCR_DONE:\n"
        "}\n" ::"r"(bcr_saddr(bar)),
        "r"(parity)
        : "memory");
}
// generic-proxy accesses (earlier shared-memory reads of the destination, global data acquired through the item flags)
// are ordered before the async-proxy copies issued after this fence
__device__ __forceinline__ void bcr_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- the persistent kernel ------------------------------------------------------------------------------------------
// Shared memory: tiles Dm | X | Z | opA0 | opB0 [| opA1 | opB1], vectors, Cholesky panel buffers.  With seven tiles
// (nbuf = 7, M <= 60) both sides' operands are fetched with the first bulk batch; with five, side 1 reuses side 0's.
__global__ void __launch_bounds__(BCR_THREADS, 1) k_bcr_run(BcrView s) {
    extern __shared__ __align__(128) double bsm[];
    const int M = s.M, MM = M * M, tid = threadIdx.x, nt = blockDim.x;
    const unsigned tile_bytes = (unsigned)MM * 8u;
    double *Dm = bsm, *X = Dm + MM, *Z = X + MM;
    // operand tiles as OFFSETS into bsm (a pointer array would make the compiler lose the shared address space and fall
    // back to generic loads): side sd uses bsm + opo(sd) and bsm + opo(sd) + MM
    const int opo1 = s.nbuf == 7 ? 5 * MM : 3 * MM;
#define BCR_OPA(sd_) (bsm + ((sd_) == 0 ? 3 * MM : opo1))
#define BCR_OPB(sd_) (bsm + ((sd_) == 0 ? 4 * MM : opo1 + MM))
    double *bk = bsm + (size_t)s.nbuf * MM, *ye0 = bk + M, *ye1 = ye0 + M, *vv = ye1 + M;  // then V[2][M][4] + 32 panel scalars
    __shared__ BcrItem it_s;
    __shared__ unsigned idx_s;
    __shared__ __align__(8) unsigned long long ldbar;
    if (tid == 0) bcr_mbar_init(&ldbar, 1);
    unsigned ph = 0;
    for (;;) {
        __syncthreads();  // the previous item's shared-memory traffic is over (and the barrier is initialised)
        if (tid == 0) idx_s = atomicAdd(s.counter, 1u);
        __syncthreads();
        const unsigned idx = idx_s;
        if (idx >= (unsigned)s.n_items) return;
        if (tid < (int)(sizeof(BcrItem) / sizeof(int))) reinterpret_cast<int *>(&it_s)[tid] = reinterpret_cast<const int *>(s.items + idx)[tid];
        __syncthreads();
        long long pt_ = (s.prof && tid == 0) ? clock64() : 0;
#define BCR_MARK(slot_) do { if (s.prof && tid == 0) { const long long n_ = clock64(); atomicAdd(s.prof + (slot_), (unsigned long long)(n_ - pt_)); pt_ = n_; } } while (0)
        if (tid < 6 && it_s.dep[tid] >= 0) bcr_wait(s.flags + it_s.dep[tid], s.epoch);
        __syncthreads();
        BCR_MARK(0);
        const BcrItem &it = it_s;
        const size_t node_off = (size_t)it.node * MM;
        if (it.kind & BCR_BACKSUB) {
            // x_k = U (y_k - W_l x_l - W_r x_r): the three tiles come in as one bulk batch, every dot product is one warp wide
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            const bool hl = it.left >= 0, hr = it.right >= 0;
            if (tid == 0) {
                bcr_fence_async();
                bcr_mbar_expect(&ldbar, tile_bytes * (1u + (hl ? 1u : 0u) + (hr ? 1u : 0u)));
                bcr_bulk_load(Dm, s.pool + node_off, tile_bytes, &ldbar);
                if (hl) bcr_bulk_load(BCR_OPA(0), s.pool + (size_t)it.cl_slot * MM, tile_bytes, &ldbar);
                if (hr) bcr_bulk_load(BCR_OPB(0), s.pool + (size_t)it.cr_slot * MM, tile_bytes, &ldbar);
            }
            for (int i = tid; i < M; i += nt) {
                bk[i] = __ldcg(s.bv + (size_t)it.node * M + i);
                ye0[i] = hl ? __ldcg(s.xv + (size_t)it.left * M + i) : 0.0;
                ye1[i] = hr ? __ldcg(s.xv + (size_t)it.right * M + i) : 0.0;
            }
            bcr_mbar_wait(&ldbar, ph);
            ph ^= 1u;
            __syncthreads();
            const double *Wl = BCR_OPA(0), *Wr = BCR_OPB(0);
            for (int i = warp; i < M; i += nw) {
                double a = 0.0;
                for (int c = lane; c < M; c += 32) {
                    if (hl) a += Wl[i * M + c] * ye0[c];
                    if (hr) a += Wr[i * M + c] * ye1[c];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) vv[i] = bk[i] - a;
            }
            __syncthreads();
            for (int i = warp; i < M; i += nw) {
                double a = 0.0;
                for (int r = i + lane; r < M; r += 32) a += Dm[i * M + r] * vv[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) __stcg(s.xv + (size_t)it.node * M + i, a);
            }
            __syncthreads();
            BCR_MARK(4);
        } else {
            const bool elim = (it.kind & BCR_ELIM) != 0;
            // per side: Schur update from the neighbour eliminated one level earlier (W = upd_slot), and - when this node is
            // being eliminated - the coupling tile with rows = this node (side 0 -> X, side 1 -> Z)
            int us[2], mode[2], ca[2], cb[2];
            bool fused[2], act[2];
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                us[sd] = it.upd_slot[sd];
                mode[sd] = elim ? (sd == 0 ? it.cl_mode : it.cr_mode) : 0;
                ca[sd] = sd == 0 ? it.cl_a : it.cr_a;
                cb[sd] = sd == 0 ? it.cl_b : it.cr_b;
                fused[sd] = mode[sd] == 2 && us[sd] >= 0 && ca[sd] == us[sd];  // the usual case: update and coupling share W
                act[sd] = us[sd] >= 0 || mode[sd] != 0;
            }
            // bulk loads of one side's operands (thread 0); returns the bytes issued
            auto side_bytes = [&](int sd) -> unsigned {
                return tile_bytes * ((us[sd] >= 0 ? 1u : 0u) + (fused[sd] ? 1u : 0u) + ((mode[sd] == 1 && cb[sd] == 0) ? 1u : 0u));
            };
            auto side_issue = [&](int sd) {
                if (us[sd] >= 0) bcr_bulk_load(BCR_OPA(sd), s.pool + (size_t)us[sd] * MM, tile_bytes, &ldbar);
                if (fused[sd]) bcr_bulk_load(BCR_OPB(sd), s.pool + (size_t)cb[sd] * MM, tile_bytes, &ldbar);
                if (mode[sd] == 1 && cb[sd] == 0) bcr_bulk_load(sd == 0 ? X : Z, s.pool + (size_t)ca[sd] * MM, tile_bytes, &ldbar);
            };
            const bool both_now = s.nbuf == 7;
            if (tid == 0) {
                bcr_fence_async();
                bcr_mbar_expect(&ldbar, tile_bytes + side_bytes(0) + (both_now ? side_bytes(1) : 0u));
                bcr_bulk_load(Dm, s.pool + node_off, tile_bytes, &ldbar);
                side_issue(0);
                if (both_now) side_issue(1);
            }
            for (int i = tid; i < M; i += nt) {
                bk[i] = __ldcg(s.bv + (size_t)it.node * M + i);
                if (us[0] >= 0) ye0[i] = __ldcg(s.bv + (size_t)it.upd_node[0] * M + i);
                if (us[1] >= 0) ye1[i] = __ldcg(s.bv + (size_t)it.upd_node[1] * M + i);
            }
            bcr_mbar_wait(&ldbar, ph);
            ph ^= 1u;
            __syncthreads();
#pragma unroll 1
            for (int sd = 0; sd < 2; ++sd) {
                if (!act[sd]) continue;
                double *A1 = BCR_OPA(sd), *A2 = BCR_OPB(sd), *OUT = sd == 0 ? X : Z;
                const double *ye = sd == 0 ? ye0 : ye1;
                if (sd == 1 && !both_now) {
                    __syncthreads();  // side 0 is done with the operand tiles
                    if (side_bytes(1) != 0u) {
                        if (tid == 0) {
                            bcr_fence_async();
                            bcr_mbar_expect(&ldbar, side_bytes(1));
                            side_issue(1);
                        }
                        bcr_mbar_wait(&ldbar, ph);
                        ph ^= 1u;
                    }
                }
                if (mode[sd] == 1 && cb[sd] != 0) bcr_load_tile(OUT, s.pool + (size_t)ca[sd] * MM, M, true);  // rare: transposed coupling
                auto upd = [&](int i, int j, double v0, double v1) {
                    double2 *d = reinterpret_cast<double2 *>(Dm + i * M + j);
                    double2 c = *d;
                    c.x -= v0; c.y -= v1;
                    *d = c;
                };
                auto neg_out = [&](int i, int j, double v0, double v1) { *reinterpret_cast<double2 *>(OUT + i * M + j) = make_double2(-v0, -v1); };
                if (fused[sd]) {
                    bcr_tn_pair(A1, A2, M, upd, neg_out);
                    bcr_gemv_t_sub(A1, ye, bk, vv, M);
                } else if (us[sd] >= 0) {
                    bcr_tn<false>(A1, A1, M, upd);
                    bcr_gemv_t_sub(A1, ye, bk, vv, M);
                }
                if (mode[sd] == 2 && !fused[sd]) {  // rare: a coupling carried over a level, its factors are not this level's W
                    __syncthreads();
                    bcr_load_tile(A1, s.pool + (size_t)ca[sd] * MM, M, false);
                    bcr_load_tile(A2, s.pool + (size_t)cb[sd] * MM, M, false);
                    __syncthreads();
                    bcr_tn<false>(A1, A2, M, neg_out);
                }
            }
            __syncthreads();
            if (!elim) {
                bcr_store_tile(s.pool + node_off, Dm, M);
                for (int i = tid; i < M; i += nt) __stcg(s.bv + (size_t)it.node * M + i, bk[i]);
                __syncthreads();
                BCR_MARK(5);
            } else {
                BCR_MARK(1);
                const bool hasL = it.cl_mode != 0 && !(it.kind & BCR_MERGE), hasR = it.cr_mode != 0;
                if (it.kind & BCR_MERGE) {
                    for (int t = tid; t < MM; t += nt) Z[t] += X[t];
                    __syncthreads();
                }
                double *U = BCR_OPA(0);
                if (M <= 64) bcr_chol_inv<2>(Dm, U, vv, vv + 8 * M, M, s.info, s.prof);
                else bcr_chol_inv<3>(Dm, U, vv, vv + 8 * M, M, s.info, s.prof);
                BCR_MARK(2);
                // W_l = U^T X, W_r = U^T Z -> their pool tiles; y = U^T b; U -> the node's tile
                if (hasL) {
                    double *out = s.pool + (size_t)it.cl_slot * MM;
                    bcr_tn<true>(U, X, M, [&](int i, int j, double v0, double v1) { __stcg(reinterpret_cast<double2 *>(out + (size_t)i * M + j), make_double2(v0, v1)); });
                }
                if (hasR) {
                    double *out = s.pool + (size_t)it.cr_slot * MM;
                    bcr_tn<true>(U, Z, M, [&](int i, int j, double v0, double v1) { __stcg(reinterpret_cast<double2 *>(out + (size_t)i * M + j), make_double2(v0, v1)); });
                }
                for (int i = tid; i < M; i += nt) {
                    double a = 0.0;
                    for (int r = 0; r <= i; ++r) a += U[r * M + i] * bk[r];
                    __stcg(s.bv + (size_t)it.node * M + i, a);
                }
                bcr_store_tile(s.pool + node_off, U, M);
                __syncthreads();
                BCR_MARK(3);
                if (s.prof && tid == 0) atomicAdd(s.prof + 6, 1ull);
            }
        }
        if (s.prof && tid == 0) atomicAdd(s.prof + 7, 1ull);
#undef BCR_MARK
        // publish: every thread's global stores are done and visible before the flag
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(s.flags + idx), "r"(s.epoch) : "memory");
        }
    }
}
