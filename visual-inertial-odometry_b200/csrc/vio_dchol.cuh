// vio_dchol.cuh — blocked dense Cholesky solve for the sliding-window sizes (P = 120 .. 192), one CTA per system:
//     (S + lambda I) x = b,   replacing S.ldlt().solve  (A17/src/backend/problem.cc:434-440)
// The column-by-column kernel (k_dense_chol_smem) pays 3 barriers and a sqrt + divide per column (350 us at P = 171);
// here the factorisation is blocked by PANELS of 4 rows with one-panel look-ahead, like the node factorisation of the
// block cyclic reduction (vio_bcr.cuh):
//   * warp 0 owns the critical path: it brings the next panel's 4 rows (and right-hand side) up to date in registers,
//     factorises the 4x4 leading block redundantly in every lane (no shuffles), scales its columns and publishes
//     V[k][0..3] = the 4 new rows of R (S + lambda I = R^T R), which it also writes back as rows of R;
//   * the other warps apply the current panel as rank-4 DMMA (mma.sync.m8n8k4.f64) updates to the 8x8 tiles of the
//     trailing UPPER triangle, and the same rank-4 update to the right-hand side (forward substitution rides along).
// One barrier per panel.  R is kept in shared memory in a pair-packed upper layout: rows 2j and 2j+1 both start at
// column 2j, so every row starts on a 16-byte boundary (P^2/2 + P doubles: 120 KB at P = 172).  The back-substitution
// R x = y runs in blocks of 4 rows with the reciprocal pivots kept from the factorisation.
#pragma once
#include "vio_dev.h"

#define DCH_THREADS 256
#define DCH_MAX_P 192

__device__ __forceinline__ void dch_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// pair-packed upper storage: row i (pair j = i / 2) holds columns 2j .. Pp-1
__host__ __device__ __forceinline__ int dch_row(int i, int Pp) {
    const int j = i >> 1;
    return 2 * j * Pp - 2 * j * (j - 1) + (i & 1) * (Pp - 2 * j) - 2 * j;  // + k gives the address of element (i, k)
}
__host__ __device__ inline size_t dch_smem_bytes(int P) {
    const int Pp = (P + 3) & ~3;
    return ((size_t)Pp * Pp / 2 + Pp + 2 * 4 * (size_t)Pp + 2 * 16 + 2 * (size_t)Pp) * sizeof(double);  // R, V[2], sc[2], y, pinv
}

// factorise 4 panel rows held in registers; see bcr_panel_factor.  ry[q] = the rows' right-hand side entries.
template <int NPASS>
__device__ __forceinline__ void dch_panel_factor(double (&r)[4][NPASS], double (&ry)[4], const double (&blk)[10], int j0, int lane, int Pp,
                                                 double *__restrict__ R, double *__restrict__ Vn, double *__restrict__ sc,
                                                 double *__restrict__ y, double *__restrict__ pinv, int *info) {
    const bool bad0 = !(blk[0] > 0.0);
    const double p0 = rsqrt(bad0 ? 1.0 : blk[0]);
    const double l10 = blk[1] * p0, l20 = blk[3] * p0, l30 = blk[6] * p0;
    const double d1 = blk[2] - l10 * l10;
    const bool bad1 = !(d1 > 0.0);
    const double p1 = rsqrt(bad1 ? 1.0 : d1);
    const double l21 = (blk[4] - l20 * l10) * p1, l31 = (blk[7] - l30 * l10) * p1;
    const double d2 = blk[5] - l20 * l20 - l21 * l21;
    const bool bad2 = !(d2 > 0.0);
    const double p2 = rsqrt(bad2 ? 1.0 : d2);
    const double l32 = (blk[8] - l30 * l20 - l31 * l21) * p2;
    const double d3 = blk[9] - l30 * l30 - l31 * l31 - l32 * l32;
    const bool bad3 = !(d3 > 0.0);
    const double p3 = rsqrt(bad3 ? 1.0 : d3);
    if (lane == 0 && (bad0 || bad1 || bad2 || bad3)) *info = j0 + 1 + (bad0 ? 0 : (bad1 ? 1 : (bad2 ? 2 : 3)));
#pragma unroll
    for (int sp = 0; sp < NPASS; ++sp) {
        const double v0 = r[0][sp] * p0;
        const double v1 = (r[1][sp] - l10 * v0) * p1;
        const double v2 = (r[2][sp] - l20 * v0 - l21 * v1) * p2;
        const double v3 = (r[3][sp] - l30 * v0 - l31 * v1 - l32 * v2) * p3;
        r[0][sp] = v0; r[1][sp] = v1; r[2][sp] = v2; r[3][sp] = v3;
    }
    if (lane < 1) r[1][0] = 0.0;
    if (lane < 2) r[2][0] = 0.0;
    if (lane < 3) r[3][0] = 0.0;
    // forward substitution of the panel's right-hand side entries
    const double y0 = ry[0] * p0;
    const double y1 = (ry[1] - l10 * y0) * p1;
    const double y2 = (ry[2] - l20 * y0 - l21 * y1) * p2;
    const double y3 = (ry[3] - l30 * y0 - l31 * y1 - l32 * y2) * p3;
    // every lane has read the rows, the 4x4 block and y of this panel (in the caller) before any lane overwrites them
    __syncwarp();
#pragma unroll
    for (int sp = 0; sp < NPASS; ++sp) {
        const int k = j0 + lane + 32 * sp;
        if (k < Pp) {
            *reinterpret_cast<double2 *>(Vn + 4 * k) = make_double2(r[0][sp], r[1][sp]);
            *reinterpret_cast<double2 *>(Vn + 4 * k + 2) = make_double2(r[2][sp], r[3][sp]);
            // rows of R (columns left of a row's first stored column are skipped; j0 is a multiple of 4, so rows j0, j0+1
            // start at column j0 and rows j0+2, j0+3 at column j0+2)
            R[dch_row(j0, Pp) + k] = r[0][sp];
            R[dch_row(j0 + 1, Pp) + k] = r[1][sp];
            if (k >= j0 + 2) { R[dch_row(j0 + 2, Pp) + k] = r[2][sp]; R[dch_row(j0 + 3, Pp) + k] = r[3][sp]; }
        }
    }
    if (lane == 0) {
        sc[0] = y0; sc[1] = y1; sc[2] = y2; sc[3] = y3;
        y[j0] = y0; y[j0 + 1] = y1; y[j0 + 2] = y2; y[j0 + 3] = y3;
        pinv[j0] = p0; pinv[j0 + 1] = p1; pinv[j0 + 2] = p2; pinv[j0 + 3] = p3;
    }
}

// S: P x P row-major (both triangles), b: P.  sm: dch_smem_bytes(P) of shared memory.  All DCH_THREADS threads call it.
template <int NPASS>
__device__ __forceinline__ void dch_solve_impl(const double *__restrict__ S, const double *__restrict__ b, double lambda, int P,
                                               double *__restrict__ x, int *info, double *sm) {
    const int Pp = (P + 3) & ~3;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5, g = lane >> 2, q = lane & 3;
    double *R = sm;                                   // pair-packed upper triangle
    double *V = R + (size_t)Pp * Pp / 2 + Pp;        // [2][Pp][4]
    double *sc = V + 2 * 4 * (size_t)Pp;              // [2][16]: y of the panel rows
    double *y = sc + 32;                              // [Pp]
    double *pinv = y + Pp;                            // [Pp] reciprocal pivots
    // ---- load: upper triangle of S + lambda I (identity on the padding), rhs
    for (int i = warp; i < Pp; i += nw) {
        const int c0 = i & ~1, ro = dch_row(i, Pp);
        for (int k = c0 + lane; k < Pp; k += 32) {
            double v = 0.0;
            if (i < P && k < P) v = S[(size_t)i * P + k] + (i == k ? lambda : 0.0);
            else if (i == k) v = 1.0;
            R[ro + k] = v;
        }
    }
    for (int i = tid; i < Pp; i += nt) y[i] = i < P ? b[i] : 0.0;
    __syncthreads();
    if (warp == 0) {
        double r[4][NPASS], ry[4], blk[10];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const int ro = dch_row(qq, Pp), c0 = qq & ~1;
#pragma unroll
            for (int sp = 0; sp < NPASS; ++sp) {
                const int k = lane + 32 * sp;
                r[qq][sp] = (k < Pp && k >= c0) ? R[ro + k] : 0.0;
            }
            ry[qq] = y[qq];
        }
        {
            int e = 0;
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 <= a2; ++b2) blk[e++] = R[dch_row(b2, Pp) + a2];  // (a2, b2) = (b2, a2) of the stored upper triangle
        }
        dch_panel_factor<NPASS>(r, ry, blk, 0, lane, Pp, R, V, sc, y, pinv, info);
    }
    __syncthreads();
    const int np = Pp / 4;
    for (int pnl = 0; pnl < np; ++pnl) {
        const int j0 = 4 * pnl, k0 = j0 + 4, cur = pnl & 1;
        const double *Vc = V + cur * 4 * Pp, *scc = sc + cur * 16;
        if (warp == 0) {
            if (k0 < Pp) {
                double vk[NPASS][4], r[4][NPASS], ry[4], fq[4][4], blk[10];
#pragma unroll
                for (int sp = 0; sp < NPASS; ++sp) {
                    const int k = k0 + lane + 32 * sp;
                    if (k < Pp) {
                        const double2 a = *reinterpret_cast<const double2 *>(Vc + 4 * k), b2 = *reinterpret_cast<const double2 *>(Vc + 4 * k + 2);
                        vk[sp][0] = a.x; vk[sp][1] = a.y; vk[sp][2] = b2.x; vk[sp][3] = b2.y;
                    } else {
                        vk[sp][0] = vk[sp][1] = vk[sp][2] = vk[sp][3] = 0.0;
                    }
                }
                const double y0 = scc[0], y1 = scc[1], y2 = scc[2], y3 = scc[3];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const double *fp = Vc + 4 * (k0 + qq);
                    const double2 f01 = *reinterpret_cast<const double2 *>(fp), f23 = *reinterpret_cast<const double2 *>(fp + 2);
                    fq[qq][0] = f01.x; fq[qq][1] = f01.y; fq[qq][2] = f23.x; fq[qq][3] = f23.y;
                    const int ro = dch_row(k0 + qq, Pp), c0 = (k0 + qq) & ~1;
#pragma unroll
                    for (int sp = 0; sp < NPASS; ++sp) {
                        const int k = k0 + lane + 32 * sp;
                        r[qq][sp] = (k < Pp && k >= c0) ? R[ro + k] - ((f01.x * vk[sp][0] + f01.y * vk[sp][1]) + (f23.x * vk[sp][2] + f23.y * vk[sp][3])) : 0.0;
                    }
                    ry[qq] = y[k0 + qq] - ((f01.x * y0 + f01.y * y1) + (f23.x * y2 + f23.y * y3));
                }
                {
                    int e = 0;
#pragma unroll
                    for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                        for (int b2 = 0; b2 <= a2; ++b2)
                            blk[e++] = R[dch_row(k0 + b2, Pp) + k0 + a2] -
                                       ((fq[a2][0] * fq[b2][0] + fq[a2][1] * fq[b2][1]) + (fq[a2][2] * fq[b2][2] + fq[a2][3] * fq[b2][3]));
                }
                dch_panel_factor<NPASS>(r, ry, blk, k0, lane, Pp, R, V + (cur ^ 1) * 4 * Pp, sc + (cur ^ 1) * 16, y, pinv, info);
            }
        } else {
            const int r_lo = k0 + 4;                      // first trailing row below the next panel
            const int nb = (Pp - r_lo + 7) >> 3;          // 8-row bands
            const int wk = warp - 1, nwk = nw - 1;
            // right-hand side: y[i] -= V[i][0..3] . y_panel
            const double y0 = scc[0], y1 = scc[1], y2 = scc[2], y3 = scc[3];
            for (int i = r_lo + wk * 32 + lane; i < Pp; i += nwk * 32) {
                const double2 f01 = *reinterpret_cast<const double2 *>(Vc + 4 * i), f23 = *reinterpret_cast<const double2 *>(Vc + 4 * i + 2);
                y[i] -= (f01.x * y0 + f01.y * y1) + (f23.x * y2 + f23.y * y3);
            }
            for (int bnd = wk; bnd < nb; bnd += nwk) {
                const int rb0 = r_lo + 8 * bnd, i = rb0 + g;
                const bool rv = i < Pp;
                const int ir = rv ? i : Pp - 1;
                double *C = R + dch_row(ir, Pp);
                const int cfirst = ir & ~1;               // first stored column of this lane's row
                const double a = rv ? -Vc[4 * i + q] : 0.0;
                const int nct = (Pp - rb0 + 7) >> 3;      // column tiles from column rb0 (tiles left of it are below the diagonal)
                for (int ct0 = 0; ct0 < nct; ct0 += 4) {
                    double bb[4];
                    double2 c[4];
                    bool cv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int kb = rb0 + 8 * (ct0 + u) + g, kc = rb0 + 8 * (ct0 + u) + 2 * q;
                        cv[u] = rv && kc < Pp && kc >= cfirst;
                        bb[u] = Vc[4 * min(kb, Pp - 1) + q];
                        if (kb >= Pp) bb[u] = 0.0;
                        c[u] = cv[u] ? *reinterpret_cast<const double2 *>(C + kc) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) dch_dmma(c[u].x, c[u].y, a, bb[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cv[u]) *reinterpret_cast<double2 *>(C + rb0 + 8 * (ct0 + u) + 2 * q) = c[u];
                }
            }
        }
        __syncthreads();
    }
    // ---- back-substitution R x = y, 4 rows at a time from the bottom; x overwrites y
    for (int j0 = Pp - 4; j0 >= 0; j0 -= 4) {
        if (tid == 0) {
            // rows j0..j0+3 of R: the 4x4 upper triangle on the diagonal, reciprocal pivots kept from the factorisation
            const double *r0 = R + dch_row(j0, Pp), *r1 = R + dch_row(j0 + 1, Pp), *r2 = R + dch_row(j0 + 2, Pp);
            const double x3 = y[j0 + 3] * pinv[j0 + 3];
            const double x2 = (y[j0 + 2] - r2[j0 + 3] * x3) * pinv[j0 + 2];
            const double x1 = (y[j0 + 1] - r1[j0 + 2] * x2 - r1[j0 + 3] * x3) * pinv[j0 + 1];
            const double x0 = (y[j0] - r0[j0 + 1] * x1 - r0[j0 + 2] * x2 - r0[j0 + 3] * x3) * pinv[j0];
            y[j0] = x0; y[j0 + 1] = x1; y[j0 + 2] = x2; y[j0 + 3] = x3;
        }
        __syncthreads();
        const double x0 = y[j0], x1 = y[j0 + 1], x2 = y[j0 + 2], x3 = y[j0 + 3];
        for (int i = tid; i < j0; i += nt) {
            const double *ri = R + dch_row(i, Pp) + j0;   // 4 consecutive columns, 16-byte aligned (j0 multiple of 4, row start even)
            const double2 a01 = *reinterpret_cast<const double2 *>(ri), a23 = *reinterpret_cast<const double2 *>(ri + 2);
            y[i] -= (a01.x * x0 + a01.y * x1) + (a23.x * x2 + a23.y * x3);
        }
        __syncthreads();
    }
    for (int i = tid; i < P; i += nt) x[i] = y[i];
}

__device__ __forceinline__ void dch_solve(const double *__restrict__ S, const double *__restrict__ b, double lambda, int P, double *__restrict__ x,
                                          int *info, double *sm) {
    const int Pp = (P + 3) & ~3;
    if (Pp <= 128) dch_solve_impl<4>(S, b, lambda, P, x, info, sm);
    else if (Pp <= 160) dch_solve_impl<5>(S, b, lambda, P, x, info, sm);
    else dch_solve_impl<6>(S, b, lambda, P, x, info, sm);
}

// single system (vio_solve on a dense sliding-window problem).  MAXT = 256 (176 registers) or 512 (128 registers: more worker
// warps for the trailing update, a few spills on the look-ahead warp)
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_dense_chol_blocked(const double *__restrict__ S, const double *__restrict__ b, double lambda,
                                                                       int P, double *__restrict__ x, int *info, const double *lam_p = nullptr) {
    extern __shared__ __align__(16) double dch_sm[];
    if (lam_p) lambda = *lam_p;
    if (threadIdx.x == 0) *info = 0;
    dch_solve(S, b, lambda, P, x, info, dch_sm);
}

// one CTA per problem of a lock-step batch: (S_k + lambda_k I) dx_k = bS_k
__global__ void __launch_bounds__(DCH_THREADS, 1) k_chol_batch_blocked(const double *__restrict__ S, const double *__restrict__ b,
                                                                       const double *lambdas, const uint8_t *act, int P, double *__restrict__ x,
                                                                       int *info) {
    extern __shared__ __align__(16) double dch_sm[];
    const int k = blockIdx.x;
    if (act && !act[k]) return;
    dch_solve(S + (size_t)k * P * P, b + (size_t)k * P, lambdas[k], P, x + (size_t)k * P, info, dch_sm);
}
