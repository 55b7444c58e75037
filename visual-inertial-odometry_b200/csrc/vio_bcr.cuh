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

#ifndef BCR_THREADS
#define BCR_THREADS 384
#endif

struct BcrView {
    int n, M, ld, n_items;  // ld: row stride of a tile (BcrPlan::ld)
    int nbuf;            // 7 or 5 operand tiles in shared memory (see k_bcr_run)
    const BcrItem *items;
    double *pool;        // [n_slots][M*M]
    double *bv, *xv;     // [n][M]
    unsigned *flags;     // [n_items] = epoch when the item is complete
    unsigned *counter;   // work queue head (zeroed before the launch)
    int first_item;      // this launch executes items [first_item, n_items)
    double *xpool;       // BCR_EXPORT target tiles (the interface system of the multi-GPU solve), else nullptr
    unsigned epoch;
    int *info;           // != 0: a pivot was not positive
    unsigned long long *prof;  // optional [8] cycle counters (VIO_B200_PROFILE=1): dependency wait, updates + couplings, Cholesky,
                               // W products + stores, back-substitution items, kept-node items, #eliminations, #items
};

// ---- loader ------------------------------------------------------------------------------------------------------
// Only the upper block triangle of S (and the upper element triangle of its diagonal blocks) has to be valid: a lower
// block is read as the transpose of its partner `tr` (BSR blocks are stored row-major, so tr[k] < k marks a lower block) -
// the k_mirror_bsr pass over the 60 MB of S is not needed in front of this solver.
__global__ void __launch_bounds__(256) k_bcr_load(const double *__restrict__ val, const long long *__restrict__ dst, const int *__restrict__ tr, long long nnzb,
                                                  const double *__restrict__ b, const int *__restrict__ blk_node,
                                                  const int *__restrict__ blk_loc, const int *__restrict__ node_size, int nb, int n, int M, int LD,
                                                  double lambda, double *__restrict__ pool, double *__restrict__ bv, const double *lam_p = nullptr) {
    if (lam_p) lambda = *lam_p;
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // one thread per (block, row): 32-bit index arithmetic (a 64-bit division per element used to dominate this kernel), the
    // six values of a row are independent loads
    for (long long t = t0; t < nnzb * 6; t += stride) {
        const unsigned tu = (unsigned)t;  // nnzb * 6 < 2^32 (checked by the plan)
        const unsigned k = tu / 6u;
        const int r = (int)(tu - 6u * k);
        const long long d = dst[k];
        if (d < 0) continue;
        const bool dg = (d & BCR_DST_DIAG) != 0;
        const unsigned kt = (unsigned)tr[k];
        double x[6];
        if (kt < k) {
            const double *src = val + 36 * (long long)kt + r;
#pragma unroll
            for (int c = 0; c < 6; ++c) x[c] = src[6 * c];
        } else {
            const double *src = val + 36 * (long long)k;
#pragma unroll
            for (int c = 0; c < 6; ++c) x[c] = (kt == k && r > c) ? src[6 * c + r] : src[6 * r + c];
        }
        double *out = pool + (d & ~BCR_DST_DIAG) + (long long)r * LD;
#pragma unroll
        for (int c = 0; c < 6; ++c) out[c] = x[c] + ((dg && c == r) ? lambda : 0.0);
    }
    for (long long t = t0; t < (long long)n * M; t += stride) {  // identity padding of ragged nodes
        const int a = (int)(t / M), q = (int)(t % M);
        if (q >= 6 * node_size[a]) pool[(long long)a * M * LD + (long long)q * LD + q] = 1.0;
    }
    for (long long t = t0; t < 6LL * nb; t += stride) {
        const int i = (int)(t / 6), c = (int)(t % 6);
        if (blk_node[i] >= 0) bv[(long long)blk_node[i] * M + 6 * blk_loc[i] + c] = b[t];
    }
}

// ---- finish: gather x, solve the isolated 6x6 blocks ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bcr_finish(const double *__restrict__ xv, const int *__restrict__ blk_gnode, const int *__restrict__ blk_node,
                                                    const int *__restrict__ blk_loc, int nb, int M, int own_hi, int do_iso,
                                                    const double *__restrict__ val, const int *__restrict__ diag,
                                                    const double *__restrict__ b, double lambda, double *__restrict__ x, int *info,
                                                    const double *lam_p = nullptr) {
    if (lam_p) lambda = *lam_p;
    // blk_gnode: node of the global partition (-1: isolated block); blk_node: node index into xv (multi-GPU: the rank's LOCAL
    // node, owned when < own_hi; other blocks are left untouched - x was zeroed and is summed over the ranks afterwards)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    if (blk_gnode[i] >= 0) {
        if (blk_node[i] >= 0 && blk_node[i] < own_hi) {
#pragma unroll
            for (int c = 0; c < 6; ++c) x[6 * (size_t)i + c] = xv[(size_t)blk_node[i] * M + 6 * blk_loc[i] + c];
        }
        return;
    }
    if (!do_iso) return;
    double A[36], y[6];
    const double *d = val + 36 * (size_t)diag[i];
#pragma unroll
    for (int e = 0; e < 36; ++e) {  // only the upper element triangle of a diagonal block needs to be valid (see k_bcr_load)
        const int r = e / 6, c = e % 6;
        A[e] = (r > c ? d[6 * c + r] : d[e]) + (r == c ? lambda : 0.0);
    }
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
// A tile is M rows x LD doubles (row stride LD >= M, see BcrPlan::ld); TS = M * LD elements.
__device__ __forceinline__ void bcr_load_tile(double *dst, const double *src, int M, int LD, bool transpose) {
    const int TS = M * LD;
    if (!transpose) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (int t = threadIdx.x; t < (TS >> 1); t += blockDim.x) d2[t] = __ldcg(s2 + t);
    } else {
        for (int t = threadIdx.x; t < TS; t += blockDim.x) {
            const int r = t / LD, c = t - r * LD;
            if (c < M) dst[c * LD + r] = __ldcg(src + t);
        }
    }
}
__device__ __forceinline__ void bcr_store_tile(double *dst, const double *src, int TS) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (int t = threadIdx.x; t < (TS >> 1); t += blockDim.x) __stcg(d2 + t, s2[t]);
}

// ---- FP64 tensor-core products -------------------------------------------------------------------------------------
// mma.sync.m8n8k4.f64 (DMMA): C(8x8) += A(8x4) B(4x8) per warp instruction, i.e. 256 FMAs for three operand registers -
// the dense M x M products run on the FP64 pipe at its full rate (37 TFLOP/s measured on B200 against 34 for DFMA) with
// a quarter of the shared-memory traffic and a tenth of the instructions of a register-tiled DFMA loop.
//   lane l = 4 g + q:  a = A[g][q],  b = B[q][g],  c0, c1 = C[g][2q], C[g][2q+1]
__device__ __forceinline__ void bcr_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#define BCR_MAXTC ((BCR_MAX_M + 7) / 8)  // 8-wide tile columns of a node tile

// C1 = A^T B1 [and C2 = A^T B2] for row-major M x M tiles in shared memory: both operands are read along rows, the
// fragments are 4 rows x 8 consecutive columns (conflict free with the padded row stride LD).  A warp owns a band of 8
// rows of the result and all its tile columns.  TRI: A is upper triangular (A[r][i] = 0 for r > i): the band stops at its
// diagonal.  epi(i, j, v0, v1) receives elements (i, j), (i, j + 1) (j even).
template <bool TRI, bool PAIR, int TC, class Epi1, class Epi2>
__device__ __forceinline__ void bcr_mma_tn_tc(const double *__restrict__ A, const double *__restrict__ B1, const double *__restrict__ B2, int M, int LD,
                                              Epi1 epi1, Epi2 epi2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, q = lane & 3;
    // TC = ceil(M / 8) is a compile-time constant: the tile loop is fully unrolled WITHOUT branches, so the fragment loads
    // of a k-step are all issued before its first DMMA (a guarded loop serialises load latency and DMMA per tile)
    bool cv[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) cv[t] = 8 * t + g < M;
    for (int ti = warp; ti < TC; ti += nw) {
        const int i0 = 8 * ti;
        const bool rv = i0 + g < M;
        double c1[TC][2], c2[PAIR ? TC : 1][2];
#pragma unroll
        for (int t = 0; t < TC; ++t) { c1[t][0] = c1[t][1] = 0.0; if (PAIR) { c2[t][0] = c2[t][1] = 0.0; } }
        const int rend = TRI ? min(M, i0 + 8) : M;
        const double *ap = A + q * LD + (rv ? i0 + g : 0), *b1p = B1 + q * LD + g, *b2p = B2 + q * LD + g;
#pragma unroll 1
        for (int r0 = 0; r0 < rend; r0 += 4, ap += 4 * LD, b1p += 4 * LD, b2p += 4 * LD) {
            double a = *ap, b1[TC], b2[PAIR ? TC : 1];
            if (!rv) a = 0.0;
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                // the last tile may hang over the edge of the row: the address is clamped, the value masked
                b1[t] = b1p[cv[t] ? 8 * t : 0];
                if (PAIR) b2[t] = b2p[cv[t] ? 8 * t : 0];
            }
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                bcr_dmma(c1[t][0], c1[t][1], a, cv[t] ? b1[t] : 0.0);
                if (PAIR) bcr_dmma(c2[t][0], c2[t][1], a, cv[t] ? b2[t] : 0.0);
            }
        }
        if (rv) {
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                const int j = 8 * t + 2 * q;
                if (j < M) {
                    epi1(i0 + g, j, c1[t][0], c1[t][1]);
                    if (PAIR) epi2(i0 + g, j, c2[t][0], c2[t][1]);
                }
            }
        }
    }
}
template <bool TRI, bool PAIR, class Epi1, class Epi2>
__device__ __forceinline__ void bcr_mma_tn(const double *__restrict__ A, const double *__restrict__ B1, const double *__restrict__ B2, int M, int LD,
                                           Epi1 epi1, Epi2 epi2) {
    switch ((M + 7) >> 3) {  // M is a multiple of 12 up to BCR_MAX_M
        case 2: bcr_mma_tn_tc<TRI, PAIR, 2>(A, B1, B2, M, LD, epi1, epi2); break;
        case 3: bcr_mma_tn_tc<TRI, PAIR, 3>(A, B1, B2, M, LD, epi1, epi2); break;
        case 5: bcr_mma_tn_tc<TRI, PAIR, 5>(A, B1, B2, M, LD, epi1, epi2); break;
        case 6: bcr_mma_tn_tc<TRI, PAIR, 6>(A, B1, B2, M, LD, epi1, epi2); break;
        case 8: bcr_mma_tn_tc<TRI, PAIR, 8>(A, B1, B2, M, LD, epi1, epi2); break;
        default: bcr_mma_tn_tc<TRI, PAIR, 9>(A, B1, B2, M, LD, epi1, epi2); break;
    }
}

// v[i] -= sum_r A[r][i] * y[r]   (A: shared tile, y: shared vector).  Four threads per element (r = part mod 4), combined
// through `scratch` ([4][M]); contains two barriers - every thread of the CTA must call it.
__device__ __forceinline__ void bcr_gemv_t_sub(const double *A, const double *y, double *v, double *scratch, int M, int LD) {
    for (int t = threadIdx.x; t < 4 * M; t += blockDim.x) {
        const int part = t / M, i = t - part * M;
        double a = 0.0;
        for (int r = part; r < M; r += 4) a += A[r * LD + i] * y[r];
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
//   * the other warps apply the current panel to everything else as rank-4 DMMA updates on 8x8 tiles,
//     C[rows][cols] -= F[rows][0..3] V[cols][0..3]^T: for the trailing rows of D the multipliers F are rows of V, for the
//     rows of U they are columns j0..j0+3 of U after the panel's 4x4 triangular transform (done in registers by the
//     lanes that need them, written back once).
// One barrier per panel (M/4 panels); V and the panel scalars are double buffered.
#define BCR_PANEL 4
#define BCR_CHUNK 8  // column tiles of the trailing update in flight per band

// Factorise 4 panel rows held in registers (r[q][sp] = row j0+q, column j0 + lane + 32 sp).  blk = the panel's 4x4
// leading block (lower triangle: 00 10 11 20 21 22 30 31 32 33), up to date and identical in every lane: each lane
// factorises it redundantly in registers, so the critical path (4 dependent reciprocal square roots) has no shuffles and
// no shared-memory round trips; the lane's own columns follow with independent FMAs.
// Publishes V'[k][0..3] (scaled rows), sc[0..3] = 1/sqrt(pivot), sc[4..9] = L[j0+q][a] for a < q in the order
// (0,1) (0,2) (1,2) (0,3) (1,3) (2,3).
template <int NPASS>
__device__ __forceinline__ void bcr_panel_factor(double (&r)[BCR_PANEL][NPASS], const double (&blk)[10], int j0, int lane, int M,
                                                 double *__restrict__ Vn, double *__restrict__ sc, int *info) {
    // 4x4 Cholesky of the leading block
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
    // own columns: v_0 = r_0 p0 ; v_1 = (r_1 - l10 v_0) p1 ; ...   (columns left of the diagonal come out as the zeros of L^T)
#pragma unroll
    for (int sp = 0; sp < NPASS; ++sp) {
        const double v0 = r[0][sp] * p0;
        const double v1 = (r[1][sp] - l10 * v0) * p1;
        const double v2 = (r[2][sp] - l20 * v0 - l21 * v1) * p2;
        const double v3 = (r[3][sp] - l30 * v0 - l31 * v1 - l32 * v2) * p3;
        r[0][sp] = v0; r[1][sp] = v1; r[2][sp] = v2; r[3][sp] = v3;
    }
    // exact zeros left of the diagonal inside the panel (rounding leaves ~1e-17 there; they are never read as multipliers,
    // but V feeds the DMMA updates as a whole)
    if (lane < 1) r[1][0] = 0.0;
    if (lane < 2) r[2][0] = 0.0;
    if (lane < 3) r[3][0] = 0.0;
#pragma unroll
    for (int sp = 0; sp < NPASS; ++sp) {
        const int k = j0 + lane + 32 * sp;
        if (k < M) {
            *reinterpret_cast<double2 *>(Vn + 4 * k) = make_double2(r[0][sp], r[1][sp]);
            *reinterpret_cast<double2 *>(Vn + 4 * k + 2) = make_double2(r[2][sp], r[3][sp]);
        }
    }
    if (lane == 0) {
        sc[0] = p0; sc[1] = p1; sc[2] = p2; sc[3] = p3;
        sc[4] = l10; sc[5] = l20; sc[6] = l21; sc[7] = l30; sc[8] = l31; sc[9] = l32;
    }
}

template <int NPASS>  // ceil(M / 32)
__device__ __forceinline__ void bcr_chol_inv(double *__restrict__ D, double *__restrict__ U, double *__restrict__ V /* [2][M][4] */,
                                             double *__restrict__ sc /* [2][16] */, int M, int LD, int *info, unsigned long long *prof) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5, g = lane >> 2, q = lane & 3;
    long long cw0 = 0, cw1 = 0, cb0 = 0, cb1 = 0;
    for (int t = tid; t < M * LD; t += nt) U[t] = (t / LD == t % LD) ? 1.0 : 0.0;
    if (warp == 0) {
        double r[BCR_PANEL][NPASS];
#pragma unroll
        for (int qq = 0; qq < BCR_PANEL; ++qq)
#pragma unroll
            for (int sp = 0; sp < NPASS; ++sp) {
                const int k = lane + 32 * sp;
                r[qq][sp] = k < M ? D[qq * LD + k] : 0.0;
            }
        double blk[10];
        {
            int e = 0;
#pragma unroll
            for (int a2 = 0; a2 < BCR_PANEL; ++a2)
#pragma unroll
                for (int b2 = 0; b2 <= a2; ++b2) blk[e++] = D[a2 * LD + b2];
        }
        bcr_panel_factor<NPASS>(r, blk, 0, lane, M, V, sc, info);
    }
    __syncthreads();
    const int np = M / BCR_PANEL;
    for (int pnl = 0; pnl < np; ++pnl) {
        const int j0 = BCR_PANEL * pnl, k0 = j0 + BCR_PANEL, cur = pnl & 1;
        const long long tA = prof ? clock64() : 0;
        const double *Vc = V + cur * 4 * M, *scc = sc + cur * 16;
        if (warp == 0) {
            if (k0 < M) {
                double vk[NPASS][BCR_PANEL], r[BCR_PANEL][NPASS];
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
                double fq[BCR_PANEL][BCR_PANEL];  // V[k0+a][q]: the panel rows' own multipliers (the same in every lane)
#pragma unroll
                for (int qq = 0; qq < BCR_PANEL; ++qq) {
                    const double *fp = Vc + 4 * (k0 + qq);
                    const double2 f01 = *reinterpret_cast<const double2 *>(fp), f23 = *reinterpret_cast<const double2 *>(fp + 2);
                    fq[qq][0] = f01.x; fq[qq][1] = f01.y; fq[qq][2] = f23.x; fq[qq][3] = f23.y;
#pragma unroll
                    for (int sp = 0; sp < NPASS; ++sp) {
                        const int k = k0 + lane + 32 * sp;
                        r[qq][sp] = k < M ? D[(k0 + qq) * LD + k] - ((f01.x * vk[sp][0] + f01.y * vk[sp][1]) + (f23.x * vk[sp][2] + f23.y * vk[sp][3])) : 0.0;
                    }
                }
                double blk[10];
                {
                    int e = 0;
#pragma unroll
                    for (int a2 = 0; a2 < BCR_PANEL; ++a2)
#pragma unroll
                        for (int b2 = 0; b2 <= a2; ++b2)
                            blk[e++] = D[(k0 + a2) * LD + k0 + b2] -
                                       ((fq[a2][0] * fq[b2][0] + fq[a2][1] * fq[b2][1]) + (fq[a2][2] * fq[b2][2] + fq[a2][3] * fq[b2][3]));
                }
                bcr_panel_factor<NPASS>(r, blk, k0, lane, M, V + (cur ^ 1) * 4 * M, sc + (cur ^ 1) * 16, info);
            }
        } else if ((warp & 3) != 0) {
            // workers = the warps that do NOT share warp 0's scheduler (warp id mod 4 picks the SM sub-partition): a DMMA holds
            // the FP64 pipe of its sub-partition for 16 cycles, and every one issued next to warp 0 lengthens the dependent
            // DFMA chain of the look-ahead factorisation, which is the critical path of the whole panel loop
            const int wk = warp - 1 - (warp >> 2), nwk = nw - ((nw + 3) >> 2);  // worker index / count
            const int nD = max(0, M - k0 - BCR_PANEL);   // trailing rows of D below the next panel: rows k0+4 ..
            const int nbD = (nD + 7) >> 3, nbU = (k0 + 7) >> 3;  // 8-row bands of D, and of rows 0 .. j0+3 of U
            const int nct = (M - k0 + 7) >> 3;               // 8-column tiles from column k0
            const double p0 = scc[0], p1 = scc[1], p2 = scc[2], p3 = scc[3];
            const double s01 = scc[4], s02 = scc[5], s12 = scc[6], s03 = scc[7], s13 = scc[8], s23 = scc[9];
            for (int bnd = wk; bnd < nbD + nbU; bnd += nwk) {
                double *C;
                double a;  // this lane's multiplier F[g][q], negated
                bool rv;
                if (bnd < nbD) {
                    const int i = k0 + BCR_PANEL + 8 * bnd + g;
                    rv = i < M;
                    C = D + (rv ? i : M - 1) * LD;
                    a = rv ? -Vc[4 * i + q] : 0.0;
                } else {
                    const int c = 8 * (bnd - nbD) + g;
                    rv = c < k0;
                    C = U + (rv ? c : 0) * LD;
                    // lanes without a row read nothing (a dummy read of row 0 would race with the band that owns it)
                    const double2 zz = make_double2(0.0, 0.0);
                    const double2 u01 = rv ? *reinterpret_cast<const double2 *>(C + j0) : zz, u23 = rv ? *reinterpret_cast<const double2 *>(C + j0 + 2) : zz;
                    const double g0 = u01.x * p0;
                    const double g1 = (u01.y - s01 * g0) * p1;
                    const double g2 = (u23.x - s02 * g0 - s12 * g1) * p2;
                    const double g3 = (u23.y - s03 * g0 - s13 * g1 - s23 * g2) * p3;
                    const double gq = q == 0 ? g0 : (q == 1 ? g1 : (q == 2 ? g2 : g3));
                    __syncwarp();  // the four lanes of a row have read columns j0..j0+3 before they are overwritten
                    if (rv) C[j0 + q] = gq;
                    a = rv ? -gq : 0.0;
                }
                // column tiles BCR_CHUNK at a time, loads before the DMMAs before the stores (no branches in between)
                for (int ct0 = 0; ct0 < nct; ct0 += BCR_CHUNK) {
                    double b[BCR_CHUNK];
                    double2 c[BCR_CHUNK];
                    bool cv[BCR_CHUNK];
#pragma unroll
                    for (int u = 0; u < BCR_CHUNK; ++u) {
                        const int kb = k0 + 8 * (ct0 + u) + g, kc = k0 + 8 * (ct0 + u) + 2 * q;
                        cv[u] = rv && kc < M;
                        b[u] = Vc[4 * min(kb, M - 1) + q];
                        if (kb >= M) b[u] = 0.0;
                        c[u] = cv[u] ? *reinterpret_cast<const double2 *>(C + kc) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < BCR_CHUNK; ++u) bcr_dmma(c[u].x, c[u].y, a, b[u]);
#pragma unroll
                    for (int u = 0; u < BCR_CHUNK; ++u)
                        if (cv[u]) *reinterpret_cast<double2 *>(C + k0 + 8 * (ct0 + u) + 2 * q) = c[u];
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
    const int M = s.M, LD = s.ld, MM = M * LD, tid = threadIdx.x, nt = blockDim.x;  // MM: elements of a tile (M rows x LD)
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
        const unsigned idx = idx_s + (unsigned)s.first_item;
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
        if (it.kind & BCR_EXPORT) {
            // open chain: the coupling left between the two pinned ends, -pool[a]^T pool[b], goes to the interface system
            if (tid == 0) {
                bcr_fence_async();
                bcr_mbar_expect(&ldbar, 2u * tile_bytes);
                bcr_bulk_load(BCR_OPA(0), s.pool + (size_t)it.cl_a * MM, tile_bytes, &ldbar);
                bcr_bulk_load(BCR_OPB(0), s.pool + (size_t)it.cl_b * MM, tile_bytes, &ldbar);
            }
            bcr_mbar_wait(&ldbar, ph);
            ph ^= 1u;
            __syncthreads();
            double *out = s.xpool + (size_t)it.cl_slot * MM;
            auto st = [&](int i, int j, double v0, double v1) { __stcg(reinterpret_cast<double2 *>(out + (size_t)i * LD + j), make_double2(-v0, -v1)); };
            bcr_mma_tn<false, false>(BCR_OPA(0), BCR_OPB(0), BCR_OPB(0), M, LD, st, st);
        } else if (it.kind & BCR_BACKSUB) {
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
                    if (hl) a += Wl[i * LD + c] * ye0[c];
                    if (hr) a += Wr[i * LD + c] * ye1[c];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) vv[i] = bk[i] - a;
            }
            __syncthreads();
            for (int i = warp; i < M; i += nw) {
                double a = 0.0;
                for (int r = i + lane; r < M; r += 32) a += Dm[i * LD + r] * vv[r];
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
            if (elim) BCR_MARK(12);
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
                if (mode[sd] == 1 && cb[sd] != 0) bcr_load_tile(OUT, s.pool + (size_t)ca[sd] * MM, M, LD, true);  // rare: transposed coupling
                auto upd = [&](int i, int j, double v0, double v1) {
                    double2 *d = reinterpret_cast<double2 *>(Dm + i * LD + j);
                    double2 c = *d;
                    c.x -= v0; c.y -= v1;
                    *d = c;
                };
                auto neg_out = [&](int i, int j, double v0, double v1) { *reinterpret_cast<double2 *>(OUT + i * LD + j) = make_double2(-v0, -v1); };
                if (fused[sd]) {
                    if (elim) BCR_MARK(15);
                    bcr_mma_tn<false, true>(A1, A1, A2, M, LD, upd, neg_out);
                    __syncthreads();
                    if (elim) BCR_MARK(13);
                    bcr_gemv_t_sub(A1, ye, bk, vv, M, LD);
                    if (elim) BCR_MARK(14);
                } else if (us[sd] >= 0) {
                    bcr_mma_tn<false, false>(A1, A1, A1, M, LD, upd, upd);
                    bcr_gemv_t_sub(A1, ye, bk, vv, M, LD);
                }
                if (mode[sd] == 2 && !fused[sd]) {  // rare: a coupling carried over a level, its factors are not this level's W
                    __syncthreads();
                    bcr_load_tile(A1, s.pool + (size_t)ca[sd] * MM, M, LD, false);
                    bcr_load_tile(A2, s.pool + (size_t)cb[sd] * MM, M, LD, false);
                    __syncthreads();
                    bcr_mma_tn<false, false>(A1, A2, A2, M, LD, neg_out, neg_out);
                }
            }
            __syncthreads();
            if (!elim) {
                bcr_store_tile(s.pool + node_off, Dm, MM);
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
                if (M <= 64) bcr_chol_inv<2>(Dm, U, vv, vv + 8 * M, M, LD, s.info, s.prof);
                else bcr_chol_inv<3>(Dm, U, vv, vv + 8 * M, M, LD, s.info, s.prof);
                BCR_MARK(2);
                // W_l = U^T X, W_r = U^T Z -> their pool tiles; y = U^T b; U -> the node's tile
                {
                    double *outL = s.pool + (size_t)(hasL ? it.cl_slot : 0) * MM, *outR = s.pool + (size_t)(hasR ? it.cr_slot : 0) * MM;
                    auto stL = [&](int i, int j, double v0, double v1) { __stcg(reinterpret_cast<double2 *>(outL + (size_t)i * LD + j), make_double2(v0, v1)); };
                    auto stR = [&](int i, int j, double v0, double v1) { __stcg(reinterpret_cast<double2 *>(outR + (size_t)i * LD + j), make_double2(v0, v1)); };
                    if (hasL && hasR) bcr_mma_tn<true, true>(U, X, Z, M, LD, stL, stR);
                    else if (hasL) bcr_mma_tn<true, false>(U, X, X, M, LD, stL, stL);
                    else if (hasR) bcr_mma_tn<true, false>(U, Z, Z, M, LD, stR, stR);
                }
                for (int i = tid; i < M; i += nt) {
                    double a = 0.0;
                    for (int r = 0; r <= i; ++r) a += U[r * LD + i] * bk[r];
                    __stcg(s.bv + (size_t)it.node * M + i, a);
                }
                bcr_store_tile(s.pool + node_off, U, MM);
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
