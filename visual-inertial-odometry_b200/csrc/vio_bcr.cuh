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
static_assert(2 * BCR_MAX_M <= BCR_THREADS, "bcr_chol_inv gives every row of D and U its own thread in the pivot phase");

struct BcrView {
    int n, M, n_items;
    const BcrItem *items;
    double *pool;        // [n_slots][M*M]
    double *bv, *xv;     // [n][M]
    unsigned *flags;     // [n_items] = epoch when the item is complete
    unsigned *counter;   // work queue head (zeroed before the launch)
    unsigned epoch;
    int *info;           // != 0: a pivot was not positive
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

// C = A^T B over 4x4 register tiles; TRI: A is upper triangular (A[r][i] = 0 for r > i), the sum stops at the diagonal.
// epi(i0, j0, acc) receives the finished tile.
template <bool TRI, class Epi>
__device__ __forceinline__ void bcr_tn(const double *__restrict__ A, const double *__restrict__ B, int M, Epi epi) {
    const int T = M >> 2;
    for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
        const int ti = t / T, tj = t - ti * T;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        const int rend = TRI ? min(M, 4 * ti + 4) : M;
        const double *ap = A + 4 * ti, *bp = B + 4 * tj;
#pragma unroll 4
        for (int r = 0; r < rend; ++r, ap += M, bp += M) {
            const double2 a01 = *reinterpret_cast<const double2 *>(ap), a23 = *reinterpret_cast<const double2 *>(ap + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(bp), b23 = *reinterpret_cast<const double2 *>(bp + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bw[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bw[b];
        }
        epi(4 * ti, 4 * tj, acc);
    }
}

// fused pair sharing the A operand:  C1 = A^T A (symmetric update) and C2 = A^T B
template <class Epi1, class Epi2>
__device__ __forceinline__ void bcr_tn_pair(const double *__restrict__ A, const double *__restrict__ B, int M, Epi1 epi1, Epi2 epi2) {
    const int T = M >> 2;
    for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
        const int ti = t / T, tj = t - ti * T;
        double c1[4][4], c2[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { c1[a][b] = 0.0; c2[a][b] = 0.0; }
        const double *ap = A + 4 * ti, *aq = A + 4 * tj, *bp = B + 4 * tj;
#pragma unroll 2
        for (int r = 0; r < M; ++r, ap += M, aq += M, bp += M) {
            const double2 a01 = *reinterpret_cast<const double2 *>(ap), a23 = *reinterpret_cast<const double2 *>(ap + 2);
            const double2 q01 = *reinterpret_cast<const double2 *>(aq), q23 = *reinterpret_cast<const double2 *>(aq + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(bp), b23 = *reinterpret_cast<const double2 *>(bp + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, qv[4] = {q01.x, q01.y, q23.x, q23.y}, bw[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { c1[a][b] += av[a] * qv[b]; c2[a][b] += av[a] * bw[b]; }
        }
        epi1(4 * ti, 4 * tj, c1);
        epi2(4 * ti, 4 * tj, c2);
    }
}

// v[i] -= sum_r A[r][i] * y[r]   (A: shared tile, y: shared vector)
__device__ __forceinline__ void bcr_gemv_t_sub(const double *A, const double *y, double *v, int M) {
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        double a0 = 0.0, a1 = 0.0;
        int r = 0;
        for (; r + 1 < M; r += 2) { a0 += A[r * M + i] * y[r]; a1 += A[(r + 1) * M + i] * y[r + 1]; }
        if (r < M) a0 += A[r * M + i] * y[r];
        v[i] -= a0 + a1;
    }
}

// D (shared, symmetric positive definite, destroyed) -> U = L^-T (shared) with D = L L^T, by forward elimination on
// [D | I]: step j scales row j by 1/sqrt(d_jj) and subtracts it from the rows below; the identity part, kept
// transposed, turns into L^-T.  Two barriers per column.
__device__ __forceinline__ void bcr_chol_inv(double *D, double *U, double *v, int M, int *info) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int t = tid; t < M * M; t += nt) U[t] = (t / M == t % M) ? 1.0 : 0.0;
    __syncthreads();
    for (int j = 0; j < M; ++j) {
        const double d = D[j * M + j];
        if (tid == 0 && !(d > 0.0)) *info = j + 1;
        const double p = rsqrt(d > 0.0 ? d : 1.0);
        if (tid < M) {
            if (tid >= j) v[tid] = D[j * M + tid] * p;
        } else if (tid - M <= j && tid - M >= 0) {
            U[(tid - M) * M + j] *= p;
        }
        __syncthreads();
        // rows j+1..M-1 of D and rows 0..j of U, columns j+1..M-1:  row[col] -= f_row * v[col]
        const int len = M - j - 1;
        for (int r = warp; r < M; r += nw) {
            double *row;
            double f;
            if (r < len) { row = D + (j + 1 + r) * M; f = v[j + 1 + r]; }
            else { const int c = r - len; row = U + c * M; f = U[c * M + j]; }
            for (int k = j + 1 + lane; k < M; k += 32) row[k] -= f * v[k];
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void bcr_wait(const unsigned *flag, unsigned epoch) {
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    } while (v != epoch);
}

// ---- the persistent kernel ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BCR_THREADS, 1) k_bcr_run(BcrView s) {
    extern __shared__ double bsm[];
    const int M = s.M, MM = M * M, tid = threadIdx.x, nt = blockDim.x;
    double *Dm = bsm, *A1 = Dm + MM, *A2 = A1 + MM, *X = A2 + MM, *Z = X + MM;
    double *bk = Z + MM, *ye = bk + M, *tv = ye + M, *vv = tv + M;  // 4 x M vectors
    __shared__ BcrItem it_s;
    __shared__ unsigned idx_s;
    for (;;) {
        __syncthreads();  // the previous item's shared-memory traffic is over
        if (tid == 0) idx_s = atomicAdd(s.counter, 1u);
        __syncthreads();
        const unsigned idx = idx_s;
        if (idx >= (unsigned)s.n_items) return;
        if (tid < (int)(sizeof(BcrItem) / sizeof(int))) reinterpret_cast<int *>(&it_s)[tid] = reinterpret_cast<const int *>(s.items + idx)[tid];
        __syncthreads();
        if (tid < 6 && it_s.dep[tid] >= 0) bcr_wait(s.flags + it_s.dep[tid], s.epoch);
        __syncthreads();
        const BcrItem &it = it_s;
        const size_t node_off = (size_t)it.node * MM;
        if (it.kind & BCR_BACKSUB) {
            // x_k = U (y_k - W_l x_l - W_r x_r)
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int i = tid; i < M; i += nt) {
                bk[i] = __ldcg(s.bv + (size_t)it.node * M + i);
                ye[i] = it.left >= 0 ? __ldcg(s.xv + (size_t)it.left * M + i) : 0.0;
                tv[i] = it.right >= 0 ? __ldcg(s.xv + (size_t)it.right * M + i) : 0.0;
            }
            __syncthreads();
            const double *Wl = it.left >= 0 ? s.pool + (size_t)it.cl_slot * MM : nullptr;
            const double *Wr = it.right >= 0 ? s.pool + (size_t)it.cr_slot * MM : nullptr;
            for (int i = warp; i < M; i += nw) {
                double a = 0.0;
                for (int c = lane; c < M; c += 32) {
                    if (Wl) a += __ldcg(Wl + (size_t)i * M + c) * ye[c];
                    if (Wr) a += __ldcg(Wr + (size_t)i * M + c) * tv[c];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) vv[i] = bk[i] - a;
            }
            __syncthreads();
            const double *Uk = s.pool + node_off;
            for (int i = warp; i < M; i += nw) {
                double a = 0.0;
                for (int r = i + lane; r < M; r += 32) a += __ldcg(Uk + (size_t)i * M + r) * vv[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) __stcg(s.xv + (size_t)it.node * M + i, a);
            }
        } else {
            const bool elim = (it.kind & BCR_ELIM) != 0;
            bcr_load_tile(Dm, s.pool + node_off, M, false);
            for (int i = tid; i < M; i += nt) bk[i] = __ldcg(s.bv + (size_t)it.node * M + i);
            // ---- per side: Schur update from the neighbour eliminated one level earlier, and (when this node is being
            // eliminated) the coupling tile with rows = this node:  side 0 -> X, side 1 -> Z
#pragma unroll 1
            for (int side = 0; side < 2; ++side) {
                const int us = it.upd_slot[side];
                const int mode = elim ? (side == 0 ? it.cl_mode : it.cr_mode) : 0;
                const int ca = side == 0 ? it.cl_a : it.cr_a, cb = side == 0 ? it.cl_b : it.cr_b;
                double *OUT = side == 0 ? X : Z;
                if (us < 0 && mode == 0) continue;
                const bool fused = mode == 2 && us >= 0 && ca == us;  // the usual case: update and coupling share W
                __syncthreads();  // A1 / A2 / ye of the other side are no longer read
                if (us >= 0) {
                    bcr_load_tile(A1, s.pool + (size_t)us * MM, M, false);
                    for (int i = tid; i < M; i += nt) ye[i] = __ldcg(s.bv + (size_t)it.upd_node[side] * M + i);
                }
                if (fused) bcr_load_tile(A2, s.pool + (size_t)cb * MM, M, false);
                if (mode == 1) bcr_load_tile(OUT, s.pool + (size_t)ca * MM, M, cb != 0);
                __syncthreads();
                auto upd = [&](int i0, int j0, double (&acc)[4][4]) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) Dm[(i0 + a) * M + j0 + b] -= acc[a][b];
                };
                auto neg_out = [&](int i0, int j0, double (&acc)[4][4]) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) OUT[(i0 + a) * M + j0 + b] = -acc[a][b];
                };
                if (fused) {
                    bcr_tn_pair(A1, A2, M, upd, neg_out);
                    bcr_gemv_t_sub(A1, ye, bk, M);
                } else if (us >= 0) {
                    bcr_tn<false>(A1, A1, M, upd);
                    bcr_gemv_t_sub(A1, ye, bk, M);
                }
                if (mode == 2 && !fused) {  // rare: a coupling carried over a level, its factors are not this level's W
                    __syncthreads();
                    bcr_load_tile(A1, s.pool + (size_t)ca * MM, M, false);
                    bcr_load_tile(A2, s.pool + (size_t)cb * MM, M, false);
                    __syncthreads();
                    bcr_tn<false>(A1, A2, M, neg_out);
                }
            }
            __syncthreads();
            if (!elim) {
                bcr_store_tile(s.pool + node_off, Dm, M);
                for (int i = tid; i < M; i += nt) __stcg(s.bv + (size_t)it.node * M + i, bk[i]);
            } else {
                const bool hasL = it.cl_mode != 0 && !(it.kind & BCR_MERGE), hasR = it.cr_mode != 0;
                if (it.kind & BCR_MERGE) {
                    for (int t = tid; t < MM; t += nt) Z[t] += X[t];
                    __syncthreads();
                }
                double *U = A1;
                bcr_chol_inv(Dm, U, vv, M, s.info);
                // W_l = U^T X, W_r = U^T Z -> their pool tiles; y = U^T b; U -> the node's tile
                if (hasL) {
                    double *out = s.pool + (size_t)it.cl_slot * MM;
                    bcr_tn<true>(U, X, M, [&](int i0, int j0, double (&acc)[4][4]) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            __stcg(reinterpret_cast<double2 *>(out + (size_t)(i0 + a) * M + j0), make_double2(acc[a][0], acc[a][1]));
                            __stcg(reinterpret_cast<double2 *>(out + (size_t)(i0 + a) * M + j0 + 2), make_double2(acc[a][2], acc[a][3]));
                        }
                    });
                }
                if (hasR) {
                    double *out = s.pool + (size_t)it.cr_slot * MM;
                    bcr_tn<true>(U, Z, M, [&](int i0, int j0, double (&acc)[4][4]) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            __stcg(reinterpret_cast<double2 *>(out + (size_t)(i0 + a) * M + j0), make_double2(acc[a][0], acc[a][1]));
                            __stcg(reinterpret_cast<double2 *>(out + (size_t)(i0 + a) * M + j0 + 2), make_double2(acc[a][2], acc[a][3]));
                        }
                    });
                }
                for (int i = tid; i < M; i += nt) {
                    double a = 0.0;
                    for (int r = 0; r <= i; ++r) a += U[r * M + i] * bk[r];
                    __stcg(s.bv + (size_t)it.node * M + i, a);
                }
                bcr_store_tile(s.pool + node_off, U, M);
            }
        }
        // publish: every thread's global stores are done and visible before the flag
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(s.flags + idx), "r"(s.epoch) : "memory");
        }
    }
}
