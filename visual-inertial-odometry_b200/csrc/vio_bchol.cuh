// vio_bchol.cuh — block-sparse Cholesky of the reduced camera system on the 6x6 BSR pattern ("block-Cholesky reduced
// solve", BASELINE config 4): (S + lambda I) = L L^T, then L y = b_S, L^T dx = y.  The symbolic part (fill pattern,
// update map) is host code (vio_bchol.h).  Numeric part: right-looking over block columns by ONE CTA - a band + border
// matrix has a sequential column dependency, each column costs a 6x6 Cholesky, <= ~20 triangular 6x6 solves and
// <= ~210 rank-6 block updates, spread over the CTA's threads; three barriers per column.  Exact (no tolerance), so it
// doubles as the reference for the PCG variants at sizes where the dense solver is too slow.
#pragma once
#include <stdint.h>

struct BcholView {
    int nb;
    const int *colptr, *rowidx;
    double *L;                              // [nnzL][36] row-major 6x6 blocks, column-compressed
    const long long *upd_ptr, *upd_dst;
    const int *upd_a, *upd_b;
};

// L <- lower block triangle of S (+ lambda on the diagonal); fill blocks were zeroed by a memset
__global__ void k_bchol_init(const double *__restrict__ val, const long long *__restrict__ a_to_l, const int *__restrict__ bsr_col,
                             const int *__restrict__ bsr_rowptr, int nb, long long nnzb, double lambda, double *__restrict__ L) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnzb * 36) return;
    const long long k = t / 36;
    const int e = (int)(t % 36);
    const long long dst = a_to_l[k];
    if (dst < 0) return;
    double v = val[t];
    if (e % 7 == 0) {
        // diagonal block <=> its column index equals its row: find the row by searching rowptr
        int lo = 0, hi = nb - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (bsr_rowptr[mid] <= k) lo = mid; else hi = mid - 1; }
        if (bsr_col[k] == lo) v += lambda;
    }
    L[36 * dst + e] = v;
}

__global__ void __launch_bounds__(1024, 1) k_bchol_factor(BcholView Y, int *info) {
    __shared__ double D[36];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *info = 0;
    for (int j = 0; j < Y.nb; ++j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        double *Ljj = Y.L + 36 * (size_t)base;
        if (tid < 36) D[tid] = Ljj[tid];
        __syncthreads();
        if (tid == 0) {  // dense 6x6 Cholesky, lower, in place; upper part zeroed
            for (int c = 0; c < 6; ++c) {
                double d = D[7 * c];
                for (int k = 0; k < c; ++k) d -= D[6 * c + k] * D[6 * c + k];
                if (!(d > 0.0)) *info = j + 1;
                d = sqrt(d);
                D[7 * c] = d;
                for (int r = c + 1; r < 6; ++r) {
                    double t = D[6 * r + c];
                    for (int k = 0; k < c; ++k) t -= D[6 * r + k] * D[6 * c + k];
                    D[6 * r + c] = t / d;
                }
                for (int r = 0; r < c; ++r) D[6 * r + c] = 0.0;
            }
        }
        __syncthreads();
        if (tid < 36) Ljj[tid] = D[tid];
        // L_ij <- L_ij L_jj^-T : one thread per (sub-diagonal block, row)
        for (int t = tid; t < 6 * cnt; t += nt) {
            double *row = Y.L + 36 * (size_t)(base + 1 + t / 6) + 6 * (t % 6);
            double x[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double a = row[c];
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if (k < c) a -= x[k] * D[6 * c + k];
                x[c] = a / D[7 * c];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) row[c] = x[c];
        }
        __syncthreads();
        // trailing update: block (row_a, row_b) -= L_aj L_bj^T for every pair a >= b of the column's sub-diagonal blocks
        const long long q0 = Y.upd_ptr[j], q1 = Y.upd_ptr[j + 1];
        for (long long t = q0 * 36 + tid; t < q1 * 36; t += nt) {
            const long long q = t / 36;
            const int e = (int)(t - q * 36), r = e / 6, c = e % 6;
            const double *La = Y.L + 36 * (size_t)(base + Y.upd_a[q]) + 6 * r;
            const double *Lb = Y.L + 36 * (size_t)(base + Y.upd_b[q]) + 6 * c;
            const double s = La[0] * Lb[0] + La[1] * Lb[1] + La[2] * Lb[2] + La[3] * Lb[3] + La[4] * Lb[4] + La[5] * Lb[5];
            Y.L[36 * (size_t)Y.upd_dst[q] + e] -= s;
        }
        __syncthreads();
    }
}

// x <- (L L^T)^-1 b   (x may alias nothing; one CTA)
__global__ void __launch_bounds__(256, 1) k_bchol_solve(BcholView Y, const double *__restrict__ b, double *__restrict__ x) {
    __shared__ double xj[6];
    __shared__ double part[40][6];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < 6 * Y.nb; t += nt) x[t] = b[t];
    __syncthreads();
    // forward: L y = b, column oriented
    for (int j = 0; j < Y.nb; ++j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        const double *D = Y.L + 36 * (size_t)base;
        if (tid == 0) {
            double y[6];
            for (int c = 0; c < 6; ++c) {
                double a = x[6 * (size_t)j + c];
                for (int k = 0; k < c; ++k) a -= D[6 * c + k] * y[k];
                y[c] = a / D[7 * c];
            }
            for (int c = 0; c < 6; ++c) { xj[c] = y[c]; x[6 * (size_t)j + c] = y[c]; }
        }
        __syncthreads();
        for (int t = tid; t < 6 * cnt; t += nt) {
            const int s = t / 6, r = t % 6;
            const double *Lr = Y.L + 36 * (size_t)(base + 1 + s) + 6 * r;
            x[6 * (size_t)Y.rowidx[base + 1 + s] + r] -= Lr[0] * xj[0] + Lr[1] * xj[1] + Lr[2] * xj[2] + Lr[3] * xj[3] + Lr[4] * xj[4] + Lr[5] * xj[5];
        }
        __syncthreads();
    }
    // backward: L^T dx = y
    for (int j = Y.nb - 1; j >= 0; --j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        const double *D = Y.L + 36 * (size_t)base;
        // t_c = sum over sub blocks s, rows r of L_s[r][c] x[row_s][r] : thread (s, c), then a fixed-order sum over s
        for (int s0 = 0; s0 < cnt; s0 += 40) {
            const int ns = min(40, cnt - s0);
            for (int t = tid; t < 6 * ns; t += nt) {
                const int s = s0 + t / 6, c = t % 6;
                const double *Ls = Y.L + 36 * (size_t)(base + 1 + s);
                const double *xs = x + 6 * (size_t)Y.rowidx[base + 1 + s];
                part[t / 6][c] = Ls[c] * xs[0] + Ls[6 + c] * xs[1] + Ls[12 + c] * xs[2] + Ls[18 + c] * xs[3] + Ls[24 + c] * xs[4] + Ls[30 + c] * xs[5];
            }
            __syncthreads();
            if (tid < 6) {
                double a = 0.0;
                for (int s = 0; s < ns; ++s) a += part[s][tid];
                x[6 * (size_t)j + tid] -= a;
            }
            __syncthreads();
        }
        if (tid == 0) {
            double y[6];
            for (int c = 5; c >= 0; --c) {
                double a = x[6 * (size_t)j + c];
                for (int k = c + 1; k < 6; ++k) a -= D[6 * k + c] * y[k];
                y[c] = a / D[7 * c];
            }
            for (int c = 0; c < 6; ++c) x[6 * (size_t)j + c] = y[c];
        }
        __syncthreads();
    }
}
