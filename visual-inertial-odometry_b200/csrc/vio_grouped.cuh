// vio_grouped.cuh — the production linearise + J^T W J + Schur kernel: one CTA per landmark GROUP.
//
// A group is a run of landmarks that share the host pose and whose observers fit NS_MAX pose "slots"
// (slot 0 = host).  Everything a group touches in the reduced camera system is a set of 6x6 blocks
// between its slots, so the CTA
//   1. stages the slots' R,t in shared memory and builds the per-landmark host chain (p_w, g, G),
//   2. evaluates edges slot-major (warp = slot, lanes = landmarks: coalesced ELL loads, uniform pose),
//      keeps the per-slot J^T W J / J^T W r sums in registers and reduces them ONCE per slot with a
//      recursive-halving shuffle reduce-scatter (no atomics),
//   3. reduces the per-landmark sums (H_ll, b_l, H_lp) and forms the Schur outer products
//      H_pl H_ll^-1 H_lp with every thread owning one 6x6 block pair in registers,
//   4. flushes the finished group tile with one RED.F64 per touched element of S.
// Reference dataflow replaced: Problem::MakeHessian + the Schur part of SolveLinearSystem
// (A15/backend/problem.cc:280-337,353-399; A17/src/backend/problem.cc:303-389,406-437).
//
// Edge algebra.  With B = reduce * Ric^T Rj^T (2x3) every Jacobian of EdgeReprojection factors through B:
//   J_lambda = B g,  J_i = B [I  G],  J_j = B [-I  N]   (g = Ri Ric pts_i (-1/l^2), G = -Ri hat(p_bi), N = Rj hat(p_bj))
// so with M = B^T W B (3x3) all blocks are products of M, G, N, g (A15/backend/edge_reprojection.cc:47-91).
#pragma once
#include "vio_dev.h"
#include "vio_kernels.cuh"

#define VIO_NS_MAX 22
#define VIO_GROUP_THREADS_MAX 320

struct GroupHdr {
    int host, ns, lm0, nlm, ell0, pair0, slot0, pad;
};

struct GroupView {
    int n_groups, ld;
    const GroupHdr *hdr;
    const int *slot_pose;        // [slot0 + s]
    const long long *pairinfo;   // [pair0 + p] = (offset << 2) | flags   (0 normal, 1 transposed, 2 diagonal, 3 skip)
    const double *ell_pjx, *ell_pjy;  // [ell0 + (s-1)*nlm + l], NaN = no observation
    const int *ell_edge;         // packed edge index (for H_lp observer rows), -1 = none
    unsigned long long *prof;    // optional [8] per-phase cycle counters (VIO_B200_PROFILE=1), else nullptr
};
#define VIO_PROF_MARK(slot)                                                          \
    do {                                                                             \
        if (gv.prof && threadIdx.x == 0) {                                           \
            const long long now_ = clock64();                                        \
            atomicAdd(gv.prof + (slot), (unsigned long long)(now_ - prof_t_));       \
            prof_t_ = now_;                                                          \
        }                                                                            \
    } while (0)

__host__ __device__ inline size_t group_smem_doubles(int ns, int nlm) {
    const int npairs = ns * (ns + 1) / 2;
    // pc[ns][12] lmh[nlm][17] w[nlm][ns][6] lmM[nlm][LMS] hinv[nlm] blv[nlm] red[ns][48] T[npairs][36] bvec[ns][18]
    // (lmh and lmM use ODD per-landmark strides so that lane = landmark accesses are bank-conflict free)
    return (size_t)ns * 12 + (size_t)nlm * 17 + 1 + (size_t)nlm * ns * 6 + (size_t)nlm * (((ns - 1) * 9) | 1) + 2 * (size_t)nlm + 1 +
           (size_t)ns * 48 + (size_t)npairs * 36 + (size_t)ns * 18 + (size_t)ns * 9 + (size_t)npairs;
}
__host__ __device__ inline size_t group_smem_bytes(int ns, int nlm) {
    const int npairs = ns * (ns + 1) / 2;
    // + ints: pose_id[ns] fixed[ns] off[ns] pair_a[npairs] pair_b[npairs]
    return group_smem_doubles(ns, nlm) * sizeof(double) + (3 * (size_t)ns + 2 * (size_t)npairs) * sizeof(int);
}

// rs_step / reduce_scatter48 / reduce_scatter32 (warp reduce-scatter by recursive halving) live in vio_kernels.cuh

__device__ __forceinline__ int sym6_index(int r, int c) {  // upper triangle, row-major
    if (r > c) { const int t = r; r = c; c = t; }
    return r * 6 - (r * (r - 1)) / 2 + (c - r);
}
__device__ __forceinline__ int sym3_index(int r, int c) {  // 00 01 02 11 12 22
    if (r > c) { const int t = r; r = c; c = t; }
    return r * 3 - (r * (r - 1)) / 2 + (c - r);
}

template <bool WITH_SCHUR>
__global__ void __launch_bounds__(VIO_GROUP_THREADS_MAX) k_linearize_grouped(DevView v, GroupView gv) {
    extern __shared__ double sm[];
    const GroupHdr h = gv.hdr[blockIdx.x];
    const int ns = h.ns, nlm = h.nlm, npairs = ns * (ns + 1) / 2;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    double *pc = sm;                                  // [ns][12]
    constexpr int LHS = 17;                           // odd strides: lane = landmark accesses hit 32 distinct banks
    const int LMS = ((ns - 1) * 9) | 1;
    double *lmh = pc + (size_t)ns * 12;               // [nlm][17]  pw(3) g(3) G(9)
    double *w = lmh + (((size_t)nlm * LHS + 1) & ~(size_t)1);  // [nlm][ns][6]  (16-byte aligned for double2 loads)
    double *lmM = w + (size_t)nlm * ns * 6;           // [nlm][LMS]  per (landmark, slot): M(6) m(3)
    double *hinv = lmM + (size_t)nlm * LMS;           // [nlm]
    double *blv = hinv + nlm;                         // [nlm]
    double *red = blv + nlm + ((nlm * (LMS + 2)) & 1);  // [ns][48] (kept 16-byte aligned)
    double *T = red + (size_t)ns * 48;                // [npairs][36]
    double *bvec = T + (size_t)npairs * 36;           // [ns][18]  bp(6) bcorr(6) hdiag(6)
    double *rjric = bvec + (size_t)ns * 18;           // [ns][9]   Rj * Ric per slot (shared by all edges of the slot)
    long long *pinfo_s = (long long *)(rjric + (size_t)ns * 9);  // [npairs] staged pair table
    int *pose_id = (int *)(pinfo_s + npairs);         // [ns]
    int *pfix = pose_id + ns;
    int *poff = pfix + ns;
    int *pair_a = poff + ns;                          // [npairs]
    int *pair_b = pair_a + npairs;

    long long prof_t_ = gv.prof ? clock64() : 0;
    unsigned long long prof_ns0_ = 0;
    if (gv.prof && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_ns0_));
    // ---- phase 0: slot table, pose cache, zero fills ---------------------------------------------------------
    for (int s = tid; s < ns; s += nt) {
        const int pid = s == 0 ? h.host : gv.slot_pose[h.slot0 + s];
        pose_id[s] = pid;
        pfix[s] = v.pose_fixed[pid];
        poff[s] = v.pose_off[pid];
    }
    for (int p = tid; p < npairs; p += nt) {
        // invert p = a*ns - a(a-1)/2 + (b-a)
        int a = 0, rem = p;
        while (rem >= ns - a) { rem -= ns - a; ++a; }
        pair_a[p] = a;
        pair_b[p] = a + rem;
    }
    if (!(h.pad & 1)) {  // bit 0: every landmark of the group is observed from every slot -> all entries get written
        for (size_t i = tid; i < (size_t)nlm * ns * 6; i += nt) w[i] = 0.0;
        for (size_t i = tid; i < (size_t)nlm * LMS; i += nt) lmM[i] = 0.0;
    }
    for (int i = tid; i < ns * 48; i += nt) red[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < ns * 12; i += nt) {
        const int s = i / 12, k = i % 12;
        pc[i] = v.poseRT[16 * (size_t)pose_id[s] + k];
    }
    __syncthreads();
    for (int s = tid; s < ns; s += nt) mat3_mul(pc + 12 * (size_t)s, v.Ric, rjric + 9 * (size_t)s);
    __syncthreads();
    const bool hfix = pfix[0] != 0;
    VIO_PROF_MARK(0);

    // ---- phase 0.5: per-landmark host chain ------------------------------------------------------------------
    for (int l = tid; l < nlm; l += nt) {
        const int gl = h.lm0 + l;
        const double lam = v.invdep[gl];
        const double pts_i[3] = {v.lm_pix[gl], v.lm_piy[gl], v.lm_piz[gl]};
        const double pci[3] = {pts_i[0] / lam, pts_i[1] / lam, pts_i[2] / lam};
        double pbi[3], pw[3], tmp[3], g[3], G[9];
        mat3_mul_vec(v.Ric, pci, pbi);
        pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
        mat3_mul_vec(pc, pbi, pw);
        pw[0] += pc[9]; pw[1] += pc[10]; pw[2] += pc[11];
        mat3_mul_vec(v.Ric, pts_i, tmp);
        mat3_mul_vec(pc, tmp, g);
        const double il2 = (v.lm_fixed && v.lm_fixed[gl]) ? 0.0 : -1.0 / (lam * lam);  // fixed landmark: J_lambda = 0
        mat3_mul_hat(pc, pbi, G);
        double *o = lmh + LHS * (size_t)l;
        o[0] = pw[0]; o[1] = pw[1]; o[2] = pw[2];
        o[3] = g[0] * il2; o[4] = g[1] * il2; o[5] = g[2] * il2;
#pragma unroll
        for (int k = 0; k < 9; ++k) o[6 + k] = -G[k];
    }
    __syncthreads();
    VIO_PROF_MARK(1);

    // ---- phase 1: edges, slot-major.  warp <-> slot, lanes <-> landmarks ----------------------------------------
    for (int s = 1 + wid; s < ns; s += nw) {
        const double *RTj = pc + 12 * (size_t)s;
        const bool jfix = pfix[s] != 0;
        double acc[48];
#pragma unroll
        for (int k = 0; k < 48; ++k) acc[k] = 0.0;
        // software pipeline: the next round's observation is in flight while this one is evaluated
        const size_t ebase = (size_t)h.ell0 + (size_t)(s - 1) * nlm;
        double n_pjx = 0.0, n_pjy = 0.0;
        int n_edge = -1;
        if (lane < nlm) { n_pjx = gv.ell_pjx[ebase + lane]; n_pjy = gv.ell_pjy[ebase + lane]; n_edge = gv.ell_edge[ebase + lane]; }
        for (int l = lane; l < nlm; l += 32) {
            const double pjx = n_pjx, pjy = n_pjy;
            const int edge = n_edge;
            if (l + 32 < nlm) { n_pjx = gv.ell_pjx[ebase + l + 32]; n_pjy = gv.ell_pjy[ebase + l + 32]; n_edge = gv.ell_edge[ebase + l + 32]; }
            if (pjx != pjx) continue;  // NaN: this landmark is not observed from slot s
            const double *lh = lmh + LHS * (size_t)l;
            const double pw[3] = {lh[0], lh[1], lh[2]};
            double pcj[3], pbj[3], r[2];
            reproj_residual(v.Ric, v.tic, RTj, pw, pjx, pjy, pcj, pbj, r);
            const double iz = 1.0 / pcj[2];
            const double rx = -pcj[0] * iz * iz, ry = -pcj[1] * iz * iz;
            // A = Ric^T Rj^T = (Rj Ric)^T ;  B = reduce * A
            const double *RjRic = rjric + 9 * (size_t)s;
            double B[6];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                B[c] = iz * RjRic[3 * c + 0] + rx * RjRic[3 * c + 2];
                B[3 + c] = iz * RjRic[3 * c + 1] + ry * RjRic[3 * c + 2];
            }
            double rho0, drho, W[3];
            robust_weights(v.rp_loss, v.rp_delta, v.rp_info, r, rho0, drho, W);
            double WB[6];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                WB[c] = W[0] * B[c] + W[1] * B[3 + c];
                WB[3 + c] = W[1] * B[c] + W[2] * B[3 + c];
            }
            double M[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) M[3 * a + b] = B[a] * WB[b] + B[3 + a] * WB[3 + b];
            const double dc = drho * v.rp_info;
            const double m[3] = {dc * (B[0] * r[0] + B[3] * r[1]), dc * (B[1] * r[0] + B[4] * r[1]),
                                 dc * (B[2] * r[0] + B[5] * r[1])};
            double *lm9 = lmM + (size_t)l * LMS + (s - 1) * 9;
            lm9[0] = M[0]; lm9[1] = M[1]; lm9[2] = M[2]; lm9[3] = M[4]; lm9[4] = M[5]; lm9[5] = M[8];
            lm9[6] = m[0]; lm9[7] = m[1]; lm9[8] = m[2];
            double *wog = v.wo + 6 * (size_t)edge;
            if (jfix) {
                double *wz = w + ((size_t)l * ns + s) * 6;
#pragma unroll
                for (int k = 0; k < 6; ++k) { wog[k] = 0.0; wz[k] = 0.0; }
                continue;
            }
            double N[9], MN[9], NMN[9], Ntm[3], Mg[3], NtMg[3];
            mat3_mul_hat(RTj, pbj, N);
            mat3_mul(M, N, MN);
            mat3t_mul(N, MN, NMN);
            mat3t_mul_vec(N, m, Ntm);
            const double g[3] = {lh[3], lh[4], lh[5]};
            mat3_mul_vec(M, g, Mg);
            mat3t_mul_vec(N, Mg, NtMg);
            double *ws = w + ((size_t)l * ns + s) * 6;
            const double wj[6] = {-Mg[0], -Mg[1], -Mg[2], NtMg[0], NtMg[1], NtMg[2]};
#pragma unroll
            for (int k = 0; k < 6; ++k) { ws[k] = wj[k]; wog[k] = wj[k]; }
            acc[0] += M[0]; acc[1] += M[1]; acc[2] += M[2]; acc[3] += M[4]; acc[4] += M[5]; acc[5] += M[8];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[6 + k] += MN[k];
            acc[33] += NMN[0]; acc[34] += NMN[1]; acc[35] += NMN[2]; acc[36] += NMN[4]; acc[37] += NMN[5]; acc[38] += NMN[8];
            acc[39] += m[0]; acc[40] += m[1]; acc[41] += m[2];
            acc[42] += Ntm[0]; acc[43] += Ntm[1]; acc[44] += Ntm[2];
            if (!hfix) {
                const double *G = lh + 6;
                double GtM[9], GtMN[9];
                mat3t_mul(G, M, GtM);
                mat3t_mul(G, MN, GtMN);
#pragma unroll
                for (int k = 0; k < 9; ++k) { acc[15 + k] += GtM[k]; acc[24 + k] += GtMN[k]; }
            }
        }
        const int base = reduce_scatter48(acc, lane);
        red[48 * (size_t)s + base] = acc[0];
        if (!(lane & 1)) red[48 * (size_t)s + base + 1] = acc[1];
    }
    __syncthreads();
    VIO_PROF_MARK(2);

    // ---- phase 1.5: per-landmark sums -> H_ll, b_l, host row of H_lp, host blocks ------------------------------
    {
        double hb[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) hb[k] = 0.0;
        for (int l = tid; l < nlm; l += nt) {
            double Ms[6] = {0, 0, 0, 0, 0, 0}, ms[3] = {0, 0, 0};
            const double *q = lmM + (size_t)l * LMS;
            for (int s = 0; s < ns - 1; ++s) {
#pragma unroll
                for (int k = 0; k < 6; ++k) Ms[k] += q[9 * s + k];
                ms[0] += q[9 * s + 6]; ms[1] += q[9 * s + 7]; ms[2] += q[9 * s + 8];
            }
            const double *lh = lmh + LHS * (size_t)l;
            const double g[3] = {lh[3], lh[4], lh[5]};
            const double *G = lh + 6;
            const double Msf[9] = {Ms[0], Ms[1], Ms[2], Ms[1], Ms[3], Ms[4], Ms[2], Ms[4], Ms[5]};
            double Mg[3];
            mat3_mul_vec(Msf, g, Mg);
            const double Hll = g[0] * Mg[0] + g[1] * Mg[1] + g[2] * Mg[2];
            const double bl = -(g[0] * ms[0] + g[1] * ms[1] + g[2] * ms[2]);
            const int gl = h.lm0 + l;
            v.Hll[gl] = Hll;
            v.bl[gl] = bl;
            hinv[l] = (v.lm_fixed && v.lm_fixed[gl]) ? 0.0 : 1.0 / Hll;
            blv[l] = bl;
            double wh[6] = {0, 0, 0, 0, 0, 0};
            if (!hfix) {
                double GtMg[3], MG[9], GtMG[9], Gtm[3];
                mat3t_mul_vec(G, Mg, GtMg);
                wh[0] = Mg[0]; wh[1] = Mg[1]; wh[2] = Mg[2]; wh[3] = GtMg[0]; wh[4] = GtMg[1]; wh[5] = GtMg[2];
                mat3_mul(Msf, G, MG);
                mat3t_mul(G, MG, GtMG);
                mat3t_mul_vec(G, ms, Gtm);
                // host block [[Ms, Ms G],[G^T Ms, G^T Ms G]] upper triangle, row-major: rows 0..2 then 3..5
                hb[0] += Msf[0]; hb[1] += Msf[1]; hb[2] += Msf[2]; hb[3] += MG[0]; hb[4] += MG[1]; hb[5] += MG[2];
                hb[6] += Msf[4]; hb[7] += Msf[5]; hb[8] += MG[3]; hb[9] += MG[4]; hb[10] += MG[5];
                hb[11] += Msf[8]; hb[12] += MG[6]; hb[13] += MG[7]; hb[14] += MG[8];
                hb[15] += GtMG[0]; hb[16] += GtMG[1]; hb[17] += GtMG[2];
                hb[18] += GtMG[4]; hb[19] += GtMG[5];
                hb[20] += GtMG[8];
                hb[21] -= ms[0]; hb[22] -= ms[1]; hb[23] -= ms[2]; hb[24] -= Gtm[0]; hb[25] -= Gtm[1]; hb[26] -= Gtm[2];
            }
            double *whg = v.wh + 6 * (size_t)gl;
            double *ws = w + (size_t)l * ns * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) { whg[k] = wh[k]; ws[k] = wh[k]; }
        }
        // block reduction of the 27 host values: warp reduce-scatter, then one shared-memory atomic per warp and value
        if (wid * 32 < nlm) {
            const int base = reduce_scatter32(hb, lane);
            if (base < 27) atomicAdd(&red[base], hb[0]);
        }
    }
    __syncthreads();
    VIO_PROF_MARK(3);

    // ---- phase 1.9: the direct (J^T W J) part of the group tile is read straight from the reduced vectors by the flush
    // (direct_value below); T only receives the Schur part.  Stage the pair table while we are at it.
    for (int p = tid; p < npairs; p += nt) pinfo_s[p] = gv.pairinfo[h.pair0 + p];
    if (!WITH_SCHUR)
        for (int idx = tid; idx < npairs * 36; idx += nt) T[idx] = 0.0;
    for (int idx = tid; idx < ns * 6; idx += nt) {
        const int s = idx / 6, k = idx % 6;
        double bp, hd;
        if (s == 0) {
            bp = red[21 + k];
            hd = red[sym6_index(k, k)];
        } else {
            const double *R = red + 48 * (size_t)s;
            bp = k < 3 ? R[39 + k] : -R[42 + (k - 3)];
            hd = k < 3 ? R[sym3_index(k, k)] : R[33 + sym3_index(k - 3, k - 3)];
        }
        bvec[18 * s + k] = bp;
        bvec[18 * s + 12 + k] = hd;
        bvec[18 * s + 6 + k] = 0.0;
    }
    __syncthreads();
    VIO_PROF_MARK(4);

    // ---- phase 2: Schur outer products, thread <-> (block pair, landmark subset) --------------------------------
    int nsub = 1;
    if (WITH_SCHUR) {
        // thread <-> (pair p, landmark subset q).  Lanes of a warp hold DIFFERENT pairs of the SAME subset, so the
        // w_a / w_b reads of a warp hit few distinct shared-memory words (broadcast).  Subset q > 0 parks its
        // partial tile in the (now dead) lmM region; the flush adds the copies up - no barrier rounds, no atomics.
        nsub = max(1, min(nt / npairs, 1 + (int)(((size_t)nlm * LMS) / ((size_t)npairs * 36))));
        for (int p0 = 0; p0 < npairs; p0 += nt) {  // one pass unless npairs > blockDim
            const int p = p0 + (nsub > 1 ? tid % npairs : tid);
            const int q = nsub > 1 ? tid / npairs : 0;
            const bool active = p < npairs && q < nsub;
            if (active) {
                double acc[36];
#pragma unroll
                for (int k = 0; k < 36; ++k) acc[k] = 0.0;
                const int a = pair_a[p], b = pair_b[p];
                const int lstride = ns * 6;
                const double *wa = w + (size_t)q * lstride + a * 6;
                const double *wb = w + (size_t)q * lstride + b * 6;
#pragma unroll 2
                for (int l = q; l < nlm; l += nsub, wa += (size_t)nsub * lstride, wb += (size_t)nsub * lstride) {
                    const double inv = hinv[l];
                    // rows of w are 48 B: three 16-byte shared loads each
                    const double2 a0 = *reinterpret_cast<const double2 *>(wa), a1 = *reinterpret_cast<const double2 *>(wa + 2),
                                  a2 = *reinterpret_cast<const double2 *>(wa + 4);
                    const double2 b0 = *reinterpret_cast<const double2 *>(wb), b1 = *reinterpret_cast<const double2 *>(wb + 2),
                                  b2 = *reinterpret_cast<const double2 *>(wb + 4);
                    const double x[6] = {a0.x * inv, a0.y * inv, a1.x * inv, a1.y * inv, a2.x * inv, a2.y * inv};
                    const double y[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 6; ++c) acc[6 * r + c] += x[r] * y[c];
                }
                if (q == 0) {
                    double *t = T + 36 * (size_t)p;
#pragma unroll
                    for (int k = 0; k < 36; ++k) t[k] = -acc[k];
                } else {
                    double *t = lmM + ((size_t)(q - 1) * npairs + p) * 36;
#pragma unroll
                    for (int k = 0; k < 36; ++k) t[k] = acc[k];
                }
            }
        }
        for (int idx = tid; idx < ns * 6; idx += nt) {
            const int s = idx / 6, k = idx % 6;
            double t = 0.0;
            for (int l = 0; l < nlm; ++l) t += w[((size_t)l * ns + s) * 6 + k] * hinv[l] * blv[l];
            bvec[18 * s + 6 + k] = t;
        }
        __syncthreads();
    }

    VIO_PROF_MARK(5);
    // ---- flush: one RED.F64 per touched element of the reduced system -------------------------------------------
    for (int idx = tid; idx < npairs * 36; idx += nt) {
        const int p = idx / 36, k = idx - 36 * p, r = k / 6, c = k - 6 * r;
        const long long info = pinfo_s[p];
        const int flags = (int)(info & 3);
        if (flags == 3) continue;
        if (flags == 2 && r > c) continue;
        const size_t off = (size_t)(info >> 2);
        const size_t e = flags == 1 ? (size_t)c * gv.ld + r : (size_t)r * gv.ld + c;
        double val = T[idx];
        {
            const int a = pair_a[p], b = pair_b[p];
            if (a == 0 && b == 0) {
                val += red[sym6_index(r, c)];
            } else if (a == 0) {
                const double *R = red + 48 * (size_t)b;  // (0,s) = [[-A1, A2],[-A3, A4]]
                if (r < 3 && c < 3) val -= R[sym3_index(r, c)];
                else if (r < 3) val += R[6 + 3 * r + (c - 3)];
                else if (c < 3) val -= R[15 + 3 * (r - 3) + c];
                else val += R[24 + 3 * (r - 3) + (c - 3)];
            } else if (a == b) {
                const double *R = red + 48 * (size_t)a;  // (s,s) = [[A1, -A2],[-A2^T, A5]]
                if (r < 3 && c < 3) val += R[sym3_index(r, c)];
                else if (r < 3) val -= R[6 + 3 * r + (c - 3)];
                else if (c < 3) val -= R[6 + 3 * c + (r - 3)];
                else val += R[33 + sym3_index(r - 3, c - 3)];
            }
        }
        if (WITH_SCHUR)
            for (int q = 1; q < nsub; ++q) val -= lmM[((size_t)(q - 1) * npairs + p) * 36 + k];
        if (val != 0.0) atomicAdd(v.S + off + e, val);
    }
    for (int idx = tid; idx < ns * 6; idx += nt) {
        const int s = idx / 6, k = idx % 6;
        if (pfix[s]) continue;
        atomicAdd(v.bp + poff[s] + k, bvec[18 * s + k]);
        atomicAdd(v.hdiag + poff[s] + k, bvec[18 * s + 12 + k]);
        if (WITH_SCHUR) atomicAdd(v.bcorr + poff[s] + k, bvec[18 * s + 6 + k]);
    }
    __syncthreads();
    VIO_PROF_MARK(6);
    if (gv.prof && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        atomicAdd(gv.prof + 7, ns1 - prof_ns0_);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_lin_edges — the edge half of the split pipeline (round 2): residuals, Jacobian algebra, J^T W J / J^T W r of every group,
// the per-landmark H_ll, b_l and the rows of H_lp; the Schur complement itself is k_schur_groups' (below).
// Same per-edge arithmetic and the same warp <-> observer slot, lane <-> landmark mapping as k_linearize_grouped, but sized
// for SEVERAL small CTAs per SM instead of one big one, so that the serial phases of one group (setup, host chain,
// landmark sums, flush) overlap the edge phase of its neighbours:
//   * 2..5 warps per CTA, at most 168 registers (launch bounds keep 12 warps resident per SM),
//   * no group tile T and no shared copy of H_lp (the rows go straight to HBM; back-substitution and the Schur kernel read them),
//   * the per-(landmark, slot) matrices M, m are not kept: every warp adds them into ITS OWN partial per-landmark sums
//     lmS[warp][landmark][9] (a warp walks its slots one after the other, lanes are distinct landmarks: no races, no
//     atomics, fixed order), the landmark phase adds the <= 5 partials.
// Shared memory: ~45 KB for a 100-landmark, 11-slot group with 4 warps (the fused kernel: 169 KB).
// Reference dataflow replaced: Problem::MakeHessian (A15/backend/problem.cc:280-337; A17/src/backend/problem.cc:303-389).
// ------------------------------------------------------------------------------------------------------------------
#define VIO_EDGE_WARPS_MAX 5
__host__ __device__ inline size_t edges_smem_bytes(int ns, int nlm, int nw) {
    // doubles: red[ns][48] lmh[nlm][15] lmS[nw][nlm][9] pc[ns][12] rjric[ns][9] pinfo[2 ns - 1]; ints: fixed[ns] off[ns]
    const size_t dbl = (size_t)ns * 48 + (size_t)nlm * 15 + (size_t)nw * nlm * 9 + (size_t)ns * 12 + (size_t)ns * 9 + 2 * (size_t)ns;
    return dbl * sizeof(double) + 2 * (size_t)ns * sizeof(int);
}
__device__ __forceinline__ int pair_index(int a, int b, int ns) { return a * ns - (a * (a - 1)) / 2 + (b - a); }
// sum of a 48-vector over the 16 lanes of a half-warp, scattered: afterwards the lane holds elements base .. base + 2 in v[0..2]
__device__ __forceinline__ int reduce_scatter48_half(double *v, int lane) {
    int base = 0;
    rs_step<48>(v, lane & 8, 8); base += (lane & 8) ? 24 : 0;
    rs_step<24>(v, lane & 4, 4); base += (lane & 4) ? 12 : 0;
    rs_step<12>(v, lane & 2, 2); base += (lane & 2) ? 6 : 0;
    rs_step<6>(v, lane & 1, 1);  base += (lane & 1) ? 3 : 0;
    return base;
}

// SPW = observer slots a warp evaluates per round: 1 (32 landmarks x 1 slot) or 2 (16 landmarks x 2 slots: half-warp = slot).
// The host picks whichever needs fewer rounds for the graph (100 landmarks x 10 observer slots: 40 rounds vs 35).
template <int MAXT, int MINB, int SPW>
__global__ void __launch_bounds__(MAXT, MINB) k_lin_edges(DevView v, GroupView gv) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LPW = 32 / SPW;
    const GroupHdr h = gv.hdr[blockIdx.x];
    const int ns = h.ns, nlm = h.nlm;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int sub = lane / LPW, ll = lane - sub * LPW;
    constexpr int LHS = 15;                           // odd stride: lane = landmark accesses hit distinct banks
    double *red = sm;                                 // [ns][48]
    double *lmh = red + (size_t)ns * 48;              // [nlm][15]  pw(3) g(3) G(9)
    double *lmS = lmh + (size_t)nlm * LHS;            // [nw][nlm][9]  per-warp partial sums of M(6) m(3)
    double *pc = lmS + (size_t)nw * nlm * 9;          // [ns][12]
    double *rjric = pc + (size_t)ns * 12;             // [ns][9]
    long long *pinfo_s = (long long *)(rjric + (size_t)ns * 9);  // [2 ns - 1]: (0,0), (0,s), (s,s)
    int *pfix = (int *)(pinfo_s + 2 * ns);            // [ns]
    int *poff = pfix + ns;

    long long prof_t_ = gv.prof ? clock64() : 0;
    unsigned long long prof_ns0_ = 0;
    if (gv.prof && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_ns0_));
    // ---- phase 0: slot table, pose cache ------------------------------------------------------------------------
    const int nsp = (ns - 1 + SPW - 1) / SPW;  // rounds of SPW observer slots; warp w takes w, w + nw, ...
    for (int s = tid; s < ns; s += nt) {
        const int pid = s == 0 ? h.host : gv.slot_pose[h.slot0 + s];
        pfix[s] = v.pose_fixed[pid];
        poff[s] = v.pose_off[pid];
    }
    for (int i = tid; i < ns * 12; i += nt) {
        const int s = i / 12, k = i - 12 * s;
        const int pid = s == 0 ? h.host : gv.slot_pose[h.slot0 + s];
        pc[i] = v.poseRT[16 * (size_t)pid + k];
    }
    for (int bi = tid; bi < 2 * ns - 1; bi += nt) {
        const int a = bi < ns ? 0 : bi - ns + 1, b = bi < ns ? bi : a;
        pinfo_s[bi] = gv.pairinfo[h.pair0 + pair_index(a, b, ns)];
    }
    for (int i = tid; i < 48; i += nt) red[i] = 0.0;
    // a warp without any slot leaves its partial sums untouched: zero them
    if (wid >= nsp)
        for (int i = lane; i < nlm * 9; i += 32) lmS[(size_t)wid * nlm * 9 + i] = 0.0;
    __syncthreads();
    for (int s = tid; s < ns; s += nt) mat3_mul(pc + 12 * (size_t)s, v.Ric, rjric + 9 * (size_t)s);
    const bool hfix = pfix[0] != 0;
    VIO_PROF_MARK(0);

    // ---- phase 0.5: per-landmark host chain (needs pc only) ----------------------------------------------------------
    for (int l = tid; l < nlm; l += nt) {
        const int gl = h.lm0 + l;
        const double il = vio_rcp(v.invdep[gl]);
        const double pts_i[3] = {v.lm_pix[gl], v.lm_piy[gl], v.lm_piz[gl]};
        const double pci[3] = {pts_i[0] * il, pts_i[1] * il, pts_i[2] * il};
        double pbi[3], pw[3], tmp[3], g[3], G[9];
        mat3_mul_vec(v.Ric, pci, pbi);
        pbi[0] += v.tic[0]; pbi[1] += v.tic[1]; pbi[2] += v.tic[2];
        mat3_mul_vec(pc, pbi, pw);
        pw[0] += pc[9]; pw[1] += pc[10]; pw[2] += pc[11];
        mat3_mul_vec(v.Ric, pts_i, tmp);
        mat3_mul_vec(pc, tmp, g);
        const double il2 = (v.lm_fixed && v.lm_fixed[gl]) ? 0.0 : -(il * il);  // fixed landmark: J_lambda = 0
        mat3_mul_hat(pc, pbi, G);
        double *o = lmh + LHS * (size_t)l;
        o[0] = pw[0]; o[1] = pw[1]; o[2] = pw[2];
        o[3] = g[0] * il2; o[4] = g[1] * il2; o[5] = g[2] * il2;
#pragma unroll
        for (int k = 0; k < 9; ++k) o[6 + k] = -G[k];
    }
    __syncthreads();
    VIO_PROF_MARK(1);

    // ---- phase 1: edges.  warp <-> SPW observer slots at a time, lanes <-> landmarks (x slot) -------------------------
    double *myS = lmS + (size_t)wid * nlm * 9;
    bool first = true;  // the warp's first round initialises its partial sums
    for (int sp = wid; sp < nsp; sp += nw, first = false) {
        const int s = 1 + sp * SPW + sub;
        const bool sv = SPW == 1 || s < ns;  // SPW = 2 and an odd number of observer slots: the upper half idles in the last round
        const int sc = sv ? s : ns - 1;
        const double *RTj = pc + 12 * (size_t)sc;
        const double *RjRic = rjric + 9 * (size_t)sc;
        const bool jfix = pfix[sc] != 0;
        double acc[48];
#pragma unroll
        for (int k = 0; k < 48; ++k) acc[k] = 0.0;
        const size_t ebase = (size_t)h.ell0 + (size_t)(sc - 1) * nlm;
        double n_pjx = 0.0, n_pjy = 0.0;
        int n_edge = -1;
        if (ll < nlm) { n_pjx = gv.ell_pjx[ebase + ll]; n_pjy = gv.ell_pjy[ebase + ll]; n_edge = gv.ell_edge[ebase + ll]; }
        // SPW = 2: uniform trip count (the exchange between the half-warps below needs every lane)
        for (int l0 = 0; SPW == 2 ? l0 < nlm : l0 + ll < nlm; l0 += LPW) {
            const int l = l0 + ll;
            const double pjx = n_pjx, pjy = n_pjy;
            const int edge = n_edge;
            if (l + LPW < nlm) { n_pjx = gv.ell_pjx[ebase + l + LPW]; n_pjy = gv.ell_pjy[ebase + l + LPW]; n_edge = gv.ell_edge[ebase + l + LPW]; }
            const bool inr = SPW == 1 || l < nlm;
            const bool valid = inr && sv && pjx == pjx;  // NaN: this landmark is not observed from slot s
            const double *lh = lmh + LHS * (size_t)(inr ? l : 0);
            double pcj[3], pbj[3], r[2];
            double M[9], m[3];
            if (valid) {
                const double pw[3] = {lh[0], lh[1], lh[2]};
                const double d[3] = {pw[0] - RTj[9], pw[1] - RTj[10], pw[2] - RTj[11]};
                mat3t_mul_vec(RTj, d, pbj);
                const double e[3] = {pbj[0] - v.tic[0], pbj[1] - v.tic[1], pbj[2] - v.tic[2]};
                mat3t_mul_vec(v.Ric, e, pcj);
                const double iz = vio_rcp(pcj[2]);
                const double xz = pcj[0] * iz, yz = pcj[1] * iz;
                r[0] = xz - pjx;
                r[1] = yz - pjy;
                const double rx = -xz * iz, ry = -yz * iz;
                double B[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    B[c] = iz * RjRic[3 * c + 0] + rx * RjRic[3 * c + 2];
                    B[3 + c] = iz * RjRic[3 * c + 1] + ry * RjRic[3 * c + 2];
                }
                double rho0, drho, W[3];
                robust_weights(v.rp_loss, v.rp_delta, v.rp_info, r, rho0, drho, W);
                double WB[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    WB[c] = W[0] * B[c] + W[1] * B[3 + c];
                    WB[3 + c] = W[1] * B[c] + W[2] * B[3 + c];
                }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = a; b < 3; ++b) M[3 * a + b] = B[a] * WB[b] + B[3 + a] * WB[3 + b];
                M[3] = M[1]; M[6] = M[2]; M[7] = M[5];
                const double dc = drho * v.rp_info;
                const double r0 = dc * r[0], r1 = dc * r[1];
                m[0] = B[0] * r0 + B[3] * r1;
                m[1] = B[1] * r0 + B[4] * r1;
                m[2] = B[2] * r0 + B[5] * r1;
            } else if (SPW == 2 || first) {
#pragma unroll
                for (int k = 0; k < 9; ++k) M[k] = 0.0;
                m[0] = m[1] = m[2] = 0.0;
            }
            // per-landmark sums of M (6) and m (3) over the slots: into the warp's own partial sums
            if (SPW == 2) {
                double q9[9] = {M[0], M[1], M[2], M[4], M[5], M[8], m[0], m[1], m[2]};
#pragma unroll
                for (int k = 0; k < 9; ++k) q9[k] += __shfl_xor_sync(0xffffffffu, q9[k], 16);
                if (inr) {  // both halves hold the sum: the lower half stores 0..4, the upper half 5..8
                    double *ls = myS + 9 * (size_t)l;
                    if (sub == 0) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) ls[k] = first ? q9[k] : ls[k] + q9[k];
                    } else {
#pragma unroll
                        for (int k = 5; k < 9; ++k) ls[k] = first ? q9[k] : ls[k] + q9[k];
                    }
                }
            } else if (valid || first) {
                double *ls = myS + 9 * (size_t)l;
                const double q9[9] = {M[0], M[1], M[2], M[4], M[5], M[8], m[0], m[1], m[2]};
#pragma unroll
                for (int k = 0; k < 9; ++k) ls[k] = first ? q9[k] : ls[k] + q9[k];
            }
            if (!valid) continue;
            double2 *wog = reinterpret_cast<double2 *>(v.wo + 6 * (size_t)edge);
            if (jfix) {
                wog[0] = make_double2(0.0, 0.0); wog[1] = make_double2(0.0, 0.0); wog[2] = make_double2(0.0, 0.0);
                continue;
            }
            double N[9], MN[9], Mg[3], NtMg[3];
            mat3_mul_hat(RTj, pbj, N);
            mat3_mul(M, N, MN);
            const double g[3] = {lh[3], lh[4], lh[5]};
            mat3_mul_vec(M, g, Mg);
            mat3t_mul_vec(N, Mg, NtMg);
            wog[0] = make_double2(-Mg[0], -Mg[1]); wog[1] = make_double2(-Mg[2], NtMg[0]); wog[2] = make_double2(NtMg[1], NtMg[2]);
            acc[0] += M[0]; acc[1] += M[1]; acc[2] += M[2]; acc[3] += M[4]; acc[4] += M[5]; acc[5] += M[8];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[6 + k] += MN[k];
            // products that are only ever summed are accumulated by their FMA chains directly (no separate DMUL / DADD)
#define VIO_ACC3T(dst, A_, i_, B_, j_) dst = fma(A_[i_], B_[j_], fma(A_[3 + i_], B_[3 + j_], fma(A_[6 + i_], B_[6 + j_], dst)))
            VIO_ACC3T(acc[33], N, 0, MN, 0); VIO_ACC3T(acc[34], N, 0, MN, 1); VIO_ACC3T(acc[35], N, 0, MN, 2);  // N^T M N, upper triangle
            VIO_ACC3T(acc[36], N, 1, MN, 1); VIO_ACC3T(acc[37], N, 1, MN, 2); VIO_ACC3T(acc[38], N, 2, MN, 2);
            acc[39] += m[0]; acc[40] += m[1]; acc[41] += m[2];
            acc[42] = fma(N[0], m[0], fma(N[3], m[1], fma(N[6], m[2], acc[42])));  // N^T m
            acc[43] = fma(N[1], m[0], fma(N[4], m[1], fma(N[7], m[2], acc[43])));
            acc[44] = fma(N[2], m[0], fma(N[5], m[1], fma(N[8], m[2], acc[44])));
            if (!hfix) {
                const double *G = lh + 6;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        VIO_ACC3T(acc[15 + 3 * a + b], G, a, M, b);   // G^T M
                        VIO_ACC3T(acc[24 + 3 * a + b], G, a, MN, b);  // G^T M N
                    }
            }
#undef VIO_ACC3T
        }
        if (SPW == 1) {
            const int base = reduce_scatter48(acc, lane);
            red[48 * (size_t)s + base] = acc[0];
            if (!(lane & 1)) red[48 * (size_t)s + base + 1] = acc[1];
        } else {
            const int base = reduce_scatter48_half(acc, lane);
            if (sv) {
                double *R = red + 48 * (size_t)s + base;
                R[0] = acc[0]; R[1] = acc[1]; R[2] = acc[2];
            }
        }
    }
    __syncthreads();
    VIO_PROF_MARK(2);

    // ---- phase 1.5: per-landmark sums -> H_ll, b_l, host row of H_lp, host blocks ------------------------------
    {
        const int ncopy = nw;  // every warp's copy is initialised (its first round, or the zero fill of phase 0)
        double hb[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) hb[k] = 0.0;
        for (int l = tid; l < nlm; l += nt) {
            double Ms[6] = {0, 0, 0, 0, 0, 0}, ms[3] = {0, 0, 0};
            for (int c = 0; c < ncopy; ++c) {
                const double *q = lmS + ((size_t)c * nlm + l) * 9;
#pragma unroll
                for (int k = 0; k < 6; ++k) Ms[k] += q[k];
                ms[0] += q[6]; ms[1] += q[7]; ms[2] += q[8];
            }
            const double *lh = lmh + LHS * (size_t)l;
            const double g[3] = {lh[3], lh[4], lh[5]};
            const double *G = lh + 6;
            const double Msf[9] = {Ms[0], Ms[1], Ms[2], Ms[1], Ms[3], Ms[4], Ms[2], Ms[4], Ms[5]};
            double Mg[3];
            mat3_mul_vec(Msf, g, Mg);
            const double Hll = g[0] * Mg[0] + g[1] * Mg[1] + g[2] * Mg[2];
            const double bl = -(g[0] * ms[0] + g[1] * ms[1] + g[2] * ms[2]);
            const int gl = h.lm0 + l;
            v.Hll[gl] = Hll;
            v.bl[gl] = bl;
            double wh[6] = {0, 0, 0, 0, 0, 0};
            if (!hfix) {
                double GtMg[3], MG[9], GtMG[9], Gtm[3];
                mat3t_mul_vec(G, Mg, GtMg);
                wh[0] = Mg[0]; wh[1] = Mg[1]; wh[2] = Mg[2]; wh[3] = GtMg[0]; wh[4] = GtMg[1]; wh[5] = GtMg[2];
                mat3_mul(Msf, G, MG);
                mat3t_mul(G, MG, GtMG);
                mat3t_mul_vec(G, ms, Gtm);
                hb[0] += Msf[0]; hb[1] += Msf[1]; hb[2] += Msf[2]; hb[3] += MG[0]; hb[4] += MG[1]; hb[5] += MG[2];
                hb[6] += Msf[4]; hb[7] += Msf[5]; hb[8] += MG[3]; hb[9] += MG[4]; hb[10] += MG[5];
                hb[11] += Msf[8]; hb[12] += MG[6]; hb[13] += MG[7]; hb[14] += MG[8];
                hb[15] += GtMG[0]; hb[16] += GtMG[1]; hb[17] += GtMG[2];
                hb[18] += GtMG[4]; hb[19] += GtMG[5];
                hb[20] += GtMG[8];
                hb[21] -= ms[0]; hb[22] -= ms[1]; hb[23] -= ms[2]; hb[24] -= Gtm[0]; hb[25] -= Gtm[1]; hb[26] -= Gtm[2];
            }
            double2 *whg = reinterpret_cast<double2 *>(v.wh + 6 * (size_t)gl);
            whg[0] = make_double2(wh[0], wh[1]); whg[1] = make_double2(wh[2], wh[3]); whg[2] = make_double2(wh[4], wh[5]);
        }
        // block reduction of the 27 host values: warp reduce-scatter, then the warps add their parts one after the other
        // (fixed order: the result does not depend on scheduling)
        const int base = reduce_scatter32(hb, lane);
        for (int w2 = 0; w2 < nw; ++w2) {
            if (w2 == wid && wid * 32 < nlm && base < 27) red[base] += hb[0];
            __syncthreads();
        }
    }
    VIO_PROF_MARK(3);

    // ---- b_p and diag(H_pp) of every slot ----------------------------------------------------------------------
    for (int idx = tid; idx < ns * 6; idx += nt) {
        const int s = idx / 6, k = idx - 6 * s;
        double bp, hd;
        if (s == 0) {
            bp = red[21 + k];
            hd = red[sym6_index(k, k)];
        } else {
            const double *R = red + 48 * (size_t)s;
            bp = k < 3 ? R[39 + k] : -R[42 + (k - 3)];
            hd = k < 3 ? R[sym3_index(k, k)] : R[33 + sym3_index(k - 3, k - 3)];
        }
        if (!pfix[s]) {
            atomicAdd(v.bp + poff[s] + k, bp);
            atomicAdd(v.hdiag + poff[s] + k, hd);
        }
    }
    VIO_PROF_MARK(4);
    VIO_PROF_MARK(5);
    // ---- flush of the direct (J^T W J) blocks: (0,0), (0,s), (s,s) -- one RED.F64 per element ----------------------------
    const int nblk = 2 * ns - 1;  // block 0 = (0,0); 1..ns-1 = (0,s); ns..2ns-2 = (s,s)
    for (int idx = tid; idx < nblk * 36; idx += nt) {
        const int bi = idx / 36, k = idx - 36 * bi, r = k / 6, c = k - 6 * r;
        const long long info = pinfo_s[bi];
        const int flags = (int)(info & 3);
        if (flags == 3) continue;
        if (flags == 2 && r > c) continue;
        const size_t off = (size_t)(info >> 2);
        const size_t e = flags == 1 ? (size_t)c * gv.ld + r : (size_t)r * gv.ld + c;
        double val;
        if (bi == 0) {
            val = red[sym6_index(r, c)];
        } else if (bi < ns) {
            const double *R = red + 48 * (size_t)bi;  // (0,s) = [[-A1, A2],[-A3, A4]]
            if (r < 3 && c < 3) val = -R[sym3_index(r, c)];
            else if (r < 3) val = R[6 + 3 * r + (c - 3)];
            else if (c < 3) val = -R[15 + 3 * (r - 3) + c];
            else val = R[24 + 3 * (r - 3) + (c - 3)];
        } else {
            const double *R = red + 48 * (size_t)(bi - ns + 1);  // (s,s) = [[A1, -A2],[-A2^T, A5]]
            if (r < 3 && c < 3) val = R[sym3_index(r, c)];
            else if (r < 3) val = -R[6 + 3 * r + (c - 3)];
            else if (c < 3) val = -R[6 + 3 * c + (r - 3)];
            else val = R[33 + sym3_index(r - 3, c - 3)];
        }
        if (val != 0.0) atomicAdd(v.S + off + e, val);
    }
    VIO_PROF_MARK(6);
    if (gv.prof && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        atomicAdd(gv.prof + 7, ns1 - prof_ns0_);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_schur_groups — the Schur complement of a landmark group as ONE symmetric rank-nlm update on the FP64 tensor cores
// (round 2; the second half of the split pipeline).  Per group, with the landmarks' rows of H_lp
//     V = [ wo (nlm x 6 (ns - 1), observer slots 1 .. ns - 1) | wh (nlm x 6, host slot) | b_l ]      and  h = 1 / H_ll:
//     -V^T diag(h) V  -> the 6x6 blocks of the reduced camera system   (A17/src/backend/problem.cc:412-431)
//     and, in its last column, b_corr = W^T diag(h) b_l                 (:424-427)
// * Operands: the observer rows of a group are ONE contiguous slab of `wo` when every landmark stores its edges in slot
//   order (header bit 1, set by the packer - true for tracked features), the host rows one slab of `wh`: a single thread
//   issues two TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) and the CTA waits on the barrier once, while the
//   other threads compute h and the lookup tables.  The DMMA fragments are read from that very layout (row stride
//   6 (ns - 1) doubles; 60 at 11 slots is conflict-free), no transposition, no scaling pass: the A fragments are multiplied
//   by h on the fly (3 DMUL per 9 DMMA).  Other groups (ragged tracks, strides that would bank-conflict) are gathered by
//   hand into a padded copy of the same layout, six 16-byte chunks per thread in flight.
// * Work unit = a 3 x 3 block of 8x8 tiles (tile sets si <= sj): 6 fragment loads feed 9 DMMAs per k step.
//   In V^T h V the A fragment of tile row t and the B fragment of tile column t are the same shared-memory words
//   (lane (g, q) holds V[k0 + q][8 t + g]).
// * Flush: every lane looks its rows / columns up once per unit in a per-column table (slot, row-in-block) and a full
//   ns x ns table of block offsets (both orientations), then issues one RED.F64 per element straight from the accumulators.
// ~55 KB of shared memory: four CTAs per SM, so load, DMMA and flush phases of different groups overlap.
// ------------------------------------------------------------------------------------------------------------------
#define VIO_SCHUR_THREADS 192
__host__ __device__ inline bool schur_stride_ok(int ld) { return ld % 16 == 4 || ld % 16 == 12; }
__host__ __device__ inline int schur_ldo(int ns, bool direct) {
    int ld = 6 * (ns - 1);
    if (direct && schur_stride_ok(ld)) return ld;  // the TMA path keeps the global layout
    ld = (ld + 1) & ~1;
    while (!schur_stride_ok(ld)) ld += 2;  // the 4 rows of a fragment land on distinct bank groups
    return ld;
}
__host__ __device__ inline size_t schur_smem_bytes(int ns, int nlm) {
    const int kp = (nlm + 3) & ~3, tr = (6 * ns + 1 + 7) >> 3;
    const int ldo = schur_ldo(ns, false) > 6 * (ns - 1) ? schur_ldo(ns, false) : 6 * (ns - 1);
    // doubles: wo_s[kp][ldo] wh_s[kp][6] hq[kp] bq[kp] zero[2]; i64: tab[ns][ns] mbar; ints: colinfo[8 tr] poff[ns] pfix[ns]
    return ((size_t)kp * ldo + (size_t)kp * 6 + 2 * (size_t)kp + 2) * sizeof(double) + ((size_t)ns * ns + 1) * sizeof(long long) +
           ((size_t)8 * tr + 2 * (size_t)ns) * sizeof(int);
}
__device__ __forceinline__ void schur_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned grp_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void grp_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(grp_saddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void grp_mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(grp_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void grp_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(grp_saddr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(grp_saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void grp_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GRP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GRP_DONE;\n"
        "bra GRP_WAIT;\n"
        "GRP_DONE:\n"
        "}\n" ::"r"(grp_saddr(bar)),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(VIO_SCHUR_THREADS, 4) k_schur_groups(DevView v, GroupView gv) {
    extern __shared__ __align__(16) double ssm[];
    const GroupHdr h = gv.hdr[blockIdx.x];
    const int ns = h.ns, nlm = h.nlm, npairs = ns * (ns + 1) / 2;
    const int kp = (nlm + 3) & ~3, NO = 6 * (ns - 1), D = 6 * ns;  // columns: [0, NO) observers, [NO, D) host, D = b_l
    const int TR = (D + 1 + 7) >> 3;
    const bool direct = (h.pad & 2) != 0 && schur_stride_ok(NO);
    const int LDO = direct ? NO : schur_ldo(ns, false);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5, g = lane >> 2, q = lane & 3;
    double *wo_s = ssm;                                 // [kp][LDO]
    double *wh_s = wo_s + (size_t)kp * LDO;             // [kp][6]
    double *hq = wh_s + (size_t)kp * 6;                 // [kp]  1 / H_ll (0 for padding rows and empty landmarks)
    double *bq = hq + kp;                               // [kp]  b_l
    double *zero = bq + kp;                             // [2]
    long long *tab = (long long *)(zero + 2);           // [ns][ns]  (offset << 2) | flags of block (a, b), both orientations
    unsigned long long *mbar = (unsigned long long *)(tab + (size_t)ns * ns);
    int *colinfo = (int *)(mbar + 1);                   // [8 TR]  slot << 8 | row-in-block; -1 = b column, -2 = padding
    int *poff = colinfo + 8 * TR;                       // [ns]
    int *pfix = poff + ns;
    if (tid == 0) grp_mbar_init(mbar, 1);
    __syncthreads();
    if (direct) {
        if (tid == 0) {
            const int e0 = gv.ell_edge[h.ell0];
            const unsigned bo = (unsigned)nlm * NO * sizeof(double), bh = (unsigned)nlm * 6 * sizeof(double);
            grp_mbar_expect(mbar, bo + bh);
            grp_bulk_load(wo_s, v.wo + 6 * (size_t)e0, bo, mbar);
            grp_bulk_load(wh_s, v.wh + 6 * (size_t)h.lm0, bh, mbar);
        }
        // rows nlm .. kp - 1 are outside the copies: zero them
        for (int i = tid; i < (kp - nlm) * NO; i += nt) wo_s[(size_t)nlm * NO + i] = 0.0;
        for (int i = tid; i < (kp - nlm) * 6; i += nt) wh_s[(size_t)nlm * 6 + i] = 0.0;
    } else {
        // gather: thread <-> one 16-byte chunk (row l, slot s, part 0..2); nt is a multiple of 3, so (l, s) advance by a fixed
        // (dq, dr) per round; six chunks per thread are in flight at a time (index loads, then row loads, then stores)
        constexpr int U = 6;
        const int per = nt / 3, dq = per / ns, dr = per - dq * ns;
        const int part = tid % 3, rs0 = tid / 3;
        int l = rs0 / ns, s = rs0 - l * ns;
        while (l < kp) {
            int lu[U], su[U], eu[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                lu[u] = l; su[u] = s;
                l += dq; s += dr;
                if (s >= ns) { s -= ns; ++l; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                eu[u] = -1;
                if (lu[u] < nlm && su[u] > 0) eu[u] = gv.ell_edge[(size_t)h.ell0 + (size_t)(su[u] - 1) * nlm + lu[u]];
            }
            double2 wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                wv[u] = make_double2(0.0, 0.0);
                if (lu[u] < nlm) {
                    if (su[u] == 0) wv[u] = reinterpret_cast<const double2 *>(v.wh + 6 * (size_t)(h.lm0 + lu[u]))[part];
                    else if (eu[u] >= 0) wv[u] = reinterpret_cast<const double2 *>(v.wo + 6 * (size_t)eu[u])[part];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (lu[u] >= kp) continue;
                double *dst = su[u] == 0 ? wh_s + (size_t)lu[u] * 6 : wo_s + (size_t)lu[u] * LDO + 6 * (su[u] - 1);
                reinterpret_cast<double2 *>(dst)[part] = wv[u];
            }
        }
    }
    // ---- tables (overlap the copies)
    for (int s = tid; s < ns; s += nt) {
        const int pid = s == 0 ? h.host : gv.slot_pose[h.slot0 + s];
        pfix[s] = v.pose_fixed[pid];
        poff[s] = v.pose_off[pid];
    }
    for (int p = tid; p < npairs; p += nt) {
        int a = 0, rem = p;
        while (rem >= ns - a) { rem -= ns - a; ++a; }
        const int b = a + rem;
        const long long info = gv.pairinfo[h.pair0 + p];
        tab[a * ns + b] = info;
        if (a != b) tab[b * ns + a] = (info & 3) < 2 ? (info ^ 1) : info;  // the same block seen from the other side: transposed
    }
    for (int c = tid; c < 8 * TR; c += nt) {
        int ci;
        if (c < NO) { const int sl = c / 6; ci = ((sl + 1) << 8) | (c - 6 * sl); }
        else if (c < D) ci = c - NO;  // host slot 0
        else ci = c == D ? -1 : -2;
        colinfo[c] = ci;
    }
    for (int l = tid; l < kp; l += nt) {
        double x = 0.0, b = 0.0;
        if (l < nlm) {
            const double hll = v.Hll[h.lm0 + l];
            if (hll != 0.0) x = 1.0 / hll;
            b = v.bl[h.lm0 + l];
        }
        hq[l] = x; bq[l] = b;
    }
    if (tid == 0) { zero[0] = 0.0; zero[1] = 0.0; }
    __syncthreads();
    if (direct) grp_mbar_wait(mbar, 0);
    // ---- units: tile sets si <= sj of three tile rows / columns each
    const int nsets = (TR + 2) / 3, nunits = nsets * (nsets + 1) / 2;
    for (int u = warp; u < nunits; u += nw) {
        int si = 0, rem = u;
        while (rem >= nsets - si) { rem -= nsets - si; ++si; }
        const int sj = si + rem;
        // per-lane fragment pointers (row q, column 8 t + g of V) and row strides of the unit's six tiles
        const double *fp[6];
        int fs[6];
#pragma unroll
        for (int x = 0; x < 6; ++x) {
            const int t = x < 3 ? 3 * si + x : 3 * sj + (x - 3), c = 8 * t + g;
            if (t >= TR || c > D) { fp[x] = zero; fs[x] = 0; }
            else if (c < NO) { fp[x] = wo_s + (size_t)q * LDO + c; fs[x] = 4 * LDO; }
            else if (c < D) { fp[x] = wh_s + (size_t)q * 6 + (c - NO); fs[x] = 24; }
            else { fp[x] = bq + q; fs[x] = 4; }
        }
        double c[3][3][2];
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y) c[x][y][0] = c[x][y][1] = 0.0;
        const double *hp = hq + q;
        if (si == sj) {
#pragma unroll 2
            for (int k0 = 0; k0 < kp; k0 += 4) {
                const double hh = hp[k0];
                const double f0 = *fp[0], f1 = *fp[1], f2 = *fp[2];
                fp[0] += fs[0]; fp[1] += fs[1]; fp[2] += fs[2];
                const double a0 = f0 * hh, a1 = f1 * hh, a2 = f2 * hh;
                schur_dmma(c[0][0][0], c[0][0][1], a0, f0);
                schur_dmma(c[0][1][0], c[0][1][1], a0, f1);
                schur_dmma(c[0][2][0], c[0][2][1], a0, f2);
                schur_dmma(c[1][1][0], c[1][1][1], a1, f1);
                schur_dmma(c[1][2][0], c[1][2][1], a1, f2);
                schur_dmma(c[2][2][0], c[2][2][1], a2, f2);
            }
        } else {
#pragma unroll 2
            for (int k0 = 0; k0 < kp; k0 += 4) {
                const double hh = hp[k0];
                const double a0 = *fp[0] * hh, a1 = *fp[1] * hh, a2 = *fp[2] * hh;
                const double b0 = *fp[3], b1 = *fp[4], b2 = *fp[5];
#pragma unroll
                for (int x = 0; x < 6; ++x) fp[x] += fs[x];
                schur_dmma(c[0][0][0], c[0][0][1], a0, b0);
                schur_dmma(c[0][1][0], c[0][1][1], a0, b1);
                schur_dmma(c[0][2][0], c[0][2][1], a0, b2);
                schur_dmma(c[1][0][0], c[1][0][1], a1, b0);
                schur_dmma(c[1][1][0], c[1][1][1], a1, b1);
                schur_dmma(c[1][2][0], c[1][2][1], a1, b2);
                schur_dmma(c[2][0][0], c[2][0][1], a2, b0);
                schur_dmma(c[2][1][0], c[2][1][1], a2, b1);
                schur_dmma(c[2][2][0], c[2][2][1], a2, b2);
            }
        }
        // ---- flush: lane (g, q) holds rows 8 ti + g and columns 8 tj + 2 q + {0, 1} of every tile
        int ri[3], cj[3][2];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const int ti = 3 * si + x, tj = 3 * sj + x;
            ri[x] = ti < TR ? colinfo[8 * ti + g] : -2;
            cj[x][0] = tj < TR ? colinfo[8 * tj + 2 * q] : -2;
            cj[x][1] = tj < TR ? colinfo[8 * tj + 2 * q + 1] : -2;
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            if (ri[x] < 0) continue;  // padding or the b_l column as a row
            const int sa = ri[x] >> 8, r = ri[x] & 255;
            const bool afix = pfix[sa] != 0;
            const int aoff = poff[sa] + r;
#pragma unroll
            for (int y = 0; y < 3; ++y) {
                if (si == sj && x > y) continue;
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                    const int cc2 = cj[y][e2];
                    const double val = c[x][y][e2];
                    if (cc2 == -2 || val == 0.0) continue;
                    if (cc2 == -1) {  // b_corr of row i
                        if (!afix) atomicAdd(v.bcorr + aoff, val);
                        continue;
                    }
                    // diagonal tiles hold (i, j) and (j, i): keep i <= j in column order
                    if (si == sj && x == y && 2 * q + e2 < g) continue;
                    const int sb = cc2 >> 8, cc = cc2 & 255;
                    const long long info = tab[sa * ns + sb];
                    const int flags = (int)(info & 3);
                    if (flags == 3) continue;
                    const size_t off = (size_t)(info >> 2);
                    const size_t el = flags == 1 ? (size_t)cc * gv.ld + r : (size_t)r * gv.ld + cc;
                    atomicAdd(v.S + off + el, -val);
                }
            }
        }
    }
}
