// host_emul.cu — TEST HARNESS: runs the per-landmark / per-edge __host__ __device__ bodies of
// visual-inertial-odometry_b200/csrc/vio_kernels.cuh on the CPU, single threaded, over host arrays.
// It exists so the kernel arithmetic and the packer can be unit-tested against oracle/_ref in the
// GPU-less container (pytest -m "not gpu").  It is NOT part of the product library and nothing in the
// package loads it; the product path is CUDA only.
#include <cstring>
#include <string>
#include <vector>

#include "../include/vio_b200.h"
#include "../visual-inertial-odometry_b200/csrc/vio_pack.h"
#include "../visual-inertial-odometry_b200/csrc/vio_kernels.cuh"

namespace {
struct HostProblem {
    PackedGraph K;
    std::vector<double> pose, pose_bak, sb, poseRT, Hll, bl, wh, wo, sys, bS, dxp, dxl, invd_bak;
    DevView v;
};

int setup(const vio_graph *g, HostProblem &H, std::string &err) {
    int rc = pack_graph(g, 0, 1, H.K, err);
    if (rc) return rc;
    PackedGraph &K = H.K;
    H.pose.assign(g->pose, g->pose + 7 * (size_t)K.C);
    H.pose_bak = H.pose;
    H.sb.assign(9 * (size_t)K.NSB, 0.0);
    H.poseRT.assign(16 * (size_t)K.C, 0.0);
    H.Hll.assign(K.L, 0); H.bl.assign(K.L, 0); H.wh.assign(6 * (size_t)K.L, 0); H.wo.assign(6 * (size_t)K.E, 0);
    H.sys.assign(K.s_count + 3 * (size_t)K.P, 0.0);
    H.bS.assign(K.P, 0); H.dxp.assign(K.P, 0); H.dxl.assign(K.L, 0); H.invd_bak.assign(K.L, 0);
    DevView &v = H.v;
    memset(&v, 0, sizeof(v));
    v.C = K.C; v.NSB = K.NSB; v.L = K.L; v.P = K.P; v.NB = K.NB; v.E = K.E; v.storage = K.storage; v.nnzb = K.nnzb;
    v.batch = 1; v.Pper = K.P; v.Cper = K.C > 0 ? K.C : 1; v.NSBper = K.NSB > 0 ? K.NSB : 1;
    v.pose = H.pose.data(); v.pose_bak = H.pose_bak.data(); v.sb = H.sb.data(); v.invdep = K.invd.data();
    v.invdep_bak = H.invd_bak.data();
    v.pose_fixed = K.pose_fixed.data(); v.sb_fixed = K.sb_fixed.data();
    v.pose_off = K.pose_off.data(); v.sb_off = K.sb_off.data(); v.pose_blk = K.pose_blk.data();
    v.poseRT = H.poseRT.data();
    quat_to_R(K.qic, v.Ric);
    for (int k = 0; k < 3; ++k) v.tic[k] = K.tic[k];
    v.lm_host = K.lm_host.data(); v.lm_eptr = K.lm_eptr.data();
    v.lm_pix = K.pix.data(); v.lm_piy = K.piy.data(); v.lm_piz = K.piz.data();
    v.e_pose_j = K.e_pose_j.data(); v.e_pjx = K.pjx.data(); v.e_pjy = K.pjy.data();
    v.rp_info = g->rp_info; v.rp_loss = g->rp_loss; v.rp_delta = g->rp_loss_delta;
    v.Hll = H.Hll.data(); v.bl = H.bl.data(); v.wh = H.wh.data(); v.wo = H.wo.data();
    v.S = H.sys.data(); v.bcorr = v.S + K.s_count; v.bp = v.bcorr + K.P; v.hdiag = v.bp + K.P; v.bS = H.bS.data();
    v.bsr_rowptr = K.rowptr.data(); v.bsr_col = K.col.data(); v.bsr_tr = K.tr.data();
    v.dxp = H.dxp.data(); v.dxl = H.dxl.data();
    return VIO_OK;
}
}  // namespace

extern "C" {

// MakeHessian + Schur with the device bodies; S is returned dense P x P (undamped), mirrored.
int emul_linearize(const vio_graph *g, int with_schur, double *S, double *bp, double *bS, double *hdiag, double *Hll,
                   double *bl, double *wh, double *wo) {
    HostProblem H;
    std::string err;
    int rc = setup(g, H, err);
    if (rc) { fprintf(stderr, "emul: %s\n", err.c_str()); return rc; }
    DevView &v = H.v;
    const PackedGraph &K = H.K;
    for (int i = 0; i < K.C; ++i) pose_prep(v, i);
    for (int l = 0; l < K.L; ++l) {
        if (with_schur) linearize_landmark<true>(v, l);
        else linearize_landmark<false>(v, l);
    }
    Se3PriorView s;
    s.n = g->n_se3prior; s.pose = g->sp_pose; s.p = g->sp_p; s.q = g->sp_q; s.info = g->sp_info;
    for (int i = 0; i < s.n; ++i) se3prior_edge(v, s, i);
    const int P = K.P;
    if (S) {
        memset(S, 0, (size_t)P * P * sizeof(double));
        if (K.storage == VIO_STORAGE_DENSE) {
            memcpy(S, v.S, (size_t)P * P * sizeof(double));
        } else {
            for (int a = 0; a < K.NB; ++a)
                for (int k = K.rowptr[a]; k < K.rowptr[a + 1]; ++k)
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 6; ++c) S[(size_t)(6 * a + r) * P + 6 * K.col[k] + c] = v.S[36 * (size_t)k + 6 * r + c];
        }
        for (int r = 0; r < P; ++r)
            for (int c = 0; c < r; ++c) S[(size_t)r * P + c] = S[(size_t)c * P + r];
    }
    for (int i = 0; i < P; ++i) {
        if (bp) bp[i] = v.bp[i];
        if (bS) bS[i] = v.bp[i] - v.bcorr[i];
        if (hdiag) hdiag[i] = v.hdiag[i];
    }
    for (int l = 0; l < K.L; ++l) {
        if (Hll) Hll[K.lm_global[l]] = v.Hll[l];
        if (bl) bl[K.lm_global[l]] = v.bl[l];
    }
    if (wh) memcpy(wh, v.wh, 6 * (size_t)K.L * sizeof(double));
    if (wo) memcpy(wo, v.wo, 6 * (size_t)K.E * sizeof(double));
    return VIO_OK;
}

// full (P+M)^2 Hessian_ / b_ assembled like vio_get_hessian
int emul_hessian(const vio_graph *g, double *Hout, double *bout) {
    HostProblem H;
    std::string err;
    int rc = setup(g, H, err);
    if (rc) return rc;
    const PackedGraph &K = H.K;
    const int P = K.P, M = K.L, n = P + M;
    std::vector<double> S((size_t)P * P), bp(P), Hll(M), bl(M), wh(6 * (size_t)M), wo(6 * (size_t)K.E);
    rc = emul_linearize(g, 0, S.data(), bp.data(), nullptr, nullptr, Hll.data(), bl.data(), wh.data(), wo.data());
    if (rc) return rc;
    memset(Hout, 0, (size_t)n * n * sizeof(double));
    for (int r = 0; r < P; ++r) memcpy(Hout + (size_t)r * n, S.data() + (size_t)r * P, P * sizeof(double));
    for (int l = 0; l < M; ++l) {
        const int gl = P + K.lm_global[l];
        Hout[(size_t)gl * n + gl] = Hll[K.lm_global[l]];
        auto put = [&](int pose, const double *w) {
            const int off = K.pose_off[pose];
            for (int k = 0; k < 6; ++k) {
                Hout[(size_t)(off + k) * n + gl] += w[k];
                Hout[(size_t)gl * n + off + k] += w[k];
            }
        };
        if (K.lm_eptr[l] != K.lm_eptr[l + 1]) put(K.lm_host[l], &wh[6 * (size_t)l]);
        for (int e = K.lm_eptr[l]; e < K.lm_eptr[l + 1]; ++e) put(K.e_pose_j[e], &wo[6 * (size_t)e]);
    }
    for (int i = 0; i < P; ++i) bout[i] = bp[i];
    for (int l = 0; l < M; ++l) bout[P + K.lm_global[l]] = bl[K.lm_global[l]];
    return VIO_OK;
}

// VertexPose::Plus on every pose with the device body
int emul_update_pose(int n_pose, const double *pose_in, const double *dx6, double sign, double *pose_out) {
    std::vector<int> off(n_pose);
    for (int i = 0; i < n_pose; ++i) off[i] = 6 * i;
    std::vector<double> pose(pose_in, pose_in + 7 * (size_t)n_pose), bak(7 * (size_t)n_pose);
    DevView v;
    memset(&v, 0, sizeof(v));
    v.C = n_pose; v.pose = pose.data(); v.pose_bak = bak.data(); v.pose_off = off.data(); v.dxp = const_cast<double *>(dx6);
    for (int i = 0; i < n_pose; ++i) update_pose(v, i, sign, 1);
    memcpy(pose_out, pose.data(), pose.size() * sizeof(double));
    return VIO_OK;
}

// Σ rho(c r.r) over reprojection edges + SE3-prior chi2 (no flavour factor)
int emul_chi2(const vio_graph *g, double *out) {
    HostProblem H;
    std::string err;
    int rc = setup(g, H, err);
    if (rc) return rc;
    DevView &v = H.v;
    const PackedGraph &K = H.K;
    for (int i = 0; i < K.C; ++i) pose_prep(v, i);
    double chi = 0.0;
    for (int l = 0; l < K.L; ++l) {
        const int e0 = K.lm_eptr[l], e1 = K.lm_eptr[l + 1];
        if (e0 == e1) continue;
        const double lam = v.invdep[l];
        const double *RTh = v.poseRT + 16 * (size_t)v.lm_host[l];
        const double pci[3] = {v.lm_pix[l] / lam, v.lm_piy[l] / lam, v.lm_piz[l] / lam};
        double pbi[3], pw[3];
        mat3_mul_vec(v.Ric, pci, pbi);
        for (int k = 0; k < 3; ++k) pbi[k] += v.tic[k];
        mat3_mul_vec(RTh, pbi, pw);
        for (int k = 0; k < 3; ++k) pw[k] += RTh[9 + k];
        for (int e = e0; e < e1; ++e) {
            double pcj[3], pbj[3], r[2];
            reproj_residual(v.Ric, v.tic, v.poseRT + 16 * (size_t)v.e_pose_j[e], pw, v.e_pjx[e], v.e_pjy[e], pcj, pbj, r);
            double rho[3];
            loss_compute(v.rp_loss, v.rp_delta, v.rp_info * (r[0] * r[0] + r[1] * r[1]), rho);
            chi += rho[0];
        }
    }
    for (int i = 0; i < g->n_se3prior; ++i) {
        double r[6];
        se3prior_residual(v.pose + 7 * (size_t)g->sp_pose[i], g->sp_p + 3 * i, g->sp_q + 4 * i, r);
        const double *Om = g->sp_info + 36 * (size_t)i;
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) chi += r[a] * Om[6 * a + b] * r[b];
    }
    *out = chi;
    return VIO_OK;
}
}
