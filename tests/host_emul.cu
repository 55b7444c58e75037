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
#include "../visual-inertial-odometry_b200/csrc/vio_bchol.h"
#include "../visual-inertial-odometry_b200/csrc/vio_bcr.h"

namespace {
struct HostProblem {
    PackedGraph K;
    std::vector<double> pose, pose_bak, sb, poseRT, Hll, bl, wh, wo, we, sys, bS, dxp, dxl, invd_bak;
    DevView v;
};

void setup_packed(const vio_graph *g, const double *pose, HostProblem &H);
int setup(const vio_graph *g, HostProblem &H, std::string &err) {
    int rc = pack_graph(g, 0, 1, H.K, err);
    if (rc) return rc;
    setup_packed(g, g->pose, H);
    return VIO_OK;
}
// host view over an already packed graph (H.K); g supplies the factor constants, pose the C x 7 vertex values
void setup_packed(const vio_graph *g, const double *pose, HostProblem &H) {
    PackedGraph &K = H.K;
    H.pose.assign(pose, pose + 7 * (size_t)K.C);
    H.pose_bak = H.pose;
    H.sb.assign(9 * (size_t)K.NSB, 0.0);
    H.poseRT.assign(16 * (size_t)K.C, 0.0);
    H.Hll.assign(K.L, 0); H.bl.assign(K.L, 0); H.wh.assign(6 * (size_t)K.L, 0); H.wo.assign(6 * (size_t)K.E, 0);
    H.sys.assign(K.s_count + 3 * (size_t)K.P, 0.0);
    H.bS.assign(K.P, 0); H.dxp.assign(K.P, 0); H.dxl.assign(K.L, 0); H.invd_bak.assign(K.L, 0);
    DevView &v = H.v;
    memset(&v, 0, sizeof(v));
    v.C = K.C; v.NSB = K.NSB; v.L = K.L; v.P = K.P; v.NB = K.NB; v.E = K.E; v.storage = K.storage; v.nnzb = K.nnzb;
    v.batch = K.batch; v.Pper = K.Pper > 0 ? K.Pper : K.P;
    v.Cper = K.C > 0 ? K.C / K.batch : 1; v.NSBper = K.NSB > 0 ? K.NSB / K.batch : 1;
    v.pose = H.pose.data(); v.pose_bak = H.pose_bak.data(); v.sb = H.sb.data(); v.invdep = K.invd.data();
    v.invdep_bak = H.invd_bak.data();
    v.pose_fixed = K.pose_fixed.data(); v.sb_fixed = K.sb_fixed.data();
    v.pose_off = K.pose_off.data(); v.sb_off = K.sb_off.data(); v.pose_blk = K.pose_blk.data();
    v.poseRT = H.poseRT.data();
    quat_to_R(K.qic, v.Ric);
    for (int k = 0; k < 3; ++k) v.tic[k] = K.tic[k];
    v.lm_host = K.lm_host.data(); v.lm_eptr = K.lm_eptr.data();
    v.lm_fixed = K.lm_fixed.empty() ? nullptr : K.lm_fixed.data();
    v.lm_pix = K.pix.data(); v.lm_piy = K.piy.data(); v.lm_piz = K.piz.data();
    v.e_pose_j = K.e_pose_j.data(); v.e_pjx = K.pjx.data(); v.e_pjy = K.pjy.data();
    v.rp_info = g->rp_info; v.rp_loss = g->rp_loss; v.rp_delta = g->rp_loss_delta;
    v.Hll = H.Hll.data(); v.bl = H.bl.data(); v.wh = H.wh.data(); v.wo = H.wo.data();
    H.we.assign(6 * (size_t)(K.L > 0 ? K.L : 1), 0.0);
    v.we = H.we.data(); v.ext_pose = K.ext_free ? g->ext_pose : -1;
    v.S = H.sys.data(); v.bcorr = v.S + K.s_count; v.bp = v.bcorr + K.P; v.hdiag = v.bp + K.P; v.bS = H.bS.data();
    v.bsr_rowptr = K.rowptr.data(); v.bsr_col = K.col.data(); v.bsr_tr = K.tr.data();
    v.dxp = H.dxp.data(); v.dxl = H.dxl.data();
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

// Lock-step batches (vio_solve_batched_lockstep): pack every item on its own and merge (PackedMerge, the production
// path), pack the caller-concatenated graph with batch = B (pack_graph's own batch mode), and
//  (1) compare every table of the two packs (n_diff = number of differing arrays; landmark-order dependent arrays are
//      only compared when no item has an edge-less landmark, where the two orders coincide),
//  (2) run MakeHessian + Schur with the device bodies over the MERGED pack: S_out = B stacked Pper x Pper systems
//      (upper triangles mirrored), bS_out = B x Pper.
int emul_merge_check(const vio_graph *const *items, int B, const vio_graph *concat, int *n_diff, double *S_out, double *bS_out) {
    std::string err;
    std::vector<PackedGraph> Ks(B);
    for (int k = 0; k < B; ++k) {
        vio_graph gk = *items[k];
        gk.storage = VIO_STORAGE_DENSE;
        gk.n_se3prior = 0;
        int rc = pack_graph(&gk, 0, 1, Ks[k], err);
        if (rc) { fprintf(stderr, "emul: item %d: %s\n", k, err.c_str()); return rc; }
    }
    HostProblem H;
    PackedMerge mg;
    int rc = mg.prepare(Ks, H.K, err);
    if (rc) { fprintf(stderr, "emul: merge: %s\n", err.c_str()); return rc; }
    mg.fill(0, B);
    PackedGraph K2;
    vio_graph gc = *concat;
    gc.storage = VIO_STORAGE_DENSE;
    rc = pack_graph(&gc, 0, 1, K2, err, B);
    if (rc) { fprintf(stderr, "emul: concat: %s\n", err.c_str()); return rc; }
    const PackedGraph &M = H.K;
    int nd = 0;
    bool edgeless = false;
    for (int l = 0; l < M.L; ++l) edgeless = edgeless || M.lm_eptr[l] == M.lm_eptr[l + 1];
#define CMP(field) do { if (!(M.field == K2.field)) { ++nd; fprintf(stderr, "emul: merged pack differs in %s\n", #field); } } while (0)
    CMP(C); CMP(NSB); CMP(NB); CMP(P); CMP(L); CMP(E); CMP(storage); CMP(batch); CMP(Pper); CMP(s_count);
    CMP(pose_off); CMP(sb_off); CMP(pose_blk); CMP(blk_off); CMP(blk_dim); CMP(blk_fixed); CMP(pose_fixed); CMP(sb_fixed); CMP(row_fixed);
    if (!edgeless) {
        CMP(lm_global); CMP(lm_host); CMP(lm_eptr); CMP(e_pose_j); CMP(pix); CMP(piy); CMP(piz); CMP(pjx); CMP(pjy); CMP(invd);
        CMP(grouped_ok); CMP(n_groups); CMP(group_threads); CMP(group_smem_max);
        CMP(g_hdr); CMP(g_slot_pose); CMP(g_pairinfo); CMP(ell_edge); CMP(ell_pjy);
        // ell_pjx marks missing entries with NaN: compare bit patterns
        if (M.ell_pjx.size() != K2.ell_pjx.size() ||
            memcmp(M.ell_pjx.data(), K2.ell_pjx.data(), M.ell_pjx.size() * sizeof(double)) != 0) { ++nd; fprintf(stderr, "emul: merged pack differs in ell_pjx\n"); }
    }
#undef CMP
    if (n_diff) *n_diff = nd;
    setup_packed(concat, concat->pose, H);
    DevView &v = H.v;
    for (int i = 0; i < M.C; ++i) pose_prep(v, i);
    for (int l = 0; l < M.L; ++l) linearize_landmark<true>(v, l);
    const int P = M.P, Pper = M.Pper;
    if (S_out) {
        memcpy(S_out, v.S, (size_t)P * Pper * sizeof(double));
        for (int k = 0; k < B; ++k) {
            double *Sk = S_out + (size_t)k * Pper * Pper;
            for (int r = 0; r < Pper; ++r)
                for (int c = 0; c < r; ++c) Sk[(size_t)r * Pper + c] = Sk[(size_t)c * Pper + r];
        }
    }
    if (bS_out)
        for (int i = 0; i < P; ++i) bS_out[i] = v.bp[i] - v.bcorr[i];
    return VIO_OK;
}

// Block-sparse Cholesky: the host symbolic factorisation of the product (vio_bchol.h) driven by a plain CPU restatement of
// the device numeric loops (k_bchol_init / k_bchol_factor / k_bchol_solve walk the same colptr / rowidx / update map).
// val: BSR values (nnzb x 36) of a symmetric positive definite block matrix; returns x = (A + lambda I)^-1 b, nnz(L).
int emul_bchol_solve(int nb, const int *rowptr_, const int *col_, const double *val, double lambda, const double *b, double *x,
                     long long *nnzL) {
    std::vector<int> rowptr(rowptr_, rowptr_ + nb + 1), col(col_, col_ + rowptr_[nb]);
    BcholSymbolic Y;
    if (!bchol_symbolic(nb, rowptr, col, 8LL * 1000 * 1000, Y)) return VIO_ERR_UNSUPPORTED;
    if (nnzL) *nnzL = Y.nnzL;
    std::vector<double> L(36 * (size_t)Y.nnzL, 0.0);
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (Y.a_to_l[k] < 0) continue;
            for (int e = 0; e < 36; ++e) L[36 * (size_t)Y.a_to_l[k] + e] = val[36 * (size_t)k + e] + ((col[k] == i && e % 7 == 0) ? lambda : 0.0);
        }
    for (int j = 0; j < nb; ++j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        double *D = &L[36 * (size_t)base];
        for (int c = 0; c < 6; ++c) {
            double d = D[7 * c];
            for (int k = 0; k < c; ++k) d -= D[6 * c + k] * D[6 * c + k];
            if (!(d > 0.0)) return VIO_ERR_INVALID;
            d = sqrt(d);
            D[7 * c] = d;
            for (int r = c + 1; r < 6; ++r) {
                double t = D[6 * r + c];
                for (int k = 0; k < c; ++k) t -= D[6 * r + k] * D[6 * c + k];
                D[6 * r + c] = t / d;
            }
            for (int r = 0; r < c; ++r) D[6 * r + c] = 0.0;
        }
        for (int t = 0; t < 6 * cnt; ++t) {
            double *row = &L[36 * (size_t)(base + 1 + t / 6) + 6 * (t % 6)], xr[6];
            for (int c = 0; c < 6; ++c) {
                double a = row[c];
                for (int k = 0; k < c; ++k) a -= xr[k] * D[6 * c + k];
                xr[c] = a / D[7 * c];
            }
            for (int c = 0; c < 6; ++c) row[c] = xr[c];
        }
        for (long long q = Y.upd_ptr[j]; q < Y.upd_ptr[j + 1]; ++q)
            for (int e = 0; e < 36; ++e) {
                const double *La = &L[36 * (size_t)(base + Y.upd_a[q]) + 6 * (e / 6)], *Lb = &L[36 * (size_t)(base + Y.upd_b[q]) + 6 * (e % 6)];
                double sacc = 0.0;
                for (int k = 0; k < 6; ++k) sacc += La[k] * Lb[k];
                L[36 * (size_t)Y.upd_dst[q] + e] -= sacc;
            }
    }
    for (int t = 0; t < 6 * nb; ++t) x[t] = b[t];
    for (int j = 0; j < nb; ++j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        const double *D = &L[36 * (size_t)base];
        double y[6];
        for (int c = 0; c < 6; ++c) {
            double a = x[6 * (size_t)j + c];
            for (int k = 0; k < c; ++k) a -= D[6 * c + k] * y[k];
            y[c] = a / D[7 * c];
        }
        for (int c = 0; c < 6; ++c) x[6 * (size_t)j + c] = y[c];
        for (int sb = 0; sb < cnt; ++sb)
            for (int r = 0; r < 6; ++r) {
                const double *Lr = &L[36 * (size_t)(base + 1 + sb) + 6 * r];
                double a = 0.0;
                for (int c = 0; c < 6; ++c) a += Lr[c] * y[c];
                x[6 * (size_t)Y.rowidx[base + 1 + sb] + r] -= a;
            }
    }
    for (int j = nb - 1; j >= 0; --j) {
        const int base = Y.colptr[j], cnt = Y.colptr[j + 1] - base - 1;
        const double *D = &L[36 * (size_t)base];
        for (int sb = 0; sb < cnt; ++sb) {
            const double *Ls = &L[36 * (size_t)(base + 1 + sb)], *xs = &x[6 * (size_t)Y.rowidx[base + 1 + sb]];
            for (int c = 0; c < 6; ++c) {
                double a = 0.0;
                for (int r = 0; r < 6; ++r) a += Ls[6 * r + c] * xs[r];
                x[6 * (size_t)j + c] -= a;
            }
        }
        double y[6];
        for (int c = 5; c >= 0; --c) {
            double a = x[6 * (size_t)j + c];
            for (int k = c + 1; k < 6; ++k) a -= D[6 * k + c] * y[k];
            y[c] = a / D[7 * c];
        }
        for (int c = 0; c < 6; ++c) x[6 * (size_t)j + c] = y[c];
    }
    return VIO_OK;
}

// Block cyclic reduction (vio_bcr.h): the host plan of the product interpreted sequentially on the CPU with plain dense
// loops - the same items, slots, couplings and update rules the persistent device kernel (vio_bcr.cuh) executes.
namespace {
// items [first, last) of a schedule over tiles of M rows x LD; xpool = export pool (open chains), done = per-item flags
int bcr_run_items(const std::vector<BcrItem> &items, int first, int last, std::vector<char> &done, int M, int LD, std::vector<double> &pool,
                  std::vector<double> &bv, std::vector<double> &xv, std::vector<double> *xpool) {
    const size_t MM = (size_t)M * LD;
    std::vector<double> Dm(MM), X(MM), Z(MM), U(MM), t(M);
    auto tn = [&](const double *A, const double *B, double *C, double sign, bool accumulate) {  // C (+)= sign * A^T B
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) {
                double acc = 0.0;
                for (int r = 0; r < M; ++r) acc += A[(size_t)r * LD + i] * B[(size_t)r * LD + j];
                C[(size_t)i * LD + j] = (accumulate ? C[(size_t)i * LD + j] : 0.0) + sign * acc;
            }
    };
    for (int q = first; q < last; ++q) {
        const BcrItem &it = items[q];
        for (int d : it.dep)
            if (d >= 0 && !done[d]) return VIO_ERR_STATE;  // the order must satisfy every dependency
        if (it.kind & BCR_EXPORT) {
            if (!xpool) return VIO_ERR_STATE;
            tn(&pool[(size_t)it.cl_a * MM], &pool[(size_t)it.cl_b * MM], &(*xpool)[(size_t)it.cl_slot * MM], -1.0, false);
            done[q] = 1;
            continue;
        }
        double *bk = &bv[(size_t)it.node * M];
        if (it.kind & BCR_BACKSUB) {
            const double *Uk = &pool[(size_t)it.node * MM];
            for (int i = 0; i < M; ++i) {
                double a = bk[i];
                if (it.left >= 0) for (int c = 0; c < M; ++c) a -= pool[(size_t)it.cl_slot * MM + (size_t)i * LD + c] * xv[(size_t)it.left * M + c];
                if (it.right >= 0) for (int c = 0; c < M; ++c) a -= pool[(size_t)it.cr_slot * MM + (size_t)i * LD + c] * xv[(size_t)it.right * M + c];
                t[i] = a;
            }
            for (int i = 0; i < M; ++i) {
                double a = 0.0;
                for (int r = i; r < M; ++r) a += Uk[(size_t)i * LD + r] * t[r];
                xv[(size_t)it.node * M + i] = a;
            }
            done[q] = 1;
            continue;
        }
        std::copy(pool.begin() + (size_t)it.node * MM, pool.begin() + (size_t)(it.node + 1) * MM, Dm.begin());
        for (int u = 0; u < 2; ++u) {
            if (it.upd_slot[u] < 0) continue;
            const double *W = &pool[(size_t)it.upd_slot[u] * MM], *ye = &bv[(size_t)it.upd_node[u] * M];
            tn(W, W, Dm.data(), -1.0, true);
            for (int i = 0; i < M; ++i) {
                double a = 0.0;
                for (int r = 0; r < M; ++r) a += W[(size_t)r * LD + i] * ye[r];
                bk[i] -= a;
            }
        }
        if (!(it.kind & BCR_ELIM)) {
            std::copy(Dm.begin(), Dm.end(), pool.begin() + (size_t)it.node * MM);
            done[q] = 1;
            continue;
        }
        auto coupling = [&](int mode, int a, int bb, double *out) {
            if (mode == 0) { std::fill(out, out + MM, 0.0); return; }
            if (mode == 1) {
                const double *src = &pool[(size_t)a * MM];
                for (int i = 0; i < M; ++i)
                    for (int j = 0; j < M; ++j) out[(size_t)i * LD + j] = bb ? src[(size_t)j * LD + i] : src[(size_t)i * LD + j];
                return;
            }
            tn(&pool[(size_t)a * MM], &pool[(size_t)bb * MM], out, -1.0, false);
        };
        coupling(it.cl_mode, it.cl_a, it.cl_b, X.data());
        coupling(it.cr_mode, it.cr_a, it.cr_b, Z.data());
        if (it.kind & BCR_MERGE)
            for (size_t e = 0; e < MM; ++e) Z[e] += X[e];
        // D = L L^T ; U = L^-T by forward elimination on [D | I]
        std::fill(U.begin(), U.end(), 0.0);
        for (int i = 0; i < M; ++i) U[(size_t)i * LD + i] = 1.0;
        for (int j = 0; j < M; ++j) {
            const double d = Dm[(size_t)j * LD + j];
            if (!(d > 0.0)) return VIO_ERR_INVALID;
            const double pinv = 1.0 / sqrt(d);
            std::vector<double> v(M, 0.0);
            for (int k = j; k < M; ++k) v[k] = Dm[(size_t)j * LD + k] * pinv;
            for (int c = 0; c <= j; ++c) U[(size_t)c * LD + j] *= pinv;
            for (int i = j + 1; i < M; ++i) {
                for (int k = j + 1; k < M; ++k) Dm[(size_t)i * LD + k] -= v[i] * v[k];
                for (int c = 0; c <= j; ++c) U[(size_t)c * LD + i] -= v[i] * U[(size_t)c * LD + j];
            }
        }
        if (it.cl_slot >= 0) tn(U.data(), X.data(), &pool[(size_t)it.cl_slot * MM], 1.0, false);
        if (it.cr_slot >= 0) tn(U.data(), Z.data(), &pool[(size_t)it.cr_slot * MM], 1.0, false);
        for (int i = 0; i < M; ++i) {
            double a = 0.0;
            for (int r = 0; r <= i; ++r) a += U[(size_t)r * LD + i] * bk[r];
            t[i] = a;
        }
        for (int i = 0; i < M; ++i) bk[i] = t[i];
        std::copy(U.begin(), U.end(), pool.begin() + (size_t)it.node * MM);
        done[q] = 1;
    }
    return VIO_OK;
}

// isolated pose block i: 6x6 solve of (S_ii + lambda I) x = b_i
int bcr_solve_iso(int i, const std::vector<int> &rowptr, const std::vector<int> &col, const double *val, double lambda, const double *b, double *x) {
    double A[36], y[6];
    int kd = -1;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) if (col[k] == i) kd = k;
    for (int e = 0; e < 36; ++e) A[e] = (kd >= 0 ? val[36 * (size_t)kd + e] : 0.0) + (e % 7 == 0 ? lambda : 0.0);
    for (int c = 0; c < 6; ++c) y[c] = b[6 * (size_t)i + c];
    for (int c = 0; c < 6; ++c) {  // Gaussian elimination without pivoting (SPD)
        if (!(A[7 * c] > 0.0)) return VIO_ERR_INVALID;
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[6 * r + c] / A[7 * c];
            for (int k = c; k < 6; ++k) A[6 * r + k] -= f * A[6 * c + k];
            y[r] -= f * y[c];
        }
    }
    for (int c = 5; c >= 0; --c) {
        double a = y[c];
        for (int k = c + 1; k < 6; ++k) a -= A[6 * c + k] * x[6 * (size_t)i + k];
        x[6 * (size_t)i + c] = a / A[7 * c];
    }
    return VIO_OK;
}
}  // namespace

// val: BSR values (nnzb x 36); returns x = (A + lambda I)^-1 b.  info[0..5] = n, w, M, levels, items, slots.
int emul_bcr_solve(int nb, const int *rowptr_, const int *col_, const double *val, double lambda, const double *b, double *x,
                   int *info) {
    std::vector<int> rowptr(rowptr_, rowptr_ + nb + 1), col(col_, col_ + rowptr_[nb]);
    BcrPlan Y;
    bcr_plan(nb, rowptr, col, Y);
    if (!Y.ok) return VIO_ERR_UNSUPPORTED;
    const int n = Y.n, M = Y.M, LD = Y.ld;  // tiles are M x LD (row stride LD)
    const size_t MM = (size_t)M * LD;
    if (info) { info[0] = n; info[1] = Y.w; info[2] = M; info[3] = Y.n_levels; info[4] = (int)Y.items.size(); info[5] = Y.n_slots; }
    std::vector<double> pool(MM * Y.n_slots, 0.0), bv((size_t)n * M, 0.0), xv((size_t)n * M, 0.0);
    // loader: BSR blocks -> node tiles, lambda and identity padding on the diagonal
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (Y.dst[k] < 0) continue;
            const bool dg = (Y.dst[k] & BCR_DST_DIAG) != 0;
            const size_t off = (size_t)(Y.dst[k] & ~BCR_DST_DIAG);
            for (int e = 0; e < 36; ++e)
                pool[off + (size_t)(e / 6) * LD + e % 6] = val[36 * (size_t)k + e] + ((dg && e % 7 == 0) ? lambda : 0.0);
        }
    for (int a = 0; a < n; ++a)
        for (int q = 6 * Y.node_size[a]; q < M; ++q) pool[(size_t)a * MM + (size_t)q * LD + q] = 1.0;
    for (int i = 0; i < nb; ++i)
        if (Y.blk_node[i] >= 0)
            for (int c = 0; c < 6; ++c) bv[(size_t)Y.blk_node[i] * M + 6 * Y.blk_loc[i] + c] = b[6 * (size_t)i + c];
    std::vector<char> done(Y.items.size(), 0);
    int rc = bcr_run_items(Y.items, 0, (int)Y.items.size(), done, M, LD, pool, bv, xv, nullptr);
    if (rc) return rc;
    for (int i = 0; i < nb; ++i) {
        if (Y.blk_node[i] >= 0) {
            for (int c = 0; c < 6; ++c) x[6 * (size_t)i + c] = xv[(size_t)Y.blk_node[i] * M + 6 * Y.blk_loc[i] + c];
        } else {
            rc = bcr_solve_iso(i, rowptr, col, val, lambda, b, x);
            if (rc) return rc;
        }
    }
    return VIO_OK;
}

// The multi-GPU variant (BcrDistPlan) with the `world` ranks played one after the other: every rank gets a SHARE of S and b
// (blocks inside an interface node are split between its two neighbouring ranks, like the landmark shards split them),
// eliminates its own open chain, exports its part of the interface system; the parts are summed (the all-reduce), the
// interface system is solved, every rank back-substitutes its interior.
int emul_bcr_dist_solve(int nb, const int *rowptr_, const int *col_, const double *val, double lambda, const double *b, int world, double *x) {
    std::vector<int> rowptr(rowptr_, rowptr_ + nb + 1), col(col_, col_ + rowptr_[nb]);
    BcrPlan P;
    bcr_plan(nb, rowptr, col, P);
    if (!P.ok) return VIO_ERR_UNSUPPORTED;
    const int n = P.n, M = P.M, LD = P.ld;
    const size_t MM = (size_t)M * LD;
    std::vector<BcrDistPlan> D(world);
    for (int r = 0; r < world; ++r) {
        bcr_dist_plan(P, rowptr, col, r, world, D[r]);
        if (!D[r].ok) return VIO_ERR_UNSUPPORTED;
    }
    const BcrSched &I = D[0].iface;
    // interface pool: tiles [0, world) = D of the interface nodes, [world, 2 world) = their couplings, then W tiles
    std::vector<double> ipool(MM * I.n_slots, 0.0), ibv((size_t)world * M, 0.0), ixv((size_t)world * M, 0.0);
    struct RankState { std::vector<double> pool, bv, xv; std::vector<char> done; };
    std::vector<RankState> R(world);
    // which rank's share a block (or b entry) of an interface node goes to: 30 % to the rank that sees it as its LAST local node
    for (int r = 0; r < world; ++r) {
        const BcrDistPlan &d = D[r];
        const int m = d.m;
        RankState &S = R[r];
        S.pool.assign(MM * d.local.n_slots, 0.0); S.bv.assign((size_t)(m + 1) * M, 0.0); S.xv.assign((size_t)(m + 1) * M, 0.0);
        S.done.assign(d.local.items.size(), 0);
        for (int i = 0; i < nb; ++i)
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                if (d.dst[k] < 0) continue;
                const bool dg = (d.dst[k] & BCR_DST_DIAG) != 0;
                const size_t off = (size_t)(d.dst[k] & ~BCR_DST_DIAG);
                const int a = d.blk_lnode[i], bb = d.blk_lnode[col[k]];
                double share = 1.0;
                if (a == bb && a == 0) share = 0.7;
                if (a == bb && a == m) share = 0.3;
                for (int e = 0; e < 36; ++e)
                    S.pool[off + (size_t)(e / 6) * LD + e % 6] = share * val[36 * (size_t)k + e] + ((dg && e % 7 == 0) ? lambda : 0.0);
            }
        for (int j = 0; j < m; ++j) {  // identity padding of ragged OWNED nodes
            const int a = (d.lo + j) % n;
            for (int q = 6 * P.node_size[a]; q < M; ++q) S.pool[(size_t)j * MM + (size_t)q * LD + q] = 1.0;
        }
        for (int i = 0; i < nb; ++i) {
            const int j = d.blk_lnode[i];
            if (j < 0) continue;
            const double share = j == 0 ? 0.7 : (j == m ? 0.3 : 1.0);
            for (int c = 0; c < 6; ++c) S.bv[(size_t)j * M + 6 * P.blk_loc[i] + c] = share * b[6 * (size_t)i + c];
        }
        std::vector<double> contrib(MM * 2 * world, 0.0);
        int rc = bcr_run_items(d.local.items, 0, d.local.n_elim_items, S.done, M, LD, S.pool, S.bv, S.xv, &contrib);
        if (rc) return rc;
        // the two end nodes' D and b are this rank's share of interface nodes r and r+1
        const int ia = r, ib = (r + 1) % world;
        for (size_t e = 0; e < MM; ++e) { contrib[(size_t)ia * MM + e] += S.pool[e]; contrib[(size_t)ib * MM + e] += S.pool[(size_t)m * MM + e]; }
        for (size_t e = 0; e < contrib.size(); ++e) ipool[e] += contrib[e];  // the all-reduce
        for (int c = 0; c < M; ++c) { ibv[(size_t)ia * M + c] += S.bv[c]; ibv[(size_t)ib * M + c] += S.bv[(size_t)m * M + c]; }
    }
    {
        std::vector<char> done(I.items.size(), 0);
        int rc = bcr_run_items(I.items, 0, (int)I.items.size(), done, M, LD, ipool, ibv, ixv, nullptr);
        if (rc) return rc;
    }
    std::vector<char> have(nb, 0);
    for (int r = 0; r < world; ++r) {
        const BcrDistPlan &d = D[r];
        RankState &S = R[r];
        const int m = d.m;
        for (int c = 0; c < M; ++c) { S.xv[c] = ixv[(size_t)r * M + c]; S.xv[(size_t)m * M + c] = ixv[(size_t)((r + 1) % world) * M + c]; }
        int rc = bcr_run_items(d.local.items, d.local.n_elim_items, (int)d.local.items.size(), S.done, M, LD, S.pool, S.bv, S.xv, nullptr);
        if (rc) return rc;
        for (int i = 0; i < nb; ++i) {
            const int j = d.blk_lnode[i];
            if (j < 0 || j >= m) continue;  // owned nodes only
            for (int c = 0; c < 6; ++c) x[6 * (size_t)i + c] = S.xv[(size_t)j * M + 6 * P.blk_loc[i] + c];
            have[i] = 1;
        }
    }
    for (int i = 0; i < nb; ++i) {
        if (P.blk_node[i] >= 0) { if (!have[i]) return VIO_ERR_STATE; continue; }
        int rc = bcr_solve_iso(i, rowptr, col, val, lambda, b, x);
        if (rc) return rc;
    }
    return VIO_OK;
}

// packer facts for host-side tests: were all landmarks grouped for the shared-memory linearise kernel, how many groups
int emul_pack_info(const vio_graph *g, int *grouped_ok, int *n_groups, int *group_threads, long long *smem_max) {
    PackedGraph K;
    std::string err;
    int rc = pack_graph(g, 0, 1, K, err);
    if (rc) { fprintf(stderr, "emul: %s\n", err.c_str()); return rc; }
    if (grouped_ok) *grouped_ok = K.grouped_ok ? 1 : 0;
    if (n_groups) *n_groups = K.n_groups;
    if (group_threads) *group_threads = K.group_threads;
    if (smem_max) *smem_max = (long long)K.group_smem_max;
    return VIO_OK;
}
}
