/* vio_oracle.c — plain-C CPU restatement of the reference's backend::Problem LM path.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see vio_oracle.h for who may load it and how it is pinned).
 *
 * Reference paths: /root/reference/workspace/assignments/...
 *   A15 = 15-vio-backend, A17 = 17-vins-initialization/vins-mono, EIG = 02-kinematics-in-3D-space/workspace/Eigen
 */
#define _GNU_SOURCE
#include "vio_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ---- Eigen quaternion semantics (xyzw storage) ------------------------------------------------- */
/* EIG/Eigen/src/Geometry/Quaternion.h:470-480  _transformVector: v + w*uv + qv x uv, uv = 2 qv x v */
static void q_rot(const double q[4], const double v[3], double o[3]) {
    double uv[3] = {2 * (q[1] * v[2] - q[2] * v[1]), 2 * (q[2] * v[0] - q[0] * v[2]), 2 * (q[0] * v[1] - q[1] * v[0])};
    o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}
/* EIG/Eigen/src/Geometry/Quaternion.h:659-670  inverse = conjugate / squaredNorm */
static void q_inv(const double q[4], double o[4]) {
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}
/* EIG/Eigen/src/Geometry/Quaternion.h:430-446 */
static void q_mul(const double a[4], const double b[4], double c[4]) {
    c[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    c[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    c[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    c[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
/* EIG/Eigen/src/Geometry/Quaternion.h:531-563 toRotationMatrix */
static void q_toR(const double q[4], double R[9]) {
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void m3_mul(const double A[9], const double B[9], double C[9]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
static void m3_T(const double A[9], double T[9]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[3 * r + c] = A[3 * c + r];
}
static void m3_vec(const double A[9], const double v[3], double o[3]) {
    for (int r = 0; r < 3; ++r) o[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
}
static void hat3(const double v[3], double S[9]) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}

/* ---- Sophus SO3 (A15/thirdparty/Sophus/sophus/so3.hpp) ------------------------------------------- */
/* so3.hpp:393-420 expAndTheta + :682-685 normalising constructor */
static void so3_exp_q(const double w[3], double q[4]) {
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2), im, re;
    if (th < 1e-10) {
        double th4 = th2 * th2;
        im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        re = 1.0 - 0.5 * th2 + (1.0 / 384.0) * th4;
    } else {
        im = sin(0.5 * th) / th;
        re = cos(0.5 * th);
    }
    q[0] = im * w[0]; q[1] = im * w[1]; q[2] = im * w[2]; q[3] = re;
    double len = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= len;
}
/* so3.hpp:541-580 logAndTheta */
static void so3_log_q(const double q[4], double w[3]) {
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], n = sqrt(n2), qw = q[3], f;
    if (n < 1e-10) f = 2.0 / qw - 2.0 * n2 / (qw * qw * qw);
    else if (fabs(qw) < 1e-10) f = (qw > 0 ? M_PI : -M_PI) / n;
    else f = 2.0 * atan(n / qw) / n;
    w[0] = f * q[0]; w[1] = f * q[1]; w[2] = f * q[2];
}
/* so3.hpp:130-145 JacobianRInv (as written: 0.5*hat(k) with the UNIT vector k) */
static void so3_jrinv(const double w[3], double J[9]) {
    double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int k = 0; k < 9; ++k) J[k] = (k % 4 == 0) ? 1.0 : 0.0;
    if (th < 1e-10) return;
    double k3[3] = {w[0] / th, w[1] / th, w[2] / th}, K[9], KK[9];
    hat3(k3, K);
    m3_mul(K, K, KK);
    double c = 1.0 - (1.0 + cos(th)) * th / (2.0 * sin(th));
    for (int k = 0; k < 9; ++k) J[k] += 0.5 * K[k] + c * KK[k];
}

/* VertexPose::Plus — A15/backend/vertex_pose.cc:7-16 (q.normalized() result is discarded there) */
void orc_pose_plus(double *p, const double *d) {
    p[0] += d[0]; p[1] += d[1]; p[2] += d[2];
    double dq[4], q[4] = {p[3], p[4], p[5], p[6]}, qn[4];
    so3_exp_q(d + 3, dq);
    q_mul(q, dq, qn);
    p[3] = qn[0]; p[4] = qn[1]; p[5] = qn[2]; p[6] = qn[3];
}

/* ---- loss functions: A17/src/backend/loss_function.cc:10-47, A17/include/backend/loss_function.h:36-44 */
void orc_loss(int kind, double delta, double e2, double rho[3]) {
    if (kind == VIO_LOSS_HUBER) {
        double dsqr = delta * delta;
        if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1; rho[2] = 0; }
        else { double sq = sqrt(e2); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e2; }
    } else if (kind == VIO_LOSS_CAUCHY) {
        double dsqr = delta * delta, rec = 1.0 / dsqr, aux = rec * e2 + 1.0;
        rho[0] = dsqr * log(aux); rho[1] = 1.0 / aux; rho[2] = -rec * pow(rho[1], 2);
    } else if (kind == VIO_LOSS_TUKEY) {
        double e = sqrt(e2), d2 = delta * delta;
        if (e <= delta) {
            double aux = e2 / d2;
            rho[0] = d2 * (1.0 - pow(1.0 - aux, 3)) / 3.0; rho[1] = pow(1.0 - aux, 2); rho[2] = -2.0 * (1.0 - aux) / d2;
        } else { rho[0] = d2 / 3.0; rho[1] = 0; rho[2] = 0; }
    } else { rho[0] = e2; rho[1] = 1; rho[2] = 0; }
}

/* ---- EdgeReprojection: A15/backend/edge_reprojection.cc:20-40 (residual), :47-91 (Jacobians);
 * the v17 4-vertex edge (A17/src/backend/edge_reprojection.cc:18-108) computes the same three blocks with
 * qic/tic read from the extrinsic vertex (its 4th Jacobian belongs to a fixed vertex and is skipped). */
void orc_reproj(double inv_dep, const double *pi7, const double *pj7, const double *qic, const double *tic,
                const double *pts_i, const double *pts_j, double r[2], double Jl[2], double Ji[12], double Jj[12]) {
    const double Qi[4] = {pi7[3], pi7[4], pi7[5], pi7[6]}, Qj[4] = {pj7[3], pj7[4], pj7[5], pj7[6]};
    double pci[3] = {pts_i[0] / inv_dep, pts_i[1] / inv_dep, pts_i[2] / inv_dep};
    double pbi[3], pw[3], pbj[3], pcj[3], t[3], Qji[4], qici[4];
    q_rot(qic, pci, pbi);
    for (int k = 0; k < 3; ++k) pbi[k] += tic[k];
    q_rot(Qi, pbi, pw);
    for (int k = 0; k < 3; ++k) pw[k] += pi7[k];
    for (int k = 0; k < 3; ++k) t[k] = pw[k] - pj7[k];
    q_inv(Qj, Qji);
    q_rot(Qji, t, pbj);
    for (int k = 0; k < 3; ++k) t[k] = pbj[k] - tic[k];
    q_inv(qic, qici);
    q_rot(qici, t, pcj);
    double dep = pcj[2];
    r[0] = pcj[0] / dep - pts_j[0];
    r[1] = pcj[1] / dep - pts_j[1];
    if (!Jl) return;
    double Ri[9], Rj[9], ric[9], ricT[9], RjT[9];
    q_toR(Qi, Ri); q_toR(Qj, Rj); q_toR(qic, ric);
    m3_T(ric, ricT); m3_T(Rj, RjT);
    double red[6] = {1.0 / dep, 0, -pcj[0] / (dep * dep), 0, 1.0 / dep, -pcj[1] / (dep * dep)};
    double A[9], AR[9], H[9], T[9];
    m3_mul(ricT, RjT, A); /* ric^T Rj^T */
    double ji[18], jj[18];
    m3_mul(A, Ri, AR);
    hat3(pbi, H);
    for (int k = 0; k < 9; ++k) H[k] = -H[k];
    m3_mul(AR, H, T); /* ric^T Rj^T Ri * -hat(pts_imu_i) */
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) { ji[6 * a + b] = A[3 * a + b]; ji[6 * a + 3 + b] = T[3 * a + b]; }
    hat3(pbj, H);
    m3_mul(ricT, H, T);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) { jj[6 * a + b] = -A[3 * a + b]; jj[6 * a + 3 + b] = T[3 * a + b]; }
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 6; ++c) {
            Ji[6 * a + c] = red[3 * a] * ji[c] + red[3 * a + 1] * ji[6 + c] + red[3 * a + 2] * ji[12 + c];
            Jj[6 * a + c] = red[3 * a] * jj[c] + red[3 * a + 1] * jj[6 + c] + red[3 * a + 2] * jj[12 + c];
        }
    /* jacobian_feature = reduce * ric^T Rj^T Ri ric * pts_i * -1/(inv_dep^2) */
    double ARr[9], v3[3];
    m3_mul(AR, ric, ARr);
    m3_vec(ARr, pts_i, v3);
    for (int a = 0; a < 2; ++a)
        Jl[a] = (red[3 * a] * v3[0] + red[3 * a + 1] * v3[1] + red[3 * a + 2] * v3[2]) * -1.0 / (inv_dep * inv_dep);
}

/* Fourth Jacobian of the v17 4-vertex EdgeReprojection, w.r.t. the extrinsic VertexPose
 * (A17/src/backend/edge_reprojection.cc:97-103).  Jex: 2x6 row-major, translation columns first. */
void orc_reproj_jext(double inv_dep, const double *pi7, const double *pj7, const double *qic, const double *tic,
                     const double *pts_i, double Jex[12]) {
    const double Qi[4] = {pi7[3], pi7[4], pi7[5], pi7[6]}, Qj[4] = {pj7[3], pj7[4], pj7[5], pj7[6]};
    double pci[3] = {pts_i[0] / inv_dep, pts_i[1] / inv_dep, pts_i[2] / inv_dep};
    double pbi[3], pw[3], pbj[3], pcj[3], t[3], Qji[4], qici[4];
    q_rot(qic, pci, pbi);
    for (int k = 0; k < 3; ++k) pbi[k] += tic[k];
    q_rot(Qi, pbi, pw);
    for (int k = 0; k < 3; ++k) pw[k] += pi7[k];
    for (int k = 0; k < 3; ++k) t[k] = pw[k] - pj7[k];
    q_inv(Qj, Qji);
    q_rot(Qji, t, pbj);
    for (int k = 0; k < 3; ++k) t[k] = pbj[k] - tic[k];
    q_inv(qic, qici);
    q_rot(qici, t, pcj);
    const double dep = pcj[2];
    const double red[6] = {1.0 / dep, 0, -pcj[0] / (dep * dep), 0, 1.0 / dep, -pcj[1] / (dep * dep)};
    double Ri[9], Rj[9], ric[9], ricT[9], RjT[9], RjTRi[9], M1[9], L[9], tmp_r[9], T1[9], S1[9], S2[9], S3[9], Rr[9], v[3], w[3], u[3];
    q_toR(Qi, Ri); q_toR(Qj, Rj); q_toR(qic, ric);
    m3_T(ric, ricT); m3_T(Rj, RjT);
    m3_mul(RjT, Ri, RjTRi);
    for (int k = 0; k < 9; ++k) M1[k] = RjTRi[k] - ((k % 4 == 0) ? 1.0 : 0.0);
    m3_mul(ricT, M1, L);                       /* ric^T (Rj^T Ri - I) */
    m3_mul(ricT, RjTRi, T1); m3_mul(T1, ric, tmp_r); /* tmp_r = ric^T Rj^T Ri ric */
    hat3(pci, S1);
    m3_mul(tmp_r, S1, Rr);                     /* tmp_r skew(p_ci) */
    m3_vec(tmp_r, pci, v); hat3(v, S2);        /* skew(tmp_r p_ci) */
    m3_vec(Ri, tic, w);
    for (int k = 0; k < 3; ++k) w[k] += pi7[k] - pj7[k];
    m3_vec(RjT, w, u);
    for (int k = 0; k < 3; ++k) u[k] -= tic[k];
    m3_vec(ricT, u, v); hat3(v, S3);           /* skew(ric^T (Rj^T (Ri tic + Pi - Pj) - tic)) */
    double je[18];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) { je[6 * a + b] = L[3 * a + b]; je[6 * a + 3 + b] = -Rr[3 * a + b] + S2[3 * a + b] + S3[3 * a + b]; }
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 6; ++c) Jex[6 * a + c] = red[3 * a] * je[c] + red[3 * a + 1] * je[6 + c] + red[3 * a + 2] * je[12 + c];
}

/* ---- EdgeSE3Prior: A15/backend/edge_prior.cpp:39-80 (USE_SO3_JACOBIAN) ------------------------------ */
void orc_se3prior(const double *pose, const double *pp, const double *qp, double r[6], double J[36]) {
    double qi[4] = {pose[3], pose[4], pose[5], pose[6]}, qn[4] = {qp[0], qp[1], qp[2], qp[3]};
    double ni = sqrt(qi[0] * qi[0] + qi[1] * qi[1] + qi[2] * qi[2] + qi[3] * qi[3]);
    double np = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int k = 0; k < 4; ++k) { qi[k] /= ni; qn[k] /= np; }
    double qc[4] = {-qn[0], -qn[1], -qn[2], qn[3]}, qr[4];
    q_mul(qc, qi, qr);
    double nr = sqrt(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
    for (int k = 0; k < 4; ++k) qr[k] /= nr;
    so3_log_q(qr, r);
    for (int k = 0; k < 3; ++k) r[3 + k] = pose[k] - pp[k];
    if (!J) return;
    double Jr[9];
    so3_jrinv(r, Jr);
    memset(J, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) J[6 * a + 3 + b] = Jr[3 * a + b];
    J[18] = 1; J[25] = 1; J[32] = 1;
}

/* ---- EdgeImu: residual = IntegrationBase::evaluate (A17/include/factor/integration_base.h:160-186),
 * Jacobians = EdgeImu::ComputeJacobians (A17/src/backend/edge_imu.cc:38-157); helpers from
 * A17/include/utility/utility.h:11-64.  J is 15 x 30: [pose_i(6) | speedbias_i(9) | pose_j(6) | speedbias_j(9)] */
static void qleft33(const double q[4], double M[9]) {
    hat3(q, M);
    M[0] += q[3]; M[4] += q[3]; M[8] += q[3];
}
void orc_imu(const double *pi, const double *si, const double *pj, const double *sj, double dt, const double *dp,
             const double *dq, const double *dv, const double *lba, const double *lbg, const double *jac,
             const double *G, double r[15], double *J) {
    const double Qi[4] = {pi[3], pi[4], pi[5], pi[6]}, Qj[4] = {pj[3], pj[4], pj[5], pj[6]};
    double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            dp_dba[3 * a + b] = jac[15 * a + 9 + b];
            dp_dbg[3 * a + b] = jac[15 * a + 12 + b];
            dq_dbg[3 * a + b] = jac[15 * (3 + a) + 12 + b];
            dv_dba[3 * a + b] = jac[15 * (6 + a) + 9 + b];
            dv_dbg[3 * a + b] = jac[15 * (6 + a) + 12 + b];
        }
    double dba[3], dbg[3], th[3], t1[3], t2[3], cdp[3], cdv[3];
    for (int k = 0; k < 3; ++k) { dba[k] = si[3 + k] - lba[k]; dbg[k] = si[6 + k] - lbg[k]; }
    m3_vec(dq_dbg, dbg, th);
    double dQ[4] = {th[0] / 2.0, th[1] / 2.0, th[2] / 2.0, 1.0}, cdq[4];
    q_mul(dq, dQ, cdq);
    m3_vec(dv_dba, dba, t1); m3_vec(dv_dbg, dbg, t2);
    for (int k = 0; k < 3; ++k) cdv[k] = dv[k] + t1[k] + t2[k];
    m3_vec(dp_dba, dba, t1); m3_vec(dp_dbg, dbg, t2);
    for (int k = 0; k < 3; ++k) cdp[k] = dp[k] + t1[k] + t2[k];
    double Qii[4], a1[3], a2[3], u1[3], u2[3];
    q_inv(Qi, Qii);
    for (int k = 0; k < 3; ++k) {
        a1[k] = 0.5 * G[k] * dt * dt + pj[k] - pi[k] - si[k] * dt;
        a2[k] = G[k] * dt + sj[k] - si[k];
    }
    q_rot(Qii, a1, u1);
    q_rot(Qii, a2, u2);
    double cdqi[4], QiiQj[4], qr[4];
    q_inv(cdq, cdqi);
    q_mul(Qii, Qj, QiiQj);
    q_mul(cdqi, QiiQj, qr);
    for (int k = 0; k < 3; ++k) {
        r[k] = u1[k] - cdp[k];
        r[3 + k] = 2 * qr[k];
        r[6 + k] = u2[k] - cdv[k];
        r[9 + k] = sj[3 + k] - si[3 + k];
        r[12 + k] = sj[6 + k] - si[6 + k];
    }
    if (!J) return;
    memset(J, 0, 450 * sizeof(double));
#define PUT(r0, c0, M, sgn)                                                   \
    for (int a_ = 0; a_ < 3; ++a_)                                            \
        for (int b_ = 0; b_ < 3; ++b_) J[30 * ((r0) + a_) + (c0) + b_] = (sgn) * (M)[3 * a_ + b_];
    double Rii[9], S1[9], S2[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    q_toR(Qii, Rii);
    hat3(u1, S1); hat3(u2, S2);
    PUT(0, 0, Rii, -1.0);
    PUT(0, 3, S1, 1.0);
    {
        double Qji[4], QjiQi[4], Lq[9], Rr[9], T[9], q3[4], L3[9], T3[9];
        q_inv(Qj, Qji);
        q_mul(Qji, Qi, QjiQi);
        qleft33(QjiQi, Lq);
        hat3(cdq, Rr);
        for (int k = 0; k < 9; ++k) Rr[k] = -Rr[k];
        Rr[0] += cdq[3]; Rr[4] += cdq[3]; Rr[8] += cdq[3];
        m3_mul(Lq, Rr, T);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) T[3 * a + b] -= QjiQi[a] * cdq[b];
        PUT(3, 3, T, -1.0);
        q_mul(QjiQi, dq, q3);
        qleft33(q3, L3);
        m3_mul(L3, dq_dbg, T3);
        PUT(3, 12, T3, -1.0);
    }
    PUT(6, 3, S2, 1.0);
    {
        double Rdt[9];
        for (int k = 0; k < 9; ++k) Rdt[k] = Rii[k] * dt;
        PUT(0, 6, Rdt, -1.0);
        PUT(0, 9, dp_dba, -1.0);
        PUT(0, 12, dp_dbg, -1.0);
        PUT(6, 6, Rii, -1.0);
        PUT(6, 9, dv_dba, -1.0);
        PUT(6, 12, dv_dbg, -1.0);
        PUT(9, 9, I3, -1.0);
        PUT(12, 12, I3, -1.0);
    }
    PUT(0, 15, Rii, 1.0);
    {
        double Lq[9];
        qleft33(qr, Lq);
        PUT(3, 18, Lq, 1.0);
    }
    PUT(6, 21, Rii, 1.0);
    PUT(9, 24, I3, 1.0);
    PUT(12, 27, I3, 1.0);
#undef PUT
}

/* ---- small dense helpers ---------------------------------------------------------------------------- */
/* general inverse with partial pivoting (Eigen MatrixXd::inverse() is PartialPivLU based) */
static int mat_inverse(int n, const double *A, double *Ai) {
    double *W = (double *)malloc(sizeof(double) * n * 2 * n);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) { W[r * 2 * n + c] = A[r * n + c]; W[r * 2 * n + n + c] = (r == c); }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (fabs(W[r * 2 * n + k]) > fabs(W[piv * 2 * n + k])) piv = r;
        if (piv != k)
            for (int c = 0; c < 2 * n; ++c) { double t = W[k * 2 * n + c]; W[k * 2 * n + c] = W[piv * 2 * n + c]; W[piv * 2 * n + c] = t; }
        double d = 1.0 / W[k * 2 * n + k];
        for (int c = 0; c < 2 * n; ++c) W[k * 2 * n + c] *= d;
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            double f = W[r * 2 * n + k];
            if (f != 0.0)
                for (int c = 0; c < 2 * n; ++c) W[r * 2 * n + c] -= f * W[k * 2 * n + c];
        }
    }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) Ai[r * n + c] = W[r * 2 * n + n + c];
    free(W);
    return 0;
}
/* Cholesky solve of an SPD system (stands in for Eigen LDLT: any backward-stable factorisation agrees to k(S) eps) */
static int chol_solve(int n, const double *A, const double *b, double *x) {
    double *L = (double *)malloc(sizeof(double) * (size_t)n * n);
    memcpy(L, A, sizeof(double) * (size_t)n * n);
    int bad = 0;
    for (int k = 0; k < n; ++k) {
        double d = L[(size_t)k * n + k];
        if (!(d > 0)) bad = 1;
        d = sqrt(d);
        L[(size_t)k * n + k] = d;
        for (int i = k + 1; i < n; ++i) L[(size_t)i * n + k] /= d;
        for (int i = k + 1; i < n; ++i) {
            double lik = L[(size_t)i * n + k];
            if (lik == 0.0) continue;
            for (int j = k + 1; j <= i; ++j) L[(size_t)i * n + j] -= lik * L[(size_t)j * n + k];
        }
    }
    for (int i = 0; i < n; ++i) {
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= L[(size_t)i * n + k] * x[k];
        x[i] = t / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double t = x[i];
        for (int k = i + 1; k < n; ++k) t -= L[(size_t)k * n + i] * x[k];
        x[i] = t / L[(size_t)i * n + i];
    }
    free(L);
    return bad;
}

/* Problem::PCGSolver — A15/backend/problem.cc:530-560, verbatim including the missing first `x += alpha*p` */
static int ref_pcg(int n, const double *A, const double *b, int max_iter, double *x) {
    double *minv = malloc(sizeof(double) * n), *r0 = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n),
           *w = malloc(sizeof(double) * n), *r1 = malloc(sizeof(double) * n);
    double r0z0 = 0, pw = 0, r0n = 0;
    for (int i = 0; i < n; ++i) {
        x[i] = 0; minv[i] = 1.0 / A[(size_t)i * n + i]; r0[i] = b[i];
        p[i] = minv[i] * r0[i]; r0z0 += r0[i] * p[i]; r0n += r0[i] * r0[i];
    }
    for (int i = 0; i < n; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += A[(size_t)i * n + j] * p[j]; w[i] = t; pw += p[i] * t; }
    double alpha = r0z0 / pw, thr = 1e-6 * sqrt(r0n), rn = 0;
    for (int i = 0; i < n; ++i) { r1[i] = r0[i] - alpha * w[i]; rn += r1[i] * r1[i]; }
    int it = 0;
    while (sqrt(rn) > thr && it < max_iter) {
        it++;
        double r1z1 = 0;
        for (int i = 0; i < n; ++i) r1z1 += r1[i] * (minv[i] * r1[i]);
        double beta = r1z1 / r0z0;
        r0z0 = r1z1;
        for (int i = 0; i < n; ++i) p[i] = beta * p[i] + minv[i] * r1[i];
        pw = 0;
        for (int i = 0; i < n; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += A[(size_t)i * n + j] * p[j]; w[i] = t; pw += p[i] * t; }
        alpha = r1z1 / pw;
        rn = 0;
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r1[i] -= alpha * w[i]; rn += r1[i] * r1[i]; }
    }
    free(minv); free(r0); free(p); free(w); free(r1);
    return it;
}

/* ---- ordering: Problem::SetOrdering — A15/backend/problem.cc:224-262, A17/src/backend/problem.cc:256-285 */
typedef struct {
    int P, M, NB;
    int *pose_off, *sb_off;
    unsigned char *row_fixed;
} ordering;

static int make_ordering(const vio_graph *g, ordering *o) {
    int C = g->n_pose, NSB = g->n_speedbias, NB = C + NSB, P = 0;
    o->pose_off = (int *)malloc(sizeof(int) * (C > 0 ? C : 1));
    o->sb_off = (int *)malloc(sizeof(int) * (NSB > 0 ? NSB : 1));
    for (int i = 0; i < C; ++i) o->pose_off[i] = -1;
    for (int i = 0; i < NSB; ++i) o->sb_off[i] = -1;
    for (int k = 0; k < NB; ++k) {
        int ent = g->pclass_order ? g->pclass_order[k] : (k < C ? k : ~(k - C));
        if (ent >= 0) { if (ent >= C || o->pose_off[ent] >= 0) return VIO_ERR_INVALID; o->pose_off[ent] = P; P += 6; }
        else { int i = ~ent; if (i >= NSB || o->sb_off[i] >= 0) return VIO_ERR_INVALID; o->sb_off[i] = P; P += 9; }
    }
    o->P = P; o->M = g->n_landmark + 3 * g->n_point; o->NB = NB; /* [inverse depths | 3 per VertexPointXYZ] */
    o->row_fixed = (unsigned char *)calloc(P > 0 ? P : 1, 1);
    for (int i = 0; i < C; ++i)
        if (g->pose_fixed && g->pose_fixed[i]) for (int d = 0; d < 6; ++d) o->row_fixed[o->pose_off[i] + d] = 1;
    for (int i = 0; i < NSB; ++i)
        if (g->speedbias_fixed && g->speedbias_fixed[i]) for (int d = 0; d < 9; ++d) o->row_fixed[o->sb_off[i] + d] = 1;
    return VIO_OK;
}
static void free_ordering(ordering *o) { free(o->pose_off); free(o->sb_off); free(o->row_fixed); }

int orc_dims(const vio_graph *g, int32_t *P, int32_t *M) {
    ordering o;
    int rc = make_ordering(g, &o);
    if (rc) return rc;
    *P = o.P; *M = o.M;
    free_ordering(&o);
    return VIO_OK;
}

static void get_ext(const vio_graph *g, double qic[4], double tic[3]) {
    if (g->ext_pose >= 0) {
        const double *e = g->pose + 7 * (size_t)g->ext_pose;
        tic[0] = e[0]; tic[1] = e[1]; tic[2] = e[2]; qic[0] = e[3]; qic[1] = e[4]; qic[2] = e[5]; qic[3] = e[6];
    } else {
        for (int k = 0; k < 4; ++k) qic[k] = g->q_ic[k];
        for (int k = 0; k < 3; ++k) tic[k] = g->t_ic[k];
    }
}

/* Edge::RobustInfo for information = c*I2 — A17/src/backend/edge.cc:50-74.  W row-major 2x2. */
static void robust_info2(int loss, double delta, double c, const double r[2], double *drho, double W[4], double *rho0) {
    double e2 = c * (r[0] * r[0] + r[1] * r[1]);
    if (loss == VIO_LOSS_TRIVIAL) { *drho = 1; W[0] = c; W[1] = 0; W[2] = 0; W[3] = c; *rho0 = e2; return; }
    double rho[3];
    orc_loss(loss, delta, e2, rho);
    double sc = sqrt(c), we[2] = {sc * r[0], sc * r[1]};
    double ri[4] = {rho[1], 0, 0, rho[1]};
    if (rho[1] + 2 * rho[2] * e2 > 0.0) {
        ri[0] += 2 * rho[2] * we[0] * we[0]; ri[1] += 2 * rho[2] * we[0] * we[1];
        ri[2] += 2 * rho[2] * we[1] * we[0]; ri[3] += 2 * rho[2] * we[1] * we[1];
    }
    for (int k = 0; k < 4; ++k) W[k] = ri[k] * c;
    *drho = rho[1];
    *rho0 = rho[0];
}

/* accumulate one edge into dense H/b exactly like the double loop of MakeHessian:
 * A15/backend/problem.cc:296-325 / A17/src/backend/problem.cc:319-358.
 * nv vertices with Jacobians Jv[i] (d x dim_i, row-major, leading dim ld_i), offsets off[i] (<0: fixed -> skipped),
 * W (d x d) for H, and Wb (d x d) with factor for b: b_i -= Jv_i^T * Wb * r */
static void add_edge_dense(double *H, double *b, int n, int d, int nv, const double *const *Jv, const int *ldj,
                           const int *dim, const int *off, const double *W, const double *Wb, double bscale,
                           const double *r) {
    double JtW[15 * 9];
    for (int i = 0; i < nv; ++i) {
        if (off[i] < 0) continue;
        for (int a = 0; a < dim[i]; ++a)
            for (int c = 0; c < d; ++c) {
                double t = 0;
                for (int k = 0; k < d; ++k) t += Jv[i][k * ldj[i] + a] * W[k * d + c];
                JtW[a * d + c] = t;
            }
        for (int j = i; j < nv; ++j) {
            if (off[j] < 0) continue;
            for (int a = 0; a < dim[i]; ++a)
                for (int c = 0; c < dim[j]; ++c) {
                    double t = 0;
                    for (int k = 0; k < d; ++k) t += JtW[a * d + k] * Jv[j][k * ldj[j] + c];
                    H[(size_t)(off[i] + a) * n + off[j] + c] += t;
                    if (j != i) H[(size_t)(off[j] + c) * n + off[i] + a] += t;
                }
        }
        for (int a = 0; a < dim[i]; ++a) {
            double t = 0;
            for (int k = 0; k < d; ++k) {
                double wr = 0;
                for (int c = 0; c < d; ++c) wr += Wb[k * d + c] * r[c];
                t += Jv[i][k * ldj[i] + a] * wr;
            }
            b[off[i] + a] -= bscale * t;
        }
    }
}

/* EdgeReprojectionXYZ — A15/backend/edge_reprojection.cc:113-163 (A17/src/backend/edge_reprojection.cc:130-180):
 * vertices [X_w(3), T_i]; r (2), JX (2x3 row-major), JT (2x6 row-major, translation columns first). */
void orc_reproj_xyz(const double *X, const double *pose_i, const double qic[4], const double tic[3], const double *obs,
                    double r[2], double *JX, double *JT) {
    const double *Pi = pose_i, *Qi = pose_i + 3;
    double qinv[4], qicinv[4], d[3] = {X[0] - Pi[0], X[1] - Pi[1], X[2] - Pi[2]}, pb[3], e[3], pc[3];
    q_inv(Qi, qinv); q_inv(qic, qicinv);
    q_rot(qinv, d, pb);
    e[0] = pb[0] - tic[0]; e[1] = pb[1] - tic[1]; e[2] = pb[2] - tic[2];
    q_rot(qicinv, e, pc);
    const double dep = pc[2];
    r[0] = pc[0] / dep - obs[0];
    r[1] = pc[1] / dep - obs[1];
    if (!JX && !JT) return;
    double Ri[9], ric[9], ricT[9], RiT[9], A[9], H3[9], Hh[9];
    q_toR(Qi, Ri); q_toR(qic, ric);
    m3_T(ric, ricT); m3_T(Ri, RiT);
    const double red[6] = {1.0 / dep, 0, -pc[0] / (dep * dep), 0, 1.0 / dep, -pc[1] / (dep * dep)};
    m3_mul(ricT, RiT, A);      /* ric^T Ri^T */
    hat3(pb, H3);
    m3_mul(ricT, H3, Hh);      /* ric^T hat(p_b) */
    for (int c = 0; c < 3; ++c) {
        const double jx0 = red[0] * A[c] + red[1] * A[3 + c] + red[2] * A[6 + c];
        const double jx1 = red[3] * A[c] + red[4] * A[3 + c] + red[5] * A[6 + c];
        if (JX) { JX[c] = jx0; JX[3 + c] = jx1; }
        if (JT) {
            JT[c] = -jx0; JT[6 + c] = -jx1;
            JT[3 + c] = red[0] * Hh[c] + red[1] * Hh[3 + c] + red[2] * Hh[6 + c];
            JT[9 + c] = red[3] * Hh[c] + red[4] * Hh[3 + c] + red[5] * Hh[6 + c];
        }
    }
}

int orc_make_hessian(const vio_graph *g, const orc_prior *prior, int flavour, double *H, double *b) {
    ordering o;
    int rc = make_ordering(g, &o);
    if (rc) return rc;
    const int P = o.P, M = o.M, n = P + M;
    memset(H, 0, sizeof(double) * (size_t)n * n);
    memset(b, 0, sizeof(double) * n);
    double qic[4], tic[3];
    get_ext(g, qic, tic);
    for (int64_t e = 0; e < g->n_reproj; ++e) {
        int l = g->rp_landmark[e], i = g->rp_pose_i[e], j = g->rp_pose_j[e];
        double r[2], Jl[2], Ji[12], Jj[12];
        orc_reproj(g->inv_depth[l], g->pose + 7 * (size_t)i, g->pose + 7 * (size_t)j, qic, tic, g->rp_pts_i + 3 * e,
                   g->rp_pts_j + 2 * e, r, Jl, Ji, Jj);
        double W[4], Om[4] = {g->rp_info, 0, 0, g->rp_info}, drho = 1.0, rho0;
        if (flavour == VIO_LM_V17) robust_info2(g->rp_loss, g->rp_loss_delta, g->rp_info, r, &drho, W, &rho0);
        else memcpy(W, Om, sizeof(W));
        const int ext_free = g->ext_pose >= 0 && !(g->pose_fixed && g->pose_fixed[g->ext_pose]);
        double Jex[12];
        if (ext_free) orc_reproj_jext(g->inv_depth[l], g->pose + 7 * (size_t)i, g->pose + 7 * (size_t)j, qic, tic, g->rp_pts_i + 3 * e, Jex);
        const double *Jv[4] = {Jl, Ji, Jj, Jex};
        int ldj[4] = {1, 6, 6, 6}, dim[4] = {1, 6, 6, 6};
        int off[4] = {P + l, (g->pose_fixed && g->pose_fixed[i]) ? -1 : o.pose_off[i],
                      (g->pose_fixed && g->pose_fixed[j]) ? -1 : o.pose_off[j], ext_free ? o.pose_off[g->ext_pose] : -1};
        /* v15: b -= JtW r with W = information; v17: b -= drho * J^T * information * r */
        add_edge_dense(H, b, n, 2, ext_free ? 4 : 3, Jv, ldj, dim, off, W, flavour == VIO_LM_V17 ? Om : W, drho, r);
    }
    for (int64_t e = 0; e < g->n_reproj_xyz; ++e) {
        int l = g->rx_point[e], i = g->rx_pose[e];
        double r[2], JX[6], JT[12];
        orc_reproj_xyz(g->point_xyz + 3 * (size_t)l, g->pose + 7 * (size_t)i, qic, tic, g->rx_obs + 2 * e, r, JX, JT);
        double W[4], Om[4] = {g->rp_info, 0, 0, g->rp_info}, drho = 1.0, rho0;
        if (flavour == VIO_LM_V17) robust_info2(g->rp_loss, g->rp_loss_delta, g->rp_info, r, &drho, W, &rho0);
        else memcpy(W, Om, sizeof(W));
        const double *Jv[2] = {JX, JT};
        int ldj[2] = {3, 6}, dim[2] = {3, 6};
        int off[2] = {P + g->n_landmark + 3 * l, (g->pose_fixed && g->pose_fixed[i]) ? -1 : o.pose_off[i]};
        add_edge_dense(H, b, n, 2, 2, Jv, ldj, dim, off, W, flavour == VIO_LM_V17 ? Om : W, drho, r);
    }
    for (int k = 0; k < g->n_se3prior; ++k) {
        int i = g->sp_pose[k];
        double r[6], J[36];
        orc_se3prior(g->pose + 7 * (size_t)i, g->sp_p + 3 * k, g->sp_q + 4 * k, r, J);
        const double *Jv[1] = {J};
        int ldj[1] = {6}, dim[1] = {6}, off[1] = {(g->pose_fixed && g->pose_fixed[i]) ? -1 : o.pose_off[i]};
        add_edge_dense(H, b, n, 6, 1, Jv, ldj, dim, off, g->sp_info + 36 * k, g->sp_info + 36 * k, 1.0, r);
    }
    for (int k = 0; k < g->n_imu; ++k) {
        int pi = g->imu_pose_i[k], si = g->imu_sb_i[k], pj = g->imu_pose_j[k], sj = g->imu_sb_j[k];
        double r[15], J[450], info[225];
        orc_imu(g->pose + 7 * (size_t)pi, g->speedbias + 9 * (size_t)si, g->pose + 7 * (size_t)pj,
                g->speedbias + 9 * (size_t)sj, g->imu_sum_dt[k], g->imu_delta_p + 3 * k, g->imu_delta_q + 4 * k,
                g->imu_delta_v + 3 * k, g->imu_lin_ba + 3 * k, g->imu_lin_bg + 3 * k, g->imu_jacobian + 225 * k,
                g->gravity, r, J);
        mat_inverse(15, g->imu_covariance + 225 * k, info); /* SetInformation(covariance.inverse()), edge_imu.cc:35 */
        const double *Jv[4] = {J, J + 6, J + 15, J + 21};
        int ldj[4] = {30, 30, 30, 30}, dim[4] = {6, 9, 6, 9};
        int off[4] = {(g->pose_fixed && g->pose_fixed[pi]) ? -1 : o.pose_off[pi],
                      (g->speedbias_fixed && g->speedbias_fixed[si]) ? -1 : o.sb_off[si],
                      (g->pose_fixed && g->pose_fixed[pj]) ? -1 : o.pose_off[pj],
                      (g->speedbias_fixed && g->speedbias_fixed[sj]) ? -1 : o.sb_off[sj]};
        add_edge_dense(H, b, n, 15, 4, Jv, ldj, dim, off, info, info, 1.0, r);
    }
    /* prior: A17/src/backend/problem.cc:365-384 (rows/cols of fixed pose-class vertices zeroed in a copy) */
    if (flavour == VIO_LM_V17 && prior && prior->dim > 0) {
        if (prior->dim != P) { free_ordering(&o); return VIO_ERR_INVALID; }
        for (int r = 0; r < P; ++r) {
            if (o.row_fixed[r]) continue;
            for (int c = 0; c < P; ++c)
                if (!o.row_fixed[c]) H[(size_t)r * n + c] += prior->H[(size_t)r * P + c];
            b[r] += prior->b[r];
        }
    }
    free_ordering(&o);
    return VIO_OK;
}

static double vec_norm(const double *x, int n) {
    double t = 0;
    for (int i = 0; i < n; ++i) t += x[i] * x[i];
    return sqrt(t);
}

int orc_chi2(const vio_graph *g, const orc_prior *prior, int flavour, double *chi2) {
    double qic[4], tic[3], chi = 0;
    get_ext(g, qic, tic);
    for (int64_t e = 0; e < g->n_reproj; ++e) {
        double r[2];
        orc_reproj(g->inv_depth[g->rp_landmark[e]], g->pose + 7 * (size_t)g->rp_pose_i[e],
                   g->pose + 7 * (size_t)g->rp_pose_j[e], qic, tic, g->rp_pts_i + 3 * e, g->rp_pts_j + 2 * e, r, NULL, NULL, NULL);
        double e2 = g->rp_info * (r[0] * r[0] + r[1] * r[1]);
        if (flavour == VIO_LM_V17 && g->rp_loss != VIO_LOSS_TRIVIAL) {
            double rho[3];
            orc_loss(g->rp_loss, g->rp_loss_delta, e2, rho);
            e2 = rho[0];
        }
        chi += e2;
    }
    for (int64_t e = 0; e < g->n_reproj_xyz; ++e) {
        double r[2];
        orc_reproj_xyz(g->point_xyz + 3 * (size_t)g->rx_point[e], g->pose + 7 * (size_t)g->rx_pose[e], qic, tic, g->rx_obs + 2 * e,
                       r, NULL, NULL);
        double e2 = g->rp_info * (r[0] * r[0] + r[1] * r[1]);
        if (flavour == VIO_LM_V17 && g->rp_loss != VIO_LOSS_TRIVIAL) {
            double rho[3];
            orc_loss(g->rp_loss, g->rp_loss_delta, e2, rho);
            e2 = rho[0];
        }
        chi += e2;
    }
    for (int k = 0; k < g->n_se3prior; ++k) {
        double r[6];
        orc_se3prior(g->pose + 7 * (size_t)g->sp_pose[k], g->sp_p + 3 * k, g->sp_q + 4 * k, r, NULL);
        const double *Om = g->sp_info + 36 * k;
        for (int a = 0; a < 6; ++a) { double t = 0; for (int c = 0; c < 6; ++c) t += Om[6 * a + c] * r[c]; chi += r[a] * t; }
    }
    for (int k = 0; k < g->n_imu; ++k) {
        double r[15], info[225];
        orc_imu(g->pose + 7 * (size_t)g->imu_pose_i[k], g->speedbias + 9 * (size_t)g->imu_sb_i[k],
                g->pose + 7 * (size_t)g->imu_pose_j[k], g->speedbias + 9 * (size_t)g->imu_sb_j[k], g->imu_sum_dt[k],
                g->imu_delta_p + 3 * k, g->imu_delta_q + 4 * k, g->imu_delta_v + 3 * k, g->imu_lin_ba + 3 * k,
                g->imu_lin_bg + 3 * k, g->imu_jacobian + 225 * k, g->gravity, r, NULL);
        mat_inverse(15, g->imu_covariance + 225 * k, info);
        for (int a = 0; a < 15; ++a) { double t = 0; for (int c = 0; c < 15; ++c) t += info[15 * a + c] * r[c]; chi += r[a] * t; }
    }
    /* err_prior_.norm() — norm, not squared (A15/backend/problem.cc:461-462, A17/src/backend/problem.cc:505-506) */
    if (flavour == VIO_LM_V17 && prior && prior->err_dim > 0) chi += vec_norm(prior->err, prior->err_dim);
    if (flavour == VIO_LM_V17) chi *= 0.5;
    *chi2 = chi;
    return VIO_OK;
}

/* Problem::SolveLinearSystem (SLAM branch) — A15/backend/problem.cc:353-421, A17/src/backend/problem.cc:406-449.
 * Hmm is diagonal for inverse-depth landmarks, so Hpm*Hmm_inv is a column scaling. */
/* Landmark block structure [M1 scalars | Mx blocks of 3]: Hmm_inv = per-landmark block inverse
 * (A15/backend/problem.cc:383-388: Hmm.block(idx, idx, size, size).inverse()). */
int orc_solve_linear_blocks(const double *H, const double *b, int P, int M1, int Mx, double lambda, int solver, double *S,
                            double *bS, double *dx, int64_t *pcg_iters) {
    const int M = M1 + 3 * Mx, n = P + M;
    double *Sl = S ? S : (double *)malloc(sizeof(double) * (size_t)P * P);
    double *bl = bS ? bS : (double *)malloc(sizeof(double) * P);
    /* tempH = Hpm * Hmm_inv (P x M), built block by block */
    double *T = (double *)malloc(sizeof(double) * (size_t)P * (M > 0 ? M : 1));
    double *hinv = (double *)malloc(sizeof(double) * (M1 + 9 * (size_t)Mx + 1));
    for (int l = 0; l < M1; ++l) hinv[l] = 1.0 / H[(size_t)(P + l) * n + P + l];
    for (int l = 0; l < Mx; ++l) {
        double A[9], Ai[9];
        const int g0 = P + M1 + 3 * l;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) A[3 * r + c] = H[(size_t)(g0 + r) * n + g0 + c];
        mat_inverse(3, A, Ai);
        memcpy(hinv + M1 + 9 * (size_t)l, Ai, sizeof(Ai));
    }
    for (int r = 0; r < P; ++r) {
        const double *hr = H + (size_t)r * n;
        double *tr = T + (size_t)r * M;
        for (int l = 0; l < M1; ++l) tr[l] = hr[P + l] * hinv[l];
        for (int l = 0; l < Mx; ++l) {
            const double *Ai = hinv + M1 + 9 * (size_t)l, *h3 = hr + P + M1 + 3 * l;
            for (int c = 0; c < 3; ++c) tr[M1 + 3 * l + c] = h3[0] * Ai[c] + h3[1] * Ai[3 + c] + h3[2] * Ai[6 + c];
        }
    }
    for (int r = 0; r < P; ++r) {
        const double *tr = T + (size_t)r * M;
        for (int c = 0; c < P; ++c) {
            double t = 0;
            const double *hc = H + (size_t)c * n + P; /* Hmp = Hpm^T */
            for (int l = 0; l < M; ++l) t += tr[l] * hc[l];
            Sl[(size_t)r * P + c] = H[(size_t)r * n + c] - t;
        }
        double t = 0;
        for (int l = 0; l < M; ++l) t += tr[l] * b[P + l];
        bl[r] = b[r] - t;
    }
    for (int i = 0; i < P; ++i) Sl[(size_t)i * P + i] += lambda;
    int64_t it = 0;
    if (solver == VIO_SOLVER_REF_PCG) it = ref_pcg(P, Sl, bl, 2 * P, dx);
    else chol_solve(P, Sl, bl, dx);
    if (pcg_iters) *pcg_iters = it;
    /* delta_x_ll = Hmm_inv * (bmm - Hmp * delta_x_pp) */
    for (int l = 0; l < M1; ++l) {
        double t = b[P + l];
        const double *hl = H + (size_t)(P + l) * n;
        for (int c = 0; c < P; ++c) t -= hl[c] * dx[c];
        dx[P + l] = hinv[l] * t;
    }
    for (int l = 0; l < Mx; ++l) {
        const int g0 = P + M1 + 3 * l;
        double t[3];
        for (int a = 0; a < 3; ++a) {
            t[a] = b[g0 + a];
            const double *hl = H + (size_t)(g0 + a) * n;
            for (int c = 0; c < P; ++c) t[a] -= hl[c] * dx[c];
        }
        const double *Ai = hinv + M1 + 9 * (size_t)l;
        for (int a = 0; a < 3; ++a) dx[g0 + a] = Ai[3 * a] * t[0] + Ai[3 * a + 1] * t[1] + Ai[3 * a + 2] * t[2];
    }
    if (!S) free(Sl);
    if (!bS) free(bl);
    free(T); free(hinv);
    return VIO_OK;
}
int orc_solve_linear(const double *H, const double *b, int P, int M, double lambda, int solver, double *S, double *bS,
                     double *dx, int64_t *pcg_iters) {
    return orc_solve_linear_blocks(H, b, P, M, 0, lambda, solver, S, bS, dx, pcg_iters);
}

/* ---- Problem::Solve — A15/backend/problem.cc:155-222 (v15), A17/src/backend/problem.cc:169-250 (v17) ---- */
typedef struct {
    vio_graph g; /* shallow copy whose state arrays point at the buffers below */
    double *pose, *sb, *invd, *pose_bak, *sb_bak, *invd_bak, *pts, *pts_bak;
    orc_prior prior;
    double *bprior, *err, *bprior_bak, *err_bak;
} lm_state;

static void update_states(lm_state *s, const ordering *o, const double *dx, int flavour, double sign, int backup) {
    const vio_graph *g = &s->g;
    if (backup) {
        memcpy(s->pose_bak, s->pose, sizeof(double) * 7 * (size_t)g->n_pose);
        memcpy(s->sb_bak, s->sb, sizeof(double) * 9 * (size_t)g->n_speedbias);
        memcpy(s->invd_bak, s->invd, sizeof(double) * (size_t)g->n_landmark);
        memcpy(s->pts_bak, s->pts, sizeof(double) * 3 * (size_t)g->n_point);
    }
    for (int i = 0; i < g->n_pose; ++i) {
        double d[6];
        for (int k = 0; k < 6; ++k) d[k] = sign * dx[o->pose_off[i] + k];
        orc_pose_plus(s->pose + 7 * (size_t)i, d);
    }
    for (int i = 0; i < g->n_speedbias; ++i)
        for (int k = 0; k < 9; ++k) s->sb[9 * (size_t)i + k] += sign * dx[o->sb_off[i] + k];
    for (int l = 0; l < g->n_landmark; ++l) s->invd[l] += sign * dx[o->P + l];
    for (int l = 0; l < 3 * g->n_point; ++l) s->pts[l] += sign * dx[o->P + g->n_landmark + l]; /* Vertex::Plus */
    /* prior update: A17/src/backend/problem.cc:465-474 */
    if (flavour == VIO_LM_V17 && s->prior.dim > 0 && s->prior.err_dim > 0 && backup) {
        const int P = o->P, ed = s->prior.err_dim;
        memcpy(s->bprior_bak, s->bprior, sizeof(double) * P);
        memcpy(s->err_bak, s->err, sizeof(double) * ed);
        for (int r = 0; r < P; ++r) {
            double t = 0;
            for (int c = 0; c < P; ++c) t += s->prior.H[(size_t)r * P + c] * dx[c];
            s->bprior[r] -= t;
        }
        for (int r = 0; r < ed; ++r) {
            double t = 0;
            for (int c = 0; c < ed; ++c) t += s->prior.jt_inv[(size_t)r * ed + c] * s->bprior[c];
            s->err[r] = -t;
        }
    }
}

int orc_solve(const vio_graph *g0, const orc_prior *prior, int iterations, const vio_lm_opts *opts, double *pose,
              double *speedbias, double *inv_depth, double *b_prior_out, double *err_prior_out, orc_result *res) {
    return orc_solve_points(g0, prior, iterations, opts, pose, speedbias, inv_depth, NULL, b_prior_out, err_prior_out, res);
}
int orc_solve_points(const vio_graph *g0, const orc_prior *prior, int iterations, const vio_lm_opts *opts, double *pose,
                     double *speedbias, double *inv_depth, double *point_xyz, double *b_prior_out, double *err_prior_out,
                     orc_result *res) {
    const int flavour = opts ? opts->flavour : VIO_LM_V17;
    int solver = opts ? opts->solver : VIO_SOLVER_AUTO;
    const int fixed_it = opts ? opts->fixed_iterations : 0;
    if (solver == VIO_SOLVER_AUTO) solver = flavour == VIO_LM_V15 ? VIO_SOLVER_REF_PCG : VIO_SOLVER_DENSE_CHOL;
    const int v15 = flavour == VIO_LM_V15;
    ordering o;
    int rc = make_ordering(g0, &o);
    if (rc) return rc;
    const int P = o.P, M = o.M, n = P + M;
    lm_state s;
    memset(&s, 0, sizeof(s));
    s.g = *g0;
    size_t np = 7 * (size_t)g0->n_pose, ns = 9 * (size_t)g0->n_speedbias, nl = (size_t)g0->n_landmark, nx = 3 * (size_t)g0->n_point;
    s.pose = malloc(sizeof(double) * (np + 1)); s.pose_bak = malloc(sizeof(double) * (np + 1));
    s.sb = malloc(sizeof(double) * (ns + 1)); s.sb_bak = malloc(sizeof(double) * (ns + 1));
    s.invd = malloc(sizeof(double) * (nl + 1)); s.invd_bak = malloc(sizeof(double) * (nl + 1));
    memcpy(s.pose, g0->pose, sizeof(double) * np);
    if (ns) memcpy(s.sb, g0->speedbias, sizeof(double) * ns);
    if (nl) memcpy(s.invd, g0->inv_depth, sizeof(double) * nl);
    s.pts = malloc(sizeof(double) * (nx + 1)); s.pts_bak = malloc(sizeof(double) * (nx + 1));
    if (nx) memcpy(s.pts, g0->point_xyz, sizeof(double) * nx);
    s.g.pose = s.pose; s.g.speedbias = s.sb; s.g.inv_depth = s.invd; s.g.point_xyz = s.pts;
    if (prior && prior->dim > 0 && !v15) {
        s.prior = *prior;
        s.bprior = malloc(sizeof(double) * P); s.bprior_bak = malloc(sizeof(double) * P);
        memcpy(s.bprior, prior->b, sizeof(double) * P);
        s.prior.b = s.bprior;
        if (prior->err_dim > 0) {
            s.err = malloc(sizeof(double) * prior->err_dim); s.err_bak = malloc(sizeof(double) * prior->err_dim);
            memcpy(s.err, prior->err, sizeof(double) * prior->err_dim);
            s.prior.err = s.err;
        }
    }
    double *H = malloc(sizeof(double) * (size_t)n * n), *b = malloc(sizeof(double) * n), *dx = calloc(n, sizeof(double));
    if (res) memset(res, 0, sizeof(*res));
    double t0 = now_ms(), t_h = 0, th0;
#define MAKE_H()                                                  \
    do {                                                          \
        th0 = now_ms();                                           \
        orc_make_hessian(&s.g, &s.prior, flavour, H, b);          \
        t_h += now_ms() - th0;                                    \
        if (res) res->linearizations++;                           \
    } while (0)
    MAKE_H();
    /* ComputeLambdaInitLM — A15/backend/problem.cc:453-474, A17/src/backend/problem.cc:497-522 */
    double chi = 0, ni = 2.0, maxd = 0;
    orc_chi2(&s.g, &s.prior, flavour, &chi);
    for (int i = 0; i < n; ++i) maxd = fmax(maxd, fabs(H[(size_t)i * n + i]));
    if (!v15) maxd = fmin(5e10, maxd);
    double lambda = 1e-5 * maxd, stop_thr = 1e-6 * chi, last_chi = 1e20;
    if (res) { res->chi2_initial = chi; res->lambda_initial = lambda; }
    int stop = 0, iter = 0;
    while (!stop && iter < iterations) {
        if (res && iter < VIO_TRACE_MAX) { res->chi2_trace[iter] = chi; res->lambda_trace[iter] = lambda; }
        int ok = 0, false_cnt = 0;
        while (!ok && (v15 || false_cnt < 10)) {
            int64_t pit = 0;
            orc_solve_linear_blocks(H, b, P, g0->n_landmark, g0->n_point, lambda, solver, NULL, NULL, dx, &pit);
            if (res) { res->trial_steps++; res->pcg_iterations += pit; }
            double dx2 = 0, dot = 0;
            for (int i = 0; i < n; ++i) { dx2 += dx[i] * dx[i]; dot += dx[i] * (lambda * dx[i] + b[i]); }
            if (v15 && ((!fixed_it && dx2 <= 1e-6) || false_cnt > 10)) { stop = 1; break; }
            update_states(&s, &o, dx, flavour, 1.0, 1);
            /* IsGoodStepInLM — A15/backend/problem.cc:493-523, A17/src/backend/problem.cc:541-573 */
            double scale = v15 ? dot + 1e-3 : 0.5 * dot + 1e-6, temp_chi = 0;
            orc_chi2(&s.g, &s.prior, flavour, &temp_chi);
            double rho = (chi - temp_chi) / scale;
            if (rho > 0 && isfinite(temp_chi)) {
                double alpha = 1.0 - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2.0 / 3.0);
                lambda *= fmax(1.0 / 3.0, alpha);
                ni = 2; chi = temp_chi; ok = 1;
            } else { lambda *= ni; ni *= 2; ok = 0; }
            if (ok) { MAKE_H(); false_cnt = 0; }
            else {
                false_cnt++;
                if (v15) update_states(&s, &o, dx, flavour, -1.0, 0); /* Plus(-delta): A15/backend/problem.cc:439-450 */
                else {
                    memcpy(s.pose, s.pose_bak, sizeof(double) * np);
                    memcpy(s.sb, s.sb_bak, sizeof(double) * ns);
                    memcpy(s.invd, s.invd_bak, sizeof(double) * nl);
                    memcpy(s.pts, s.pts_bak, sizeof(double) * nx);
                    if (s.prior.dim > 0 && s.prior.err_dim > 0) {
                        memcpy(s.bprior, s.bprior_bak, sizeof(double) * P);
                        memcpy(s.err, s.err_bak, sizeof(double) * s.prior.err_dim);
                    }
                }
            }
        }
        iter++;
        if (!fixed_it) {
            if (v15) { if (sqrt(chi) <= stop_thr) stop = 1; }
            else if (last_chi - chi < 1e-5) stop = 1;
        }
        last_chi = chi;
    }
#undef MAKE_H
    if (res) {
        res->iterations = iter; res->chi2_final = chi; res->lambda_final = lambda;
        res->ms_total = now_ms() - t0; res->ms_hessian = t_h;
    }
    if (pose) memcpy(pose, s.pose, sizeof(double) * np);
    if (speedbias && ns) memcpy(speedbias, s.sb, sizeof(double) * ns);
    if (inv_depth && nl) memcpy(inv_depth, s.invd, sizeof(double) * nl);
    if (point_xyz && nx) memcpy(point_xyz, s.pts, sizeof(double) * nx);
    if (b_prior_out && s.bprior) memcpy(b_prior_out, s.bprior, sizeof(double) * P);
    if (err_prior_out && s.err) memcpy(err_prior_out, s.err, sizeof(double) * s.prior.err_dim);
    free(H); free(b); free(dx);
    free(s.pose); free(s.pose_bak); free(s.sb); free(s.sb_bak); free(s.invd); free(s.invd_bak); free(s.pts); free(s.pts_bak);
    free(s.bprior); free(s.bprior_bak); free(s.err); free(s.err_bak);
    free_ordering(&o);
    return VIO_OK;
}

/* ---- block-sparse accumulation for the large synthetic BA -------------------------------------------
 * Same per-edge arithmetic (orc_reproj + the MakeHessian block loop); the dense (P+M)^2 container of
 * A17/src/backend/problem.cc:306-307 is replaced by per-landmark scratch + 6x6 block storage, and the Schur
 * complement of :406-437 is applied landmark by landmark (Hmm is diagonal). */
static int find_block(const int32_t *rowptr, const int32_t *col, int a, int b) {
    int lo = rowptr[a], hi = rowptr[a + 1] - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (col[mid] < b) lo = mid + 1; else hi = mid; }
    return (lo <= rowptr[a + 1] - 1 && col[lo] == b) ? lo : -1;
}

typedef struct { int pose; double w[6]; } lm_slot;

static int linearize_range(const vio_graph *g, const int32_t *rowptr, const int32_t *col, double *val, double *bS,
                           double *Hll_out, double *bl_out, int64_t lm_begin, int64_t lm_end, double *checksum) {
    /* edges must be grouped by landmark in the caller's arrays (true for every generator in this repo) */
    double qic[4], tic[3], cs = 0;
    get_ext(g, qic, tic);
    const int base = 0;
    (void)base;
    int64_t e = 0;
    /* find first edge of lm_begin by scanning (arrays are landmark-sorted) */
    {
        int64_t lo = 0, hi = g->n_reproj;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (g->rp_landmark[mid] < lm_begin) lo = mid + 1; else hi = mid; }
        e = lo;
    }
    lm_slot slots[64];
    for (int64_t l = lm_begin; l < lm_end; ++l) {
        int ns = 0;
        double Hll = 0, bl = 0;
        double hostblk[36], hostb[6];
        memset(hostblk, 0, sizeof(hostblk)); memset(hostb, 0, sizeof(hostb));
        int host = -1;
        for (; e < g->n_reproj && g->rp_landmark[e] == l; ++e) {
            int i = g->rp_pose_i[e], j = g->rp_pose_j[e];
            double r[2], Jl[2], Ji[12], Jj[12];
            orc_reproj(g->inv_depth[l], g->pose + 7 * (size_t)i, g->pose + 7 * (size_t)j, qic, tic, g->rp_pts_i + 3 * e,
                       g->rp_pts_j + 2 * e, r, Jl, Ji, Jj);
            double W[4], drho, rho0;
            robust_info2(g->rp_loss, g->rp_loss_delta, g->rp_info, r, &drho, W, &rho0);
            const double c = g->rp_info;
            int ifix = g->pose_fixed && g->pose_fixed[i], jfix = g->pose_fixed && g->pose_fixed[j];
            /* JtW rows for each vertex */
            double WJl[2] = {W[0] * Jl[0] + W[1] * Jl[1], W[2] * Jl[0] + W[3] * Jl[1]};
            double WJi[12], WJj[12];
            for (int k = 0; k < 6; ++k) {
                WJi[k] = W[0] * Ji[k] + W[1] * Ji[6 + k]; WJi[6 + k] = W[2] * Ji[k] + W[3] * Ji[6 + k];
                WJj[k] = W[0] * Jj[k] + W[1] * Jj[6 + k]; WJj[6 + k] = W[2] * Jj[k] + W[3] * Jj[6 + k];
            }
            Hll += Jl[0] * WJl[0] + Jl[1] * WJl[1];
            bl -= drho * c * (Jl[0] * r[0] + Jl[1] * r[1]);
            if (host < 0) { host = i; slots[0].pose = i; memset(slots[0].w, 0, sizeof(slots[0].w)); ns = 1; }
            if (ns >= 64) return VIO_ERR_UNSUPPORTED;
            lm_slot *sj = &slots[ns++];
            sj->pose = j;
            for (int k = 0; k < 6; ++k) {
                slots[0].w[k] += ifix ? 0.0 : (Ji[k] * WJl[0] + Ji[6 + k] * WJl[1]);
                sj->w[k] = jfix ? 0.0 : (Jj[k] * WJl[0] + Jj[6 + k] * WJl[1]);
            }
            if (!ifix)
                for (int a = 0; a < 6; ++a) {
                    for (int b2 = 0; b2 < 6; ++b2) hostblk[6 * a + b2] += Ji[a] * WJi[b2] + Ji[6 + a] * WJi[6 + b2];
                    hostb[a] -= drho * c * (Ji[a] * r[0] + Ji[6 + a] * r[1]);
                }
            if (val) {
                if (!jfix) {
                    int id = find_block(rowptr, col, j, j);
                    for (int a = 0; a < 6; ++a) {
                        for (int b2 = 0; b2 < 6; ++b2) val[36 * (size_t)id + 6 * a + b2] += Jj[a] * WJj[b2] + Jj[6 + a] * WJj[6 + b2];
                        bS[6 * j + a] -= drho * c * (Jj[a] * r[0] + Jj[6 + a] * r[1]);
                    }
                }
                if (!ifix && !jfix) {
                    int id = find_block(rowptr, col, i, j), idt = find_block(rowptr, col, j, i);
                    for (int a = 0; a < 6; ++a)
                        for (int b2 = 0; b2 < 6; ++b2) {
                            double t = Ji[a] * WJj[b2] + Ji[6 + a] * WJj[6 + b2];
                            val[36 * (size_t)id + 6 * a + b2] += t;
                            val[36 * (size_t)idt + 6 * b2 + a] += t;
                        }
                }
            } else {
                cs += Jj[0] * WJj[0] + Ji[0] * WJj[3];
            }
        }
        if (ns == 0) continue;
        if (Hll_out) Hll_out[l] = Hll;
        if (bl_out) bl_out[l] = bl;
        const double inv = 1.0 / Hll;
        if (val) {
            int id = find_block(rowptr, col, host, host);
            for (int k = 0; k < 36; ++k) val[36 * (size_t)id + k] += hostblk[k];
            for (int k = 0; k < 6; ++k) bS[6 * host + k] += hostb[k];
            for (int a = 0; a < ns; ++a) {
                for (int k = 0; k < 6; ++k) bS[6 * slots[a].pose + k] -= slots[a].w[k] * inv * bl;
                for (int b2 = 0; b2 < ns; ++b2) {
                    int idb = find_block(rowptr, col, slots[a].pose, slots[b2].pose);
                    for (int r2 = 0; r2 < 6; ++r2)
                        for (int c2 = 0; c2 < 6; ++c2)
                            val[36 * (size_t)idb + 6 * r2 + c2] -= (slots[a].w[r2] * inv) * slots[b2].w[c2];
                }
            }
        } else {
            for (int a = 0; a < ns; ++a)
                for (int b2 = a; b2 < ns; ++b2)
                    for (int r2 = 0; r2 < 6; ++r2)
                        for (int c2 = 0; c2 < 6; ++c2) cs += (slots[a].w[r2] * inv) * slots[b2].w[c2];
            cs += hostblk[0] + hostb[0];
        }
    }
    if (checksum) *checksum = cs;
    return VIO_OK;
}

int orc_linearize_bsr(const vio_graph *g, const int32_t *rowptr, const int32_t *col, double *val, double *bS,
                      double *Hll, double *bl, int64_t lm_begin, int64_t lm_end) {
    if (g->n_speedbias != 0 || g->n_imu != 0) return VIO_ERR_UNSUPPORTED;
    /* SE3 priors (pose-only) */
    for (int k = 0; k < g->n_se3prior && lm_begin == 0; ++k) {
        int i = g->sp_pose[k];
        if (g->pose_fixed && g->pose_fixed[i]) continue;
        double r[6], J[36], JtW[36];
        orc_se3prior(g->pose + 7 * (size_t)i, g->sp_p + 3 * k, g->sp_q + 4 * k, r, J);
        const double *Om = g->sp_info + 36 * k;
        int id = find_block(rowptr, col, i, i);
        for (int a = 0; a < 6; ++a)
            for (int c = 0; c < 6; ++c) { double t = 0; for (int q = 0; q < 6; ++q) t += J[6 * q + a] * Om[6 * q + c]; JtW[6 * a + c] = t; }
        for (int a = 0; a < 6; ++a) {
            for (int c = 0; c < 6; ++c) { double t = 0; for (int q = 0; q < 6; ++q) t += JtW[6 * a + q] * J[6 * q + c]; val[36 * (size_t)id + 6 * a + c] += t; }
            double t = 0;
            for (int q = 0; q < 6; ++q) t += JtW[6 * a + q] * r[q];
            bS[6 * i + a] -= t;
        }
    }
    return linearize_range(g, rowptr, col, val, bS, Hll, bl, lm_begin, lm_end, NULL);
}

int orc_linearize_sample(const vio_graph *g, int64_t lm_begin, int64_t lm_end, double *checksum) {
    return linearize_range(g, NULL, NULL, NULL, NULL, NULL, NULL, lm_begin, lm_end, checksum);
}

/* ================================================================================================
 * IntegrationBase (A17/include/factor/integration_base.h)
 *   ctor :13-30  jacobian = I, covariance = 0, noise = diag(ACC_N^2, GYR_N^2, ACC_N^2, GYR_N^2, ACC_W^2, GYR_W^2) x I3
 *   midPointIntegration :55-130, propagate :132-158 (delta_q.normalize() after every sample)
 * ================================================================================================ */
int orc_preintegrate(int32_t n, const double *dt, const double *acc, const double *gyr, const double *ba, const double *bg,
                     const double *noise, double *sum_dt, double *delta_p, double *delta_q_xyzw, double *delta_v,
                     double *jac225, double *cov225) {
    double J[225], P[225], F[225], V[270], T[225], Q[18];
    double p[3] = {0, 0, 0}, v[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, a0[3], g0[3], sdt = 0.0;
    if (n < 1) return VIO_ERR_INVALID;
    for (int i = 0; i < 225; ++i) { J[i] = (i / 15 == i % 15) ? 1.0 : 0.0; P[i] = 0.0; }
    for (int k = 0; k < 3; ++k) {
        Q[k] = noise[0] * noise[0]; Q[3 + k] = noise[2] * noise[2]; Q[6 + k] = noise[0] * noise[0];
        Q[9 + k] = noise[2] * noise[2]; Q[12 + k] = noise[1] * noise[1]; Q[15 + k] = noise[3] * noise[3];
        a0[k] = acc[k]; g0[k] = gyr[k];
    }
    for (int i = 1; i < n; ++i) {
        const double h = dt[i];
        const double *a1 = acc + 3 * i, *g1 = gyr + 3 * i;
        double a0x[3], a1x[3], w[3], ua0[3], ua1[3], q1[4], dqh[4], R0[9], R1[9], A0[9], A1[9], W[9], M0[9], M1[9], IW[9], M1w[9];
        for (int k = 0; k < 3; ++k) { a0x[k] = a0[k] - ba[k]; a1x[k] = a1[k] - ba[k]; w[k] = 0.5 * (g0[k] + g1[k]) - bg[k]; }
        q_rot(q, a0x, ua0);
        dqh[0] = w[0] * h / 2; dqh[1] = w[1] * h / 2; dqh[2] = w[2] * h / 2; dqh[3] = 1.0;
        q_mul(q, dqh, q1);
        q_rot(q1, a1x, ua1);
        q_toR(q, R0); q_toR(q1, R1);
        hat3(a0x, A0); hat3(a1x, A1); hat3(w, W);
        m3_mul(R0, A0, M0); m3_mul(R1, A1, M1);
        for (int k = 0; k < 9; ++k) IW[k] = ((k % 4 == 0) ? 1.0 : 0.0) - W[k] * h;
        m3_mul(M1, IW, M1w);
        memset(F, 0, sizeof(F)); memset(V, 0, sizeof(V));
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const double id = r == c ? 1.0 : 0.0, m0 = M0[3 * r + c], m1 = M1[3 * r + c], m1w = M1w[3 * r + c];
                const double r0 = R0[3 * r + c], r1 = R1[3 * r + c];
#define FB(br, bc) F[(3 * (br) + r) * 15 + 3 * (bc) + c]
#define VB(br, bc) V[(3 * (br) + r) * 18 + 3 * (bc) + c]
                FB(0, 0) = id; FB(0, 1) = -0.25 * m0 * h * h + -0.25 * m1w * h * h; FB(0, 2) = id * h;
                FB(0, 3) = -0.25 * (r0 + r1) * h * h; FB(0, 4) = -0.25 * m1 * h * h * -h;
                FB(1, 1) = IW[3 * r + c]; FB(1, 4) = -1.0 * id * h;
                FB(2, 1) = -0.5 * m0 * h + -0.5 * m1w * h; FB(2, 2) = id; FB(2, 3) = -0.5 * (r0 + r1) * h;
                FB(2, 4) = -0.5 * m1 * h * -h; FB(3, 3) = id; FB(4, 4) = id;
                VB(0, 0) = 0.25 * r0 * h * h; VB(0, 1) = 0.25 * -m1 * h * h * 0.5 * h; VB(0, 2) = 0.25 * r1 * h * h;
                VB(0, 3) = VB(0, 1); VB(1, 1) = 0.5 * id * h; VB(1, 3) = 0.5 * id * h;
                VB(2, 0) = 0.5 * r0 * h; VB(2, 1) = 0.5 * -m1 * h * 0.5 * h; VB(2, 2) = 0.5 * r1 * h; VB(2, 3) = VB(2, 1);
                VB(3, 4) = id * h; VB(4, 5) = id * h;
#undef FB
#undef VB
            }
        /* jacobian = F jacobian ; covariance = F covariance F^T + V noise V^T */
        for (int r = 0; r < 15; ++r)
            for (int c = 0; c < 15; ++c) {
                double a = 0.0, b = 0.0;
                for (int k = 0; k < 15; ++k) { a += F[15 * r + k] * J[15 * k + c]; b += F[15 * r + k] * P[15 * k + c]; }
                T[15 * r + c] = b;
                cov225[15 * r + c] = a; /* scratch: new jacobian */
            }
        memcpy(J, cov225, sizeof(J));
        for (int r = 0; r < 15; ++r)
            for (int c = 0; c < 15; ++c) {
                double a = 0.0;
                for (int k = 0; k < 15; ++k) a += T[15 * r + k] * F[15 * c + k];
                for (int k = 0; k < 18; ++k) a += V[18 * r + k] * Q[k] * V[18 * c + k];
                P[15 * r + c] = a;
            }
        for (int k = 0; k < 3; ++k) {
            const double ua = 0.5 * (ua0[k] + ua1[k]);
            p[k] = p[k] + v[k] * h + 0.5 * ua * h * h;
            v[k] = v[k] + ua * h;
        }
        {
            const double nq = sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
            for (int k = 0; k < 4; ++k) q[k] = q1[k] / nq;
        }
        sdt += h;
        for (int k = 0; k < 3; ++k) { a0[k] = a1[k]; g0[k] = g1[k]; }
    }
    *sum_dt = sdt;
    for (int k = 0; k < 3; ++k) { delta_p[k] = p[k]; delta_v[k] = v[k]; }
    for (int k = 0; k < 4; ++k) delta_q_xyzw[k] = q[k];
    memcpy(jac225, J, sizeof(J));
    memcpy(cov225, P, sizeof(P));
    return VIO_OK;
}

/* ================================================================================================
 * Problem::Marginalize(margVertexs = {pose[marg_pose], speedbias[marg_sb]}, pose_dim)
 * A17/src/backend/problem.cc:617-795.  Eigen's SelfAdjointEigenSolver is restated as a cyclic Jacobi eigen-solver
 * (ascending eigenvalues, eigenvector signs arbitrary - as with Eigen); see DESIGN 6 for what is and is not
 * reproducible across backward-stable eigen-solvers.
 * ================================================================================================ */
static void jacobi_eigh(int n, double *A /* in: symmetric, destroyed */, double *w /* n ascending */, double *V /* n x n, columns */) {
    for (int i = 0; i < n * n; ++i) V[i] = 0.0;
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < n; ++p) { diag += A[p * n + p] * A[p * n + p]; for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q]; }
        if (off <= 1e-32 * (diag + off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
    for (int i = 0; i < n - 1; ++i) {  /* selection sort, ascending, columns of V follow */
        int m = i;
        for (int j = i + 1; j < n; ++j) if (w[j] < w[m]) m = j;
        if (m != i) {
            double t = w[i]; w[i] = w[m]; w[m] = t;
            for (int k = 0; k < n; ++k) { t = V[k * n + i]; V[k * n + i] = V[k * n + m]; V[k * n + m] = t; }
        }
    }
}

int orc_marginalize(const vio_graph *g, const orc_prior *prior, int32_t marg_pose, int32_t marg_sb, int32_t *dim_out,
                    double *H_out, double *b_out, double *err_out, double *jt_inv_out) {
    ordering o;
    int rc = make_ordering(g, &o);
    if (rc) return rc;
    const int P = o.P; /* pose_dim */
    if (marg_pose < 0 || marg_pose >= g->n_pose || marg_sb < 0 || marg_sb >= g->n_speedbias || g->n_point > 0) { free_ordering(&o); return VIO_ERR_INVALID; }
    double qic[4], tic[3];
    get_ext(g, qic, tic);
    /* landmarks of the frame's edges get ordering ids pose_dim, pose_dim+1, ... (:626-639; here in landmark-index order) */
    int *lm_slot = (int *)malloc(sizeof(int) * (g->n_landmark > 0 ? g->n_landmark : 1));
    for (int l = 0; l < g->n_landmark; ++l) lm_slot[l] = -1;
    int nl = 0;
    for (int64_t e = 0; e < g->n_reproj; ++e)
        if (g->rp_pose_i[e] == marg_pose || g->rp_pose_j[e] == marg_pose) lm_slot[g->rp_landmark[e]] = 0;
    for (int l = 0; l < g->n_landmark; ++l) if (lm_slot[l] == 0) lm_slot[l] = nl++;
    const int cols = P + nl;
    double *H = (double *)calloc((size_t)cols * cols, sizeof(double)), *b = (double *)calloc(cols, sizeof(double));
    /* H_marg, b_marg over the frame's edges: NO fixed-vertex test here, the extrinsic vertex contributes too (:645-680) */
    for (int64_t e = 0; e < g->n_reproj; ++e) {
        const int l = g->rp_landmark[e], i = g->rp_pose_i[e], j = g->rp_pose_j[e];
        if (i != marg_pose && j != marg_pose) continue;
        double r[2], Jl[2], Ji[12], Jj[12], Jex[12], W[4], Om[4] = {g->rp_info, 0, 0, g->rp_info}, drho = 1.0, rho0;
        orc_reproj(g->inv_depth[l], g->pose + 7 * (size_t)i, g->pose + 7 * (size_t)j, qic, tic, g->rp_pts_i + 3 * e, g->rp_pts_j + 2 * e, r, Jl, Ji, Jj);
        robust_info2(g->rp_loss, g->rp_loss_delta, g->rp_info, r, &drho, W, &rho0);
        const double *Jv[4] = {Jl, Ji, Jj, Jex};
        int ldj[4] = {1, 6, 6, 6}, dim[4] = {1, 6, 6, 6}, off[4] = {P + lm_slot[l], o.pose_off[i], o.pose_off[j], -1};
        int nv = 3;
        if (g->ext_pose >= 0) {
            orc_reproj_jext(g->inv_depth[l], g->pose + 7 * (size_t)i, g->pose + 7 * (size_t)j, qic, tic, g->rp_pts_i + 3 * e, Jex);
            off[3] = o.pose_off[g->ext_pose];
            nv = 4;
        }
        add_edge_dense(H, b, cols, 2, nv, Jv, ldj, dim, off, W, Om, drho, r);
    }
    for (int k = 0; k < g->n_imu; ++k) {
        const int pi = g->imu_pose_i[k], si = g->imu_sb_i[k], pj = g->imu_pose_j[k], sj = g->imu_sb_j[k];
        if (pi != marg_pose && pj != marg_pose) continue;
        double r[15], J[450], info[225];
        orc_imu(g->pose + 7 * (size_t)pi, g->speedbias + 9 * (size_t)si, g->pose + 7 * (size_t)pj, g->speedbias + 9 * (size_t)sj,
                g->imu_sum_dt[k], g->imu_delta_p + 3 * k, g->imu_delta_q + 4 * k, g->imu_delta_v + 3 * k, g->imu_lin_ba + 3 * k,
                g->imu_lin_bg + 3 * k, g->imu_jacobian + 225 * k, g->gravity, r, J);
        mat_inverse(15, g->imu_covariance + 225 * k, info);
        const double *Jv[4] = {J, J + 6, J + 15, J + 21};
        int ldj[4] = {30, 30, 30, 30}, dim[4] = {6, 9, 6, 9}, off[4] = {o.pose_off[pi], o.sb_off[si], o.pose_off[pj], o.sb_off[sj]};
        add_edge_dense(H, b, cols, 15, 4, Jv, ldj, dim, off, info, info, 1.0, r);
    }
    for (int k = 0; k < g->n_se3prior; ++k) {
        const int i = g->sp_pose[k];
        if (i != marg_pose) continue;
        double r[6], J[36];
        orc_se3prior(g->pose + 7 * (size_t)i, g->sp_p + 3 * k, g->sp_q + 4 * k, r, J);
        const double *Jv[1] = {J};
        int ldj[1] = {6}, dim[1] = {6}, off[1] = {o.pose_off[i]};
        add_edge_dense(H, b, cols, 6, 1, Jv, ldj, dim, off, g->sp_info + 36 * k, g->sp_info + 36 * k, 1.0, r);
    }
    /* marg landmarks (:686-708) */
    double *Hm = (double *)calloc((size_t)P * P, sizeof(double)), *bm = (double *)calloc(P, sizeof(double));
    for (int r = 0; r < P; ++r) {
        for (int c = 0; c < P; ++c) {
            double t = 0.0;
            for (int l = 0; l < nl; ++l) t += (H[(size_t)r * cols + P + l] / H[(size_t)(P + l) * cols + P + l]) * H[(size_t)(P + l) * cols + c];
            Hm[(size_t)r * P + c] = H[(size_t)r * cols + c] - t;
        }
        double t = 0.0;
        for (int l = 0; l < nl; ++l) t += (H[(size_t)r * cols + P + l] / H[(size_t)(P + l) * cols + P + l]) * b[P + l];
        bm[r] = b[r] - t;
    }
    /* + prior (:710-715) */
    if (prior && prior->dim > 0) {
        if (prior->dim != P) { free(H); free(b); free(Hm); free(bm); free(lm_slot); free_ordering(&o); return VIO_ERR_INVALID; }
        for (int i = 0; i < P * P; ++i) Hm[i] += prior->H[i];
        for (int i = 0; i < P; ++i) bm[i] += prior->b[i];
    }
    /* move the marginalised blocks to the bottom right, larger index first (:720-745): a stable permutation */
    int *perm = (int *)malloc(sizeof(int) * P), n_keep = 0;
    const int off_p = o.pose_off[marg_pose], off_s = o.sb_off[marg_sb];
    int first_off = off_p < off_s ? off_p : off_s, first_dim = off_p < off_s ? 6 : 9;
    int second_off = off_p < off_s ? off_s : off_p, second_dim = off_p < off_s ? 9 : 6;
    /* the loop moves margVertexs[1] (speed-bias) first, then margVertexs[0] (pose): final tail order = [speed-bias | pose] */
    for (int r = 0; r < P; ++r) {
        const int in_first = r >= first_off && r < first_off + first_dim, in_second = r >= second_off && r < second_off + second_dim;
        if (!in_first && !in_second) perm[n_keep++] = r;
    }
    const int m2 = 15, n2 = P - m2;
    for (int d = 0; d < 9; ++d) perm[n2 + d] = off_s + d;
    for (int d = 0; d < 6; ++d) perm[n2 + 9 + d] = off_p + d;
    double *Hp = (double *)malloc(sizeof(double) * (size_t)P * P), *bp = (double *)malloc(sizeof(double) * P);
    for (int r = 0; r < P; ++r) { bp[r] = bm[perm[r]]; for (int c = 0; c < P; ++c) Hp[(size_t)r * P + c] = Hm[(size_t)perm[r] * P + perm[c]]; }
    /* Amm pseudo-inverse and Schur (:747-766) */
    const double eps = 1e-8;
    double Amm[225], wv[15], Vv[225], Ainv[225];
    for (int r = 0; r < m2; ++r)
        for (int c = 0; c < m2; ++c) Amm[r * m2 + c] = 0.5 * (Hp[(size_t)(n2 + r) * P + n2 + c] + Hp[(size_t)(n2 + c) * P + n2 + r]);
    jacobi_eigh(m2, Amm, wv, Vv);
    for (int r = 0; r < m2; ++r)
        for (int c = 0; c < m2; ++c) {
            double t = 0.0;
            for (int k = 0; k < m2; ++k) t += Vv[r * m2 + k] * (wv[k] > eps ? 1.0 / wv[k] : 0.0) * Vv[c * m2 + k];
            Ainv[r * m2 + c] = t;
        }
    double *Hpr = (double *)malloc(sizeof(double) * (size_t)n2 * n2), *bpr = (double *)malloc(sizeof(double) * n2);
    double *tB = (double *)malloc(sizeof(double) * (size_t)n2 * m2);
    for (int r = 0; r < n2; ++r)
        for (int c = 0; c < m2; ++c) {
            double t = 0.0;
            for (int k = 0; k < m2; ++k) t += Hp[(size_t)r * P + n2 + k] * Ainv[k * m2 + c];
            tB[(size_t)r * m2 + c] = t;
        }
    for (int r = 0; r < n2; ++r) {
        for (int c = 0; c < n2; ++c) {
            double t = 0.0;
            for (int k = 0; k < m2; ++k) t += tB[(size_t)r * m2 + k] * Hp[(size_t)(n2 + k) * P + c];
            Hpr[(size_t)r * n2 + c] = Hp[(size_t)r * P + c] - t;
        }
        double t = 0.0;
        for (int k = 0; k < m2; ++k) t += tB[(size_t)r * m2 + k] * bp[n2 + k];
        bpr[r] = bp[r] - t;
    }
    /* eigen-decomposition of H_prior: Jt_prior_inv, err_prior, H_prior = J^T J, |.| < 1e-9 -> 0 (:768-782) */
    double *Acopy = (double *)malloc(sizeof(double) * (size_t)n2 * n2), *w2 = (double *)malloc(sizeof(double) * n2);
    double *V2 = (double *)malloc(sizeof(double) * (size_t)n2 * n2);
    memcpy(Acopy, Hpr, sizeof(double) * (size_t)n2 * n2);
    jacobi_eigh(n2, Acopy, w2, V2);
    for (int r = 0; r < n2; ++r) {
        const double sis = w2[r] > eps ? sqrt(1.0 / w2[r]) : 0.0;
        for (int c = 0; c < n2; ++c) jt_inv_out[(size_t)r * n2 + c] = sis * V2[(size_t)c * n2 + r];
    }
    for (int r = 0; r < n2; ++r) {
        double t = 0.0;
        for (int c = 0; c < n2; ++c) t += jt_inv_out[(size_t)r * n2 + c] * bpr[c];
        err_out[r] = -t;
    }
    for (int r = 0; r < n2; ++r)
        for (int c = 0; c < n2; ++c) {
            double t = 0.0;
            for (int k = 0; k < n2; ++k) t += V2[(size_t)r * n2 + k] * (w2[k] > eps ? w2[k] : 0.0) * V2[(size_t)c * n2 + k];
            H_out[(size_t)r * n2 + c] = fabs(t) > 1e-9 ? t : 0.0;
        }
    memcpy(b_out, bpr, sizeof(double) * n2);
    *dim_out = n2;
    free(H); free(b); free(Hm); free(bm); free(lm_slot); free(perm); free(Hp); free(bp); free(Hpr); free(bpr); free(tB);
    free(Acopy); free(w2); free(V2);
    free_ordering(&o);
    return VIO_OK;
}
