// vio_preint.cuh — IMU pre-integration on the device: IntegrationBase::push_back / propagate /
// repropagate (A17/include/factor/integration_base.h:30-158) for a BATCH of segments (one per keyframe pair).
// One CTA per segment; the sample loop is sequential (each step needs the previous delta_q), the 15x15 products
// jacobian = F jacobian and covariance = F covariance F^T + V noise V^T are spread over 225 threads with the matrices in
// shared memory.  Outputs are laid out like the EdgeImu constants of vio_graph, so they can be passed on unchanged.
#pragma once
#include "vio_math.cuh"

struct PreintView {
    int n_seg;
    const int *seg_ptr;                  // [n_seg + 1] sample ranges; the first sample of a segment is (acc_0, gyr_0)
    const double *dt, *acc, *gyr;        // per sample: dt, acc[3], gyr[3]
    const double *ba, *bg;               // per segment linearisation biases
    double acc_n, acc_w, gyr_n, gyr_w;   // ACC_N, ACC_W, GYR_N, GYR_W (A17/include/parameters.h)
    double *sum_dt, *dp, *dq, *dv, *jac, *cov;
};

// Eigen's QuaternionBase::_transformVector (EIG/Eigen/src/Geometry/Quaternion.h:470-483): v + w*2(u x v) + u x 2(u x v).
// It is NOT a pure rotation for the un-normalised result_delta_q the reference feeds it.
VIO_HD void quat_transform(const double q[4] /*xyzw*/, const double v[3], double out[3]) {
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

__global__ void __launch_bounds__(256) k_preintegrate(PreintView s) {
    __shared__ double J[225], P[225], F[225], T[225], V[15 * 18];
    __shared__ double st[32];  // 0..2 dp, 3..6 dq (xyzw), 7..9 dv, 10..12 acc_0, 13..15 gyr_0, 16 sum_dt
    __shared__ double R0[9], R1[9], Ra0[9], Ra1[9], Rw[9], dts;
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int i0 = s.seg_ptr[seg], i1 = s.seg_ptr[seg + 1];
    const double ba[3] = {s.ba[3 * seg], s.ba[3 * seg + 1], s.ba[3 * seg + 2]};
    const double bg[3] = {s.bg[3 * seg], s.bg[3 * seg + 1], s.bg[3 * seg + 2]};
    if (tid < 225) { J[tid] = (tid / 15 == tid % 15) ? 1.0 : 0.0; P[tid] = 0.0; }
    if (tid == 0) {
        for (int k = 0; k < 17; ++k) st[k] = 0.0;
        st[6] = 1.0;
        if (i1 > i0)
            for (int k = 0; k < 3; ++k) { st[10 + k] = s.acc[3 * (size_t)i0 + k]; st[13 + k] = s.gyr[3 * (size_t)i0 + k]; }
    }
    __syncthreads();
    const double q_an = s.acc_n * s.acc_n, q_gn = s.gyr_n * s.gyr_n, q_aw = s.acc_w * s.acc_w, q_gw = s.gyr_w * s.gyr_w;
    for (int i = i0 + 1; i < i1; ++i) {
        // ---- midPointIntegration: state by one thread, F and V blocks from the shared 3x3 pieces -------------------------
        if (tid == 0) {
            const double dt = s.dt[i];
            const double a1[3] = {s.acc[3 * (size_t)i], s.acc[3 * (size_t)i + 1], s.acc[3 * (size_t)i + 2]};
            const double g1[3] = {s.gyr[3 * (size_t)i], s.gyr[3 * (size_t)i + 1], s.gyr[3 * (size_t)i + 2]};
            const double a0x[3] = {st[10] - ba[0], st[11] - ba[1], st[12] - ba[2]};
            const double a1x[3] = {a1[0] - ba[0], a1[1] - ba[1], a1[2] - ba[2]};
            const double w[3] = {0.5 * (st[13] + g1[0]) - bg[0], 0.5 * (st[14] + g1[1]) - bg[1], 0.5 * (st[15] + g1[2]) - bg[2]};
            const double q0[4] = {st[3], st[4], st[5], st[6]};
            double un_acc_0[3], un_acc_1[3], q1[4];
            quat_transform(q0, a0x, un_acc_0);
            const double dqh[4] = {w[0] * dt / 2, w[1] * dt / 2, w[2] * dt / 2, 1.0};
            quat_mul(q0, dqh, q1);
            quat_transform(q1, a1x, un_acc_1);
            quat_to_R(q0, R0);
            quat_to_R(q1, R1);
            const double hat0[9] = {0, -a0x[2], a0x[1], a0x[2], 0, -a0x[0], -a0x[1], a0x[0], 0};
            const double hat1[9] = {0, -a1x[2], a1x[1], a1x[2], 0, -a1x[0], -a1x[1], a1x[0], 0};
            const double hatw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
            for (int k = 0; k < 9; ++k) { Ra0[k] = hat0[k]; Ra1[k] = hat1[k]; Rw[k] = hatw[k]; }
            for (int k = 0; k < 3; ++k) {
                const double un_acc = 0.5 * (un_acc_0[k] + un_acc_1[k]);
                st[k] = st[k] + st[7 + k] * dt + 0.5 * un_acc * dt * dt;
                st[7 + k] = st[7 + k] + un_acc * dt;
            }
            // delta_q = result_delta_q.normalized() (Eigen: coeffs / norm)
            const double nq = sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
            st[3] = q1[0] / nq; st[4] = q1[1] / nq; st[5] = q1[2] / nq; st[6] = q1[3] / nq;
            st[16] += dt;
            for (int k = 0; k < 3; ++k) { st[10 + k] = a1[k]; st[13 + k] = g1[k]; }
            dts = dt;
        }
        if (tid < 225) F[tid] = 0.0;
        for (int t = tid; t < 270; t += blockDim.x) V[t] = 0.0;
        __syncthreads();
        if (tid < 9) {
            // every thread owns element (r, c) of the 3x3 blocks
            const int r = tid / 3, c = tid % 3;
            const double dt = dts, id = (r == c) ? 1.0 : 0.0;
            // M0 = R0 * hat(a0),  M1 = R1 * hat(a1),  M1w = R1 * hat(a1) * (I - hat(w) dt)
            double m0 = 0.0, m1 = 0.0, m1w = 0.0;
            for (int k = 0; k < 3; ++k) { m0 += R0[3 * r + k] * Ra0[3 * k + c]; m1 += R1[3 * r + k] * Ra1[3 * k + c]; }
            for (int k = 0; k < 3; ++k) {
                double m1rk = 0.0;
                for (int q = 0; q < 3; ++q) m1rk += R1[3 * r + q] * Ra1[3 * q + k];
                m1w += m1rk * ((k == c ? 1.0 : 0.0) - Rw[3 * k + c] * dt);
            }
            const double r0 = R0[3 * r + c], r1 = R1[3 * r + c];
#define FB(br, bc) F[(3 * (br) + r) * 15 + 3 * (bc) + c]
#define VB(br, bc) V[(3 * (br) + r) * 18 + 3 * (bc) + c]
            FB(0, 0) = id;
            FB(0, 1) = -0.25 * m0 * dt * dt + -0.25 * m1w * dt * dt;
            FB(0, 2) = id * dt;
            FB(0, 3) = -0.25 * (r0 + r1) * dt * dt;
            FB(0, 4) = -0.25 * m1 * dt * dt * -dt;
            FB(1, 1) = id - Rw[3 * r + c] * dt;
            FB(1, 4) = -1.0 * id * dt;
            FB(2, 1) = -0.5 * m0 * dt + -0.5 * m1w * dt;
            FB(2, 2) = id;
            FB(2, 3) = -0.5 * (r0 + r1) * dt;
            FB(2, 4) = -0.5 * m1 * dt * -dt;
            FB(3, 3) = id;
            FB(4, 4) = id;
            VB(0, 0) = 0.25 * r0 * dt * dt;
            VB(0, 1) = 0.25 * -m1 * dt * dt * 0.5 * dt;
            VB(0, 2) = 0.25 * r1 * dt * dt;
            VB(0, 3) = VB(0, 1);
            VB(1, 1) = 0.5 * id * dt;
            VB(1, 3) = 0.5 * id * dt;
            VB(2, 0) = 0.5 * r0 * dt;
            VB(2, 1) = 0.5 * -m1 * dt * 0.5 * dt;
            VB(2, 2) = 0.5 * r1 * dt;
            VB(2, 3) = VB(2, 1);
            VB(3, 4) = id * dt;
            VB(4, 5) = id * dt;
#undef FB
#undef VB
        }
        __syncthreads();
        // ---- jacobian = F jacobian ; covariance = F covariance F^T + V noise V^T -----------------------------------------
        double jn = 0.0, tn = 0.0;
        const int r = tid / 15, c = tid % 15;
        if (tid < 225) {
            for (int k = 0; k < 15; ++k) { jn += F[15 * r + k] * J[15 * k + c]; tn += F[15 * r + k] * P[15 * k + c]; }
        }
        __syncthreads();
        if (tid < 225) { J[tid] = jn; T[tid] = tn; }
        __syncthreads();
        if (tid < 225) {
            double pn = 0.0;
            for (int k = 0; k < 15; ++k) pn += T[15 * r + k] * F[15 * c + k];
            double vn = 0.0;
            for (int k = 0; k < 18; ++k) {
                const double qk = k < 3 ? q_an : k < 6 ? q_gn : k < 9 ? q_an : k < 12 ? q_gn : k < 15 ? q_aw : q_gw;
                vn += V[18 * r + k] * qk * V[18 * c + k];
            }
            P[tid] = pn + vn;
        }
        __syncthreads();
    }
    if (tid < 225) { s.jac[225 * (size_t)seg + tid] = J[tid]; s.cov[225 * (size_t)seg + tid] = P[tid]; }
    if (tid == 0) {
        s.sum_dt[seg] = st[16];
        for (int k = 0; k < 3; ++k) { s.dp[3 * (size_t)seg + k] = st[k]; s.dv[3 * (size_t)seg + k] = st[7 + k]; }
        for (int k = 0; k < 4; ++k) s.dq[4 * (size_t)seg + k] = st[3 + k];
    }
}
