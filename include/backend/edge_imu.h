// EdgeImu(IntegrationBase*) for VINS drivers: include after the integrator's own factor/integration_base.h
// (it only reads the public constants the reference's EdgeImu reads:
//  17-vins-initialization/vins-mono/src/backend/edge_imu.cc:13-65).
#pragma once
#include "myslam_backend_b200.h"
namespace myslam {
namespace backend {
template <typename IntegrationBaseT>
inline ImuPreintegrationB200 MakeImuPreintegrationB200(const IntegrationBaseT *p) {
    ImuPreintegrationB200 q;
    q.sum_dt = p->sum_dt; q.delta_p = p->delta_p; q.delta_q = p->delta_q; q.delta_v = p->delta_v;
    q.linearized_ba = p->linearized_ba; q.linearized_bg = p->linearized_bg;
    q.jacobian = p->jacobian; q.covariance = p->covariance;
    return q;
}
#ifdef MYSLAM_B200_HAVE_INTEGRATION_BASE
class EdgeImu : public EdgeImuB200 {
public:
    explicit EdgeImu(IntegrationBase *pre) : EdgeImuB200(MakeImuPreintegrationB200(pre)) {}
};
#endif
}  // namespace backend
}  // namespace myslam
