// Drop-in for the hot-path subset of the reference's WalkingFK
// (modules/Walking_module/include/WalkingForwardKinematics.hpp:73-220): base selection on the stance foot,
// setInternalRobotState, frame transforms and MIXED-representation Jacobians, computed by the CUDA FK
// kernel (wc_fk_host) on a wc_model table (the iDynTree::Model / URDF is not part of the reference repo).
#ifndef WC_HOST_WALKING_FK_HPP
#define WC_HOST_WALKING_FK_HPP

#include <vector>

#include "../../include/wc_capi.h"
#include "wc_compat.hpp"

class WalkingFK {
    wc_ik* m_handle{nullptr};
    wc_model m_model{};
    iDynTree::Transform m_worldToBaseSole; // world_H_sole of the current stance foot
    bool m_prevContactLeft{false}, m_firstStep{true}, m_evaluated{false};
    int m_nJ{0};
    std::vector<double> m_q, m_lfH, m_rfH, m_neckR, m_com, m_Jl, m_Jr, m_Jn, m_Jc;
    bool evaluate();

public:
    ~WalkingFK();
    bool initialize(const yarp::os::Searchable& config, const wc_model& model);
    // WalkingForwardKinematics.cpp:217-256
    bool evaluateWorldToBaseTransformation(const iDynTree::Transform& leftFootTransform,
                                           const iDynTree::Transform& rightFootTransform, const bool& isLeftFixedFrame);
    // :258-276 (joint velocities only enter the CoM velocity, which is not on the hot path)
    bool setInternalRobotState(const iDynTree::VectorDynSize& positionFeedbackInRadians,
                               const iDynTree::VectorDynSize& velocityFeedbackInRadians);
    iDynTree::Transform getLeftFootToWorldTransform();
    iDynTree::Transform getRightFootToWorldTransform();
    iDynTree::Rotation getNeckOrientation();
    bool getCoMPosition(iDynTree::Position& comPosition);
    bool getLeftFootJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getRightFootJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getNeckJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getCoMJacobian(iDynTree::MatrixDynSize& jacobian);
};

#endif
