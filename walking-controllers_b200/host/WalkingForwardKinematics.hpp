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
    std::vector<double> m_q, m_dq, m_lfH, m_rfH, m_neckR, m_com, m_Jl, m_Jr, m_Jn, m_Jc;
    // evaluateCoM / evaluateDCM (WalkingForwardKinematics.cpp:278-340): CoM velocity, low-pass filters, measured DCM
    wc_glue* m_glue{nullptr};
    double m_cutFrequency{0.0}, m_comOut[3]{}, m_comVel[3]{}, m_dcmOut[2]{};
    bool m_useFilters{false}, m_comEvaluated{false}, m_dcmEvaluated{false};
    bool evaluate();

public:
    ~WalkingFK();
    bool initialize(const yarp::os::Searchable& config, const wc_model& model);
    // WalkingForwardKinematics.cpp:217-256
    bool evaluateWorldToBaseTransformation(const iDynTree::Transform& leftFootTransform,
                                           const iDynTree::Transform& rightFootTransform, const bool& isLeftFixedFrame);
    // :258-276 (the joint velocities enter the CoM velocity, hence the measured DCM)
    bool setInternalRobotState(const iDynTree::VectorDynSize& positionFeedbackInRadians,
                               const iDynTree::VectorDynSize& velocityFeedbackInRadians);
    iDynTree::Transform getLeftFootToWorldTransform();
    iDynTree::Transform getRightFootToWorldTransform();
    iDynTree::Rotation getNeckOrientation();
    bool getCoMPosition(iDynTree::Position& comPosition); // filtered if use_filters and evaluateCoM() was called
    bool evaluateCoM();                                      // :278-310
    bool evaluateDCM();                                      // :312-340
    bool getDCM(iDynTree::Vector2& dcm);                     // :342-352
    bool getCoMVelocity(iDynTree::Vector3& comVelocity);     // :376-394
    bool getLeftFootJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getRightFootJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getNeckJacobian(iDynTree::MatrixDynSize& jacobian);
    bool getCoMJacobian(iDynTree::MatrixDynSize& jacobian);
};

#endif
