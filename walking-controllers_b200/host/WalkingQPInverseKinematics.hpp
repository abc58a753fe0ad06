// Drop-in for the reference's WalkingQPIK / WalkingQPIK_qpOASES
// (modules/Walking_module/include/WalkingQPInverseKinematics.hpp:25-209,
//  modules/Walking_module/include/WalkingQPInverseKinematics_qpOASES.hpp:15-113): same virtual interface,
// setters with the same dimension checks and bool conventions, getSolution latch.  The QP is built and
// solved by the CUDA library (include/wc_capi.h, wc_ik_solve_with_jacobians_host) with a batch of one.
#ifndef WC_HOST_WALKING_QPIK_HPP
#define WC_HOST_WALKING_QPIK_HPP

#include <vector>

#include "../../include/wc_capi.h"
#include "wc_compat.hpp"

class WalkingQPIK {
protected:
    iDynTree::MatrixDynSize m_comJacobian, m_neckJacobian, m_leftFootJacobian, m_rightFootJacobian;
    iDynTree::Twist m_leftFootTwist, m_rightFootTwist;
    iDynTree::Vector3 m_comVelocity;
    iDynTree::Position m_desiredComPosition, m_comPosition;
    iDynTree::Transform m_desiredLeftFootToWorldTransform, m_desiredRightFootToWorldTransform;
    iDynTree::Transform m_leftFootToWorldTransform, m_rightFootToWorldTransform;
    iDynTree::Rotation m_desiredNeckOrientation, m_neckOrientation; // desired: as passed (additional_rotation is applied on the device)
    iDynTree::VectorDynSize m_regularizationTerm, m_jointPosition;
    int m_numberOfVariables{0}, m_numberOfConstraints{0}, m_actuatedDOFs{0};
    bool m_isSolutionEvaluated{false};
    bool m_useCoMAsConstraint{false};
    wc_ik_config m_cfg{};

    virtual bool initializeMatrices(const yarp::os::Searchable& config);
    // shared body of the two subclasses' initialize(): configuration parsing + device handle with the chosen solver
    bool initializeDevice(const yarp::os::Searchable& config, const int& actuatedDOFs, const iDynTree::VectorDynSize& minJointsLimit,
                          const iDynTree::VectorDynSize& maxJointsLimit, int solver, wc_ik*& handle, wc_model& model,
                          std::vector<double>& dq, std::vector<double>& lfErr, std::vector<double>& rfErr);

public:
    virtual ~WalkingQPIK();
    virtual bool initialize(const yarp::os::Searchable& config, const int& actuatedDOFs,
                            const iDynTree::VectorDynSize& minJointsLimit, const iDynTree::VectorDynSize& maxJointsLimit) = 0;
    bool setRobotState(const iDynTree::VectorDynSize& jointPosition, const iDynTree::Transform& leftFootToWorldTransform,
                       const iDynTree::Transform& rightFootToWorldTransform, const iDynTree::Rotation& neckOrientation,
                       const iDynTree::Position& comPosition);
    bool setCoMJacobian(const iDynTree::MatrixDynSize& comJacobian);
    bool setLeftFootJacobian(const iDynTree::MatrixDynSize& leftFootJacobian);
    bool setRightFootJacobian(const iDynTree::MatrixDynSize& rightFootJacobian);
    bool setNeckJacobian(const iDynTree::MatrixDynSize& neckJacobian);
    bool setDesiredJointPosition(const iDynTree::VectorDynSize& regularizationTerm);
    void setDesiredFeetTwist(const iDynTree::Twist& leftFootTwist, const iDynTree::Twist& rightFootTwist);
    void setDesiredCoMVelocity(const iDynTree::Vector3& comVelocity);
    void setDesiredFeetTransformation(const iDynTree::Transform& desiredLeftFootToWorldTransform,
                                      const iDynTree::Transform& desiredRightFootToWorldTransform);
    void setDesiredNeckOrientation(const iDynTree::Rotation& desiredNeckOrientation);
    void setDesiredCoMPosition(const iDynTree::Position& desiredComPosition);
    virtual bool solve() = 0;
    virtual bool getSolution(iDynTree::VectorDynSize& output) = 0;
    virtual bool getLeftFootError(iDynTree::VectorDynSize& output) = 0;
    virtual bool getRightFootError(iDynTree::VectorDynSize& output) = 0;

    // north_star aliases (names of a later upstream version)
    bool setJointsState(const iDynTree::VectorDynSize& q, const iDynTree::Transform& l, const iDynTree::Transform& r,
                        const iDynTree::Rotation& n, const iDynTree::Position& c) { return setRobotState(q, l, r, n, c); }
    void setDesiredCoM(const iDynTree::Position& p, const iDynTree::Vector3& v) { setDesiredCoMPosition(p); setDesiredCoMVelocity(v); }
    void setDesiredFeet(const iDynTree::Transform& l, const iDynTree::Transform& r, const iDynTree::Twist& lt, const iDynTree::Twist& rt)
    { setDesiredFeetTransformation(l, r); setDesiredFeetTwist(lt, rt); }
};

class WalkingQPIK_qpOASES : public WalkingQPIK {
    wc_ik* m_handle{nullptr};
    wc_model m_model{};
    std::vector<double> m_dq, m_lfErr, m_rfErr;
    int32_t m_status{0}, m_nwsr{0};

public:
    ~WalkingQPIK_qpOASES() override;
    bool initialize(const yarp::os::Searchable& config, const int& actuatedDOFs, const iDynTree::VectorDynSize& minJointsLimit,
                    const iDynTree::VectorDynSize& maxJointsLimit) final;
    bool solve() final;
    bool getSolution(iDynTree::VectorDynSize& output) final;
    bool getLeftFootError(iDynTree::VectorDynSize& output) final;
    bool getRightFootError(iDynTree::VectorDynSize& output) final;
    int32_t qpStatus() const { return m_status; }
    int32_t workingSetRecalculations() const { return m_nwsr; }
};

// Drop-in for WalkingQPIK_osqp (modules/Walking_module/include/WalkingQPInverseKinematics_osqp.hpp): the same handle with
// wc_ik_set_solver(WC_IK_SOLVER_OSQP); the OSQP workspace lives on the device between solve() calls as osqp-eigen's does.
// solve() is false when OSQP reports an error or isSolutionFeasible fails (_osqp.cpp:349-393), NOT when OSQP merely stops
// at max_iter: that is the reference's behaviour (it only tests m_optimizerSolver->solve()).
class WalkingQPIK_osqp : public WalkingQPIK {
    wc_ik* m_handle{nullptr};
    wc_model m_model{};
    std::vector<double> m_dq, m_lfErr, m_rfErr;
    int32_t m_status{0}, m_iters{0}, m_flags{0}, m_setups{0};

public:
    ~WalkingQPIK_osqp() override;
    bool initialize(const yarp::os::Searchable& config, const int& actuatedDOFs, const iDynTree::VectorDynSize& minJointsLimit,
                    const iDynTree::VectorDynSize& maxJointsLimit) final;
    bool solve() final;
    bool getSolution(iDynTree::VectorDynSize& output) final;
    bool getLeftFootError(iDynTree::VectorDynSize& output) final;
    bool getRightFootError(iDynTree::VectorDynSize& output) final;
    int32_t osqpStatus() const { return m_status; }
    int32_t osqpIterations() const { return m_iters; }
    int32_t solverSetups() const { return m_setups; }
};

#endif
