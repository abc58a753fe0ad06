// See WalkingQPInverseKinematics.hpp.  Reference behaviour mirrored from
// modules/Walking_module/src/WalkingQPInverseKinematics.cpp and ..._qpOASES.cpp (line numbers in comments).
#include "WalkingQPInverseKinematics.hpp"

#include <cmath>
#include <cstring>
#include <iostream>

namespace {
bool triplets3(const yarp::os::Searchable& cfg, const char* key, double* M)
{
    const auto& t = wc_compat::cfg_list(cfg, key);
    if (t.empty() || t.size() % 3) return false;
    for (int k = 0; k < 9; k++) M[k] = 0.0;
    for (size_t k = 0; k < t.size(); k += 3) {
        const int row = static_cast<int>(t[k]), col = static_cast<int>(t[k + 1]);
        if (row < 0 || col < 0 || row >= 3 || col >= 3) return false;
        M[3 * col + row] += t[k + 2]; // getTripletsFromValues pushes (col, row, value): Utils.cpp:93-101
    }
    return true;
}
} // namespace

WalkingQPIK::~WalkingQPIK() {}

bool WalkingQPIK::initializeMatrices(const yarp::os::Searchable& config) // QPIK.cpp:25-116
{
    if (!m_useCoMAsConstraint && !triplets3(config, "comWeightTriplets", m_cfg.com_weight)) {
        std::cerr << "Initialization failed while reading comWeightTriplets vector.\n"; return false;
    }
    if (!triplets3(config, "neckWeightTriplets", m_cfg.neck_weight)) {
        std::cerr << "Initialization failed while reading neckWeightTriplets vector.\n"; return false;
    }
    const auto& w = wc_compat::cfg_list(config, "jointRegularizationWeights");
    const auto& g = wc_compat::cfg_list(config, "jointRegularizationGains");
    if (static_cast<int>(w.size()) != m_actuatedDOFs) { std::cerr << "Initialization failed while reading jointRegularizationWeights vector.\n"; return false; }
    if (static_cast<int>(g.size()) != m_actuatedDOFs) { std::cerr << "Initialization failed while reading jointRegularizationGains vector.\n"; return false; }
    for (int i = 0; i < m_actuatedDOFs; i++) { m_cfg.joint_reg_weight[i] = w[i]; m_cfg.joint_reg_gain[i] = g[i]; }
    m_comJacobian.resize(3, m_numberOfVariables); m_neckJacobian.resize(6, m_numberOfVariables);
    m_leftFootJacobian.resize(6, m_numberOfVariables); m_rightFootJacobian.resize(6, m_numberOfVariables);
    const char* keys[4] = {"k_posFoot", "k_attFoot", "k_neck", "k_posCom"};
    double* dst[4] = {&m_cfg.k_pos_foot, &m_cfg.k_att_foot, &m_cfg.k_neck, &m_cfg.k_pos_com};
    for (int k = 0; k < 4; k++) {
        if (!config.check(keys[k])) { std::cerr << "Initialization failed while reading " << keys[k] << ".\n"; return false; }
        *dst[k] = wc_compat::cfg_number(config, keys[k], 0.0);
    }
    return true;
}

bool WalkingQPIK::setRobotState(const iDynTree::VectorDynSize& jointPosition, const iDynTree::Transform& l,
                                const iDynTree::Transform& r, const iDynTree::Rotation& neckOrientation,
                                const iDynTree::Position& comPosition)
{
    if (static_cast<int>(jointPosition.size()) != m_actuatedDOFs) { // :127-132
        std::cerr << "[setRobotState] The size of the jointPosition vector is not coherent with the number of the actuated Joint\n";
        return false;
    }
    m_jointPosition = jointPosition; m_leftFootToWorldTransform = l; m_rightFootToWorldTransform = r;
    m_neckOrientation = neckOrientation; m_comPosition = comPosition;
    return true;
}

void WalkingQPIK::setDesiredNeckOrientation(const iDynTree::Rotation& d) { m_desiredNeckOrientation = d; } // R_des * R_add on the device (:143-146)

static bool check_jac(const iDynTree::MatrixDynSize& J, size_t rows, size_t cols, const char* who)
{
    if (J.rows() != rows) { std::cerr << "[" << who << "] the number of rows has to be equal to " << rows << ".\n"; return false; }
    if (J.cols() != cols) { std::cerr << "[" << who << "] the number of rows has to be equal to" << cols << "\n"; return false; }
    return true;
}
bool WalkingQPIK::setCoMJacobian(const iDynTree::MatrixDynSize& J) { if (!check_jac(J, 3, m_actuatedDOFs + 6, "setCoMJacobian")) return false; m_comJacobian = J; return true; }
bool WalkingQPIK::setLeftFootJacobian(const iDynTree::MatrixDynSize& J) { if (!check_jac(J, 6, m_actuatedDOFs + 6, "setCoMJacobian")) return false; m_leftFootJacobian = J; return true; }
bool WalkingQPIK::setRightFootJacobian(const iDynTree::MatrixDynSize& J) { if (!check_jac(J, 6, m_actuatedDOFs + 6, "setCoMJacobian")) return false; m_rightFootJacobian = J; return true; }
bool WalkingQPIK::setNeckJacobian(const iDynTree::MatrixDynSize& J) { if (!check_jac(J, 6, m_actuatedDOFs + 6, "setNeckMJacobian")) return false; m_neckJacobian = J; return true; } // rows 3..5 are used (:214-215)

bool WalkingQPIK::setDesiredJointPosition(const iDynTree::VectorDynSize& regularizationTerm)
{
    if (static_cast<int>(regularizationTerm.size()) != m_actuatedDOFs) { // :222-227
        std::cerr << "[setDesiredJointPosition] The number of the desired joint position has to be equal to the number of actuated joints\n";
        return false;
    }
    m_regularizationTerm = regularizationTerm;
    return true;
}
void WalkingQPIK::setDesiredFeetTransformation(const iDynTree::Transform& l, const iDynTree::Transform& r) { m_desiredLeftFootToWorldTransform = l; m_desiredRightFootToWorldTransform = r; }
void WalkingQPIK::setDesiredFeetTwist(const iDynTree::Twist& l, const iDynTree::Twist& r) { m_leftFootTwist = l; m_rightFootTwist = r; }
void WalkingQPIK::setDesiredCoMVelocity(const iDynTree::Vector3& v) { m_comVelocity = v; }
void WalkingQPIK::setDesiredCoMPosition(const iDynTree::Position& p) { m_desiredComPosition = p; }

// ------------------------------------------------------------------------------------------ qpOASES variant
WalkingQPIK_qpOASES::~WalkingQPIK_qpOASES() { wc_ik_destroy(m_handle); }

bool WalkingQPIK::initializeDevice(const yarp::os::Searchable& config, const int& actuatedDOFs,
                                   const iDynTree::VectorDynSize& minJointsLimit, const iDynTree::VectorDynSize& maxJointsLimit,
                                   int solver, wc_ik*& m_handle, wc_model& m_model, std::vector<double>& m_dq,
                                   std::vector<double>& m_lfErr, std::vector<double>& m_rfErr)
{
    m_actuatedDOFs = actuatedDOFs;
    if (config.isNull()) { std::cerr << "[initialize] Empty configuration for QP-IK solver.\n"; return false; } // _qpOASES.cpp:60-64
    if (actuatedDOFs < 1 || actuatedDOFs > WC_MAX_DOFS) return false;
    wc_ik_default_config(&m_cfg);
    m_cfg.n_dofs = actuatedDOFs;
    m_useCoMAsConstraint = wc_compat::cfg_number(config, "useCoMAsConstraint", 0.0) != 0.0;                                     // :66
    m_cfg.use_com_as_constraint = m_useCoMAsConstraint ? 1 : 0;
    m_numberOfVariables = m_actuatedDOFs + 6;                                                                   // :73
    m_numberOfConstraints = m_useCoMAsConstraint ? 15 : 12;                                                     // :77-80
    const auto& reg = wc_compat::cfg_list(config, "jointRegularization");                                                       // :94-103
    if (static_cast<int>(reg.size()) != m_actuatedDOFs) {
        std::cerr << "[initialize] Unable to convert a YARP list to an iDynTree::VectorDynSize, joint regularization\n"; return false;
    }
    m_regularizationTerm.resize(m_actuatedDOFs); m_jointPosition.resize(m_actuatedDOFs);
    for (int i = 0; i < m_actuatedDOFs; i++) {
        m_regularizationTerm(i) = reg[i] * (3.14159265358979323846 / 180.0);
        m_cfg.joint_reg_rad[i] = m_regularizationTerm(i);
    }
    if (!initializeMatrices(config)) { std::cerr << "[initialize] Unable to Initialize the constant matrix.\n"; return false; }
    if (minJointsLimit.size() != maxJointsLimit.size() || static_cast<int>(minJointsLimit.size()) != m_actuatedDOFs) { // :26-37
        std::cerr << "[setVelocityBounds] The size of the vector limits has to be equal to the number of the joint\n"; return false;
    }
    for (int i = 0; i < m_actuatedDOFs; i++) { m_cfg.min_vel[i] = minJointsLimit(i); m_cfg.max_vel[i] = maxJointsLimit(i); }
    const auto& add = wc_compat::cfg_list(config, "additional_rotation");                                                       // :118
    if (add.size() != 9) { std::cerr << "[initialize] Unable to set the additional rotation.\n"; return false; }
    for (int k = 0; k < 9; k++) m_cfg.additional_rotation[k] = add[k];
    m_cfg.nwsr = 100;                                                                                           // :312
    // the device handle needs a kinematic table only for wc_ik_tick / wc_fk; this class feeds Jacobians itself
    wc_model_icub_like(&m_model);
    if (m_model.n_dofs != actuatedDOFs) {
        // any other joint set (12 .. 26 joints): a placeholder table of the right size -- a pelvis with two serial chains -- is enough,
        // since the Jacobians and transforms come from the caller (setRobotState / set*Jacobian)
        if (actuatedDOFs < 12 || actuatedDOFs > 26) { std::cerr << "[initialize] actuatedDOFs must be in [12, 26].\n"; return false; }
        std::memset(&m_model, 0, sizeof(m_model));
        m_model.n_links = actuatedDOFs + 1; m_model.n_dofs = actuatedDOFs;
        m_model.parent[0] = -1; m_model.dof[0] = -1; m_model.mass[0] = 1.0;
        m_model.joint_R[0][0] = m_model.joint_R[0][4] = m_model.joint_R[0][8] = 1.0;
        const int half = actuatedDOFs / 2;
        for (int i = 1; i <= actuatedDOFs; i++) {
            m_model.parent[i] = (i == 1 || i == half + 1) ? 0 : i - 1;
            m_model.dof[i] = i - 1; m_model.mass[i] = 1.0;
            m_model.joint_R[i][0] = m_model.joint_R[i][4] = m_model.joint_R[i][8] = 1.0;
            m_model.joint_p[i][2] = -0.1; m_model.axis[i][1] = 1.0;
        }
        m_model.l_sole_link = half; m_model.r_sole_link = actuatedDOFs; m_model.neck_link = 0;
        m_model.l_sole_R[0] = m_model.l_sole_R[4] = m_model.l_sole_R[8] = 1.0;
        m_model.r_sole_R[0] = m_model.r_sole_R[4] = m_model.r_sole_R[8] = 1.0;
        m_model.neck_R[0] = m_model.neck_R[4] = m_model.neck_R[8] = 1.0;
    }
    wc_ik_destroy(m_handle); m_handle = nullptr;
    const wc_status rc = wc_ik_create(&m_cfg, &m_model, 1, 0, &m_handle);
    if (rc != WC_OK) { std::cerr << "[initialize] wc_ik_create failed with status " << rc << "\n"; return false; }
    if (solver != WC_IK_SOLVER_QPOASES && wc_ik_set_solver(m_handle, solver, nullptr) != WC_OK) return false; // OSQP defaults (_osqp.cpp:127-131)
    if (solver == WC_IK_SOLVER_OSQP) m_numberOfConstraints += m_actuatedDOFs;                                   // _osqp.cpp:77-82
    m_dq.assign(m_actuatedDOFs, 0.0); m_lfErr.assign(6, 0.0); m_rfErr.assign(6, 0.0);
    return true;
}

bool WalkingQPIK_qpOASES::initialize(const yarp::os::Searchable& config, const int& actuatedDOFs,
                                     const iDynTree::VectorDynSize& minJointsLimit, const iDynTree::VectorDynSize& maxJointsLimit)
{
    return initializeDevice(config, actuatedDOFs, minJointsLimit, maxJointsLimit, WC_IK_SOLVER_QPOASES, m_handle, m_model, m_dq, m_lfErr, m_rfErr);
}

bool WalkingQPIK_qpOASES::solve()
{
    m_isSolutionEvaluated = false;
    if (!m_handle) return false;
    // the regularisation target may have been changed through setDesiredJointPosition: re-create is avoided
    // by keeping it in the per-call configuration only when it differs (rare; not on the tick path)
    bool changed = false;
    for (int i = 0; i < m_actuatedDOFs; i++) if (m_cfg.joint_reg_rad[i] != m_regularizationTerm(i)) { m_cfg.joint_reg_rad[i] = m_regularizationTerm(i); changed = true; }
    if (changed) {
        wc_ik_destroy(m_handle); m_handle = nullptr;
        if (wc_ik_create(&m_cfg, &m_model, 1, 0, &m_handle) != WC_OK) return false;
    }
    double lfH[12], rfH[12], lfd[12], rfd[12];
    wc_compat::toH12(m_leftFootToWorldTransform, lfH); wc_compat::toH12(m_rightFootToWorldTransform, rfH);
    wc_compat::toH12(m_desiredLeftFootToWorldTransform, lfd); wc_compat::toH12(m_desiredRightFootToWorldTransform, rfd);
    double ltw[6], rtw[6]; // iDynTree::Twist has no data(): element access only
    wc_compat::twist6(m_leftFootTwist, ltw); wc_compat::twist6(m_rightFootTwist, rtw);
    const wc_status rc = wc_ik_solve_with_jacobians_host(
        m_handle, m_jointPosition.data(), lfH, rfH, m_neckOrientation.data(), m_comPosition.data(),
        m_leftFootJacobian.data(), m_rightFootJacobian.data(), m_neckJacobian.data(), m_comJacobian.data(),
        lfd, rfd, ltw, rtw, m_desiredComPosition.data(), m_comVelocity.data(),
        m_desiredNeckOrientation.data(), m_dq.data(), m_lfErr.data(), m_rfErr.data(), &m_status, &m_nwsr);
    if (rc != WC_OK || m_status != WC_QP_SUCCESSFUL_RETURN) { std::cerr << "[solve] Unable to solve the problem.\n"; return false; } // :316-335
    m_isSolutionEvaluated = true;
    return true;
}

bool WalkingQPIK_qpOASES::getSolution(iDynTree::VectorDynSize& output)
{
    if (!m_isSolutionEvaluated) { // :343-348
        std::cerr << "[getSolution] The solution is not evaluated. Please call 'solve()' method.\n"; return false;
    }
    if (static_cast<int>(output.size()) != m_actuatedDOFs) output.resize(m_actuatedDOFs);
    for (int i = 0; i < m_actuatedDOFs; i++) output(i) = m_dq[i];
    m_isSolutionEvaluated = false;
    return true;
}
bool WalkingQPIK_qpOASES::getLeftFootError(iDynTree::VectorDynSize& output) { if (output.size() != 6) output.resize(6); for (int i = 0; i < 6; i++) output(i) = m_lfErr[i]; return true; }
bool WalkingQPIK_qpOASES::getRightFootError(iDynTree::VectorDynSize& output) { if (output.size() != 6) output.resize(6); for (int i = 0; i < 6; i++) output(i) = m_rfErr[i]; return true; }

// ------------------------------------------------------------------------------------------ osqp variant
WalkingQPIK_osqp::~WalkingQPIK_osqp() { wc_ik_destroy(m_handle); }

bool WalkingQPIK_osqp::initialize(const yarp::os::Searchable& config, const int& actuatedDOFs,
                                  const iDynTree::VectorDynSize& minJointsLimit, const iDynTree::VectorDynSize& maxJointsLimit)
{
    return initializeDevice(config, actuatedDOFs, minJointsLimit, maxJointsLimit, WC_IK_SOLVER_OSQP, m_handle, m_model, m_dq, m_lfErr, m_rfErr);
}

bool WalkingQPIK_osqp::solve()
{
    m_isSolutionEvaluated = false;                                                                              // _osqp.cpp:351
    if (!m_handle) return false;
    bool changed = false;
    for (int i = 0; i < m_actuatedDOFs; i++) if (m_cfg.joint_reg_rad[i] != m_regularizationTerm(i)) { m_cfg.joint_reg_rad[i] = m_regularizationTerm(i); changed = true; }
    if (changed) { // a new regularisation target needs a new device configuration; the OSQP workspace starts over
        wc_ik_destroy(m_handle); m_handle = nullptr;
        if (wc_ik_create(&m_cfg, &m_model, 1, 0, &m_handle) != WC_OK || wc_ik_set_solver(m_handle, WC_IK_SOLVER_OSQP, nullptr) != WC_OK) return false;
    }
    double lfH[12], rfH[12], lfd[12], rfd[12];
    wc_compat::toH12(m_leftFootToWorldTransform, lfH); wc_compat::toH12(m_rightFootToWorldTransform, rfH);
    wc_compat::toH12(m_desiredLeftFootToWorldTransform, lfd); wc_compat::toH12(m_desiredRightFootToWorldTransform, rfd);
    double ltw[6], rtw[6]; // iDynTree::Twist has no data(): element access only
    wc_compat::twist6(m_leftFootTwist, ltw); wc_compat::twist6(m_rightFootTwist, rtw);
    const wc_status rc = wc_ik_solve_with_jacobians_host(
        m_handle, m_jointPosition.data(), lfH, rfH, m_neckOrientation.data(), m_comPosition.data(),
        m_leftFootJacobian.data(), m_rightFootJacobian.data(), m_neckJacobian.data(), m_comJacobian.data(),
        lfd, rfd, ltw, rtw, m_desiredComPosition.data(), m_comVelocity.data(),
        m_desiredNeckOrientation.data(), m_dq.data(), m_lfErr.data(), m_rfErr.data(), &m_status, &m_iters);
    if (rc != WC_OK || wc_ik_osqp_info_host(m_handle, &m_flags, &m_setups) != WC_OK) return false;
    if (m_flags & 1) { std::cerr << "[solve] Unable to solve the problem.\n"; return false; }                 // :377-381
    if (m_flags & 2) { std::cerr << "[solve] The solution is not feasible.\n"; return false; }                // :384-388
    m_isSolutionEvaluated = true;
    return true;
}

bool WalkingQPIK_osqp::getSolution(iDynTree::VectorDynSize& output)
{
    if (!m_isSolutionEvaluated) {                                                                               // :412-417
        std::cerr << "[getSolution] The solution is not evaluated. Please call 'solve()' method.\n"; return false;
    }
    if (static_cast<int>(output.size()) != m_actuatedDOFs) output.resize(m_actuatedDOFs);
    for (int i = 0; i < m_actuatedDOFs; i++) output(i) = m_dq[i];                                               // no latch reset in this variant
    return true;
}
bool WalkingQPIK_osqp::getLeftFootError(iDynTree::VectorDynSize& output)
{
    if (!m_isSolutionEvaluated) { std::cerr << "[getLeftFootError] The solution is not evaluated. Please call 'solve()' method.\n"; return false; } // :432-437
    if (output.size() != 6) output.resize(6);
    for (int i = 0; i < 6; i++) output(i) = m_lfErr[i];
    return true;
}
bool WalkingQPIK_osqp::getRightFootError(iDynTree::VectorDynSize& output)
{
    if (!m_isSolutionEvaluated) { std::cerr << "[getRightFootError] The solution is not evaluated. Please call 'solve()' method.\n"; return false; } // :445-450
    if (output.size() != 6) output.resize(6);
    for (int i = 0; i < 6; i++) output(i) = m_rfErr[i];
    return true;
}
