#include "WalkingForwardKinematics.hpp"

#include <cstring>
#include <iostream>

WalkingFK::~WalkingFK() { wc_ik_destroy(m_handle); wc_glue_destroy(m_glue); }

bool WalkingFK::initialize(const yarp::os::Searchable& config, const wc_model& model)
{
    if (config.isNull()) { std::cerr << "[initialize] Empty configuration for fk solver.\n"; return false; }
    m_model = model; m_nJ = model.n_dofs;
    wc_ik_config cfg; wc_ik_default_config(&cfg); cfg.n_dofs = m_nJ;
    wc_ik_destroy(m_handle); m_handle = nullptr;
    if (wc_ik_create(&cfg, &m_model, 1, 0, &m_handle) != WC_OK) { std::cerr << "[initialize] Unable to set the robot model.\n"; return false; }
    const int nV = m_nJ + 6;
    m_q.assign(m_nJ, 0.0); m_lfH.assign(12, 0.0); m_rfH.assign(12, 0.0); m_neckR.assign(9, 0.0); m_com.assign(3, 0.0);
    m_Jl.assign(6 * nV, 0.0); m_Jr.assign(6 * nV, 0.0); m_Jn.assign(6 * nV, 0.0); m_Jc.assign(3 * nV, 0.0);
    m_firstStep = true; m_prevContactLeft = false;
    // CoM filters and the DCM (:131-162): sampling_time, com_height (omega), cut_frequency, use_filters
    m_dq.assign(m_nJ, 0.0);
    wc_glue_config gc; wc_glue_default_config(&gc);
    gc.n_dofs = m_nJ;
    gc.sampling_time = wc_compat::cfg_number(config, "sampling_time", gc.sampling_time);
    gc.com_height = wc_compat::cfg_number(config, "com_height", gc.com_height);
    gc.gravity = wc_compat::cfg_number(config, "gravity_acceleration", 9.81);
    m_cutFrequency = wc_compat::cfg_number(config, "cut_frequency", 0.0);
    m_useFilters = wc_compat::cfg_number(config, "use_filters", 0.0) != 0.0;
    wc_glue_destroy(m_glue); m_glue = nullptr;
    if (wc_glue_create(&gc, 1, 0, &m_glue) != WC_OK) { std::cerr << "[initialize] Unable to create the CoM filters.\n"; return false; }
    m_comEvaluated = m_dcmEvaluated = false;
    return true;
}

bool WalkingFK::evaluateWorldToBaseTransformation(const iDynTree::Transform& leftFootTransform,
                                                  const iDynTree::Transform& rightFootTransform, const bool& isLeftFixedFrame)
{
    if (isLeftFixedFrame) {
        if (!m_prevContactLeft || m_firstStep) { m_worldToBaseSole = leftFootTransform; m_prevContactLeft = true; m_evaluated = false; }
    } else {
        if (m_prevContactLeft || m_firstStep) { m_worldToBaseSole = rightFootTransform; m_prevContactLeft = false; m_evaluated = false; }
    }
    m_firstStep = false;
    return true;
}

bool WalkingFK::setInternalRobotState(const iDynTree::VectorDynSize& q, const iDynTree::VectorDynSize& dq)
{
    if (static_cast<int>(q.size()) != m_nJ) { std::cerr << "[setInternalRobotState] Error while updating the state.\n"; return false; }
    for (int i = 0; i < m_nJ; i++) m_q[i] = q(i);
    for (int i = 0; i < m_nJ; i++) m_dq[i] = static_cast<int>(dq.size()) == m_nJ ? dq(i) : 0.0;
    m_firstStep = true; // as the reference (:273)
    m_evaluated = false;
    return true;
}

bool WalkingFK::evaluate()
{
    if (m_evaluated) return true;
    if (!m_handle) return false;
    double bH[12]; wc_compat::toH12(m_worldToBaseSole, bH);
    const uint8_t left = m_prevContactLeft ? 1 : 0;
    if (wc_fk_host(m_handle, m_q.data(), bH, &left, m_lfH.data(), m_rfH.data(), m_neckR.data(), m_com.data(), m_Jl.data(),
                   m_Jr.data(), m_Jn.data(), m_Jc.data()) != WC_OK)
        return false;
    m_evaluated = true;
    return true;
}

iDynTree::Transform WalkingFK::getLeftFootToWorldTransform() { evaluate(); return wc_compat::fromH12(m_lfH.data()); }
iDynTree::Transform WalkingFK::getRightFootToWorldTransform() { evaluate(); return wc_compat::fromH12(m_rfH.data()); }
iDynTree::Rotation WalkingFK::getNeckOrientation() { evaluate(); iDynTree::Rotation R; std::memcpy(R.data(), m_neckR.data(), 9 * sizeof(double)); return R; }
bool WalkingFK::getCoMPosition(iDynTree::Position& c)
{
    if (!evaluate()) return false;
    for (int k = 0; k < 3; k++) c(k) = (m_comEvaluated && m_useFilters) ? m_comOut[k] : m_com[k];
    return true;
}

bool WalkingFK::evaluateCoM()
{
    m_comEvaluated = false;
    if (!evaluate() || !m_glue) { std::cerr << "[evaluateCoM] The KinDynComputations solver is not initialized.\n"; return false; }
    if (wc_glue_evaluate_com_host(m_glue, m_com.data(), m_Jc.data(), m_dq.data(), m_cutFrequency, m_useFilters ? 1 : 0, m_comOut, m_comVel,
                                  m_dcmOut) != WC_OK)
        return false;
    m_comEvaluated = true;
    return true;
}

bool WalkingFK::evaluateDCM()
{
    m_dcmEvaluated = false;
    if (!m_comEvaluated) {
        std::cerr << "[evaluateDCM] The CoM position and velocity is not evaluated. Please call evaluateCoM method.\n"; return false;
    }
    m_dcmEvaluated = true; // the DCM of this CoM / CoM velocity pair was computed by the same device call
    return true;
}

bool WalkingFK::getDCM(iDynTree::Vector2& dcm)
{
    if (!m_dcmEvaluated) { std::cerr << "[getDCM] The dcm is not evaluated. Please call evaluateDCM() method.\n"; return false; }
    dcm(0) = m_dcmOut[0]; dcm(1) = m_dcmOut[1];
    return true;
}

bool WalkingFK::getCoMVelocity(iDynTree::Vector3& v)
{
    if (!m_comEvaluated) { std::cerr << "[getCoMVelocity] The CoM is not evaluated. Please call evaluateCoM() method.\n"; return false; }
    for (int k = 0; k < 3; k++) v(k) = m_comVel[k];
    return true;
}

static bool copy_jac(const std::vector<double>& src, size_t rows, size_t cols, iDynTree::MatrixDynSize& J)
{
    if (J.rows() != rows || J.cols() != cols) J.resize(rows, cols);
    std::memcpy(J.data(), src.data(), rows * cols * sizeof(double));
    return true;
}
bool WalkingFK::getLeftFootJacobian(iDynTree::MatrixDynSize& J) { return evaluate() && copy_jac(m_Jl, 6, m_nJ + 6, J); }
bool WalkingFK::getRightFootJacobian(iDynTree::MatrixDynSize& J) { return evaluate() && copy_jac(m_Jr, 6, m_nJ + 6, J); }
bool WalkingFK::getNeckJacobian(iDynTree::MatrixDynSize& J) { return evaluate() && copy_jac(m_Jn, 6, m_nJ + 6, J); }
bool WalkingFK::getCoMJacobian(iDynTree::MatrixDynSize& J) { return evaluate() && copy_jac(m_Jc, 3, m_nJ + 6, J); }
