#include "WalkingForwardKinematics.hpp"

#include <cstring>
#include <iostream>

WalkingFK::~WalkingFK() { wc_ik_destroy(m_handle); }

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

bool WalkingFK::setInternalRobotState(const iDynTree::VectorDynSize& q, const iDynTree::VectorDynSize&)
{
    if (static_cast<int>(q.size()) != m_nJ) { std::cerr << "[setInternalRobotState] Error while updating the state.\n"; return false; }
    for (int i = 0; i < m_nJ; i++) m_q[i] = q(i);
    m_firstStep = true; // as the reference (:273)
    m_evaluated = false;
    return true;
}

bool WalkingFK::evaluate()
{
    if (m_evaluated) return true;
    if (!m_handle) return false;
    double bH[12]; m_worldToBaseSole.toH12(bH);
    const uint8_t left = m_prevContactLeft ? 1 : 0;
    if (wc_fk_host(m_handle, m_q.data(), bH, &left, m_lfH.data(), m_rfH.data(), m_neckR.data(), m_com.data(), m_Jl.data(),
                   m_Jr.data(), m_Jn.data(), m_Jc.data()) != WC_OK)
        return false;
    m_evaluated = true;
    return true;
}

iDynTree::Transform WalkingFK::getLeftFootToWorldTransform() { evaluate(); return iDynTree::Transform::fromH12(m_lfH.data()); }
iDynTree::Transform WalkingFK::getRightFootToWorldTransform() { evaluate(); return iDynTree::Transform::fromH12(m_rfH.data()); }
iDynTree::Rotation WalkingFK::getNeckOrientation() { evaluate(); iDynTree::Rotation R; std::memcpy(R.data(), m_neckR.data(), 9 * sizeof(double)); return R; }
bool WalkingFK::getCoMPosition(iDynTree::Position& c) { if (!evaluate()) return false; for (int k = 0; k < 3; k++) c(k) = m_com[k]; return true; }

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
