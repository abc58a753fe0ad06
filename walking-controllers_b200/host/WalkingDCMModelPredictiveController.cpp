// See WalkingDCMModelPredictiveController.hpp.  Reference behaviour mirrored from
// modules/Walking_module/src/WalkingDCMModelPredictiveController.cpp (line numbers in comments).
#include "WalkingDCMModelPredictiveController.hpp"

#include <cmath>
#include <iostream>

WalkingController::~WalkingController() { wc_mpc_destroy(m_handle); }

bool WalkingController::initialize(const yarp::os::Searchable& config)
{
    if (config.isNull()) { std::cerr << "[initialize] Empty configuration for walking controller.\n"; return false; } // :175-179
    const auto& zmp = wc_compat::cfg_list(config, "initial_zmp_position");                                                                 // :317-344
    if (zmp.size() != 2) { std::cerr << "[initialize] The dimension set in the configuration file is not 2.\n"; return false; }
    wc_mpc_default_config(&m_cfg, 1);
    const double dT = wc_compat::cfg_number(config, "sampling_time", 0.016);                                                               // :182
    const double horizonSeconds = wc_compat::cfg_number(config, "controllerHorizon", 2.0);                                                 // :185
    m_cfg.horizon = static_cast<int32_t>(std::round(horizonSeconds / dT));                                                 // :187
    m_cfg.sampling_time = dT;
    if (!config.check("com_height")) { std::cerr << "[initialize] Unable to get the double from searchable.\n"; return false; } // :225
    m_cfg.com_height = wc_compat::cfg_number(config, "com_height", 0.0);
    m_cfg.gravity = wc_compat::cfg_number(config, "gravity_acceleration", 9.81);                                                           // :230
    // weight triplets (row, col, value); getTripletsFromValues stores them TRANSPOSED (Utils.cpp:93-101)
    auto triplets = [&](const char* key, double* M) -> bool {
        const auto& t = wc_compat::cfg_list(config, key);
        if (t.empty() || t.size() % 3) return false;
        for (int k = 0; k < 4; k++) M[k] = 0.0;
        for (size_t k = 0; k < t.size(); k += 3) {
            const int row = static_cast<int>(t[k]), col = static_cast<int>(t[k + 1]);
            if (row >= 2 || col >= 2 || row < 0 || col < 0) return false;
            M[2 * col + row] += t[k + 2];
        }
        return true;
    };
    if (!triplets("stateWeightTriplets", m_cfg.Q)) { std::cerr << "Initialization failed while reading stateWeightTriplets vector.\n"; return false; }
    if (!triplets("inputWeightTriplets", m_cfg.R)) { std::cerr << "Initialization failed while reading inputWeightTriplets vector.\n"; return false; }
    const auto& fs = wc_compat::cfg_list(config, "foot_size");                                                                             // :247-303
    if (fs.size() != 4) { std::cerr << "Please set the foot_size in the configuration file.\n"; return false; }
    m_cfg.foot_x[0] = fs[0]; m_cfg.foot_x[1] = fs[1]; m_cfg.foot_y[0] = fs[2]; m_cfg.foot_y[1] = fs[3];
    m_cfg.hull_tolerance = wc_compat::cfg_number(config, "convex_hull_tolerance", 0.01);                                                   // :306
    m_cfg.initial_zmp[0] = zmp[0]; m_cfg.initial_zmp[1] = zmp[1];
    m_output(0) = zmp[0]; m_output(1) = zmp[1];
    wc_mpc_destroy(m_handle); m_handle = nullptr;
    const wc_status rc = wc_mpc_create(&m_cfg, 1, 0, &m_handle);
    if (rc != WC_OK) { std::cerr << "[initialize] wc_mpc_create failed with status " << rc << "\n"; return false; }
    m_ref.assign(2 * (m_cfg.horizon + 1), 0.0);
    reset();
    return true;
}

bool WalkingController::setConvexHullConstraint(const std::deque<iDynTree::Transform>& leftFoot,
                                                const std::deque<iDynTree::Transform>& rightFoot,
                                                const std::deque<bool>& leftInContact, const std::deque<bool>& rightInContact)
{
    if (leftFoot.empty() || rightFoot.empty() || leftInContact.empty() || rightInContact.empty()) return false;
    // only .front() of each deque is read (:369,378-403); hull + solver are (re)built on the device
    wc_compat::toH12(leftFoot.front(), m_lfH);
    wc_compat::toH12(rightFoot.front(), m_rfH);
    m_lc = leftInContact.front(); m_rc = rightInContact.front();
    const auto feetStatus = std::make_pair(leftInContact.front(), rightInContact.front());
    if (m_feetStatus == feetStatus) return true;                                                                          // :373-374
    m_feetStatus = feetStatus;
    if (!feetStatus.first && !feetStatus.second) {                                                                        // :404-408
        std::cerr << "[setConvexHullConstraint] None foot is in contact How is it possible?.\n";
        return false;
    }
    m_hasSolver = true;
    return true;
}

bool WalkingController::setFeedback(const iDynTree::Vector2& currentState)
{
    if (!m_hasSolver) return false;
    m_dcm[0] = currentState(0); m_dcm[1] = currentState(1);
    return true;
}

bool WalkingController::setReferenceSignal(const std::deque<iDynTree::Vector2>& referenceSignal, const bool& resetTrajectory)
{
    if (!m_hasSolver || referenceSignal.empty()) return false;
    const int N = m_cfg.horizon;
    for (int k = 0; k <= N; k++) { // a shorter deque is continued with its last sample (MPCSolver.cpp:201-215,232-238)
        const iDynTree::Vector2& r = k < static_cast<int>(referenceSignal.size()) ? referenceSignal[k] : referenceSignal.back();
        m_ref[2 * k] = r(0); m_ref[2 * k + 1] = r(1);
    }
    m_reset = resetTrajectory ? 1 : 0;
    return true;
}

bool WalkingController::solve()
{
    m_isSolutionEvaluated = false;                                                                                         // :493
    if (!m_handle || !m_hasSolver) { std::cerr << "[solve] Unable to initialize the solver.\n"; return false; }
    double zmp[2];
    const wc_status rc = wc_mpc_tick_host(m_handle, m_lfH, m_rfH, &m_lc, &m_rc, m_dcm, m_ref.data(), &m_reset, zmp,
                                          &m_lastStatus, &m_lastIters, &m_lastFlags);
    if (rc != WC_OK) { std::cerr << "[solve] wc_mpc_tick_host failed with status " << rc << "\n"; return false; }
    if (m_lastFlags & (WC_MPC_FLAG_NO_CONTACT | WC_MPC_FLAG_NO_SOLVER | WC_MPC_FLAG_NOT_SOLVED)) {
        std::cerr << "[solve] Unable to solve the problem.\n";                                                             // :503-507
        return false;
    }
    m_output(0) = zmp[0]; m_output(1) = zmp[1];                                                                            // :510-511
    if (m_lastFlags & WC_MPC_FLAG_OUTSIDE_HULL) {                                                                          // :513-517
        std::cerr << "[solve] The evaluated ZMP is outside the convexHull.\n";
        return false;
    }
    m_isSolutionEvaluated = true;
    return true;
}

bool WalkingController::getControllerOutput(iDynTree::Vector2& controllerOutput)
{
    if (!m_isSolutionEvaluated) {                                                                                          // :525-530
        std::cerr << "[getControllerOutput] The solution is not evaluated. Please call 'solve()' method.\n";
        return false;
    }
    m_isSolutionEvaluated = false;
    controllerOutput = m_output;
    return true;
}

void WalkingController::reset()
{
    m_feetStatus = std::make_pair(false, false);                                                                           // :540
    m_isSolutionEvaluated = false;
    if (m_handle) wc_mpc_reset(m_handle, nullptr);
}
