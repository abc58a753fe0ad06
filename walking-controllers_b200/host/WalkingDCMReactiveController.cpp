// See WalkingDCMReactiveController.hpp.  Reference behaviour mirrored from
// modules/Walking_module/src/WalkingDCMReactiveController.cpp (line numbers in comments).
#include "WalkingDCMReactiveController.hpp"

#include <iostream>

WalkingDCMReactiveController::~WalkingDCMReactiveController() { wc_glue_destroy(m_handle); }

bool WalkingDCMReactiveController::initialize(const yarp::os::Searchable& config)
{
    if (config.isNull()) { std::cerr << "[initialize] Empty configuration for DCM controller.\n"; return false; }            // :24-28
    if (!config.check("kDCM")) { std::cerr << "[initialize] Unable to get the double from searchable.\n"; return false; }    // :31-35
    m_kDCM = wc_compat::cfg_number(config, "kDCM", 0.0);
    if (!config.check("com_height")) { std::cerr << "[initialize] Unable to get the double from searchable.\n"; return false; } // :37-42
    wc_glue_config cfg;
    wc_glue_default_config(&cfg);
    cfg.com_height = wc_compat::cfg_number(config, "com_height", 0.0);
    cfg.gravity = wc_compat::cfg_number(config, "gravity_acceleration", 9.81);                                                                 // :43-44
    wc_glue_destroy(m_handle); m_handle = nullptr;
    const wc_status rc = wc_glue_create(&cfg, 1, 0, &m_handle);
    if (rc != WC_OK) { std::cerr << "[initialize] wc_glue_create failed with status " << rc << "\n"; return false; }
    m_isInitialized = true;
    return true;
}

void WalkingDCMReactiveController::setFeedback(const iDynTree::Vector2& dcmFeedback) { m_dcmFeedback = dcmFeedback; }         // :51-54

void WalkingDCMReactiveController::setReferenceSignal(const iDynTree::Vector2& dcmPositionDesired,
                                                      const iDynTree::Vector2& dcmVelocityDesired)                            // :56-61
{
    m_dcmPositionDesired = dcmPositionDesired;
    m_dcmVelocityDesired = dcmVelocityDesired;
}

bool WalkingDCMReactiveController::evaluateControl()
{
    m_controlEvaluated = false;                                                                                                // :65
    if (!m_isInitialized) {
        std::cerr << "[evaluateControl] The controller is not initialized. Please call 'initialize()'\n";                     // :67-72
        return false;
    }
    const double xd[2] = {m_dcmPositionDesired(0), m_dcmPositionDesired(1)};
    const double vd[2] = {m_dcmVelocityDesired(0), m_dcmVelocityDesired(1)};
    const double x[2] = {m_dcmFeedback(0), m_dcmFeedback(1)};
    double u[2];
    if (wc_glue_reactive_dcm_host(m_handle, m_kDCM, nullptr, xd, vd, x, u) != WC_OK) return false;                           // :74-78
    m_controllerOutput(0) = u[0]; m_controllerOutput(1) = u[1];
    m_controlEvaluated = true;
    return true;
}

bool WalkingDCMReactiveController::getControllerOutput(iDynTree::Vector2& controllerOutput)
{
    if (!m_controlEvaluated) {
        std::cerr << "[getControllerOutput] Please before call evaluateControl() method.\n";                                  // :85-89
        return false;
    }
    controllerOutput = m_controllerOutput;
    return true;
}
