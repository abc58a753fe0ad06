// See TrajectoryGenerator.hpp.
#include "TrajectoryGenerator.hpp"

#include <cmath>
#include <iostream>

TrajectoryGenerator::~TrajectoryGenerator() { wc_traj_destroy(m_handle); }

bool TrajectoryGenerator::initialize(const yarp::os::Searchable& config) { return configurePlanner(config); }

bool TrajectoryGenerator::configurePlanner(const yarp::os::Searchable& config)
{
    if (config.isNull()) { std::cerr << "[configurePlanner] Empty configuration for unicycle planner.\n"; return false; } // :46-50
    m_dT = wc_compat::cfg_number(config, "sampling_time", 0.016);                                                                          // :54
    m_plannerHorizon = wc_compat::cfg_number(config, "plannerHorizon", 20.0);                                                              // :55
    const auto& l = wc_compat::cfg_list(config, "leftZMPDelta");                                                                           // :65-79
    const auto& r = wc_compat::cfg_list(config, "rightZMPDelta");
    if (l.size() != 2 || r.size() != 2) { std::cerr << "[configurePlanner] Initialization failed while reading the ZMP deltas.\n"; return false; }
    const double nominalDuration = wc_compat::cfg_number(config, "nominalDuration", 4.0);                                                  // :97
    const double switchOverSwingRatio = wc_compat::cfg_number(config, "switchOverSwingRatio", 0.4);                                        // :99-100
    m_samples = static_cast<int>(std::round(m_plannerHorizon / m_dT));
    wc_traj_default_config(&m_cfg, m_samples, 0);
    m_cfg.sampling_time = m_dT;
    m_cfg.step_ticks = static_cast<int32_t>(std::round(nominalDuration / m_dT));
    m_cfg.ds_ticks = static_cast<int32_t>(nominalDuration * switchOverSwingRatio / (1.0 + switchOverSwingRatio) / m_dT);
    m_cfg.n_steps = (m_samples + m_cfg.step_ticks) / m_cfg.step_ticks + 8 + 3;
    m_cfg.nominal_width = wc_compat::cfg_number(config, "nominalWidth", 0.04);                                                             // :103
    m_cfg.step_height = wc_compat::cfg_number(config, "stepHeight", 0.005);                                                                // :93
    m_cfg.com_height = wc_compat::cfg_number(config, "com_height", 0.49);                                                                  // :96
    m_cfg.zmp_delta[0] = 0.5 * (l[0] + r[0]); m_cfg.zmp_delta[1] = 0.5 * (r[1] - l[1]);
    m_swingLeft = wc_compat::cfg_number(config, "swingLeft", 1.0) != 0.0;                                                                  // :105
    wc_traj_destroy(m_handle); m_handle = nullptr;
    const wc_status rc = wc_traj_create(&m_cfg, 1, 0, &m_handle);
    if (rc != WC_OK) { std::cerr << "[configurePlanner] wc_traj_create failed with status " << rc << "\n"; return false; }
    m_computed = false;
    return true;
}

void TrajectoryGenerator::setGait(double stepLength, double headingRatePerStep, double initialHeading)
{
    m_stepLength = stepLength; m_headingRate = headingRatePerStep; m_heading = initialHeading;
}

bool TrajectoryGenerator::generate(int firstSample)
{
    if (!m_handle) return false;
    const size_t n = static_cast<size_t>(m_samples);
    m_dcmPos.resize(2 * n); m_dcmVel.resize(2 * n); m_lfH.resize(12 * n); m_rfH.resize(12 * n); m_lfTw.resize(6 * n); m_rfTw.resize(6 * n);
    m_comH.resize(n); m_comHV.resize(n); m_lc.resize(n); m_rc.resize(n); m_lfix.resize(n);
    const int32_t phase = 0;
    const uint8_t firstLeft = m_swingLeft ? 0 : 1; // swinging the left foot first = standing on the right one
    const wc_status rc = wc_traj_generate_host(m_handle, &phase, &m_stepLength, &m_headingRate, &m_heading, &firstLeft, firstSample, m_samples,
                                               m_dcmPos.data(), m_dcmVel.data(), m_lfH.data(), m_rfH.data(), m_lfTw.data(), m_rfTw.data(),
                                               m_lc.data(), m_rc.data(), m_lfix.data(), m_comH.data(), m_comHV.data());
    m_computed = rc == WC_OK;
    return m_computed;
}

bool TrajectoryGenerator::generateFirstTrajectories() { return generate(0); }
bool TrajectoryGenerator::updateTrajectories(int firstTick) { return generate(firstTick); }

#define WC_REQUIRE_COMPUTED(name) \
    if (!isTrajectoryComputed()) { std::cerr << "[" name "] No trajectories are available\n"; return false; }

bool TrajectoryGenerator::getDCMPositionTrajectory(std::vector<iDynTree::Vector2>& out)
{
    WC_REQUIRE_COMPUTED("getDCMPositionTrajectory")
    out.resize(m_samples);
    for (int s = 0; s < m_samples; s++) { out[s](0) = m_dcmPos[2 * s]; out[s](1) = m_dcmPos[2 * s + 1]; }
    return true;
}

bool TrajectoryGenerator::getDCMVelocityTrajectory(std::vector<iDynTree::Vector2>& out)
{
    WC_REQUIRE_COMPUTED("getDCMVelocityTrajectory")
    out.resize(m_samples);
    for (int s = 0; s < m_samples; s++) { out[s](0) = m_dcmVel[2 * s]; out[s](1) = m_dcmVel[2 * s + 1]; }
    return true;
}

bool TrajectoryGenerator::getFeetTrajectories(std::vector<iDynTree::Transform>& l, std::vector<iDynTree::Transform>& r)
{
    WC_REQUIRE_COMPUTED("getFeetTrajectories")
    l.resize(m_samples); r.resize(m_samples);
    for (int s = 0; s < m_samples; s++) { l[s] = wc_compat::fromH12(&m_lfH[12 * s]); r[s] = wc_compat::fromH12(&m_rfH[12 * s]); }
    return true;
}

bool TrajectoryGenerator::getFeetTwist(std::vector<iDynTree::Twist>& l, std::vector<iDynTree::Twist>& r)
{
    WC_REQUIRE_COMPUTED("getFeetTwist")
    l.resize(m_samples); r.resize(m_samples);
    for (int s = 0; s < m_samples; s++)
        for (int e = 0; e < 6; e++) { l[s](e) = m_lfTw[6 * s + e]; r[s](e) = m_rfTw[6 * s + e]; }
    return true;
}

bool TrajectoryGenerator::getWhenUseLeftAsFixed(std::vector<bool>& out)
{
    WC_REQUIRE_COMPUTED("getWhenUseLeftAsFixed")
    out.assign(m_lfix.begin(), m_lfix.end());
    return true;
}

bool TrajectoryGenerator::getFeetStandingPeriods(std::vector<bool>& l, std::vector<bool>& r)
{
    WC_REQUIRE_COMPUTED("getFeetStandingPeriods")
    l.assign(m_lc.begin(), m_lc.end()); r.assign(m_rc.begin(), m_rc.end());
    return true;
}

bool TrajectoryGenerator::getCoMHeightTrajectory(std::vector<double>& out)
{
    WC_REQUIRE_COMPUTED("getCoMHeightTrajectory")
    out = m_comH;
    return true;
}

bool TrajectoryGenerator::getCoMHeightVelocity(std::vector<double>& out)
{
    WC_REQUIRE_COMPUTED("getCoMHeightVelocity")
    out = m_comHV;
    return true;
}
