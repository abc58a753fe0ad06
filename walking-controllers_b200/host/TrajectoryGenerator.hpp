// Host-side stand-in for the reference's TrajectoryGenerator (modules/Walking_module/include/TrajectoryGenerator.hpp) with
// the getters WalkingModule::updateTrajectories reads (WalkingModule.cpp:1107-1121; TrajectoryGenerator.cpp:501-598).
// NOT a drop-in for the planner: the reference wraps the UnicyclePlanner library (external), which plans footsteps towards a
// joypad goal; this class produces the periodic gait of the synthetic workload from the same
// plannerParams.ini values, on the GPU through wc_traj_generate_host (include/wc_capi.h).  Gait shape (step length, heading
// rate, initial heading, phase) replaces the unicycle's desired position; there is no asynchronous thread: trajectories are
// computed when generateFirstTrajectories / updateTrajectories is called.
#ifndef WC_HOST_TRAJECTORY_GENERATOR_HPP
#define WC_HOST_TRAJECTORY_GENERATOR_HPP

#include <vector>

#include "../../include/wc_capi.h"
#include "wc_compat.hpp"

class TrajectoryGenerator {
    wc_traj* m_handle{nullptr};
    wc_traj_config m_cfg{};
    double m_dT{0.01}, m_plannerHorizon{10.0};
    double m_stepLength{0.1}, m_headingRate{0.0}, m_heading{0.0};
    bool m_swingLeft{true}, m_computed{false};
    int m_samples{0};
    std::vector<double> m_dcmPos, m_dcmVel, m_lfH, m_rfH, m_lfTw, m_rfTw, m_comH, m_comHV;
    std::vector<uint8_t> m_lc, m_rc, m_lfix;

    bool generate(int firstSample);

public:
    TrajectoryGenerator() = default;
    ~TrajectoryGenerator();
    TrajectoryGenerator(const TrajectoryGenerator&) = delete;
    TrajectoryGenerator& operator=(const TrajectoryGenerator&) = delete;

    bool initialize(const yarp::os::Searchable& config);       // = configurePlanner (TrajectoryGenerator.cpp:33-42)
    bool configurePlanner(const yarp::os::Searchable& config); // :44-171
    // the periodic gait that stands in for the unicycle goal (metres per step, radians per step, radians)
    void setGait(double stepLength, double headingRatePerStep, double initialHeading);
    bool generateFirstTrajectories();                          // :232-283: plannerHorizon seconds from tick 0
    bool updateTrajectories(int firstTick);                    // :285-330 (merge point = firstTick): plannerHorizon seconds from there
    bool isTrajectoryComputed() const { return m_computed; }

    bool getDCMPositionTrajectory(std::vector<iDynTree::Vector2>& DCMPositionTrajectory);
    bool getDCMVelocityTrajectory(std::vector<iDynTree::Vector2>& DCMVelocityTrajectory);
    bool getFeetTrajectories(std::vector<iDynTree::Transform>& lFootTrajectory, std::vector<iDynTree::Transform>& rFootTrajectory);
    bool getFeetTwist(std::vector<iDynTree::Twist>& lFootTwist, std::vector<iDynTree::Twist>& rFootTwist);
    bool getWhenUseLeftAsFixed(std::vector<bool>& isLeftFixedFrame);
    bool getFeetStandingPeriods(std::vector<bool>& lFootContacts, std::vector<bool>& rFootContacts);
    bool getCoMHeightTrajectory(std::vector<double>& CoMHeightTrajectory);
    bool getCoMHeightVelocity(std::vector<double>& CoMHeightVelocity);
    void reset() { m_computed = false; }
};

#endif
