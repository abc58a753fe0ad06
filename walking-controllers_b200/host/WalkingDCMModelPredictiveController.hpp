// Drop-in for the reference's WalkingController
// (modules/Walking_module/include/WalkingDCMModelPredictiveController.hpp:190-250): same method names,
// argument meaning, bool return convention and getControllerOutput latch; the work is done by the
// CUDA library behind the C ABI (include/wc_capi.h) with a batch of one instance.
#ifndef WC_HOST_WALKING_CONTROLLER_HPP
#define WC_HOST_WALKING_CONTROLLER_HPP

#include <deque>
#include <utility>
#include <vector>

#include "../../include/wc_capi.h"
#include "wc_compat.hpp"

class WalkingController {
    wc_mpc* m_handle{nullptr};
    wc_mpc_config m_cfg{};
    std::pair<bool, bool> m_feetStatus{false, false};
    bool m_hasSolver{false};
    bool m_isSolutionEvaluated{false};
    iDynTree::Vector2 m_output;
    // inputs buffered by the setters until solve()
    double m_lfH[12]{}, m_rfH[12]{}, m_dcm[2]{};
    uint8_t m_lc{0}, m_rc{0}, m_reset{0};
    std::vector<double> m_ref;
    int32_t m_lastStatus{WC_OSQP_UNSOLVED}, m_lastIters{0}, m_lastFlags{0};

public:
    WalkingController() = default;
    ~WalkingController();
    WalkingController(const WalkingController&) = delete;
    WalkingController& operator=(const WalkingController&) = delete;

    bool initialize(const yarp::os::Searchable& config);
    bool setConvexHullConstraint(const std::deque<iDynTree::Transform>& leftFoot, const std::deque<iDynTree::Transform>& rightFoot,
                                 const std::deque<bool>& leftInContact, const std::deque<bool>& rightInContact);
    bool setFeedback(const iDynTree::Vector2& currentState);
    bool setReferenceSignal(const std::deque<iDynTree::Vector2>& referenceSignal, const bool& resetTrajectory);
    bool solve();
    bool getControllerOutput(iDynTree::Vector2& controllerOutput);
    void reset();

    // extras (not in the reference): integer solver information of the last solve
    int32_t osqpStatus() const { return m_lastStatus; }
    int32_t osqpIterations() const { return m_lastIters; }
    int32_t flags() const { return m_lastFlags; }
    int horizon() const { return m_cfg.horizon; }
};

#endif
