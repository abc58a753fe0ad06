// Drop-in for the reference's WalkingDCMReactiveController
// (modules/Walking_module/include/WalkingDCMReactiveController.hpp): same method names, argument meaning, bool return
// convention and evaluateControl / getControllerOutput latch; the control law runs in the CUDA library behind the C ABI
// (wc_glue_reactive_dcm_host, include/wc_capi.h) with a batch of one instance.
#ifndef WC_HOST_WALKING_DCM_REACTIVE_CONTROLLER_HPP
#define WC_HOST_WALKING_DCM_REACTIVE_CONTROLLER_HPP

#include "../../include/wc_capi.h"
#include "wc_compat.hpp"

class WalkingDCMReactiveController {
    wc_glue* m_handle{nullptr};
    double m_kDCM{0.0};
    bool m_isInitialized{false};
    bool m_controlEvaluated{false};
    iDynTree::Vector2 m_dcmFeedback, m_dcmPositionDesired, m_dcmVelocityDesired, m_controllerOutput;

public:
    WalkingDCMReactiveController() = default;
    ~WalkingDCMReactiveController();
    WalkingDCMReactiveController(const WalkingDCMReactiveController&) = delete;
    WalkingDCMReactiveController& operator=(const WalkingDCMReactiveController&) = delete;

    bool initialize(const yarp::os::Searchable& config);
    void setFeedback(const iDynTree::Vector2& dcmFeedback);
    void setReferenceSignal(const iDynTree::Vector2& dcmPositionDesired, const iDynTree::Vector2& dcmVelocityDesired);
    bool evaluateControl();
    bool getControllerOutput(iDynTree::Vector2& controllerOutput);
};

#endif
