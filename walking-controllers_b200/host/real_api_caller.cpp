// Caller of the drop-in classes written ONLY against genuine iDynTree / YARP signatures (built by `make check_real_api` with
// -DWC_USE_REAL_IDYNTREE -Imock_include; with the real libraries installed, the same file compiles against them).  It walks the
// call sequence of WalkingModule::updateModule (WM/src/WalkingModule.cpp:601-745): configuration from yarp::os::Property,
// setConvexHullConstraint -> setFeedback -> setReferenceSignal -> solve -> getControllerOutput, then the QP-IK setters / solve /
// getSolution.  Run on a GPU box it performs one tick (exit code 0 = every call returned true); without a device it stops at
// initialize(), which is all the build check needs.
#include <cstdio>
#include <deque>

#include <iDynTree/Core/MatrixDynSize.h>
#include <iDynTree/Core/Transform.h>
#include <iDynTree/Core/Twist.h>
#include <iDynTree/Core/VectorDynSize.h>
#include <yarp/os/Bottle.h>
#include <yarp/os/Property.h>
#include <yarp/os/Value.h>

#include "WalkingDCMModelPredictiveController.hpp"
#include "WalkingDCMReactiveController.hpp"
#include "WalkingForwardKinematics.hpp"
#include "WalkingLogger.hpp"
#include "WalkingQPInverseKinematics.hpp"

static yarp::os::Value* list_of(std::initializer_list<double> v)
{
    yarp::os::Value* l = yarp::os::Value::makeList();
    for (double x : v) l->asList()->addDouble(x);
    return l;
}

int main()
{
    yarp::os::Property mpc;
    mpc.put("sampling_time", yarp::os::Value(0.01));
    mpc.put("controllerHorizon", yarp::os::Value(0.25));
    mpc.put("com_height", yarp::os::Value(0.53));
    mpc.put("gravity_acceleration", yarp::os::Value(9.81));
    mpc.put("convex_hull_tolerance", yarp::os::Value(0.05));
    mpc.put("initial_zmp_position", list_of({0.0, 0.0}));
    {
        // ((0 0 7500) (1 1 7500)): a list of triplets (row, col, value)
        yarp::os::Value* q = yarp::os::Value::makeList();
        yarp::os::Bottle& a = q->asList()->addList(); a.addInt(0); a.addInt(0); a.addDouble(7500.0);
        yarp::os::Bottle& b = q->asList()->addList(); b.addInt(1); b.addInt(1); b.addDouble(7500.0);
        mpc.put("stateWeightTriplets", q);
        yarp::os::Value* r = yarp::os::Value::makeList();
        yarp::os::Bottle& c = r->asList()->addList(); c.addInt(0); c.addInt(0); c.addDouble(9e6);
        yarp::os::Bottle& d = r->asList()->addList(); d.addInt(1); d.addInt(1); d.addDouble(9e6);
        mpc.put("inputWeightTriplets", r);
        yarp::os::Value* fs = yarp::os::Value::makeList();
        yarp::os::Bottle& fx = fs->asList()->addList(); fx.addDouble(-0.02); fx.addDouble(0.05);
        yarp::os::Bottle& fy = fs->asList()->addList(); fy.addDouble(-0.025); fy.addDouble(0.025);
        mpc.put("foot_size", fs);
    }
    WalkingController controller;
    if (!controller.initialize(mpc)) { std::printf("WalkingController::initialize returned false (no CUDA device?)\n"); return 2; }

    std::deque<iDynTree::Transform> lf, rf;
    iDynTree::Transform L(iDynTree::Rotation::Identity(), iDynTree::Position(0.0, 0.07, 0.0));
    iDynTree::Transform R(iDynTree::Rotation::RotZ(0.05), iDynTree::Position(0.0, -0.07, 0.0));
    lf.push_back(L); rf.push_back(R);
    std::deque<bool> lc(1, true), rc(1, true);
    bool ok = controller.setConvexHullConstraint(lf, rf, lc, rc);
    iDynTree::Vector2 dcm; dcm(0) = 0.01; dcm(1) = 0.0;
    ok = ok && controller.setFeedback(dcm);
    std::deque<iDynTree::Vector2> ref(26);
    for (auto& v : ref) { v(0) = 0.0; v(1) = 0.0; }
    ok = ok && controller.setReferenceSignal(ref, true) && controller.solve();
    iDynTree::Vector2 zmp;
    ok = ok && controller.getControllerOutput(zmp);
    std::printf("MPC: ok=%d zmp=(%g, %g)\n", (int)ok, zmp(0), zmp(1));

    yarp::os::Property ikc;
    ikc.put("useCoMAsConstraint", yarp::os::Value(1));
    ikc.put("neckWeightTriplets", list_of({0, 0, 5, 1, 1, 5, 2, 2, 5}));
    ikc.put("comWeightTriplets", list_of({0, 0, 100, 1, 1, 100, 2, 2, 100}));
    ikc.put("additional_rotation", list_of({0, 0, 1, 1, 0, 0, 0, 1, 0}));
    {
        yarp::os::Value* v = yarp::os::Value::makeList(); yarp::os::Value* w = yarp::os::Value::makeList(); yarp::os::Value* g = yarp::os::Value::makeList();
        for (int i = 0; i < 23; i++) { v->asList()->addDouble(0.0); w->asList()->addDouble(1.0); g->asList()->addDouble(5.0); }
        ikc.put("jointRegularization", v); ikc.put("jointRegularizationWeights", w); ikc.put("jointRegularizationGains", g);
    }
    ikc.put("k_posFoot", yarp::os::Value(4.0)); ikc.put("k_attFoot", yarp::os::Value(2.0)); ikc.put("k_neck", yarp::os::Value(1.0));
    ikc.put("k_posCom", yarp::os::Value(2.0));
    iDynTree::VectorDynSize vmin(23), vmax(23);
    for (unsigned i = 0; i < 23; i++) { vmin(i) = -1.0; vmax(i) = 1.0; }
    WalkingQPIK_qpOASES ik;
    ok = ik.initialize(ikc, 23, vmin, vmax) && ok;
    iDynTree::VectorDynSize q(23), dq(23);
    iDynTree::MatrixDynSize Jl(6, 29), Jr(6, 29), Jn(6, 29), Jc(3, 29);
    for (unsigned r = 0; r < 6; r++) { Jl(r, r) = 1.0; Jr(r, r) = 1.0; Jr(r, 6 + 12 + r) = 0.3; Jn(r, r) = 1.0; }
    for (unsigned r = 0; r < 3; r++) { Jc(r, r) = 1.0; Jc(r, 6 + r) = 0.1; Jn(3 + r, 6 + r) = 1.0; }
    ok = ik.setRobotState(q, L, R, iDynTree::Rotation::Identity(), iDynTree::Position(0.0, 0.0, 0.5)) && ok;
    ok = ik.setLeftFootJacobian(Jl) && ik.setRightFootJacobian(Jr) && ik.setNeckJacobian(Jn) && ik.setCoMJacobian(Jc) && ok;
    ok = ik.setDesiredJointPosition(q) && ok;
    iDynTree::Twist lt, rt; lt.zero(); rt.zero(); rt(0) = 0.02;
    ik.setDesiredFeetTwist(lt, rt);
    ik.setDesiredFeetTransformation(L, R);
    iDynTree::Vector3 cv; cv.zero();
    ik.setDesiredCoMVelocity(cv);
    ik.setDesiredCoMPosition(iDynTree::Position(0.0, 0.0, 0.5));
    ik.setDesiredNeckOrientation(iDynTree::Rotation::Identity());
    ok = ik.solve() && ik.getSolution(dq) && ok;
    std::printf("QP-IK: ok=%d dq[12]=%g\n", (int)ok, dq(12));

    yarp::os::Property rc_;
    rc_.put("kDCM", yarp::os::Value(4.0)); rc_.put("com_height", yarp::os::Value(0.53)); rc_.put("gravity_acceleration", yarp::os::Value(9.81));
    WalkingDCMReactiveController reactive;
    ok = reactive.initialize(rc_) && ok;
    reactive.setFeedback(dcm);
    iDynTree::Vector2 z2; z2.zero();
    reactive.setReferenceSignal(z2, z2);
    ok = reactive.evaluateControl() && reactive.getControllerOutput(zmp) && ok;
    return ok ? 0 : 1;
}
