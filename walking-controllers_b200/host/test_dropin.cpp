// test_dropin.cpp — exercises the drop-in C++ classes exactly in the reference's per-tick call order
// (WalkingModule::updateModule, modules/Walking_module/src/WalkingModule.cpp:601-637 and :709-745 with
// solveQPIK :367-425) for one robot on a synthetic straight walk, and checks every tick against the CPU
// oracle (TEST ONLY: the oracle is linked into this test binary, never into the product library).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <string>
#include <vector>

#include "../../oracle/wc_oracle.h"
#include "WalkingDCMModelPredictiveController.hpp"
#include "WalkingDCMReactiveController.hpp"
#include "WalkingLogger.hpp"
#include "TrajectoryGenerator.hpp"
#include "WalkingForwardKinematics.hpp"
#include "WalkingQPInverseKinematics.hpp"

static iDynTree::Transform planar(double x, double y, double z, double yaw)
{
    iDynTree::Position p; p(0) = x; p(1) = y; p(2) = z;
    return iDynTree::Transform(iDynTree::Rotation::RotZ(yaw), p);
}

int main(int argc, char** argv)
{
    const int N = argc > 1 ? std::atoi(argv[1]) : 25;
    const int T = argc > 2 ? std::atoi(argv[2]) : 200;
    const double dT = 0.01;
    // ---- configuration: the values of app/robots/iCubGazeboV2_5/*.ini
    yarp::os::Searchable mpcCfg, ikCfg, fkCfg;
    mpcCfg.put("sampling_time", dT); mpcCfg.put("controllerHorizon", N * dT); mpcCfg.put("com_height", 0.53);
    mpcCfg.put("stateWeightTriplets", std::vector<double>{0, 0, 7500, 1, 1, 7500});
    mpcCfg.put("inputWeightTriplets", std::vector<double>{0, 0, 9000000, 1, 1, 9000000});
    mpcCfg.put("foot_size", std::vector<double>{-0.02, 0.05, -0.025, 0.025});
    mpcCfg.put("initial_zmp_position", std::vector<double>{0.0, 0.0});
    mpcCfg.put("convex_hull_tolerance", 0.05);
    ikCfg.put("useCoMAsConstraint", 1.0);
    ikCfg.put("neckWeightTriplets", std::vector<double>{0, 0, 5, 1, 1, 5, 2, 2, 5});
    ikCfg.put("additional_rotation", std::vector<double>{0, 0, 1, 1, 0, 0, 0, 1, 0});
    ikCfg.put("jointRegularization", std::vector<double>{15, 0, 0, -7, 22, 11, 30, -7, 22, 11, 30, 5.082, 0.406, -0.131, -45.249,
                                                         -26.454, -0.351, 5.082, 0.406, -0.131, -45.249, -26.454, -0.351});
    ikCfg.put("jointRegularizationWeights", std::vector<double>{1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1});
    ikCfg.put("jointRegularizationGains", std::vector<double>(23, 5.0));
    ikCfg.put("k_posCom", 1.0); ikCfg.put("k_posFoot", 4.0); ikCfg.put("k_attFoot", 2.0); ikCfg.put("k_neck", 1.0);
    fkCfg.put("left_foot_frame", std::string("l_sole"));

    WalkingController controller;
    WalkingQPIK_qpOASES ik;
    WalkingQPIK_osqp ikOsqp;            // `use_osqp on` (iCubGenova04 / iCubGazeboV2_5 dcmWalkingCoordinator.ini)
    WalkingFK fk;
    wc_model model; wc_model_icub_like(&model);
    iDynTree::VectorDynSize vmin(23), vmax(23);
    for (int i = 0; i < 23; i++) { vmin(i) = -20.0; vmax(i) = 20.0; } // wide limits: the synthetic swing-foot targets start far from the posture
    if (!controller.initialize(mpcCfg) || !ik.initialize(ikCfg, 23, vmin, vmax) || !ikOsqp.initialize(ikCfg, 23, vmin, vmax) || !fk.initialize(fkCfg, model)) {
        std::fprintf(stderr, "initialisation failed\n"); return 2;
    }
    // ---- oracle twins
    wc_mpc_config ocfg; wc_mpc_default_config(&ocfg, N);
    wco_mpc* ompc = wco_mpc_create(&ocfg);
    wc_ik_config oik; wc_ik_default_config(&oik);
    for (int i = 0; i < 23; i++) { oik.min_vel[i] = -20.0; oik.max_vel[i] = 20.0; }

    // ---- synthetic straight walk: step 0.1 m every 90 ticks, 37 ticks of double support
    const int STEP = 90, DS = 37;
    auto foot_x = [&](int k) { return 0.05 * k; };
    auto foot_y = [&](int k) { return (k % 2 == 0) ? 0.08 : -0.08; };
    const double omega = std::sqrt(9.81 / 0.53);
    auto dcm_ref = [&](int s, double* r) { // smooth path between successive footstep ZMPs
        const int k = s / STEP; const double tau = (s % STEP) / double(STEP);
        const double w = tau * tau * (3 - 2 * tau);
        r[0] = foot_x(k) + 0.03 + w * (foot_x(k + 1) - foot_x(k));
        r[1] = foot_y(k) + w * (foot_y(k + 1) - foot_y(k));
        (void)omega;
    };
    std::deque<iDynTree::Vector2> dcmDeque;
    for (int s = 0; s <= N; s++) { double r[2]; dcm_ref(s, r); iDynTree::Vector2 v; v(0) = r[0]; v(1) = r[1]; dcmDeque.push_back(v); }

    iDynTree::VectorDynSize qDesired(23), dqDesired(23), qMeas(23);
    for (int i = 0; i < 23; i++) qDesired(i) = oik.joint_reg_rad[i];
    double worst_zmp = 0.0, worst_dq = 0.0, worst_dq_osqp = 0.0, worst_fk_dcm = 0.0;
    wco_ikosqp* oikOsqp = wco_ikosqp_create(&oik, nullptr);
    int osqpSetups = 0;
    int failures = 0;
    for (int t = 0; t < T; t++) {
        const int k = t / STEP, i = t % STEP;
        const bool ds = i < DS, stanceLeft = (k % 2 == 0);
        const double u = ds ? 0.0 : (i - DS) / double(STEP - DS), bl = u * u * (3 - 2 * u);
        const double ox = foot_x(k - 1) + bl * (foot_x(k + 1) - foot_x(k - 1)), oz = ds ? 0.0 : 4 * 0.025 * u * (1 - u);
        iDynTree::Transform stance = planar(foot_x(k), foot_y(k), 0, 0), other = planar(ox, foot_y(k - 1), oz, 0);
        std::deque<iDynTree::Transform> leftTraj{stanceLeft ? stance : other}, rightTraj{stanceLeft ? other : stance};
        std::deque<bool> leftInContact{stanceLeft || ds}, rightInContact{!stanceLeft || ds};
        iDynTree::Vector2 measuredDCM; measuredDCM(0) = dcmDeque.front()(0) + 0.002 * std::sin(0.37 * t);
        measuredDCM(1) = dcmDeque.front()(1) + 0.002 * std::cos(0.21 * t);

        // ======== MPC block (WalkingModule.cpp:605-634)
        iDynTree::Vector2 desiredZMP;
        bool ok = controller.setConvexHullConstraint(leftTraj, rightTraj, leftInContact, rightInContact) &&
                  controller.setFeedback(measuredDCM) && controller.setReferenceSignal(dcmDeque, false) && controller.solve() &&
                  controller.getControllerOutput(desiredZMP);
        double lfH[12], rfH[12], dcm[2] = {measuredDCM(0), measuredDCM(1)}, zo[2];
        wc_compat::toH12(leftTraj.front(), lfH); wc_compat::toH12(rightTraj.front(), rfH);
        std::vector<double> ref(2 * (N + 1));
        for (int s = 0; s <= N; s++) { ref[2 * s] = dcmDeque[s](0); ref[2 * s + 1] = dcmDeque[s](1); }
        int ost = 0, oit = 0;
        const int ofl = wco_mpc_tick(ompc, lfH, rfH, leftInContact.front(), rightInContact.front(), dcm, ref.data(), 0, zo, &ost, &oit);
        const bool ook = (ofl & ~WC_MPC_FLAG_NEW_SOLVER) == 0;
        if (ok != ook || controller.osqpStatus() != ost || controller.osqpIterations() != oit) {
            std::fprintf(stderr, "tick %d: MPC status mismatch (ok %d/%d status %d/%d iters %d/%d)\n", t, ok, ook, controller.osqpStatus(), ost,
                         controller.osqpIterations(), oit);
            failures++;
        }
        for (int a = 0; a < 2; a++) worst_zmp = std::fmax(worst_zmp, std::fabs(desiredZMP(a) - zo[a]) / std::fmax(1.0, std::fabs(zo[a])));
        iDynTree::Vector2 dummy;
        if (controller.getControllerOutput(dummy)) { std::fprintf(stderr, "tick %d: output latch did not reset\n", t); failures++; }

        // ======== QP-IK block (WalkingModule.cpp:715-744, solveQPIK :367-425)
        for (int j = 0; j < 23; j++) qMeas(j) = qDesired(j) + 0.003 * std::sin(0.1 * t + j);
        ok = fk.evaluateWorldToBaseTransformation(leftTraj.front(), rightTraj.front(), stanceLeft) && fk.setInternalRobotState(qDesired, dqDesired);
        iDynTree::Position comMeas; fk.getCoMPosition(comMeas);
        { // WalkingModule::evaluateCoM / evaluateDCM (WalkingModule.cpp:559-570): the measured DCM of the MPC feedback
            iDynTree::Vector2 dcmFk; iDynTree::Vector3 comVelFk; iDynTree::MatrixDynSize jc;
            if (!fk.evaluateCoM() || !fk.evaluateDCM() || !fk.getDCM(dcmFk) || !fk.getCoMVelocity(comVelFk) || !fk.getCoMJacobian(jc)) {
                std::fprintf(stderr, "tick %d: evaluateCoM / evaluateDCM failed\n", t); failures++;
            } else {
                for (int a = 0; a < 2; a++) {
                    double v = 0.0;
                    for (int j = 0; j < 23; j++) v += jc(a, 6 + j) * dqDesired(j);
                    worst_fk_dcm = std::fmax(worst_fk_dcm, std::fabs(dcmFk(a) - (comMeas(a) + v / omega)));
                }
            }
        }
        iDynTree::Position comDes = comMeas; comDes(0) += 0.002; comDes(1) -= 0.001;
        iDynTree::Vector3 comVel; comVel(0) = 0.02; comVel(1) = 0.01 * std::sin(0.05 * t); comVel(2) = 0.0;
        iDynTree::Twist lTw, rTw; lTw.zero(); rTw.zero();
        if (!ds) { (stanceLeft ? rTw : lTw)(0) = 0.1 / ((STEP - DS) * dT) * 6 * u * (1 - u); (stanceLeft ? rTw : lTw)(2) = 4 * 0.025 * (1 - 2 * u) / ((STEP - DS) * dT); }
        const iDynTree::Rotation yawRotation = iDynTree::Rotation::RotZ(0.0).inverse();
        ok = ok && ik.setRobotState(qMeas, fk.getLeftFootToWorldTransform(), fk.getRightFootToWorldTransform(), fk.getNeckOrientation(), comMeas);
        ik.setDesiredNeckOrientation(yawRotation.inverse());
        ik.setDesiredFeetTransformation(leftTraj.front(), rightTraj.front());
        ik.setDesiredFeetTwist(lTw, rTw);
        ik.setDesiredCoMVelocity(comVel);
        ik.setDesiredCoMPosition(comDes);
        iDynTree::MatrixDynSize jac, comJac;
        fk.getLeftFootJacobian(jac); ok = ok && ik.setLeftFootJacobian(jac);
        fk.getRightFootJacobian(jac); ok = ok && ik.setRightFootJacobian(jac);
        fk.getNeckJacobian(jac); ok = ok && ik.setNeckJacobian(jac);
        fk.getCoMJacobian(comJac); ok = ok && ik.setCoMJacobian(comJac);
        ok = ok && ik.solve() && ik.getSolution(dqDesired);
        // oracle
        double bH[12], olf[12], orf[12], onR[9], ocom[3], Jl[6 * 29], Jr[6 * 29], Jn[6 * 29], Jc[3 * 29], odq[23], ole[6], ore[6];
        wc_compat::toH12((stanceLeft ? leftTraj : rightTraj).front(), bH);
        wco_fk(&model, qDesired.data(), bH, stanceLeft, olf, orf, onR, ocom, Jl, Jr, Jn, Jc);
        double lfd[12], rfd[12]; wc_compat::toH12(leftTraj.front(), lfd); wc_compat::toH12(rightTraj.front(), rfd);
        int onw = 0;
        double lTw6[6], rTw6[6]; wc_compat::twist6(lTw, lTw6); wc_compat::twist6(rTw, rTw6);
        const int ost2 = wco_ik_solve(&oik, qMeas.data(), olf, orf, onR, comMeas.data(), Jl, Jr, Jn, Jc, lfd, rfd, lTw6, rTw6,
                                      comDes.data(), comVel.data(), yawRotation.inverse().data(), odq, ole, ore, &onw, nullptr, nullptr,
                                      nullptr, nullptr, nullptr);
        if (ok != (ost2 == 0) || ik.qpStatus() != ost2) { std::fprintf(stderr, "tick %d: IK status mismatch %d/%d\n", t, ik.qpStatus(), ost2); failures++; }
        for (int j = 0; j < 23; j++) worst_dq = std::fmax(worst_dq, std::fabs(dqDesired(j) - odq[j]));
        // ======== the same tick through WalkingQPIK_osqp (WalkingModule.cpp:721-731), against its restatement
        {
            iDynTree::VectorDynSize dqOsqp(23);
            bool ok2 = ikOsqp.setRobotState(qMeas, fk.getLeftFootToWorldTransform(), fk.getRightFootToWorldTransform(), fk.getNeckOrientation(), comMeas);
            ikOsqp.setDesiredNeckOrientation(yawRotation.inverse());
            ikOsqp.setDesiredFeetTransformation(leftTraj.front(), rightTraj.front());
            ikOsqp.setDesiredFeetTwist(lTw, rTw);
            ikOsqp.setDesiredCoMVelocity(comVel);
            ikOsqp.setDesiredCoMPosition(comDes);
            // The restatement gets the very Jacobians the class is given (the WalkingFK getters, i.e. the device FK): which entries
            // of the Hessian are EXACTLY zero decides whether osqp-eigen re-initialises the solver, so parity of the solver life
            // cycle is only defined on bit-identical inputs (noise-level entries flip between 0 and 1e-17 as the posture moves).
            iDynTree::MatrixDynSize jl, jr, jn;
            fk.getLeftFootJacobian(jl); ok2 = ok2 && ikOsqp.setLeftFootJacobian(jl);
            fk.getRightFootJacobian(jr); ok2 = ok2 && ikOsqp.setRightFootJacobian(jr);
            fk.getNeckJacobian(jn); ok2 = ok2 && ikOsqp.setNeckJacobian(jn);
            ok2 = ok2 && ikOsqp.setCoMJacobian(comJac);
            ok2 = ok2 && ikOsqp.solve() && ikOsqp.getSolution(dqOsqp);
            double odq2[23], ole2[6], ore2[6], glf[12], grf[12]; int ost3 = 0, oit3 = 0, ofl3 = 0;
            wc_compat::toH12(fk.getLeftFootToWorldTransform(), glf); wc_compat::toH12(fk.getRightFootToWorldTransform(), grf);
            wco_ikosqp_solve(oikOsqp, qMeas.data(), glf, grf, fk.getNeckOrientation().data(), comMeas.data(), jl.data(), jr.data(), jn.data(),
                             comJac.data(), lfd, rfd, lTw6, rTw6, comDes.data(), comVel.data(), yawRotation.inverse().data(),
                             odq2, ole2, ore2, &ost3, &oit3, &ofl3);
            if (ok2 != (ofl3 == 0) || ikOsqp.osqpStatus() != ost3 || ikOsqp.osqpIterations() != oit3 || ikOsqp.solverSetups() != wco_ikosqp_setups(oikOsqp)) {
                std::fprintf(stderr, "tick %d: osqp IK mismatch (ok %d/%d status %d/%d iters %d/%d setups %d/%d)\n", t, ok2, ofl3 == 0,
                             ikOsqp.osqpStatus(), ost3, ikOsqp.osqpIterations(), oit3, ikOsqp.solverSetups(), wco_ikosqp_setups(oikOsqp));
                failures++;
            }
            for (int j = 0; j < 23; j++) worst_dq_osqp = std::fmax(worst_dq_osqp, std::fabs(dqOsqp(j) - odq2[j]));
            osqpSetups = ikOsqp.solverSetups();
        }
        for (int j = 0; j < 23; j++) qDesired(j) += dT * dqDesired(j);   // stand-in for iCub::ctrl::Integrator (:741-744)

        // advanceReferenceSignals (WalkingModule.cpp:35-96)
        dcmDeque.pop_front();
        { double r[2]; dcm_ref(t + N + 1, r); iDynTree::Vector2 v; v(0) = r[0]; v(1) = r[1]; dcmDeque.push_back(v); }
    }
    // ======== use_mpc off: WalkingDCMReactiveController (WalkingModule.cpp:638-655) + the logger row (:793-803)
    double worst_reactive = 0.0;
    {
        yarp::os::Searchable rcCfg, logCfg;
        rcCfg.put("kDCM", 1.5); rcCfg.put("com_height", 0.53);
        WalkingDCMReactiveController reactive;
        iDynTree::Vector2 out;
        if (reactive.evaluateControl() || reactive.getControllerOutput(out)) { std::fprintf(stderr, "reactive: latches\n"); failures++; }
        if (!reactive.initialize(rcCfg)) { std::fprintf(stderr, "reactive: initialize failed\n"); failures++; }
        logCfg.put("sampling_time", dT); logCfg.put("file_name", std::string("/tmp/wc_dropin_dataset.txt"));
        WalkingLogger logger;
        bool lok = logger.configure(logCfg, "test") && logger.startRecord({"record", "dcm_x", "dcm_y", "dcm_des_x", "dcm_des_y", "zmp_des_x", "zmp_des_y"});
        for (int t = 0; t < 20; t++) {
            iDynTree::Vector2 xd, vd, x;
            double r0[2], r1[2]; dcm_ref(t, r0); dcm_ref(t + 1, r1);
            for (int a = 0; a < 2; a++) { xd(a) = r0[a]; vd(a) = (r1[a] - r0[a]) / dT; x(a) = r0[a] + 0.002 * std::sin(0.3 * t + a); }
            reactive.setFeedback(x); reactive.setReferenceSignal(xd, vd);
            if (!reactive.evaluateControl() || !reactive.getControllerOutput(out)) { std::fprintf(stderr, "reactive: tick %d failed\n", t); failures++; continue; }
            double want[2];
            wco_reactive_dcm(1, omega, 1.5, nullptr, xd.data(), vd.data(), x.data(), want);
            for (int a = 0; a < 2; a++) worst_reactive = std::fmax(worst_reactive, std::fabs(out(a) - want[a]));
            logger.sendData(x, xd, out);
        }
        logger.quit();
        std::ifstream f(logger.fileName()); std::string line; int rows = 0;
        std::getline(f, line);
        if (!lok || line != "time dcm_x dcm_y dcm_des_x dcm_des_y zmp_des_x zmp_des_y ") { std::fprintf(stderr, "logger: header '%s'\n", line.c_str()); failures++; }
        while (std::getline(f, line)) rows++;
        if (rows != 20) { std::fprintf(stderr, "logger: %d rows\n", rows); failures++; }
    }
    // ======== TrajectoryGenerator stand-in (periodic gait on the device) against the CPU twin
    double worst_traj = 0.0;
    {
        yarp::os::Searchable plCfg;
        plCfg.put("sampling_time", dT); plCfg.put("plannerHorizon", 3.0); plCfg.put("nominalDuration", 0.9); plCfg.put("switchOverSwingRatio", 0.7);
        plCfg.put("nominalWidth", 0.16); plCfg.put("stepHeight", 0.025); plCfg.put("com_height", 0.53);
        plCfg.put("leftZMPDelta", std::vector<double>{0.03, -0.005}); plCfg.put("rightZMPDelta", std::vector<double>{0.03, 0.005});
        TrajectoryGenerator gen;
        std::vector<iDynTree::Vector2> dcmPos;
        if (gen.getDCMPositionTrajectory(dcmPos)) { std::fprintf(stderr, "trajectory: getter before generation\n"); failures++; }
        if (!gen.initialize(plCfg)) { std::fprintf(stderr, "trajectory: initialize failed\n"); failures++; }
        gen.setGait(0.12, 0.05, 0.3);
        std::vector<iDynTree::Transform> lT, rT; std::vector<bool> lC, rC, lFix; std::vector<double> comH;
        if (!gen.generateFirstTrajectories() || !gen.getDCMPositionTrajectory(dcmPos) || !gen.getFeetTrajectories(lT, rT) ||
            !gen.getFeetStandingPeriods(lC, rC) || !gen.getWhenUseLeftAsFixed(lFix) || !gen.getCoMHeightTrajectory(comH)) {
            std::fprintf(stderr, "trajectory: generation failed\n"); failures++;
        } else {
            wc_traj_config tc; wc_traj_default_config(&tc, 300, 0);
            wco_traj* tw = wco_traj_create(&tc, 1);
            const int32_t ph = 0; const double L = 0.12, dth = 0.05, th0 = 0.3; const uint8_t fl = 0;
            wco_traj_set_gait(tw, &ph, &L, &dth, &th0, &fl);
            std::vector<double> pos(2 * 300);
            wco_traj_dcm(tw, 0, 300, pos.data(), nullptr);
            if (dcmPos.size() != 300 || lT.size() != 300) { std::fprintf(stderr, "trajectory: %zu samples\n", dcmPos.size()); failures++; }
            for (size_t s = 0; s < dcmPos.size() && s < 300; s++) {
                for (int a = 0; a < 2; a++) worst_traj = std::fmax(worst_traj, std::fabs(dcmPos[s](a) - pos[2 * s + a]));
                double lf[12], rf[12], got[12]; uint8_t lc, rc, lfix;
                wc_traj_feet f = {lf, rf, nullptr, nullptr, &lc, &rc, &lfix, nullptr, nullptr};
                wco_traj_feet(tw, (int)s, &f);
                wc_compat::toH12(lT[s], got);
                for (int e = 0; e < 12; e++) worst_traj = std::fmax(worst_traj, std::fabs(got[e] - lf[e]));
                if (lC[s] != (lc != 0) || rC[s] != (rc != 0) || lFix[s] != (lfix != 0)) { std::fprintf(stderr, "trajectory: flags at %zu\n", s); failures++; break; }
            }
            wco_traj_free(tw);
        }
    }
    std::printf("trajectory generator: worst abs err %.3e\n", worst_traj);
    if (worst_traj > 1e-12) failures++;
    wco_mpc_free(ompc);
    wco_ikosqp_free(oikOsqp);
    std::printf("WalkingFK::evaluateDCM: worst abs err %.3e\n", worst_fk_dcm);
    if (worst_fk_dcm > 1e-12) failures++;
    std::printf("WalkingQPIK_osqp: worst dq err %.3e rad/s, %d solver set-ups (stance changes + 1)\n", worst_dq_osqp, osqpSetups);
    if (worst_dq_osqp > 1e-5) failures++;
    std::printf("reactive controller: worst abs err %.3e\n", worst_reactive);
    if (worst_reactive > 1e-13) failures++;
    std::printf("drop-in test: %d ticks, horizon %d, worst rel ZMP err %.3e, worst dq err %.3e rad/s, failures %d\n", T, N, worst_zmp,
                worst_dq, failures);
    return (failures == 0 && worst_zmp <= 1e-6 && worst_dq <= 1e-5) ? 0 : 1;
}
