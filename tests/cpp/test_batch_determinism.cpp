// GPU re-host of the reference's tests/test_determinism.cpp at the reference's scale and shape
// (/root/reference/tests/test_determinism.cpp:24-138): 4 sub-tests {arm.yaml, squat.yaml} x {talos.urdf, talos_fast.urdf};
// each builds `nexp` = 5 FRESH controllers from the same YAML with stabilizer / collision detection / closed loop forced off
// (:66-69), creates the behaviour, runs behavior->update() `duration` = 50 000 times recording controller->q(false) and the
// mimic-filtered controllable tail after every tick (:26-42), and requires pairwise ||q_i - q_j|| < 1e-8 on every tick for both
// (:44-57, 109-115).  Here every update is a device tick of the CUDA path; the command trace comes back through
// iwbc_batch_behavior_run_trace.  B = 2 identical robots per controller additionally pin "same input, same output" inside a batch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "inria_wbc_b200/batch_pos_tracker.hpp"

using namespace inria_wbc;

struct Run {
    std::vector<double> cmds;          // [duration][B][nv]
    std::vector<int> keep;             // non-mimic indexes (filter_cmd)
    int nv = 0, ncontrollable = 0, B = 0, final_status = 0;
};

static Run test_behavior(const std::string& etc, const std::string& robots, const std::string& behavior_yaml, const std::string& urdf, int duration)
{
    controllers::Config c;
    c.controller_yaml = etc + "/talos/talos_pos_tracker.yaml";
    c.robots_path = robots;
    c.etc_path = etc;
    c.overrides_yaml = "stabilizer: {activated: false}\ncollision_detection: {activated: false}\nclosed_loop: false\nurdf: " + urdf + "\n";
    c.batch = 2;
    auto controller = controllers::Factory::instance().create("talos-pos-tracker", c);
    auto p_controller = std::dynamic_pointer_cast<controllers::BatchPosTracker>(controller);
    if (!p_controller || p_controller->tasks().empty())
        throw IWBC_B200_EXCEPTION("not a PosTracker / no tasks");
    behaviors::BatchBehavior behavior(controller, etc + "/talos/" + behavior_yaml);
    Run r;
    r.nv = controller->nv(), r.B = c.batch;
    r.keep = controller->non_mimic_indexes();
    r.ncontrollable = (int)controller->controllable_dofs().size();
    behavior.update(duration, &r.cmds);
    // A QP that is not solved to optimality is data in the batched path: the robot keeps its state from that tick on and the
    // status sticks (the reference would throw out of behavior->update()).  On the stand-in Talos models of this repo the squat
    // drifts into torso_2_joint's lower limit after ~30 000 ticks (TSID's position bound returns 1e6 at q == qmin: infeasible);
    // determinism is still required over all `duration` ticks, including WHEN that happens.
    r.final_status = controller->status()[0];
    return r;
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: test_batch_determinism <etc_path> <robots_path> [duration] [nexp]\n");
        return 2;
    }
    const int duration = argc > 3 ? std::atoi(argv[3]) : 50000, nexp = argc > 4 ? std::atoi(argv[4]) : 5;
    const char* behaviors_paths[] = {"arm.yaml", "squat.yaml"};
    const char* urdfs[] = {"talos/talos.urdf", "talos/talos_fast.urdf"};
    int errors = 0;
    for (const char* bp : behaviors_paths)
        for (const char* urdf : urdfs) {
            // one fresh controller per host thread, all on the same GPU at the same time (the reference's harness runs one controller
            // per std::thread too, tests/utest.hpp:79-92): handles share nothing on the device
            std::vector<Run> results(nexp);
            std::vector<std::string> failures(nexp);
            std::vector<std::thread> pool;
            for (int e = 0; e < nexp; ++e)
                pool.emplace_back([&, e]() {
                    try {
                        results[e] = test_behavior(argv[1], argv[2], bp, urdf, duration);
                    }
                    catch (const std::exception& ex) {
                        failures[e] = ex.what();
                    }
                });
            for (auto& t : pool)
                t.join();
            for (int e = 0; e < nexp; ++e)
                if (!failures[e].empty()) {
                    std::printf("exception in determinism_%s_%s: %s\n", bp, urdf, failures[e].c_str());
                    return 1;
                }
            const Run& a = results[0];
            const size_t row = (size_t)a.B * a.nv;
            double worst = 0, worst_f = 0, moved = 0;
            int bad = 0;
            for (int i = 0; i < nexp; ++i)
                for (int j = 0; j < nexp; ++j) {
                    if (i == j)
                        continue;
                    const Run& b = results[j];
                    for (int t = 0; t < duration; ++t)
                        for (int r = 0; r < a.B; ++r) {
                            const double* x = &results[i].cmds[t * row + (size_t)r * a.nv];
                            const double* y = &b.cmds[t * row + (size_t)r * a.nv];
                            double n2 = 0, n2f = 0;
                            for (int k = 0; k < a.nv; ++k)
                                n2 += (x[k] - y[k]) * (x[k] - y[k]);
                            // filter_cmd(cmd).tail(ncontrollable)
                            for (size_t k = a.keep.size() - a.ncontrollable; k < a.keep.size(); ++k)
                                n2f += (x[a.keep[k]] - y[a.keep[k]]) * (x[a.keep[k]] - y[a.keep[k]]);
                            worst = std::max(worst, std::sqrt(n2)), worst_f = std::max(worst_f, std::sqrt(n2f));
                            if (std::sqrt(n2) >= 1e-8 || std::sqrt(n2f) >= 1e-8)
                                ++bad;
                        }
                }
            // the two identical robots of one controller; and the behaviour really moves the robot
            for (int t = 0; t < duration; ++t)
                for (int k = 0; k < a.nv; ++k) {
                    if (a.cmds[t * row + k] != a.cmds[t * row + a.nv + k])
                        ++bad;
                    moved = std::max(moved, std::fabs(a.cmds[t * row + k] - a.cmds[k]));
                }
            if (!(moved > 1e-3))
                ++bad;
            int frozen_from = duration; // first tick from which the command no longer changes (only meaningful when the final status is not OPTIMAL)
            for (int i = 0; i < nexp; ++i)
                if (results[i].final_status != a.final_status)
                    ++bad;
            if (a.final_status != controllers::OPTIMAL)
                for (int t = duration - 1; t > 0; --t) {
                    bool same = true;
                    for (int k = 0; k < a.nv && same; ++k)
                        same = a.cmds[t * row + k] == a.cmds[(t - 1) * row + k];
                    if (!same)
                        break;
                    frozen_from = t;
                }
            (void)frozen_from;
            std::printf("determinism_%s_%s: %d controllers x %d ticks, nv %d (%d controllable after the mimic filter): max pairwise |dq| %.3g (filtered %.3g), moved %.4f, "
                        "final status %d%s -> %s\n",
                bp, urdf, nexp, duration, a.nv, a.ncontrollable, worst, worst_f, moved, a.final_status,
                a.final_status ? " (sticky: at some tick the QP was not solved to optimality; the robot kept its state over such ticks, identically in every run)" : "",
                bad ? "FAILED" : "ok");
            errors += bad;
        }
    std::printf(errors ? "FAILED: %d differing ticks\n" : "OK\n", errors);
    return errors ? 1 : 0;
}
