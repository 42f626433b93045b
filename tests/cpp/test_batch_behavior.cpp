// The behaviours through the C++ host mirror (behaviors::BatchBehavior): the reference drives a controller with
//     auto behavior = behaviors::Factory::instance().create(name, controller, yaml);  for (tick) behavior->update();
// (src/robot_dart/talos.cpp:300-390).  Here: Talos walking, four robots at different phases of the walk, 60 ticks crossing the first
// contact removal; run twice in one chunk and once tick by tick -- the three command sequences must agree within 1e-8
// (tests/test_determinism.cpp:51,63), every QP must be solved, and the contact masks must follow walk.cpp:203-221.
#include <cmath>
#include <cstdio>
#include <vector>

#include "inria_wbc_b200/batch_pos_tracker.hpp"

using namespace inria_wbc;

struct Out {
    std::vector<double> q, tau;
    std::vector<unsigned> masks;
};

static Out run(const std::string& etc, const std::string& robots, int ticks, int chunk)
{
    controllers::Config c;
    c.controller_yaml = etc + "/talos/talos_pos_tracker.yaml";
    c.robots_path = robots;
    c.etc_path = etc;
    c.overrides_yaml = "stabilizer: {activated: false}\ncollision_detection: {activated: false}\nclosed_loop: false\n";
    c.batch = 4;
    auto controller = controllers::Factory::instance().create("talos-pos-tracker", c);
    behaviors::BatchBehavior walk(controller, etc + "/talos/walk.yaml", "num_of_cycles: 1\n");
    if (walk.total_ticks() != 10000)
        throw IWBC_B200_EXCEPTION("walk with one cycle must last 10 s");
    walk.set_phase({0, 0, 960, 500});
    for (int t = 0; t < ticks; t += chunk) {
        walk.update(chunk);
        if (!controller->all_optimal())
            throw IWBC_B200_EXCEPTION("QP not solved to optimality");
    }
    if (walk.tick() != ticks || std::fabs(controller->t() - ticks * controller->dt()) > 1e-12)
        throw IWBC_B200_EXCEPTION("tick / time bookkeeping");
    return {controller->q(), controller->tau(), controller->contact_masks()};
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: test_batch_behavior <etc_path> <robots_path>\n");
        return 2;
    }
    try {
        const int ticks = 60;
        Out a = run(argv[1], argv[2], ticks, ticks), b = run(argv[1], argv[2], ticks, ticks), c = run(argv[1], argv[2], ticks, 1);
        double e_ab = 0, e_ac = 0;
        for (size_t k = 0; k < a.q.size(); ++k) {
            e_ab = std::fmax(e_ab, std::fabs(a.q[k] - b.q[k]));
            e_ac = std::fmax(e_ac, std::fabs(a.q[k] - c.q[k]));
        }
        const size_t nv = a.q.size() / 4;
        double same01 = 0;
        for (size_t k = 0; k < nv; ++k)
            same01 = std::fmax(same01, std::fabs(a.q[k] - a.q[nv + k]));
        // robot 2 is at tick 960 + 59 = 1019 of the walk: LF_INIT removed the left-foot contact at tick 1000 (walk.cpp:203-208)
        const unsigned ds = a.masks[0];
        const bool masks_ok = a.masks[1] == ds && a.masks[3] == ds && a.masks[2] != ds && (a.masks[2] & ds) == a.masks[2] && a.masks == c.masks;
        if (e_ab > 1e-8 || e_ac > 1e-8 || same01 != 0.0 || !masks_ok) {
            std::printf("FAILED run-to-run %g chunked-vs-ticks %g robots0-1 %g masks %u %u %u %u\n", e_ab, e_ac, same01, a.masks[0], a.masks[1], a.masks[2],
                a.masks[3]);
            return 1;
        }
        // unknown behaviour name: an exception naming what exists (factory.hpp:51-63 behaviour)
        bool thrown = false;
        try {
            controllers::Config cc;
            cc.controller_yaml = std::string(argv[1]) + "/talos/talos_pos_tracker.yaml";
            cc.robots_path = argv[2];
            cc.etc_path = argv[1];
            cc.batch = 1;
            auto ctl = controllers::Factory::instance().create("talos-pos-tracker", cc);
            behaviors::BatchBehavior bad(ctl, std::string(argv[1]) + "/talos/walk.yaml", "name: humanoid::no-such-behavior\n");
        }
        catch (const inria_wbc::Exception& e) {
            thrown = std::string(e.what()).find("humanoid::walk") != std::string::npos;
        }
        if (!thrown) {
            std::printf("FAILED: unknown behaviour did not throw\n");
            return 1;
        }
        // cost(task) / objective_value() (controller.hpp:148-153, pos_tracker.hpp:43-44): after one open-loop update, of every robot;
        // identical robots give identical numbers, the posture cost equals the norm built from ddq and the gains by hand
        {
            controllers::Config cc;
            cc.controller_yaml = std::string(argv[1]) + "/talos/talos_pos_tracker.yaml";
            cc.robots_path = argv[2];
            cc.etc_path = argv[1];
            cc.overrides_yaml = "stabilizer: {activated: false}\ncollision_detection: {activated: false}\nclosed_loop: false\n";
            cc.batch = 2;
            auto ctl = controllers::Factory::instance().create("talos-pos-tracker", cc);
            ctl->update();
            const double f0 = ctl->objective_value(0), f1 = ctl->objective_value(1);
            const double c_com = ctl->cost("com", 0), c_lh = ctl->cost("lh", 1), c_post = ctl->cost("posture", 0);
            // at the start pose the posture task asks for zero acceleration: its cost is the norm of the actuated accelerations
            double want = 0;
            const std::vector<double>& a = ctl->ddq();
            const size_t nv1 = a.size() / 2;
            for (size_t k = 6; k < nv1; ++k)
                want += a[k] * a[k];
            want = std::sqrt(want);
            if (!(f0 == f1) || !std::isfinite(f0) || !(c_com >= 0) || !(c_lh >= 0) || std::fabs(c_post - want) > 1e-9 * std::fmax(1.0, want)) {
                std::printf("FAILED getters: objective %g %g cost com %g lh %g posture %g (want %g)\n", f0, f1, c_com, c_lh, c_post, want);
                return 1;
            }
        }
        std::printf("OK run-to-run %g chunked-vs-ticks %g\n", e_ab, e_ac);
        return 0;
    }
    catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 1;
    }
}
