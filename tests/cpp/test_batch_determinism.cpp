// GPU re-host of the reference's tests/test_determinism.cpp on the batched controller: several controllers are built
// from the same YAML (stabilizer / collision detection / closed loop off, as test_determinism.cpp:66-69), each runs the
// same open-loop squat-like CoM trajectory, and the command sequences must agree within 1e-8 (test_determinism.cpp:51,63).
// In addition every robot of the batch that receives identical inputs must produce identical commands.
#include <cmath>
#include <cstdio>
#include <vector>

#include "inria_wbc_b200/batch_pos_tracker.hpp"

using namespace inria_wbc::controllers;

static std::vector<std::vector<double>> run(const std::string& etc, const std::string& robots, int B, int duration)
{
    Config c;
    c.controller_yaml = etc + "/talos/talos_pos_tracker.yaml";
    c.robots_path = robots;
    c.etc_path = etc;
    c.overrides_yaml = "stabilizer: {activated: false}\ncollision_detection: {activated: false}\nclosed_loop: false\n";
    c.batch = B;
    auto controller = Factory::instance().create("talos-pos-tracker", c);
    auto p = std::dynamic_pointer_cast<BatchPosTracker>(controller);
    double com0[3];
    p->get_com_ref(0, com0);
    std::vector<std::vector<double>> cmds;
    for (int t = 0; t < duration; ++t) {
        // min-jerk descent of the CoM by 5 cm (robots 0 and 1 identical, the others by 2.5 cm)
        const double s = std::min(1.0, (t + 1.0) / duration), mj = s * s * s * (10 - 15 * s + 6 * s * s);
        for (int i = 0; i < B; ++i) {
            double com[3] = {com0[0], com0[1], com0[2] - (i < 2 ? 0.05 : 0.025) * mj};
            p->set_com_ref(i, com);
        }
        p->update(); // default (no) sensors: open loop on the integrated state
        if (!p->all_optimal())
            throw IWBC_B200_EXCEPTION("QP not solved to optimality");
        cmds.push_back(p->q());
    }
    return cmds;
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: test_batch_determinism <etc_path> <robots_path> [duration]\n");
        return 2;
    }
    const int B = 4, nexp = 3, duration = argc > 3 ? std::atoi(argv[3]) : 300;
    std::vector<std::vector<std::vector<double>>> results;
    try {
        for (int e = 0; e < nexp; ++e)
            results.push_back(run(argv[1], argv[2], B, duration));
    }
    catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 1;
    }
    int errors = 0;
    const size_t nv = results[0][0].size() / B;
    for (int i = 0; i < nexp; ++i)
        for (int j = 0; j < nexp; ++j) {
            if (i == j)
                continue;
            for (int t = 0; t < duration; ++t) {
                double n2 = 0;
                for (size_t k = 0; k < results[i][t].size(); ++k) {
                    const double d = results[i][t][k] - results[j][t][k];
                    n2 += d * d;
                }
                if (std::sqrt(n2) >= 1e-8)
                    ++errors;
            }
        }
    // identical robots inside one batch
    double moved = 0;
    for (int t = 0; t < duration; ++t) {
        for (size_t k = 0; k < nv; ++k) {
            if (results[0][t][k] != results[0][t][nv + k])
                ++errors;
            moved = std::max(moved, std::fabs(results[0][t][k] - results[0][0][k]));
        }
    }
    if (!(moved > 1e-3)) { // the squat really happens
        std::printf("the robot did not move (%g)\n", moved);
        ++errors;
    }
    std::printf(errors ? "FAILED: %d differing ticks\n" : "OK: %d experiments x %d ticks x %d robots identical (|dq| moved %.4f)\n", errors ? errors : nexp,
        duration, B, moved);
    return errors ? 1 : 0;
}
