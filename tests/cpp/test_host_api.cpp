// Host-side (no GPU) checks of the C++ mirror of the reference's controller surface
// (include/inria_wbc_b200/batch_pos_tracker.hpp): factory names and error behaviour (factory.hpp:40-63),
// construction from the reference's YAML files, task / contact lookup errors (pos_tracker.hpp:28-39).
#include <cstdio>
#include <cstring>
#include <iostream>

#include "inria_wbc_b200/batch_pos_tracker.hpp"

using namespace inria_wbc::controllers;

static int failures = 0;
#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            ++failures;                                                 \
        }                                                               \
    } while (0)

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: test_host_api <etc_path> <robots_path>\n");
        return 2;
    }
    const std::string etc = argv[1], robots = argv[2];
    Config c;
    c.controller_yaml = etc + "/talos/talos_pos_tracker.yaml";
    c.robots_path = robots;
    c.etc_path = etc;
    c.batch = 4;
    // unknown controller name: exception listing the registered names
    try {
        Factory::instance().create("no-such-controller", c);
        CHECK(false);
    }
    catch (const inria_wbc::Exception& e) {
        CHECK(std::strstr(e.what(), "talos-pos-tracker") != nullptr);
    }
    // missing YAML file: exception, not a crash
    try {
        Config bad = c;
        bad.controller_yaml = etc + "/talos/does_not_exist.yaml";
        Factory::instance().create("talos-pos-tracker", bad);
        CHECK(false);
    }
    catch (const inria_wbc::Exception&) {
    }
    auto controller = Factory::instance().create("talos-pos-tracker", c);
    auto p = std::dynamic_pointer_cast<BatchPosTracker>(controller);
    CHECK(p.get() != nullptr);
    CHECK(!p->tasks().empty()); // tests/test_determinism.cpp:93
    CHECK(p->nq() == 51 && p->nv() == 50 && p->na() == 44);
    CHECK(p->has_task("com") && p->has_task("lh") && !p->has_task("nope"));
    CHECK(p->has_contact("contact_lfoot") && p->has_contact("contact_rfoot"));
    CHECK(p->controllable_dofs(false).size() == 44 && p->controllable_dofs(true).size() == 32);
    double com[3];
    p->get_com_ref(0, com);
    com[2] -= 0.05;
    p->set_com_ref(1, com);
    double back[3];
    p->get_com_ref(1, back);
    CHECK(back[2] == com[2]);
    p->get_com_ref(0, back);
    CHECK(back[2] != com[2]); // per-instance references
    try {
        double pos[12] = {0};
        p->set_se3_ref(0, pos, "unknown_task");
        CHECK(false);
    }
    catch (const inria_wbc::Exception& e) {
        CHECK(std::strstr(e.what(), "not found") != nullptr);
    }
    try {
        p->remove_contact(0, "contact_nofoot");
        CHECK(false);
    }
    catch (const inria_wbc::Exception&) {
    }
    try {
        std::vector<double> q(3), dq(3);
        p->update(q, dq); // wrong sizes: the reference asserts (controller.cpp:163-205)
        CHECK(false);
    }
    catch (const inria_wbc::Exception&) {
    }
    // Controller::update(const SensorData&): the reference's assertions, in its order, before anything touches the GPU
    try { // the shipped Talos file has stabilizer.activated: true (humanoid_pos_tracker.cpp:180)
        p->update(SensorData{});
        CHECK(false);
    }
    catch (const inria_wbc::Exception& e) {
        CHECK(std::strstr(e.what(), "the stabilizer needs the LF force") != nullptr);
    }
    try {
        SensorData s;
        s["lf_force"] = std::vector<double>(12, 0.0);
        p->update(s);
        CHECK(false);
    }
    catch (const inria_wbc::Exception& e) {
        CHECK(std::strstr(e.what(), "the stabilizer needs the RF force") != nullptr);
    }
    {
        Config cl = c;
        cl.overrides_yaml = "closed_loop: true\nstabilizer: {activated: false}\n";
        auto pc = Factory::instance().create("talos-pos-tracker", cl);
        try { // controller.cpp:165-168
            pc->update(SensorData{});
            CHECK(false);
        }
        catch (const inria_wbc::Exception& e) {
            CHECK(std::strstr(e.what(), "we need the joint positions in closed loop mode!") != nullptr);
        }
        try {
            SensorData s;
            s["positions"] = std::vector<double>(4 * 44, 0.0);
            s["joint_velocities"] = std::vector<double>(4 * 44, 0.0);
            pc->update(s);
            CHECK(false);
        }
        catch (const inria_wbc::Exception& e) {
            CHECK(std::strstr(e.what(), "we need the floating base position in closed loop mode!") != nullptr);
        }
        CHECK(pc->non_mimic_indexes().size() == 6 + 32);
        CHECK(pc->q(true).size() == (size_t)4 * 38 && pc->q(false).size() == (size_t)4 * 50);
        try {
            pc->set_behavior_type(0, 7); // controller.cpp:438-439
            CHECK(false);
        }
        catch (const inria_wbc::Exception&) {
        }
    }
    // AutoRegister (factory.hpp:66-76): a plugin registers its controller under a new name with a static object
    struct MyController : BatchPosTracker {
        explicit MyController(const Config& cc) : BatchPosTracker(cc) {}
    };
    static Register<MyController> reg("my-controller");
    CHECK(std::dynamic_pointer_cast<MyController>(Factory::instance().create("my-controller", c)).get() != nullptr);
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
