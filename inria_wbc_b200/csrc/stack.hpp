// stack.hpp -- YAML controller config -> compiled problem (iwbc_problem_t) + default reference vector.
//
// Mirrors, step by step, what the reference does once at construction:
//   Controller ctor        /root/reference/src/controllers/controller.cpp:65-130  (URDF, free-flyer, dt)
//   PosTracker ctor        /root/reference/src/controllers/pos_tracker.cpp:40-159 (frames, SRDF posture,
//                          quaternion normalisation :71-77, first computeProblemData at v = 0 :105-106,
//                          parse_tasks :161-189 in YAML order, joint_range_reduction :115-148)
//   task factories         /root/reference/src/controllers/tasks.cpp:38-404 (Kd = 2 sqrt(Kp), masks,
//                          level 0 for bounds / actuation-bounds / contacts, level 1 otherwise)
//   HumanoidPosTracker     /root/reference/src/controllers/humanoid_pos_tracker.cpp:50-79 (init_com)
//   TalosPosTracker        /root/reference/src/controllers/talos_pos_tracker.cpp:41-59 (torso_max_roll)
// Upstream TSID constants (Contact6d generators, friction pyramid, regularisation weights) follow
// SURVEY.md Appendix A.3.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/iwbc_problem.h"
#include "model.hpp"
#include "yaml_lite.hpp"

namespace iwbc {
namespace host {

    struct ConfigError : std::runtime_error {
        explicit ConfigError(const std::string& m) : std::runtime_error(m) {}
    };

#define IWBC_HCHECK(expr, label_)                                                                \
    [&]() {                                                                                     \
        try {                                                                                   \
            return (expr);                                                                      \
        }                                                                                       \
        catch (std::exception & e_) {                                                           \
            throw ::iwbc::host::ConfigError(std::string("config error reading ") + label_ + ": " + e_.what()); \
        }                                                                                       \
    }()

    struct CompiledStack {
        iwbc_problem_t prob;
        Model model;
        std::vector<double> ref0;                 // default per-instance reference vector (construction-time values)
        std::vector<std::string> task_names;      // index = task id
        std::vector<std::string> contact_names;   // index = contact id
        std::vector<std::string> frame_names;     // index = frame id in prob
        std::vector<std::string> mimic_dof_names;
        std::string controller_name;
        bool closed_loop = false;
        bool torque_collision_detection = false; // CONTROLLER.collision_detection.activated (talos-pos-tracker)
    };

    namespace detail {
        inline void put_se3(const SE3& m, double* out12)
        { // jplace/fplace layout: R row-major (9), p (3)
            for (int i = 0; i < 9; ++i)
                out12[i] = m.R[i];
            for (int i = 0; i < 3; ++i)
                out12[9 + i] = m.p[i];
        }
        inline void put_sample_se3(const SE3& m, double* ref24)
        { // TrajectorySample(12,6): p, R column-major; vel, acc zero
            std::memset(ref24, 0, sizeof(double) * IWBC_REF_SE3);
            for (int i = 0; i < 3; ++i)
                ref24[i] = m.p[i];
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r)
                    ref24[3 + 3 * c + r] = m.R[3 * r + c];
        }
        inline int use_frame(CompiledStack& cs, const std::string& name)
        {
            for (size_t i = 0; i < cs.frame_names.size(); ++i)
                if (cs.frame_names[i] == name)
                    return (int)i;
            int id = cs.model.frame_id(name);
            if (id < 0)
                throw ConfigError("Unknown frame or joint [" + name + "]");
            if ((int)cs.frame_names.size() >= IWBC_MAXF)
                throw ConfigError("too many task frames (IWBC_MAXF)");
            int k = (int)cs.frame_names.size();
            cs.frame_names.push_back(name);
            cs.prob.fparent[k] = cs.model.frames[id].parent;
            put_se3(cs.model.frames[id].placement, cs.prob.fplace[k]);
            cs.prob.nframes = k + 1;
            return k;
        }
        inline void parse_mask(const std::string& s, size_t expect, unsigned char* out, int& nrows, const std::string& what)
        {
            if (s.size() != expect)
                throw ConfigError(what + " mask needs " + std::to_string(expect) + " entries, here: " + s);
            nrows = 0;
            for (size_t i = 0; i < s.size(); ++i) {
                out[i] = s[i] == '0' ? 0 : 1;
                nrows += out[i];
            }
        }
        inline int alloc_ref(CompiledStack& cs, int n)
        {
            int off = (int)cs.ref0.size();
            cs.ref0.resize(off + n, 0.0);
            return off;
        }
    } // namespace detail

    // `robots_path`: directory the `urdf:` key is resolved against (stands in for robot_dart's utheque);
    // `etc_path`: directory substituted for a leading "@ETC@" in base_path.
    inline CompiledStack compile_stack(const yaml::Node& config, const std::string& robots_path, const std::string& etc_path)
    {
        using namespace detail;
        CompiledStack cs;
        std::memset(&cs.prob, 0, sizeof(cs.prob));
        iwbc_problem_t& P = cs.prob;
        P.magic = IWBC_PROBLEM_MAGIC;
        P.version = IWBC_PROBLEM_VERSION;

        const yaml::Node& c = config["CONTROLLER"];
        if (!c.is_map())
            throw ConfigError("config error: missing CONTROLLER section");
        cs.controller_name = IWBC_HCHECK(c["name"].as<std::string>(), "CONTROLLER.name");
        std::string base_path = IWBC_HCHECK(c["base_path"].as<std::string>(), "CONTROLLER.base_path");
        if (base_path.rfind("@ETC@", 0) == 0)
            base_path = etc_path + base_path.substr(5);
        std::string solver = IWBC_HCHECK(c["solver"].as<std::string>(), "CONTROLLER.solver");
        if (solver != "eiquadprog")
            throw ConfigError("solver in configuration file must be 'eiquadprog' (the 'qpmad' backend is out of scope)");
        P.dt = IWBC_HCHECK(c["dt"].as<double>(), "CONTROLLER.dt");
        bool floating_base = IWBC_HCHECK(c["floating_base"].as<bool>(), "CONTROLLER.floating_base");
        std::string fb_name = IWBC_HCHECK(c["floating_base_joint_name"].as<std::string>(), "CONTROLLER.floating_base_joint_name");
        if (floating_base && !fb_name.empty())
            throw ConfigError("floating_base_joint_name: a floating base already present in the URDF is not supported");
        cs.mimic_dof_names = IWBC_HCHECK(c["mimic_dof_names"].as<std::vector<std::string>>(), "CONTROLLER.mimic_dof_names");
        cs.closed_loop = IWBC_HCHECK(c["closed_loop"].as<bool>(), "CONTROLLER.closed_loop");
        (void)IWBC_HCHECK(c["verbose"].as<bool>(), "CONTROLLER.verbose");
        std::string urdf = IWBC_HCHECK(c["urdf"].as<std::string>(), "CONTROLLER.urdf");
        std::string urdf_path = (!urdf.empty() && urdf[0] == '/') ? urdf : robots_path + "/" + urdf;

        cs.model = load_urdf(urdf_path, floating_base);
        Model& model = cs.model;
        if (c["frames"]) {
            yaml::Node fr = yaml::LoadFile(base_path + "/" + c["frames"].as<std::string>());
            for (auto& kv : fr.map) {
                std::string ref = IWBC_HCHECK(kv.second["ref"].as<std::string>(), "frames." + kv.first + ".ref");
                auto pos = IWBC_HCHECK(kv.second["pos"].as<std::vector<double>>(), "frames." + kv.first + ".pos");
                if (pos.size() != 3)
                    throw ConfigError("frames." + kv.first + ".pos needs 3 numbers");
                model.add_virtual_frame(kv.first, ref, pos.data());
            }
        }
        if ((int)model.joints.size() > IWBC_MAXJ || model.nv > IWBC_MAXNV)
            throw ConfigError("model too large for IWBC_MAXJ / IWBC_MAXNV");

        // ---- reference posture
        std::string srdf = IWBC_HCHECK(c["configurations"].as<std::string>(), "CONTROLLER.configurations");
        std::string ref_config = IWBC_HCHECK(c["ref_config"].as<std::string>(), "CONTROLLER.ref_config");
        load_srdf_configurations(model, base_path + "/" + srdf);
        if (!model.reference_configurations.count(ref_config))
            throw ConfigError("The following reference config is not in ref_map : " + ref_config);
        std::vector<double> q0 = model.reference_configurations[ref_config];
        if (floating_base) {
            double n = std::sqrt(q0[3] * q0[3] + q0[4] * q0[4] + q0[5] * q0[5] + q0[6] * q0[6]);
            for (int i = 3; i < 7; ++i)
                q0[i] /= n;
        }

        // ---- tree
        P.nj = (int)model.joints.size();
        P.nq = model.nq;
        P.nv = model.nv;
        P.na = model.na();
        P.floating_base = floating_base ? 1 : 0;
        P.gravity[0] = 0;
        P.gravity[1] = 0;
        P.gravity[2] = -9.81;
        P.hess_reg = 1e-8;
        P.sc_jacobian_world = 0;
        if (c["self_collision_jacobian"]) // extension key (not in the reference): "world" | "local_world_aligned"
            P.sc_jacobian_world = c["self_collision_jacobian"].as<std::string>() == "world" ? 1 : 0;
        for (int j = 0; j < P.nj; ++j) {
            const Joint& J = model.joints[j];
            P.jtype[j] = J.type;
            P.parent[j] = J.parent;
            P.idx_q[j] = J.idx_q;
            P.idx_v[j] = J.idx_v;
            P.depth[j] = J.parent >= 0 ? P.depth[J.parent] + 1 : 0;
            put_se3(J.placement, P.jplace[j]);
            for (int i = 0; i < 3; ++i) {
                P.axis[j][i] = J.axis[i];
                P.com[j][i] = J.Y.c[i];
            }
            P.mass[j] = J.Y.m;
            P.total_mass += J.Y.m;
            P.inertia[j][0] = J.Y.I[0];
            P.inertia[j][1] = J.Y.I[1];
            P.inertia[j][2] = J.Y.I[2];
            P.inertia[j][3] = J.Y.I[4];
            P.inertia[j][4] = J.Y.I[5];
            P.inertia[j][5] = J.Y.I[8];
        }
        for (int j = P.nj - 1; j >= 0; --j) {
            if (P.last_desc[j] < j)
                P.last_desc[j] = j;
            int p = P.parent[j];
            if (p >= 0 && P.last_desc[p] < P.last_desc[j])
                P.last_desc[p] = P.last_desc[j];
        }
        {
            int k = 0;
            for (int j = 0; j < P.nj; ++j) {
                if (P.jtype[j] == IWBC_J_FREEFLYER)
                    continue;
                const Joint& J = model.joints[j];
                P.q_lo[k] = J.q_lo;
                P.q_hi[k] = J.q_hi;
                P.v_max[k] = J.v_max;
                P.tau_max[k] = J.tau_max;
                ++k;
            }
        }
        for (int i = 0; i < P.nq; ++i)
            P.q0[i] = q0[i];

        // ---- "compute problem data at init": FK at q0, v = 0 (pos_tracker.cpp:105-106)
        Kinematics kin = forward_kinematics(model, q0);

        // ---- base dynamics is the first level-0 block of a floating-base formulation
        if (floating_base)
            P.l0[P.nl0++] = {IWBC_B_BASEDYN, 0};

        // ---- tasks, in file order
        std::string task_file = IWBC_HCHECK(c["tasks"].as<std::string>(), "CONTROLLER.tasks");
        std::string task_path = (!task_file.empty() && task_file[0] == '/') ? task_file : base_path + "/" + task_file;
        yaml::Node tasks = yaml::LoadFile(task_path);
        if (!tasks.is_map())
            throw ConfigError("task file " + task_path + " is not a map");
        const int na = P.na, nv = P.nv;
        int bounds_task = -1;
        for (auto& kv : tasks.map) {
            const std::string& name = kv.first;
            const yaml::Node& node = kv.second;
            std::string type = IWBC_HCHECK(node["type"].as<std::string>(), name + ".type");
            auto new_task = [&](int ttype, int level) -> iwbc_task_t& {
                if (P.ntasks >= IWBC_MAXT)
                    throw ConfigError("too many tasks (IWBC_MAXT)");
                iwbc_task_t& t = P.tasks[P.ntasks];
                std::memset(&t, 0, sizeof(t));
                t.type = ttype;
                t.level = level;
                t.ref_off = -1;
                t.frame = -1;
                cs.task_names.push_back(name);
                int id = P.ntasks++;
                if (level == 0) {
                    if (P.nl0 >= IWBC_MAXL0)
                        throw ConfigError("too many level-0 blocks");
                    P.l0[P.nl0++] = {IWBC_B_TASK, id};
                }
                else {
                    if (P.nl1 >= IWBC_MAXL1)
                        throw ConfigError("too many level-1 blocks");
                    P.l1[P.nl1++] = {IWBC_B_TASK, id};
                }
                return t;
            };
            if (type == "se3") {
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                std::string tracked = IWBC_HCHECK(node["tracked"].as<std::string>(), name + ".tracked");
                std::string mask = IWBC_HCHECK(node["mask"].as<std::string>(), name + ".mask");
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                bool joint = model.exist_joint(tracked), body = model.exist_body(tracked);
                if (joint && body)
                    throw ConfigError("Ambiguous name to track for task " + name + ": this is both a joint and a frame [" + tracked + "]");
                if (!joint && !body)
                    throw ConfigError("Unknown frame or joint [" + tracked + "]");
                iwbc_task_t& t = new_task(IWBC_T_SE3, 1);
                parse_mask(mask, 6, t.mask, t.nrows, "SE3");
                t.weight = weight;
                for (int i = 0; i < 6; ++i) {
                    t.kp[i] = kp;
                    t.kd[i] = 2.0 * std::sqrt(kp);
                }
                t.frame = use_frame(cs, tracked);
                t.ref_off = alloc_ref(cs, IWBC_REF_SE3);
                SE3 ref = joint ? kin.oMi[model.joint_id(tracked)] : frame_placement(model, kin, model.frame_id(tracked));
                put_sample_se3(ref, &cs.ref0[t.ref_off]);
            }
            else if (type == "com") {
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                std::string mask = IWBC_HCHECK(node["mask"].as<std::string>(), name + ".mask");
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                iwbc_task_t& t = new_task(IWBC_T_COM, 1);
                parse_mask(mask, 3, t.mask, t.nrows, "CoM");
                t.weight = weight;
                for (int i = 0; i < 3; ++i) {
                    t.kp[i] = kp;
                    t.kd[i] = 2.0 * std::sqrt(kp);
                }
                t.ref_off = alloc_ref(cs, IWBC_REF_COM);
                for (int i = 0; i < 3; ++i)
                    cs.ref0[t.ref_off + i] = kin.com[i];
            }
            else if (type == "momentum") {
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                std::string mask = IWBC_HCHECK(node["mask"].as<std::string>(), name + ".mask");
                iwbc_task_t& t = new_task(IWBC_T_MOMENTUM, 1);
                parse_mask(mask, 6, t.mask, t.nrows, "Momentum");
                t.weight = weight;
                for (int i = 0; i < 6; ++i) {
                    t.kp[i] = kp;
                    t.kd[i] = 2.0 * std::sqrt(kp);
                }
                t.ref_off = alloc_ref(cs, IWBC_REF_MOM); // zeros
            }
            else if (type == "posture") {
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                std::string ref_name = IWBC_HCHECK(node["ref"].as<std::string>(), name + ".ref");
                if (!model.reference_configurations.count(ref_name))
                    throw ConfigError("Reference name " + ref_name + " not found");
                iwbc_task_t& t = new_task(IWBC_T_POSTURE, 1);
                t.weight = weight;
                t.kp[0] = kp;
                t.kd[0] = 2.0 * std::sqrt(kp);
                if (node["mask"]) {
                    std::string mask = node["mask"].as<std::string>();
                    parse_mask(mask, (size_t)na, t.mask, t.nrows, "posture");
                }
                else {
                    for (int i = 0; i < na; ++i)
                        t.mask[i] = 1;
                    t.nrows = na;
                }
                t.ref_off = alloc_ref(cs, 3 * na);
                const std::vector<double>& rq = model.reference_configurations[ref_name];
                for (int i = 0; i < na; ++i)
                    cs.ref0[t.ref_off + i] = rq[model.nq - na + i];
            }
            else if (type == "bounds") {
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                iwbc_task_t& t = new_task(IWBC_T_BOUNDS, 0);
                t.weight = weight;
                t.nrows = na;
                bounds_task = P.ntasks - 1;
            }
            else if (type == "actuation-bounds") {
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                iwbc_task_t& t = new_task(IWBC_T_ACTBOUNDS, 0);
                t.weight = weight;
                t.nrows = na;
            }
            else if (type == "self-collision") {
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                double kd = IWBC_HCHECK(node["kd"].as<double>(), name + ".kd");
                double m = IWBC_HCHECK(node["m"].as<double>(), name + ".m");
                std::string tracked = IWBC_HCHECK(node["tracked"].as<std::string>(), name + ".tracked");
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                double margin = IWBC_HCHECK(node["margin"].as<double>(), name + ".margin");
                double radius = IWBC_HCHECK(node["radius"].as<double>(), name + ".radius");
                const yaml::Node& av = node["avoided"];
                if (!av.is_map())
                    throw ConfigError("config error reading " + name + ".avoided");
                iwbc_task_t& t = new_task(IWBC_T_SELFCOL, 1);
                t.weight = weight;
                t.nrows = 1;
                t.kp[0] = kp;
                t.kd[0] = kd;
                t.radius = radius;
                t.sc_m = m;
                // 5PL constants (task-self-collision.cpp:163-164)
                t.sc_k = -std::log(std::pow(-1e-5 + 1., -1. / m) - 1.) / margin;
                t.sc_sp = -1. / t.sc_k * std::log(-1 + std::pow(2, 1. / m));
                t.frame = use_frame(cs, tracked);
                // the reference iterates an std::unordered_map (order unspecified): we fix YAML order
                for (auto& a : av.map) {
                    if (!model.exist_frame(a.first))
                        throw ConfigError("Frame " + a.first + " in " + name + " does not exists.");
                    if (t.nav >= IWBC_MAXAV)
                        throw ConfigError("too many avoided frames in " + name);
                    t.av_frame[t.nav] = use_frame(cs, a.first);
                    t.av_r0[t.nav] = a.second.as<double>();
                    ++t.nav;
                }
            }
            else if (type == "contact") {
                if (P.ncontacts >= IWBC_MAXC)
                    throw ConfigError("too many contacts (IWBC_MAXC)");
                double kp = IWBC_HCHECK(node["kp"].as<double>(), name + ".kp");
                std::string jn = IWBC_HCHECK(node["joint"].as<std::string>(), name + ".joint");
                double lxn = IWBC_HCHECK(node["lxn"].as<double>(), name + ".lxn");
                double lyn = IWBC_HCHECK(node["lyn"].as<double>(), name + ".lyn");
                double lxp = IWBC_HCHECK(node["lxp"].as<double>(), name + ".lxp");
                double lyp = IWBC_HCHECK(node["lyp"].as<double>(), name + ".lyp");
                double lz = IWBC_HCHECK(node["lz"].as<double>(), name + ".lz");
                double mu = IWBC_HCHECK(node["mu"].as<double>(), name + ".mu");
                auto normal = IWBC_HCHECK(node["normal"].as<std::vector<double>>(), name + ".normal");
                double fmin = IWBC_HCHECK(node["fmin"].as<double>(), name + ".fmin");
                double fmax = IWBC_HCHECK(node["fmax"].as<double>(), name + ".fmax");
                if (normal.size() != 3)
                    throw ConfigError("normal size:" + std::to_string(normal.size()));
                if (!model.exist_frame(jn) || !model.exist_joint(jn))
                    throw ConfigError(jn + " does not exist!");
                iwbc_contact_t& ct = P.contacts[P.ncontacts];
                std::memset(&ct, 0, sizeof(ct));
                cs.contact_names.push_back(name);
                int cid = P.ncontacts++;
                ct.frame = use_frame(cs, jn);
                ct.active_default = 1;
                for (int i = 0; i < 6; ++i) {
                    ct.kp[i] = kp;
                    ct.kd[i] = 2.0 * std::sqrt(kp);
                }
                // contact points (tasks.cpp:353-356): columns (-lxn,-lyn,lz) (-lxn,lyp,lz) (lxp,-lyn,lz) (lxp,lyp,lz)
                double Pt[4][3] = {{-lxn, -lyn, lz}, {-lxn, lyp, lz}, {lxp, -lyn, lz}, {lxp, lyp, lz}};
                for (int i = 0; i < 4; ++i) {
                    // T[:, 3i:3i+3] = [I3 ; skew(P_i)]
                    for (int r = 0; r < 3; ++r)
                        ct.T[r * 12 + 3 * i + r] = 1.0;
                    const double* p = Pt[i];
                    double S[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
                    for (int r = 0; r < 3; ++r)
                        for (int cc = 0; cc < 3; ++cc)
                            ct.T[(3 + r) * 12 + 3 * i + cc] = S[3 * r + cc];
                }
                // friction pyramid (Contact6d::updateForceInequalityConstraints)
                double n[3] = {normal[0], normal[1], normal[2]};
                auto cross = [](const double* a, const double* b, double* o) {
                    o[0] = a[1] * b[2] - a[2] * b[1];
                    o[1] = a[2] * b[0] - a[0] * b[2];
                    o[2] = a[0] * b[1] - a[1] * b[0];
                };
                double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, t1[3], t2[3];
                cross(n, ex, t1);
                if (std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]) < 1e-5)
                    cross(n, ey, t1);
                cross(n, t1, t2);
                double n1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                double n2 = std::sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
                for (int i = 0; i < 3; ++i) {
                    t1[i] /= n1;
                    t2[i] /= n2;
                }
                for (int i = 0; i < 4; ++i) {
                    for (int k = 0; k < 3; ++k) {
                        ct.B[(4 * i + 0) * 12 + 3 * i + k] = -t1[k] - mu * n[k];
                        ct.B[(4 * i + 1) * 12 + 3 * i + k] = t1[k] - mu * n[k];
                        ct.B[(4 * i + 2) * 12 + 3 * i + k] = -t2[k] - mu * n[k];
                        ct.B[(4 * i + 3) * 12 + 3 * i + k] = t2[k] - mu * n[k];
                        ct.B[16 * 12 + 3 * i + k] = n[k];
                    }
                }
                for (int i = 0; i < 16; ++i) {
                    ct.lb[i] = -1e10;
                    ct.ub[i] = 0.0;
                }
                ct.lb[16] = fmin;
                ct.ub[16] = fmax;
                const double wreg[6] = {1, 1, 1e-3, 2, 2, 2};
                for (int i = 0; i < 6; ++i)
                    ct.wreg[i] = wreg[i];
                ct.w_forcereg = 1e-3; // cst::w_force_feet, include/inria_wbc/controllers/tasks.hpp:23
                ct.ref_off = alloc_ref(cs, IWBC_REF_CONTACT); // SE3 sample, then f_ref (0) and the weight-vector index (0 = Contact6d's default)
                put_sample_se3(kin.oMi[model.joint_id(jn)], &cs.ref0[ct.ref_off]);
                // addRigidContact: motion (eq, level 0), force (ineq, level 0), regularisation (level 1)
                if (P.nl0 + 2 > IWBC_MAXL0 || P.nl1 + 1 > IWBC_MAXL1)
                    throw ConfigError("too many blocks");
                P.l0[P.nl0++] = {IWBC_B_CONTACT_MOTION, cid};
                P.l0[P.nl0++] = {IWBC_B_CONTACT_FORCE, cid};
                P.l1[P.nl1++] = {IWBC_B_CONTACT_REG, cid};
            }
            else if (type == "torque") {
                // make_torque (tasks.cpp:227-271): TaskActuationEquality, mask over the actuated joints (default all), optional
                // per-joint `scaling` (setWeightVector), zero reference, addActuationTask(task, weight, 1)
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                iwbc_task_t& t = new_task(IWBC_T_TORQUE, 1);
                t.weight = weight;
                if (node["mask"]) {
                    std::string mask = IWBC_HCHECK(node["mask"].as<std::string>(), name + ".mask");
                    if ((int)mask.size() != na)
                        throw ConfigError("wrong size in torque mask, expected:" + std::to_string(na) + " got:" + std::to_string(mask.size()));
                    parse_mask(mask, (size_t)na, t.mask, t.nrows, "torque");
                }
                else {
                    for (int i = 0; i < na; ++i)
                        t.mask[i] = 1;
                    t.nrows = na;
                }
                t.ref_off = alloc_ref(cs, 2 * na); // tau_ref = 0, then the weight vector (TaskActuationEquality's default: ones)
                for (int i = 0; i < na; ++i)
                    cs.ref0[t.ref_off + na + i] = 1.0;
                if (node["scaling"]) {
                    auto scaling = IWBC_HCHECK(node["scaling"].as<std::vector<double>>(), name + ".scaling");
                    if ((int)scaling.size() != na)
                        throw ConfigError("wrong size in torque scaling, expected:" + std::to_string(na) + " got:" + std::to_string(scaling.size()));
                    for (int i = 0; i < na; ++i)
                        cs.ref0[t.ref_off + na + i] = scaling[i];
                }
            }
            else if (type == "cop") {
                // make_cop (tasks.cpp:156-178): TaskCopEquality, reference (0, 0, 0), addForceTask(task, weight, 1)
                double weight = IWBC_HCHECK(node["weight"].as<double>(), name + ".weight");
                iwbc_task_t& t = new_task(IWBC_T_COP, 1);
                t.weight = weight;
                t.nrows = 3;
                t.ref_off = alloc_ref(cs, 6); // cop reference (0, 0, 0), contact normal (TaskCopEquality's default (0, 0, 1))
                cs.ref0[t.ref_off + 5] = 1.0;
            }
            else
                throw ConfigError("unknown task type '" + type + "' for task " + name);
        }
        P.nref = (int)cs.ref0.size();
        // force-regularisation weight vectors: [0] = Contact6d's default
        P.nwreg = 1;
        {
            const double wreg0[6] = {1, 1, 1e-3, 2, 2, 2};
            for (int i = 0; i < 6; ++i)
                P.wreg_variants[0][i] = wreg0[i];
        }

        // ---- joint_range_reduction (pos_tracker.cpp:115-148)
        if (c["joint_range_reduction"]) {
            double red = c["joint_range_reduction"].as<double>() / 180.0 * M_PI;
            if (red < 0)
                throw ConfigError("Joint range reduction: reduction must be positive.");
            for (int i = 0; i < na; ++i) {
                if (P.q_hi[i] - P.q_lo[i] < 2 * red)
                    throw ConfigError("Joint range reduction: reduction cannot be greater than actual range of motion");
                P.q_lo[i] += red;
                P.q_hi[i] -= red;
                double qi = q0[model.nq - na + i];
                if (qi < P.q_lo[i] || qi > P.q_hi[i])
                    throw ConfigError("Joint range reduction: q0 falls outside reduced joint space.");
            }
        }

        // ---- humanoid: CoM reference to the middle of the feet (humanoid_pos_tracker.cpp:50-79)
        const bool humanoid = cs.controller_name == "humanoid-pos-tracker" || cs.controller_name == "talos-pos-tracker";
        if (humanoid) {
            auto tid = [&](const std::string& n) {
                for (size_t i = 0; i < cs.task_names.size(); ++i)
                    if (cs.task_names[i] == n)
                        return (int)i;
                return -1;
            };
            auto cid = [&](const std::string& n) {
                for (size_t i = 0; i < cs.contact_names.size(); ++i)
                    if (cs.contact_names[i] == n)
                        return (int)i;
                return -1;
            };
            int com = tid("com"), lf = tid("lf"), rf = tid("rf");
            if (com < 0)
                throw ConfigError("init_com: no com task");
            if (lf < 0 || rf < 0 || cid("contact_lfoot") < 0 || cid("contact_rfoot") < 0)
                throw ConfigError("init_com: contact_rfoot or contact_lfoot or lf or rf is missing, cannot compute initial com reference");
            double* cref = &cs.ref0[P.tasks[com].ref_off];
            const double* l = &cs.ref0[P.tasks[lf].ref_off];
            const double* r = &cs.ref0[P.tasks[rf].ref_off];
            double fx = 0.5 * (l[0] + r[0]), fy = 0.5 * (l[1] + r[1]);
            double d = std::sqrt((kin.com[0] - fx) * (kin.com[0] - fx) + (kin.com[1] - fy) * (kin.com[1] - fy));
            if (d > 0.01)
                throw ConfigError("Wrong starting configuration: the CoM needs to be between the two feet with less than 1cm difference, but the distance is " + std::to_string(d));
            cref[0] = fx;
            cref[1] = fy;
            // ---- stabilizer parameter sets (HumanoidPosTracker::parse_stabilizer, humanoid_pos_tracker.cpp:81-114; parse_stab_conf,
            // include/inria_wbc/stabilizers/stabilizer_conf.hpp:46-92): one file per behaviour type
            const yaml::Node& sn = c["stabilizer"];
            if (!sn.is_map())
                throw ConfigError("config error reading CONTROLLER.stabilizer");
            P.stabilizer_activated = IWBC_HCHECK(sn["activated"].as<bool>(), "CONTROLLER.stabilizer.activated") ? 1 : 0;
            const char* keys[3] = {"params_fixed_base", "params_ss", "params_ds"}; // indexed by iwbc_behavior_type
            bool have_files = true;
            for (const char* k : keys)
                have_files = have_files && sn[k];
            if (have_files) {
                for (int bt = 0; bt < 3; ++bt) {
                    const std::string file = base_path + "/" + IWBC_HCHECK(sn[keys[bt]].as<std::string>(), std::string("CONTROLLER.stabilizer.") + keys[bt]);
                    yaml::Node st = yaml::LoadFile(file);
                    iwbc_stab_conf_t& sc = P.stab[bt];
                    auto vec = [&](const char* key, int n, double* out, const char* what) {
                        auto v = IWBC_HCHECK(st[key].as<std::vector<double>>(), file + ":" + key);
                        if ((int)v.size() != n)
                            throw ConfigError(std::string("you need ") + std::to_string(n) + " coefficient in " + what);
                        for (int i = 0; i < n; ++i)
                            out[i] = v[i];
                    };
                    vec("com", 6, sc.com_gains, "p for the com stabilizer");
                    vec("ankle", 6, sc.ankle_gains, "d for the ankle stabilizer");
                    vec("ffda", 3, sc.ffda_gains, "p for the ffda stabilizer");
                    vec("zmp_p", 6, sc.zmp_p, "p for the zmp stabilizer");
                    vec("zmp_d", 6, sc.zmp_d, "d for the zmp stabilizer");
                    vec("zmp_w", 6, sc.zmp_w, "w for the zmp stabilizer");
                    vec("momentum_p", 6, sc.momentum_p, "p for the momentum stabilizer");
                    vec("momentum_d", 6, sc.momentum_d, "d for the momentum stabilizer");
                    sc.use_zmp = IWBC_HCHECK(st["use_zmp"].as<bool>(), file + ":use_zmp") ? 1 : 0;
                    sc.use_momentum = IWBC_HCHECK(st["use_momentum"].as<bool>(), file + ":use_momentum") ? 1 : 0;
                    sc.filter_size = IWBC_HCHECK(st["filter_size"].as<int>(), file + ":filter_size");
                    if (sc.filter_size < 1 || sc.filter_size > 16)
                        throw ConfigError(file + ": filter_size must be in 1 .. 16 on the batched path");
                    sc.zmp_w_variant = 0;
                    if (sc.use_zmp) { // zmp_w becomes the contacts' regularisation weight vector: one more (deduplicated) variant
                        int v = -1;
                        for (int k = 0; k < P.nwreg && v < 0; ++k) {
                            bool same = true;
                            for (int i = 0; i < 6; ++i)
                                same = same && P.wreg_variants[k][i] == sc.zmp_w[i];
                            if (same)
                                v = k;
                        }
                        if (v < 0) {
                            if (P.nwreg >= IWBC_MAXWV)
                                throw ConfigError("too many distinct zmp_w vectors (IWBC_MAXWV)");
                            v = P.nwreg++;
                            for (int i = 0; i < 6; ++i)
                                P.wreg_variants[v][i] = sc.zmp_w[i];
                        }
                        sc.zmp_w_variant = v;
                    }
                }
                P.stab_present = 1;
            }
            else if (P.stabilizer_activated)
                throw ConfigError("config error reading CONTROLLER.stabilizer.params_fixed_base / params_ss / params_ds");
            P.stab_task_com = com, P.stab_task_lf = lf, P.stab_task_rf = rf;
            P.stab_task_torso = tid("torso"), P.stab_task_momentum = tid("momentum");
            P.stab_contact_lf = cid("contact_lfoot"), P.stab_contact_rf = cid("contact_rfoot");
            P.stab_imu_frame = -1;
            P.stab_talos = cs.controller_name == "talos-pos-tracker" ? 1 : 0;
            if (P.stab_talos && model.exist_frame("imu_link")) // TalosPosTracker::update reads frameVelocity("imu_link")
                P.stab_imu_frame = use_frame(cs, "imu_link");
        }
        else {
            P.stab_task_com = P.stab_task_lf = P.stab_task_rf = P.stab_task_torso = P.stab_task_momentum = -1;
            P.stab_contact_lf = P.stab_contact_rf = P.stab_imu_frame = -1;
        }
        // ---- Talos: limit the hip-roll range around the start posture (talos_pos_tracker.cpp:41-59)
        if (cs.controller_name == "talos-pos-tracker") {
            // `as<float>() / 180 * M_PI`: the division is done in float, then promoted (talos_pos_tracker.cpp:48)
            double roll = (double)(IWBC_HCHECK(c["torso_max_roll"].as<float>(), "CONTROLLER.torso_max_roll") / 180) * M_PI;
            auto names = model.actuated_names();
            for (const char* n : {"leg_left_2_joint", "leg_right_2_joint"}) {
                auto it = std::find(names.begin(), names.end(), std::string(n));
                if (it == names.end())
                    throw ConfigError(std::string("Talos should have ") + n);
                int id = (int)(it - names.begin());
                double qi = q0[model.nq - na + id];
                if (!(P.q_lo[id] <= qi && P.q_hi[id] >= qi))
                    throw ConfigError("Error in bounds, the torso limits are not viable");
                P.q_lo[id] = qi - roll;
                P.q_hi[id] = qi + roll;
            }
            if (bounds_task < 0)
                throw ConfigError("talos-pos-tracker needs a 'bounds' task");
            // parse_torque_safety (talos_pos_tracker.cpp:62-108): the torque-discrepancy monitor itself is outside the batched path
            // (SURVEY.md 2.1 row 10); an activated one makes every update demand "joints_torque" sensor data, so it is remembered
            const yaml::Node& cd = c["collision_detection"];
            if (!cd.is_map())
                throw ConfigError("config error reading CONTROLLER.collision_detection");
            cs.torque_collision_detection = IWBC_HCHECK(cd["activated"].as<bool>(), "CONTROLLER.collision_detection.activated");
            (void)IWBC_HCHECK(cd["filter_size"].as<int>(), "CONTROLLER.collision_detection.filter_size");
            (void)IWBC_HCHECK(cd["max_invalid"].as<int>(), "CONTROLLER.collision_detection.max_invalid");
        }
        (void)nv;
        return cs;
    }

} // namespace host
} // namespace iwbc
