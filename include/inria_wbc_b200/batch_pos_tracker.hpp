// batch_pos_tracker.hpp -- C++ host-side mirror of the reference's controller surface for the batched path.
//
// Header-only, STL-only wrapper over the C ABI (include/iwbc_batch.h).  It keeps the reference's names, argument
// meaning and error behaviour for the ONE path this library replaces, with a batch index added where the reference
// has one robot:
//   inria_wbc::controllers::Controller      include/inria_wbc/controllers/controller.hpp:46-237   -> BatchController
//   inria_wbc::controllers::PosTracker      include/inria_wbc/controllers/pos_tracker.hpp:17-98   -> BatchPosTracker
//   inria_wbc::utils::Factory               include/inria_wbc/utils/factory.hpp:14-76             -> Factory (same names:
//                                           "pos-tracker", "humanoid-pos-tracker", "talos-pos-tracker")
//   tsid::solvers::SolverHQPBase            (used at src/controllers/pos_tracker.cpp:86-102,
//                                           src/controllers/controller.cpp:247)                     -> BatchSolverHQuadProgFast
//   inria_wbc::Exception                    include/inria_wbc/exceptions.hpp:52-65                 -> inria_wbc::Exception
// Arrays are instance-major std::vector<double> (row i = robot i), the layout of the C ABI.
#ifndef INRIA_WBC_B200_BATCH_POS_TRACKER_HPP
#define INRIA_WBC_B200_BATCH_POS_TRACKER_HPP

#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

extern "C" {
#include "../iwbc_batch.h"
}

namespace inria_wbc {

    // include/inria_wbc/exceptions.hpp:52-65: runtime_error carrying message + file:line
    class Exception : public std::runtime_error {
    public:
        Exception(const std::string& msg, const char* file, int line) : std::runtime_error(make(msg, file, line)) {}

    private:
        static std::string make(const std::string& msg, const char* file, int line)
        {
            std::ostringstream o;
            o << msg << " (" << file << ":" << line << ")";
            return o.str();
        }
    };
#define IWBC_B200_EXCEPTION(msg) ::inria_wbc::Exception((msg), __FILE__, __LINE__)
#define IWBC_B200_ASSERT(cond, msg)       \
    do {                                  \
        if (!(cond))                      \
            throw IWBC_B200_EXCEPTION(msg); \
    } while (0)

    namespace controllers {

        // HQP status codes surfaced at src/controllers/controller.cpp:287-304
        enum Status { UNKNOWN = -1, OPTIMAL = 0, INFEASIBLE = 1, UNBOUNDED = 2, MAX_ITER_REACHED = 3, ERROR = 4 };

        // Stands in for the `YAML::Node` the reference hands to Factory::create (yaml-cpp is not a dependency of this library): the file
        // the node was loaded from plus the keys the caller mutated afterwards.
        struct Config {
            std::string controller_yaml;      // what the reference loads with YAML::LoadFile before Factory::create
            std::string robots_path;          // where CONTROLLER.urdf is resolved (utheque's role, controller.cpp:71-79)
            std::string etc_path;             // replaces "@ETC@" in CONTROLLER.base_path ("" = next to the yaml)
            std::string overrides_yaml;       // CONTROLLER keys mutated after loading, as tests/test_determinism.cpp:66-70 does
            int batch = 1;
            int device = 0;
            bool throw_on_failure = false;    // true: update() throws like Controller::_solve (controller.cpp:284-308) when a robot's QP is not OPTIMAL
        };

        // controller.hpp:39 with instance-major matrices: key -> [batch][dim] row-major.  Keys read here: "positions" (na),
        // "joint_velocities" (na), "floating_base_position" (7: xyz + quaternion xyzw), "floating_base_velocity" (6), and for a
        // controller whose stabilizer is activated "lf_force", "rf_force", "lf_torque", "rf_torque", "imu_vel" (3 each).
        using SensorData = std::unordered_map<std::string, std::vector<double>>;

        class BatchController {
        public:
            explicit BatchController(const Config& c) : config_(c)
            {
                prob_ = iwbc_problem_load(c.controller_yaml.c_str(), c.robots_path.empty() ? nullptr : c.robots_path.c_str(),
                    c.etc_path.empty() ? nullptr : c.etc_path.c_str(), c.overrides_yaml.empty() ? nullptr : c.overrides_yaml.c_str());
                if (!prob_)
                    throw IWBC_B200_EXCEPTION(std::string("cannot build the controller: ") + iwbc_last_error());
                IWBC_B200_ASSERT(c.batch > 0, "batch must be positive");
                iwbc_problem_dims(prob_, &nq_, &nv_, &na_, &nc_, &nref_);
                B_ = c.batch;
                dt_ = iwbc_problem_data(prob_)->dt;
                std::vector<double> q0(nq_), r0(nref_);
                iwbc_problem_q0(prob_, q0.data());
                iwbc_problem_default_refs(prob_, r0.data());
                q_tsid_.resize((size_t)B_ * nq_);
                refs_.resize((size_t)B_ * nref_);
                for (int i = 0; i < B_; ++i) {
                    std::copy(q0.begin(), q0.end(), q_tsid_.begin() + (size_t)i * nq_);
                    std::copy(r0.begin(), r0.end(), refs_.begin() + (size_t)i * nref_);
                }
                v_tsid_.assign((size_t)B_ * nv_, 0.0);
                mask_.assign(B_, iwbc_problem_default_contact_mask(prob_));
                tau_.assign((size_t)B_ * na_, 0.0), a_tsid_.assign((size_t)B_ * nv_, 0.0), q_.assign((size_t)B_ * nv_, 0.0);
                forces_.assign((size_t)B_ * 12 * (nc_ > 0 ? nc_ : 1), 0.0), momentum_.assign((size_t)B_ * 3, 0.0);
                status_.assign(B_, UNKNOWN), iters_.assign(B_, 0);
                qn_.resize(q_tsid_.size()), vn_.resize(v_tsid_.size());
                btype_.assign(B_, IWBC_BT_FIXED_BASE); // humanoid_pos_tracker.cpp:39
                closed_loop_ = iwbc_problem_closed_loop(prob_) != 0;
            }
            BatchController(const BatchController&) = delete; // controller.hpp:50-51
            BatchController& operator=(const BatchController&) = delete;
            virtual ~BatchController()
            {
                if (stab_)
                    iwbc_stabilizer_free(stab_);
                if (batch_)
                    iwbc_batch_free(batch_);
                if (prob_)
                    iwbc_problem_free(prob_);
            }

            // Controller::update(sensor_data), closed-loop form (controller.cpp:172-205): q[B][nq] = [floating_base_position(7); positions],
            // dq[B][nv] = [floating_base_velocity(6); joint_velocities].  Size mismatch throws like the reference's asserts.
            virtual void update(const std::vector<double>& q, const std::vector<double>& dq)
            {
                IWBC_B200_ASSERT(q.size() == (size_t)B_ * nq_, "update: q must have batch * nq entries");
                IWBC_B200_ASSERT(dq.size() == (size_t)B_ * nv_, "update: dq must have batch * nv entries");
                solve(q.data(), dq.data());
            }
            // Controller::update() open loop (controller.cpp:163-170): the controller's own integrated state
            virtual void update() { solve(q_tsid_.data(), v_tsid_.data()); }
            // Controller::update(const SensorData&) (controller.hpp:57, controller.cpp:161-206; HumanoidPosTracker::update,
            // humanoid_pos_tracker.cpp:134-302): closed loop reads the state from the sensors, open loop the controller's own; an
            // activated stabilizer needs the F/T and IMU keys ("the stabilizer needs the LF force", :180-184)
            virtual void update(const SensorData& sensors)
            {
                const iwbc_problem_t* P = iwbc_problem_data(prob_);
                const double *q = q_tsid_.data(), *dq = v_tsid_.data();
                if (closed_loop_) {
                    const auto& pos = need(sensors, "positions", "we need the joint positions in closed loop mode!");
                    const auto& vel = need(sensors, "joint_velocities", "we need the joint velocities in closed loop mode!");
                    IWBC_B200_ASSERT(pos.size() == (size_t)B_ * na_, "Joint positions do not have the correct size");
                    IWBC_B200_ASSERT(vel.size() == (size_t)B_ * na_, "Joint velocities do not have the correct size");
                    qs_.resize((size_t)B_ * nq_), vs_.resize((size_t)B_ * nv_);
                    const std::vector<double>*fbp = nullptr, *fbv = nullptr;
                    if (nv_ > na_) {
                        fbp = &need(sensors, "floating_base_position", "we need the floating base position in closed loop mode!");
                        fbv = &need(sensors, "floating_base_velocity", "we need the floating base velocity in closed loop mode!");
                        IWBC_B200_ASSERT(fbp->size() == (size_t)B_ * 7 && fbv->size() == (size_t)B_ * 6, "floating base sensors do not have the correct size");
                    }
                    for (int i = 0; i < B_; ++i) {
                        double* qi = &qs_[(size_t)i * nq_];
                        double* vi = &vs_[(size_t)i * nv_];
                        if (fbp) {
                            std::copy(fbp->begin() + 7 * i, fbp->begin() + 7 * i + 7, qi);
                            std::copy(fbv->begin() + 6 * i, fbv->begin() + 6 * i + 6, vi);
                        }
                        std::copy(pos.begin() + (size_t)i * na_, pos.begin() + (size_t)(i + 1) * na_, qi + (nq_ - na_));
                        std::copy(vel.begin() + (size_t)i * na_, vel.begin() + (size_t)(i + 1) * na_, vi + (nv_ - na_));
                    }
                    q = qs_.data(), dq = vs_.data();
                }
                if (P->stabilizer_activated) {
                    const char* keys[5] = {"lf_force", "rf_force", "lf_torque", "rf_torque", "imu_vel"};
                    const char* msgs[5] = {"the stabilizer needs the LF force", "the stabilizer needs the RF force", "the stabilizer needs the LF torque",
                        "the stabilizer needs the RF torque", "the stabilizer imu needs angular velocity"};
                    sens_.resize((size_t)B_ * 15);
                    for (int k = 0; k < 5; ++k) {
                        const auto& sv = need(sensors, keys[k], msgs[k]);
                        IWBC_B200_ASSERT(sv.size() == (size_t)B_ * 3, std::string(keys[k]) + " needs batch * 3 entries");
                        for (int i = 0; i < B_; ++i)
                            std::copy(sv.begin() + 3 * i, sv.begin() + 3 * i + 3, sens_.begin() + (size_t)i * 15 + 3 * k);
                    }
                    solve(q, dq, sens_.data());
                }
                else
                    solve(q, dq);
            }
            // Controller::set_behavior_type (controller.cpp:436-443): selects the stabilizer parameter set (and filter windows) of a robot
            void set_behavior_type(int i, int behavior_type)
            {
                IWBC_B200_ASSERT(behavior_type >= IWBC_BT_FIXED_BASE && behavior_type <= IWBC_BT_DOUBLE_SUPPORT,
                    "behavior type is either behavior_types::FIXED_BASE , behavior_types::SINGLE_SUPPORT or behavior_types::DOUBLE_SUPPORT ");
                btype_.at(i) = behavior_type;
            }

            // filter_mimics = true forms of the getters (controller.hpp:117-121, controller.cpp:372-400): the entries of non_mimic_indexes()
            std::vector<double> q(bool filter_mimics) const { return filter_mimics ? slice(q_, nv_) : q_; }
            std::vector<double> dq(bool filter_mimics) const { return filter_mimics ? slice(v_tsid_, nv_) : v_tsid_; }
            std::vector<double> ddq(bool filter_mimics) const { return filter_mimics ? slice(a_tsid_, nv_) : a_tsid_; }
            // (tau_ of the reference carries six leading zeros for a floating base: controller.cpp:268,275; here [B][na])
            // results, instance-major (controller.hpp:117-122, controller.cpp:250-282)
            const std::vector<double>& tau() const { return tau_; }             // [B][na]
            const std::vector<double>& ddq() const { return a_tsid_; }          // [B][nv]
            const std::vector<double>& dq() const { return v_tsid_; }           // [B][nv]
            const std::vector<double>& q() const { return q_; }                 // [B][nv] base as angle-axis (the "DART" layout)
            const std::vector<double>& q_tsid() const { return q_tsid_; }       // [B][nq]
            const std::vector<double>& activated_contacts_forces() const { return forces_; } // [B][12 * ncontacts]
            const std::vector<double>& momentum() const { return momentum_; }   // [B][3]
            // Controller::cost(task_name) (controller.hpp:148-153, pos_tracker.hpp:43) and PosTracker::objective_value()
            // (pos_tracker.hpp:44) of robot i at its last solve.  Logging getters in the reference (src/robot_dart/talos.cpp:502-505):
            // re-evaluated on demand from a stage dump of that ONE robot at the state and references its QP was solved at (the
            // arrays given to update(q, dq) must still hold that state; update() keeps its own).
            double cost(const std::string& task_name, int i = 0) const
            {
                IWBC_B200_ASSERT(solved_q_ && i >= 0 && i < B_, "cost(): no solve yet, or robot index out of range");
                double c = 0;
                if (iwbc_batch_task_cost(batch_, solved_q_ + (size_t)i * nq_, solved_v_ + (size_t)i * nv_, refs_.data() + (size_t)i * nref_, mask_[i],
                        task_name.c_str(), a_tsid_.data() + (size_t)i * nv_, tau_.data() + (size_t)i * na_, &c)
                    != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("cost(") + task_name + "): " + iwbc_last_error());
                return c;
            }
            double objective_value(int i = 0) const
            {
                IWBC_B200_ASSERT(solved_q_ && i >= 0 && i < B_, "objective_value(): no solve yet, or robot index out of range");
                double f = 0;
                if (iwbc_batch_objective_value(batch_, solved_q_ + (size_t)i * nq_, solved_v_ + (size_t)i * nv_, refs_.data() + (size_t)i * nref_, mask_[i], &f)
                    != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("objective_value(): ") + iwbc_last_error());
                return f;
            }
            const std::vector<int>& status() const { return status_; }          // per-instance HQP status (data, not an exception)
            const std::vector<int>& solver_iterations() const { return iters_; }
            bool all_optimal() const
            {
                for (int s : status_)
                    if (s != OPTIMAL)
                        return false;
                return true;
            }
            double t() const { return t_; }
            double dt() const { return dt_; }
            int batch() const { return B_; }
            int nq() const { return nq_; }
            int nv() const { return nv_; }
            int na() const { return na_; }
            std::vector<std::string> controllable_dofs(bool filter_mimics = true) const
            {
                std::vector<std::string> all = names(0), mim = names(3), out;
                for (const auto& n : all) {
                    bool is_mimic = false;
                    for (const auto& m : mim)
                        is_mimic = is_mimic || m == n;
                    if (!filter_mimics || !is_mimic)
                        out.push_back(n);
                }
                return out;
            }
            const iwbc_problem* problem() const { return prob_; }

        protected:
            std::vector<std::string> names(int kind) const
            {
                std::vector<std::string> out;
                const int n = iwbc_problem_name(prob_, kind, 0, nullptr, 0);
                char buf[256];
                for (int i = 0; i < n; ++i) {
                    iwbc_problem_name(prob_, kind, i, buf, sizeof(buf));
                    out.emplace_back(buf);
                }
                return out;
            }
        public:
            // `ticks` closed device-side ticks of a compiled behaviour on the controller's own (open-loop) state: what a loop of
            // Behavior::update() calls does in the reference, without crossing PCIe in between.  Results = those of the last tick.
            // cmd_trace (optional): resized to [ticks][B][nv], every tick's q(false) -- what the reference's tests record after each update
            void run_behavior(iwbc_behavior* h, const int* phase, int tick0, int ticks, std::vector<double>* cmd_trace = nullptr)
            {
                ensure_batch();
                iwbc_results_t out = {tau_.data(), a_tsid_.data(), forces_.data(), nullptr, nullptr, q_.data(), momentum_.data(), status_.data(),
                    iters_.data(), nullptr};
                int rc;
                if (cmd_trace) {
                    cmd_trace->assign((size_t)ticks * B_ * nv_, 0.0);
                    rc = iwbc_batch_behavior_run_trace(batch_, h, B_, q_tsid_.data(), v_tsid_.data(), phase, tick0, ticks, &out, mask_.data(), cmd_trace->data());
                }
                else
                    rc = iwbc_batch_behavior_run(batch_, h, B_, q_tsid_.data(), v_tsid_.data(), phase, tick0, ticks, &out, mask_.data());
                if (rc != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("behaviour roll-out failed: ") + iwbc_last_error());
                t_ += dt_ * ticks;
            }
            // Controller::filter_cmd / the filter_mimics = true getters (controller.cpp:372-400): positions of `row` ([nv] "DART" layout,
            // 6 base entries first) that are not mimic dofs
            std::vector<int> non_mimic_indexes() const
            {
                std::vector<std::string> all = names(0), mim = names(3);
                std::vector<int> idx;
                for (int k = 0; k < nv_ - na_; ++k)
                    idx.push_back(k);
                for (int k = 0; k < na_; ++k) {
                    bool is_mimic = false;
                    for (const auto& m : mim)
                        is_mimic = is_mimic || m == all[k];
                    if (!is_mimic)
                        idx.push_back(nv_ - na_ + k);
                }
                return idx;
            }
            iwbc_problem* problem() { return prob_; }
            const std::vector<unsigned>& contact_masks() const { return mask_; } // activated contacts per instance (bit = contact index)

        protected:
            void ensure_batch()
            {
                if (!batch_) {
                    batch_ = iwbc_batch_create(prob_, config_.device, B_);
                    if (!batch_)
                        throw IWBC_B200_EXCEPTION(std::string("cannot create the GPU batch: ") + iwbc_last_error());
                }
            }
            void solve(const double* q, const double* dq, const double* sensors15 = nullptr)
            {
                ensure_batch();
                iwbc_results_t out = {tau_.data(), a_tsid_.data(), forces_.data(), qn_.data(), vn_.data(), q_.data(), momentum_.data(), status_.data(),
                    iters_.data(), nullptr};
                int rc;
                if (sensors15) {
                    if (!stab_) {
                        stab_ = iwbc_stabilizer_create(batch_);
                        if (!stab_)
                            throw IWBC_B200_EXCEPTION(std::string("cannot create the stabilizer state: ") + iwbc_last_error());
                    }
                    rc = iwbc_batch_update_sensors(batch_, stab_, B_, q, dq, refs_.data(), nref_, mask_.data(), btype_.data(), 0, sensors15, &out, nullptr);
                }
                else
                    rc = iwbc_batch_update(batch_, B_, q, dq, refs_.data(), nref_, mask_.data(), &out);
                if (rc != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("update failed: ") + iwbc_last_error());
                if (config_.throw_on_failure)
                    for (int i = 0; i < B_; ++i)
                        if (status_[i] != OPTIMAL) { // the reference's message (controller.cpp:285-307), with the robot's index
                            static const char* what[] = {" => Unknown", "", " => Infeasible ", " => Unbounded ", " => Max iter reached ", " => Error "};
                            const int s = status_[i];
                            throw IWBC_B200_EXCEPTION("Controller failed, can't solve problem. Status : " + std::to_string(s)
                                + (s >= -1 && s <= 4 ? what[s + 1] : " => Uknown status") + " (t=" + std::to_string(t_) + ", robot " + std::to_string(i) + ")");
                        }
                solved_q_ = q, solved_v_ = dq; // (update(): these are the buffers that become qn_ / vn_ below and keep the solved state)
                q_tsid_.swap(qn_);
                v_tsid_.swap(vn_);
                t_ += dt_;
            }
            static const std::vector<double>& need(const SensorData& s, const char* key, const char* msg)
            {
                auto it = s.find(key);
                if (it == s.end())
                    throw IWBC_B200_EXCEPTION(msg);
                return it->second;
            }
            std::vector<double> slice(const std::vector<double>& full, int width) const
            {
                const std::vector<int> idx = non_mimic_indexes();
                std::vector<double> out;
                out.reserve((size_t)B_ * idx.size());
                for (int i = 0; i < B_; ++i)
                    for (int k : idx)
                        out.push_back(full[(size_t)i * width + k]);
                return out;
            }

            Config config_;
            iwbc_problem* prob_ = nullptr;
            iwbc_batch* batch_ = nullptr;
            iwbc_stabilizer* stab_ = nullptr;
            bool closed_loop_ = false;
            std::vector<double> qs_, vs_, sens_;
            std::vector<int> btype_;
            int B_ = 0, nq_ = 0, nv_ = 0, na_ = 0, nc_ = 0, nref_ = 0;
            double t_ = 0.0, dt_ = 0.0;
            std::vector<double> q_tsid_, v_tsid_, a_tsid_, tau_, q_, forces_, momentum_, refs_, qn_, vn_;
            std::vector<int> status_, iters_;
            std::vector<unsigned> mask_;
            const double *solved_q_ = nullptr, *solved_v_ = nullptr; // state of the last solve (cost / objective_value)
        };

        // PosTracker's reference setters, per instance (pos_tracker.hpp:52-67, pos_tracker.cpp:220-260)
        class BatchPosTracker : public BatchController {
        public:
            explicit BatchPosTracker(const Config& c) : BatchController(c) {}

            bool has_task(const std::string& n) const { return slot(n, false).second > 0; }
            bool has_contact(const std::string& n) const { return iwbc_problem_contact_index(prob_, n.c_str()) >= 0; }
            // TrajectorySample layout (include/inria_wbc/trajs/utils.hpp:28-36): pos = p(3) ++ R column-major(9), vel(6), acc(6)
            void set_se3_ref(int i, const double* pos12, const std::string& task_name, const double* vel6 = nullptr, const double* acc6 = nullptr)
            {
                double* r = ref(i, task_name, 24);
                std::copy(pos12, pos12 + 12, r);
                for (int k = 0; k < 6; ++k) {
                    r[12 + k] = vel6 ? vel6[k] : 0.0;
                    r[18 + k] = acc6 ? acc6[k] : 0.0;
                }
            }
            void set_contact_se3_ref(int i, const double* pos12, const std::string& contact_name) { set_se3_ref(i, pos12, contact_name); }
            void get_se3_ref(int i, const std::string& task_name, double* pos12) const
            {
                const double* r = const_cast<BatchPosTracker*>(this)->ref(i, task_name, 24);
                std::copy(r, r + 12, pos12);
            }
            void set_com_ref(int i, const double* com3, const double* vel3 = nullptr, const double* acc3 = nullptr)
            {
                double* r = ref(i, "com", 9);
                for (int k = 0; k < 3; ++k) {
                    r[k] = com3[k];
                    r[3 + k] = vel3 ? vel3[k] : 0.0;
                    r[6 + k] = acc3 ? acc3[k] : 0.0;
                }
            }
            void get_com_ref(int i, double* com3) const
            {
                const double* r = const_cast<BatchPosTracker*>(this)->ref(i, "com", 9);
                std::copy(r, r + 3, com3);
            }
            void set_momentum_ref(int i, const double* L6, const double* dL6)
            {
                double* r = ref(i, "momentum", 18);
                for (int k = 0; k < 6; ++k) {
                    r[6 + k] = L6[k];
                    r[12 + k] = dL6 ? dL6[k] : 0.0;
                }
            }
            void set_posture_ref(int i, const double* q_na, const std::string& task_name = "posture")
            {
                double* r = ref(i, task_name, 3 * na_);
                std::copy(q_na, q_na + na_, r);
            }
            // add_contact / remove_contact (pos_tracker.cpp:246-287): the contact's bit in the instance's mask; instances with
            // different masks are bucketed inside update()
            void add_contact(int i, const std::string& contact_name) { mask_.at(i) |= 1u << contact_bit(contact_name); }
            void remove_contact(int i, const std::string& contact_name) { mask_.at(i) &= ~(1u << contact_bit(contact_name)); }
            // set_task_weight (pos_tracker.hpp:67): shared by the whole batch (it is part of the compiled problem)
            void set_task_weight(const std::string& task_name, double w)
            {
                if (iwbc_problem_set_task_weight(prob_, task_name.c_str(), w) != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("set_task_weight: ") + iwbc_last_error());
            }
            std::vector<std::string> tasks() const { return names(1); }
            std::vector<std::string> contacts() const { return names(2); }

        private:
            std::pair<int, int> slot(const std::string& n, bool must) const
            {
                int off = -1, len = 0;
                if (iwbc_problem_ref_slot(prob_, n.c_str(), &off, &len) != IWBC_OK) {
                    if (must)
                        throw IWBC_B200_EXCEPTION("Task [" + n + "] not found"); // pos_tracker.hpp:34-39
                    return {-1, 0};
                }
                return {off, len};
            }
            double* ref(int i, const std::string& n, int need)
            {
                IWBC_B200_ASSERT(i >= 0 && i < B_, "instance index out of range");
                auto s = slot(n, true);
                IWBC_B200_ASSERT(s.second >= need, "Task [" + n + "] has no reference of this type");
                return refs_.data() + (size_t)i * nref_ + s.first;
            }
            int contact_bit(const std::string& n) const
            {
                const int c = iwbc_problem_contact_index(prob_, n.c_str());
                if (c < 0)
                    throw IWBC_B200_EXCEPTION("Contact [" + n + "] not found"); // pos_tracker.hpp:28-33
                return c;
            }
        };

        // utils::Factory (factory.hpp:14-76): name -> creator; unknown name throws listing what is registered
        class Factory {
        public:
            using creator_t = std::function<std::shared_ptr<BatchController>(const Config&)>;
            static Factory& instance()
            {
                static Factory f;
                return f;
            }
            void register_creator(const std::string& name, creator_t c) { map_[name] = c; }
            // AutoRegister (factory.hpp:66-76): `static Register<MyController> __my("my-controller");` in the plugin's translation unit
            template <class T>
            struct AutoRegister {
                explicit AutoRegister(const std::string& name)
                {
                    Factory::instance().register_creator(name, [](const Config& c) { return std::static_pointer_cast<BatchController>(std::make_shared<T>(c)); });
                }
            };
            std::shared_ptr<BatchController> create(const std::string& name, const Config& c) const
            {
                auto it = map_.find(name);
                if (it == map_.end()) {
                    std::string msg = "Error :" + name + " is not in the factory. Available:";
                    for (const auto& kv : map_)
                        msg += " " + kv.first;
                    throw IWBC_B200_EXCEPTION(msg);
                }
                return it->second(c);
            }

        private:
            Factory()
            {
                for (const char* n : {"pos-tracker", "humanoid-pos-tracker", "talos-pos-tracker"})
                    map_[n] = [](const Config& c) { return std::make_shared<BatchPosTracker>(c); };
            }
            std::map<std::string, creator_t> map_;
        };
        template <class T>
        using Register = Factory::AutoRegister<T>; // controller.hpp:236-237
    } // namespace controllers

    namespace behaviors {
        // behaviors::Behavior (include/inria_wbc/behaviors/behavior.hpp:17-30): constructed from a controller and the behaviour's YAML,
        // `update()` = set this tick's references / contacts, then controller->update().  Here the behaviour is compiled into a table the
        // GPU evaluates; every instance may sit at its own phase of the behaviour.
        class BatchBehavior {
        public:
            using controller_ptr_t = std::shared_ptr<controllers::BatchController>;
            BatchBehavior(const controller_ptr_t& controller, const std::string& behavior_yaml, const std::string& overrides_yaml = "")
                : controller_(controller)
            {
                IWBC_B200_ASSERT(controller_ != nullptr, "Invalid controller pointer"); // behavior.cpp:25
                h_ = iwbc_behavior_load(controller_->problem(), behavior_yaml.c_str(), overrides_yaml.empty() ? nullptr : overrides_yaml.c_str());
                if (!h_)
                    throw IWBC_B200_EXCEPTION(std::string("cannot build the behavior: ") + iwbc_last_error());
                iwbc_behavior_info(h_, &total_ticks_, nullptr, nullptr, nullptr, nullptr);
            }
            BatchBehavior(const BatchBehavior&) = delete;
            BatchBehavior& operator=(const BatchBehavior&) = delete;
            virtual ~BatchBehavior() { iwbc_behavior_free(h_); }
            // per-instance offset into the behaviour (instance i runs tick `t + phase[i]`); default: all in step
            void set_phase(const std::vector<int>& phase)
            {
                IWBC_B200_ASSERT((int)phase.size() == controller_->batch(), "set_phase: one offset per instance");
                phase_ = phase;
            }
            // `ticks` behaviour + controller updates on the device; cmd_trace (optional) receives q(false) after every one of them
            virtual void update(int ticks = 1, std::vector<double>* cmd_trace = nullptr)
            {
                controller_->run_behavior(h_, phase_.empty() ? nullptr : phase_.data(), tick_, ticks, cmd_trace);
                tick_ += ticks;
            }
            int tick() const { return tick_; }
            int total_ticks() const { return total_ticks_; }
            controller_ptr_t controller() { return controller_; }

        private:
            controller_ptr_t controller_;
            iwbc_behavior* h_ = nullptr;
            std::vector<int> phase_;
            int tick_ = 0, total_ticks_ = 0;
        };
    } // namespace behaviors

    namespace solvers {
        // tsid::solvers::SolverHQPBase through the three virtuals the reference uses (resize / solve / status), for B stacked QPs:
        //   min 1/2 x'Hx + g'x  s.t.  CE x + ce0 = 0,  CI x + ci0 >= 0   (row-major, instance-major)
        struct BatchHQPOutput {
            std::vector<double> x;
            std::vector<int> status, iterations;
            std::vector<unsigned> activeSet; // [B][IWBC_ACTIVE_WORDS] bitmask over CI rows
        };
        class BatchSolverHQuadProgFast {
        public:
            explicit BatchSolverHQuadProgFast(const std::string& name = "solver-eiquadprog", int device = 0) : name_(name), device_(device) {}
            void resize(unsigned n, unsigned neq, unsigned nin) { n_ = n, neq_ = neq, nin_ = nin; }
            const BatchHQPOutput& solve(int B, const double* H, const double* g, const double* CE, const double* ce0, const double* CI, const double* ci0)
            {
                out_.x.assign((size_t)B * n_, 0.0), out_.status.assign(B, -1), out_.iterations.assign(B, 0);
                out_.activeSet.assign((size_t)B * IWBC_ACTIVE_WORDS, 0u);
                if (iwbc_qp_solve_batch(device_, B, n_, neq_, nin_, H, g, CE, ce0, CI, ci0, out_.x.data(), out_.status.data(), out_.iterations.data(),
                        out_.activeSet.data())
                    != IWBC_OK)
                    throw IWBC_B200_EXCEPTION(std::string("solve failed: ") + iwbc_last_error());
                return out_;
            }
            const std::string& name() const { return name_; }

        private:
            std::string name_;
            int device_;
            unsigned n_ = 0, neq_ = 0, nin_ = 0;
            BatchHQPOutput out_;
        };
    } // namespace solvers
} // namespace inria_wbc
#endif
