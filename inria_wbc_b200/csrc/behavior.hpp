// behavior.hpp -- behaviours (the per-tick reference generators that call the hot path) compiled into a SEGMENT TABLE
// that a device kernel can evaluate for any (instance, tick) in closed form.
//
// The reference pre-computes std::vector tables of waypoints at construction and indexes them with (current trajectory,
// time) every tick on the host (src/behaviors/humanoid/walk.cpp:42-246, src/behaviors/humanoid/move_com.cpp:8-61,
// src/behaviors/generic/cartesian.cpp:8-94).  Every table entry is a constant or a minimum-jerk polynomial between two
// waypoints (include/inria_wbc/trajs/trajectory_generator.hpp:29-160), so here a behaviour is
//     segments  x  channels  ->  { start, end, duration, kind }        (+ the contact mask of each segment)
// and tick t of instance i is evaluated on the GPU from its own phase offset: no per-tick host work and no per-tick
// upload of references (1.2 kB / instance / tick for Talos).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/iwbc_batch.h"
#include "host_state.hpp"
#include "stack.hpp"
#include "yaml_lite.hpp"

// one channel (a task reference the behaviour drives) inside one segment; POD, shared with the device
struct iwbc_beh_chanseg_t {
    double p0[3], p1[3]; // translation (SE3) or vector (CoM) waypoints
    double R0[9];        // SE3: start rotation in the reference-vector layout (column-major, as TrajectorySample holds it)
    double axis[3];      // SE3: axis / angle of R0' R1 (Eigen::AngleAxisd, trajectory_generator.hpp:124-126)
    double angle;
    double duration;     // trajectory_duration of THIS channel's polynomial (seconds)
    int minjerk;         // 0: constant_traj(p0)   1: min_jerk_trajectory(p0 -> p1)
    int rot;             // 1: the rotation is interpolated (angle != 0)
};
struct iwbc_beh_channel_t {
    int kind;    // 0: 3-vector reference (CoM: pos(3) vel(3) acc(3));  1: SE3 reference (pos(12) vel(6) acc(6))
    int ff;      // write the polynomial's first / second derivative into vel / acc (cartesian, move_com); 0: zeros (walk)
    int slot[2]; // offsets into the reference vector (task, and the contact that shares the frame); -1: none
    // recorded-trajectory channels (generic::cartesian_traj): rows of 3 (CoM) or 12 (translation + column-major rotation) doubles,
    // replayed forwards then backwards; table_rows == 0: a waypoint channel (segments)
    int table_off, table_rows;
    double scale; // applied to the translation / CoM (cartesian_traj.cpp:24-35)
};
struct iwbc_beh_seg_t {
    int start, ticks;   // first global tick, length
    unsigned mask;      // active contacts while the segment runs
    unsigned mask_last; // ... at its last tick (walk adds the landing foot's contact before that tick's update)
};

struct iwbc_behavior {
    std::string name;
    const iwbc_problem* prob = nullptr;
    double dt = 0;
    int total_ticks = 0;
    bool loop = false;
    int loop_start = 0; // a looping behaviour repeats the ticks [loop_start, total_ticks) (walk-on-spot plays its INIT segment once)
    std::vector<iwbc_beh_channel_t> channels;
    std::vector<iwbc_beh_seg_t> segs;
    std::vector<iwbc_beh_chanseg_t> cs; // [seg][channel]
    std::vector<unsigned> masks;        // distinct contact masks, most frequent first
    std::vector<double> table;          // recorded trajectories of the table channels
    // device copies (one device; created on first use)
    int device = -1;
    void* d_blob = nullptr;
    iwbc_beh_channel_t* d_channels = nullptr;
    iwbc_beh_seg_t* d_segs = nullptr;
    iwbc_beh_chanseg_t* d_cs = nullptr;
    double* d_ref0 = nullptr;
    double* d_table = nullptr;
    int *d_bcount = nullptr, *d_blist = nullptr;
    int blist_cap = 0;
};

namespace iwbc {
namespace host {

    inline int beh_n_steps(double dt, double duration) { return (int)std::floor(duration / dt); } // trajectory_generator.hpp:73

    struct BehaviorBuilder {
        iwbc_behavior& b;
        const iwbc_problem& prob;
        explicit BehaviorBuilder(iwbc_behavior& b_, const iwbc_problem& p) : b(b_), prob(p)
        {
            b.prob = &p;
            b.dt = p.cs.prob.dt;
        }
        int slot_of(const std::string& name) const
        {
            int off = -1, len = 0;
            if (iwbc_problem_ref_slot(&prob, name.c_str(), &off, &len) != IWBC_OK)
                throw ConfigError("behavior: the controller has no task or contact named '" + name + "'");
            return off;
        }
        int add_channel(int kind, int ff, int slot0, int slot1)
        {
            b.channels.push_back({kind, ff, {slot0, slot1}, 0, 0, 1.0});
            return (int)b.channels.size() - 1;
        }
        void begin_segment(int ticks, unsigned mask, unsigned mask_last)
        {
            b.segs.push_back({b.total_ticks, ticks, mask, mask_last});
            b.total_ticks += ticks;
            b.cs.resize(b.segs.size() * b.channels.size());
        }
        iwbc_beh_chanseg_t& at(int ch) { return b.cs[(b.segs.size() - 1) * b.channels.size() + ch]; }
        // translation-only waypoints; the rotation (SE3 channels) stays the construction-time one
        void set(int ch, const double* p0, const double* p1, double duration, bool minjerk)
        {
            iwbc_beh_chanseg_t& c = at(ch);
            std::memset(&c, 0, sizeof(c));
            for (int k = 0; k < 3; ++k)
                c.p0[k] = p0[k], c.p1[k] = minjerk ? p1[k] : p0[k];
            const int s = b.channels[ch].slot[0];
            if (b.channels[ch].kind == 1)
                for (int k = 0; k < 9; ++k)
                    c.R0[k] = prob.cs.ref0[s + 3 + k];
            c.axis[0] = 1.0;
            c.duration = duration;
            c.minjerk = minjerk ? 1 : 0;
        }
        void finish()
        {
            // distinct masks by how many ticks they cover
            std::vector<std::pair<long, unsigned>> cnt;
            auto bump = [&](unsigned m, long n) {
                for (auto& e : cnt)
                    if (e.second == m) {
                        e.first += n;
                        return;
                    }
                cnt.push_back({n, m});
            };
            for (auto& s : b.segs) {
                bump(s.mask, s.ticks - 1);
                bump(s.mask_last, 1);
            }
            std::sort(cnt.begin(), cnt.end(), [](const std::pair<long, unsigned>& a, const std::pair<long, unsigned>& c) { return a.first > c.first; });
            for (auto& e : cnt)
                b.masks.push_back(e.second);
            if (b.masks.size() > 8)
                throw ConfigError("behavior: more than 8 distinct contact masks");
        }
    };

    // humanoid::walk (src/behaviors/humanoid/walk.cpp).  The walk is a list of states; each state moves at most one foot,
    // the CoM or one hand between two waypoints and holds everything else.  States are described here by which waypoint
    // pair each of the five channels follows; the waypoint bookkeeping (feet leap-frogging by step_length, CoM over the
    // support foot, hands following the opposite... same-side foot) is walk.cpp:63-196.
    inline void compile_walk(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c)
    {
        BehaviorBuilder bb(out, prob);
        const double com_dur = IWBC_HCHECK(c["traj_com_duration"].as<float>(), "traj_com_duration");
        const double foot_dur = IWBC_HCHECK(c["traj_foot_duration"].as<float>(), "traj_foot_duration");
        const float step_height = IWBC_HCHECK(c["step_height"].as<float>(), "step_height");
        const float step_length = IWBC_HCHECK(c["step_length"].as<float>(), "step_length");
        const int ncyc = IWBC_HCHECK(c["num_of_cycles"].as<int>(), "num_of_cycles");
        if (ncyc <= 0)
            throw ConfigError("num_of_cycles needs to be more than 0"); // walk.cpp:34-35
        const char* need[] = {"lf", "rf", "lh", "rh", "com", "contact_lfoot", "contact_rfoot"}; // walk.cpp:14-20
        for (const char* n : need)
            bb.slot_of(n);
        const int lfb = iwbc_problem_contact_index(&prob, "contact_lfoot"), rfb = iwbc_problem_contact_index(&prob, "contact_rfoot");
        const unsigned LF = 1u << lfb, RF = 1u << rfb, DS = iwbc_problem_default_contact_mask(&prob) | LF | RF;
        enum { COM, LFT, RFT, LH, RH };
        bb.add_channel(0, 0, bb.slot_of("com"), -1);
        bb.add_channel(1, 0, bb.slot_of("lf"), bb.slot_of("contact_lfoot")); // set_se3_ref + set_contact_se3_ref (walk.cpp:224-227)
        bb.add_channel(1, 0, bb.slot_of("rf"), bb.slot_of("contact_rfoot"));
        bb.add_channel(1, 0, bb.slot_of("lh"), -1);
        bb.add_channel(1, 0, bb.slot_of("rh"), -1);

        struct P3 {
            double v[3];
        };
        auto ref_pos = [&](int ch) {
            P3 p;
            for (int k = 0; k < 3; ++k)
                p.v[k] = prob.cs.ref0[out.channels[ch].slot[0] + k];
            return p;
        };
        P3 lf_low = ref_pos(LFT), rf_low = ref_pos(RFT), com_mid = ref_pos(COM), lh = ref_pos(LH), rh = ref_pos(RH);
        P3 lf_high = lf_low, rf_high = rf_low;
        lf_high.v[2] += step_height;
        rf_high.v[2] += step_height;
        P3 com_lf = lf_low, com_rf = rf_low; // CoM over a foot, at the initial height
        com_lf.v[2] = com_rf.v[2] = com_mid.v[2];
        P3 lh_fwd = lh, rh_fwd = rh;

        const int com_ticks = beh_n_steps(out.dt, com_dur), foot_ticks = beh_n_steps(out.dt, foot_dur);
        // one state: (segment length, masks, and per channel: from, to, duration, moving?)
        struct Move {
            const P3 *from, *to;
        };
        auto emit = [&](bool foot_state, unsigned mask, unsigned mask_last, Move com, Move lfm, Move rfm, Move lhm, Move rhm) {
            bb.begin_segment(foot_state ? foot_ticks : com_ticks, mask, mask_last);
            const double seg_dur = foot_state ? foot_dur : com_dur;
            const Move mv[5] = {com, lfm, rfm, lhm, rhm};
            for (int ch = 0; ch < 5; ++ch) {
                // hands always use traj_com_duration, also inside foot states (walk.cpp:101-102,119-120,...)
                const double d = (ch == LH || ch == RH) ? com_dur : seg_dur;
                bb.set(ch, mv[ch].from->v, mv[ch].to ? mv[ch].to->v : mv[ch].from->v, d, mv[ch].to != nullptr);
            }
        };
        auto hold = [](const P3& p) { return Move{&p, nullptr}; };
        auto go = [](const P3& a, const P3& b2) { return Move{&a, &b2}; };

        // INIT: CoM from the middle over the right foot; LF_INIT: left foot up (its contact is removed at the first tick)
        emit(false, DS, DS, go(com_mid, com_rf), hold(lf_low), hold(rf_low), hold(lh), hold(rh));
        emit(true, DS & ~LF, DS & ~LF, hold(com_rf), go(lf_low, lf_high), hold(rf_low), hold(lh), hold(rh));
        for (int cyc = 0; cyc <= ncyc; ++cyc) {
            const bool last = cyc == ncyc; // LIFT_DOWN_LF_FINAL + MOVE_COM_CENTER_FINAL instead of a full cycle
            // left foot lands ahead of the right one (the first time it also catches up: float arithmetic as in walk.cpp:116-118)
            const float diff = (float)(rf_low.v[0] - lf_low.v[0]);
            const double adv = last ? (double)step_length : (double)(float)(diff + step_length);
            const P3 lf_from = lf_high, lh_from = lh;
            lf_low.v[0] += adv;
            lh_fwd = lh;
            lh_fwd.v[0] += adv;
            emit(true, DS & ~LF, DS, hold(com_rf), go(lf_from, lf_low), hold(rf_low), go(lh_from, lh_fwd), hold(rh));
            if (last) {
                com_mid.v[0] = (rf_low.v[0] + lf_low.v[0]) / 2.0;
                com_mid.v[1] = (rf_low.v[1] + lf_low.v[1]) / 2.0;
                emit(false, DS, DS, go(com_rf, com_mid), hold(lf_low), hold(rf_low), hold(lh_fwd), hold(rh_fwd));
                break;
            }
            lh = lh_fwd;
            // CoM over the left foot, right foot up, forward and down, CoM over the right foot, left foot up
            const P3 com_from = com_rf;
            com_lf.v[0] = lf_low.v[0];
            emit(false, DS, DS, go(com_from, com_lf), hold(lf_low), hold(rf_low), hold(lh), hold(rh));
            rf_high.v[0] = lf_low.v[0];
            emit(true, DS & ~RF, DS & ~RF, hold(com_lf), hold(lf_low), go(rf_low, rf_high), hold(lh), hold(rh));
            const P3 rf_from = rf_high, rh_from = rh;
            const double adv2 = (double)(2 * step_length);
            rf_low.v[0] += adv2;
            rh_fwd = rh;
            rh_fwd.v[0] += adv2;
            emit(true, DS & ~RF, DS, hold(com_lf), hold(lf_low), go(rf_from, rf_low), hold(lh), go(rh_from, rh_fwd));
            rh = rh_fwd;
            com_rf.v[0] = rf_low.v[0];
            emit(false, DS, DS, go(com_lf, com_rf), hold(lf_low), hold(rf_low), hold(lh), hold(rh));
            lf_high.v[0] = rf_low.v[0];
            emit(true, DS & ~LF, DS & ~LF, hold(com_rf), go(lf_low, lf_high), hold(rf_low), hold(lh), hold(rh));
        }
        bb.finish();
    }

    // humanoid::walk-on-spot (src/behaviors/humanoid/walk_on_spot.cpp:36-208): CoM over the right foot once (INIT), then for ever
    // { left foot up, down, CoM over the left foot, right foot up, down, CoM over the right foot }; a foot's contact is removed at
    // the first tick of its LIFT_UP state and re-added at the last tick of its LIFT_DOWN state (walk_on_spot.cpp:163-182).
    inline void compile_walk_on_spot(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c)
    {
        BehaviorBuilder bb(out, prob);
        const double com_dur = IWBC_HCHECK(c["traj_com_duration"].as<float>(), "traj_com_duration");
        const double foot_dur = IWBC_HCHECK(c["traj_foot_duration"].as<float>(), "traj_foot_duration");
        const float step_height = IWBC_HCHECK(c["step_height"].as<float>(), "step_height");
        const int lfb = iwbc_problem_contact_index(&prob, "contact_lfoot"), rfb = iwbc_problem_contact_index(&prob, "contact_rfoot");
        if (lfb < 0 || rfb < 0)
            throw ConfigError("Walk on spot requires a humanoid controller!"); // walk_on_spot.cpp:46
        const unsigned LF = 1u << lfb, RF = 1u << rfb, DS = iwbc_problem_default_contact_mask(&prob) | LF | RF;
        enum { COM, LFT, RFT };
        bb.add_channel(0, 0, bb.slot_of("com"), -1);
        bb.add_channel(1, 0, bb.slot_of("lf"), bb.slot_of("contact_lfoot"));
        bb.add_channel(1, 0, bb.slot_of("rf"), bb.slot_of("contact_rfoot"));
        double lf_low[3], rf_low[3], lf_high[3], rf_high[3], com_init[3], com_lf[3], com_rf[3];
        for (int k = 0; k < 3; ++k) {
            lf_low[k] = lf_high[k] = com_lf[k] = prob.cs.ref0[out.channels[LFT].slot[0] + k];
            rf_low[k] = rf_high[k] = com_rf[k] = prob.cs.ref0[out.channels[RFT].slot[0] + k];
            com_init[k] = prob.cs.ref0[out.channels[COM].slot[0] + k];
        }
        lf_high[2] += step_height, rf_high[2] += step_height;
        com_lf[2] = com_rf[2] = com_init[2];
        const int com_ticks = beh_n_steps(out.dt, com_dur), foot_ticks = beh_n_steps(out.dt, foot_dur);
        auto seg = [&](bool foot_state, unsigned mask, unsigned mask_last, const double* c0, const double* c1, const double* l0, const double* l1,
                       const double* r0, const double* r1) {
            const double d = foot_state ? foot_dur : com_dur;
            bb.begin_segment(foot_state ? foot_ticks : com_ticks, mask, mask_last);
            bb.set(COM, c0, c1 ? c1 : c0, d, c1 != nullptr);
            bb.set(LFT, l0, l1 ? l1 : l0, d, l1 != nullptr);
            bb.set(RFT, r0, r1 ? r1 : r0, d, r1 != nullptr);
        };
        seg(false, DS, DS, com_init, com_rf, lf_low, nullptr, rf_low, nullptr);          // INIT (once)
        seg(true, DS & ~LF, DS & ~LF, com_rf, nullptr, lf_low, lf_high, rf_low, nullptr); // LIFT_UP_LF
        seg(true, DS & ~LF, DS, com_rf, nullptr, lf_high, lf_low, rf_low, nullptr);       // LIFT_DOWN_LF
        seg(false, DS, DS, com_rf, com_lf, lf_low, nullptr, rf_low, nullptr);             // MOVE_COM_LEFT
        seg(true, DS & ~RF, DS & ~RF, com_lf, nullptr, lf_low, nullptr, rf_low, rf_high); // LIFT_UP_RF
        seg(true, DS & ~RF, DS, com_lf, nullptr, lf_low, nullptr, rf_high, rf_low);       // LIFT_DOWN_RF
        seg(false, DS, DS, com_lf, com_rf, lf_low, nullptr, rf_low, nullptr);             // MOVE_COM_RIGHT
        out.loop = true;
        out.loop_start = com_ticks; // "we skip the init_traj" (walk_on_spot.cpp:202-204)
        bb.finish();
    }

    // humanoid::clapping (src/behaviors/humanoid/clapping.cpp:8-52): both hands move `motion_size` towards each other along y and back, for ever
    inline void compile_clapping(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c)
    {
        BehaviorBuilder bb(out, prob);
        const double duration = IWBC_HCHECK(c["trajectory_duration"].as<float>(), "trajectory_duration");
        const float motion_size = IWBC_HCHECK(c["motion_size"].as<float>(), "motion_size");
        bb.add_channel(1, 0, bb.slot_of("lh"), -1);
        bb.add_channel(1, 0, bb.slot_of("rh"), -1);
        double lh0[3], rh0[3], lh1[3], rh1[3];
        for (int k = 0; k < 3; ++k) {
            lh0[k] = lh1[k] = prob.cs.ref0[out.channels[0].slot[0] + k];
            rh0[k] = rh1[k] = prob.cs.ref0[out.channels[1].slot[0] + k];
        }
        lh1[1] -= motion_size;
        rh1[1] += motion_size;
        const unsigned mask = iwbc_problem_default_contact_mask(&prob);
        const int ticks = beh_n_steps(out.dt, duration);
        bb.begin_segment(ticks, mask, mask);
        bb.set(0, lh0, lh1, duration, true);
        bb.set(1, rh0, rh1, duration, true);
        bb.begin_segment(ticks, mask, mask);
        bb.set(0, lh1, lh0, duration, true);
        bb.set(1, rh1, rh0, duration, true);
        out.loop = true;
        bb.finish();
    }

    namespace detail {
        struct Rot3 {
            double m[9]; // row-major
        };
        inline Rot3 rmul(const Rot3& a, const Rot3& b)
        {
            Rot3 c;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
            return c;
        }
        inline Rot3 rtrans(const Rot3& a)
        {
            Rot3 t;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    t.m[3 * i + j] = a.m[3 * j + i];
            return t;
        }
        inline Rot3 raxis(int axis, double th)
        {
            const double c = std::cos(th), s2 = std::sin(th);
            Rot3 r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
            const int a = (axis + 1) % 3, b = (axis + 2) % 3;
            r.m[3 * a + a] = c, r.m[3 * a + b] = -s2, r.m[3 * b + a] = s2, r.m[3 * b + b] = c;
            return r;
        }
        // Eigen::AngleAxisd(R): through the quaternion (Eigen's Quaternion(Matrix3) then AngleAxis(Quaternion)), angle in [0, pi]
        inline void angle_axis(const Rot3& R, double& angle, double axis[3])
        {
            double w, q[3];
            const double tr = R.m[0] + R.m[4] + R.m[8];
            if (tr > 0) {
                double t = std::sqrt(tr + 1.0);
                w = 0.5 * t;
                t = 0.5 / t;
                q[0] = (R.m[7] - R.m[5]) * t, q[1] = (R.m[2] - R.m[6]) * t, q[2] = (R.m[3] - R.m[1]) * t;
            }
            else {
                int i = 0;
                if (R.m[4] > R.m[0])
                    i = 1;
                if (R.m[8] > R.m[4 * i])
                    i = 2;
                const int j = (i + 1) % 3, k = (i + 2) % 3;
                double t = std::sqrt(R.m[4 * i] - R.m[4 * j] - R.m[4 * k] + 1.0);
                q[i] = 0.5 * t;
                t = 0.5 / t;
                w = (R.m[3 * k + j] - R.m[3 * j + k]) * t;
                q[j] = (R.m[3 * j + i] + R.m[3 * i + j]) * t;
                q[k] = (R.m[3 * k + i] + R.m[3 * i + k]) * t;
            }
            double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            if (n == 0.0) {
                angle = 0.0;
                axis[0] = 1.0, axis[1] = 0.0, axis[2] = 0.0;
                return;
            }
            angle = 2.0 * std::atan2(n, std::fabs(w));
            if (w < 0)
                n = -n;
            for (int c = 0; c < 3; ++c)
                axis[c] = q[c] / n;
        }
        inline Rot3 slot_rotation(const double* ref24)
        { // TrajectorySample(12,6): R column-major after the translation
            Rot3 R;
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r)
                    R.m[3 * r + c] = ref24[3 + 3 * c + r];
            return R;
        }
    } // namespace detail

    // generic::cartesian (src/behaviors/generic/cartesian.cpp:8-94): every listed SE3 task goes from its construction-time
    // reference to a relative target along a minimum-jerk path (position and rotation angle about the fixed relative axis),
    // with the polynomial's derivatives as feed-forward velocity / acceleration; `loop` appends the way back.
    // humanoid::move-feet (src/behaviors/humanoid/move_feet.cpp:8-102) is the same generator with each sample also written into the
    // contact that shares the task's frame (`contact_names`, move_feet.cpp:87): with_contacts = true.
    inline void compile_cartesian(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c, bool with_contacts = false)
    {
        BehaviorBuilder bb(out, prob);
        const double duration = IWBC_HCHECK(c["trajectory_duration"].as<float>(), "trajectory_duration");
        out.loop = IWBC_HCHECK(c["loop"].as<bool>(), "loop");
        const std::vector<std::string> names = IWBC_HCHECK(c["task_names"].as<std::vector<std::string>>(), "task_names");
        const yaml::Node& ts = c["relative_targets_pos"];
        const yaml::Node& to = c["relative_targets_rpy"];
        if (!ts.is_seq() || !to.is_seq())
            throw ConfigError("config error reading relative_targets_pos / relative_targets_rpy");
        if (names.size() != ts.size())
            throw ConfigError(with_contacts ? "MoveFeet behavior needs the same number of tasks and targets" : "cartesian behavior needs the same number of tasks and targets"); // cartesian.cpp:22-23, move_feet.cpp:31-32
        if (to.size() < names.size())
            throw ConfigError("cartesian behavior needs one relative_targets_rpy entry per task");
        struct Way {
            double p0[3], p1[3];
            detail::Rot3 R0, R1;
        };
        std::vector<Way> ways;
        std::vector<std::string> contacts;
        if (with_contacts) {
            contacts = IWBC_HCHECK(c["contact_names"].as<std::vector<std::string>>(), "contact_names");
            if (contacts.size() < names.size())
                throw ConfigError("MoveFeet behavior needs one contact per task");
        }
        for (size_t i = 0; i < names.size(); ++i) {
            const int slot = bb.slot_of(names[i]);
            bb.add_channel(1, 1, slot, with_contacts ? bb.slot_of(contacts[i]) : -1);
            Way w;
            const double* ref = prob.cs.ref0.data() + slot;
            for (int k = 0; k < 3; ++k)
                w.p0[k] = w.p1[k] = ref[k];
            w.R0 = w.R1 = detail::slot_rotation(ref);
            const std::vector<double> dp = ts[i].as<std::vector<double>>(), rpy = to[i].as<std::vector<double>>();
            if (dp.size() == 3)
                for (int k = 0; k < 3; ++k)
                    w.p1[k] = dp[k] + w.p0[k];
            if (rpy.size() == 3) // AngleAxis(rz, Z) * AngleAxis(ry, Y) * AngleAxis(rx, X) * R_init (cartesian.cpp:35-40)
                w.R1 = detail::rmul(detail::rmul(detail::rmul(detail::raxis(2, rpy[2]), detail::raxis(1, rpy[1])), detail::raxis(0, rpy[0])), w.R0);
            ways.push_back(w);
        }
        const unsigned mask = iwbc_problem_default_contact_mask(&prob);
        const int ticks = beh_n_steps(out.dt, duration);
        for (int leg = 0; leg < (out.loop ? 2 : 1); ++leg) {
            bb.begin_segment(ticks, mask, mask);
            for (size_t i = 0; i < ways.size(); ++i) {
                const Way& w = ways[i];
                iwbc_beh_chanseg_t& cs = bb.at((int)i);
                std::memset(&cs, 0, sizeof(cs));
                const detail::Rot3 &Ra = leg == 0 ? w.R0 : w.R1, &Rb = leg == 0 ? w.R1 : w.R0;
                for (int k = 0; k < 3; ++k) {
                    cs.p0[k] = leg == 0 ? w.p0[k] : w.p1[k];
                    cs.p1[k] = leg == 0 ? w.p1[k] : w.p0[k];
                }
                for (int col = 0; col < 3; ++col)
                    for (int r = 0; r < 3; ++r)
                        cs.R0[3 * col + r] = Ra.m[3 * r + col];
                detail::angle_axis(detail::rmul(detail::rtrans(Ra), Rb), cs.angle, cs.axis); // trajectory_generator.hpp:124-126
                cs.duration = duration;
                cs.minjerk = 1;
                cs.rot = cs.angle != 0.0 ? 1 : 0;
            }
        }
        bb.finish();
    }

    // humanoid::move_com (src/behaviors/humanoid/move_com.cpp:8-61): the CoM reference visits the listed targets (masked,
    // relative or absolute) along minimum-jerk segments with feed-forward derivatives; `loop` appends the way back to the start.
    inline void compile_move_com(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c)
    {
        BehaviorBuilder bb(out, prob);
        const double duration = IWBC_HCHECK(c["trajectory_duration"].as<float>(), "trajectory_duration");
        out.loop = IWBC_HCHECK(c["loop"].as<bool>(), "loop");
        const yaml::Node& tg = c["targets"];
        const std::string mask = IWBC_HCHECK(c["mask"].as<std::string>(), "mask");
        const bool absolute = IWBC_HCHECK(c["absolute"].as<bool>(), "absolute");
        if (!tg.is_seq())
            throw ConfigError("config error reading targets");
        if (mask.size() != 3)
            throw ConfigError("The mask for the CoM should be 3-dimensional"); // move_com.cpp:20
        const int slot = bb.slot_of("com");
        bb.add_channel(0, 1, slot, -1);
        double init[3];
        for (int k = 0; k < 3; ++k)
            init[k] = prob.cs.ref0[slot + k];
        std::vector<std::vector<double>> targets;
        for (size_t i = 0; i < tg.size(); ++i)
            targets.push_back(tg[i].as<std::vector<double>>());
        if (out.loop)
            targets.push_back(absolute ? std::vector<double>{init[0], init[1], init[2]} : std::vector<double>{0.0, 0.0, 0.0});
        const unsigned cm = iwbc_problem_default_contact_mask(&prob);
        const int ticks = beh_n_steps(out.dt, duration);
        double start[3] = {init[0], init[1], init[2]};
        for (const std::vector<double>& t : targets) {
            if (t.size() != 3)
                throw ConfigError("references need to be 3-dimensional"); // move_com.cpp:32
            double end[3] = {init[0], init[1], init[2]};
            for (int j = 0; j < 3; ++j)
                if (mask[j] == '1')
                    end[j] = absolute ? t[j] : t[j] + init[j];
            bb.begin_segment(ticks, cm, cm);
            bb.set(0, start, end, duration, true);
            for (int k = 0; k < 3; ++k)
                start[k] = end[k];
        }
        bb.finish();
    }

    // generic::cartesian_traj (src/behaviors/generic/cartesian_traj.cpp:8-46, src/trajs/loader.cpp:11-100): recorded SE3 / CoM
    // trajectories (one CSV per task: 3 + 9 numbers per row, the rotation in Eigen's column-major order; 3 per row for "com") replayed
    // row by row, forwards then backwards for ever (`step_` flips at both ends; the `loop` key is read but not used by the reference).
    inline void compile_cartesian_traj(iwbc_behavior& out, const iwbc_problem& prob, const yaml::Node& c, const std::string& dir)
    {
        BehaviorBuilder bb(out, prob);
        (void)IWBC_HCHECK(c["loop"].as<bool>(), "loop");
        const std::string rel = IWBC_HCHECK(c["trajectories"].as<std::string>(), "trajectories");
        const double scale = IWBC_HCHECK(c["scale"].as<float>(), "scale");
        const std::string ypath = dir + "/" + rel;
        const yaml::Node root = yaml::LoadFile(ypath);
        const yaml::Node& refs = root["refs"];
        if (!refs.is_map())
            throw ConfigError("trajectory file without a 'refs' map: " + ypath);
        const std::string tdir = ypath.substr(0, ypath.find_last_of('/'));
        int rows = -1;
        for (auto& kv : refs.map) {
            if (kv.first == "posture")
                throw ConfigError("cartesian_traj: posture trajectories have no device-side form yet");
            const bool com = kv.first == "com";
            const int width = com ? 3 : 12;
            const std::string file = tdir + "/" + kv.second.as<std::string>();
            std::ifstream ifs(file.c_str());
            if (!ifs.good())
                throw ConfigError("Error when loading trajectory:" + file); // loader.cpp:15
            std::vector<double> data;
            double x;
            while (ifs >> x)
                data.push_back(x);
            if (data.empty() || data.size() % width != 0)
                throw ConfigError("wrong number of values in " + file);
            const int n = (int)(data.size() / width);
            if (rows >= 0 && n != rows)
                throw ConfigError("wrong number of rows in" + file); // loader.cpp:76,88
            rows = n;
            const int ch = bb.add_channel(com ? 0 : 1, 0, bb.slot_of(kv.first), -1);
            out.channels[ch].table_off = (int)out.table.size();
            out.channels[ch].table_rows = n;
            out.channels[ch].scale = scale;
            out.table.insert(out.table.end(), data.begin(), data.end());
        }
        if (rows < 2)
            throw ConfigError("cartesian_traj needs at least two recorded rows");
        const unsigned mask = iwbc_problem_default_contact_mask(&prob);
        bb.begin_segment(2 * rows - 2, mask, mask); // 0, 1, ..., rows-1, rows-2, ..., 1 and again
        for (size_t ch = 0; ch < out.channels.size(); ++ch)
            std::memset(&bb.at((int)ch), 0, sizeof(iwbc_beh_chanseg_t));
        out.loop = true;
        bb.finish();
    }

    inline void compile_behavior(iwbc_behavior& out, const iwbc_problem& prob, const std::string& yaml_path, const std::string& overrides)
    {
        yaml::Node root = yaml::LoadFile(yaml_path);
        if (!root["BEHAVIOR"])
            throw ConfigError("behavior file without a BEHAVIOR node: " + yaml_path);
        yaml::Node c = root["BEHAVIOR"];
        if (!overrides.empty()) {
            yaml::Node ov = yaml::Load(overrides);
            for (auto& kv : ov.map)
                c.at(kv.first) = kv.second;
        }
        out.name = IWBC_HCHECK(c["name"].as<std::string>(), "BEHAVIOR.name");
        // Behavior::_customize_tasks (src/behaviors/behavior.cpp:7-22): constructing a behaviour re-weights the controller's tasks
        if (c["customize_task_weights"].is_map())
            for (auto& kv : c["customize_task_weights"].map) {
                const double w = IWBC_HCHECK(kv.second.as<double>(), "customize_task_weights." + kv.first);
                if (iwbc_problem_set_task_weight(const_cast<iwbc_problem*>(&prob), kv.first.c_str(), w) != IWBC_OK)
                    throw ConfigError("customize_task_weights: unknown task '" + kv.first + "'");
            }
        if (out.name == "humanoid::walk")
            compile_walk(out, prob, c);
        else if (out.name == "humanoid::walk-on-spot")
            compile_walk_on_spot(out, prob, c);
        else if (out.name == "humanoid::clapping")
            compile_clapping(out, prob, c);
        else if (out.name == "generic::cartesian")
            compile_cartesian(out, prob, c);
        else if (out.name == "generic::cartesian_traj")
            compile_cartesian_traj(out, prob, c, yaml_path.find_last_of('/') == std::string::npos ? std::string(".") : yaml_path.substr(0, yaml_path.find_last_of('/')));
        else if (out.name == "humanoid::move-feet")
            compile_cartesian(out, prob, c, true);
        else if (out.name == "humanoid::move_com")
            compile_move_com(out, prob, c);
        else
            throw ConfigError("behavior '" + out.name + "' has no device-side form yet (available: humanoid::walk, humanoid::walk-on-spot, humanoid::clapping, humanoid::move_com, humanoid::move-feet, generic::cartesian, generic::cartesian_traj)");
    }

} // namespace host
} // namespace iwbc
