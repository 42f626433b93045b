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
    std::vector<iwbc_beh_channel_t> channels;
    std::vector<iwbc_beh_seg_t> segs;
    std::vector<iwbc_beh_chanseg_t> cs; // [seg][channel]
    std::vector<unsigned> masks;        // distinct contact masks, most frequent first
    // device copies (one device; created on first use)
    int device = -1;
    void* d_blob = nullptr;
    iwbc_beh_channel_t* d_channels = nullptr;
    iwbc_beh_seg_t* d_segs = nullptr;
    iwbc_beh_chanseg_t* d_cs = nullptr;
    double* d_ref0 = nullptr;
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
            b.channels.push_back({kind, ff, {slot0, slot1}});
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
        if (out.name == "humanoid::walk")
            compile_walk(out, prob, c);
        else
            throw ConfigError("behavior '" + out.name + "' has no device-side form yet (available: humanoid::walk)");
    }

} // namespace host
} // namespace iwbc
