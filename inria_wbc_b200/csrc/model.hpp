// model.hpp -- host-side robot model: URDF -> joint tree with pinocchio's conventions.
//
// Stands in for `pinocchio::urdf::buildModel` (+ `JointModelFreeFlyer` root) and
// `pinocchio::srdf::loadReferenceConfigurations` as the reference calls them
// (/root/reference/src/controllers/controller.cpp:101-118, src/controllers/pos_tracker.cpp:57-77,
// virtual frames: src/controllers/pos_tracker.cpp:191-209).  Conventions kept:
//   * joints are numbered depth-first, children visited in alphabetical joint-name order (urdfdom
//     stores joints in a std::map and pinocchio walks child_links in that order);
//   * fixed joints are folded into the supporting movable joint: inertias merged, a FIXED_JOINT frame
//     and a BODY frame are added; every movable joint gets a JOINT frame of the same name;
//   * free-flyer: q = [p(3), quat xyzw(4)], v = [v_lin; w] in the body frame; `root_joint`;
//   * <mimic> tags are ignored (pinocchio 2.x), so mimic joints are independent dofs.
// Construction-time only (not the hot path): the first forward kinematics that fixes the initial
// task references runs here on the host, once.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "xml_lite.hpp"

namespace iwbc {
namespace host {

    struct SE3 {
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; // row-major
        double p[3] = {0, 0, 0};
    };

    inline SE3 mul(const SE3& a, const SE3& b)
    {
        SE3 c;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k)
                    s += a.R[3 * i + k] * b.R[3 * k + j];
                c.R[3 * i + j] = s;
            }
            c.p[i] = a.p[i] + a.R[3 * i] * b.p[0] + a.R[3 * i + 1] * b.p[1] + a.R[3 * i + 2] * b.p[2];
        }
        return c;
    }
    inline void rot_apply(const double* R, const double* v, double* out)
    {
        for (int i = 0; i < 3; ++i)
            out[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
    }
    inline SE3 from_rpy_xyz(const double rpy[3], const double xyz[3])
    {
        // URDF: R = Rz(yaw) Ry(pitch) Rx(roll)
        double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]);
        double cp = std::cos(rpy[1]), sp = std::sin(rpy[1]);
        double cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
        SE3 m;
        m.R[0] = cy * cp;
        m.R[1] = cy * sp * sr - sy * cr;
        m.R[2] = cy * sp * cr + sy * sr;
        m.R[3] = sy * cp;
        m.R[4] = sy * sp * sr + cy * cr;
        m.R[5] = sy * sp * cr - cy * sr;
        m.R[6] = -sp;
        m.R[7] = cp * sr;
        m.R[8] = cp * cr;
        for (int i = 0; i < 3; ++i)
            m.p[i] = xyz[i];
        return m;
    }
    inline void axis_angle_R(const double a[3], double th, double* R)
    {
        double c = std::cos(th), s = std::sin(th), v = 1 - c;
        R[0] = a[0] * a[0] * v + c;
        R[1] = a[0] * a[1] * v - a[2] * s;
        R[2] = a[0] * a[2] * v + a[1] * s;
        R[3] = a[0] * a[1] * v + a[2] * s;
        R[4] = a[1] * a[1] * v + c;
        R[5] = a[1] * a[2] * v - a[0] * s;
        R[6] = a[0] * a[2] * v - a[1] * s;
        R[7] = a[1] * a[2] * v + a[0] * s;
        R[8] = a[2] * a[2] * v + c;
    }
    inline void quat_R(double x, double y, double z, double w, double* R)
    {
        R[0] = 1 - 2 * (y * y + z * z);
        R[1] = 2 * (x * y - z * w);
        R[2] = 2 * (x * z + y * w);
        R[3] = 2 * (x * y + z * w);
        R[4] = 1 - 2 * (x * x + z * z);
        R[5] = 2 * (y * z - x * w);
        R[6] = 2 * (x * z - y * w);
        R[7] = 2 * (y * z + x * w);
        R[8] = 1 - 2 * (x * x + y * y);
    }

    // rigid-body inertia: mass, CoM and 3x3 rotational inertia about the CoM, all in one frame
    struct Inertia {
        double m = 0;
        double c[3] = {0, 0, 0};
        double I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    };
    inline Inertia transform(const SE3& X, const Inertia& a)
    { // express `a` (given in frame B) in frame A, X = A_M_B
        Inertia o;
        o.m = a.m;
        rot_apply(X.R, a.c, o.c);
        for (int i = 0; i < 3; ++i)
            o.c[i] += X.p[i];
        double t[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k)
                    s += X.R[3 * i + k] * a.I[3 * k + j];
                t[3 * i + j] = s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k)
                    s += t[3 * i + k] * X.R[3 * j + k];
                o.I[3 * i + j] = s;
            }
        return o;
    }
    inline Inertia add(const Inertia& a, const Inertia& b)
    {
        if (a.m == 0)
            return b;
        if (b.m == 0)
            return a;
        Inertia o;
        o.m = a.m + b.m;
        for (int i = 0; i < 3; ++i)
            o.c[i] = (a.m * a.c[i] + b.m * b.c[i]) / o.m;
        auto shift = [&](const Inertia& x) {
            double d[3] = {x.c[0] - o.c[0], x.c[1] - o.c[1], x.c[2] - o.c[2]};
            double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    o.I[3 * i + j] += x.I[3 * i + j] + x.m * ((i == j ? dd : 0.0) - d[i] * d[j]);
        };
        shift(a);
        shift(b);
        return o;
    }

    enum JointType { FREEFLYER = 0, REVOLUTE = 1, PRISMATIC = 2 };
    enum FrameType { F_JOINT = 0, F_FIXED_JOINT = 1, F_BODY = 2, F_OP = 3 };

    struct Joint {
        std::string name;
        int type = REVOLUTE;
        int parent = -1; // -1 = universe
        SE3 placement;
        double axis[3] = {0, 0, 1};
        int idx_q = 0, idx_v = 0;
        Inertia Y;
        double q_lo = 0, q_hi = 0, v_max = 0, tau_max = 0;
    };
    struct Frame {
        std::string name;
        int parent = -1; // supporting joint, -1 = universe
        SE3 placement;
        int type = F_BODY;
    };

    struct Model {
        std::string name;
        std::vector<Joint> joints;
        std::vector<Frame> frames;
        int nq = 0, nv = 0;
        bool floating_base = false;
        std::map<std::string, std::vector<double>> reference_configurations;

        int na() const { return floating_base ? nv - 6 : nv; }
        int joint_id(const std::string& n) const
        {
            for (size_t i = 0; i < joints.size(); ++i)
                if (joints[i].name == n)
                    return (int)i;
            return -1;
        }
        bool exist_joint(const std::string& n) const { return joint_id(n) >= 0; }
        int frame_id(const std::string& n) const
        {
            for (size_t i = 0; i < frames.size(); ++i)
                if (frames[i].name == n)
                    return (int)i;
            return -1;
        }
        bool exist_frame(const std::string& n) const { return frame_id(n) >= 0; }
        bool exist_body(const std::string& n) const
        {
            for (auto& f : frames)
                if (f.name == n && f.type == F_BODY)
                    return true;
            return false;
        }
        // names of the actuated dofs (tail na of v), pinocchio order
        std::vector<std::string> actuated_names() const
        {
            std::vector<std::string> out;
            for (auto& j : joints)
                if (j.type != FREEFLYER)
                    out.push_back(j.name);
            return out;
        }
        // PosTracker::parse_frames: a virtual frame rigidly attached to an existing frame
        void add_virtual_frame(const std::string& name, const std::string& ref, const double pos[3])
        {
            int id = frame_id(ref);
            if (id < 0)
                throw std::runtime_error("virtual frame '" + name + "': unknown reference frame '" + ref + "'");
            Frame f;
            f.name = name;
            f.parent = frames[id].parent;
            SE3 t;
            for (int i = 0; i < 3; ++i)
                t.p[i] = pos[i];
            f.placement = mul(frames[id].placement, t);
            f.type = F_FIXED_JOINT;
            frames.push_back(f);
        }
    };

    namespace detail {
        inline std::vector<double> numbers(const std::string& s)
        {
            std::vector<double> v;
            std::istringstream in(s);
            double d;
            while (in >> d)
                v.push_back(d);
            return v;
        }
        struct ULink {
            std::string name;
            Inertia Y;
        };
        struct UJoint {
            std::string name, type, parent, child;
            SE3 origin;
            double axis[3] = {1, 0, 0};
            double lo = 0, hi = 0, vel = 0, eff = 0;
            bool has_limit = false;
        };
    } // namespace detail

    inline Model load_urdf(const std::string& path, bool floating_base)
    {
        using namespace detail;
        xml::Element robot = xml::LoadFile(path);
        if (robot.name != "robot")
            throw std::runtime_error("urdf: root element is <" + robot.name + ">, expected <robot> in " + path);
        std::map<std::string, ULink> links;
        std::map<std::string, UJoint> joints; // sorted by name, like urdfdom
        for (auto& e : robot.children) {
            if (e.name == "link") {
                ULink l;
                l.name = e.attr_or("name", "");
                if (auto* in = e.child("inertial")) {
                    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
                    if (auto* o = in->child("origin")) {
                        auto a = numbers(o->attr_or("xyz", "0 0 0")), b = numbers(o->attr_or("rpy", "0 0 0"));
                        for (int i = 0; i < 3 && i < (int)a.size(); ++i)
                            xyz[i] = a[i];
                        for (int i = 0; i < 3 && i < (int)b.size(); ++i)
                            rpy[i] = b[i];
                    }
                    Inertia y;
                    if (auto* m = in->child("mass"))
                        y.m = std::atof(m->attr_or("value", "0").c_str());
                    if (auto* it = in->child("inertia")) {
                        auto g = [&](const char* k) { return std::atof(it->attr_or(k, "0").c_str()); };
                        double xx = g("ixx"), xy = g("ixy"), xz = g("ixz"), yy = g("iyy"), yz = g("iyz"), zz = g("izz");
                        double I[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
                        for (int i = 0; i < 9; ++i)
                            y.I[i] = I[i];
                    }
                    SE3 X = from_rpy_xyz(rpy, xyz);
                    l.Y = transform(X, y); // inertia in the link frame (CoM at xyz, axes rotated by rpy)
                }
                links[l.name] = l;
            }
            else if (e.name == "joint") {
                UJoint j;
                j.name = e.attr_or("name", "");
                j.type = e.attr_or("type", "fixed");
                if (auto* p = e.child("parent"))
                    j.parent = p->attr_or("link", "");
                if (auto* c = e.child("child"))
                    j.child = c->attr_or("link", "");
                double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
                if (auto* o = e.child("origin")) {
                    auto a = numbers(o->attr_or("xyz", "0 0 0")), b = numbers(o->attr_or("rpy", "0 0 0"));
                    for (int i = 0; i < 3 && i < (int)a.size(); ++i)
                        xyz[i] = a[i];
                    for (int i = 0; i < 3 && i < (int)b.size(); ++i)
                        rpy[i] = b[i];
                }
                j.origin = from_rpy_xyz(rpy, xyz);
                if (auto* ax = e.child("axis")) {
                    auto a = numbers(ax->attr_or("xyz", "1 0 0"));
                    double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
                    for (int i = 0; i < 3; ++i)
                        j.axis[i] = a[i] / n;
                }
                if (auto* lim = e.child("limit")) {
                    j.has_limit = true;
                    j.lo = std::atof(lim->attr_or("lower", "0").c_str());
                    j.hi = std::atof(lim->attr_or("upper", "0").c_str());
                    j.vel = std::atof(lim->attr_or("velocity", "0").c_str());
                    j.eff = std::atof(lim->attr_or("effort", "0").c_str());
                }
                joints[j.name] = j;
            }
        }
        if (links.empty())
            throw std::runtime_error("urdf: no <link> in " + path);
        // root link = the one that is nobody's child
        std::string root;
        for (auto& l : links) {
            bool is_child = false;
            for (auto& j : joints)
                if (j.second.child == l.first)
                    is_child = true;
            if (!is_child) {
                if (!root.empty())
                    throw std::runtime_error("urdf: several root links (" + root + ", " + l.first + ")");
                root = l.first;
            }
        }
        Model model;
        model.name = robot.attr_or("name", "robot");
        model.floating_base = floating_base;
        Frame uni;
        uni.name = "universe";
        uni.parent = -1;
        uni.type = F_FIXED_JOINT;
        model.frames.push_back(uni);

        int root_joint = -1;
        if (floating_base) {
            Joint j;
            j.name = "root_joint";
            j.type = FREEFLYER;
            j.parent = -1;
            j.idx_q = 0;
            j.idx_v = 0;
            model.nq = 7;
            model.nv = 6;
            model.joints.push_back(j);
            root_joint = 0;
            Frame f;
            f.name = "root_joint";
            f.parent = 0;
            f.type = F_JOINT;
            model.frames.push_back(f);
        }
        // recursive descent
        struct Ctx {
            Model& model;
            std::map<std::string, ULink>& links;
            std::map<std::string, UJoint>& joints;
            void visit(const std::string& link, int sup_joint, const SE3& link_in_joint)
            {
                // body frame + inertia of this link
                Frame bf;
                bf.name = link;
                bf.parent = sup_joint;
                bf.placement = link_in_joint;
                bf.type = F_BODY;
                model.frames.push_back(bf);
                if (sup_joint >= 0)
                    model.joints[sup_joint].Y = add(model.joints[sup_joint].Y, transform(link_in_joint, links[link].Y));
                for (auto& kv : joints) {
                    const UJoint& uj = kv.second;
                    if (uj.parent != link)
                        continue;
                    if (!links.count(uj.child))
                        throw std::runtime_error("urdf: joint '" + uj.name + "' has unknown child link '" + uj.child + "'");
                    SE3 jplace = mul(link_in_joint, uj.origin);
                    if (uj.type == "fixed") {
                        Frame jf;
                        jf.name = uj.name;
                        jf.parent = sup_joint;
                        jf.placement = jplace;
                        jf.type = F_FIXED_JOINT;
                        model.frames.push_back(jf);
                        visit(uj.child, sup_joint, jplace);
                    }
                    else if (uj.type == "revolute" || uj.type == "prismatic" || uj.type == "continuous") {
                        if (uj.type == "continuous")
                            throw std::runtime_error("urdf: continuous joint '" + uj.name + "' is not supported (pinocchio maps it to a cos/sin pair); use a no-wheel model");
                        Joint j;
                        j.name = uj.name;
                        j.type = uj.type == "prismatic" ? PRISMATIC : REVOLUTE;
                        j.parent = sup_joint;
                        j.placement = jplace;
                        for (int i = 0; i < 3; ++i)
                            j.axis[i] = uj.axis[i];
                        j.idx_q = model.nq++;
                        j.idx_v = model.nv++;
                        j.q_lo = uj.lo;
                        j.q_hi = uj.hi;
                        j.v_max = uj.vel;
                        j.tau_max = uj.eff;
                        model.joints.push_back(j);
                        int id = (int)model.joints.size() - 1;
                        Frame jf;
                        jf.name = uj.name;
                        jf.parent = id;
                        jf.type = F_JOINT;
                        model.frames.push_back(jf);
                        visit(uj.child, id, SE3());
                    }
                    else
                        throw std::runtime_error("urdf: unsupported joint type '" + uj.type + "' (" + uj.name + ")");
                }
            }
        } ctx{model, links, joints};
        ctx.visit(root, root_joint, SE3());
        return model;
    }

    // pinocchio::srdf::loadReferenceConfigurations: <group_state name=..><joint name= value=/>
    inline void load_srdf_configurations(Model& model, const std::string& path)
    {
        xml::Element robot = xml::LoadFile(path);
        for (auto& gs : robot.children) {
            if (gs.name != "group_state")
                continue;
            std::vector<double> q(model.nq, 0.0);
            if (model.floating_base)
                q[6] = 1.0; // neutral quaternion
            for (auto& j : gs.children) {
                if (j.name != "joint")
                    continue;
                std::string jn = j.attr_or("name", "");
                auto vals = detail::numbers(j.attr_or("value", ""));
                int id = model.joint_id(jn);
                if (id < 0)
                    continue; // pinocchio ignores joints that are not in the model (e.g. wheels of a no-wheel URDF)
                int nqj = model.joints[id].type == FREEFLYER ? 7 : 1;
                if ((int)vals.size() != nqj)
                    throw std::runtime_error("srdf: joint '" + jn + "' expects " + std::to_string(nqj) + " values");
                for (int k = 0; k < nqj; ++k)
                    q[model.joints[id].idx_q + k] = vals[k];
            }
            model.reference_configurations[gs.attr_or("name", "")] = q;
        }
    }

    // construction-time forward kinematics (positions only) and centre of mass
    struct Kinematics {
        std::vector<SE3> oMi;
        double com[3];
        double mass;
    };
    inline SE3 joint_transform(const Joint& j, const double* q)
    {
        SE3 m;
        if (j.type == FREEFLYER) {
            const double* p = q + j.idx_q;
            quat_R(p[3], p[4], p[5], p[6], m.R);
            for (int i = 0; i < 3; ++i)
                m.p[i] = p[i];
        }
        else if (j.type == REVOLUTE)
            axis_angle_R(j.axis, q[j.idx_q], m.R);
        else
            for (int i = 0; i < 3; ++i)
                m.p[i] = j.axis[i] * q[j.idx_q];
        return m;
    }
    inline Kinematics forward_kinematics(const Model& model, const std::vector<double>& q)
    {
        Kinematics k;
        k.oMi.resize(model.joints.size());
        k.mass = 0;
        k.com[0] = k.com[1] = k.com[2] = 0;
        for (size_t i = 0; i < model.joints.size(); ++i) {
            const Joint& j = model.joints[i];
            SE3 li = mul(j.placement, joint_transform(j, q.data()));
            k.oMi[i] = j.parent >= 0 ? mul(k.oMi[j.parent], li) : li;
            double c[3];
            rot_apply(k.oMi[i].R, j.Y.c, c);
            for (int a = 0; a < 3; ++a)
                k.com[a] += j.Y.m * (c[a] + k.oMi[i].p[a]);
            k.mass += j.Y.m;
        }
        for (int a = 0; a < 3; ++a)
            k.com[a] /= k.mass;
        return k;
    }
    inline SE3 frame_placement(const Model& model, const Kinematics& k, int frame)
    {
        const Frame& f = model.frames[frame];
        return f.parent >= 0 ? mul(k.oMi[f.parent], f.placement) : f.placement;
    }

} // namespace host
} // namespace iwbc
