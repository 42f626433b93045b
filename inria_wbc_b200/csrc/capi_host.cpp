// capi_host.cpp -- host-only half of the C ABI (include/iwbc_batch.h): config -> compiled problem.
// No CUDA here; the batch half is batch.cu.
#include <cstring>
#include <functional>
#include <string>

#include "../../include/iwbc_batch.h"
#include "host_state.hpp"

namespace iwbc {
thread_local std::string g_last_error;
void set_error(const std::string& e) { g_last_error = e; }
} // namespace iwbc

using namespace iwbc;

extern "C" {

const char* iwbc_last_error(void) { return g_last_error.c_str(); }
int iwbc_abi_version(void) { return 1; }

iwbc_problem* iwbc_problem_load(const char* controller_yaml, const char* robots_path, const char* etc_path, const char* overrides_yaml)
{
    try {
        if (!controller_yaml || !robots_path)
            throw std::runtime_error("iwbc_problem_load: NULL path");
        std::string path(controller_yaml);
        yaml::Node config = yaml::LoadFile(path);
        if (!config["CONTROLLER"].is_map())
            throw host::ConfigError("config error: missing CONTROLLER section in " + path);
        if (overrides_yaml && *overrides_yaml) {
            yaml::Node ov = yaml::Load(overrides_yaml);
            if (!ov.is_map())
                throw host::ConfigError("overrides must be a YAML map of CONTROLLER keys");
            // nested keys are merged, not replaced: the reference's tests write `c["CONTROLLER"]["stabilizer"]["activated"] = false`
            // and keep the map's other entries (tests/test_determinism.cpp:67-69)
            std::function<void(yaml::Node&, const yaml::Node&)> merge = [&](yaml::Node& dst, const yaml::Node& src) {
                for (auto& kv : src.map) {
                    if (kv.second.is_map() && dst[kv.first].is_map())
                        merge(dst.at(kv.first), kv.second);
                    else
                        dst.at(kv.first) = kv.second;
                }
            };
            merge(config.at("CONTROLLER"), ov);
        }
        std::string etc;
        if (etc_path && *etc_path)
            etc = etc_path;
        else {
            // <etc>/<robot>/<file>.yaml -> <etc>
            size_t a = path.find_last_of('/');
            std::string dir = a == std::string::npos ? "." : path.substr(0, a);
            size_t b = dir.find_last_of('/');
            etc = b == std::string::npos ? "." : dir.substr(0, b);
        }
        auto* p = new iwbc_problem;
        p->cs = host::compile_stack(config, robots_path, etc);
        return p;
    }
    catch (std::exception& e) {
        set_error(e.what());
        return nullptr;
    }
}

void iwbc_problem_free(iwbc_problem* p) { delete p; }

const iwbc_problem_t* iwbc_problem_data(const iwbc_problem* p) { return p ? &p->cs.prob : nullptr; }

int iwbc_problem_dims(const iwbc_problem* p, int* nq, int* nv, int* na, int* ncontacts, int* nref)
{
    if (!p) {
        set_error("NULL problem");
        return IWBC_ERR_ARG;
    }
    const iwbc_problem_t& P = p->cs.prob;
    if (nq)
        *nq = P.nq;
    if (nv)
        *nv = P.nv;
    if (na)
        *na = P.na;
    if (ncontacts)
        *ncontacts = P.ncontacts;
    if (nref)
        *nref = P.nref;
    return IWBC_OK;
}

int iwbc_problem_qp_shape(const iwbc_problem* p, unsigned contact_mask, int* n, int* me, int* mi, int* cost_rows)
{
    if (!p) {
        set_error("NULL problem");
        return IWBC_ERR_ARG;
    }
    host::QPShape s = host::qp_shape(p->cs.prob, contact_mask);
    if (n)
        *n = s.n;
    if (me)
        *me = s.me;
    if (mi)
        *mi = s.mi;
    if (cost_rows)
        *cost_rows = s.rows;
    return IWBC_OK;
}

unsigned iwbc_problem_default_contact_mask(const iwbc_problem* p)
{
    unsigned m = 0;
    if (p)
        for (int c = 0; c < p->cs.prob.ncontacts; ++c)
            if (p->cs.prob.contacts[c].active_default)
                m |= 1u << c;
    return m;
}

int iwbc_problem_default_refs(const iwbc_problem* p, double* out)
{
    if (!p || !out) {
        set_error("NULL argument");
        return IWBC_ERR_ARG;
    }
    std::memcpy(out, p->cs.ref0.data(), sizeof(double) * p->cs.ref0.size());
    return IWBC_OK;
}

int iwbc_problem_ref_slot(const iwbc_problem* p, const char* name, int* offset, int* length)
{
    if (!p || !name) {
        set_error("NULL argument");
        return IWBC_ERR_ARG;
    }
    const iwbc_problem_t& P = p->cs.prob;
    for (size_t i = 0; i < p->cs.task_names.size(); ++i)
        if (p->cs.task_names[i] == name) {
            const iwbc_task_t& t = P.tasks[i];
            if (t.ref_off < 0)
                break;
            int len = t.type == IWBC_T_SE3 ? IWBC_REF_SE3 : t.type == IWBC_T_COM ? IWBC_REF_COM
                : t.type == IWBC_T_MOMENTUM                                      ? IWBC_REF_MOM
                : t.type == IWBC_T_TORQUE                                        ? 2 * P.na
                : t.type == IWBC_T_COP                                           ? 6
                                                                                 : 3 * P.na;
            if (offset)
                *offset = t.ref_off;
            if (length)
                *length = len;
            return IWBC_OK;
        }
    for (size_t i = 0; i < p->cs.contact_names.size(); ++i)
        if (p->cs.contact_names[i] == name) {
            if (offset)
                *offset = P.contacts[i].ref_off;
            if (length)
                *length = IWBC_REF_CONTACT; // SE3 sample (24) + f_ref (6) + weight-vector index + pad
            return IWBC_OK;
        }
    set_error(std::string("no task or contact with a reference named '") + name + "'");
    return IWBC_ERR_ARG;
}

int iwbc_problem_set_task_weight(iwbc_problem* p, const char* task, double weight)
{
    if (!p || !task) {
        set_error("NULL argument");
        return IWBC_ERR_ARG;
    }
    for (size_t i = 0; i < p->cs.task_names.size(); ++i)
        if (p->cs.task_names[i] == task) {
            p->cs.prob.tasks[i].weight = weight;
            p->revision++;
            return IWBC_OK;
        }
    set_error(std::string("Unknown task: ") + task);
    return IWBC_ERR_ARG;
}

int iwbc_problem_contact_index(const iwbc_problem* p, const char* contact)
{
    if (p && contact)
        for (size_t i = 0; i < p->cs.contact_names.size(); ++i)
            if (p->cs.contact_names[i] == contact)
                return (int)i;
    set_error("unknown contact");
    return IWBC_ERR_ARG;
}

int iwbc_problem_name(const iwbc_problem* p, int kind, int i, char* buf, int buflen)
{
    if (!p)
        return IWBC_ERR_ARG;
    std::vector<std::string> names;
    switch (kind) {
    case 0:
        names = p->cs.model.actuated_names();
        break;
    case 1:
        names = p->cs.task_names;
        break;
    case 2:
        names = p->cs.contact_names;
        break;
    case 3:
        names = p->cs.mimic_dof_names;
        break;
    default:
        return IWBC_ERR_ARG;
    }
    if (buf && buflen > 0 && i >= 0 && i < (int)names.size()) {
        std::strncpy(buf, names[i].c_str(), buflen - 1);
        buf[buflen - 1] = 0;
    }
    return (int)names.size();
}

int iwbc_problem_closed_loop(const iwbc_problem* p) { return p && p->cs.closed_loop ? 1 : 0; }

int iwbc_problem_q0(const iwbc_problem* p, double* out)
{
    if (!p || !out)
        return IWBC_ERR_ARG;
    std::memcpy(out, p->cs.prob.q0, sizeof(double) * p->cs.prob.nq);
    return IWBC_OK;
}

} // extern "C"
