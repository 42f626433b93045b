// host_state.hpp -- the opaque handle behind `iwbc_problem` and the QP-shape bookkeeping shared by
// the two halves of the C ABI.
#pragma once
#include <string>

#include "stack.hpp"

struct iwbc_problem {
    iwbc::host::CompiledStack cs;
    unsigned long revision = 0; // bumped when a weight changes so that batches re-upload the problem
};

namespace iwbc {
void set_error(const std::string& e);

namespace host {
    struct QPShape {
        int n, me, mi, rows, nk;
    };
    // nVar / nEq / nIn of InverseDynamicsFormulationAccForce for a given set of active contacts
    // (the reference prints them at src/controllers/pos_tracker.cpp:150-158)
    inline QPShape qp_shape(const iwbc_problem_t& P, unsigned mask)
    {
        QPShape s = {0, 0, 0, 0, 0};
        for (int c = 0; c < P.ncontacts; ++c)
            if (mask & (1u << c))
                s.nk += 12;
        s.n = P.nv + s.nk;
        for (int b = 0; b < P.nl0; ++b) {
            const iwbc_block_t& B = P.l0[b];
            bool on = (B.kind != IWBC_B_CONTACT_MOTION && B.kind != IWBC_B_CONTACT_FORCE) || (mask & (1u << B.index));
            if (!on)
                continue;
            if (B.kind == IWBC_B_BASEDYN)
                s.me += P.nv - P.na;
            else if (B.kind == IWBC_B_CONTACT_MOTION)
                s.me += 6;
            else if (B.kind == IWBC_B_CONTACT_FORCE)
                s.mi += 34;
            else
                s.mi += 2 * P.tasks[B.index].nrows;
        }
        for (int b = 0; b < P.nl1; ++b) {
            const iwbc_block_t& B = P.l1[b];
            if (B.kind == IWBC_B_CONTACT_REG)
                s.rows += (mask & (1u << B.index)) ? 6 : 0;
            else
                s.rows += P.tasks[B.index].nrows;
        }
        return s;
    }
} // namespace host
} // namespace iwbc
