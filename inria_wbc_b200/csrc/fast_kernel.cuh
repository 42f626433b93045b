// fast_kernel.cuh -- the "first pass" of the batched control cycle for floating-base robots with contacts.
//
// eiquadprog's Goldfarb-Idnani loop starts from the minimiser of the cost on the
// EQUALITY constraints and only then looks at the inequalities.  On the reference's own workloads most
// states end there (Talos double-support squat: 91 % of instances finish with an empty active set), so
// this kernel computes exactly that point -- dynamics, task/contact rows, H, the equality-constrained
// minimiser x_eq, s = CI x_eq + ci0 -- and
//   * if eiquadprog's termination test passes at x_eq (same formula: |sum min(0, s)| <= mi eps c1 c2 100),
//     decodes and writes the results (status OPTIMAL, 1 iteration, empty active set);
//   * otherwise appends the instance to a list that the active-set kernel (cycle_kernel.cuh) then steps
//     through the full Goldfarb-Idnani iteration.
// The point of the split is occupancy: without the n x n matrix J and the triangular factor R of the
// active-set method an instance needs ~63 kB of shared memory instead of 134 kB, so three CTAs fit on an
// SM instead of one, and the equality phase collapses from 18 sequential Givens sweeps over J to
//   1. a left-looking Cholesky of the dv-block of H (H is block diagonal: the contact-force blocks are
//      constants of the problem and are factorised once on the host) whose matrix is AUGMENTED with the
//      constraint normals and the gradient as extra rows, so that B = L^-1 N and y = L^-1 g fall out of
//      the same sweep,
//   2. a Householder QR of B (n x me),
//   3. x_eq = L^-T w,  w = -Q [R^-T ce0 ; (Q' y)_2]   (compact-WY application of Q via the Gram matrix
//      of the reflectors, then one back-substitution).
// x_eq, the multipliers and s are the same quantities the reference computes (they do not depend on how
// Q is built); parity is checked against the oracle in tests/.
//
// Kinematics / dynamics are evaluated in world coordinates without level-by-level recursion: every joint
// composes the transforms along its own root path (the path table is static), velocities and drift
// accelerations are prefix sums of per-joint terms along the same path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

#include "cycle_kernel.cuh"

namespace iwbc {
namespace dev {

#define IWBC_FMAXD 20   /* max tree depth + 1 handled by the fast path */
#define IWBC_FMAXJOBS 96
#define IWBC_FMAXBLK 40
#define IWBC_FMAXROWS 96
#define IWBC_FMAXME 24
#define IWBC_FMAXCHUNK 3 /* row chunks of the Hessian assembly */
#define IWBC_FMAXPAIRS ((IWBC_FMAXME + 1) * (IWBC_FMAXME + 2) / 2)

    enum { FJ_SE3 = 0, FJ_CONTACT = 1, FJ_SELFCOL = 2, FJ_COM = 3, FJ_MOM = 4 };
    enum { FB_SE3 = 0, FB_COM = 1, FB_MOM = 2, FB_SELFCOL = 3, FB_CONTACT = 4 };

    struct FJob {
        short kind, task, a, slot; // slot: rhs slot (SE3 / COM / MOM) or pair slot (SELFCOL)
    };
    struct FBlock {
        short kind, index, row0, nrows, slot0;
        // flattened from the task / contact record so that the row loop reads ONE record per block instead of a chain of dependent
        // global loads (block -> task -> frame -> parent joint): frame and its parent joint (-1: none), component mask bits, and for
        // a self-collision block the number of avoided frames (their frames / parent joints: FPlan::pair_frame / pair_pj by pair slot)
        short frame, pj, maskbits, nav, pad0, pad1, pad2;
    };
    struct FRow {
        short kind, task, comp, slot; // kind: FB_*; comp: component of the 6-/3-vector; slot: rhs slot (or first pair slot)
    };

    // static tables of one (problem, contact mask): global memory, read through L1
    struct FPlan {
        unsigned char path[IWBC_MAXJ][IWBC_FMAXD]; // joints from the root to j (path[j][depth[j]] = j)
        unsigned char dofj[IWBC_MAXNV];
        unsigned char tile_i[136], tile_j[136];
        unsigned char pair_i[IWBC_FMAXPAIRS], pair_k[IWBC_FMAXPAIRS];
        unsigned char tri_r[1830], tri_c[1830]; // packed lower triangle (nv <= 60): entry e = r (r + 1) / 2 + c -> (r, c)
        unsigned char tri_sup[1830];            // ... and whether joint(c) supports joint(r), i.e. M(r, c) is structurally non-zero
        FJob jobs[IWBC_FMAXJOBS];
        FBlock blocks[IWBC_FMAXBLK];
        FRow rows[IWBC_FMAXROWS];
        double hdiag[IWBC_MAXNV];   // posture weights on the diagonal of H + the 1e-8 regularisation
        double blk_sw[IWBC_FMAXBLK], row_sw[IWBC_FMAXROWS]; // sqrt(task weight) per row block / per row
        short pair_frame[IWBC_FMAXJOBS], pair_pj[IWBC_FMAXJOBS]; // self-collision pair slot -> avoided frame, its parent joint
        // Force blocks of H, one per (regularisation weight vector, active-contact slot): H_ff = w T' W^2 T + 1e-8 I depends on the
        // problem and on WHICH weight vector W the instance has installed (iwbc_problem_t::wreg_variants: Contact6d's default or a
        // stabilizer's zmp_w), so it is factorised on the host for every variant and the kernel picks by the instance's index.
        double Uff[IWBC_MAXWV][IWBC_MAXC][144]; // L_ff^-T (upper triangular, row-major)
        double Gf[IWBC_MAXWV][IWBC_MAXC][72];   // L_ff^-1 g_f = Gf f_ref (12 x 6, row-major): g_f = -w T' W^2 f_ref is per instance
        double UT[IWBC_MAXWV][IWBC_MAXC][72];   // U_ff' T' (12 x 6, row-major): (U_ff' N_f)[k][r] = -sum_a UT[k][a] Jft[a][r]
        double c1v[IWBC_MAXWV][IWBC_MAXC], c2v[IWBC_MAXWV][IWBC_MAXC]; // trace(H_ff), trace(L_ff^-T)
    };

    struct FShape {
        int ok;    // the fast path applies
        int nca;   // active contacts
        int cact[IWBC_MAXC];
        int me, cb0, nf, NC; // equalities; first base-dynamics column (= 6 nca); force-block columns (= nu + 1); QR columns (= me + 1)
        int ldh, nrh;        // leading dimension / rows of the augmented matrix (nv + me + 1)
        int nvp;             // row stride of A (nv rounded up to 4)
        int njobs, nblocks, nrows, ntiles, npairs, nposture, nslots;
        int ci_bounds, ci_act; // index of the bounds / actuation-bounds block in Shape::ci (-1: none)
        int posture[4];
        // shared-memory offsets (doubles)
        int o_qs, o_vs, o_X, o_Y, o_Z, o_Jft, o_cb, o_nle, o_lb, o_ub, o_g, o_x, o_w, o_sc, o_fr, o_rhs, o_sc4, o_P, o_Rm, o_tv, o_hd, o_il, o_tau, o_int, o_Md, total;
        int d_liMi, d_oMi, d_Sw, d_vw, d_aw, d_Yw, d_fw, d_Fc, d_oMf, d_vf, d_af;
        int z_Bf, z_G;
        int s_J, s_R, s_vec, s_int, s_M, s_total; // extra regions of the active-set continuation (k_slow)
        // The dense level-1 rows are formed and folded into H in chunks of whole row blocks (the row buffer is what decides whether
        // four CTAs fit on an SM): chunk c = blocks [ch_b0, ch_b1), rows [ch_r0, ch_r0 + ch_nr); crows = the largest chunk.
        int nchunk, crows;
        int ch_b0[IWBC_FMAXCHUNK], ch_b1[IWBC_FMAXCHUNK], ch_r0[IWBC_FMAXCHUNK], ch_nr[IWBC_FMAXCHUNK];
    };


    // ------------------------------------------------------------------------------------------ host: plan
    inline bool build_fast_plan(const iwbc_problem_t& P, const Shape& S, int T, FShape& F, FPlan& pl)
    {
        std::memset(&F, 0, sizeof(F));
        std::memset(&pl, 0, sizeof(pl));
        const int nv = P.nv, nj = P.nj, na = P.na, nu = nv - na;
        if (!P.floating_base || nu != 6 || S.me == 0 || S.me > IWBC_FMAXME)
            return false;
        for (int c = 0; c < P.ncontacts; ++c)
            if (S.cidx[c] >= 0)
                F.cact[F.nca++] = c;
        if (F.nca == 0 || S.me != 6 * F.nca + nu || 2 * S.me > nv)
            return false;
        for (int j = 0; j < nj; ++j) {
            if (P.depth[j] + 1 > IWBC_FMAXD)
                return false;
            int a = j;
            for (int d = P.depth[j]; d >= 0; --d) {
                pl.path[j][d] = (unsigned char)a;
                a = P.parent[a];
            }
            const int c0 = P.idx_v[j], cn = P.jtype[j] == IWBC_J_FREEFLYER ? 6 : 1;
            for (int k = 0; k < cn; ++k)
                pl.dofj[c0 + k] = (unsigned char)j;
        }
        F.me = S.me;
        F.cb0 = 6 * F.nca;
        F.nf = nu + 1;
        F.NC = S.me + 1;
        F.nrh = nv + S.me + 1;
        F.ldh = (F.nrh + 1) & ~1; // even: the Cholesky panel reads pairs of rows as double2
        F.nvp = (nv + 3) & ~3;
        // ---- jobs, row blocks, rows (level-1 order = Hessian summation order)
        int nslots = 0, npairslots = 0, row = 0;
        for (int k = 0; k < nv; ++k)
            pl.hdiag[k] = P.hess_reg;
        for (int bl = 0; bl < P.nl1; ++bl) {
            if (P.l1[bl].kind != IWBC_B_TASK)
                continue;
            const int t = P.l1[bl].index;
            const iwbc_task_t& Tk = P.tasks[t];
            if (Tk.type == IWBC_T_POSTURE) {
                if (F.nposture >= 4)
                    return false;
                F.posture[F.nposture++] = t;
                for (int i = 0; i < na; ++i)
                    if (Tk.mask[i])
                        pl.hdiag[nu + i] += Tk.weight;
                continue;
            }
            if (F.nblocks >= IWBC_FMAXBLK || row + Tk.nrows > IWBC_FMAXROWS)
                return false;
            FBlock& B = pl.blocks[F.nblocks++];
            B.index = (short)t, B.row0 = (short)row, B.nrows = (short)Tk.nrows;
            if (Tk.type == IWBC_T_SE3 || Tk.type == IWBC_T_COM || Tk.type == IWBC_T_MOMENTUM) {
                if (F.njobs >= IWBC_FMAXJOBS)
                    return false;
                const int ncomp = Tk.type == IWBC_T_COM ? 3 : 6;
                B.kind = Tk.type == IWBC_T_SE3 ? FB_SE3 : Tk.type == IWBC_T_COM ? FB_COM : FB_MOM;
                B.slot0 = (short)nslots;
                pl.jobs[F.njobs++] = {(short)(Tk.type == IWBC_T_SE3 ? FJ_SE3 : Tk.type == IWBC_T_COM ? FJ_COM : FJ_MOM), (short)t, 0, (short)nslots};
                int idx = 0;
                for (int i = 0; i < ncomp; ++i)
                    if (Tk.mask[i])
                        pl.rows[row + idx++] = {B.kind, (short)t, (short)i, (short)nslots};
                if (idx != Tk.nrows)
                    return false;
                nslots++;
            }
            else if (Tk.type == IWBC_T_SELFCOL) {
                B.kind = FB_SELFCOL;
                B.slot0 = (short)npairslots;
                for (int a = 0; a < Tk.nav; ++a) {
                    if (F.njobs >= IWBC_FMAXJOBS)
                        return false;
                    pl.jobs[F.njobs++] = {(short)FJ_SELFCOL, (short)t, (short)a, (short)(npairslots + a)};
                }
                pl.rows[row] = {FB_SELFCOL, (short)t, 0, (short)npairslots};
                npairslots += Tk.nav;
            }
            else
                return false;
            row += Tk.nrows;
        }
        if (row != S.rows_dense)
            return false;
        F.nrows = row;
        for (int s = 0; s < F.nca; ++s) {
            if (F.njobs >= IWBC_FMAXJOBS || F.nblocks >= IWBC_FMAXBLK)
                return false;
            pl.jobs[F.njobs++] = {(short)FJ_CONTACT, (short)F.cact[s], 0, 0};
            FBlock& B = pl.blocks[F.nblocks++];
            B.kind = FB_CONTACT, B.index = (short)F.cact[s], B.row0 = 0, B.nrows = 6, B.slot0 = 0;
        }
        F.nslots = nslots;
        for (int bi = 0; bi < F.nblocks; ++bi) {
            FBlock& B = pl.blocks[bi];
            pl.blk_sw[bi] = B.kind == FB_CONTACT ? 1.0 : std::sqrt(P.tasks[B.index].weight);
            B.frame = B.pj = -1, B.maskbits = 0, B.nav = 0;
            if (B.kind == FB_CONTACT)
                B.frame = (short)P.contacts[B.index].frame;
            else {
                const iwbc_task_t& Tk = P.tasks[B.index];
                if (B.kind == FB_SE3 || B.kind == FB_SELFCOL)
                    B.frame = (short)Tk.frame;
                for (int i = 0; i < 6; ++i)
                    if ((B.kind == FB_SE3 || B.kind == FB_MOM || (B.kind == FB_COM && i < 3)) && Tk.mask[i])
                        B.maskbits |= (short)(1 << i);
                if (B.kind == FB_SELFCOL) {
                    B.nav = (short)Tk.nav;
                    for (int a = 0; a < Tk.nav; ++a) {
                        pl.pair_frame[B.slot0 + a] = (short)Tk.av_frame[a];
                        pl.pair_pj[B.slot0 + a] = (short)P.fparent[Tk.av_frame[a]];
                    }
                }
            }
            if (B.frame >= 0)
                B.pj = (short)P.fparent[B.frame];
        }
        for (int r = 0; r < F.nrows; ++r)
            pl.row_sw[r] = std::sqrt(P.tasks[pl.rows[r].task].weight);
        // ---- row chunks: the smallest capacity that packs the row blocks (in order) into at most IWBC_FMAXCHUNK chunks; the contact
        // blocks (no rows of A: they fill Jft) ride in the last chunk
        {
            const int ntask = F.nblocks - F.nca;
            for (int cap = 1; cap <= IWBC_FMAXROWS; ++cap) {
                int nch = 0, used = 0;
                bool fits = true;
                for (int bi = 0; bi < ntask && fits; ++bi) {
                    const int nr = pl.blocks[bi].nrows;
                    if (nr > cap)
                        fits = false;
                    else if (nch == 0 || used + nr > cap) {
                        if (nch == IWBC_FMAXCHUNK)
                            fits = false;
                        else {
                            F.ch_b0[nch] = bi, F.ch_r0[nch] = pl.blocks[bi].row0, F.ch_nr[nch] = 0;
                            ++nch;
                            used = 0;
                        }
                    }
                    if (fits) {
                        used += nr;
                        F.ch_nr[nch - 1] = used;
                        F.ch_b1[nch - 1] = bi + 1;
                    }
                }
                if (fits) {
                    if (nch == 0) {
                        F.ch_b0[0] = F.ch_b1[0] = 0, F.ch_r0[0] = F.ch_nr[0] = 0;
                        nch = 1;
                    }
                    F.nchunk = nch;
                    F.crows = cap;
                    break;
                }
            }
            if (F.nchunk == 0)
                return false;
            F.ch_b1[F.nchunk - 1] = F.nblocks;
            F.crows = 0;
            for (int c = 0; c < F.nchunk; ++c)
                F.crows = std::max(F.crows, F.ch_nr[c]);
        }
        F.ci_bounds = F.ci_act = -1;
        for (int b = 0; b < S.nci; ++b) {
            if (S.ci[b].kind == 0) {
                if (F.ci_bounds >= 0 || S.ci[b].rows != na)
                    return false;
                F.ci_bounds = b;
            }
            else if (S.ci[b].kind == 1) {
                if (F.ci_act >= 0 || S.ci[b].rows != na)
                    return false;
                F.ci_act = b;
            }
        }
        // ---- H tiles (4 x 4, lower triangle) and reflector pairs
        const int ntb = F.nvp / 4;
        for (int ti = 0; ti < ntb; ++ti)
            for (int tj = 0; tj <= ti; ++tj) {
                if (F.ntiles >= 136)
                    return false;
                pl.tile_i[F.ntiles] = (unsigned char)ti, pl.tile_j[F.ntiles] = (unsigned char)tj;
                F.ntiles++;
            }
        if (T > 1 && F.ntiles > T)
            return false;
        for (int r = 0, e = 0; r < nv && r < 60; ++r)
            for (int c = 0; c <= r; ++c, ++e) {
                pl.tri_r[e] = (unsigned char)r, pl.tri_c[e] = (unsigned char)c;
                const int jc = pl.dofj[c], jr = pl.dofj[r];
                pl.tri_sup[e] = (unsigned char)(jc <= jr && jr <= P.last_desc[jc]);
            }
        for (int k = 0; k <= F.me; ++k) // k == me: the transformed right-hand side t
            for (int i = 0; i < k && i < F.me; ++i) {
                pl.pair_i[F.npairs] = (unsigned char)i, pl.pair_k[F.npairs] = (unsigned char)k;
                F.npairs++;
            }
        // ---- constant force blocks: H_ff = w T' W^2 T + reg I, per weight-vector variant and active contact
        for (int wv = 0; wv < (P.nwreg > 0 ? P.nwreg : 1); ++wv)
        for (int s = 0; s < F.nca; ++s) {
            const iwbc_contact_t& C = P.contacts[F.cact[s]];
            const double* wreg = P.nwreg > 0 ? P.wreg_variants[wv] : C.wreg;
            double H[144], L[144];
            double c1 = 0, c2 = 0;
            for (int i = 0; i < 12; ++i) {
                for (int j = 0; j < 12; ++j) {
                    double a = 0;
                    for (int r = 0; r < 6; ++r)
                        a += (wreg[r] * C.T[r * 12 + i]) * (wreg[r] * C.T[r * 12 + j]);
                    H[i * 12 + j] = C.w_forcereg * a + (i == j ? P.hess_reg : 0.0);
                }
                c1 += H[i * 12 + i];
            }
            std::memset(L, 0, sizeof(L));
            for (int k = 0; k < 12; ++k) {
                double d = H[k * 12 + k];
                for (int j = 0; j < k; ++j)
                    d -= L[k * 12 + j] * L[k * 12 + j];
                if (!(d > 0))
                    return false;
                L[k * 12 + k] = std::sqrt(d);
                for (int i = k + 1; i < 12; ++i) {
                    double a = H[i * 12 + k];
                    for (int j = 0; j < k; ++j)
                        a -= L[i * 12 + j] * L[k * 12 + j];
                    L[i * 12 + k] = a / L[k * 12 + k];
                }
            }
            // Linv (lower) by forward substitution, U = Linv^T
            double Li[144];
            std::memset(Li, 0, sizeof(Li));
            for (int c = 0; c < 12; ++c)
                for (int i = c; i < 12; ++i) {
                    double a = (i == c) ? 1.0 : 0.0;
                    for (int j = c; j < i; ++j)
                        a -= L[i * 12 + j] * Li[j * 12 + c];
                    Li[i * 12 + c] = a / L[i * 12 + i];
                }
            for (int k2 = 0; k2 < 12; ++k2)
                for (int a = 0; a < 6; ++a) {
                    double w = 0;
                    for (int i = 0; i <= k2; ++i)
                        w += Li[k2 * 12 + i] * C.T[a * 12 + i]; // U[i][k2] = Li[k2][i]
                    pl.UT[wv][s][k2 * 6 + a] = w;
                }
            for (int i = 0; i < 12; ++i) {
                for (int j = 0; j < 12; ++j)
                    pl.Uff[wv][s][i * 12 + j] = Li[j * 12 + i];
                c2 += Li[i * 12 + i];
                // (L^-1 g_f)_i = sum_j Li[i][j] g_f[j],  g_f[j] = -w sum_r (W_r T[r][j]) (W_r f_ref[r])
                for (int r = 0; r < 6; ++r) {
                    double a = 0;
                    for (int j = 0; j <= i; ++j)
                        a += Li[i * 12 + j] * (wreg[r] * C.T[r * 12 + j]) * wreg[r];
                    pl.Gf[wv][s][i * 6 + r] = -C.w_forcereg * a;
                }
            }
            pl.c1v[wv][s] = c1, pl.c2v[wv][s] = c2;
        }
        // ---- shared-memory layout
        const int nq = P.nq, nf = P.nframes, n = S.n;
        int o = 0;
        auto take = [&](int cnt) {
            int r = o;
            o += align2(cnt);
            return r;
        };
        F.o_qs = take(nq);
        F.o_vs = take(nv);
        F.o_X = take(F.crows * F.nvp + align2(F.crows)); // one chunk of A and its right-hand side
        int d = 0;
        auto dt = [&](int cnt) {
            int r = d;
            d += align2(cnt);
            return r;
        };
        F.d_liMi = dt(12 * nj), F.d_oMi = dt(12 * nj), F.d_Sw = dt(6 * nv), F.d_vw = dt(6 * nj), F.d_aw = dt(6 * nj), F.d_Yw = dt(10 * nj);
        F.d_fw = dt(6 * nj), F.d_Fc = dt(6 * nv);
        F.d_oMf = dt(12 * nf), F.d_vf = dt(6 * nf), F.d_af = dt(6 * nf);
        if (12 * nf + 6 * nf < 6 * nj) // vJ scratch lives in [oMf, vf]
            d += 6 * nj;
        const int hs = nv * F.ldh;
        F.o_Y = take(d > hs ? d : hs);
        F.z_Bf = 0;
        F.z_G = align2(12 * F.nca * F.nf);
        const int zq = F.z_G + align2((F.me + 1) * (F.me + 1));
        // arrays that stay alive through the active-set continuation
        F.o_Jft = take(P.ncontacts * 6 * nv);
        F.o_cb = take(12 * P.ncontacts);
        F.o_nle = take(nv);
        F.o_lb = take(na); // (acceleration bounds and torques exist for the actuated joints only; every byte counts here: Talos double
        F.o_ub = take(na); //  support sits 100 B under the two-CTAs-per-SM limit of the continuation kernel)
        F.o_x = take(n + 2);
        F.o_sc = take(48);
        F.o_fr = take(8 * F.nca); // per active-contact slot: the instance's f_ref (6), its weight-vector index, pad
        F.o_int = take((nv + 16) / 2 + 1);
        F.o_Md = take(nv); // diagonal of the mass matrix (its strict lower triangle sits in the upper triangle of the augmented matrix)
        // arrays that are dead once J has been formed: the iteration vectors of the continuation alias them
        const int dead0 = o;
        F.o_Z = take(P.nref > zq ? P.nref : zq);
        F.o_g = take(n + 2);
        F.o_w = take(n + 2);
        F.o_rhs = take(6 * (nslots > 0 ? nslots : 1));
        F.o_sc4 = take(4 * (npairslots > 0 ? npairslots : 1));
        F.o_P = take(6 * F.NC > 16 ? 6 * F.NC : 16); // QR partial sums (SEG x NC); the Cholesky keeps its 4 x 4 diagonal block copy here
        F.o_Rm = take(F.me * F.NC + 3 * F.me + 8); // R (me x NC), tau, scale, invbeta
        F.o_tv = take(3 * (F.me + 2) + 4);
        F.o_hd = take(nv);
        F.o_il = take(nv);
        F.o_tau = take(na);
        F.total = o;
        // continuation: R aliases the augmented matrix (dead after J), the vectors alias the dead block
        const int rsz = n * (n + 1) / 2;
        F.s_R = F.o_Y;
        if (rsz > (d > hs ? d : hs))
            return false;
        const int vsz = 11 * align2(n + 2) + align2(S.mi) + 32;
        F.s_vec = dead0;
        o = dead0 + (vsz > F.total - dead0 ? vsz : F.total - dead0);
        F.s_int = take((2 * (n + 2) + S.mi + 1) / 2 + 1);
        F.s_J = take(n * S.ldJ);
        F.s_M = take(nv * (nv + 1) / 2); // packed lower triangle of M: R is about to overwrite the augmented matrix that held it
        F.s_total = o;
        F.ok = 1;
        return true;
    }

    // ------------------------------------------------------------------------------------------ device
#ifndef IWBC_OPT_QR
#define IWBC_OPT_QR 1   // Householder steps fully unrolled (static dead rows)
#endif
#ifndef IWBC_OPT_MTAB
#define IWBC_OPT_MTAB 1 // (row, column) of the packed mass-matrix entries from a table instead of a square root
#endif
#ifndef IWBC_OPT_BF
#define IWBC_OPT_BF 1   // constants of the problem folded on the host: U_ff' T' for the force rows of B, sqrt(task weights)
#endif
#ifndef IWBC_OPT_RELOAD
#define IWBC_OPT_RELOAD 1 // QR: pivot column double-buffered and re-read in the update instead of held in registers
#endif
#ifndef IWBC_OPT_SIGMA
#define IWBC_OPT_SIGMA 1 // QR: the next pivot column's norm is computed by its owners one step ahead
#endif
#ifndef IWBC_OPT_PREFETCH
#define IWBC_OPT_PREFETCH 1 // L2 prefetch of the next instance's inputs
#endif
#ifndef IWBC_OPT_XBS
#define IWBC_OPT_XBS 1  // back-substitution x = L^-T w unrolled over its 50 / 38 steps
#endif
#ifndef IWBC_OPT_CHOL
#define IWBC_OPT_CHOL 1 // 16-byte loads of the row buffer in the H tiles
#endif
#ifndef IWBC_FAST_MINB
#define IWBC_FAST_MINB 4 // CTAs of k_fast per SM the register budget is compiled for
#endif
    template <int T>
    struct FT {
        static constexpr int PAIR = T >= 64 ? 2 : 1;                   // threads per row in the Cholesky
        static constexpr int SEG = T >= 128 ? 6 : (T >= 64 ? 3 : 1);   // row segments in the QR
        static constexpr int NTPT = T > 1 ? 1 : 136;                   // H tiles per thread
        static constexpr int NCH = (IWBC_MAXJ * 16 + T - 1) / T;       // subtree-sum items per thread
    };

    template <int T>
    __device__ __forceinline__ void fsync()
    {
        if (T > 1)
            __syncthreads();
    }

    // block-wide sum of one value (T > 32: via shared scratch red[T/32]); result broadcast
    template <int T>
    __device__ __forceinline__ double fblock_sum(double v, double* red, int tid)
    {
        if (T == 1)
            return v;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0)
            red[tid >> 5] = v;
        fsync<T>();
        double s = red[0];
        for (int w = 1; w < (T + 31) / 32; ++w)
            s += red[w];
        fsync<T>();
        return s;
    }

    // One instance.  Returns 1 when the instance is finished here (results written), 0 when it has to go through the
    // active-set iteration (state left in shared memory for slow_continue()), -1 when H is not positive definite.
    // NV_, ME_: compile-time velocity dimension / number of equalities of the common shapes (0: read them at run time); with
    // them every leading dimension and trip count below is a constant and the inner loops unroll.
    template <int T, int NV_ = 0, int ME_ = 0>
    __device__ int fast_instance(const iwbc_problem_t* __restrict__ gP, const FPlan* __restrict__ gF, const Shape& S, const FShape& F, double* smem,
        const IO& io, int inst, int tid, long long* cyc)
    {
        const int nv = NV_ ? NV_ : gP->nv, me = ME_ ? ME_ : F.me, NC = me + 1;
        const int nq = gP->nq, nj = gP->nj, nu = 6, na = nv - nu, nfr = gP->nframes, n = ME_ ? nv + 2 * (me - 6) : S.n;
        const int ld = (nv + me + 2) & ~1, nvp = (nv + 3) & ~3; // = F.ldh, F.nvp
        double* const qs = smem + F.o_qs;
        double* const vs = smem + F.o_vs;
        double* const X = smem + F.o_X;
        double* const Y = smem + F.o_Y;
        double* const Z = smem + F.o_Z;
        double* const A = X;
        double* const bvec = X + F.crows * nvp;
        double* const Md = smem + F.o_Md;
        double* const Hs = Y;
        double* const liMi = Y + F.d_liMi;
        double* const vJ = Y + F.d_oMf; // scratch until the frames are computed
        double* const gvec = smem + F.o_g;
        double* const xv = smem + F.o_x;
        double* const wv = smem + F.o_w;
        double* const sc = smem + F.o_sc;
        double* const rhs = smem + F.o_rhs;
        double* const sc4 = smem + F.o_sc4;
        double* const Pp = smem + F.o_P;
        double* const Rm = smem + F.o_Rm;
        double* const tauv = Rm + me * NC;
        double* const scalev = tauv + me;
        double* const invb = scalev + me;
        double* const tv = smem + F.o_tv;
        double* const hd = smem + F.o_hd;
        double* const il = smem + F.o_il;
        double* const taut = smem + F.o_tau;
        double* const Bf = Z + F.z_Bf;
        double* const Gm = Z + F.z_G;
        double* const frv = smem + F.o_fr;
        Sm sm;
        sm.qs = qs, sm.vs = vs, sm.ref = Z;
        sm.oMi = Y + F.d_oMi, sm.Sw = Y + F.d_Sw, sm.vw = Y + F.d_vw, sm.aw = Y + F.d_aw, sm.Yw = Y + F.d_Yw, sm.fw = Y + F.d_fw, sm.Fc = Y + F.d_Fc;
        sm.oMf = Y + F.d_oMf, sm.vf = Y + F.d_vf, sm.af = Y + F.d_af;
        sm.M = Hs, sm.Md = Md, sm.Jft = smem + F.o_Jft, sm.cb = smem + F.o_cb, sm.nle = smem + F.o_nle, sm.lb = smem + F.o_lb, sm.ub = smem + F.o_ub;
        sm.x = xv, sm.g = gvec, sm.sc = sc;
        sm.dofj = reinterpret_cast<int*>(smem + F.o_int);

        // ---------------------------------------------------------------- load
        {
            const double* gq = io.q + (size_t)inst * nq;
            const double* gv = io.v + (size_t)inst * nv;
            const double* gr = io.refs + (size_t)inst * io.refs_stride;
            for (int k = tid; k < nq; k += T)
                qs[k] = gq[k];
            for (int k = tid; k < nv; k += T)
                vs[k] = gv[k];
            for (int k = tid; k < gP->nref; k += T)
                Z[k] = gr[k];
            // the contacts' force references and weight-vector indices outlive the reference row (Z is reused below)
            for (int e = tid; e < 8 * F.nca; e += T) {
                const int s = e >> 3, a = e & 7;
                const int off = gP->contacts[F.cact[s]].ref_off;
                double val = 0.0;
                if (a < 6)
                    val = gr[off + IWBC_REF_CONTACT_FREF + a];
                else if (a == 6) {
                    const int wv = (int)gr[off + IWBC_REF_CONTACT_WREG];
                    val = (wv >= 0 && wv < gP->nwreg) ? (double)wv : 0.0;
                }
                frv[e] = val;
            }
        }
        fsync<T>();
        // per active-contact slot: the force-block factor of the weight vector this instance has installed (registers)
        const double* const Uq0 = gF->Uff[(int)frv[6]][0];
        const double* const Uq1 = F.nca > 1 ? gF->Uff[(int)frv[14]][1] : Uq0;
        auto Usel = [&](int s) { return s == 0 ? Uq0 : (s == 1 ? Uq1 : gF->Uff[(int)frv[8 * s + 6]][s]); }; // (humanoids: two contacts)
        if (cyc && tid == 0)
            cyc[0] = clock64();

        // ---------------------------------------------------------------- joint transforms (one joint per thread) + acceleration bounds
        for (int j = tid; j < nj + na; j += T) {
            if (j >= nj) {
                const int i = j - nj;
                acc_bounds(gP, i, qs[nq - na + i], vs[nv - na + i], sm.lb[i], sm.ub[i]);
                continue;
            }
            const double* jp = gP->jplace[j];
            const int ty = gP->jtype[j];
            double jm[9];
            V3 jt = mk(0, 0, 0);
            if (ty == IWBC_J_FREEFLYER) {
                const double* qq = qs + gP->idx_q[j];
                quat_R(qq[3], qq[4], qq[5], qq[6], jm);
                jt = mk(qq[0], qq[1], qq[2]);
            }
            else if (ty == IWBC_J_REVOLUTE)
                axis_angle_R(ld3(gP->axis[j]), qs[gP->idx_q[j]], jm);
            else {
                jm[0] = jm[4] = jm[8] = 1;
                jm[1] = jm[2] = jm[3] = jm[5] = jm[6] = jm[7] = 0;
                jt = qs[gP->idx_q[j]] * ld3(gP->axis[j]);
            }
            double* li = liMi + 12 * j;
            mm3(jp, jm, li);
            st3(li + 9, rmul(jp, jt) + ld3(jp + 9));
        }
        fsync<T>();
        // ---------------------------------------------------------------- world placements along the root path, motion subspaces, joint velocities
        for (int j = tid; j < nj; j += T) {
            const unsigned char* path = gF->path[j];
            const int dep = gP->depth[j];
            double o[12];
            {
                const double* l0 = liMi + 12 * path[0];
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    o[k] = l0[k];
            }
            for (int d = 1; d <= dep; ++d) {
                const double* li = liMi + 12 * path[d];
                double t[9];
                mm3(o, li, t);
                V3 p = rmul(o, ld3(li + 9)) + ld3(o + 9);
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    o[k] = t[k];
                o[9] = p.x, o[10] = p.y, o[11] = p.z;
            }
            double* oo = sm.oMi + 12 * j;
#pragma unroll
            for (int k = 0; k < 12; ++k)
                oo[k] = o[k];
            const V3 p = ld3(o + 9);
            const int ty = gP->jtype[j], c0 = gP->idx_v[j];
            V3 vJl = mk(0, 0, 0), vJa = mk(0, 0, 0);
            if (ty == IWBC_J_FREEFLYER) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    V3 col = mk(o[k], o[3 + k], o[6 + k]);
                    st3(sm.Sw + 6 * (c0 + k), col);
                    st3(sm.Sw + 6 * (c0 + k) + 3, mk(0, 0, 0));
                    V3 lin = cross(p, col);
                    st3(sm.Sw + 6 * (c0 + 3 + k), lin);
                    st3(sm.Sw + 6 * (c0 + 3 + k) + 3, col);
                    vJl = vJl + vs[c0 + k] * col + vs[c0 + 3 + k] * lin;
                    vJa = vJa + vs[c0 + 3 + k] * col;
                }
            }
            else {
                V3 aw = rmul(o, ld3(gP->axis[j]));
                V3 lin, ang;
                if (ty == IWBC_J_REVOLUTE) {
                    ang = aw;
                    lin = cross(p, aw);
                }
                else {
                    ang = mk(0, 0, 0);
                    lin = aw;
                }
                st3(sm.Sw + 6 * c0, lin);
                st3(sm.Sw + 6 * c0 + 3, ang);
                vJl = vs[c0] * lin;
                vJa = vs[c0] * ang;
            }
            st3(vJ + 6 * j, vJl);
            st3(vJ + 6 * j + 3, vJa);
        }
        fsync<T>();
        // ---------------------------------------------------------------- velocities / drift accelerations (prefix sums along the path), inertias, momenta, forces
        for (int j = tid; j < nj; j += T) {
            const unsigned char* path = gF->path[j];
            const int dep = gP->depth[j];
            V3 vl = mk(0, 0, 0), w = mk(0, 0, 0), al = mk(0, 0, 0), aa = mk(0, 0, 0);
            for (int d = 0; d <= dep; ++d) {
                const double* pj = vJ + 6 * path[d];
                const V3 jl = ld3(pj), ja = ld3(pj + 3);
                al = al + cross(vl, ja) + cross(w, jl);
                aa = aa + cross(w, ja);
                vl = vl + jl;
                w = w + ja;
            }
            st3(sm.vw + 6 * j, vl);
            st3(sm.vw + 6 * j + 3, w);
            st3(sm.aw + 6 * j, al);
            st3(sm.aw + 6 * j + 3, aa);
            const double* o = sm.oMi + 12 * j;
            const double m = gP->mass[j];
            V3 c = rmul(o, ld3(gP->com[j])) + ld3(o + 9);
            const double* I = gP->inertia[j];
            double Il[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
            double t[9], Iw[9];
            mm3(o, Il, t);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb)
                    Iw[3 * a + bb] = t[3 * a] * o[3 * bb] + t[3 * a + 1] * o[3 * bb + 1] + t[3 * a + 2] * o[3 * bb + 2];
            const double cc = dot(c, c);
            const double Ixx = Iw[0] + m * (cc - c.x * c.x), Ixy = Iw[1] - m * c.x * c.y, Ixz = Iw[2] - m * c.x * c.z;
            const double Iyy = Iw[4] + m * (cc - c.y * c.y), Iyz = Iw[5] - m * c.y * c.z, Izz = Iw[8] + m * (cc - c.z * c.z);
            V3 h = m * c;
            double* Yj = sm.Yw + 10 * j;
            Yj[0] = m;
            st3(Yj + 1, h);
            Yj[4] = Ixx, Yj[5] = Ixy, Yj[6] = Ixz, Yj[7] = Iyy, Yj[8] = Iyz, Yj[9] = Izz;
            V3 alg = al - ld3(gP->gravity);
            auto Iapply = [&](V3 x) { return mk(Ixx * x.x + Ixy * x.y + Ixz * x.z, Ixy * x.x + Iyy * x.y + Iyz * x.z, Ixz * x.x + Iyz * x.y + Izz * x.z); };
            V3 hl = m * vl - cross(h, w), ha = Iapply(w) + cross(h, vl);
            V3 fl = m * alg - cross(h, aa) + cross(w, hl);
            V3 fa = Iapply(aa) + cross(h, alg) + cross(w, ha) + cross(vl, hl);
            st3(sm.fw + 6 * j, fl);
            st3(sm.fw + 6 * j + 3, fa);
            st3(sm.Fc + 6 * j, hl); // body momenta (only their total is needed): Fc is free until the dof pass
            st3(sm.Fc + 6 * j + 3, ha);
        }
        fsync<T>();
        // ---------------------------------------------------------------- subtree sums (composite inertia, subtree force) + momentum total
        {
            double acc[FT<T>::NCH];
            const int items = nj * 16;
#pragma unroll
            for (int c8 = 0; c8 < FT<T>::NCH; ++c8) {
                const int it = c8 * T + tid;
                double s = 0;
                if (it < items) {
                    const int j = it >> 4, k = it & 15;
                    const int last = gP->last_desc[j];
                    // (two chains per item: the root's items sum over every joint)
                    double s1 = 0;
                    if (k < 10) {
                        int jj = j;
                        for (; jj + 1 <= last; jj += 2) {
                            s += sm.Yw[10 * jj + k];
                            s1 += sm.Yw[10 * (jj + 1) + k];
                        }
                        if (jj <= last)
                            s += sm.Yw[10 * jj + k];
                    }
                    else {
                        int jj = j;
                        for (; jj + 1 <= last; jj += 2) {
                            s += sm.fw[6 * jj + (k - 10)];
                            s1 += sm.fw[6 * (jj + 1) + (k - 10)];
                        }
                        if (jj <= last)
                            s += sm.fw[6 * jj + (k - 10)];
                    }
                    s += s1;
                }
                acc[c8] = s;
            }
            for (int k = tid; k < 6; k += T) { // total momentum -> sc[10..15]
                double s = 0;
                for (int j = 0; j < nj; ++j)
                    s += sm.Fc[6 * j + k];
                sc[10 + k] = s;
            }
            fsync<T>();
#pragma unroll
            for (int c8 = 0; c8 < FT<T>::NCH; ++c8) {
                const int it = c8 * T + tid;
                if (it < items) {
                    const int j = it >> 4, k = it & 15;
                    if (k < 10)
                        sm.Yw[10 * j + k] = acc[c8];
                    else
                        sm.fw[6 * j + (k - 10)] = acc[c8];
                }
            }
            fsync<T>();
        }
        // ---------------------------------------------------------------- per dof: nle, F = Yc S ; frames ; centroidal totals
        for (int it = tid; it < nv + nfr + 1; it += T) {
            if (it < nv) {
                const int c = it, j = gF->dofj[c];
                V3 sl = ld3(sm.Sw + 6 * c), sa = ld3(sm.Sw + 6 * c + 3);
                sm.nle[c] = dot(sl, ld3(sm.fw + 6 * j)) + dot(sa, ld3(sm.fw + 6 * j + 3));
                const double* Yj = sm.Yw + 10 * j;
                V3 h = ld3(Yj + 1);
                V3 Fl = Yj[0] * sl - cross(h, sa);
                V3 Fa = mk(Yj[4] * sa.x + Yj[5] * sa.y + Yj[6] * sa.z, Yj[5] * sa.x + Yj[7] * sa.y + Yj[8] * sa.z, Yj[6] * sa.x + Yj[8] * sa.y + Yj[9] * sa.z)
                    + cross(h, sl);
                st3(sm.Fc + 6 * c, Fl);
                st3(sm.Fc + 6 * c + 3, Fa);
            }
            else if (it < nv + nfr) {
                const int f = it - nv;
                const int pj = gP->fparent[f];
                const double* fp = gP->fplace[f];
                double o[12];
                V3 vl = mk(0, 0, 0), w = mk(0, 0, 0), al = mk(0, 0, 0), aa = mk(0, 0, 0);
                if (pj >= 0) {
                    const double* pm = sm.oMi + 12 * pj;
                    mm3(pm, fp, o);
                    st3(o + 9, rmul(pm, ld3(fp + 9)) + ld3(pm + 9));
                    vl = ld3(sm.vw + 6 * pj);
                    w = ld3(sm.vw + 6 * pj + 3);
                    al = ld3(sm.aw + 6 * pj);
                    aa = ld3(sm.aw + 6 * pj + 3);
                }
                else {
                    for (int k = 0; k < 12; ++k)
                        o[k] = fp[k];
                }
                V3 p = ld3(o + 9);
                V3 wl = rtmul(o, w), vll = rtmul(o, vl - cross(p, w));
                V3 aal = rtmul(o, aa), all = rtmul(o, al - cross(p, aa)) + cross(wl, vll);
                // (vJ aliases [oMf, vf]: it is dead since the previous barrier)
                double* of = sm.oMf + 12 * f;
                for (int k = 0; k < 12; ++k)
                    of[k] = o[k];
                st3(sm.vf + 6 * f, vll);
                st3(sm.vf + 6 * f + 3, wl);
                st3(sm.af + 6 * f, all);
                st3(sm.af + 6 * f + 3, aal);
            }
            else {
                // totals = subtree sums of the roots; sc: [0..9] mass h I, [10..15] momentum, [16..21] force,
                // [22..24] com, [25..27] vcom, [28..30] acom drift, [31..36] hg, [37..42] dhg drift
                double tot[16];
                for (int k = 0; k < 16; ++k)
                    tot[k] = 0;
                for (int j = 0; j < nj; ++j)
                    if (gP->depth[j] == 0) {
                        for (int k = 0; k < 10; ++k)
                            tot[k] += sm.Yw[10 * j + k];
                        for (int k = 0; k < 6; ++k)
                            tot[10 + k] += sm.fw[6 * j + k];
                    }
                for (int k = 0; k < 10; ++k)
                    sc[k] = tot[k];
                for (int k = 0; k < 6; ++k)
                    sc[16 + k] = tot[10 + k];
                const double mt = tot[0];
                V3 hh = mk(tot[1], tot[2], tot[3]);
                V3 com = (1.0 / mt) * hh;
                V3 hl = ld3(sc + 10), ha = ld3(sc + 13);
                V3 fl = mk(tot[10], tot[11], tot[12]), fa = mk(tot[13], tot[14], tot[15]);
                V3 ag = -1.0 * ld3(gP->gravity);
                V3 dl = fl - mt * ag, da = fa - cross(hh, ag);
                st3(sc + 22, com);
                st3(sc + 25, (1.0 / mt) * hl);
                st3(sc + 28, (1.0 / mt) * dl);
                st3(sc + 31, hl);
                st3(sc + 34, ha - cross(com, hl));
                st3(sc + 37, dl);
                st3(sc + 40, da - cross(com, dl));
            }
        }
        fsync<T>();
        if (cyc && tid == 0)
            cyc[1] = clock64();

        // ---------------------------------------------------------------- scalar jobs: right-hand sides of every task, self-collision pairs
        for (int jb = tid; jb < F.njobs; jb += T) {
            const FJob J = gF->jobs[jb];
            if (J.kind == FJ_SE3) {
                const iwbc_task_t& Tk = gP->tasks[J.task];
                se3_rhs(sm, Tk.frame, Tk.kp, Tk.kd, Z + Tk.ref_off, rhs + 6 * J.slot);
            }
            else if (J.kind == FJ_CONTACT) {
                const iwbc_contact_t& C = gP->contacts[J.task];
                se3_rhs(sm, C.frame, C.kp, C.kd, Z + C.ref_off, sm.cb + 6 * J.task);
            }
            else if (J.kind == FJ_COM) {
                const iwbc_task_t& Tk = gP->tasks[J.task];
                const double* r = Z + Tk.ref_off;
                for (int i = 0; i < 3; ++i) {
                    const double a_des = -Tk.kp[i] * (sc[22 + i] - r[i]) - Tk.kd[i] * (sc[25 + i] - r[3 + i]) + r[6 + i];
                    rhs[6 * J.slot + i] = a_des - sc[28 + i];
                }
            }
            else if (J.kind == FJ_MOM) {
                const iwbc_task_t& Tk = gP->tasks[J.task];
                const double* r = Z + Tk.ref_off;
                for (int i = 0; i < 6; ++i) {
                    const double dL = -Tk.kp[i] * (sc[31 + i] - r[6 + i]) + r[12 + i];
                    rhs[6 * J.slot + i] = dL - sc[37 + i];
                }
            }
            else {
                const iwbc_task_t& Tk = gP->tasks[J.task];
                const int f0 = Tk.frame, fa = Tk.av_frame[J.a];
                const int lwa = !gP->sc_jacobian_world;
                const double* o0 = sm.oMf + 12 * f0;
                const double* oa = sm.oMf + 12 * fa;
                V3 diff = ld3(o0 + 9) - ld3(oa + 9);
                const double sq = dot(diff, diff), nrm = sqrt(sq);
                const double aa = Tk.av_r0[J.a] + Tk.radius;
                V3 drift = ld3(sm.af + 6 * f0) - ld3(sm.af + 6 * fa);
                auto point_vel = [&](int f, const double* o) {
                    int pj = gP->fparent[f];
                    if (pj < 0)
                        return mk(0, 0, 0);
                    V3 vl = ld3(sm.vw + 6 * pj), w = ld3(sm.vw + 6 * pj + 3);
                    return lwa ? vl + cross(w, ld3(o + 9)) : vl;
                };
                V3 Jv = point_vel(f0, o0) - point_vel(fa, oa);
                const double m = Tk.sc_m, k = Tk.sc_k;
                const double x = k * (nrm - aa + Tk.sc_sp);
                const double e_p = exp(-x);
                const double l1 = log1p(e_p);
                const double pw0 = exp(-m * l1);
                const double pw1 = exp((-m - 1) * l1);
                const double pw2 = exp((-m - 2) * l1);
                const double Cv = 1. - pw0;
                const double gs = -1. / nrm * k * m * e_p * pw1;
                const double h = 1. / sq * k * k * (-m - 1.) * m * exp(-2 * x) * pw2 + 1. / sq * k * k * m * e_p * pw1
                    + 1. / (nrm * sqrt(nrm)) * k * m * e_p * pw1;
                V3 grad = gs * diff;
                V3 HJv = (h * dot(diff, Jv)) * diff + gs * Jv;
                const double Bc = -(dot(HJv, Jv) + dot(grad, Tk.kd[0] * Jv - drift) + Tk.kp[0] * Cv);
                st3(sc4 + 4 * J.slot, grad);
                sc4[4 * J.slot + 3] = Bc;
            }
        }
        fsync<T>();
        // ---------------------------------------------------------------- rows: A (dense level-1 rows, scaled by sqrt(w)), b, contact Jacobians;
        // H = A'A in 4 x 4 register tiles (lower triangle), g = -A'b - posture.  The rows are formed and folded in F.nchunk chunks of
        // whole row blocks (the accumulators stay in registers across the chunks; summation order = row order, as before).
        double hacc[FT<T>::NTPT][16];
#pragma unroll 1
        for (int tt = 0; tt < FT<T>::NTPT; ++tt)
#pragma unroll
            for (int e = 0; e < 16; ++e)
                hacc[tt][e] = 0;
        for (int k = (T - 1 - tid); k < nv; k += T)
            gvec[k] = 0.0;
        // (this thread's tile: read from the plan once, not once per chunk)
        const int tile0_i = (T > 1 && tid < F.ntiles) ? 4 * gF->tile_i[tid] : 0, tile0_j = (T > 1 && tid < F.ntiles) ? 4 * gF->tile_j[tid] : 0;
#pragma unroll 1
        for (int ch = 0; ch < F.nchunk; ++ch) {
            const int b0 = F.ch_b0[ch], b1 = F.ch_b1[ch], r0 = F.ch_r0[ch], R = F.ch_nr[ch];
            const int NH = (T / nv) > 0 ? (T / nv) : 1;
            for (int item = tid; item < nv * NH; item += T) {
                const int hh = item / nv, k = item - hh * nv;
                const int jk = gF->dofj[k], jk_last = gP->last_desc[jk]; // joint jk supports joint j iff jk <= j <= jk_last
                const V3 sl = ld3(sm.Sw + 6 * k), sa = ld3(sm.Sw + 6 * k + 3);
                // local-frame Jacobian column of a frame whose parent joint is pj (frame_jac_col with the chain of lookups folded)
                auto jac_col = [&](int f, int pj, V3& lin, V3& ang) {
                    lin = mk(0, 0, 0);
                    ang = mk(0, 0, 0);
                    if (pj < 0 || pj < jk || pj > jk_last)
                        return;
                    const double* o = sm.oMf + 12 * f;
                    ang = rtmul(o, sa);
                    lin = rtmul(o, sl - cross(ld3(o + 9), sa));
                };
                for (int bi = b0 + hh; bi < b1; bi += NH) {
                    const FBlock Bk = gF->blocks[bi];
                    if (Bk.kind == FB_CONTACT) {
                        const int c = Bk.index;
                        V3 l, a;
                        jac_col(Bk.frame, Bk.pj, l, a);
                        double* Jc = sm.Jft + (c * 6) * nv + k;
                        Jc[0] = l.x, Jc[nv] = l.y, Jc[2 * nv] = l.z, Jc[3 * nv] = a.x, Jc[4 * nv] = a.y, Jc[5 * nv] = a.z;
                        continue;
                    }
                    const double sw = gF->blk_sw[bi];
                    double* const Ab = A + (Bk.row0 - r0) * nvp + k;
                    if (Bk.kind == FB_SE3) {
                        V3 l, a;
                        jac_col(Bk.frame, Bk.pj, l, a);
                        const double col[6] = {l.x, l.y, l.z, a.x, a.y, a.z};
                        int idx = 0;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
                            if (Bk.maskbits >> i & 1)
                                Ab[(idx++) * nvp] = sw * col[i];
                    }
                    else if (Bk.kind == FB_COM) {
                        const double imt = 1.0 / sc[0];
                        int idx = 0;
                        for (int i = 0; i < 3; ++i)
                            if (Bk.maskbits >> i & 1)
                                Ab[(idx++) * nvp] = sw * imt * sm.Fc[6 * k + i];
                    }
                    else if (Bk.kind == FB_MOM) {
                        const V3 com = ld3(sc + 22);
                        V3 Fl = ld3(sm.Fc + 6 * k), Fa = ld3(sm.Fc + 6 * k + 3) - cross(com, Fl);
                        const double col[6] = {Fl.x, Fl.y, Fl.z, Fa.x, Fa.y, Fa.z};
                        int idx = 0;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
                            if (Bk.maskbits >> i & 1)
                                Ab[(idx++) * nvp] = sw * col[i];
                    }
                    else { // FB_SELFCOL
                        const int lwa = !gP->sc_jacobian_world;
                        auto col = [&](int f, int pj) {
                            if (pj < 0 || pj < jk || pj > jk_last)
                                return mk(0, 0, 0);
                            return lwa ? sl + cross(sa, ld3(sm.oMf + 12 * f + 9)) : sl;
                        };
                        const V3 c0 = col(Bk.frame, Bk.pj);
                        double acc = 0;
                        for (int a = 0; a < Bk.nav; ++a)
                            acc += dot(ld3(sc4 + 4 * (Bk.slot0 + a)), c0 - col(gF->pair_frame[Bk.slot0 + a], gF->pair_pj[Bk.slot0 + a]));
                        Ab[0] = sw * acc;
                    }
                }
            }
            for (int r = tid; r < R; r += T) {
                const FRow Rw = gF->rows[r0 + r];
                const iwbc_task_t& Tk = gP->tasks[Rw.task];
#if IWBC_OPT_BF
                const double sw = gF->row_sw[r0 + r];
#else
                const double sw = sqrt(Tk.weight);
#endif
                double val;
                if (Rw.kind == FB_SELFCOL) {
                    val = 0;
                    for (int a = 0; a < Tk.nav; ++a)
                        val += sc4[4 * (Rw.slot + a) + 3];
                }
                else
                    val = rhs[6 * Rw.slot + Rw.comp];
                bvec[r] = sw * val;
                for (int k = nv; k < nvp; ++k) // zero padding of the row
                    A[r * nvp + k] = 0.0;
            }
            fsync<T>();
#pragma unroll 1
            for (int tt = 0; tt < FT<T>::NTPT; ++tt) {
                const int tile = tt * T + tid;
                double* h = hacc[tt];
                if (tile < F.ntiles) {
                    const double* Ai = A + (T > 1 ? tile0_i : 4 * gF->tile_i[tile]);
                    const double* Aj = A + (T > 1 ? tile0_j : 4 * gF->tile_j[tile]);
                    for (int r = 0; r < R; ++r) {
#if IWBC_OPT_CHOL
                        // (tile columns are multiples of 4 and nvp % 4 == 0: 16-byte loads)
                        const double2 a01 = *reinterpret_cast<const double2*>(Ai + r * nvp), a23 = *reinterpret_cast<const double2*>(Ai + r * nvp + 2);
                        const double2 b01 = *reinterpret_cast<const double2*>(Aj + r * nvp), b23 = *reinterpret_cast<const double2*>(Aj + r * nvp + 2);
                        const double a0 = a01.x, a1 = a01.y, a2 = a23.x, a3 = a23.y, b0 = b01.x, b1 = b01.y, b2 = b23.x, b3 = b23.y;
#else
                        const double a0 = Ai[r * nvp], a1 = Ai[r * nvp + 1], a2 = Ai[r * nvp + 2], a3 = Ai[r * nvp + 3];
                        const double b0 = Aj[r * nvp], b1 = Aj[r * nvp + 1], b2 = Aj[r * nvp + 2], b3 = Aj[r * nvp + 3];
#endif
                        h[0] += a0 * b0, h[1] += a0 * b1, h[2] += a0 * b2, h[3] += a0 * b3;
                        h[4] += a1 * b0, h[5] += a1 * b1, h[6] += a1 * b2, h[7] += a1 * b3;
                        h[8] += a2 * b0, h[9] += a2 * b1, h[10] += a2 * b2, h[11] += a2 * b3;
                        h[12] += a3 * b0, h[13] += a3 * b1, h[14] += a3 * b2, h[15] += a3 * b3;
                    }
                }
            }
            // g: threads from the top end (the tile owners sit at the bottom end); thread-owned accumulator in shared memory
            for (int k = (T - 1 - tid); k < nv; k += T) {
                double s = 0;
                for (int r = 0; r < R; ++r)
                    s += A[r * nvp + k] * bvec[r];
                gvec[k] -= s;
            }
            fsync<T>(); // this chunk of A is dead
        }
        for (int k = (T - 1 - tid); k < nv; k += T) {
            if (k >= nu) {
                double s = gvec[k];
                const int i = k - nu;
                for (int pt = 0; pt < F.nposture; ++pt) {
                    const iwbc_task_t& Tk = gP->tasks[F.posture[pt]];
                    if (!Tk.mask[i])
                        continue;
                    const double* r = Z + Tk.ref_off;
                    const double a_des = -Tk.kp[0] * (qs[nq - na + i] - r[i]) - Tk.kd[0] * (vs[nv - na + i] - r[na + i]) + r[2 * na + i];
                    s -= Tk.weight * a_des;
                }
                gvec[k] = s;
            }
        }
        if (cyc && tid == 0)
            cyc[2] = clock64();
        // ---------------------------------------------------------------- mass matrix M[r][c] = S_lo . F_hi when joint(lo) supports joint(hi): one half,
        // held in registers until the dynamics scratch it is computed from is dead (its home is the upper triangle of the augmented
        // matrix, which overlays that scratch)
        constexpr int MNV = NV_ > 0 ? NV_ : 60; // (the tile table bounds nv to 60)
        constexpr int MPT = (MNV * (MNV + 1) / 2 + T - 1) / T;
        double mreg[MPT];
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            const int e = m * T + tid;
            mreg[m] = 0.0;
            if (e < nv * (nv + 1) / 2) {
#if IWBC_OPT_MTAB
                const int r = gF->tri_r[e], c = gF->tri_c[e];
#else
                int r = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
                while (r * (r + 1) / 2 > e)
                    --r;
                while ((r + 1) * (r + 2) / 2 <= e)
                    ++r;
                const int c = e - r * (r + 1) / 2;
#endif
#if IWBC_OPT_MTAB
                if (gF->tri_sup[e])
#else
                if (supports(gP, gF->dofj[c], gF->dofj[r]))
#endif
                    mreg[m] = dot(ld3(sm.Sw + 6 * c), ld3(sm.Fc + 6 * r)) + dot(ld3(sm.Sw + 6 * c + 3), ld3(sm.Fc + 6 * r + 3));
            }
        }
        fsync<T>(); // the dynamics scratch is dead
        // ---------------------------------------------------------------- augmented matrix: [H ; N' ; g'] (column-major, ld = ldh)
        {
#pragma unroll 1
            for (int tt = 0; tt < FT<T>::NTPT; ++tt) {
                const int tile = tt * T + tid;
                if (tile < F.ntiles) {
                    const int i0 = 4 * gF->tile_i[tile], j0 = 4 * gF->tile_j[tile];
                    const double* h = hacc[tt];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int i = i0 + a, j = j0 + b;
                            if (i < nv && j <= i) {
                                double val = h[4 * a + b];
                                if (i == j) {
                                    val += gF->hdiag[i];
                                    hd[i] = val;
                                }
                                Hs[i + j * ld] = val;
                            }
                        }
                }
            }
            // the mass matrix: strict lower triangle into the (unused) upper triangle of the augmented matrix -- M(r, c), r > c, at
            // Hs[c + r ld] --, its diagonal into Md; its first nu rows are the base-dynamics normals (rows nv+cb0 .. nv+me-1)
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                const int e = m * T + tid;
                if (e < nv * (nv + 1) / 2) {
#if IWBC_OPT_MTAB
                    const int r = gF->tri_r[e], c = gF->tri_c[e];
#else
                    int r = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
                    while (r * (r + 1) / 2 > e)
                        --r;
                    while ((r + 1) * (r + 2) / 2 <= e)
                        ++r;
                    const int c = e - r * (r + 1) / 2;
#endif
                    const double val = mreg[m];
                    if (r == c)
                        Md[r] = val;
                    else
                        Hs[c + r * ld] = val;
                    if (c < nu) {
                        Hs[(nv + F.cb0 + c) + r * ld] = val;
                        if (r < nu && r != c)
                            Hs[(nv + F.cb0 + r) + c * ld] = val;
                    }
                }
            }
            // constraint normals (dv part) of the contacts as rows nv .. nv+cb0-1, the gradient as row nv+me
            for (int e = tid; e < (F.cb0 + 1) * nv; e += T) {
                const int c0 = e / nv, k = e - c0 * nv;
                const int c = c0 < F.cb0 ? c0 : me;
                double val;
                if (c < F.cb0)
                    val = sm.Jft[(F.cact[c / 6] * 6 + (c % 6)) * nv + k];
                else
                    val = gvec[k];
                Hs[(nv + c) + k * ld] = val;
            }
            // force part of the base-dynamics normals and of the gradient, through the constant force-block factor:
            // Bf[kk][r] = (Uff' N_f)[kk][r],  N_f[12 s + i][r] = -(T' Jft)[i][r]
            for (int e = tid; e < 12 * F.nca * F.nf; e += T) {
                const int kk = e / F.nf, r = e - kk * F.nf;
                const int s = kk / 12, k2 = kk - 12 * s;
                double val;
                const int wvs = (int)frv[8 * s + 6];
                if (r == nu) { // (L_ff^-1 g_f)[k2] for this instance's force reference
                    const double* G6 = gF->Gf[wvs][s] + k2 * 6;
                    val = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        val += G6[a] * frv[8 * s + a];
                }
                else {
                    const int c = F.cact[s];
#if IWBC_OPT_BF
                    // (U_ff' T' is a constant of the problem: six products per entry instead of up to 84)
                    const double* W6 = gF->UT[wvs][s] + k2 * 6;
                    val = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        val -= W6[a] * sm.Jft[(c * 6 + a) * nv + r];
#else
                    const double* Tm = gP->contacts[c].T;
                    const double* U = Usel(s);
                    val = 0;
                    for (int i = 0; i <= k2; ++i) {
                        double npf = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a)
                            npf += Tm[a * 12 + i] * sm.Jft[(c * 6 + a) * nv + r];
                        val -= U[i * 12 + k2] * npf;
                    }
#endif
                }
                Bf[kk * F.nf + r] = val;
            }
        }
        fsync<T>();

        // ---------------------------------------------------------------- blocked left-looking Cholesky of the augmented matrix (panels of 4 columns, 2 barriers each)
        bool spd = true;
        {
            constexpr int PAIR = FT<T>::PAIR;
            const int sub = tid % PAIR, row0 = tid / PAIR, rstep = (T / PAIR) > 0 ? (T / PAIR) : 1;
            const int nrh = F.nrh;
            double* const Db = Pp; // copy of the updated diagonal block (16 doubles)
            for (int cb = 0; cb < nv; cb += 4) {
                const int pw = nv - cb < 4 ? nv - cb : 4;
                // (1) panel update: rows i >= cb of columns cb .. cb+3 minus the contribution of all previous columns
                const double* Lp = Hs + cb;
                for (int ib = cb; ib < nrh; ib += rstep) { // (uniform trip count: the pair shuffles need the whole warp)
                    const int i = ib + row0;
                    const bool act = i < nrh;
                    const double* Li = Hs + (act ? i : cb);
                    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                    for (int j = sub; j < cb; j += PAIR) {
                        const double li = Li[j * ld];
                        const double2 p01 = *reinterpret_cast<const double2*>(Lp + j * ld); // 16-byte aligned: cb % 4 == 0, ld even
                        const double2 p23 = *reinterpret_cast<const double2*>(Lp + j * ld + 2);
                        a0 += li * p01.x, a1 += li * p01.y, a2 += li * p23.x, a3 += li * p23.y;
                    }
                    // (measured and dropped: two columns per trip into two accumulator sets -- neutral)
                    if (PAIR == 2) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
                        a3 += __shfl_xor_sync(0xffffffffu, a3, 1);
                    }
                    if (act && sub == 0) {
                        double* Ri = Hs + i + cb * ld;
                        const double v0 = Ri[0] - a0, v1 = pw > 1 ? Ri[ld] - a1 : 0.0, v2 = pw > 2 ? Ri[2 * ld] - a2 : 0.0, v3 = pw > 3 ? Ri[3 * ld] - a3 : 0.0;
                        Ri[0] = v0; // (lower triangle only: the upper triangle of the dv block holds the mass matrix)
                        if (pw > 1 && i > cb)
                            Ri[ld] = v1;
                        if (pw > 2 && i > cb + 1)
                            Ri[2 * ld] = v2;
                        if (pw > 3 && i > cb + 2)
                            Ri[3 * ld] = v3;
                        if (i < cb + 4) {
                            double* dr = Db + 4 * (i - cb);
                            dr[0] = v0, dr[1] = v1, dr[2] = v2, dr[3] = v3;
                        }
                    }
                }
                // (2) factor the diagonal block (every thread, redundantly, from the copy) and solve the panel rows.  (Measured and
                // dropped: ONE thread factorising it before the barrier and publishing the ten numbers -- 5 % fewer instructions, but
                // the factorisation then sits on warp 0's path to the barrier instead of overlapping across the warps: 4.89 -> 4.79 M/s.)
                auto factor4 = [&](double& l10, double& l20, double& l30, double& l21, double& l31, double& l32, double& i0, double& i1, double& i2,
                                   double& i3) {
                    const double d00 = Db[0], d10 = Db[4], d20 = pw > 2 ? Db[8] : 0.0, d30 = pw > 3 ? Db[12] : 0.0;
                    const double d11 = pw > 1 ? Db[5] : 1.0, d21 = pw > 2 ? Db[9] : 0.0, d31 = pw > 3 ? Db[13] : 0.0;
                    const double d22 = pw > 2 ? Db[10] : 1.0, d32 = pw > 3 ? Db[14] : 0.0, d33 = pw > 3 ? Db[15] : 1.0;
                    const double d10u = pw > 1 ? d10 : 0.0;
                    bool ok = d00 > 0.0;
                    i0 = rsqrt(d00 > 0.0 ? d00 : 1.0);
                    l10 = d10u * i0, l20 = d20 * i0, l30 = d30 * i0;
                    const double t1 = d11 - l10 * l10;
                    ok = ok && t1 > 0.0;
                    i1 = rsqrt(t1 > 0.0 ? t1 : 1.0);
                    l21 = (d21 - l20 * l10) * i1, l31 = (d31 - l30 * l10) * i1;
                    const double t2 = d22 - l20 * l20 - l21 * l21;
                    ok = ok && t2 > 0.0;
                    i2 = rsqrt(t2 > 0.0 ? t2 : 1.0);
                    l32 = (d32 - l30 * l20 - l31 * l21) * i2;
                    const double t3 = d33 - l30 * l30 - l31 * l31 - l32 * l32;
                    ok = ok && t3 > 0.0;
                    i3 = rsqrt(t3 > 0.0 ? t3 : 1.0);
                    return ok;
                };
                fsync<T>();
                {
                    double l10, l20, l30, l21, l31, l32, i0, i1, i2, i3;
                    if (!factor4(l10, l20, l30, l21, l31, l32, i0, i1, i2, i3))
                        spd = false;
                    if (tid == 0) {
                        il[cb] = i0;
                        if (pw > 1)
                            il[cb + 1] = i1;
                        if (pw > 2)
                            il[cb + 2] = i2;
                        if (pw > 3)
                            il[cb + 3] = i3;
                    }
                    for (int i = cb + tid; i < nrh; i += T) {
                        double* Ri = Hs + i + cb * ld;
                        const int r = i - cb;
                        const double x0 = Ri[0] * i0;
                        Ri[0] = x0;
                        if (pw > 1 && r >= 1) {
                            const double x1 = (Ri[ld] - x0 * l10) * i1;
                            Ri[ld] = x1;
                            if (pw > 2 && r >= 2) {
                                const double x2 = (Ri[2 * ld] - x0 * l20 - x1 * l21) * i2;
                                Ri[2 * ld] = x2;
                                if (pw > 3 && r >= 3)
                                    Ri[3 * ld] = (Ri[3 * ld] - x0 * l30 - x1 * l31 - x2 * l32) * i3;
                            }
                        }
                    }
                }
                fsync<T>();
            }
        }
        if (cyc && tid == 0)
            cyc[3] = clock64();

        // ---------------------------------------------------------------- Householder QR of B = [L^-1 N_v ; B_f] (n x me), right-hand side y carried as column me
        // B(j, c): j < nv -> Hs[(nv + c) + j ld] ; j >= nv -> c >= cb0 ? Bf[(j - nv) nf + (c - cb0)] : 0.  Column k keeps its
        // sub-diagonal part unscaled (v_k = e_k + scale_k B(k+1:, k)).  Threads whose column is already factorised (c < k)
        // accumulate the Gram entries v_c'v_k needed to apply Q afterwards.
        const int nfrc = 2 * (me - 6), nfc = 7, cb0 = me - 6; // = 12 F.nca, F.nf, F.cb0
        constexpr int SEG = FT<T>::SEG;
        const int qc0 = (T > 1) ? tid % NC : 0, seg0 = (T > 1) ? tid / NC : 0;
        const int cstep = (T > 1) ? NC : 1;
        if constexpr (T > 1 && NV_ > 0 && ME_ > 0) {
            // Register-resident variant: thread (seg0, c) owns rows j = seg0 + SEG m of column c for the whole factorisation;
            // only the pivot column and the pivot row are published through shared memory (wv / gvec are free here).
            constexpr int MR = (NV_ + SEG - 1) / SEG, MF = (2 * (ME_ - 6) + SEG - 1) / SEG;
            double* const colbuf = wv;   // B(:, k); IWBC_OPT_RELOAD: double-buffered with xv (free until the solve), so that the update can
                                         // read the pivot column again instead of holding it in 26 registers across the barrier
            double* const rowbuf = gvec; // B(k, :), double-buffered: [2][NC]
            // |B(k+1:, k)|^2 of the NEXT pivot column, per row segment, double-buffered: its owners compute it right after they have
            // updated the column, so that the sigma -> rsqrt -> reciprocal chain of a step starts at the top of the step and overlaps
            // the dot products and the first barrier instead of following them (tv is free until the QR is done)
            double* const Sg = tv;
            const bool qact = seg0 < SEG;
            const int c = qc0;
            double cc[MR], cf[MF], ck[MR], ckf[MF];
            if (qact) {
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int j = seg0 + SEG * m;
                    cc[m] = j < nv ? Hs[(nv + c) + j * ld] : 0.0;
                }
#pragma unroll
                for (int m = 0; m < MF; ++m) {
                    const int jj = seg0 + SEG * m;
                    cf[m] = (c >= cb0 && jj < nfrc) ? Bf[jj * nfc + (c - cb0)] : 0.0;
                }
                if (c == 0) {
#pragma unroll
                    for (int m = 0; m < MR; ++m)
                        if (seg0 + SEG * m < nv)
                            colbuf[seg0 + SEG * m] = cc[m];
#if IWBC_OPT_SIGMA
                    double sg0 = 0;
#pragma unroll
                    for (int m = 0; m < MR; ++m)
                        if (seg0 + SEG * m > 0 && seg0 + SEG * m < nv)
                            sg0 += cc[m] * cc[m];
                    Sg[seg0] = sg0; // (column 0 has no force rows)
#endif
                }
                if (seg0 == 0)
                    rowbuf[c] = cc[0];
            }
            fsync<T>();
            // (k is a compile-time constant after unrolling: rows j = seg0 + SEG m <= k of every segment are statically dead -- on
            // average half of the dot / update work --, and the next pivot row sits in a statically known register)
#if IWBC_OPT_QR
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int k = 0; k < ME_; ++k) {
                const bool frc = k >= cb0;
                const double* rb = rowbuf + (k & 1) * NC;
                double* rbn = rowbuf + ((k + 1) & 1) * NC;
                const double* const cbk = (IWBC_OPT_RELOAD && (k & 1)) ? xv : colbuf;  // this step's pivot column
                double* const cbn = (IWBC_OPT_RELOAD && !(k & 1)) ? xv : colbuf;       // where the next one is published
                if (qact) {
                    double sacc = 0, sacc1 = 0, sacc2 = 0; // (three chains: a single accumulator is 13 dependent FMAs per step)
#pragma unroll
                    for (int m = 0; m < MR; ++m) {
                        if (SEG * (m + 1) - 1 <= k) // dead for every segment
                            continue;
                        const int j = seg0 + SEG * m;
                        ck[m] = (j > k && j < nv) ? cbk[j] : 0.0;
                        if (m & 1)
                            sacc1 += ck[m] * cc[m];
                        else
                            sacc += ck[m] * cc[m];
                    }
                    if (frc) {
#pragma unroll
                        for (int m = 0; m < MF; ++m) {
                            const int jj = seg0 + SEG * m;
                            ckf[m] = jj < nfrc ? cbk[nv + jj] : 0.0;
                            sacc2 += ckf[m] * cf[m];
                        }
                    }
                    Pp[seg0 * NC + c] = (sacc + sacc1) + sacc2;
                }
#if !IWBC_OPT_SIGMA
                fsync<T>();
#endif
                double sigma = 0;
#pragma unroll
                for (int sg = 0; sg < SEG; ++sg)
                    sigma += IWBC_OPT_SIGMA ? Sg[(k & 1) * SEG + sg] : Pp[sg * NC + k];
                const double alpha = rb[k];
                double beta = alpha, tau = 0, scale = 0, ibeta = 0;
                if (sigma != 0.0) {
                    const double nn = alpha * alpha + sigma;
                    const double inr = rsqrt(nn), nr = nn * inr;
                    beta = alpha >= 0 ? -nr : nr;
                    ibeta = alpha >= 0 ? -inr : inr;
                    tau = 1.0 + fabs(alpha) * inr; // (beta - alpha) / beta
                    scale = 1.0 / (alpha - beta);
                }
#if IWBC_OPT_SIGMA
                fsync<T>();
#endif
                if (qact) {
                    if (c == k) {
                        if (seg0 == 0) {
                            Rm[k * NC + k] = beta;
                            tauv[k] = tau;
                            scalev[k] = scale;
                            invb[k] = sigma != 0.0 ? ibeta : 1.0 / alpha; // (the division only where it is needed: one thread, off the others' path)
                        }
                    }
                    else {
                        double pc = 0;
#pragma unroll
                        for (int sg = 0; sg < SEG; ++sg)
                            pc += Pp[sg * NC + c];
                        const double bkc = rb[c];
                        if (c < k) {
                            if (seg0 == 0)
                                Gm[c * (me + 1) + k] = scalev[c] * (bkc + scale * pc);
                        }
                        else {
                            const double coef = tau * (bkc + scale * pc);
                            if (seg0 == 0)
                                Rm[k * NC + c] = bkc - coef;
                            const double cs = coef * scale;
#if IWBC_OPT_RELOAD
#pragma unroll
                            for (int m = 0; m < MR; ++m)
                                if (SEG * (m + 1) - 1 > k) {
                                    const int j = seg0 + SEG * m;
                                    cc[m] -= cs * ((j > k && j < nv) ? cbk[j] : 0.0);
                                }
                            if (frc) {
#pragma unroll
                                for (int m = 0; m < MF; ++m)
                                    cf[m] -= cs * (seg0 + SEG * m < nfrc ? cbk[nv + seg0 + SEG * m] : 0.0);
                            }
#else
#pragma unroll
                            for (int m = 0; m < MR; ++m)
                                if (SEG * (m + 1) - 1 > k)
                                    cc[m] -= cs * ck[m];
                            if (frc) {
#pragma unroll
                                for (int m = 0; m < MF; ++m)
                                    cf[m] -= cs * ckf[m];
                            }
#endif
                        }
                    }
                    // publish the next pivot column and pivot row
                    if (c == k + 1) {
#pragma unroll
                        for (int m = 0; m < MR; ++m)
                            if (SEG * (m + 1) - 1 > k && seg0 + SEG * m < nv) // (rows <= k + 1 of the next column are never read)
                                cbn[seg0 + SEG * m] = cc[m];
                        if (k + 1 >= cb0) {
#pragma unroll
                            for (int m = 0; m < MF; ++m)
                                if (seg0 + SEG * m < nfrc)
                                    cbn[nv + seg0 + SEG * m] = cf[m];
                        }
#if IWBC_OPT_SIGMA
                        double sgn = 0, sgn1 = 0, sgn2 = 0; // (three chains, like the dot products)
#pragma unroll
                        for (int m = 0; m < MR; ++m)
                            if (SEG * (m + 1) - 1 > k + 1 && seg0 + SEG * m > k + 1 && seg0 + SEG * m < nv) {
                                if (m & 1)
                                    sgn1 += cc[m] * cc[m];
                                else
                                    sgn += cc[m] * cc[m];
                            }
                        if (k + 1 >= cb0) {
#pragma unroll
                            for (int m = 0; m < MF; ++m)
                                if (seg0 + SEG * m < nfrc)
                                    sgn2 += cf[m] * cf[m];
                        }
                        Sg[((k + 1) & 1) * SEG + seg0] = (sgn + sgn1) + sgn2;
#endif
                    }
                    {
                        const int m1 = (k + 1) / SEG; // the row k + 1 lives in segment (k + 1) % SEG, register m1
                        if (m1 < MR && seg0 == (k + 1) - SEG * m1)
                            rbn[c] = cc[m1];
                    }
                }
                fsync<T>();
            }
            if (qact) {
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int j = seg0 + SEG * m;
                    if (j < nv)
                        Hs[(nv + c) + j * ld] = cc[m];
                }
                if (c >= cb0) {
#pragma unroll
                    for (int m = 0; m < MF; ++m) {
                        const int jj = seg0 + SEG * m;
                        if (jj < nfrc)
                            Bf[jj * nfc + (c - cb0)] = cf[m];
                    }
                }
            }
            fsync<T>();
        }
        else
        {
            for (int k = 0; k < me; ++k) {
                const double* colk = Hs + (nv + k);
                const bool frc = k >= cb0; // column k has force rows (contact columns have none and nothing fills them in)
                if (seg0 < SEG)
                    for (int c = qc0; c < NC; c += cstep) {
                        const double* colc = Hs + (nv + c);
                        double s = 0;
                        for (int j = k + 1 + seg0; j < nv; j += SEG)
                            s += colk[j * ld] * colc[j * ld];
                        if (frc && c >= cb0) {
                            const double* fk = Bf + (k - cb0);
                            const double* fc = Bf + (c - cb0);
                            for (int jj = seg0; jj < nfrc; jj += SEG)
                                s += fk[jj * nfc] * fc[jj * nfc];
                        }
                        Pp[seg0 * NC + c] = s;
                    }
                fsync<T>();
                double sigma = 0;
                for (int sg = 0; sg < SEG; ++sg)
                    sigma += Pp[sg * NC + k];
                const double alpha = colk[k * ld]; // k < me <= nv
                double beta = alpha, tau = 0, scale = 0, ibeta = 1.0 / alpha;
                if (sigma != 0.0) {
                    const double nn = alpha * alpha + sigma;
                    const double inr = rsqrt(nn), nr = nn * inr;
                    beta = alpha >= 0 ? -nr : nr;
                    ibeta = alpha >= 0 ? -inr : inr;
                    tau = 1.0 + fabs(alpha) * inr;          // (beta - alpha) / beta
                    scale = 1.0 / (alpha - beta);
                }
                if (seg0 < SEG)
                    for (int c = qc0; c < NC; c += cstep) {
                        if (c == k) {
                            if (seg0 == 0) {
                                Rm[k * NC + k] = beta;
                                tauv[k] = tau;
                                scalev[k] = scale;
                                invb[k] = ibeta;
                            }
                            continue;
                        }
                        double pc = 0;
                        for (int sg = 0; sg < SEG; ++sg)
                            pc += Pp[sg * NC + c];
                        double* colc = Hs + (nv + c);
                        const double bkc = colc[k * ld];
                        if (c < k) { // Gram entry v_c'v_k = scale_c (B(k, c) + scale_k sum_{j > k} B(j, c) B(j, k))
                            if (seg0 == 0)
                                Gm[c * (me + 1) + k] = scalev[c] * (bkc + scale * pc);
                            continue;
                        }
                        const double coef = tau * (bkc + scale * pc);
                        if (seg0 == 0)
                            Rm[k * NC + c] = bkc - coef;
                        const double cs = coef * scale;
                        for (int j = k + 1 + seg0; j < nv; j += SEG)
                            colc[j * ld] -= cs * colk[j * ld];
                        if (frc) {
                            const double* fk = Bf + (k - cb0);
                            double* fc = Bf + (c - cb0);
                            for (int jj = seg0; jj < nfrc; jj += SEG)
                                fc[jj * nfc] -= cs * fk[jj * nfc];
                        }
                    }
                fsync<T>();
            }
        }
        // ---------------------------------------------------------------- w = -Q t,  t = [R^-T ce0 ; y2]  via  Q = I - V T V',  T^-1 = striu(V'V) + diag(1/tau)
        {
            double* const zv = tv + (me + 2);
            double* const zs = tv + 2 * (me + 2) + 2;
            auto ce0_of = [&](int k) { return k < cb0 ? -sm.cb[6 * F.cact[k / 6] + (k % 6)] : sm.nle[k - cb0]; };
            // u = R^-T ce0: column-oriented forward substitution by the first warp; c1 = trace(H), c2 = trace(L^-T) by another thread
            if (T == 1) {
                for (int k = 0; k < me; ++k) {
                    double sacc = ce0_of(k);
                    for (int i = 0; i < k; ++i)
                        sacc -= Rm[i * NC + k] * tv[i];
                    tv[k] = sacc * invb[k];
                }
            }
            else if (tid < 32) {
                double rk = tid < me ? ce0_of(tid) : 0.0;
                for (int i = 0; i < me; ++i) {
                    const double ui = __shfl_sync(0xffffffffu, rk, i) * invb[i];
                    if (tid == i)
                        tv[i] = ui;
                    if (tid > i && tid < me)
                        rk -= Rm[i * NC + tid] * ui;
                }
            }
            if (tid == (T > 32 ? 32 : 0)) { // fixed summation order
                double c1 = 0, c2 = 0;
                for (int s = 0; s < F.nca; ++s) {
                    const int wvs = (int)frv[8 * s + 6];
                    c1 += gF->c1v[wvs][s];
                    c2 += gF->c2v[wvs][s];
                }
                for (int k = 0; k < nv; ++k) {
                    c1 += hd[k];
                    c2 += il[k];
                }
                tv[2 * (me + 2)] = c1;
                tv[2 * (me + 2) + 1] = c2;
            }
            fsync<T>();
            // v_i't = t_i + scale_i sum_{j > i} B(j, i) t_j  (partial sums per row segment)
            if (seg0 < SEG)
                for (int i = qc0; i < me; i += cstep) {
                    const double* coli = Hs + (nv + i);
                    const double* colt = Hs + (nv + me);
                    double sacc = 0;
                    for (int j = i + 1 + seg0; j < nv; j += SEG)
                        sacc += coli[j * ld] * (j < me ? tv[j] : colt[j * ld]);
                    if (i >= cb0) {
                        const double* fi = Bf + (i - cb0);
                        const double* ft = Bf + (me - cb0);
                        for (int jj = seg0; jj < nfrc; jj += SEG)
                            sacc += fi[jj * nfc] * ft[jj * nfc];
                    }
                    Pp[seg0 * NC + i] = sacc;
                }
            fsync<T>();
            // z = T (V't): back-substitution on T^-1 z = V't (upper triangular, diagonal 1/tau); zs = z * scale
            if (T == 1) {
                for (int i = me - 1; i >= 0; --i) {
                    double sacc = tv[i] + scalev[i] * Pp[i];
                    for (int k = i + 1; k < me; ++k)
                        sacc -= Gm[i * (me + 1) + k] * zv[k];
                    zv[i] = sacc * tauv[i];
                    zs[i] = zv[i] * scalev[i];
                }
            }
            else if (tid < 32) {
                double ri = 0;
                if (tid < me) {
                    double pacc = 0;
                    for (int sg = 0; sg < SEG; ++sg)
                        pacc += Pp[sg * NC + tid];
                    ri = tv[tid] + scalev[tid] * pacc;
                }
                for (int k = me - 1; k >= 0; --k) {
                    const double zk = __shfl_sync(0xffffffffu, ri, k) * tauv[k];
                    if (tid == k) {
                        zv[k] = zk;
                        zs[k] = zk * scalev[k];
                    }
                    if (tid < k)
                        ri -= Gm[tid * (me + 1) + k] * zk;
                }
            }
            fsync<T>();
            // w = -(t - V z)
            for (int j = tid; j < n; j += T) {
                double sacc;
                if (j < nv) {
                    const double* row = Hs + nv + j * ld; // B(j, k) for k = 0 ..
                    sacc = j < me ? tv[j] - zv[j] : row[me];
                    const int kmax = j < me ? j : me;
                    for (int k = 0; k < kmax; ++k)
                        sacc -= row[k] * zs[k];
                }
                else {
                    const double* row = Bf + (j - nv) * nfc;
                    sacc = row[me - cb0];
                    for (int k = cb0; k < me; ++k)
                        sacc -= row[k - cb0] * zs[k];
                }
                wv[j] = -sacc;
            }
            fsync<T>();
        }
        // ---------------------------------------------------------------- x = L^-T w: back-substitution on the dv block (one warp), constant factor on the force blocks
        {
            if (T == 1) {
                for (int k = nv - 1; k >= 0; --k) {
                    double s = wv[k];
                    for (int i = k + 1; i < nv; ++i)
                        s -= Hs[i + k * ld] * xv[i];
                    xv[k] = s * il[k];
                }
            }
            else if (tid < 32) {
                // lane l owns rows l and l + 32 (nv <= 64)
                double w0 = tid < nv ? wv[tid] : 0.0, w1 = tid + 32 < nv ? wv[tid + 32] : 0.0;
                auto step = [&](int k) {
                    const double mine = (k >= 32) ? w1 : w0;
                    const double xk = __shfl_sync(0xffffffffu, mine, k & 31) * il[k];
                    if (tid == (k & 31))
                        xv[k] = xk;
                    // w_i -= L[k][i] x_k for i < k
                    if (tid < k)
                        w0 -= Hs[k + tid * ld] * xk;
                    if (k > 32 && tid + 32 < k)
                        w1 -= Hs[k + (tid + 32) * ld] * xk;
                };
                if constexpr (NV_ > 0 && IWBC_OPT_XBS) {
                    // (compile-time k: the operand loads of later steps are hoisted above the shuffle chain, the upper half is
                    // skipped once k <= 32)
#pragma unroll
                    for (int k = NV_ - 1; k >= 0; --k)
                        step(k);
                }
                else {
                    for (int k = nv - 1; k >= 0; --k)
                        step(k);
                }
            }
            else {
                for (int kk = tid - 32; kk < 12 * F.nca; kk += (T > 32 ? T - 32 : 1)) {
                    const int s = kk / 12, i = kk - 12 * s;
                    const double* U = Usel(s) + i * 12;
                    double a = 0;
                    for (int j = i; j < 12; ++j)
                        a += U[j] * wv[nv + 12 * s + j];
                    xv[nv + 12 * s + i] = a;
                }
            }
            if (T == 1)
                for (int kk = 0; kk < 12 * F.nca; ++kk) {
                    const int s = kk / 12, i = kk - 12 * s;
                    const double* U = Usel(s) + i * 12;
                    double a = 0;
                    for (int j = i; j < 12; ++j)
                        a += U[j] * wv[nv + 12 * s + j];
                    xv[nv + 12 * s + i] = a;
                }
        }
        fsync<T>();
        if (cyc && tid == 0)
            cyc[4] = clock64();

        // ---------------------------------------------------------------- s = CI x + ci0: eiquadprog's termination test at x_eq
        Sizes z;
        z.nv = nv, z.nq = nq, z.nj = nj, z.nf = nfr, z.nc = gP->ncontacts, z.n = n, z.mi = S.mi, z.rows_dense = S.rows_dense, z.nref = gP->nref;
        z.ldJ = S.ldJ, z.ldM = ld, z.msym = 1;
        z.S = &S, z.P = gP;
        FusedRows::wrenches<T>(sm, z, tid);
        fsync<T>();
        double psi = 0;
        for (int i = tid; i < na; i += T) {
            const double t = FusedRows::act_row_value(sm, z, i);
            taut[i] = t;
            if (F.ci_act >= 0) {
                const double tm = gP->tau_max[i], h = sm.nle[nu + i];
                psi += fmin(0.0, t + tm + h) + fmin(0.0, -t + tm - h);
            }
            if (F.ci_bounds >= 0) {
                const double xi = xv[nu + i];
                psi += fmin(0.0, xi - sm.lb[i]) + fmin(0.0, -xi + sm.ub[i]);
            }
        }
        for (int e = T - 1 - tid; e < 17 * F.nca; e += T) {
            const int s = e / 17, r = e - 17 * s;
            const iwbc_contact_t& C = gP->contacts[F.cact[s]];
            const double* f = xv + nv + 12 * s;
            double y = 0;
#pragma unroll
            for (int k = 0; k < 12; ++k)
                y += C.B[r * 12 + k] * f[k];
            psi += fmin(0.0, y - C.lb[r]) + fmin(0.0, -y + C.ub[r]);
        }
        psi = fblock_sum<T>(psi, Pp, tid);
        const double c1 = tv[2 * (me + 2)], c2 = tv[2 * (me + 2) + 1];
        if (!spd)
            return -1;
        if (!(fabs(psi) <= S.mi * 2.220446049250313e-16 * c1 * c2 * 100.0))
            return 0;
        if (cyc && tid == 0)
            cyc[5] = clock64();
        // ---------------------------------------------------------------- decode, integrate, store (status OPTIMAL after 1 iteration, empty active set)
        sm.d = wv;
        sm.x_old = gvec;
        sm.act = nullptr;
        k_dec<T>(gP, sm, z, io, inst, tid, IWBC_STATUS_OPTIMAL, 1, S.me, taut);
        return 1;
    }

    // Active-set continuation of fast_instance() (which returned <= 0 and left L, the reflectors, R, the Gram matrix, x_eq, M,
    // the contact Jacobians and the bounds in shared memory): builds J = L^-T Q explicitly (eiquadprog's J after the
    // equality phase, up to an orthogonal change of basis inside the two column groups, which leaves every quantity of
    // the iteration -- z, r, step lengths, multipliers -- unchanged) and runs the Goldfarb-Idnani inequality loop.
    template <int T, int NV_ = 0, int ME_ = 0>
    __device__ void slow_continue(const iwbc_problem_t* __restrict__ gP, const FPlan* __restrict__ gF, const Shape& S, const FShape& F, double* smem,
        const IO& io, int inst, int tid, int code, int& flip, long long* cyc)
    {
        const int nv = NV_ ? NV_ : gP->nv, me = ME_ ? ME_ : F.me, NC = me + 1, n = ME_ ? nv + 2 * (me - 6) : S.n;
        const int ld = (nv + me + 2) & ~1, ldJ = n | 1;
        const int nfrc = 2 * (me - 6), nfc = 7, cb0 = me - 6;
        double* const Hs = smem + F.o_Y;
        double* const Z = smem + F.o_Z;
        double* const Bf = Z + F.z_Bf;
        double* const Gm = Z + F.z_G;
        double* const Rm = smem + F.o_Rm;
        double* const tauv = Rm + me * NC;
        double* const scalev = tauv + me;
        double* const tv = smem + F.o_tv;
        double* const il = smem + F.o_il;
        double* const frv = smem + F.o_fr;
        const double* const Uq0 = gF->Uff[(int)frv[6]][0];
        const double* const Uq1 = F.nca > 1 ? gF->Uff[(int)frv[14]][1] : Uq0;
        auto Usel = [&](int s) { return s == 0 ? Uq0 : (s == 1 ? Uq1 : gF->Uff[(int)frv[8 * s + 6]][s]); }; // (humanoids: two contacts)
        Sizes z;
        z.nv = nv, z.nq = gP->nq, z.nj = gP->nj, z.nf = gP->nframes, z.nc = gP->ncontacts, z.n = n, z.mi = S.mi, z.rows_dense = S.rows_dense;
        z.nref = gP->nref, z.ldJ = ldJ, z.ldM = 0, z.msym = 2;
        z.S = &S, z.P = gP;
        Sm sm;
        sm.qs = smem + F.o_qs, sm.vs = smem + F.o_vs, sm.ref = Z;
        sm.M = smem + F.s_M, sm.Md = nullptr, sm.Jft = smem + F.o_Jft, sm.cb = smem + F.o_cb, sm.nle = smem + F.o_nle, sm.lb = smem + F.o_lb, sm.ub = smem + F.o_ub;
        sm.x = smem + F.o_x, sm.g = smem + F.o_g, sm.sc = smem + F.o_sc;
        sm.dofj = reinterpret_cast<int*>(smem + F.o_int);
        sm.J = smem + F.s_J;
        sm.R = smem + F.s_R;
        {
            double* p = smem + F.s_vec;
            const int L = align2(n + 2);
            sm.d = p, p += L;
            sm.z = p, p += L;
            sm.np = p, p += L;
            sm.r = p, p += L;
            sm.u = p, p += L;
            sm.x_old = p, p += L;
            sm.u_old = p, p += L;
            sm.gc = p, p += L;
            sm.gs = p, p += L;
            sm.gx = p, p += L;
            sm.sub = p, p += L;
            sm.s = p, p += align2(S.mi);
            sm.red = p, p += 32;
            int* ip = reinterpret_cast<int*>(smem + F.s_int);
            sm.act = ip, ip += n + 2;
            sm.act_old = ip, ip += n + 2;
            sm.iai = ip;
        }
        // The mass matrix leaves the augmented matrix (fast_instance kept its strict lower triangle in that matrix's upper triangle, which
        // R and the J construction overwrite): packed lower triangle in the continuation's own region.  (Readers and overwriters are
        // all behind at least one barrier.)
        {
            const double* Md = smem + F.o_Md;
            for (int e = tid; e < nv * (nv + 1) / 2; e += T) {
#if IWBC_OPT_MTAB
                const int r = gF->tri_r[e], c = gF->tri_c[e];
#else
                int r = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
                while (r * (r + 1) / 2 > e)
                    --r;
                while ((r + 1) * (r + 2) / 2 <= e)
                    ++r;
                const int c = e - r * (r + 1) / 2;
#endif
                sm.M[e] = r == c ? Md[r] : Hs[c + r * ld];
            }
        }
        if (code < 0) { // Cholesky failed: eiquadprog cannot start (gi_solve's convention)
            for (int k = tid; k < n; k += T)
                sm.x[k] = 0.0;
            fsync<T>();
            k_dec<T>(gP, sm, z, io, inst, tid, IWBC_STATUS_INFEASIBLE, 0, 0, nullptr);
            return;
        }
        if constexpr (NV_ > 0 && ME_ > 0) {
            // ---- J = L^-T Q, one column per thread, register-resident.  Column c of Q = I - V T V' is e_c - V x_c with
            // x_c = T V[c, :]' (back-substitution on T^-1 = striu(V'V) + diag(1/tau)); its dv rows are then pushed through
            // L^-T by a right-looking back-substitution whose running right-hand side (NV_ doubles) never leaves registers.
            // Every L / reflector operand is warp-uniform (a broadcast shared-memory load), so the phase is FP64-issue bound
            // instead of waiting on the longest of NV_ unequal forward substitutions.
            for (int c = tid; c < n; c += T) {
                double* col = sm.J + c * ldJ;
                double zq[ME_];
#pragma unroll
                for (int i = ME_ - 1; i >= 0; --i) {
                    double y;
                    if (c < NV_)
                        y = (c == i) ? 1.0 : (c > i ? scalev[i] * Hs[(NV_ + i) + c * ld] : 0.0);
                    else
                        y = i >= cb0 ? scalev[i] * Bf[(c - NV_) * nfc + (i - cb0)] : 0.0;
#pragma unroll
                    for (int k = i + 1; k < ME_; ++k)
                        y -= Gm[i * (ME_ + 1) + k] * zq[k];
                    zq[i] = y * tauv[i];
                }
                double r[NV_];
#pragma unroll
                for (int j = 0; j < NV_; ++j)
                    r[j] = (j == c ? 1.0 : 0.0) - (j < ME_ ? zq[j] : 0.0); // unit diagonal of V
#pragma unroll
                for (int k = 0; k < ME_; ++k)
                    zq[k] *= scalev[k];
                // strictly-lower part of V_v: V[j][k] = scalev[k] * Hs[(NV_ + k) + j ld], k < min(j, ME_)
#pragma unroll
                for (int j = 1; j < NV_; ++j) {
                    const double* vj = Hs + NV_ + j * ld;
                    double a0 = 0, a1 = 0;
#pragma unroll
                    for (int k = 0; k < ME_; ++k)
                        if (k < j) {
                            if (k & 1)
                                a1 += vj[k] * zq[k];
                            else
                                a0 += vj[k] * zq[k];
                        }
                    r[j] -= a0 + a1;
                }
                // force rows: Q_f[:, c] first (O(1) entries, no cancellation against the 1e4-sized constant factor), then U_ff Q_f
                for (int sidx = 0; sidx < (ME_ - 6) / 6; ++sidx) {
                    double qf[12];
#pragma unroll
                    for (int j2 = 0; j2 < 12; ++j2) {
                        const int jj = 12 * sidx + j2;
                        const double* row = Bf + jj * nfc;
                        double a0 = (NV_ + jj == c) ? 1.0 : 0.0;
#pragma unroll
                        for (int k = ME_ - 6; k < ME_; ++k)
                            a0 -= row[k - (ME_ - 6)] * zq[k];
                        qf[j2] = a0;
                    }
#pragma unroll
                    for (int i2 = 0; i2 < 12; ++i2) {
                        const double* U = Usel(sidx) + i2 * 12;
                        double a0 = 0;
#pragma unroll
                        for (int j2 = 0; j2 < 12; ++j2)
                            if (j2 >= i2)
                                a0 += U[j2] * qf[j2];
                        col[NV_ + 12 * sidx + i2] = a0;
                    }
                }
                // y = L^-T r (right-looking): y_i = r_i / L_ii, then r_k -= L[i][k] y_i for k < i
#pragma unroll
                for (int i = NV_ - 1; i >= 0; --i) {
                    const double yi = r[i] * il[i];
                    col[i] = yi;
                    const double* Li = Hs + i; // L[i][k] = Li[k ld]
#pragma unroll
                    for (int k = 0; k < i; ++k)
                        r[k] -= Li[k * ld] * yi;
                }
                // (measured and dropped: two rows per step with one 16-byte load of (L[i-1][k], L[i][k]) feeding two FMAs -- neutral)
            }
        }
        else {
            // ---- J = L^-T Q = U - (U V)(T V'),  U = L^-T (block diagonal: dv block from L, constant force blocks)
            // (1) U into J: thread j builds row j of U (= column j of L^-1) by forward substitution; everything else is zero
            for (int e = tid; e < n * ldJ; e += T)
                sm.J[e] = 0.0;
            fsync<T>();
            for (int j = tid; j < n; j += T) {
                if (j < nv) {
                    double* Uj = sm.J + j; // U[j][i] = Uj[i * ldJ]
                    Uj[j * ldJ] = il[j];
                    for (int i = j + 1; i < nv; ++i) {
                        const double* Li = Hs + i; // L[i][k] = Li[k * ld]
                        double s0 = 0, s1 = 0;
                        int k = j;
                        for (; k + 1 < i; k += 2) {
                            s0 += Li[k * ld] * Uj[k * ldJ];
                            s1 += Li[(k + 1) * ld] * Uj[(k + 1) * ldJ];
                        }
                        if (k < i)
                            s0 += Li[k * ld] * Uj[k * ldJ];
                        Uj[i * ldJ] = -(s0 + s1) * il[i];
                    }
                }
                else {
                    const int ii = j - nv, sidx = ii / 12, i2 = ii - 12 * sidx;
                    const double* U = Usel(sidx) + i2 * 12;
                    for (int j2 = i2; j2 < 12; ++j2)
                        sm.J[j + (nv + 12 * sidx + j2) * ldJ] = U[j2];
                }
            }
            fsync<T>();
            // (2) W = U_vv V_v (nv x me) into the top block of the augmented matrix (L is dead): W[i][k] -> Hs[i + k ld]
            for (int e = tid; e < nv * me; e += T) {
                const int k = e / nv, i = e - k * nv;
                double val;
                {
                    const double* Ui = sm.J + i;
                    const double* colk = Hs + (nv + k);
                    double d0 = 0, d1 = 0;
                    int j = i > k + 1 ? i : k + 1;
                    for (; j + 1 < nv; j += 2) {
                        d0 += Ui[j * ldJ] * colk[j * ld];
                        d1 += Ui[(j + 1) * ldJ] * colk[(j + 1) * ld];
                    }
                    if (j < nv)
                        d0 += Ui[j * ldJ] * colk[j * ld];
                    val = (i <= k ? Ui[k * ldJ] : 0.0) + scalev[k] * (d0 + d1);
                    Hs[i + k * ld] = val;
                }
            }
            fsync<T>();
            // (3) one column per thread: x_c = T V[c, :]' (back-substitution on T^-1 = striu(V'V) + diag(1/tau)), J[:, c] -= W x_c
            for (int c = tid; c < n; c += T) {
                double* col = sm.J + c * ldJ;
                constexpr int MEC = ME_ > 0 ? ME_ : IWBC_FMAXME; // compile-time bound: zq stays in registers
                double zq[MEC];
    #pragma unroll
                for (int i = MEC - 1; i >= 0; --i) {
                    if (i >= me) {
                        zq[i] = 0.0;
                        continue;
                    }
                    double y;
                    if (c < nv)
                        y = (c == i) ? 1.0 : (c > i ? scalev[i] * Hs[(nv + i) + c * ld] : 0.0);
                    else
                        y = i >= cb0 ? scalev[i] * Bf[(c - nv) * nfc + (i - cb0)] : 0.0;
    #pragma unroll
                    for (int k = i + 1; k < MEC; ++k)
                        if (k < me)
                            y -= Gm[i * (me + 1) + k] * zq[k];
                    zq[i] = y * tauv[i];
                }
                for (int i = 0; i < nv; i += 2) {
                    const double* w0 = Hs + i;
                    double a0 = 0, a1 = 0;
                    if (i + 1 < nv) {
    #pragma unroll
                        for (int k = 0; k < MEC; ++k)
                            if (k < me) {
                                a0 += w0[k * ld] * zq[k];
                                a1 += w0[k * ld + 1] * zq[k];
                            }
                        col[i] -= a0;
                        col[i + 1] -= a1;
                    }
                    else {
    #pragma unroll
                        for (int k = 0; k < MEC; ++k)
                            if (k < me)
                                a0 += w0[k * ld] * zq[k];
                        col[i] -= a0;
                    }
                }
                // force rows: Q_f[:, c] first (O(1) entries, no cancellation against the 1e4-sized constant factor), then U_ff Q_f.
                // (U_ff e_c - (U_ff V_f) x_c loses the null-space-of-T force components, whose Hessian eigenvalue is the 1e-8
                // regularisation.)
                for (int sidx = 0; sidx < F.nca; ++sidx) {
                    double qf[12];
    #pragma unroll
                    for (int j2 = 0; j2 < 12; ++j2) {
                        const int jj = 12 * sidx + j2;
                        const double* row = Bf + jj * nfc;
                        double a0 = (nv + jj == c) ? 1.0 : 0.0;
    #pragma unroll
                        for (int k = 0; k < MEC; ++k)
                            if (k >= cb0 && k < me)
                                a0 -= row[k - cb0] * (zq[k] * scalev[k]);
                        qf[j2] = a0;
                    }
    #pragma unroll
                    for (int i2 = 0; i2 < 12; ++i2) {
                        const double* U = Usel(sidx) + i2 * 12;
                        double a0 = 0;
    #pragma unroll
                        for (int j2 = 0; j2 < 12; ++j2)
                            if (j2 >= i2)
                                a0 += U[j2] * qf[j2];
                        col[nv + 12 * sidx + i2] = a0;
                    }
                }
            }
        }
        // ---- R (packed), R_norm, bookkeeping of the equality part of the active set.  R aliases the augmented matrix and the
        // iteration vectors alias Rm / Gm / ...: every reader of those is done after the first barrier.
        double R_norm = 1.0;
        for (int k = 0; k < me; ++k)
            R_norm = fmax(R_norm, fabs(Rm[k * NC + k]));
        const double c1 = tv[2 * (me + 2)], c2 = tv[2 * (me + 2) + 1];
        fsync<T>();
        for (int e = tid; e < me * me; e += T) {
            const int i = e / me, j = e - i * me;
            if (i <= j)
                sm.R[Ridx(i, j)] = Rm[i * NC + j];
        }
        fsync<T>();
        for (int k = tid; k < me; k += T) {
            sm.act[k] = -k - 1;
            sm.u[k] = 0.0;
        }
        fsync<T>();
        if (cyc && tid == 0)
            cyc[2] = clock64();
        int iters = 0, iq = me;
        int status;
        if constexpr (T >= 64)
            status = gi_loop_cta<T>(gP, sm, z, me, S.mi, tid, flip, iters, iq, c1, c2, R_norm, (cyc && io.prof) ? io.prof + 16 : nullptr);
        else
            status = gi_loop<T, FusedRows>(gP, sm, z, (const FusedRows*)nullptr, me, S.mi, tid, flip, iters, iq, c1, c2, R_norm);
        if (cyc && tid == 0)
            cyc[3] = clock64();
        k_dec<T>(gP, sm, z, io, inst, tid, status, iters, iq, nullptr);
    }

#ifndef IWBC_CPU_SIM
    // One CTA of T threads per instance, persistent over the batch.  Instances that need the active-set iteration are
    // appended to io.slow_list (order is arbitrary; every instance's results are independent of it).
    template <int T, int NV_, int ME_>
    __global__ void __launch_bounds__(T, IWBC_FAST_MINB)
        k_fast(const iwbc_problem_t* __restrict__ gP, const FPlan* __restrict__ gF, const __grid_constant__ Shape S, const __grid_constant__ FShape F, IO io)
    {
        extern __shared__ double smem[];
        const int tid = threadIdx.x;
        {
            int* dofj = reinterpret_cast<int*>(smem + F.o_int);
            for (int k = tid; k < gP->nv; k += T)
                dofj[k] = gF->dofj[k];
        }
        __syncthreads();
        // all instances, or the device-side list a behaviour built for this contact mask (io.in_list)
        const int count = io.in_list ? *io.in_count : io.B;
        for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
            const int inst = io.in_list ? io.in_list[idx] : idx;
#if IWBC_OPT_PREFETCH
            // the next instance's state and reference rows towards L2 while this one computes (the inputs are larger than L2 and
            // the load phase of an instance otherwise waits on DRAM with nothing to overlap inside the CTA)
            if (idx + (int)gridDim.x < count) {
                const int nxt = io.in_list ? io.in_list[idx + gridDim.x] : idx + (int)gridDim.x;
                const int nq_ = gP->nq, nv_ = gP->nv, nr_ = gP->nref;
                const int lq = (nq_ * 8 + 127) / 128, lv = (nv_ * 8 + 127) / 128, lr = (nr_ * 8 + 127) / 128;
                for (int l = tid; l < lq + lv + lr; l += T) {
                    const char* pbase = l < lq ? reinterpret_cast<const char*>(io.q + (size_t)nxt * nq_) + 128 * l
                        : l < lq + lv    ? reinterpret_cast<const char*>(io.v + (size_t)nxt * nv_) + 128 * (l - lq)
                                         : reinterpret_cast<const char*>(io.refs + (size_t)nxt * io.refs_stride) + 128 * (l - lq - lv);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pbase));
                }
            }
#endif
            long long stamps[8];
            long long* cyc = io.prof ? stamps : nullptr;
            const int code = fast_instance<T, NV_, ME_>(gP, gF, S, F, smem, io, inst, tid, cyc);
            const bool done = code > 0;
            if (!done) {
                // hand the instance to the active-set kernel, together with the shared-memory image of the front end
                // (L, reflectors, R, Gram matrix, x_eq, M, Jacobians, bounds: F.total doubles) so that it is not recomputed
                __shared__ int s_pos;
                __syncthreads();
                if (tid == 0) {
                    s_pos = atomicAdd(io.slow_count, 1);
                    io.slow_list[s_pos] = inst;
                }
                __syncthreads();
                const int pos = s_pos;
                if (pos < io.scratch_cap) {
                    const int n2 = F.total >> 1; // offsets are even: whole double2's
                    double2* dst = reinterpret_cast<double2*>(io.scratch) + (size_t)pos * n2;
                    const double2* src = reinterpret_cast<const double2*>(smem);
                    for (int e = tid; e < n2; e += T)
                        dst[e] = src[e];
                    if (tid == 0)
                        io.slow_code[pos] = code;
                }
            }
            else if (tid == 0 && cyc) {
                cyc[6] = clock64();
                for (int k = 0; k < 6; ++k)
                    atomicAdd(io.prof + 8 + k, (unsigned long long)(cyc[k + 1] - cyc[k]));
                atomicAdd(io.prof + 14, 1ull);
            }
            __syncthreads();
        }
    }

    // The instances the fast kernel could not finish (io.list): same front end, then J and the active-set iteration.
    template <int T, int NV_, int ME_>
    __global__ void __launch_bounds__(T, 2)
        k_slow(const iwbc_problem_t* __restrict__ gP, const FPlan* __restrict__ gF, const __grid_constant__ Shape S, const __grid_constant__ FShape F, IO io)
    {
        extern __shared__ double smem[];
        const int tid = threadIdx.x;
        {
            int* dofj = reinterpret_cast<int*>(smem + F.o_int);
            for (int k = tid; k < gP->nv; k += T)
                dofj[k] = gF->dofj[k];
        }
        __syncthreads();
        int flip = 0;
        const int count = io.list ? *io.list_count : io.B;
        // dynamic hand-out of list positions: the listed instances differ in active-set iterations (1 .. 40), a static stride
        // leaves most CTAs idle while the longest ones finish
        __shared__ int s_next;
        for (int idx = blockIdx.x;; idx += gridDim.x) {
            if (io.work) {
                if (tid == 0)
                    s_next = atomicAdd(io.work, 1);
                __syncthreads();
                idx = s_next;
            }
            if (idx >= count)
                break;
            const int inst = io.list ? io.list[idx] : idx;
            long long stamps[8];
            long long* cyc = io.prof ? stamps : nullptr;
            if (cyc && tid == 0)
                cyc[7] = clock64();
            int code;
            if (io.list && idx < io.scratch_cap) {
                const int n2 = F.total >> 1;
                const double2* src = reinterpret_cast<const double2*>(io.scratch) + (size_t)idx * n2;
                double2* dst = reinterpret_cast<double2*>(smem);
                for (int e = tid; e < n2; e += T)
                    dst[e] = src[e];
                code = io.slow_code[idx];
                __syncthreads();
            }
            else
                code = fast_instance<T, NV_, ME_>(gP, gF, S, F, smem, io, inst, tid, nullptr);
            if (code <= 0) {
                if (cyc && tid == 0)
                    cyc[1] = clock64();
                slow_continue<T, NV_, ME_>(gP, gF, S, F, smem, io, inst, tid, code, flip, cyc);
                if (cyc && tid == 0 && code == 0) {
                    const long long t4 = clock64();
                    atomicAdd(io.prof + 0, (unsigned long long)(cyc[1] - cyc[7]));
                    atomicAdd(io.prof + 1, (unsigned long long)(cyc[2] - cyc[1]));
                    atomicAdd(io.prof + 2, (unsigned long long)(cyc[3] - cyc[2]));
                    atomicAdd(io.prof + 3, (unsigned long long)(t4 - cyc[3]));
                    atomicAdd(io.prof + 6, 1ull);
                }
            }
            __syncthreads();
        }
    }
#endif

} // namespace dev
} // namespace iwbc
