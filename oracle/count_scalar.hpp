// count_scalar.hpp -- a drop-in `double` that counts its own arithmetic.  TEST INFRASTRUCTURE ONLY.
//
// "F_dyn, F_asm to be counted by instrumenting the oracle (one counter per FMA/add/mul in the scalar
// code)".  Building wbc_oracle.cpp with -DORACLE_COUNT_FLOPS (`make -C oracle count`) replaces every `double` of the
// restatement by this type; each +, -, *, / and each sqrt / sin / cos / exp / log / atan2 / acos / pow bumps the counter of
// the cycle phase that is running (dynamics, task & contact assembly, QP, decode).  An add and a multiply count one FLOP
// each (so a fused multiply-add is 2, as SURVEY.md counts); comparisons, negations and copies are free.
#pragma once
#include <cmath>
#include <limits>
#include <type_traits>

namespace flopcount {
enum { PH_OTHER = 0, PH_DYN = 1, PH_ASM = 2, PH_QP = 3, PH_DEC = 4, NPH = 5 };
enum { OP_ADD = 0, OP_MUL = 1, OP_DIV = 2, OP_FUNC = 3, NOP = 4 };
extern thread_local long counts[NPH][NOP];
extern thread_local int phase;
inline void bump(int op) { counts[phase][op]++; }
} // namespace flopcount

struct cdouble {
    double v;
    cdouble() = default;
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    constexpr cdouble(T x) : v((double)x) {}
    constexpr operator double() const { return v; }
    cdouble& operator+=(cdouble o) { flopcount::bump(0); v += o.v; return *this; }
    cdouble& operator-=(cdouble o) { flopcount::bump(0); v -= o.v; return *this; }
    cdouble& operator*=(cdouble o) { flopcount::bump(1); v *= o.v; return *this; }
    cdouble& operator/=(cdouble o) { flopcount::bump(2); v /= o.v; return *this; }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> cdouble& operator+=(T o) { return *this += cdouble(o); }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> cdouble& operator-=(T o) { return *this -= cdouble(o); }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> cdouble& operator*=(T o) { return *this *= cdouble(o); }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> cdouble& operator/=(T o) { return *this /= cdouble(o); }
};
inline cdouble operator-(cdouble a) { return cdouble(-a.v); }
#define CD_BIN(op, slot)                                                                                                                  \
    inline cdouble operator op(cdouble a, cdouble b) { flopcount::bump(slot); return cdouble(a.v op b.v); }                               \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                                               \
    inline cdouble operator op(cdouble a, T b) { flopcount::bump(slot); return cdouble(a.v op (double)b); }                               \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                                               \
    inline cdouble operator op(T a, cdouble b) { flopcount::bump(slot); return cdouble((double)a op b.v); }
CD_BIN(+, 0)
CD_BIN(-, 0)
CD_BIN(*, 1)
CD_BIN(/, 2)
#undef CD_BIN
#define CD_CMP(op)                                                                                               \
    inline bool operator op(cdouble a, cdouble b) { return a.v op b.v; }                                         \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                      \
    inline bool operator op(cdouble a, T b) { return a.v op (double)b; }                                         \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                      \
    inline bool operator op(T a, cdouble b) { return (double)a op b.v; }
CD_CMP(<)
CD_CMP(>)
CD_CMP(<=)
CD_CMP(>=)
CD_CMP(==)
CD_CMP(!=)
#undef CD_CMP
namespace std {
template <>
struct numeric_limits<cdouble> : numeric_limits<double> {
    static constexpr cdouble infinity() noexcept { return cdouble(numeric_limits<double>::infinity()); }
    static constexpr cdouble epsilon() noexcept { return cdouble(numeric_limits<double>::epsilon()); }
    static constexpr cdouble max() noexcept { return cdouble(numeric_limits<double>::max()); }
    static constexpr cdouble min() noexcept { return cdouble(numeric_limits<double>::min()); }
    static constexpr cdouble lowest() noexcept { return cdouble(numeric_limits<double>::lowest()); }
    static constexpr cdouble quiet_NaN() noexcept { return cdouble(numeric_limits<double>::quiet_NaN()); }
};
#define CD_F1(f) inline cdouble f(cdouble a) { flopcount::bump(3); return cdouble(std::f(a.v)); }
CD_F1(sqrt) CD_F1(sin) CD_F1(cos) CD_F1(exp) CD_F1(log) CD_F1(log1p) CD_F1(acos) CD_F1(asin) CD_F1(floor)
#undef CD_F1
inline cdouble fabs(cdouble a) { return cdouble(std::fabs(a.v)); }
inline cdouble abs(cdouble a) { return cdouble(std::fabs(a.v)); }
inline cdouble atan2(cdouble a, cdouble b) { flopcount::bump(3); return cdouble(std::atan2(a.v, b.v)); }
inline cdouble pow(cdouble a, cdouble b) { flopcount::bump(3); return cdouble(std::pow(a.v, b.v)); }
template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline cdouble pow(cdouble a, T b) { flopcount::bump(3); return cdouble(std::pow(a.v, (double)b)); }
inline bool isfinite(cdouble a) { return std::isfinite(a.v); }
inline bool isnan(cdouble a) { return std::isnan(a.v); }
} // namespace std
