// Operation count of the restated reference algorithm (test / measurement infrastructure, like everything under oracle/).
//
// SURVEY.md section 8d estimates the algorithmic FP32 work of one column at ~6.8e6 flops "+- 30 % -- pin by op-counting
// the oracle".  This translation unit builds the oracle with its Float64 instantiation replaced by `Counted`, a wrapper of
// `double` (same size and layout, so the C entry points and the Python driver work unchanged) whose arithmetic operators
// and math functions increment global counters.  tools/opcount.py runs the headline workload through it and writes
// profiles/oracle_opcount.json.  Single-threaded (the counters are plain globals): call with nthreads = 1.
#include <cmath>
#include <cstdint>
#include <limits>
#include <type_traits>

struct OpCounters { unsigned long long add, mul, div, fma_like, sqrt_, exp_, expm1_, log_, trig, pow_, cmp, minmax; };
static OpCounters g_ops = {};

struct Counted {
    double v;
    Counted() = default;
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    Counted(T x) : v((double)x) {}
    explicit operator double() const { return v; }
    explicit operator float() const { return (float)v; }
    explicit operator int() const { return (int)v; }
    explicit operator long long() const { return (long long)v; }
    explicit operator bool() const { return v != 0; }
    Counted operator-() const { Counted r; r.v = -v; return r; }
    Counted& operator+=(Counted o) { ++g_ops.add; v += o.v; return *this; }
    Counted& operator-=(Counted o) { ++g_ops.add; v -= o.v; return *this; }
    Counted& operator*=(Counted o) { ++g_ops.mul; v *= o.v; return *this; }
    Counted& operator/=(Counted o) { ++g_ops.div; v /= o.v; return *this; }
};
static_assert(sizeof(Counted) == sizeof(double), "Counted must alias double arrays");

#define RB_BINOP(op, field)                                                                                          \
    inline Counted operator op(Counted a, Counted b) { ++g_ops.field; Counted r; r.v = a.v op b.v; return r; }          \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                           \
    inline Counted operator op(Counted a, T b) { ++g_ops.field; Counted r; r.v = a.v op (double)b; return r; }         \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                           \
    inline Counted operator op(T a, Counted b) { ++g_ops.field; Counted r; r.v = (double)a op b.v; return r; }
RB_BINOP(+, add) RB_BINOP(-, add) RB_BINOP(*, mul) RB_BINOP(/, div)
#define RB_CMP(op)                                                                                                    \
    inline bool operator op(Counted a, Counted b) { ++g_ops.cmp; return a.v op b.v; }                                  \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                           \
    inline bool operator op(Counted a, T b) { ++g_ops.cmp; return a.v op (double)b; }                                  \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>                           \
    inline bool operator op(T a, Counted b) { ++g_ops.cmp; return (double)a op b.v; }
RB_CMP(<) RB_CMP(>) RB_CMP(<=) RB_CMP(>=) RB_CMP(==) RB_CMP(!=)

namespace std {
inline Counted sqrt(Counted x) { ++g_ops.sqrt_; return Counted(std::sqrt(x.v)); }
inline Counted exp(Counted x) { ++g_ops.exp_; return Counted(std::exp(x.v)); }
inline Counted expm1(Counted x) { ++g_ops.expm1_; return Counted(std::expm1(x.v)); }
inline Counted log(Counted x) { ++g_ops.log_; return Counted(std::log(x.v)); }
inline Counted sin(Counted x) { ++g_ops.trig; return Counted(std::sin(x.v)); }
inline Counted cos(Counted x) { ++g_ops.trig; return Counted(std::cos(x.v)); }
inline Counted pow(Counted x, Counted y) { ++g_ops.pow_; return Counted(std::pow(x.v, y.v)); }
inline Counted abs(Counted x) { return Counted(std::abs(x.v)); }
inline Counted fabs(Counted x) { return Counted(std::fabs(x.v)); }
inline Counted floor(Counted x) { return Counted(std::floor(x.v)); }
inline Counted trunc(Counted x) { return Counted(std::trunc(x.v)); }
inline bool isfinite(Counted x) { return std::isfinite(x.v); }
inline bool isnan(Counted x) { return std::isnan(x.v); }
inline Counted max(Counted a, Counted b) { ++g_ops.minmax; return a.v < b.v ? b : a; }
inline Counted min(Counted a, Counted b) { ++g_ops.minmax; return b.v < a.v ? b : a; }
template <> struct numeric_limits<Counted> {
    static constexpr bool is_specialized = true;
    static Counted epsilon() { return Counted(numeric_limits<float>::epsilon()); }   // the Float32 guard constants: the counted path is the headline's
    static Counted max() { return Counted(numeric_limits<double>::max()); }
    static Counted min() { return Counted(numeric_limits<double>::min()); }
    static Counted lowest() { return Counted(numeric_limits<double>::lowest()); }
    static Counted infinity() { return Counted(numeric_limits<double>::infinity()); }
    static Counted quiet_NaN() { return Counted(numeric_limits<double>::quiet_NaN()); }
};
}  // namespace std

#define ORACLE_F64_T Counted
#include "rrtmgp_oracle.cpp"

extern "C" void oracle_opcount_read(unsigned long long* out12, int reset) {
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(&g_ops);
    for (int i = 0; i < 12; ++i) out12[i] = p[i];
    if (reset) g_ops = OpCounters{};
}
