/*
 * blitz/array.h -- stand-in for the subset of Blitz++ that SARAS's lib/poisson, lib/grid,
 * lib/parallel and lib/io/parser use (and that the drop-in poisson class in
 * saras_b200/host is written against).  Written from the Blitz++ user manual's semantics,
 * not from Blitz++ sources:
 *   - Array<T,N> (N = 1..3): reference-counted storage, row-major, arbitrary index bases,
 *     copy construction = reference, operator= = element-wise copy, subarrays (Range /
 *     RectDomain / mixed int-Range slicing with rank reduction) are views sharing storage
 *     and keeping the parent's index base in every surviving dimension;
 *   - TinyVector<T,N> with comma initialisation and element-wise arithmetic;
 *   - Range (with stride, Range::all()), RectDomain<N>;
 *   - the handful of expression forms the sources use (unary minus, array - array,
 *     fabs, pow(array, scalar)) evaluated eagerly into temporaries: the floating-point
 *     operations and their order are the same as lazy evaluation;
 *   - full reductions max / mean / sum in storage order (one sequential accumulator);
 *   - (for the reference's lib/field, lib/boundary and src/solvers, which the test infrastructure compiles unmodified to pin its
 *     CPU restatements of the projection step and the implicit solves) element-wise + - * / between arrays, views and scalars, and the tensor-index form
 *     `A1(range)(i) * B3` (firstIndex / secondIndex / thirdIndex): all evaluated eagerly into temporaries, which performs the same
 *     floating-point operations in the same order per element as Blitz's lazy expression templates.
 * With a real Blitz++ installation this directory is simply left off the include path.
 */
#ifndef SARAS_COMPAT_BLITZ_ARRAY_H
#define SARAS_COMPAT_BLITZ_ARRAY_H

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>

namespace blitz {

/* ------------------------------------------------------------------ TinyVector */
template <class T, int N>
class TinyVector {
    T d_[N];

   public:
    class ListInit {
        TinyVector *v_;
        int idx_;

       public:
        ListInit(TinyVector *v, int idx) : v_(v), idx_(idx) {}
        ListInit operator,(T x) {
            assert(idx_ < N);
            v_->d_[idx_] = x;
            return ListInit(v_, idx_ + 1);
        }
    };

    TinyVector() { for (int i = 0; i < N; i++) d_[i] = T(); }
    TinyVector(T x) { for (int i = 0; i < N; i++) d_[i] = x; }
    TinyVector(T a, T b) { static_assert(N == 2, "rank"); d_[0] = a; d_[1] = b; }
    TinyVector(T a, T b, T c) { static_assert(N == 3, "rank"); d_[0] = a; d_[1] = b; d_[2] = c; }
    template <class U>
    TinyVector(const TinyVector<U, N> &o) { for (int i = 0; i < N; i++) d_[i] = T(o[i]); }

    ListInit operator=(T x) {
        for (int i = 0; i < N; i++) d_[i] = x;
        return ListInit(this, 1);
    }

    T &operator[](int i) { return d_[i]; }
    const T &operator[](int i) const { return d_[i]; }
    T &operator()(int i) { return d_[i]; }
    const T &operator()(int i) const { return d_[i]; }
    int length() const { return N; }
    int lbound(int) const { return 0; }
    int ubound(int) const { return N - 1; }
    T *data() { return d_; }
    const T *data() const { return d_; }

    TinyVector &operator+=(const TinyVector &o) { for (int i = 0; i < N; i++) d_[i] += o.d_[i]; return *this; }
    TinyVector &operator-=(const TinyVector &o) { for (int i = 0; i < N; i++) d_[i] -= o.d_[i]; return *this; }
    TinyVector &operator+=(T x) { for (int i = 0; i < N; i++) d_[i] += x; return *this; }
    TinyVector &operator-=(T x) { for (int i = 0; i < N; i++) d_[i] -= x; return *this; }
};

#define SARAS_TV_BINOP(op)                                                                       \
    template <class T, int N>                                                                    \
    TinyVector<T, N> operator op(const TinyVector<T, N> &a, const TinyVector<T, N> &b) {         \
        TinyVector<T, N> r;                                                                      \
        for (int i = 0; i < N; i++) r[i] = a[i] op b[i];                                         \
        return r;                                                                                \
    }                                                                                            \
    template <class T, int N>                                                                    \
    TinyVector<T, N> operator op(const TinyVector<T, N> &a, T b) {                               \
        TinyVector<T, N> r;                                                                      \
        for (int i = 0; i < N; i++) r[i] = a[i] op b;                                            \
        return r;                                                                                \
    }                                                                                            \
    template <class T, int N>                                                                    \
    TinyVector<T, N> operator op(T a, const TinyVector<T, N> &b) {                               \
        TinyVector<T, N> r;                                                                      \
        for (int i = 0; i < N; i++) r[i] = a op b[i];                                            \
        return r;                                                                                \
    }
SARAS_TV_BINOP(+)
SARAS_TV_BINOP(-)
SARAS_TV_BINOP(*)
SARAS_TV_BINOP(/)
#undef SARAS_TV_BINOP

template <class T, int N>
TinyVector<T, N> operator-(const TinyVector<T, N> &a) {
    TinyVector<T, N> r;
    for (int i = 0; i < N; i++) r[i] = -a[i];
    return r;
}

template <class T, int N>
std::ostream &operator<<(std::ostream &os, const TinyVector<T, N> &v) {
    os << "(";
    for (int i = 0; i < N; i++) os << v[i] << (i + 1 < N ? "," : "");
    return os << ")";
}

/* ------------------------------------------------------------------ Range, RectDomain */
class Range {
    int first_, last_, stride_;

   public:
    enum { fromStart = INT_MIN, toEnd = INT_MAX };
    Range() : first_(fromStart), last_(toEnd), stride_(1) {}
    Range(int first, int last, int stride = 1) : first_(first), last_(last), stride_(stride) {}
    static Range all() { return Range(); }
    int first(int lowRange = 0) const { return first_ == fromStart ? lowRange : first_; }
    int last(int highRange = 0) const { return last_ == toEnd ? highRange : last_; }
    int stride() const { return stride_; }
    int length() const { return (last_ - first_) / stride_ + 1; }
};

template <int N>
class RectDomain {
    TinyVector<int, N> lb_, ub_;

   public:
    RectDomain() {}
    RectDomain(const TinyVector<int, N> &lb, const TinyVector<int, N> &ub) : lb_(lb), ub_(ub) {}
    const TinyVector<int, N> &lbound() const { return lb_; }
    const TinyVector<int, N> &ubound() const { return ub_; }
    TinyVector<int, N> &lbound() { return lb_; }
    TinyVector<int, N> &ubound() { return ub_; }
    int lbound(int d) const { return lb_[d]; }
    int ubound(int d) const { return ub_[d]; }
    Range operator[](int d) const { return Range(lb_[d], ub_[d]); }
};

/* ------------------------------------------------------------------ tensor-index placeholders */
struct firstIndex {};
struct secondIndex {};
struct thirdIndex {};
enum { firstDim = 0, secondDim = 1, thirdDim = 2 };
template <class T, int N> class Array;
/* `A1(i)`: the value of the 1-D array at the ABSOLUTE index of dimension D of the array it is combined with */
template <class T, int D>
struct IndexedVec { Array<T, 1> v; };

/* ------------------------------------------------------------------ Array */
template <class A> struct is_range_ : std::false_type {};
template <> struct is_range_<Range> : std::true_type {};

template <class T, int N>
class Array {
    template <class U, int M> friend class Array;
    std::shared_ptr<T> block_;
    T *zero_;                 /* address of the (possibly virtual) element with all indices 0 */
    int base_[N], ext_[N];
    std::ptrdiff_t stride_[N];

    void setup(const int *ext, const int *base) {
        std::size_t n = 1;
        for (int d = N - 1; d >= 0; d--) {
            ext_[d] = ext[d]; base_[d] = base[d];
            stride_[d] = (std::ptrdiff_t)n;
            n *= (std::size_t)ext[d];
        }
        if (n > 0) block_.reset(new T[n](), std::default_delete<T[]>()); else block_.reset();
        computeZero(block_.get());
    }
    void computeZero(T *first) {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++) off += (std::ptrdiff_t)base_[d] * stride_[d];
        zero_ = first - off;
    }
    template <class F>
    void forEachIndex(F f) const {          /* row-major traversal of POSITIONS 0..ext-1 */
        int p[N];
        for (int d = 0; d < N; d++) { p[d] = 0; if (ext_[d] <= 0) return; }
        while (true) {
            f(p);
            int d = N - 1;
            while (d >= 0) { if (++p[d] < ext_[d]) break; p[d] = 0; d--; }
            if (d < 0) break;
        }
    }
    T &atPos(const int *p) const {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++) off += (std::ptrdiff_t)(p[d] + base_[d]) * stride_[d];
        return zero_[off];
    }
    bool sameShape(const int *e) const { for (int d = 0; d < N; d++) if (ext_[d] != e[d]) return false; return true; }

   public:
    typedef T T_numtype;
    Array() : zero_(nullptr) { for (int d = 0; d < N; d++) { base_[d] = 0; ext_[d] = 0; stride_[d] = 0; } }
    explicit Array(int n0) { static_assert(N == 1, "rank"); int e[1] = {n0}, b[1] = {0}; setup(e, b); }
    Array(int n0, int n1) { static_assert(N == 2, "rank"); int e[2] = {n0, n1}, b[2] = {0, 0}; setup(e, b); }
    Array(int n0, int n1, int n2) { static_assert(N == 3, "rank"); int e[3] = {n0, n1, n2}, b[3] = {0, 0, 0}; setup(e, b); }
    explicit Array(const TinyVector<int, N> &ext) { int b[N]; for (int d = 0; d < N; d++) b[d] = 0; setup(ext.data(), b); }
    /* Blitz: the copy constructor makes a REFERENCE to the same data */
    Array(const Array &o) = default;

    void reference(const Array &o) {
        block_ = o.block_; zero_ = o.zero_;
        for (int d = 0; d < N; d++) { base_[d] = o.base_[d]; ext_[d] = o.ext_[d]; stride_[d] = o.stride_[d]; }
    }

    /* ---- sizing ---- */
    void resize(int n0) { static_assert(N == 1, "rank"); int e[1] = {n0}, b[1] = {base_[0]}; setup(e, b); }
    void resize(int n0, int n1) { static_assert(N == 2, "rank"); int e[2] = {n0, n1}, b[2] = {base_[0], base_[1]}; setup(e, b); }
    void resize(int n0, int n1, int n2) { static_assert(N == 3, "rank"); int e[3] = {n0, n1, n2}, b[3] = {base_[0], base_[1], base_[2]}; setup(e, b); }
    void resize(const TinyVector<int, N> &ext) { int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d]; setup(ext.data(), b); }
    void resize(Range r0) { static_assert(N == 1, "rank"); int e[1] = {r0.length()}, b[1] = {r0.first()}; setup(e, b); }
    void resize(Range r0, Range r1) { static_assert(N == 2, "rank"); int e[2] = {r0.length(), r1.length()}, b[2] = {r0.first(), r1.first()}; setup(e, b); }
    void resize(Range r0, Range r1, Range r2) {
        static_assert(N == 3, "rank");
        int e[3] = {r0.length(), r1.length(), r2.length()}, b[3] = {r0.first(), r1.first(), r2.first()};
        setup(e, b);
    }
    void reindexSelf(const TinyVector<int, N> &nb) {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++) { off += (std::ptrdiff_t)(nb[d] - base_[d]) * stride_[d]; base_[d] = nb[d]; }
        zero_ -= off;
    }

    /* ---- queries ---- */
    TinyVector<int, N> lbound() const { TinyVector<int, N> r; for (int d = 0; d < N; d++) r[d] = base_[d]; return r; }
    TinyVector<int, N> ubound() const { TinyVector<int, N> r; for (int d = 0; d < N; d++) r[d] = base_[d] + ext_[d] - 1; return r; }
    TinyVector<int, N> shape() const { TinyVector<int, N> r; for (int d = 0; d < N; d++) r[d] = ext_[d]; return r; }
    TinyVector<int, N> extent() const { return shape(); }
    int lbound(int d) const { return base_[d]; }
    int ubound(int d) const { return base_[d] + ext_[d] - 1; }
    int extent(int d) const { return ext_[d]; }
    int rows() const { return ext_[0]; }
    std::ptrdiff_t stride(int d) const { return stride_[d]; }
    std::size_t size() const { std::size_t n = 1; for (int d = 0; d < N; d++) n *= (std::size_t)ext_[d]; return n; }
    std::size_t numElements() const { return size(); }
    bool isStorageContiguous() const {
        std::ptrdiff_t n = 1;
        for (int d = N - 1; d >= 0; d--) { if (stride_[d] != n) return false; n *= ext_[d]; }
        return true;
    }
    T *dataFirst() { int p[N]; for (int d = 0; d < N; d++) p[d] = 0; return &atPos(p); }
    const T *dataFirst() const { int p[N]; for (int d = 0; d < N; d++) p[d] = 0; return &atPos(p); }
    T *data() { return dataFirst(); }
    const T *data() const { return dataFirst(); }
    T *dataZero() { return zero_; }

    /* ---- element access ---- */
    T &operator()(int i) { static_assert(N == 1, "rank"); return zero_[i * stride_[0]]; }
    const T &operator()(int i) const { static_assert(N == 1, "rank"); return zero_[i * stride_[0]]; }
    T &operator()(int i, int j) { static_assert(N == 2, "rank"); return zero_[i * stride_[0] + j * stride_[1]]; }
    const T &operator()(int i, int j) const { static_assert(N == 2, "rank"); return zero_[i * stride_[0] + j * stride_[1]]; }
    T &operator()(int i, int j, int k) { static_assert(N == 3, "rank"); return zero_[i * stride_[0] + j * stride_[1] + k * stride_[2]]; }
    const T &operator()(int i, int j, int k) const { static_assert(N == 3, "rank"); return zero_[i * stride_[0] + j * stride_[1] + k * stride_[2]]; }
    T &operator()(const TinyVector<int, N> &ix) {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++) off += (std::ptrdiff_t)ix[d] * stride_[d];
        return zero_[off];
    }
    const T &operator()(const TinyVector<int, N> &ix) const {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++) off += (std::ptrdiff_t)ix[d] * stride_[d];
        return zero_[off];
    }

    /* ---- tensor-index forms ---- */
    IndexedVec<T, 0> operator()(firstIndex) const { static_assert(N == 1, "rank"); IndexedVec<T, 0> r; r.v.reference(*this); return r; }
    IndexedVec<T, 1> operator()(secondIndex) const { static_assert(N == 1, "rank"); IndexedVec<T, 1> r; r.v.reference(*this); return r; }
    IndexedVec<T, 2> operator()(thirdIndex) const { static_assert(N == 1, "rank"); IndexedVec<T, 2> r; r.v.reference(*this); return r; }
    const Array &operator()(firstIndex, secondIndex, thirdIndex) const { static_assert(N == 3, "rank"); return *this; }
    /* element-wise combination with a 1-D array indexed along dimension D (absolute indices) */
    template <int D, class F>
    Array zipIndexed(const Array<T, 1> &v, F f) const {
        Array r;
        int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d];
        r.setup(ext_, b);
        forEachIndex([&](const int *p) { r.atPos(p) = f(v(p[D] + base_[D]), atPos(p)); });
        return r;
    }

    /* ---- subarrays (views) ---- */
   private:
    /* restrict dimension d of *this (in place, as a view) to Range r: the index base is kept */
    void sliceDim(int d, const Range &r) {
        const int first = r.first(base_[d]), last = r.last(base_[d] + ext_[d] - 1), st = r.stride();
        const std::ptrdiff_t off = ((std::ptrdiff_t)first - (std::ptrdiff_t)base_[d] * st) * stride_[d];
        zero_ += off;
        ext_[d] = (last - first) / st + 1;
        stride_[d] *= st;
    }
    struct SliceArg { bool isRange; int idx; Range r; };
    static SliceArg mk(int i) { SliceArg a; a.isRange = false; a.idx = i; return a; }
    static SliceArg mk(const Range &r) { SliceArg a; a.isRange = true; a.idx = 0; a.r = r; return a; }
    template <int M>
    Array<T, M> sliceTo(const SliceArg *args) const {
        Array tmp(*this);
        for (int d = 0; d < N; d++) if (args[d].isRange) tmp.sliceDim(d, args[d].r);
        Array<T, M> out;
        out.block_ = tmp.block_;
        T *z = tmp.zero_;
        int m = 0;
        for (int d = 0; d < N; d++) {
            if (args[d].isRange) { out.base_[m] = tmp.base_[d]; out.ext_[m] = tmp.ext_[d]; out.stride_[m] = tmp.stride_[d]; m++; }
            else z += (std::ptrdiff_t)args[d].idx * tmp.stride_[d];
        }
        out.zero_ = z;
        return out;
    }

   public:
    Array operator()(const RectDomain<N> &dom) const {
        Array tmp(*this);
        for (int d = 0; d < N; d++) tmp.sliceDim(d, Range(dom.lbound(d), dom.ubound(d)));
        return tmp;
    }
    Array<T, 1> operator()(const Range &r0) const {
        static_assert(N == 1, "rank");
        SliceArg a[1] = {mk(r0)};
        return sliceTo<1>(a);
    }
    template <class A0, class A1, typename std::enable_if<(is_range_<A0>::value || is_range_<A1>::value), int>::type = 0>
    Array<T, (is_range_<A0>::value ? 1 : 0) + (is_range_<A1>::value ? 1 : 0)> operator()(const A0 &a0, const A1 &a1) const {
        static_assert(N == 2, "rank");
        SliceArg a[2] = {mk(a0), mk(a1)};
        return sliceTo<(is_range_<A0>::value ? 1 : 0) + (is_range_<A1>::value ? 1 : 0)>(a);
    }
    template <class A0, class A1, class A2,
              typename std::enable_if<(is_range_<A0>::value || is_range_<A1>::value || is_range_<A2>::value), int>::type = 0>
    Array<T, (is_range_<A0>::value ? 1 : 0) + (is_range_<A1>::value ? 1 : 0) + (is_range_<A2>::value ? 1 : 0)>
    operator()(const A0 &a0, const A1 &a1, const A2 &a2) const {
        static_assert(N == 3, "rank");
        SliceArg a[3] = {mk(a0), mk(a1), mk(a2)};
        return sliceTo<(is_range_<A0>::value ? 1 : 0) + (is_range_<A1>::value ? 1 : 0) + (is_range_<A2>::value ? 1 : 0)>(a);
    }

    /* ---- assignment (element-wise; works on views, hence const-qualified like Blitz's rvalue use) ---- */
    Array &operator=(const Array &o) {
        if (block_.get() == nullptr && zero_ == nullptr && o.zero_ != nullptr && size() == 0) {
            /* assigning to a null array: Blitz requires conforming shapes; be lenient and allocate */
            int b[N]; for (int d = 0; d < N; d++) b[d] = o.base_[d];
            setup(o.ext_, b);
        }
        assert(sameShape(o.ext_));
        if (&o == this) return *this;
        forEachIndex([&](const int *p) { atPos(p) = o.atPos(p); });
        return *this;
    }
    Array &operator=(const T &x) { forEachIndex([&](const int *p) { atPos(p) = x; }); return *this; }
    Array &operator+=(const Array &o) { assert(sameShape(o.ext_)); forEachIndex([&](const int *p) { atPos(p) += o.atPos(p); }); return *this; }
    Array &operator-=(const Array &o) { assert(sameShape(o.ext_)); forEachIndex([&](const int *p) { atPos(p) -= o.atPos(p); }); return *this; }
    Array &operator*=(const Array &o) { assert(sameShape(o.ext_)); forEachIndex([&](const int *p) { atPos(p) *= o.atPos(p); }); return *this; }
    Array &operator+=(const T &x) { forEachIndex([&](const int *p) { atPos(p) += x; }); return *this; }
    Array &operator-=(const T &x) { forEachIndex([&](const int *p) { atPos(p) -= x; }); return *this; }
    Array &operator*=(const T &x) { forEachIndex([&](const int *p) { atPos(p) *= x; }); return *this; }
    Array &operator/=(const T &x) { forEachIndex([&](const int *p) { atPos(p) /= x; }); return *this; }

    /* a compact copy with the same bases */
    Array copy() const {
        Array r;
        int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d];
        r.setup(ext_, b);
        r = *this;
        return r;
    }
    /* apply f element-wise into a fresh compact array (eager "expression") */
    template <class F>
    Array map(F f) const {
        Array r;
        int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d];
        r.setup(ext_, b);
        forEachIndex([&](const int *p) { r.atPos(p) = f(atPos(p)); });
        return r;
    }
    template <class F>
    Array zip(const Array &o, F f) const {
        assert(sameShape(o.ext_));
        Array r;
        int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d];
        r.setup(ext_, b);
        forEachIndex([&](const int *p) { r.atPos(p) = f(atPos(p), o.atPos(p)); });
        return r;
    }
    /* the same with an operand of another element type (e.g. a bool mask) */
    template <class U, class F>
    Array zipOther(const Array<U, N> &o, F f) const {
        assert(sameShape(o.ext_));
        Array r;
        int b[N]; for (int d = 0; d < N; d++) b[d] = base_[d];
        r.setup(ext_, b);
        forEachIndex([&](const int *p) { r.atPos(p) = f(atPos(p), o.atPos(p)); });
        return r;
    }
    /* sequential reduction in storage (row-major) order */
    template <class Acc, class F>
    Acc fold(Acc init, F f) const {
        Acc a = init;
        forEachIndex([&](const int *p) { a = f(a, atPos(p)); });
        return a;
    }
};

/* ---- eager expressions ---- */
template <class S> struct is_scalar_ : std::integral_constant<bool, std::is_arithmetic<S>::value> {};
template <class T, int N> Array<T, N> operator-(const Array<T, N> &a) { return a.map([](const T &x) { return -x; }); }
#define SARAS_ARR_BINOP(op)                                                                                                   \
    template <class T, int N>                                                                                                 \
    Array<T, N> operator op(const Array<T, N> &a, const Array<T, N> &b) { return a.zip(b, [](const T &x, const T &y) { return x op y; }); } \
    template <class T, int N, class S, typename std::enable_if<is_scalar_<S>::value, int>::type = 0>                           \
    Array<T, N> operator op(const Array<T, N> &a, const S &s) { const T t = T(s); return a.map([&](const T &x) { return x op t; }); }      \
    template <class T, int N, class S, typename std::enable_if<is_scalar_<S>::value, int>::type = 0>                           \
    Array<T, N> operator op(const S &s, const Array<T, N> &a) { const T t = T(s); return a.map([&](const T &x) { return t op x; }); }      \
    template <class T, int D>                                                                                                  \
    Array<T, 3> operator op(const IndexedVec<T, D> &v, const Array<T, 3> &a) { return a.template zipIndexed<D>(v.v, [](const T &x, const T &y) { return x op y; }); } \
    template <class T, int D>                                                                                                  \
    Array<T, 3> operator op(const Array<T, 3> &a, const IndexedVec<T, D> &v) { return a.template zipIndexed<D>(v.v, [](const T &x, const T &y) { return y op x; }); } \
    template <class T, int D, class S, typename std::enable_if<is_scalar_<S>::value, int>::type = 0>                           \
    IndexedVec<T, D> operator op(const S &s, const IndexedVec<T, D> &v) { IndexedVec<T, D> r; const T t = T(s); r.v.reference(v.v.map([&](const T &x) { return t op x; })); return r; } \
    template <class T, int D, class S, typename std::enable_if<is_scalar_<S>::value, int>::type = 0>                           \
    IndexedVec<T, D> operator op(const IndexedVec<T, D> &v, const S &s) { IndexedVec<T, D> r; const T t = T(s); r.v.reference(v.v.map([&](const T &x) { return x op t; })); return r; }
SARAS_ARR_BINOP(+)
SARAS_ARR_BINOP(-)
SARAS_ARR_BINOP(*)
SARAS_ARR_BINOP(/)
#undef SARAS_ARR_BINOP
template <class T, class U, int N, typename std::enable_if<!std::is_same<T, U>::value, int>::type = 0>
Array<T, N> operator*(const Array<T, N> &a, const Array<U, N> &m) { return a.zipOther(m, [](const T &x, const U &y) { return x * T(y); }); }
template <class T, int N> Array<T, N> fabs(const Array<T, N> &a) { return a.map([](const T &x) { return T(std::fabs(x)); }); }
template <class T, int N> Array<T, N> abs(const Array<T, N> &a) { return a.map([](const T &x) { return T(std::fabs(x)); }); }
template <class T, int N> Array<T, N> sqrt(const Array<T, N> &a) { return a.map([](const T &x) { return T(std::sqrt(x)); }); }
template <class T, int N, class E> Array<T, N> pow(const Array<T, N> &a, E e) { return a.map([&](const T &x) { return T(std::pow(x, e)); }); }
template <class T, int N> Array<T, N> pow2(const Array<T, N> &a) { return a.map([](const T &x) { return x * x; }); }

/* Stencil operators (blitz/array/stencil-et.h) as the reference's derivative.cc uses them:
 *     out(Range(0, core_ub), all, all) = central12n(F, 0);
 * i.e. the stencil expression's domain is the operand's domain shrunk by one point at each end of `dim` -- where both
 * neighbours exist -- and it is assigned, position by position, to a destination view of exactly that shape.  Returned here as
 * a compact array of the shrunk shape whose element p along `dim` is the stencil applied at operand index lbound + 1 + p
 * (central12n: (A(+1) - A(-1)) / 2; central22n: A(-1) - 2 A + A(+1), in that order).  Not on any path this repository pins bit for
 * bit (computeDiff / computeNLin); tests/test_reference_timestep.py checks them against analytic derivatives. */
template <class T, class F>
Array<T, 3> stencil3_(const Array<T, 3> &a, int dim, F f) {
    TinyVector<int, 3> lb = a.lbound(), ext = a.extent();
    ext[dim] -= 2;
    Array<T, 3> r(ext);
    TinyVector<int, 3> nb = lb;
    nb[dim] += 1;
    r.reindexSelf(nb);
    for (int i = nb[0]; i < nb[0] + ext[0]; i++) for (int j = nb[1]; j < nb[1] + ext[1]; j++) for (int k = nb[2]; k < nb[2] + ext[2]; k++) {
        int p[3] = {i, j, k}, m[3] = {i, j, k};
        p[dim] += 1; m[dim] -= 1;
        r(i, j, k) = f(a(p[0], p[1], p[2]), a(i, j, k), a(m[0], m[1], m[2]));
    }
    return r;
}
/* the `n` variants are Blitz's NORMALISED stencils: central12 = A(+1) - A(-1) has the factor 2 h, so central12n = central12 / 2 (the
 * reference multiplies by 1/h only, derivative.cc:122-123); central22 = A(-1) - 2 A + A(+1) has the factor h^2, central22n = central22 */
template <class T>
Array<T, 3> central12n(const Array<T, 3> &a, int dim) { return stencil3_(a, dim, [](const T &p, const T &, const T &m) { return (p - m) * T(0.5); }); }
template <class T>
Array<T, 3> central22n(const Array<T, 3> &a, int dim) { return stencil3_(a, dim, [](const T &p, const T &c, const T &m) { return m - T(2) * c + p; }); }

/* ---- full reductions ---- */
template <class T, int N>
T max(const Array<T, N> &a) {
    bool first = true;
    T m = T();
    return a.fold(m, [&](T acc, const T &x) { if (first) { first = false; return x; } return x > acc ? x : acc; });
}
template <class T, int N>
T min(const Array<T, N> &a) {
    bool first = true;
    T m = T();
    return a.fold(m, [&](T acc, const T &x) { if (first) { first = false; return x; } return x < acc ? x : acc; });
}
template <class T, int N>
T sum(const Array<T, N> &a) { return a.fold(T(0), [](T acc, const T &x) { return acc + x; }); }
template <class T, int N>
T mean(const Array<T, N> &a) { return sum(a) / T(a.size()); }

/* Blitz's swap: exchanges which storage the two Array objects refer to */
template <class T, int N>
void swap(Array<T, N> &a, Array<T, N> &b) {
    Array<T, N> c(a);
    a.reference(b);
    b.reference(c);
}

}  // namespace blitz

#endif
