// xrt/geometry.h — host-side math PODs of the drop-in API (mirrors the public surface of the reference's
// geometry.h:10-702: Vec2f, Vec3f, Matrix44f, multVecMatrix/multDirMatrix, constants). Written from
// scratch; only the names, argument meaning and arithmetic ORDER that parity depends on are kept
// (row-vector * matrix with translation in row 3, geometry.h:636-669).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <ostream>

#ifndef kEpsilon
#define kEpsilon FLT_EPSILON /* cmakelists.txt:61 */
#endif
#ifndef kInfinity
#define kInfinity FLT_MAX /* cmakelists.txt:62 */
#endif

constexpr float PI = 3.14159265359; // geometry.h:10 (double literal narrowed to float)
constexpr float PI_MUL_2 = 2.0f * PI;
constexpr float PI_MUL_4 = 4.0f * PI;
constexpr float PI_DIV_2 = 0.5f * PI;
constexpr float PI_DIV_4 = 0.25f * PI;
constexpr float PI_INV = 1.0f / PI;
constexpr float PI_MUL_2_INV = 1.0f / PI_MUL_2;
constexpr float PI_MUL_4_INV = 1.0f / PI_MUL_4;
constexpr float EPS = 1e-9f;
constexpr float RAY_EPS = 1e-3f; // geometry.h:23

inline float rad2deg(float rad) { return 180.0f * rad / PI; }
inline float deg2rad(float deg) { return deg / 180.0f * PI; } // geometry.h:26 — order matters for tan(FOV/2)

enum class MaterialType { Unknow, Lambert, Metals, Glass };

template <typename T, int N>
struct VecN {
    T v[N];
    VecN() { for (int i = 0; i < N; ++i) v[i] = T(0); }
    VecN(T s) { for (int i = 0; i < N; ++i) v[i] = s; }
    template <int M = N, typename = typename std::enable_if<M == 2>::type>
    VecN(T x, T y) { v[0] = x; v[1] = y; }
    template <int M = N, typename = typename std::enable_if<M == 3>::type>
    VecN(T x, T y, T z) { v[0] = x; v[1] = y; v[2] = z; }

    T operator[](int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    VecN operator-() const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = -v[i]; return r; }
    VecN& operator+=(const VecN& o) { for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
    VecN& operator-=(const VecN& o) { for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
    VecN& operator*=(const VecN& o) { for (int i = 0; i < N; ++i) v[i] *= o.v[i]; return *this; }
    VecN& operator/=(const VecN& o) { for (int i = 0; i < N; ++i) v[i] /= o.v[i]; return *this; }
    const T* getPtr() const { return v; }
};

#define XRT_VEC_BINOP(op)                                                                                   \
    template <typename T, int N> inline VecN<T, N> operator op(const VecN<T, N>& a, const VecN<T, N>& b)   \
    { VecN<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }                      \
    template <typename T, int N> inline VecN<T, N> operator op(const VecN<T, N>& a, float k)               \
    { VecN<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = a.v[i] op k; return r; }                           \
    template <typename T, int N> inline VecN<T, N> operator op(float k, const VecN<T, N>& b)               \
    { VecN<T, N> r; for (int i = 0; i < N; ++i) r.v[i] = k op b.v[i]; return r; }
XRT_VEC_BINOP(+)
XRT_VEC_BINOP(-)
XRT_VEC_BINOP(*)
XRT_VEC_BINOP(/)
#undef XRT_VEC_BINOP

template <typename T> using Vec2 = VecN<T, 2>;
template <typename T> using Vec3 = VecN<T, 3>;
using Vec2f = Vec2<float>;
using Vec3f = Vec3<float>;
using Vec3ui = Vec3<uint32_t>;

template <typename T> inline Vec3<T> vmin(const Vec3<T>& a, const Vec3<T>& b)
{ return Vec3<T>(std::min(a[0], b[0]), std::min(a[1], b[1]), std::min(a[2], b[2])); }
template <typename T> inline Vec3<T> vmax(const Vec3<T>& a, const Vec3<T>& b)
{ return Vec3<T>(std::max(a[0], b[0]), std::max(a[1], b[1]), std::max(a[2], b[2])); }
template <typename T> inline T dot(const Vec3<T>& a, const Vec3<T>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b)
{ return Vec3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }

inline float length(const Vec3f& v) { return std::sqrt(dot(v, v)); }
inline float length2(const Vec3f& v) { return dot(v, v); }
inline Vec3f normalize(const Vec3f& v) { return v / length(v); } // geometry.cpp:13-16: divide, not rsqrt

// Row-major 4x4, identity by default, translation in row 3 (geometry.h:280-300).
template <typename T>
class Matrix44 {
public:
    T x[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    Matrix44() {}
    Matrix44(T a, T b, T c, T d, T e, T f, T g, T h, T i, T j, T k, T l, T m, T n, T o, T p)
    {
        const T vals[16] = {a, b, c, d, e, f, g, h, i, j, k, l, m, n, o, p};
        for (int r = 0; r < 4; ++r) for (int cc = 0; cc < 4; ++cc) x[r][cc] = vals[4 * r + cc];
    }
    const T* operator[](int i) const { return x[i]; }
    T* operator[](int i) { return x[i]; }

    template <typename S> void multVecMatrix(const Vec3<S>& src, Vec3<S>& dst) const
    {
        // point: (src,1) * M, then perspective divide (geometry.h:636-651)
        S a = src[0] * x[0][0] + src[1] * x[1][0] + src[2] * x[2][0] + x[3][0];
        S b = src[0] * x[0][1] + src[1] * x[1][1] + src[2] * x[2][1] + x[3][1];
        S c = src[0] * x[0][2] + src[1] * x[1][2] + src[2] * x[2][2] + x[3][2];
        S w = src[0] * x[0][3] + src[1] * x[1][3] + src[2] * x[2][3] + x[3][3];
        dst[0] = a / w; dst[1] = b / w; dst[2] = c / w;
    }
    template <typename S> void multDirMatrix(const Vec3<S>& src, Vec3<S>& dst) const
    {
        // direction: (src,0) * M (geometry.h:653-669)
        S a = src[0] * x[0][0] + src[1] * x[1][0] + src[2] * x[2][0];
        S b = src[0] * x[0][1] + src[1] * x[1][1] + src[2] * x[2][1];
        S c = src[0] * x[0][2] + src[1] * x[1][2] + src[2] * x[2][2];
        dst[0] = a; dst[1] = b; dst[2] = c;
    }
    friend std::ostream& operator<<(std::ostream& s, const Matrix44& m)
    {
        for (int r = 0; r < 4; ++r) s << (r ? " " : "[") << m.x[r][0] << " " << m.x[r][1] << " " << m.x[r][2] << " " << m.x[r][3] << (r == 3 ? "]" : "\n");
        return s;
    }
};
typedef Matrix44<float> Matrix44f;

template <typename S> inline Vec3<S> multVecMatrix(const Vec3<S>& src, const Matrix44<S>& m)
{ Vec3<S> d; m.multVecMatrix(src, d); return d; }
template <typename S> inline Vec3<S> multDirMatrix(const Vec3<S>& src, const Matrix44<S>& m)
{ Vec3<S> d; m.multDirMatrix(src, d); return d; }
