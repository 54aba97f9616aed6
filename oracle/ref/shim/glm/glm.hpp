// glm/glm.hpp — SHIM. TEST INFRASTRUCTURE ONLY (oracle/ref: the reference's own sources compiled for the parity anchor).
//
// The reference uses glm 0.9.9.8 (external/CMakeLists.txt:29-33 fetches it from GitHub; it is not vendored under
// /root/reference and there is no network), so its sources cannot see the real headers here. This file RESTATES the
// subset of glm the reference's render path uses — float vectors/matrices and a dozen functions — with glm 0.9.9.8's
// published (pure C++, non-SIMD) operation order, one IEEE rounding per written operation when compiled with
// -ffp-contract=off:
//   dot(vec3)      tmp = a*b; (tmp.x + tmp.y) + tmp.z                      (glm/detail/func_geometric.inl compute_dot)
//   normalize      v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)   (func_geometric.inl, func_exponential.inl)
//   length         sqrt(dot(v,v)); distance(a,b) = length(b - a)
//   reflect        I - N * dot(N,I) * 2
//   cross          (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   min/max/clamp  min(x,y) = (y < x) ? y : x; max(x,y) = (x < y) ? y : x; clamp = min(max(x,lo),hi)
//   mat4 * vec4    (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)           (type_mat4x4.inl, scalar path)
//   mat3 * vec3    m[0][i]*v.x + m[1][i]*v.y + m[2][i]*v.z, left to right  (type_mat3x3.inl)
//   inverse(mat4)  the cofactor expansion of func_matrix.inl compute_inverse<4,4>
//   translate / rotate / radians                                            (ext/matrix_transform.inl, func_trigonometric.inl)
// Nothing in the crender_b200 package includes this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cstring>
#include <limits>

namespace glm
{
    template<typename T>
    struct tvec2
    {
        union { T x, r, s; };
        union { T y, g, t; };
        constexpr tvec2() : x(0), y(0) {}
        template<typename A, typename B>
        constexpr tvec2(A a, B b) : x(T(a)), y(T(b)) {}
        template<typename A>
        constexpr explicit tvec2(A a) : x(T(a)), y(T(a)) {}
        template<typename U>
        constexpr tvec2(const tvec2<U> &o) : x(T(o.x)), y(T(o.y)) {}
        T       &operator[](int i) { return i == 0 ? x : y; }
        const T &operator[](int i) const { return i == 0 ? x : y; }
    };
    template<typename T>
    struct tvec4;
    template<typename T>
    struct tvec3
    {
        union { T x, r, s; };
        union { T y, g, t; };
        union { T z, b, p; };
        constexpr tvec3() : x(0), y(0), z(0) {}
        template<typename A, typename B, typename C>
        constexpr tvec3(A a, B b_, C c) : x(T(a)), y(T(b_)), z(T(c)) {}
        template<typename A>
        constexpr explicit tvec3(A a) : x(T(a)), y(T(a)), z(T(a)) {}
        template<typename U>
        constexpr tvec3(const tvec3<U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
        template<typename U>
        constexpr tvec3(const tvec4<U> &o);
        T       &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
        const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
        tvec3   &operator+=(const tvec3 &o) { x += o.x, y += o.y, z += o.z; return *this; }
        tvec3   &operator-=(const tvec3 &o) { x -= o.x, y -= o.y, z -= o.z; return *this; }
        tvec3   &operator*=(const tvec3 &o) { x *= o.x, y *= o.y, z *= o.z; return *this; }
        tvec3   &operator*=(T s_) { x *= s_, y *= s_, z *= s_; return *this; }
    };
    template<typename T>
    struct tvec4
    {
        union { T x, r, s; };
        union { T y, g, t; };
        union { T z, b, p; };
        union { T w, a, q; };
        constexpr tvec4() : x(0), y(0), z(0), w(0) {}
        template<typename A, typename B, typename C, typename D>
        constexpr tvec4(A a_, B b_, C c, D d) : x(T(a_)), y(T(b_)), z(T(c)), w(T(d)) {}
        template<typename A>
        constexpr explicit tvec4(A a_) : x(T(a_)), y(T(a_)), z(T(a_)), w(T(a_)) {}
        template<typename U, typename D>
        constexpr tvec4(const tvec3<U> &o, D d) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(d)) {}
        T       &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
        const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    };
    template<typename T>
    template<typename U>
    constexpr tvec3<T>::tvec3(const tvec4<U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z))
    {
    }

    using vec2  = tvec2<float>;
    using vec3  = tvec3<float>;
    using vec4  = tvec4<float>;
    using ivec2 = tvec2<int>;
    using ivec3 = tvec3<int>;

    // ---- vec3 arithmetic (component-wise, one operation each)
    inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
    inline vec3 operator+(const vec3 &a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
    inline vec3 operator+(float s, const vec3 &a) { return vec3(s + a.x, s + a.y, s + a.z); }
    inline vec3 operator-(const vec3 &a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
    inline vec3 operator-(float s, const vec3 &a) { return vec3(s - a.x, s - a.y, s - a.z); }
    inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
    inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
    inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
    inline vec3 operator/(float s, const vec3 &a) { return vec3(s / a.x, s / a.y, s / a.z); }
    inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
    inline bool operator==(const vec3 &a, const vec3 &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
    // ---- vec2 / vec4
    inline vec2 operator+(const vec2 &a, const vec2 &b) { return vec2(a.x + b.x, a.y + b.y); }
    inline vec2 operator-(const vec2 &a, const vec2 &b) { return vec2(a.x - b.x, a.y - b.y); }
    inline vec2 operator*(const vec2 &a, float s) { return vec2(a.x * s, a.y * s); }
    inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
    inline vec4 operator-(const vec4 &a, const vec4 &b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    inline vec4 operator*(const vec4 &a, const vec4 &b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
    inline vec4 operator*(const vec4 &a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
    inline vec4 operator*(float s, const vec4 &a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
    inline vec4 operator/(const vec4 &a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
    inline bool operator==(const vec4 &a, const vec4 &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

    // ---- scalar functions
    inline float min(float x, float y) { return (y < x) ? y : x; }
    inline float max(float x, float y) { return (x < y) ? y : x; }
    inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
    inline float abs(float x) { return std::fabs(x); }
    inline float sqrt(float x) { return std::sqrt(x); }
    inline float pow(float x, float y) { return std::pow(x, y); }
    inline float sin(float x) { return std::sin(x); }
    inline float cos(float x) { return std::cos(x); }
    inline float tan(float x) { return std::tan(x); }
    inline float acos(float x) { return std::acos(x); }
    inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
    inline constexpr float radians(float degrees) { return degrees * 0.01745329251994329576923690768489f; }

    // ---- geometric
    inline float dot(const vec3 &a, const vec3 &b)
    {
        const vec3 tmp(a * b);
        return tmp.x + tmp.y + tmp.z;
    }
    inline float dot(const vec4 &a, const vec4 &b)
    {
        const vec4 tmp(a * b);
        return (tmp.x + tmp.y) + (tmp.z + tmp.w);
    }
    inline vec3 cross(const vec3 &x, const vec3 &y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
    inline float length(const vec3 &v) { return sqrt(dot(v, v)); }
    inline float distance(const vec3 &p0, const vec3 &p1) { return length(p1 - p0); }
    inline vec3  normalize(const vec3 &v) { return v * inversesqrt(dot(v, v)); }
    inline vec4  normalize(const vec4 &v) { return v * inversesqrt(dot(v, v)); }
    inline vec3  reflect(const vec3 &I, const vec3 &N) { return I - N * dot(N, I) * 2.0f; }

    // ---- matrices, column-major: m[column][row]
    struct mat3
    {
        vec3 c[3];
        mat3() : c { vec3(1, 0, 0), vec3(0, 1, 0), vec3(0, 0, 1) } {}
        explicit mat3(float d) : c { vec3(d, 0, 0), vec3(0, d, 0), vec3(0, 0, d) } {}
        mat3(const vec3 &a, const vec3 &b, const vec3 &cc) : c { a, b, cc } {}
        vec3       &operator[](int i) { return c[i]; }
        const vec3 &operator[](int i) const { return c[i]; }
    };
    inline vec3 operator*(const mat3 &m, const vec3 &v)
    {
        return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z, m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
    }
    struct mat4
    {
        vec4 c[4];
        mat4() : c { vec4(1, 0, 0, 0), vec4(0, 1, 0, 0), vec4(0, 0, 1, 0), vec4(0, 0, 0, 1) } {}
        template<typename A>
        explicit mat4(A d) : c { vec4(float(d), 0, 0, 0), vec4(0, float(d), 0, 0), vec4(0, 0, float(d), 0), vec4(0, 0, 0, float(d)) }
        {
        }
        mat4(const vec4 &a, const vec4 &b, const vec4 &cc, const vec4 &d) : c { a, b, cc, d } {}
        vec4       &operator[](int i) { return c[i]; }
        const vec4 &operator[](int i) const { return c[i]; }
    };
    inline vec4 operator*(const mat4 &m, const vec4 &v)
    {
        const vec4 Mul0 = m[0] * vec4(v.x), Mul1 = m[1] * vec4(v.y);
        const vec4 Add0 = Mul0 + Mul1;
        const vec4 Mul2 = m[2] * vec4(v.z), Mul3 = m[3] * vec4(v.w);
        const vec4 Add1 = Mul2 + Mul3;
        return Add0 + Add1;
    }
    inline mat4 operator*(const mat4 &m1, const mat4 &m2)
    {
        mat4 r;
        for (int j = 0; j < 4; j++) r[j] = m1[0] * m2[j][0] + m1[1] * m2[j][1] + m1[2] * m2[j][2] + m1[3] * m2[j][3];
        return r;
    }
    inline mat4 inverse(const mat4 &m)
    {
        const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
        const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
        const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
        const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
        const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
        const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
        const vec4  Fac0(Coef00, Coef00, Coef02, Coef03), Fac1(Coef04, Coef04, Coef06, Coef07), Fac2(Coef08, Coef08, Coef10, Coef11);
        const vec4  Fac3(Coef12, Coef12, Coef14, Coef15), Fac4(Coef16, Coef16, Coef18, Coef19), Fac5(Coef20, Coef20, Coef22, Coef23);
        const vec4  Vec0(m[1][0], m[0][0], m[0][0], m[0][0]), Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
        const vec4  Vec2(m[1][2], m[0][2], m[0][2], m[0][2]), Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
        const vec4  Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2), Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
        const vec4  Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5), Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
        const vec4  SignA(+1, -1, +1, -1), SignB(-1, +1, -1, +1);
        const mat4  Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
        const vec4  Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
        const vec4  Dot0(m[0] * Row0);
        const float Dot1           = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
        const float OneOverDeterminant = 1.0f / Dot1;
        return mat4(Inverse[0] * OneOverDeterminant, Inverse[1] * OneOverDeterminant, Inverse[2] * OneOverDeterminant, Inverse[3] * OneOverDeterminant);
    }
    inline mat4 translate(const mat4 &m, const vec3 &v)
    {
        mat4 Result(m);
        Result[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
        return Result;
    }
    inline mat4 rotate(const mat4 &m, float angle, const vec3 &v)
    {
        const float a = angle, c = cos(a), s = sin(a);
        const vec3  axis(normalize(v));
        const vec3  temp((1.0f - c) * axis);
        mat4        Rotate;
        Rotate[0][0] = c + temp[0] * axis[0], Rotate[0][1] = temp[0] * axis[1] + s * axis[2], Rotate[0][2] = temp[0] * axis[2] - s * axis[1];
        Rotate[1][0] = temp[1] * axis[0] - s * axis[2], Rotate[1][1] = c + temp[1] * axis[1], Rotate[1][2] = temp[1] * axis[2] + s * axis[0];
        Rotate[2][0] = temp[2] * axis[0] + s * axis[1], Rotate[2][1] = temp[2] * axis[1] - s * axis[0], Rotate[2][2] = c + temp[2] * axis[2];
        mat4 Result;
        Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
        Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
        Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
        Result[3] = m[3];
        return Result;
    }
    template<typename M>
    inline const float *value_ptr(const M &m)
    {
        return &m[0][0];
    }
}    // namespace glm
