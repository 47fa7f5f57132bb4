// mvo_math.h — ORACLE (test infrastructure, not product code).
//
// Scalar fp32 vector/matrix helpers for the CPU restatement of the MultiVolumes HLSL hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/.
//
// Evaluation-order contract (SURVEY.md §7 hard part 3): every expression is evaluated in fp32 in
// the order written here, compiled with -ffp-contract=off, so that integer decisions derived from
// it (visibility, LOD, sample counts, OIT ordering) are reproducible bit for bit. HLSL `min16float`
// is modelled as fp32 with the source literals.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace mvo {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

inline f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
inline f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
inline f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f4 operator+(f4 a, f4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline f4 operator*(f4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float comp(const f3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// dot(): left-to-right sum of products, no contraction.
inline float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot2(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }
inline float length2(f2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
// normalize(): v * (1 / sqrt(dot(v, v))) — IEEE sqrt and divide (HLSL uses rsqrt; the restatement
// pins the correctly-rounded form so CPU and GPU agree).
inline f3 normalize3(f3 v) { const float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float lerp1(float a, float b, float t) { return a + (b - a) * t; }   // HLSL lerp: x + s(y - x)
inline float signf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
inline float fracf(float x) { return x - floorf(x); }
inline float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// Division in the image passes (OIT resolve, TAA): x / d is restated as x * rcp(d), rcp = the correctly rounded reciprocal.
// dxc's fast-math default does the same to the reference's shaders (its DXIL multiplies by reciprocals).
inline float rcp(float d) { return 1.0f / d; }
// a * b + c with one rounding, exactly where the restatement says so (the file is compiled with -ffp-contract=off)
inline float fma1(float a, float b, float c) { return std::fmaf(a, b, c); }
// pow(x, 0.25) and pow(x, 1.25) restated through correctly-rounded sqrt so CPU and GPU agree.
inline float pow025(float x) { return sqrtf(sqrtf(x)); }
inline float pow125(float x) { return x * sqrtf(sqrtf(x)); }

inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// Matrices: row-major storage, HLSL row-vector convention v' = mul(v, M).
struct m44 { float m[4][4]; };
struct m43 { float m[4][3]; };   // HLSL float4x3

inline f4 mul_p44(f3 p, const m44& M)   // mul(float4(p, 1), M)
{
    f4 r;
    r.x = ((p.x * M.m[0][0] + p.y * M.m[1][0]) + p.z * M.m[2][0]) + M.m[3][0];
    r.y = ((p.x * M.m[0][1] + p.y * M.m[1][1]) + p.z * M.m[2][1]) + M.m[3][1];
    r.z = ((p.x * M.m[0][2] + p.y * M.m[1][2]) + p.z * M.m[2][2]) + M.m[3][2];
    r.w = ((p.x * M.m[0][3] + p.y * M.m[1][3]) + p.z * M.m[2][3]) + M.m[3][3];
    return r;
}
inline f3 mul_p43(f3 p, const m43& M)   // mul(float4(p, 1), M) with M float4x3
{
    f3 r;
    r.x = ((p.x * M.m[0][0] + p.y * M.m[1][0]) + p.z * M.m[2][0]) + M.m[3][0];
    r.y = ((p.x * M.m[0][1] + p.y * M.m[1][1]) + p.z * M.m[2][1]) + M.m[3][1];
    r.z = ((p.x * M.m[0][2] + p.y * M.m[1][2]) + p.z * M.m[2][2]) + M.m[3][2];
    return r;
}
inline f3 mul_v33(f3 v, const m43& M)   // mul(v, (float3x3)M)
{
    f3 r;
    r.x = (v.x * M.m[0][0] + v.y * M.m[1][0]) + v.z * M.m[2][0];
    r.y = (v.x * M.m[0][1] + v.y * M.m[1][1]) + v.z * M.m[2][1];
    r.z = (v.x * M.m[0][2] + v.y * M.m[1][2]) + v.z * M.m[2][2];
    return r;
}

// Host-side matrix algebra (DirectXMath call sites: MultiRayCaster.cpp:325-350). Products and
// inverses are evaluated in double and rounded once to fp32, which makes them reproducible by any
// other implementation that does the same (the CUDA host code does).
inline m44 mul44(const m44& A, const m44& B)
{
    m44 R;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += (double)A.m[i][k] * (double)B.m[k][j];
            R.m[i][j] = (float)s;
        }
    return R;
}
inline m44 inverse44(const m44& A)
{
    double a[4][4], inv[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = A.m[i][j];
    // cofactor expansion
    auto det3 = [&](int r0, int r1, int r2, int c0, int c1, int c2) {
        return a[r0][c0] * (a[r1][c1] * a[r2][c2] - a[r1][c2] * a[r2][c1])
             - a[r0][c1] * (a[r1][c0] * a[r2][c2] - a[r1][c2] * a[r2][c0])
             + a[r0][c2] * (a[r1][c0] * a[r2][c1] - a[r1][c1] * a[r2][c0]);
    };
    double cof[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int r[3], c[3], ri = 0, ci = 0;
            for (int k = 0; k < 4; ++k) { if (k != i) r[ri++] = k; if (k != j) c[ci++] = k; }
            const double d = det3(r[0], r[1], r[2], c[0], c[1], c[2]);
            cof[i][j] = ((i + j) & 1) ? -d : d;
        }
    const double det = a[0][0] * cof[0][0] + a[0][1] * cof[0][1] + a[0][2] * cof[0][2] + a[0][3] * cof[0][3];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i][j] = cof[j][i] / det;
    m44 R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.m[i][j] = (float)inv[i][j];
    return R;
}
inline m44 from43(const m43& W)
{
    m44 R;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 3; ++j) R.m[i][j] = W.m[i][j]; R.m[i][3] = (i == 3) ? 1.0f : 0.0f; }
    return R;
}
inline m43 to43(const m44& M)
{
    m43 R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = M.m[i][j];
    return R;
}

// ---- fp16 / R11G11B10F conversions (storage formats) ----
// float -> half, round-to-nearest-even, overflow to inf, denormals kept.
inline uint16_t f32_to_f16(float f)
{
    const uint32_t x = as_uint(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));   // inf / nan
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);   // >= 65520 rounds to inf
    if (ax < 0x33000001u) return (uint16_t)sign;                // <= 2^-25 rounds to zero
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }          // half denormal
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    return (uint16_t)(sign | (base + q));   // mantissa carry propagates into the exponent correctly
}

struct HalfTable {
    float t[65536];
    HalfTable()
    {
        for (uint32_t h = 0; h < 65536; ++h) {
            const uint32_t sign = (h & 0x8000u) << 16;
            const uint32_t e = (h >> 10) & 0x1fu;
            const uint32_t m = h & 0x3ffu;
            float v;
            if (e == 0) v = ldexpf((float)m, -24);
            else if (e == 31) v = m ? NAN : INFINITY;
            else v = ldexpf((float)(m | 0x400u), (int)e - 25);
            t[h] = as_float(as_uint(v) | sign);
        }
    }
};
inline const float* half_table() { static HalfTable T; return T.t; }
inline float f16_to_f32(uint16_t h) { return half_table()[h]; }

// R11G11B10_FLOAT channel quantisation: unsigned small floats with 5 exponent bits (bias 15) and
// 6 (R, G) or 5 (B) mantissa bits. Every such value is exactly representable in fp16, so a light
// map stored as RGBA16F holding quantised values is texel-identical to the reference's format
// (MultiRayCaster.cpp:123-125). Rounding: nearest-even (DirectXMath XMStoreFloat3PK convention);
// negatives and NaN -> 0, overflow -> max finite (65024 / 64512).
inline float quantize_ufloat(float f, int mant_bits)
{
    if (!(f > 0.0f)) return 0.0f;
    // Rounded directly from the fp32 pattern (going through fp16 first would double-round).
    const uint32_t x = as_uint(f);
    const float max_finite = (mant_bits == 6) ? 65024.0f : 64512.0f;
    if (x >= 0x7f800000u) return max_finite;
    int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift = 23 - mant_bits;
    if (e < -14) shift += (-14 - e);
    if (shift > 24) return 0.0f;
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    const int e_eff = (e < -14) ? -14 : e;
    const float v = ldexpf((float)q, e_eff - mant_bits);
    return v > max_finite ? max_finite : v;
}

} // namespace mvo
