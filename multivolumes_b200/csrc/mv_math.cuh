// mv_math.cuh — fp32 vector / matrix / storage-format helpers shared by the host code and the
// sm_100a kernels of libmv_b200.so.
//
// Evaluation-order contract. HLSL leaves the association of dot(), mul(), normalize() to the shader
// compiler (dxc runs fast-math; the shipped CSVolumeCull.cso is itself not a literal evaluation of
// its source). This library fixes ONE order for each of them — written out below —
// and is compiled with --fmad=false, IEEE division and IEEE square root, so that every integer
// decision derived from fp32 geometry (visibility, face masks, LOD, sample counts, OIT layer order)
// is reproducible bit for bit by any implementation that states the same order (the test oracle does).
// Where a fused multiply-add is wanted it is written explicitly as fmaf().
//
// HLSL `min16float` is computed in fp32 with the source literals (MultiVolumes.vcxproj does not pass
// -enable-16bit-types, so min-precision is only a hint).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>

#define MV_HD __host__ __device__ __forceinline__
#define MV_D __device__ __forceinline__

namespace mv {

// SharedConsts.h:5-10, Common.hlsli:12, RayMarch.hlsli:11-12
constexpr uint32_t kGroupVolumeCount = 4;
constexpr uint32_t kNumCubeMip = 5;
constexpr uint32_t kNumOitLayers = 8;
constexpr float kZNear = 1.0f, kZFar = 1000.0f;
constexpr uint32_t kCubeMapRayMarchBit = 1u << 15;
// `min16float` literals: the values the reference's SHIPPED shaders hold (Bin/*.cso, DXIL: dxc folds a min16float literal to
// binary16), not the decimal text of the HLSL. Read off the disassembly (oracle/dxil/container.py) and confirmed by executing
// those shaders (oracle/dxil/interp.py): with these values the oracle reproduces CSRayMarchV.cso and CSRayMarchL.cso bit for
// bit; with the source decimals it does not.
constexpr float kAbsorption = 0.7998046875f;            // ABSORPTION 0.8        -> 0xH3A66
constexpr float kZeroThreshold = 0.01000213623046875f;  // ZERO_THRESHOLD 0.01   -> 0xH211F
constexpr float kMaxDist = 3.46484375f;                 // g_maxDist 2 sqrt(3)   -> 0xH42EE
constexpr float kInvTwoPi = 0.1591796875f;              // "/ (2.0 * PI)"        -> fmul by 0xH3118
constexpr float kAlphaClamp = 0.99951171875f;           // 0.9997                -> 0xH3BFF (PSResolveOIT.cso)
constexpr float kNinth = 0.111083984375f;               // "/ 9.0"               -> fmul by 0xH2F1C (CSTemporalAA.cso)
constexpr float kToneScale = 1.0498046875f;             // 1.05                  -> 0xH3C33 (PSToneMap.cso)
constexpr float kToneBias = 0.7001953125f;              // 0.7                   -> 0xH399A
constexpr float kFltMax = 3.402823466e+38f;
constexpr float kPi = 3.1415926535897f;   // SHIrradiance.hlsli:6

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

MV_HD V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
MV_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
MV_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
MV_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
MV_HD V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
MV_HD V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
MV_HD V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
MV_HD V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
MV_HD V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
MV_HD V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
MV_HD float comp(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// dot: products summed left to right
MV_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
MV_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
MV_HD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
// normalize: v * (1 / sqrt(dot(v, v))) with correctly rounded sqrt and divide
MV_HD V3 normalize(V3 v) { const float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
MV_HD float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
MV_HD float lerp(float a, float b, float t) { return a + (b - a) * t; }   // HLSL lerp: x + s(y - x)
MV_HD float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
MV_HD float frac(float x) { return x - floorf(x); }
MV_HD float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
MV_HD float clamp1(float x) { return fminf(fmaxf(x, -1.0f), 1.0f); }
// Division, where the image passes (OIT resolve, TAA) divide: x / d is evaluated as x * rcp(d) with rcp the CORRECTLY ROUNDED
// reciprocal — reproducible by any IEEE implementation as 1.0f / d — which is also what dxc's fast-math default turns the
// reference's divisions into (the shipped DXIL multiplies by reciprocals). Half the cost of an IEEE divide.
MV_HD float rcp(float d)
{
#ifdef __CUDA_ARCH__
    return __frcp_rn(d);
#else
    return 1.0f / d;
#endif
}
// a * b + c with ONE rounding, where the image passes fuse (written explicitly: the library is compiled without contraction)
MV_HD float fma1(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
// Texel coordinate of the normalised coordinate u on an n-texel axis (n <= 16384) in fixed point with 8 fractional bits, rounded
// to nearest, clamp addressing: u * n - 0.5 on the 1/256 grid (D3D11.3 functional spec 7.18.8: linear filtering uses fixed-point
// texel coordinates with at least 8 fractional bits).
MV_HD int tex_coord_q8(float u, int n)
{
    float uc = u < -1.0f ? -1.0f : (u > 2.0f ? 2.0f : u);
    if (!(uc == uc)) uc = 0.0f;
    const int q = (int)floorf(fma1(fma1(uc, (float)n, -0.5f), 256.0f, 0.5f));
    const int hi = (n - 1) * 256;
    return q < 0 ? 0 : (q > hi ? hi : q);
}
// One axis of a linear filter with such a weight: a tap of weight zero does not contribute (it may hold inf / NaN)
MV_HD float lerp_q8(float a, float b, float w) { return w == 0.0f ? a : fma1(b - a, w, a); }
// pow(x, 0.25), pow(x, 1.25) through correctly rounded square roots
MV_HD float pow025(float x) { return sqrtf(sqrtf(x)); }
MV_HD float pow125(float x) { return x * sqrtf(sqrtf(x)); }

MV_HD uint32_t as_uint(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
MV_HD float as_float(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// Matrices are row-major float arrays in the reference's row-vector convention:
// mul(float4(p, 1), M) with M 4x4 (16 floats) or 4x3 (12 floats, HLSL float4x3).
MV_HD V4 mul_p44(V3 p, const float* M)
{
    V4 r;
    r.x = ((p.x * M[0] + p.y * M[4]) + p.z * M[8]) + M[12];
    r.y = ((p.x * M[1] + p.y * M[5]) + p.z * M[9]) + M[13];
    r.z = ((p.x * M[2] + p.y * M[6]) + p.z * M[10]) + M[14];
    r.w = ((p.x * M[3] + p.y * M[7]) + p.z * M[11]) + M[15];
    return r;
}
MV_HD V3 mul_p43(V3 p, const float* M)
{
    V3 r;
    r.x = ((p.x * M[0] + p.y * M[3]) + p.z * M[6]) + M[9];
    r.y = ((p.x * M[1] + p.y * M[4]) + p.z * M[7]) + M[10];
    r.z = ((p.x * M[2] + p.y * M[5]) + p.z * M[8]) + M[11];
    return r;
}
MV_HD V3 mul_v33(V3 v, const float* M)   // mul(v, (float3x3)M) with M a 4x3
{
    V3 r;
    r.x = (v.x * M[0] + v.y * M[3]) + v.z * M[6];
    r.y = (v.x * M[1] + v.y * M[4]) + v.z * M[7];
    r.z = (v.x * M[2] + v.y * M[5]) + v.z * M[8];
    return r;
}

// ---- storage formats ----
// fp32 -> binary16, round to nearest even, overflow to infinity, denormals kept, NaN -> quiet NaN
// with the sign kept.
MV_HD uint16_t f32_to_f16(float f)
{
#ifdef __CUDA_ARCH__
    if (f != f) return (uint16_t)(((__float_as_uint(f) >> 16) & 0x8000u) | 0x7e00u);
    return __half_as_ushort(__float2half_rn(f));
#else
    const uint32_t x = as_uint(f);
    const uint32_t sgn = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sgn | 0x7e00u);
    if (ax >= 0x477ff000u) return (uint16_t)(sgn | 0x7c00u);
    if (ax < 0x33000001u) return (uint16_t)sgn;
    const int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift = 13;
    uint32_t base = 0;
    if (e < -14) shift += -14 - e; else { base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    return (uint16_t)(sgn | (base + q));
#endif
}
MV_HD float f16_to_f32(uint16_t h)
{
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half(h));
#else
    const uint32_t sgn = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0) { return as_float(as_uint((float)m * 5.9604644775390625e-8f) | sgn); }
    if (e == 31) return as_float(sgn | 0x7f800000u | (m << 13));
    return as_float(sgn | ((e + 112u) << 23) | (m << 13));
#endif
}
MV_HD uint2 pack_half4(V4 v)
{
    uint2 r;
    r.x = (uint32_t)f32_to_f16(v.x) | ((uint32_t)f32_to_f16(v.y) << 16);
    r.y = (uint32_t)f32_to_f16(v.z) | ((uint32_t)f32_to_f16(v.w) << 16);
    return r;
}
MV_HD V4 unpack_half4(uint2 p)
{
    return {f16_to_f32((uint16_t)(p.x & 0xffffu)), f16_to_f32((uint16_t)(p.x >> 16)),
            f16_to_f32((uint16_t)(p.y & 0xffffu)), f16_to_f32((uint16_t)(p.y >> 16))};
}

// One channel of DXGI_FORMAT_R11G11B10_FLOAT (the reference's light-map format,
// MultiRayCaster.cpp:123-125): unsigned, 5 exponent bits (bias 15), 6 (R, G) or 5 (B) mantissa bits.
// Round to nearest even straight from the fp32 pattern; negatives and NaN -> 0; overflow -> largest
// finite. Every result is exactly representable in binary16, so an RGBA16F light map holding these
// values is texel-identical to the reference's.
MV_HD float quantize_ufloat(float f, int mant_bits)
{
    if (!(f > 0.0f)) return 0.0f;
    const uint32_t x = as_uint(f);
    const float max_finite = (mant_bits == 6) ? 65024.0f : 64512.0f;
    if (x >= 0x7f800000u) return max_finite;
    const int e = (int)(x >> 23) - 127;
    const uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift = 23 - mant_bits;
    if (e < -14) shift += -14 - e;
    if (shift > 24) return 0.0f;
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    const int e_eff = (e < -14) ? -14 : e;
    // q * 2^(e_eff - mant_bits): q < 2^8 and the exponent is in the normal fp32 range, so this is exact
    const float v = (float)q * as_float((uint32_t)(e_eff - mant_bits + 127) << 23);
    return v > max_finite ? max_finite : v;
}

// HLSL uint(f): NaN and negatives -> 0, saturating
MV_HD uint32_t float_to_uint_sat(float f)
{
    if (!(f == f) || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}
// uint(max(log2(x), 0)) (VolumeCull.hlsli:288): floor(log2 x) for x >= 1 is the biased exponent
// minus 127, read from the bit pattern (exact); anything below 1 and NaN give 0.
MV_HD uint32_t floor_log2_clamped(float x)
{
    if (!(x >= 1.0f)) return 0u;
    return ((as_uint(x) >> 23) & 0xffu) - 127u;
}

} // namespace mv
