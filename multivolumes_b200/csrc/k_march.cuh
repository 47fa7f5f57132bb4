// k_march.cuh — device functions shared by the light march, the view march and the direct
// screen-space ray cast: RayMarch.hlsli of the reference restated for sm_100a.
//
// Volumes (RGBA16F, or R16F density alone) and light maps (RGBA16F) are CUDA 3-D arrays sampled through texture objects with
// normalised coordinates, clamp addressing and hardware trilinear filtering — the LINEAR_CLAMP
// sampler of the reference (MultiRayCaster.cpp:556-560).
#pragma once
#include "mv_internal.h"

namespace mv {

MV_D V4 sample_volume(cudaTextureObject_t tex, V3 uvw)   // GetSample, RayMarch.hlsli:44-50
{
    const float4 c = tex3D<float4>(tex, uvw.x, uvw.y, uvw.z);
    return {c.x, c.y, c.z, c.w};
}

// a texture fetch the compiler may not move (asm volatile): used to put the two fetches of one march step in flight together
MV_D float4 tex3d_issue(cudaTextureObject_t tex, float x, float y, float z)
{
    float4 r;
    asm volatile("tex.3d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(tex), "f"(x), "f"(y), "f"(z));
    return r;
}

// The density of a fetched volume texel: alpha of an RGBA16F volume, the only channel of an R16F one
// (MV_FLAG_DENSITY_ONLY; `densityOnly` is uniform over a launch).
MV_D float texel_density(float4 c, bool densityOnly) { return densityOnly ? c.x : c.w; }
MV_D float fetch_density(cudaTextureObject_t tex, float x, float y, float z, bool densityOnly)
{
    return texel_density(tex3D<float4>(tex, x, y, z), densityOnly);
}

// Store of one volume texel by the ingest kernels: RGBA16F, or the alpha channel alone into an R16F volume
MV_D void store_volume_texel(cudaSurfaceObject_t surf, uint32_t x, uint32_t y, uint32_t z, V4 rgba, bool densityOnly)
{
    if (densityOnly) surf3Dwrite((unsigned short)f32_to_f16(rgba.w), surf, (int)(x * 2), (int)y, (int)z);
    else surf3Dwrite(pack_half4(rgba), surf, (int)(x * 8), (int)y, (int)z);
}

MV_D V3 local_to_tex3d(V3 pos)   // LocalToTex3DSpace, RayMarch.hlsli:170-177
{
    return {pos.x * 0.5f + 0.5f, pos.y * 0.5f + 0.5f, pos.z * 0.5f + 0.5f};
}

MV_D bool inside_unit_box(V3 p) { return fabsf(p.x) <= 1.0f && fabsf(p.y) <= 1.0f && fabsf(p.z) <= 1.0f; }
MV_D bool outside_unit_box(V3 p) { return fabsf(p.x) > 1.0f || fabsf(p.y) > 1.0f || fabsf(p.z) > 1.0f; }

// ComputeRayOrigin, RayMarch.hlsli:128-155: an eye outside the box is moved to the entry point
// (slab test per axis, nearest non-negative hit), then clamped to the box; false on a miss.
MV_D bool compute_ray_origin(V3& rayOrigin, V3 rayDir)
{
    if (inside_unit_box(rayOrigin)) return true;
    float U = kFltMax;
    bool isHit = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        const float d = comp(rayDir, i), o = comp(rayOrigin, i);
        const float u = (-sign(d) - o) / d;
        if (u < 0.0f) continue;
        if (fabsf(comp(rayDir, j) * u + comp(rayOrigin, j)) > 1.0f) continue;
        if (fabsf(comp(rayDir, k) * u + comp(rayOrigin, k)) > 1.0f) continue;
        if (u < U) { U = u; isHit = true; }
    }
    rayOrigin = {clamp1(rayDir.x * U + rayOrigin.x), clamp1(rayDir.y * U + rayOrigin.y), clamp1(rayDir.z * U + rayOrigin.z)};
    return isHit;
}

// GetStep, RayMarch.hlsli:182-192
// factorTh = 1 - transm is handed in: in the view march and RayCast transm is itself 1 - scatter.w, and the compiled shaders
// use scatter.w there (dxc folds 1 - (1 - x) to x, which is not an identity in fp32); the product is associated as compiled:
// ((factorTh * 1.5) * factorEv) * factorUi (CSRayMarchV.cso %371-%375, CSRayMarchL.cso %510-%515).
MV_D float get_step(float dDensity, float factorTh, float density, float step)
{
    const float factorEv = fminf(1.0f / 256.0f / fabsf(dDensity), 2.0f);
    const float factorUi = fminf(1.0f - density, 1.0f);
    return fmaxf(((factorTh * 1.5f) * factorEv) * factorUi, 1.0f) * step;
}

// GetTMax, RayMarch.hlsli:82-92: ray parameter at which the scene depth occludes the ray
MV_D float get_tmax(V3 clipPos, V3 rayOrigin, V3 rayDir, const float* wvpi)
{
    if (clipPos.z >= 1.0f) return kFltMax;
    const V4 h = mul_p44(clipPos, wvpi);
    const V3 p = {h.x / h.w, h.y / h.w, h.z / h.w};
    const V3 t = (p - rayOrigin) / rayDir;
    return max3(t.x, t.y, t.z);
}

// Conservative early-out for the ray / unit-box tests: true only when the ray o + u d (u >= 0, d not
// normalised) certainly misses the box, because it stays outside the box's bounding sphere (radius
// sqrt(3)) by a margin far above the fp32 error of these few operations. Whenever it returns true the
// exact slab test (compute_ray_origin, the OIT exit test) would have reported a miss as well, so using
// it changes no result; NaNs fall through to the exact test.
MV_D bool ray_misses_box_for_sure(V3 o, V3 d)
{
    const float oo = dot(o, o);
    const float slack = 3.2f + 1.0e-4f * oo;      // sphere radius^2 + absolute and relative safety margins
    if (!(oo > slack)) return false;              // origin inside the (inflated) sphere
    const float tca = -dot(o, d);                 // projection of (centre - o) on d, times |d|
    if (tca < 0.0f) return true;                  // sphere entirely behind the ray
    return (oo - slack) * dot(d, d) > tca * tca;  // closest approach of the line outside the sphere
}

struct MarchCount { uint32_t samples, lightFetches, skipped; };

// Is the brick that holds the sample position known to be empty (Occupancy, mv_internal.h)? `bits` are the bricks of
// the ray's source volume, `pos` the sample position in the volume's local space ([-1, 1]^3). The sample's texel coordinate
// is (pos / 2 + 1/2) G; the texels its trilinear footprint touches lie within half a texel of it, inside the one-texel
// border the brick's bit accounts for — which also covers the rounding of this lookup (it is not part of the stated
// arithmetic: any conservative test gives the same results).
MV_D bool brick_is_empty(const uint32_t* __restrict__ bits, const Occupancy& occ, V3 pos)
{
    const int top = (int)occ.bricks - 1;
    const int bx = min((int)fmaf(pos.x, occ.halfBricks, occ.halfBricks), top);
    const int by = min((int)fmaf(pos.y, occ.halfBricks, occ.halfBricks), top);
    const int bz = min((int)fmaf(pos.z, occ.halfBricks, occ.halfBricks), top);
    const uint32_t b = ((uint32_t)bz * occ.bricks + (uint32_t)by) * occ.bricks + (uint32_t)bx;
    return (__ldg(bits + (b >> 5)) >> (b & 31)) & 1u;
}

// The per-ray loop of CSRayMarch.hlsl:112-155 and RayCast.hlsli:57-105 (identical bodies).
// One trilinear density fetch per step, one trilinear light-map fetch when the sample is non-empty,
// adaptive step from the density change, front-to-back accumulation, early out at transmittance < 0.01.
// `emptyBits` (nullptr = none): the empty-space bricks of the source volume. A sample inside a brick known to be empty is an
// empty sample (density <= ZERO_THRESHOLD) whatever its exact value: it advances t by the base step and changes nothing
// else, so its fetch is left out. The bricks are consulted only after an empty sample — inside a dense run the lookup
// would lengthen every step's dependent chain for nothing — so the first sample of an empty run is still fetched.
MV_D V4 march_ray(cudaTextureObject_t grid, cudaTextureObject_t light, uint32_t smpCount, V3 rayOrigin, V3 rayDir,
                  float tMax, bool densityOnly, MarchCount& mc, const uint32_t* __restrict__ emptyBits, const Occupancy& occ)
{
    const float stepScale = kMaxDist / (float)smpCount;  // g_maxDist, RayMarch.hlsli:17
    V4 scatter = {0.0f, 0.0f, 0.0f, 0.0f};
    float t = 0.0f;
    float prevDensity = 0.0f;
    // The loop is bound by its chain of dependent texture round trips. The light-map texel of a step does not depend on
    // the density fetched at that step, only its use does; dense and empty samples come in runs, so when the previous
    // sample was dense the light fetch is issued together with the density fetch (one round trip per step instead of
    // two) and its value is dropped if the sample turns out empty. No result changes.
    // (Measured and not adopted: always issuing it, +35 % texture requests, slower; also fetching the next sample of an
    // empty run, whose position is known in advance, in the same round trip: no gain. profiles/r01_notes.md)
    bool wasDense = false;
    for (uint32_t i = 0; i < smpCount; ++i) {
        const V3 pos = {rayOrigin.x + rayDir.x * t, rayOrigin.y + rayDir.y * t, rayOrigin.z + rayDir.z * t};
        if (outside_unit_box(pos)) break;
        // One step per iteration for every lane, skipped or fetched: a lane that ran through its empty samples in an inner
        // loop of its own would do so while the lanes that need a fetch sit masked off (measured: view march +39 %).
        if (emptyBits && !wasDense && brick_is_empty(emptyBits, occ, pos)) {
            ++mc.samples; ++mc.skipped;
            t += stepScale;
            if (t > tMax) break;
            continue;
        }
        const V3 uvw = local_to_tex3d(pos);
        const float4 c4 = tex3d_issue(grid, uvw.x, uvw.y, uvw.z);
        float4 l = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (wasDense) l = tex3d_issue(light, uvw.x, uvw.y, uvw.z);
        V4 color = {c4.x, c4.y, c4.z, c4.w};
        if (densityOnly) color = {1.0f, 1.0f, 1.0f, c4.x};
        ++mc.samples;
        float newStep = stepScale;
        if (color.w > kZeroThreshold) {                  // skip empty space
            if (!wasDense) l = tex3d_issue(light, uvw.x, uvw.y, uvw.z);   // GetLight, RayMarch.hlsli:235-240
            wasDense = true;
            ++mc.lightFetches;
            const float transm = 1.0f - scatter.w;
            const float dDensity = color.w - prevDensity;
            newStep = get_step(dDensity, scatter.w, color.w, stepScale);
            prevDensity = color.w;
            // colour (not pre-multiplied) x density x light x ABSORPTION x transmittance, associated as the compiled shaders
            // have it (CSRayMarchV.cso, PSCube.cso): ((transm * A) * a) once, then * colour * light per channel
            const float ka = (transm * kAbsorption) * color.w;
            scatter.x += (ka * color.x) * l.x;
            scatter.y += (ka * color.y) * l.y;
            scatter.z += (ka * color.z) * l.z;
            scatter.w += ka;
            if (transm < kZeroThreshold) break;
        }
        else wasDense = false;
        t += newStep;
        if (t > tMax) break;
    }
    scatter.x *= kInvTwoPi; scatter.y *= kInvTwoPi; scatter.z *= kInvTwoPi;   // CSRayMarch.hlsl:155, as compiled
    return scatter;
}

// CastLightRay, RayMarch.hlsli:197-230 (mip 0): transmittance toward the light / along the AO direction
MV_D void cast_light_ray(float& transm, cudaTextureObject_t grid, V3 rayOrigin, V3 rayDir, float stepScale,
                         uint32_t numSamples, bool densityOnly, uint32_t& samples)
{
    float t = stepScale;
    float step = stepScale;
    float prevDensity = 0.0f;
    for (uint32_t i = 0; i < numSamples; ++i) {
        const V3 pos = {rayOrigin.x + rayDir.x * t, rayOrigin.y + rayDir.y * t, rayOrigin.z + rayDir.z * t};
        if (outside_unit_box(pos)) break;
        const V3 uvw = local_to_tex3d(pos);
        const float density = fetch_density(grid, uvw.x, uvw.y, uvw.z, densityOnly);
        ++samples;
        const float dDensity = density - prevDensity;
        const float opacity = saturate(density * step);
        const float newStep = get_step(dDensity, 1.0f - transm, opacity, stepScale);
        prevDensity = density;
        transm *= 1.0f - density * kAbsorption;
        if (transm < kZeroThreshold) break;
        step = newStep;
        t += step;
    }
}

// EvaluateSHIrradiance, XUSG/Shaders/SHIrradianceTypeless.hlsli:16-37 (x and y negated)
MV_D V3 evaluate_sh_irradiance(const float* sh, V3 norm)
{
    const float c1 = 0.42904276540489171563379376569857f;
    const float c2 = 0.51166335397324424423977581244463f;
    const float c3 = 0.24770795610037568833406429782001f;
    const float c4 = 0.88622692545275801364908374167057f;
    const float x = -norm.x, y = -norm.y, z = norm.z;
    float irr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float t1 = (c1 * (x * x - y * y)) * sh[8 * 3 + k];
        const float t2 = (c3 * (3.0f * z * z - 1.0f)) * sh[6 * 3 + k];
        const float t3 = c4 * sh[k];
        const float t4 = 2.0f * c1 * ((sh[4 * 3 + k] * x * y + sh[7 * 3 + k] * x * z) + sh[5 * 3 + k] * y * z);
        const float t5 = 2.0f * c2 * ((sh[3 * 3 + k] * x + sh[1 * 3 + k] * y) + sh[2 * 3 + k] * z);
        irr[k] = fmaxf(0.0f, (((t1 + t2) + t3) + t4) + t5);
    }
    return {irr[0], irr[1], irr[2]};
}

} // namespace mv
