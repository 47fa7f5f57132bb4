// mvo_passes.cpp — ORACLE (test infrastructure, not product code).
//
// Scalar restatement of the reference's HLSL passes. Every function cites the shader lines it
// follows (paths relative to /root/reference/MultiVolumes/Content/Shaders unless noted).
// Compiled with -O2 -ffp-contract=off -fopenmp. Never linked into the product.
#include "mvo_core.h"
#include <omp.h>
#include <cstdio>

namespace mvo {

Min16Consts g_min16;

// ------------------------------------------------------------------------------------------
// CSInitGridData.hlsl:10-27 — procedural density (mode 0 verbatim; mode 1 = same envelope times
// seeded value noise so that sources differ)
// ------------------------------------------------------------------------------------------
static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    uint32_t h = seed ^ (x * 0x8da6b343u) ^ (y * 0xd8163841u) ^ (z * 0xcb1ab31fu);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; h *= 0x27d4eb2fu; h ^= h >> 16;
    return h;
}
static inline float lattice(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    return (float)(hash3(x, y, z, seed) >> 8) * (1.0f / 16777216.0f);
}
static float value_noise(f3 p, uint32_t seed)   // p in lattice units, p >= 0
{
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    float tx = p.x - fx, ty = p.y - fy, tz = p.z - fz;
    tx = tx * tx * (3.0f - 2.0f * tx); ty = ty * ty * (3.0f - 2.0f * ty); tz = tz * tz * (3.0f - 2.0f * tz);
    float c[2][2][2];
    for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) c[k][j][i] = lattice(ix + i, iy + j, iz + k, seed);
    const float x00 = lerp1(c[0][0][0], c[0][0][1], tx), x10 = lerp1(c[0][1][0], c[0][1][1], tx);
    const float x01 = lerp1(c[1][0][0], c[1][0][1], tx), x11 = lerp1(c[1][1][0], c[1][1][1], tx);
    return lerp1(lerp1(x00, x10, ty), lerp1(x01, x11, ty), tz);
}

void init_grid_data(Caster& c, uint32_t src, uint32_t mode, uint32_t seed)
{
    Tex3D& t = c.volumes[src];
    const int n = (int)t.n;
    const float gridSize = (float)n;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                // :17 pos = (DTid + 0.5) / gridSize * 2.0 - 1.0
                const f3 pos = {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f,
                                ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};
                const float r_sq = dot3(pos, pos);          // :18
                float a = 1.0f - r_sq;                      // :19
                a *= a;                                     // :20
                a = saturate(a * a * 2.0f);                 // :21
                if (mode == 1) {
                    const f3 q = {(pos.x + 1.0f) * 2.0f, (pos.y + 1.0f) * 2.0f, (pos.z + 1.0f) * 2.0f};
                    float nz = 0.5f * value_noise(q, seed);
                    nz += 0.3f * value_noise(q * 2.0f, seed ^ 0x68bc21ebu);
                    nz += 0.2f * value_noise(q * 4.0f, seed ^ 0x02e5be93u);
                    a = saturate(a * (0.25f + 1.5f * nz));
                }
                const f3 colorU = {1.0f, 0.6f, 0.0f}, colorD = {0.5f, 0.8f, 1.0f};   // :23-24
                const float s = saturate(pos.y * 0.5f + 0.2f);                       // :25
                uint16_t* o = &t.texels[(((size_t)z * n + y) * n + x) * 4];
                o[0] = f32_to_f16(lerp1(colorD.x, colorU.x, s));
                o[1] = f32_to_f16(lerp1(colorD.y, colorU.y, s));
                o[2] = f32_to_f16(lerp1(colorD.z, colorU.z, s));
                o[3] = f32_to_f16(a);
            }
}

// ------------------------------------------------------------------------------------------
// CSVolumeCull.hlsl:13-78 + VolumeCull.hlsli
// ------------------------------------------------------------------------------------------
static f3 project_to_viewport(uint32_t i, const m44& wvp, f2 viewport)   // VolumeCull.hlsli:27-41
{
    const f3 p3 = {(i & 1) ? 1.0f : -1.0f, ((i >> 1) & 1) ? 1.0f : -1.0f, (i >> 2) ? 1.0f : -1.0f};
    f4 p = mul_p44(p3, wvp);
    p.x /= p.w; p.y /= p.w; p.z /= p.w;
    p.x = p.x * 0.5f + 0.5f; p.y = p.y * 0.5f + 0.5f;
    p.y = 1.0f - p.y;
    return {p.x * viewport.x, p.y * viewport.y, p.z};
}

static inline uint32_t float_to_uint_sat(float f)   // HLSL uint(f): NaN -> 0, saturating
{
    if (!(f == f) || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}
// uint(max(log2(x), 0)): floor(log2 x) for x >= 1 taken from the exponent field (exact), else 0.
static inline uint32_t floor_log2_clamped(float x)
{
    if (!(x >= 1.0f)) return 0u;              // also NaN
    return ((as_uint(x) >> 23) & 0xffu) - 127u;   // +inf -> 128
}

void cull_volumes(Caster& c)
{
    const uint32_t N = c.d.num_volumes;
    c.visible.clear(); c.cubeVolumes.clear();
    // VolumeCull.hlsli:119-138 — unique edge id -> corner pair
    static const int edgeLanes[12][2] = {{0, 1}, {3, 2}, {1, 3}, {2, 0}, {6, 7}, {5, 4}, {4, 6}, {7, 5}, {4, 0}, {2, 6}, {7, 3}, {1, 5}};
    // VolumeCull.hlsli:213-223 — face (indexed by mask bit f, labelled -X,+X,-Y,+Y,-Z,+Z) -> 4 edge ids
    static const int faceEdges[6][4] = {{8, 3, 9, 6}, {10, 2, 11, 7}, {0, 8, 5, 11}, {1, 10, 4, 9}, {0, 2, 1, 3}, {4, 6, 5, 7}};
    const float sqrt3 = sqrtf(3.0f);
    for (uint32_t volumeId = 0; volumeId < N; ++volumeId) {
        const PerObject& po = c.perObject[volumeId];
        f3 v[8];
        bool anyInView = false;
        for (uint32_t i = 0; i < 8; ++i) {
            v[i] = project_to_viewport(i, po.WorldViewProj, c.cb.viewport);
            // CSVolumeCull.hlsl:32
            const bool isInView = (v[i].x <= c.cb.viewport.x && v[i].y <= c.cb.viewport.y && v[i].x >= 0.0f && v[i].y >= 0.0f)
                                  && v[i].z > 0.0f && v[i].z < 1.0f;
            anyInView = anyInView || isInView;
        }
        if (!anyInView) continue;   // :38 (attributes of culled volumes keep their previous content)

        const uint32_t volumeIn = c.volumeDescs[volumeId];
        uint32_t raySampleCount = c.d.max_ray_samples;     // :44 g_numSamples

        // GenVisibilityMask, VolumeCull.hlsli:46-66
        const f3 localEye = mul_p43(c.cb.eyePt, po.WorldI);
        uint32_t faceMask = 0;
        for (uint32_t f = 0; f < 6; ++f) {
            const float viewComp = comp(localEye, (int)(f >> 1));
            const bool vis = (f & 1) ? viewComp > -1.0f : viewComp < 1.0f;
            if (vis) faceMask |= 1u << f;
        }
        // GetCubeEdge, :90-152
        f2 e[12];
        for (int k = 0; k < 12; ++k) {
            const f3& a = v[edgeLanes[k][0]]; const f3& b = v[edgeLanes[k][1]];
            e[k] = {b.x - a.x, b.y - a.y};
        }
        // EstimateCubeMaxEdgeLength, :248-262 (max is order independent)
        float maxEdge = 0.0f;
        for (int lane = 0; lane < 6; ++lane) {
            const float ms = fmaxf(length2(e[2 * lane]), length2(e[2 * lane + 1]));
            maxEdge = lane == 0 ? ms : fmaxf(maxEdge, ms);
        }
        // EstimateCubeMapLOD, :267-294 (upscale = 2, raySampleCountScale = 2)
        const uint32_t cubeMapSize = volumeIn >> 18, numMips = (volumeIn >> 14) & 0xf;
        float s = maxEdge / 2.0f;
        float raySampleAmt = 2.0f * s / sqrt3;
        const uint32_t raySampleCnt = float_to_uint_sat(ceilf(raySampleAmt));
        raySampleCount = std::min(raySampleCnt, raySampleCount);
        raySampleAmt = fminf(raySampleAmt, (float)raySampleCount);
        s = raySampleAmt / 2.0f * sqrt3;
        const uint32_t level = floor_log2_clamped((float)cubeMapSize / s);
        const uint32_t mipLevel = std::min(level, numMips - 1);
        // EstimateProjCoverage, :299-322 — WaveActiveSum pinned to lane-ascending sequential sum
        float projCov = 0.0f;
        for (int f = 0; f < 6; ++f) {
            float faceArea = 0.0f;
            if (faceMask & (1u << f)) {
                const f2 e0 = e[faceEdges[f][0]], e1 = e[faceEdges[f][1]], e2 = e[faceEdges[f][2]], e3 = e[faceEdges[f][3]];
                const float t0 = 0.5f * fabsf(e0.x * e1.y - e0.y * e1.x);   // CalcTriangleArea :71-74
                const float t1 = 0.5f * fabsf(e2.x * e3.y - e2.y * e3.x);
                faceArea = t0 + t1;
            }
            projCov = f == 0 ? faceArea : projCov + faceArea;
        }
        // EstimateCubeMapVisiblePixels, :327-334
        const uint32_t edgeLength = cubeMapSize >> mipLevel;
        const float cubeMapPix = (float)(edgeLength * edgeLength) * (float)__builtin_popcount(faceMask);
        const bool useCubeMap = cubeMapPix <= projCov;   // CSVolumeCull.hlsl:66-67
        const uint32_t maskBits = useCubeMap ? (faceMask | kCubeMapRayMarchBit) : faceMask;
        if (useCubeMap) c.cubeVolumes.push_back(volumeId);
        uint16_t* a = &c.attribs[volumeId * 4];
        a[0] = (uint16_t)mipLevel; a[1] = (uint16_t)raySampleCount; a[2] = (uint16_t)maskBits; a[3] = (uint16_t)(volumeIn & 0x3fff);
        c.visible.push_back(volumeId);
    }
    c.stats.visible_count = (uint32_t)c.visible.size();
    c.stats.cubemap_count = (uint32_t)c.cubeVolumes.size();
}

// ------------------------------------------------------------------------------------------
// RayMarch.hlsli helpers
// ------------------------------------------------------------------------------------------
static inline f3 local_to_tex3d(f3 pos) { return {pos.x * 0.5f + 0.5f, pos.y * 0.5f + 0.5f, pos.z * 0.5f + 0.5f}; }   // :170-177

static bool compute_ray_origin(f3& rayOrigin, f3 rayDir)   // :128-155
{
    if (fabsf(rayOrigin.x) <= 1.0f && fabsf(rayOrigin.y) <= 1.0f && fabsf(rayOrigin.z) <= 1.0f) return true;
    float U = kFltMax;
    bool isHit = false;
    for (int i = 0; i < 3; ++i) {
        const float d = comp(rayDir, i), o = comp(rayOrigin, i);
        const float u = (-signf(d) - o) / d;
        if (u < 0.0f) continue;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        if (fabsf(comp(rayDir, j) * u + comp(rayOrigin, j)) > 1.0f) continue;
        if (fabsf(comp(rayDir, k) * u + comp(rayOrigin, k)) > 1.0f) continue;
        if (u < U) { U = u; isHit = true; }
    }
    f3 p = {rayDir.x * U + rayOrigin.x, rayDir.y * U + rayOrigin.y, rayDir.z * U + rayOrigin.z};
    rayOrigin = {fminf(fmaxf(p.x, -1.0f), 1.0f), fminf(fmaxf(p.y, -1.0f), 1.0f), fminf(fmaxf(p.z, -1.0f), 1.0f)};
    return isHit;
}

// :182-192. factorTh = 1 - transm is handed in: in the view march and RayCast transm is itself 1 - scatter.w and the compiled
// shaders use scatter.w there (dxc folds 1 - (1 - x) to x, not an identity in fp32); the product is associated as compiled:
// ((factorTh * 1.5) * factorEv) * factorUi (CSRayMarchV.cso %371-%375, CSRayMarchL.cso %510-%515).
static inline float get_step(float dDensity, float factorTh, float density, float step)
{
    const float factorEv = fminf(1.0f / 256.0f / fabsf(dDensity), 2.0f);
    const float factorUi = fminf(1.0f - density, 1.0f);
    return fmaxf(((factorTh * 1.5f) * factorEv) * factorUi, 1.0f) * step;
}

static float get_tmax(f3 pos, f3 rayOrigin, f3 rayDir, const m44& wvpi)   // :82-92
{
    if (pos.z >= 1.0f) return kFltMax;
    const f4 h = mul_p44(pos, wvpi);
    const f3 p = {h.x / h.w, h.y / h.w, h.z / h.w};
    const f3 t = (p - rayOrigin) / rayDir;
    return max3(t.x, t.y, t.z);
}

struct MarchCounters { uint64_t samples = 0, lightFetches = 0; };

static inline f3 get_light(const Caster& c, uint32_t volumeId, f3 pos)   // :235-240
{
    const f4 l = sample3d(c.lightMaps[volumeId], local_to_tex3d(pos), c.filterModel);
    return {l.x, l.y, l.z};
}

// Per-ray loop shared by CSRayMarch.hlsl:112-155 and RayCast.hlsli:57-105 (the two bodies are
// identical up to the tMax test, which the caller folds into `tMax`).
static f4 march_loop(const Caster& c, uint32_t volumeId, uint32_t volTexId, uint32_t smpCount, f3 rayOrigin, f3 rayDir,
                     float tMax, MarchCounters& mc)
{
    const float maxDist = g_min16.maxDist;              // g_maxDist = 2 sqrt(3), RayMarch.hlsli:17
    const float stepScale = maxDist / (float)smpCount;
    f4 scatter = {0, 0, 0, 0};
    float t = 0.0f;
    float prevDensity = 0.0f;
    const Tex3D& grid = c.volumes[volTexId];
    for (uint32_t i = 0; i < smpCount; ++i) {
        const f3 pos = {rayOrigin.x + rayDir.x * t, rayOrigin.y + rayDir.y * t, rayOrigin.z + rayDir.z * t};
        if (fabsf(pos.x) > 1.0f || fabsf(pos.y) > 1.0f || fabsf(pos.z) > 1.0f) break;
        const f3 uvw = local_to_tex3d(pos);
        f4 color = sample3d(grid, uvw, c.filterModel);   // GetSample
        ++mc.samples;
        float newStep = stepScale;
        if (color.w > kZeroThreshold) {                  // skip empty space
            const f3 light = get_light(c, volumeId, pos);
            ++mc.lightFetches;
            const float transm = 1.0f - scatter.w;
            const float dDensity = color.w - prevDensity;
            newStep = get_step(dDensity, scatter.w, color.w, stepScale);
            prevDensity = color.w;
            // colour (not pre-multiplied) x density x light x ABSORPTION x transmittance (:138-141), associated as the compiled
            // shaders have it (CSRayMarchV.cso %379-%390, PSCube.cso %305-%316): ((transm * A) * a) once, then * colour * light
            const float ka = (transm * kAbsorption) * color.w;
            scatter.x += (ka * color.x) * light.x;
            scatter.y += (ka * color.y) * light.y;
            scatter.z += (ka * color.z) * light.z;
            scatter.w += ka;
            if (transm < kZeroThreshold) break;
        }
        t += newStep;
        if (t > tMax) break;
    }
    const float twoPi = 2.0f * kPi;
    if (g_min16.invTwoPi != 0.0f) { scatter.x *= g_min16.invTwoPi; scatter.y *= g_min16.invTwoPi; scatter.z *= g_min16.invTwoPi; }
    else { scatter.x /= twoPi; scatter.y /= twoPi; scatter.z /= twoPi; }
    return scatter;
}

static inline float point_sample_depth(const Caster& c, f2 uv)   // POINT_CLAMP, CSRayMarch.hlsl:68
{
    const int W = (int)c.d.width, H = (int)c.d.height;
    int ix = (int)floorf(uv.x * (float)W), iy = (int)floorf(uv.y * (float)H);
    if (!(uv.x == uv.x)) ix = 0;
    if (!(uv.y == uv.y)) iy = 0;
    ix = std::min(std::max(ix, 0), W - 1); iy = std::min(std::max(iy, 0), H - 1);
    return c.depth[(size_t)iy * W + ix];
}

// ------------------------------------------------------------------------------------------
// CSRayMarchV (CSRayMarch.hlsl:77-158). Launch shape follows the work-graph variant: exactly
// (G >> mip)^2 texels per visible face (LibRayMarch.hlsl:120-121) — the over-dispatched threads of
// the ExecuteIndirect path write nothing.
// ------------------------------------------------------------------------------------------
static f3 get_local_pos(float px, float py, uint32_t slice, float gridSize)   // :28-53
{
    float x = (px + 0.5f) / gridSize * 2.0f - 1.0f;
    float y = (py + 0.5f) / gridSize * 2.0f - 1.0f;
    y = -y;
    switch (slice) {
    case 0: return {1.0f, y, -x};
    case 1: return {-1.0f, y, x};
    case 2: return {x, 1.0f, -y};
    case 3: return {x, -1.0f, y};
    case 4: return {x, y, 1.0f};
    case 5: return {-x, y, -1.0f};
    default: return {0, 0, 0};
    }
}

void ray_march_view(Caster& c)
{
    uint64_t rays = 0, samples = 0, lightFetches = 0;
    for (uint32_t volumeId : c.cubeVolumes) {
        if (c.shardVolumes ? !c.owns_source(c.volumeDescs[volumeId] & 0x3fff) : (volumeId % c.shardWorld != c.shardRank)) continue;   // marched by its owner rank
        const uint16_t* a = &c.attribs[volumeId * 4];
        const uint32_t mip = a[0], smpCount = a[1], maskBits = a[2], volTexId = a[3];
        const PerObject& po = c.perObject[volumeId];
        const int size = (int)(c.d.grid_size >> mip);
        CubeMap& cm = c.cubeMaps[volumeId];
        for (uint32_t face = 0; face < 6; ++face) {
            if ((maskBits & (1u << face)) == 0) continue;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays, samples, lightFetches)
            for (int y = 0; y < size; ++y)
                for (int x = 0; x < size; ++x) {
                    f3 rayOrigin = mul_p43(c.cb.eyePt, po.WorldI);                         // :88
                    const f3 target = get_local_pos((float)x, (float)y, face, (float)size); // :93
                    const f3 rayDir = normalize3(target - rayOrigin);                       // :94
                    if (!compute_ray_origin(rayOrigin, rayDir)) continue;                   // :95
                    const f3 u = (target - rayOrigin) / rayDir;                             // ComputeTargetHit
                    float tMax = max3(u.x, u.y, u.z);
                    // GetClipPos :59-71
                    const f3 p01 = {rayOrigin.x + 0.01f * rayDir.x, rayOrigin.y + 0.01f * rayDir.y, rayOrigin.z + 0.01f * rayDir.z};
                    const f4 hPos = mul_p44(p01, po.WorldViewProj);
                    const f2 xy = {hPos.x / hPos.w, hPos.y / hPos.w};
                    f2 uv = {xy.x * 0.5f + 0.5f, xy.y * 0.5f + 0.5f};
                    uv.y = 1.0f - uv.y;
                    const float z = point_sample_depth(c, uv);
                    const size_t idx = ((size_t)face * size + y) * size + x;
                    cm.depth[mip][idx] = z;                                                 // :105
                    tMax = fminf(get_tmax({xy.x, xy.y, z}, rayOrigin, rayDir, po.WorldViewProjI), tMax);   // :106
                    MarchCounters mc;
                    const f4 scatter = march_loop(c, volumeId, volTexId, smpCount, rayOrigin, rayDir, tMax, mc);
                    uint16_t* o = &cm.color[mip][idx * 4];
                    o[0] = f32_to_f16(scatter.x); o[1] = f32_to_f16(scatter.y); o[2] = f32_to_f16(scatter.z); o[3] = f32_to_f16(scatter.w);
                    if (c.debugF32) {
                        const size_t G = c.d.grid_size;
                        float* q = &c.dbgCubeF32[((((size_t)volumeId * 6 + face) * G + y) * G + x) * 4];
                        q[0] = scatter.x; q[1] = scatter.y; q[2] = scatter.z; q[3] = scatter.w;
                    }
                    ++rays; samples += mc.samples; lightFetches += mc.lightFetches;
                }
        }
    }
    c.stats.view_rays = rays; c.stats.view_samples = samples; c.stats.view_light_fetches = lightFetches;
}

// ------------------------------------------------------------------------------------------
// SHIrradianceTypeless.hlsli:16-37
// ------------------------------------------------------------------------------------------
f4 evaluate_sh_irradiance(const f3 sh[9], f3 norm)
{
    const float c1 = 0.42904276540489171563379376569857f;
    const float c2 = 0.51166335397324424423977581244463f;
    const float c3 = 0.24770795610037568833406429782001f;
    const float c4 = 0.88622692545275801364908374167057f;
    const float x = -norm.x, y = -norm.y, z = norm.z;
    float irr[3];
    for (int k = 0; k < 3; ++k) {
        auto S = [&](int i) { return comp(sh[i], k); };
        const float t1 = (c1 * (x * x - y * y)) * S(8);
        const float t2 = (c3 * (3.0f * z * z - 1.0f)) * S(6);
        const float t3 = c4 * S(0);
        const float t4 = 2.0f * c1 * ((S(4) * x * y + S(7) * x * z) + S(5) * y * z);
        const float t5 = 2.0f * c2 * ((S(3) * x + S(1) * y) + S(2) * z);
        irr[k] = fmaxf(0.0f, (((t1 + t2) + t3) + t4) + t5);
    }
    const float avgLum = (sh[0].x * 0.25f + sh[0].y * 0.5f) + sh[0].z * 0.25f;
    return {irr[0], irr[1], irr[2], avgLum};
}

// ------------------------------------------------------------------------------------------
// CSRayMarchL.hlsl:20-121
// ------------------------------------------------------------------------------------------
static float shadow_test(const Caster& c, f3 pos)   // RayMarch.hlsli:103-112, LINEAR_LESS_EQUAL comparison sampler
{
    const f4 ls = mul_p44(pos, c.cb.shadowViewProj);
    f2 uv = {ls.x * 0.5f + 0.5f, ls.y * 0.5f + 0.5f};
    uv.y = 1.0f - uv.y;
    const float ref = ls.z - 0.0027f;
    const int S = (int)c.shadowSize;
    if (S == 0) return 1.0f;
    const float fx = uv.x * (float)S - 0.5f, fy = uv.y * (float)S - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float wx = fx - flx, wy = fy - fly;
    auto tap = [&](int ix, int iy) {
        ix = std::min(std::max(ix, 0), S - 1); iy = std::min(std::max(iy, 0), S - 1);
        const float d = (float)c.shadow[(size_t)iy * S + ix] / 65535.0f;     // D16_UNORM
        return ref <= d ? 1.0f : 0.0f;
    };
    const int ix = (int)flx, iy = (int)fly;
    const float t00 = tap(ix, iy), t10 = tap(ix + 1, iy), t01 = tap(ix, iy + 1), t11 = tap(ix + 1, iy + 1);
    return lerp1(lerp1(t00, t10, wx), lerp1(t01, t11, wx), wy);
}

static void cast_light_ray(const Caster& c, float& transm, uint32_t volTexId, f3 rayOrigin, f3 rayDir, float stepScale,
                           uint32_t numSamples, uint64_t& samples)   // RayMarch.hlsli:197-230
{
    float t = stepScale;
    float step = stepScale;
    float prevDensity = 0.0f;
    const Tex3D& grid = c.density_source(volTexId);      // volume-sharded storage: another rank's volume is read through its proxy
    for (uint32_t i = 0; i < numSamples; ++i) {
        const f3 pos = {rayOrigin.x + rayDir.x * t, rayOrigin.y + rayDir.y * t, rayOrigin.z + rayDir.z * t};
        if (fabsf(pos.x) > 1.0f || fabsf(pos.y) > 1.0f || fabsf(pos.z) > 1.0f) break;
        const f3 uvw = local_to_tex3d(pos);
        const float density = sample3d(grid, uvw, c.filterModel).w;
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

static f3 density_gradient(const Caster& c, uint32_t volTexId, f3 uvw)   // RayMarch.hlsli:55-77
{
    static const int off[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    const Tex3D& grid = c.volumes[volTexId];
    const float inv = 1.0f / (float)grid.n;      // integer texel offsets applied in normalised space
    float q[6];
    for (int j = 0; j < 6; ++j) {
        const f3 p = {uvw.x + (float)off[j][0] * inv, uvw.y + (float)off[j][1] * inv, uvw.z + (float)off[j][2] * inv};
        q[j] = sample3d(grid, p, c.filterModel).w;
    }
    return {q[1] - q[0], q[3] - q[2], q[5] - q[4]};
}

void ray_march_light(Caster& c, int volumeOverride)
{
    const uint32_t N = c.d.num_volumes;
    const int L = (int)c.d.light_grid_size;
    // :29-33
    uint32_t volumeId;
    if (volumeOverride >= 0) volumeId = (uint32_t)volumeOverride;
    else if (!c.visible.empty()) volumeId = c.visible[c.cb.frameIdx % c.visible.size()];
    else volumeId = c.cb.frameIdx % N;
    c.stats.light_volume = volumeId;
    const uint32_t volTexId0 = c.volumeDescs[volumeId] & 0x3fff;
    if (!c.owns_source(volTexId0)) { c.stats.light_voxels = 0; c.stats.light_dense_voxels = 0; c.stats.light_samples = 0; return; }   // marched by its owner
    const PerObject& po0 = c.perObject[volumeId];
    const float gridSize = (float)L;
    const float maxDist = g_min16.maxDist;
    const uint32_t numSamples = c.d.max_light_samples;
    const float gStep = maxDist / (float)numSamples;     // RayMarch.hlsli:18
    Tex3D& lm = c.lightMaps[volumeId];
    uint64_t dense = 0, samples = 0;
    const int slab = c.shardVolumes ? L : (L + (int)c.shardWorld - 1) / (int)c.shardWorld;
    const int zBegin = c.shardVolumes ? 0 : std::min(L, (int)c.shardRank * slab), zEnd = std::min(L, zBegin + slab);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dense, samples)
    for (int z = zBegin; z < zEnd; ++z)
        for (int y = 0; y < L; ++y)
            for (int x = 0; x < L; ++x) {
                f3 rayOrigin = {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f,
                                ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};                       // :36
                const f3 uvw = local_to_tex3d(rayOrigin);                                           // :41
                const float density = sample3d(c.volumes[volTexId0], uvw, c.filterModel).w;         // :45
                const bool hasDensity = density >= kZeroThreshold;                                  // :46
                rayOrigin = mul_p43(rayOrigin, po0.World);                                          // :48
                float shadow = shadow_test(c, rayOrigin);                                           // :51
                float ao = 1.0f;
                f3 irradiance = {0, 0, 0};
                if (hasDensity) {
                    ++dense;
                    f3 aoRayDir = {0, 0, 0};
                    if (c.hasSH) {                                                                  // :65-75
                        aoRayDir = -density_gradient(c, volTexId0, uvw);
                        const bool nz = fabsf(aoRayDir.x) > 0.0f || fabsf(aoRayDir.y) > 0.0f || fabsf(aoRayDir.z) > 0.0f;
                        aoRayDir = nz ? aoRayDir : rayOrigin;
                        aoRayDir = mul_v33(aoRayDir, po0.World);
                        aoRayDir = normalize3(aoRayDir);
                        const f4 irr = evaluate_sh_irradiance(c.sh, normalize3(aoRayDir));          // GetIrradiance
                        irradiance = {irr.x, irr.y, irr.z};
                    }
                    for (uint32_t n = 0; n < N; ++n) {                                              // :77
                        const uint32_t volTexId = c.volumeDescs[n] & 0x3fff;
                        const PerObject& po = c.perObject[n];
                        f3 localRayOrigin = mul_p43(rayOrigin, po.WorldI);                          // :83
                        if (shadow >= kZeroThreshold) {
                            const f3 lightPos3 = {c.cb.lightPos.x, c.cb.lightPos.y, c.cb.lightPos.z};
                            const f3 rayDir = normalize3(mul_v33(lightPos3, po.WorldI));            // :91-92
                            if (!compute_ray_origin(localRayOrigin, rayDir)) continue;              // :95
                            cast_light_ray(c, shadow, volTexId, localRayOrigin, rayDir, gStep, numSamples, samples);
                        }
                        if (c.hasSH) {                                                              // :100-108
                            const f3 rayDir = normalize3(mul_v33(aoRayDir, po.WorldI));
                            if (!compute_ray_origin(localRayOrigin, rayDir)) continue;
                            float transm = 1.0f;
                            cast_light_ray(c, transm, volTexId, localRayOrigin, rayDir, gStep, numSamples, samples);
                            ao *= (n == volumeId) ? transm : pow025(saturate(transm + 0.5f));
                        }
                    }
                }
                const f3 lightColor = {c.cb.lightColor.x * c.cb.lightColor.w, c.cb.lightColor.y * c.cb.lightColor.w, c.cb.lightColor.z * c.cb.lightColor.w};
                f3 ambient = {c.cb.ambient.x * c.cb.ambient.w, c.cb.ambient.y * c.cb.ambient.w, c.cb.ambient.z * c.cb.ambient.w};
                if (c.hasSH) ambient = {ao * irradiance.x, ao * irradiance.y, ao * irradiance.z};   // :117
                const f3 out = {shadow * lightColor.x + ambient.x, shadow * lightColor.y + ambient.y, shadow * lightColor.z + ambient.z};
                if (c.debugF32) { float* q = &c.dbgLightF32[(((size_t)z * L + y) * L + x) * 3]; q[0] = out.x; q[1] = out.y; q[2] = out.z; }
                uint16_t* o = &lm.texels[(((size_t)z * L + y) * L + x) * 4];
                // R11G11B10_FLOAT store, kept in an RGBA16F texel (exactly representable)
                o[0] = f32_to_f16(quantize_ufloat(out.x, 6));
                o[1] = f32_to_f16(quantize_ufloat(out.y, 6));
                o[2] = f32_to_f16(quantize_ufloat(out.z, 5));
                o[3] = 0;
            }
    c.stats.light_voxels = (uint64_t)L * L * L; c.stats.light_dense_voxels = dense; c.stats.light_samples = samples;
}

// ------------------------------------------------------------------------------------------
// OIT: VSCube / PSDepthPeel / PSCube / CubeCast / RayCast / PSResolveOIT restated per pixel.
// The hardware rasteriser of the reference is replaced by the analytic exit point of the pixel-centre
// ray on each visible volume's box (what RTCube.hlsl:72-98 does with ray queries).
// ------------------------------------------------------------------------------------------
// Stated evaluation order of the OIT passes (the product's k_oit.cu states the same one): divisions as multiplications by the
// correctly rounded reciprocal (rcp), fused multiply-adds (fma1) exactly where written — see the note above rgb_to_ycocg.
static inline float unproject_z(float depth)   // PSCube.hlsli:21-26
{
    return (kZNear * kZFar) * rcp(fma1(depth, kZNear - kZFar, kZFar));
}
static inline f3 mul_v33_f(f3 v, const m43& M)   // mul(v, (float3x3)M), fused
{
    return {fma1(v.z, M.m[2][0], fma1(v.y, M.m[1][0], v.x * M.m[0][0])), fma1(v.z, M.m[2][1], fma1(v.y, M.m[1][1], v.x * M.m[0][1])),
            fma1(v.z, M.m[2][2], fma1(v.y, M.m[1][2], v.x * M.m[0][2]))};
}

// D3D cube-map convention: face index and (u, v) in [0,1] of a point on the unit cube surface.
static inline void cube_face_uv(f3 p, int face, float& u, float& v)
{
    switch (face) {
    case 0: u = -p.z * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    case 1: u = p.z * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    case 2: u = p.x * 0.5f + 0.5f; v = p.z * 0.5f + 0.5f; break;
    case 3: u = p.x * 0.5f + 0.5f; v = -p.z * 0.5f + 0.5f; break;
    case 4: u = p.x * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    default: u = -p.x * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    }
}

// Seamless cube addressing: texel (i, j) of `face`, where one index may be -1 or S, resolved to the
// edge-adjacent texel of the neighbouring face. Coordinates are odd integers in units of 1/S.
static inline void cube_resolve_texel(int S, int face, int i, int j, int& oface, int& oi, int& oj)
{
    const bool iOut = i < 0 || i >= S;
    if (iOut && (j < 0 || j >= S)) j = j < 0 ? 0 : S - 1;   // corner: pinned to the edge texel
    const bool jOut = j < 0 || j >= S;
    if (!iOut && !jOut) { oface = face; oi = i; oj = j; return; }
    const int a = 2 * i + 1 - S, b = 2 * j + 1 - S;     // face-plane coordinates (u, v directions)
    int X, Y, Z;
    switch (face) {     // inverse of cube_face_uv with the major axis at +-S
    case 0: X = S; Y = -b; Z = -a; break;
    case 1: X = -S; Y = -b; Z = a; break;
    case 2: X = a; Y = S; Z = b; break;
    case 3: X = a; Y = -S; Z = -b; break;
    case 4: X = a; Y = -b; Z = S; break;
    default: X = -a; Y = -b; Z = -S; break;
    }
    int P[3] = {X, Y, Z};
    const int major = face >> 1;
    int over = -1;
    for (int k = 0; k < 3; ++k) if (k != major && (P[k] > S || P[k] < -S)) over = k;
    if (over >= 0) {
        const int e = (P[over] > 0 ? P[over] : -P[over]) - S;      // = 1
        P[major] = (P[major] > 0 ? 1 : -1) * (S - e);
        P[over] = P[over] > 0 ? S : -S;
        oface = over * 2 + (P[over] > 0 ? 0 : 1);
    } else oface = face;
    int ua, vb;
    switch (oface) {
    case 0: ua = -P[2]; vb = -P[1]; break;
    case 1: ua = P[2]; vb = -P[1]; break;
    case 2: ua = P[0]; vb = P[2]; break;
    case 3: ua = P[0]; vb = -P[2]; break;
    case 4: ua = P[0]; vb = -P[1]; break;
    default: ua = -P[0]; vb = -P[1]; break;
    }
    oi = (ua + S - 1) / 2; oj = (vb + S - 1) / 2;
}

// exposed for the known-answer test of the seamless Gather addressing (mvo.h)
extern "C" void mvo_cube_resolve_texel(int S, int face, int i, int j, int* out3) { cube_resolve_texel(S, face, i, j, out3[0], out3[1], out3[2]); }

static f4 cube_cast(const Caster& c, uint32_t volumeId, uint32_t mip, int px, int py, int face, f3 pos, f3 rayDir)   // PSCube.hlsli:51-108
{
    const int S = (int)(c.d.grid_size >> mip);
    const float gridSize = (float)S;
    const CubeMap& cm = c.cubeMaps[volumeId];
    float u, v;
    cube_face_uv(pos, face, u, v);
    // gather footprint
    const float fx = fma1(u, gridSize, -0.5f), fy = fma1(v, gridSize, -0.5f);
    const float flx = floorf(fx), fly = floorf(fy);
    const int i0 = (int)flx, j0 = (int)fly;
    const int ti[4] = {i0, i0 + 1, i0 + 1, i0}, tj[4] = {j0 + 1, j0 + 1, j0, j0};   // Gather order (-,+),(+,+),(+,-),(-,-)
    f4 samples[4]; float zs[4];
    for (int k = 0; k < 4; ++k) {
        int f, i, j;
        cube_resolve_texel(S, face, ti[k], tj[k], f, i, j);
        const size_t idx = ((size_t)f * S + j) * S + i;
        const uint16_t* h = &cm.color[mip][idx * 4];
        samples[k] = {f16_to_f32(h[0]), f16_to_f32(h[1]), f16_to_f32(h[2]), f16_to_f32(h[3])};
        zs[k] = cm.depth[mip][idx];
    }
    // GetDomain :31-46
    float uvx = u * gridSize, uvy = v * gridSize;
    float domx = fracf(fma1(u, gridSize, 0.5f)), domy = fracf(fma1(v, gridSize, 0.5f));
    const float bound = gridSize - 1.0f;
    const f3 axes = pos * gridSize;
    const bool edge = (fabsf(axes.x) > bound && axes.x * rayDir.x < 0.0f) || (fabsf(axes.y) > bound && axes.y * rayDir.y < 0.0f) ||
                      (fabsf(axes.z) > bound && axes.z * rayDir.z < 0.0f);
    if (edge) {
        uvx = fminf(uvx, gridSize - 0.5f); uvy = fminf(uvy, gridSize - 0.5f);
        domx = uvx < 0.5f ? 1.0f : 0.0f; domy = uvy < 0.5f ? 1.0f : 0.0f;
    }
    const float dix = 1.0f - domx, diy = 1.0f - domy;
    const float wb[4] = {dix * domy, domx * domy, domx * diy, dix * diy};
    float depth = c.depth[(size_t)py * c.d.width + px];
    depth = unproject_z(depth);
    f4 result = {0, 0, 0, 0};
    float ws = 0.0f;
    for (int k = 0; k < 4; ++k) {
        const float zi = unproject_z(zs[k]);
        float w = fmaxf(fma1(-0.5f, fabsf(depth - zi), 1.0f), 0.0f);
        w *= wb[k];
        result = {fma1(samples[k].x, w, result.x), fma1(samples[k].y, w, result.y), fma1(samples[k].z, w, result.z), fma1(samples[k].w, w, result.w)};
        ws += w;
    }
    if (ws > 0.0f) { const float iw = rcp(ws); return {result.x * iw, result.y * iw, result.z * iw, result.w * iw}; }
    // fallback: plain bilinear SampleLevel of the same footprint (:57)
    const float bx = fx - flx, by = fy - fly;
    const float bw[4] = {(1.0f - bx) * by, bx * by, bx * (1.0f - by), (1.0f - bx) * (1.0f - by)};
    f4 col = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) col = {fma1(samples[k].x, bw[k], col.x), fma1(samples[k].y, bw[k], col.y), fma1(samples[k].z, bw[k], col.z), fma1(samples[k].w, bw[k], col.w)};
    return col;
}

struct Fragment { uint32_t key; uint32_t volumeId; int face; f3 lpt; };

void resolve_oit(Caster& c)
{
    const int W = (int)c.d.width, H = (int)c.d.height;
    uint64_t frags = 0, dRays = 0, dSamples = 0, dLight = 0;
    const size_t nvis = c.visible.size();
    std::vector<f3> eyeL(nvis);
    for (size_t k = 0; k < nvis; ++k) eyeL[k] = mul_p43(c.cb.eyePt, c.perObject[c.visible[k]].WorldI);
    // rows [row0, row1) plus a one-row halo (clipped), read by the TAA of the band's border rows
    const int rowBegin = c.row1 > c.row0 ? std::max((int)c.row0 - 1, 0) : 0, rowEnd = c.row1 > c.row0 ? std::min((int)c.row1 + 1, H) : 0;
    if (c.debugOIT) { c.dbgCount.assign((size_t)W * H, 0); c.dbgInfo.assign((size_t)W * H * 8 * 4, 0); c.dbgData.assign((size_t)W * H * 8 * 9, 0.0f); c.dbgResult.assign((size_t)W * H * 4, 0.0f); c.dbgAllKeys.assign((size_t)W * H * c.d.num_volumes, 0xffffffffu); }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : frags, dRays, dSamples, dLight)
    for (int py = rowBegin; py < rowEnd; ++py)
        for (int px = 0; px < W; ++px) {
            // pixel-centre ray (RTCube.hlsl GenerateCameraRay: unproject z = 0 through screenToWorld)
            // PSCube.hlsl:38-40: xy = (pixel centre / viewport) * 2 - 1, y up; unprojected at z = 0
            const f2 xy = {fma1((float)px + 0.5f, 2.0f / c.cb.viewport.x, -1.0f), fma1((float)py + 0.5f, -(2.0f / c.cb.viewport.y), 1.0f)};
            const m44& S = c.cb.screenToWorld;
            f4 wh;
            wh.x = fma1(xy.x, S.m[0][0], fma1(xy.y, S.m[1][0], S.m[3][0])); wh.y = fma1(xy.x, S.m[0][1], fma1(xy.y, S.m[1][1], S.m[3][1]));
            wh.z = fma1(xy.x, S.m[0][2], fma1(xy.y, S.m[1][2], S.m[3][2])); wh.w = fma1(xy.x, S.m[0][3], fma1(xy.y, S.m[1][3], S.m[3][3]));
            const float iwh = rcp(wh.w);
            const f3 wpos = {wh.x * iwh, wh.y * iwh, wh.z * iwh};
            const f3 dirW = wpos - c.cb.eyePt;
            // depth peel: the 8 nearest back-face fragments (PSDepthPeel.hlsl:12-24)
            Fragment layers[kNumOitLayers];
            int nl = 0;
            for (size_t k = 0; k < nvis; ++k) {
                const uint32_t volumeId = c.visible[k];
                const PerObject& po = c.perObject[volumeId];
                const f3 o = eyeL[k];
                const f3 d = mul_v33_f(dirW, po.WorldI);
                float tmin = -kFltMax, tmax = kFltMax; int exitAxis = -1; bool miss = false;
                for (int a = 0; a < 3; ++a) {
                    const float da = comp(d, a), oa = comp(o, a);
                    if (da == 0.0f) { if (fabsf(oa) > 1.0f) miss = true; continue; }
                    const float inv = rcp(da);
                    const float t1 = (-1.0f - oa) * inv, t2 = (1.0f - oa) * inv;
                    const float tn = fminf(t1, t2), tf = fmaxf(t1, t2);
                    if (tn > tmin) tmin = tn;
                    if (tf < tmax) { tmax = tf; exitAxis = a; }
                }
                if (miss || exitAxis < 0 || !(tmax > 0.0f) || !(tmin < tmax)) continue;
                f3 lpt = {fminf(fmaxf(fma1(d.x, tmax, o.x), -1.0f), 1.0f), fminf(fmaxf(fma1(d.y, tmax, o.y), -1.0f), 1.0f),
                          fminf(fmaxf(fma1(d.z, tmax, o.z), -1.0f), 1.0f)};
                const float sgn = comp(d, exitAxis) > 0.0f ? 1.0f : -1.0f;
                if (exitAxis == 0) lpt.x = sgn; else if (exitAxis == 1) lpt.y = sgn; else lpt.z = sgn;
                const m44& M = po.WorldViewProj;
                const float clipZ = fma1(lpt.z, M.m[2][2], fma1(lpt.y, M.m[1][2], fma1(lpt.x, M.m[0][2], M.m[3][2])));
                const float clipW = fma1(lpt.z, M.m[2][3], fma1(lpt.y, M.m[1][3], fma1(lpt.x, M.m[0][3], M.m[3][3])));
                if (!(clipW > 0.0f)) continue;
                const float z = clipZ * rcp(clipW);
                if (!(z >= 0.0f && z <= 1.0f)) continue;       // rasteriser depth clip
                ++frags;
                Fragment fr = {as_uint(z), volumeId, exitAxis * 2 + (sgn > 0.0f ? 0 : 1), lpt};
                if (c.debugOIT) c.dbgAllKeys[((size_t)py * W + px) * c.d.num_volumes + k] = fr.key;
                // keep the kNumOitLayers smallest keys; ties keep list order (stable)
                int pos = nl;
                while (pos > 0 && layers[pos - 1].key > fr.key) --pos;
                if (pos >= (int)kNumOitLayers) continue;
                const int last = nl < (int)kNumOitLayers ? nl : (int)kNumOitLayers - 1;
                for (int m = last; m > pos; --m) layers[m] = layers[m - 1];
                layers[pos] = fr;
                if (nl < (int)kNumOitLayers) ++nl;
            }
            // shade + resolve (PSCube.hlsl:30-60, PSResolveOIT.hlsl:12-26)
            f4 result = {0, 0, 0, 0};
            for (int l = 0; l < nl; ++l) {
                const Fragment& fr = layers[l];
                const uint16_t* a = &c.attribs[fr.volumeId * 4];
                const PerObject& po = c.perObject[fr.volumeId];
                const f3 localEye = mul_p43(c.cb.eyePt, po.WorldI);
                const f3 rayDir = fr.lpt - localEye;                                   // PSCube.hlsl:34
                const uint32_t smpCnt = (a[2] & kCubeMapRayMarchBit) ? 0 : a[1];       // VSCube.hlsl:73
                f4 color;
                if (smpCnt > 0) {
                    // RayCast.hlsli:42-107
                    f3 ro = localEye; const f3 rd = normalize3(rayDir);
                    if (!compute_ray_origin(ro, rd)) color = {0, 0, 0, 0};
                    else {
                        const float zd = c.depth[(size_t)py * W + px];
                        const float tMax = get_tmax({xy.x, xy.y, zd}, ro, rd, po.WorldViewProjI);
                        MarchCounters mc;
                        color = march_loop(c, fr.volumeId, a[3], smpCnt, ro, rd, tMax, mc);
                        ++dRays; dSamples += mc.samples; dLight += mc.lightFetches;
                    }
                } else color = cube_cast(c, fr.volumeId, a[0], px, py, fr.face, fr.lpt, rayDir);
                // K-colour layers are RGBA16F; unwritten layers stay cleared (PSCube.hlsl:57)
                f4 src = {0, 0, 0, 0};
                if (color.w > 0.0f && color.w <= 1.0f)
                    src = {f16_to_f32(f32_to_f16(color.x)), f16_to_f32(f32_to_f16(color.y)), f16_to_f32(f32_to_f16(color.z)), f16_to_f32(f32_to_f16(color.w))};
                if (c.debugOIT) {
                    const size_t q = ((size_t)py * W + px) * 8 + l;
                    float fu, fv; cube_face_uv(fr.lpt, fr.face, fu, fv);
                    uint32_t* di = &c.dbgInfo[q * 4]; di[0] = fr.key; di[1] = fr.volumeId; di[2] = (uint32_t)fr.face; di[3] = (color.w > 0.0f && color.w <= 1.0f) ? 1u : 0u;
                    float* dd = &c.dbgData[q * 9]; dd[0] = fr.lpt.x; dd[1] = fr.lpt.y; dd[2] = fr.lpt.z; dd[3] = fu; dd[4] = fv; dd[5] = color.x; dd[6] = color.y; dd[7] = color.z; dd[8] = color.w;
                }
                const float k = 1.0f - result.w;
                result = {src.x * k + result.x, src.y * k + result.y, src.z * k + result.z, src.w * k + result.w};   // fmul, fadd: as PSResolveOIT.cso has it
            }
            result.w = fminf(result.w, g_min16.alphaClamp);
            if (c.debugOIT) { c.dbgCount[(size_t)py * W + px] = (uint32_t)nl; float* dr = &c.dbgResult[((size_t)py * W + px) * 4]; dr[0] = result.x; dr[1] = result.y; dr[2] = result.z; dr[3] = result.w; }
            // premultiplied-alpha blend onto the colour RT (MultiRayCaster.cpp:931)
            uint16_t* dst = &c.color[((size_t)py * W + px) * 4];
            const float ia = 1.0f - result.w;
            const f4 d4 = {f16_to_f32(dst[0]), f16_to_f32(dst[1]), f16_to_f32(dst[2]), f16_to_f32(dst[3])};
            dst[0] = f32_to_f16(fma1(d4.x, ia, result.x)); dst[1] = f32_to_f16(fma1(d4.y, ia, result.y));
            dst[2] = f32_to_f16(fma1(d4.z, ia, result.z)); dst[3] = f32_to_f16(fma1(d4.w, ia, result.w));
        }
    c.stats.oit_fragments = frags; c.stats.direct_rays = dRays; c.stats.direct_samples = dSamples; c.stats.direct_light_fetches = dLight;
}

// ------------------------------------------------------------------------------------------
// CSTemporalAA.hlsl:254-336 (ALPHA_BOUND = 1.0, _USE_YCOCG_, _VARIANCE_AABB_) and PSToneMap.hlsl:19-28
// ------------------------------------------------------------------------------------------
// Stated evaluation order of this pass (the product's k_post.cu states the same one): YCoCg as sums and doublings, every
// division as a multiplication by the correctly rounded reciprocal (rcp) or by the reciprocal constant, fused multiply-adds
// (fma1) exactly where written. dxc compiles the reference with fast-math: its DXIL multiplies by reciprocals too.
static inline f3 rgb_to_ycocg(f3 rgb)   // :78-85: (1 2 1; 2 0 -2; -1 2 -1)
{
    const float g2 = rgb.y + rgb.y;
    return {(rgb.x + g2) + rgb.z, (rgb.x - rgb.z) + (rgb.x - rgb.z), (g2 - rgb.x) - rgb.z};
}
static inline f3 TM(f3 hdr) { const f3 c = rgb_to_ycocg(hdr); const float q = rcp(4.0f + c.x); return {c.x * q, c.y * q, c.z * q}; }   // :106-114
// :119-128 with :90-101 folded in: (c * (4 / (1 - c.x))) * 0.25 = c * rcp(1 - c.x) — scaling by 4 and by 0.25 is exact
static inline f3 ITM(f3 c) { const float q = rcp(1.0f - c.x); const float y = c.x * q, co = c.y * q, cg = c.z * q; return {y + co - cg, y + cg, y - co - cg}; }
static inline float lerpf(float a, float b, float t) { return fma1(b - a, t, a); }   // lerp as one fused multiply-add

void temporal_aa(Caster& c, bool taaOn)
{
    const int W = (int)c.d.width, H = (int)c.d.height;
    c.frameParity ^= 1u;                                   // ObjectRenderer.cpp:217
    std::vector<uint16_t>& out = c.taaHistory[c.frameParity];
    const std::vector<uint16_t>& hist = c.taaHistory[c.frameParity ^ 1u];
    if (!taaOn) { std::copy(c.color.begin() + (size_t)c.row0 * W * 4, c.color.begin() + (size_t)c.row1 * W * 4, out.begin() + (size_t)c.row0 * W * 4); return; }
    auto loadC = [&](const std::vector<uint16_t>& img, int x, int y) -> f4 {   // Texture2D[] load: out of bounds -> 0
        if (x < 0 || y < 0 || x >= W || y >= H) return {0, 0, 0, 0};
        const uint16_t* p = &img[((size_t)y * W + x) * 4];
        return {f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2]), f16_to_f32(p[3])};
    };
    auto loadV = [&](int x, int y) -> f2 {
        if (x < 0 || y < 0 || x >= W || y >= H) return {0, 0};
        const uint16_t* p = &c.velocity[((size_t)y * W + x) * 2];
        return {f16_to_f32(p[0]), f16_to_f32(p[1])};
    };
    static const int offs[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {1, 1}, {-1, 1}};   // :46-50
    const float historyMax = 15.0f;                                                                         // :41-43
#pragma omp parallel for schedule(static)
    for (int y = (int)c.row0; y < (int)c.row1; ++y)
        for (int x = 0; x < W; ++x) {
            const f2 texSize = {(float)W, (float)H};
            const f2 uv = {((float)x + 0.5f) / texSize.x, ((float)y + 0.5f) / texSize.y};   // :258, a division in the shipped DXIL too
            const f4 current = loadC(c.color, x, y);
            // VelocityMax :133-161
            f2 vmax = loadV(x, y);
            float speedSq = fma1(vmax.x, vmax.x, vmax.y * vmax.y);
            for (int i = 0; i < 4; ++i) {
                const f2 nb = loadV(x + offs[i + 4][0], y + offs[i + 4][1]);
                const float s = fma1(nb.x, nb.x, nb.y * nb.y);
                if (s > speedSq) { vmax = nb; speedSq = s; }
            }
            const f2 uvBack = {uv.x - vmax.x, uv.y - vmax.y};
            // history.SampleLevel(g_smpLinear, uvBack, 0): bilinear, clamp. Texel coordinates in fixed point with 8 fractional
            // bits (axis_q8): D3D's contract, and it makes a fetch at a texel centre return that texel, which the
            // `historyBlur > 0` test below depends on (with fp32 coordinates it hangs on the last ulp of u * W - 0.5; found by
            // running the reference's CSTemporalAA.cso, oracle/dxil). The blend is fp32; zero-weight taps do not contribute.
            f4 history;
            {
                const AxisFix ax = axis_q8(uvBack.x, W), ay = axis_q8(uvBack.y, H);
                const float wx = ax.ffrac, wy = ay.ffrac;
                const f4 t00 = loadC(hist, ax.i0, ay.i0), t10 = loadC(hist, ax.i1, ay.i0);
                const f4 t01 = loadC(hist, ax.i0, ay.i1), t11 = loadC(hist, ax.i1, ay.i1);
                auto L = [&](float a, float b, float cc, float d) { return lerp_q8(lerp_q8(a, b, wx), lerp_q8(cc, d, wx), wy); };
                history = {L(t00.x, t10.x, t01.x, t11.x), L(t00.y, t10.y, t01.y, t11.y), L(t00.z, t10.z, t01.z, t11.z), L(t00.w, t10.w, t01.w, t11.w)};
            }
            // :267-275
            float curHistoryBlur = fma1(fabsf(vmax.x), 4.0f * texSize.x, fabsf(vmax.y) * (4.0f * texSize.y));
            float historyBlur = 1.0f - history.w;
            historyBlur = fmaxf(historyBlur, curHistoryBlur);
            history.w = fma1(history.w, historyMax, 1.0f);
            // :278-287 (ALPHA_BOUND = 1.0)
            const f3 ctm = TM({current.x, current.y, current.z});
            const f4 currentTM = {ctm.x, ctm.y, ctm.z, current.w};
            const float gamma = (historyBlur > 0.0f || current.w < 1.0f) ? 1.0f : 16.0f;
            // NeighborMinMax :166-236
            f4 cur = currentTM;
            f3 mu = {cur.x, cur.y, cur.z};
            cur.w = cur.w < 1.0f ? 0.0f : 1.0f;
            f3 m2 = mu * mu;
            static const float weights[8] = {0.5f, 0.5f, 0.5f, 0.5f, 0.25f, 0.25f, 0.25f, 0.25f};
            for (int i = 0; i < 8; ++i) {
                const f4 nbr = loadC(c.color, x + offs[i][0], y + offs[i][1]);
                const f3 ntm = TM({nbr.x, nbr.y, nbr.z});
                const float na = nbr.w < 1.0f ? 0.0f : 1.0f;
                cur = {fma1(ntm.x, weights[i], cur.x), fma1(ntm.y, weights[i], cur.y), fma1(ntm.z, weights[i], cur.z), fma1(na, weights[i], cur.w)};
                mu = mu + ntm;
                m2 = {fma1(ntm.x, ntm.x, m2.x), fma1(ntm.y, ntm.y, m2.y), fma1(ntm.z, ntm.z, m2.z)};
            }
            const float ninth = g_min16.ninth;
            cur = cur * 0.25f;
            mu = mu * ninth;
            const f3 m2n = m2 * ninth;
            const f3 sigma = {sqrtf(fabsf(fma1(-mu.x, mu.x, m2n.x))), sqrtf(fabsf(fma1(-mu.y, mu.y, m2n.y))), sqrtf(fabsf(fma1(-mu.z, mu.z, m2n.z)))};
            const f3 gsigma = sigma * gamma;
            f4 neighborMin, neighborMax;
            neighborMin.x = fminf(mu.x - gsigma.x, cur.x); neighborMin.y = fminf(mu.y - gsigma.y, cur.y); neighborMin.z = fminf(mu.z - gsigma.z, cur.z);
            neighborMax.x = fmaxf(mu.x + gsigma.x, cur.x); neighborMax.y = fmaxf(mu.y + gsigma.y, cur.y); neighborMax.z = fmaxf(mu.z + gsigma.z, cur.z);
            neighborMin.w = mu.x - sigma.x;   // GET_LUMA4 = .x in YCoCg
            neighborMax.w = mu.x + sigma.x;
            f4 filtered = cur;
            // :290-301
            curHistoryBlur = saturate(curHistoryBlur);
            historyBlur = saturate(historyBlur);
            f3 historyTM = TM({history.x, history.y, history.z});
            historyTM = {fminf(fmaxf(historyTM.x, neighborMin.x), neighborMax.x), fminf(fmaxf(historyTM.y, neighborMin.y), neighborMax.y),
                         fminf(fmaxf(historyTM.z, neighborMin.z), neighborMax.z)};
            const float contrast = neighborMax.w - neighborMin.w;
            // :304-311
            const float lumContrastFactor = 32.0f * 4.0f;
            float addAlias = fma1(historyBlur, 0.5f, 0.25f);
            addAlias = saturate(addAlias + rcp(fma1(contrast, lumContrastFactor, 1.0f)));
            filtered.x = lerpf(filtered.x, currentTM.x, addAlias); filtered.y = lerpf(filtered.y, currentTM.y, addAlias); filtered.z = lerpf(filtered.z, currentTM.z, addAlias);
            // :314-326
            const float lumHist = historyTM.x;
            const float distToClamp = fminf(fabsf(neighborMin.w - lumHist), fabsf(neighborMax.w - lumHist));
            const float historyAmt = fminf(fma1(historyBlur, 0.125f, rcp(history.w)), 1.0f);
            float blend = 0.25f * rcp(lerpf(8.0f, distToClamp + contrast, historyAmt));
            blend = fminf(blend, 0.25f);
            blend = filtered.w > 0.0f ? blend : 1.0f;
            // :328-330
            f3 result = ITM({lerpf(historyTM.x, filtered.x, blend), lerpf(historyTM.y, filtered.y, blend), lerpf(historyTM.z, filtered.z, blend)});
            if (result.x != result.x || result.y != result.y || result.z != result.z) result = ITM({filtered.x, filtered.y, filtered.z});
            history.w = fminf(history.w * (1.0f / historyMax), 1.0f - curHistoryBlur);
            uint16_t* o = &out[((size_t)y * W + x) * 4];
            o[0] = f32_to_f16(result.x); o[1] = f32_to_f16(result.y); o[2] = f32_to_f16(result.z); o[3] = f32_to_f16(history.w);
        }
}

// ------------------------------------------------------------------------------------------
// LightProbe::RenderEnvironment (LightProbe.cpp:85-97: screen quad at z = 1, DEPTH_READ_LESS_EQUAL) + PSEnvironment.hlsl:46-69
// (infinite-size branch): the radiance cube map along the pixel's ray, alpha 0, wherever the scene depth is 1. The cube lookup
// is restated as bilinear filtering of mip 0 with fp32 weights and seamless edge taps (the sampler's fixed-point weights are
// hardware detail); same stated order as the product's k_env.cu.
// ------------------------------------------------------------------------------------------
void render_environment(Caster& c)
{
    const int W = (int)c.d.width, H = (int)c.d.height, S = (int)c.envSize;
    c.color = c.background;                                       // what the mesh pass left
    if (!S) return;
    const m44& M = c.cb.screenToWorld;
    const int rowBegin = c.row1 > c.row0 ? std::max((int)c.row0 - 1, 0) : 0, rowEnd = c.row1 > c.row0 ? std::min((int)c.row1 + 1, H) : 0;
#pragma omp parallel for schedule(static)
    for (int py = rowBegin; py < rowEnd; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t pix = (size_t)py * W + px;
            if (!(1.0f <= c.depth[pix])) continue;
            const float sx = fma1((float)px + 0.5f, 2.0f / c.cb.viewport.x, -1.0f), sy = fma1((float)py + 0.5f, -(2.0f / c.cb.viewport.y), 1.0f);
            const float whx = fma1(sx, M.m[0][0], fma1(sy, M.m[1][0], M.m[2][0] + M.m[3][0])), why = fma1(sx, M.m[0][1], fma1(sy, M.m[1][1], M.m[2][1] + M.m[3][1]));
            const float whz = fma1(sx, M.m[0][2], fma1(sy, M.m[1][2], M.m[2][2] + M.m[3][2])), whw = fma1(sx, M.m[0][3], fma1(sy, M.m[1][3], M.m[2][3] + M.m[3][3]));
            const float iw = rcp(whw);
            const f3 viewDir = normalize3(c.cb.eyePt - f3{whx * iw, why * iw, whz * iw});
            const f3 d = -viewDir;
            const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
            int face; float ma;
            if (ax >= ay && ax >= az) { face = d.x > 0.0f ? 0 : 1; ma = ax; }
            else if (ay >= az) { face = d.y > 0.0f ? 2 : 3; ma = ay; }
            else { face = d.z > 0.0f ? 4 : 5; ma = az; }
            const float im = rcp(ma);
            float u, v;
            cube_face_uv({d.x * im, d.y * im, d.z * im}, face, u, v);
            const float fx = fma1(u, (float)S, -0.5f), fy = fma1(v, (float)S, -0.5f);
            const float flx = floorf(fx), fly = floorf(fy);
            const float wx = fx - flx, wy = fy - fly;
            const int i0 = (int)flx, j0 = (int)fly;
            f4 t[4];
            for (int k = 0; k < 4; ++k) {
                int f, i, j;
                cube_resolve_texel(S, face, i0 + (k & 1), j0 + (k >> 1), f, i, j);
                const uint16_t* h = &c.envCube[(((size_t)f * S + j) * S + i) * 4];
                t[k] = {f16_to_f32(h[0]), f16_to_f32(h[1]), f16_to_f32(h[2]), f16_to_f32(h[3])};
            }
            uint16_t* o = &c.color[pix * 4];
            o[0] = f32_to_f16(lerpf(lerpf(t[0].x, t[1].x, wx), lerpf(t[2].x, t[3].x, wx), wy));
            o[1] = f32_to_f16(lerpf(lerpf(t[0].y, t[1].y, wx), lerpf(t[2].y, t[3].y, wx), wy));
            o[2] = f32_to_f16(lerpf(lerpf(t[0].z, t[1].z, wx), lerpf(t[2].z, t[3].z, wx), wy));
            o[3] = 0;
        }
}

void tone_map(Caster& c)
{
    const int W = (int)c.d.width, H = (int)c.d.height;
    const std::vector<uint16_t>& src = c.taaHistory[c.frameParity];
    (void)H;
#pragma omp parallel for schedule(static)
    for (int i = (int)c.row0 * W; i < (int)c.row1 * W; ++i) {
        float r[3];
        for (int k = 0; k < 3; ++k) {
            float v = f16_to_f32(src[(size_t)i * 4 + k]);
            v *= g_min16.toneScale / (v + g_min16.toneBias);   // PSToneMap.hlsl:23
            v = pow125(fabsf(v));                    // :24 pow(abs(result), 1.25)
            // RGBA8_UNORM render-target write: saturate, scale, round to nearest
            float s = saturate(v);
            if (!(v == v)) s = 0.0f;
            r[k] = floorf(s * 255.0f + 0.5f);
        }
        uint8_t* o = &c.backBuffer[(size_t)i * 4];
        o[0] = (uint8_t)r[0]; o[1] = (uint8_t)r[1]; o[2] = (uint8_t)r[2]; o[3] = 255;
    }
}

// ------------------------------------------------------------------------------------------
// SH projection (XUSG CSSHCubeMap / CSSHSum / CSSHNormalize — binary only in the reference, so this
// follows the published DirectXSH SHProjectCubeMap algorithm the DXIL constants point to;
// PARITY UNPINNED). cubeRGB: 6 x size x size x 3 floats, D3D face order.
// ------------------------------------------------------------------------------------------
void sh_project(const float* cubeRGB, uint32_t size, float out27[27])
{
    double acc[9][3] = {}; double wsum = 0.0;
    const float fS = (float)size;
    for (uint32_t face = 0; face < 6; ++face)
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                const float u = ((float)x + 0.5f) / fS * 2.0f - 1.0f;
                const float v = ((float)y + 0.5f) / fS * 2.0f - 1.0f;
                const f3 p = get_local_pos((float)x, (float)y, face, fS);    // same face convention as the cube maps
                const f3 d = normalize3(p);
                const float t = 1.0f + u * u + v * v;
                const float w = 4.0f / (sqrtf(t) * t);                       // differential solid angle
                // XMSHEvalDirection basis, order 3
                const float Y[9] = {0.282094792f, -0.488602512f * d.y, 0.488602512f * d.z, -0.488602512f * d.x,
                                    1.092548431f * d.x * d.y, -1.092548431f * d.y * d.z, 0.946174695f * d.z * d.z - 0.315391565f,
                                    -1.092548431f * d.x * d.z, 0.546274215f * (d.x * d.x - d.y * d.y)};
                const float* px = &cubeRGB[(((size_t)face * size + y) * size + x) * 3];
                for (int i = 0; i < 9; ++i) for (int k = 0; k < 3; ++k) acc[i][k] += (double)(px[k] * Y[i] * w);
                wsum += w;
            }
    const double norm = 4.0 * 3.14159265358979323846 / wsum;               // CSSHNormalize: coeff * 4pi / sum(w)
    for (int i = 0; i < 9; ++i) for (int k = 0; k < 3; ++k) out27[i * 3 + k] = (float)(acc[i][k] * norm);
}

} // namespace mvo
