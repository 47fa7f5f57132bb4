// k_ray_march_l.cu — light-space transmittance pre-march into one volume's light map.
//
// Replaces CSRayMarchL (MultiVolumes/Content/Shaders/CSRayMarchL.hlsl:20-121; host side
// MultiRayCaster.cpp:1299-1327). One thread per light-map voxel; voxels with density >= 0.01 cast a
// shadow ray (and, with a light probe, an ambient-occlusion ray along the negative density gradient)
// through ALL N volumes. The volume processed is picked on the device by the cull kernel
// (visible[frameIdx % visibleCount]), so the frame needs no host read-back.
//
// The light map is an RGBA16F CUDA 3-D array written through a surface object and read back by the
// view march through a texture object. The reference's R11G11B10_FLOAT has no CUDA array format; the
// values are rounded to 11/11/10-bit floats before the store, which makes the texels identical to the
// reference's format (every such value is exactly representable in binary16).
#include "k_march.cuh"
#include <cstdio>
#include <cstdlib>

namespace mv {

namespace {

constexpr int kLightThreads = 128;
constexpr uint32_t kFull = 0xffffffffu;

// ShadowTest, RayMarch.hlsli:103-112: 2x2 PCF of the D16 shadow map (LINEAR_LESS_EQUAL comparison
// sampler, clamp addressing). No shadow map bound -> lit.
MV_D float shadow_test(const DeviceScene& s, const FrameCB& cb, V3 pos)
{
    const int S = (int)cb.shadowSize;
    if (S == 0) return 1.0f;
    const V4 ls = mul_p44(pos, cb.shadowViewProj);
    const float uvx = ls.x * 0.5f + 0.5f;
    const float uvy = 1.0f - (ls.y * 0.5f + 0.5f);
    const float ref = ls.z - 0.0027f;
    const float fx = uvx * (float)S - 0.5f, fy = uvy * (float)S - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float wx = fx - flx, wy = fy - fly;
    const int ix = (int)flx, iy = (int)fly;
    auto tap = [&](int x, int y) {
        x = min(max(x, 0), S - 1); y = min(max(y, 0), S - 1);
        const float d = (float)__ldg(s.shadow + (size_t)y * S + x) / 65535.0f;   // D16_UNORM
        return ref <= d ? 1.0f : 0.0f;
    };
    const float t00 = tap(ix, iy), t10 = tap(ix + 1, iy), t01 = tap(ix, iy + 1), t11 = tap(ix + 1, iy + 1);
    return lerp(lerp(t00, t10, wx), lerp(t01, t11, wx), wy);
}

// GetDensityGradient, RayMarch.hlsli:55-77: six taps at +-1 texel (SampleLevel with integer offsets)
MV_D V3 density_gradient(cudaTextureObject_t grid, V3 uvw, float invGrid, bool densityOnly)
{
    const float q0 = fetch_density(grid, uvw.x + -1.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 0.0f * invGrid, densityOnly);
    const float q1 = fetch_density(grid, uvw.x + 1.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 0.0f * invGrid, densityOnly);
    const float q2 = fetch_density(grid, uvw.x + 0.0f * invGrid, uvw.y + -1.0f * invGrid, uvw.z + 0.0f * invGrid, densityOnly);
    const float q3 = fetch_density(grid, uvw.x + 0.0f * invGrid, uvw.y + 1.0f * invGrid, uvw.z + 0.0f * invGrid, densityOnly);
    const float q4 = fetch_density(grid, uvw.x + 0.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + -1.0f * invGrid, densityOnly);
    const float q5 = fetch_density(grid, uvw.x + 0.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 1.0f * invGrid, densityOnly);
    return {q1 - q0, q3 - q2, q5 - q4};
}

// Volume-sharded storage (mv_create_sharded): the light map of a volume is marched by the rank that holds its source — the
// frame's light volume is picked on the device, so every rank launches the passes and the others leave at once — and the
// volumes of other ranks are read through their R16F density proxies.
MV_D bool light_volume_elsewhere(const DeviceScene& s, uint32_t volumeId)
{
    return s.shardVolumes && (s.volumeDescs[volumeId] & 0x3fffu) % s.shardWorld != s.shardRank;
}
MV_D bool density_in_x(const DeviceScene& s, uint32_t volTexId, bool densityOnly)
{
    return densityOnly || (s.srcIsProxy != nullptr && s.srcIsProxy[volTexId] != 0);
}

// Where the voxels of this launch go. One GPU: straight into the light volume's 3-D array (surface
// store). Sharded: the rank fills the z-slab [z0, z1) into the linear staging buffer of its own
// exchange block and, when peers are mapped, into every peer's (NVLink stores); mv_light_commit then
// moves the assembled staging buffer into the array on every rank.
struct LightTarget {
    uint32_t z0, z1;
    uint2* staging;                 // nullptr = surface store
    uint2* peerStaging[kMaxPeers];
    uint32_t numPeers;
};

#ifndef MV_LIGHT_MIN_BLOCKS
#define MV_LIGHT_MIN_BLOCKS 8
#endif
constexpr uint32_t kMaxSharedDirs = 1024;   // volumes whose light direction and bounding sphere are staged in shared memory

MV_D void store_light_voxel(const DeviceScene& s, const LightTarget& tgt, uint32_t volumeId, uint32_t L, uint32_t x, uint32_t y, uint32_t z, V3 value)
{
    const V4 out = {quantize_ufloat(value.x, 6), quantize_ufloat(value.y, 6), quantize_ufloat(value.z, 5), 0.0f};
    const uint2 packed = pack_half4(out);
    if (tgt.staging) {
        const size_t idx = ((size_t)z * L + y) * L + x;
        tgt.staging[idx] = packed;
        for (uint32_t p = 0; p < tgt.numPeers; ++p) if (tgt.peerStaging[p]) tgt.peerStaging[p][idx] = packed;
    } else surf3Dwrite(packed, s.lightSurf[volumeId], (int)(x * 8), (int)y, (int)z);                // :120
}

// Pass 1, one thread per light-map voxel (8x4x4 bricks): density at the voxel centre and the shadow-map
// test (CSRayMarchL.hlsl:36-51). A voxel below the density threshold casts no ray: its texel
// (shadow * lightColor + ambient, with ao = 1 and irradiance = 0 under a light probe) is written here.
// The others are appended, brick by brick and in thread order inside a brick, to the dense-voxel list
// that pass 2 marches with full warps.
template <bool kDensityOnly>
__global__ void __launch_bounds__(kLightThreads) k_light_classify(DeviceScene s, FrameCB cb, int volumeOverride, LightTarget tgt)
{
    __shared__ uint32_t s_warpCount[kLightThreads / 32];
    __shared__ uint32_t s_base;
    const uint32_t z0 = tgt.z0, z1 = tgt.z1;
    const uint32_t L = cb.lightGridSize;
    const uint32_t bricksX = (L + 7) / 8, bricksY = (L + 3) / 4;
    const uint32_t brick = blockIdx.x;
    const uint32_t bz = brick / (bricksX * bricksY), rem = brick - bz * bricksX * bricksY;
    const uint32_t by = rem / bricksX, bx = rem - by * bricksX;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = bx * 8 + (lane & 7), y = by * 4 + (lane >> 3), z = z0 + bz * 4 + warp;
    const bool active = x < L && y < L && z < z1;
    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;   // :29-33
    if (light_volume_elsewhere(s, volumeId)) return;

    bool dense = false;
    float shadow = 1.0f;
    if (active) {
        const float gridSize = (float)L;
        V3 rayOrigin = {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f,
                        ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};                               // :36
        const cudaTextureObject_t grid0 = s.volumeTex[s.volumeDescs[volumeId] & 0x3fffu];
        const V3 uvw = local_to_tex3d(rayOrigin);                                                   // :41
        const float density = fetch_density(grid0, uvw.x, uvw.y, uvw.z, kDensityOnly);                          // :45
        dense = density >= kZeroThreshold;                                                          // :46
        rayOrigin = mul_p43(rayOrigin, s.perObject[volumeId].world);                                // :48
        shadow = shadow_test(s, cb, rayOrigin);                                                     // :51
        if (!dense) {
            const V3 lightColor = {cb.lightColor[0] * cb.lightColor[3], cb.lightColor[1] * cb.lightColor[3], cb.lightColor[2] * cb.lightColor[3]};
            V3 ambient = {cb.ambient[0] * cb.ambient[3], cb.ambient[1] * cb.ambient[3], cb.ambient[2] * cb.ambient[3]};
            if (cb.hasSH) ambient = {1.0f * 0.0f, 1.0f * 0.0f, 1.0f * 0.0f};                        // ao * irradiance, :117
            store_light_voxel(s, tgt, volumeId, L, x, y, z,
                              V3{shadow * lightColor.x + ambient.x, shadow * lightColor.y + ambient.y, shadow * lightColor.z + ambient.z});
        }
    }
    const uint32_t bits = __ballot_sync(kFull, dense);
    if (lane == 0) s_warpCount[warp] = __popc(bits);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kLightThreads / 32; ++w) total += s_warpCount[w];
        s_base = total ? atomicAdd(&s.lists->lightDenseCount, total) : 0u;
        if (s.stats && total) atomicAdd(&s.stats->light_dense_voxels, (unsigned long long)total);
    }
    __syncthreads();
    if (dense) {
        uint32_t slot = s_base + __popc(bits & ((1u << lane) - 1u));
        for (uint32_t w = 0; w < warp; ++w) slot += s_warpCount[w];
        s.lightDense[slot] = make_uint2((z * L + y) * L + x, __float_as_uint(shadow));
    }
}

// Conservative early-out in world space: true only when the ray o + u d (u >= 0, |d| = 1) certainly
// misses the sphere (centre sph.xyz, squared radius sph.w, already inflated by 2 %) that bounds a
// volume's box. Whenever it returns true the exact local-space tests below report a miss as well, so
// it changes no result; it only spares the transform into the volume's space. NaNs fall through.
MV_D bool ray_misses_sphere_for_sure(V3 o, V3 d, float4 sph)
{
    const V3 oc = {sph.x - o.x, sph.y - o.y, sph.z - o.z};
    const float c2 = dot(oc, oc);
    const float r2 = sph.w + 1.0e-4f * c2;
    if (!(c2 > r2)) return false;
    const float b = dot(oc, d);
    if (b < 0.0f) return true;
    return c2 - b * b > r2;
}

// Per-volume tables of a CTA (shared memory): the world-space bounding sphere of the volume's box and
// the light direction in its local space. The light is directional (CSRayMarchL.hlsl:91-92):
// normalize(mul(g_lightPos.xyz, (float3x3)WorldI)) depends on the volume only, so it is evaluated once
// per CTA and volume instead of once per voxel and volume.
MV_D void stage_volume_tables(const DeviceScene& s, const FrameCB& cb, float4* s_sph, float* s_dirS, uint32_t nShared)
{
    const V3 lightPos = {cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]};
    for (uint32_t n = threadIdx.x; n < nShared; n += kLightThreads) {
        const PerObject* po = s.perObject + n;
        const V3 d = normalize(mul_v33(lightPos, po->worldI));
        s_dirS[3 * n] = d.x; s_dirS[3 * n + 1] = d.y; s_dirS[3 * n + 2] = d.z;
        const float* W = po->world;                          // rows: the box's half-axes a, b, c and its centre
        float r2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sb = (k & 1) ? -1.0f : 1.0f, sc = (k & 2) ? -1.0f : 1.0f;
            const V3 c = {W[0] + sb * W[3] + sc * W[6], W[1] + sb * W[4] + sc * W[7], W[2] + sb * W[5] + sc * W[8]};
            r2 = fmaxf(r2, dot(c, c));
        }
        s_sph[n] = make_float4(W[9], W[10], W[11], r2 * 1.02f);
    }
    __syncthreads();
}

// Clusters of kLightCluster consecutive volumes with one bounding sphere each (shared memory, after the per-volume
// tables): a voxel's ray that certainly misses the cluster's sphere certainly misses every member's, so the per-voxel
// loop over the N volumes steps over the whole cluster. With N = 512 (BASELINE.json configs[4]) that loop, not the
// marching, was most of the pass. The cluster sphere contains the members' (already inflated) spheres with 1 % to spare.
constexpr uint32_t kLightCluster = 16;
MV_D void stage_volume_clusters(const float4* s_sph, float4* s_clu, uint32_t nShared)
{
    const uint32_t numClusters = (nShared + kLightCluster - 1) / kLightCluster;
    for (uint32_t c = threadIdx.x; c < numClusters; c += kLightThreads) {
        const uint32_t n0 = c * kLightCluster, n1 = min(n0 + kLightCluster, nShared);
        V3 ctr = {0.0f, 0.0f, 0.0f};
        for (uint32_t n = n0; n < n1; ++n) { ctr.x += s_sph[n].x; ctr.y += s_sph[n].y; ctr.z += s_sph[n].z; }
        const float inv = 1.0f / (float)(n1 - n0);
        ctr = {ctr.x * inv, ctr.y * inv, ctr.z * inv};
        float radius = 0.0f;
        for (uint32_t n = n0; n < n1; ++n) {
            const V3 d = {s_sph[n].x - ctr.x, s_sph[n].y - ctr.y, s_sph[n].z - ctr.z};
            radius = fmaxf(radius, sqrtf(dot(d, d)) + sqrtf(s_sph[n].w));
        }
        radius *= 1.01f;
        s_clu[c] = make_float4(ctr.x, ctr.y, ctr.z, radius * radius);
    }
    __syncthreads();
}

MV_D V3 light_voxel_center(uint32_t voxel, uint32_t L, uint32_t& x, uint32_t& y, uint32_t& z)   // CSRayMarchL.hlsl:36
{
    x = voxel % L; y = (voxel / L) % L; z = voxel / (L * L);
    const float gridSize = (float)L;
    return {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f, ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};
}

MV_D V3 shadow_dir_local(const FrameCB& cb, const PerObject* po, const float* s_dirS, uint32_t n)
{
    if (n < kMaxSharedDirs) return V3{s_dirS[3 * n], s_dirS[3 * n + 1], s_dirS[3 * n + 2]};
    return normalize(mul_v33(V3{cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]}, po->worldI));      // :91-92 (directional)
}

MV_D void store_rec(LightRec* dst, const LightRec& r)
{
    const uint4* w = reinterpret_cast<const uint4*>(&r);
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = w[0]; d[1] = w[1]; d[2] = w[2]; d[3] = w[3];
}
MV_D LightRec load_rec(const LightRec* src)
{
    LightRec r;
    uint4* w = reinterpret_cast<uint4*>(&r);
    const uint4* p = reinterpret_cast<const uint4*>(src);
    w[0] = __ldg(p); w[1] = __ldg(p + 1); w[2] = __ldg(p + 2); w[3] = __ldg(p + 3);
    return r;
}

// What the volume loop needs to know about one voxel.
struct LightVoxel {
    V3 rayOrigin;      // world space (:48)
    V3 aoRayDir;       // world space, normalised (:65-75)
};

// The volume loop of CSRayMarchL.hlsl:77-110 for one voxel with the shadow ray's march left out: its
// outcome is summarised by castEnd, the first volume at which the shadow ray was no longer cast. Calls
// f(n, castShadow, localRayOrigin, rayDir) for every ambient-occlusion ray that hits its volume's box,
// in ascending volume order, with exactly the geometry (origin moved by the shadow ray's entry point,
// volume skipped when the shadow ray misses it) of the full loop.
template <class F>
MV_D void for_each_ao_ray(const DeviceScene& s, const FrameCB& cb, const LightVoxel& v, uint32_t castEnd, uint32_t maxRays,
                          const float4* s_sph, const float* s_dirS, uint32_t nShared, V3 lightDirW, F&& f)
{
    uint32_t found = 0;
    for (uint32_t n = 0; n < cb.numVolumes && found < maxRays; ++n) {
        const bool castShadow = n < castEnd;
        if (n < nShared && ray_misses_sphere_for_sure(v.rayOrigin, castShadow ? lightDirW : v.aoRayDir, s_sph[n])) continue;
        const PerObject* po = s.perObject + n;
        V3 localRayOrigin = mul_p43(v.rayOrigin, po->worldI);                                       // :83
        if (castShadow) {
            const V3 rayDir = shadow_dir_local(cb, po, s_dirS, n);
            if (ray_misses_box_for_sure(localRayOrigin, rayDir)) continue;
            if (!compute_ray_origin(localRayOrigin, rayDir)) continue;                              // :95
        }
        const V3 dirU = mul_v33(v.aoRayDir, po->worldI);                                            // :102
        if (ray_misses_box_for_sure(localRayOrigin, dirU)) continue;
        const V3 rayDir = normalize(dirU);
        if (!compute_ray_origin(localRayOrigin, rayDir)) continue;                                  // :103
        f(n, castShadow, localRayOrigin, rayDir);
        ++found;
    }
}

// Pass 2, persistent: warps pull 32 dense voxels at a time from the list. Per voxel the loop over the N
// volumes of CSRayMarchL.hlsl:77-110 is walked with its exact control flow, but only the shadow ray —
// whose transmittance chains from one volume to the next — is marched here. The ambient-occlusion ray
// of a (voxel, volume) pair depends on nothing but geometry and on whether the shadow ray was still
// being cast at that volume, so every AO ray that hits its box is deferred as an independent work item
// (passes 3-5) instead of lengthening this thread's dependent fetch chain: the pass's duration is
// bounded by its longest chain, not by its throughput (profiles/r01_notes.md). The pass counts the
// deferred rays per volume (they are marched sorted by volume, one texture per warp).
template <bool kDensityOnly>
__global__ void __launch_bounds__(kLightThreads, MV_LIGHT_MIN_BLOCKS) k_ray_march_l(DeviceScene s, FrameCB cb, int volumeOverride, LightTarget tgt)
{
    extern __shared__ float4 s_tab[];                       // [nShared] spheres, then [nShared] x 3 floats of directions
    if (light_volume_elsewhere(s, volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume)) return;
    const uint32_t N = cb.numVolumes, L = cb.lightGridSize;
    const uint32_t nShared = min(N, kMaxSharedDirs);
    const uint32_t numClusters = (nShared + kLightCluster - 1) / kLightCluster;
    float4* s_clu = s_tab + nShared;                        // [numClusters] cluster spheres
    float* s_dirS = reinterpret_cast<float*>(s_clu + numClusters);
    stage_volume_tables(s, cb, s_tab, s_dirS, nShared);
    stage_volume_clusters(s_tab, s_clu, nShared);
    const V3 lightDirW = normalize(V3{cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]});
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;   // :29-33
    const uint32_t count = s.lists->lightDenseCount;
    const cudaTextureObject_t grid0 = s.volumeTex[s.volumeDescs[volumeId] & 0x3fffu];
    const PerObject* po0 = s.perObject + volumeId;
    const float gStep = kMaxDist / (float)cb.maxLightSamples;                                        // RayMarch.hlsli:18
    uint32_t samples = 0;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s.lists->lightDenseCursor, 32u);
        base = __shfl_sync(kFull, base, 0);
        if (base >= count) break;
        const bool live = base + lane < count;
        uint32_t x = 0, y = 0, z = 0, numItems = 0;
        LightRec rec;
        rec.voxel = 0; rec.castEnd = N; rec.shadow = 1.0f; rec.ao = 1.0f;
        V3 rayOrigin = {0.0f, 0.0f, 0.0f}, aoRayDir = {0.0f, 0.0f, 0.0f};
        if (live) {
            const uint2 item = __ldg(s.lightDense + base + lane);
            rec.voxel = item.x;
            float shadow = __uint_as_float(item.y);
            rayOrigin = light_voxel_center(rec.voxel, L, x, y, z);
            const V3 uvw = local_to_tex3d(rayOrigin);                                               // :41
            rayOrigin = mul_p43(rayOrigin, po0->world);                                             // :48
            if (cb.hasSH) {                                                                         // :65-75
                aoRayDir = -density_gradient(grid0, uvw, 1.0f / (float)cb.gridSize, kDensityOnly);
                const bool nz = fabsf(aoRayDir.x) > 0.0f || fabsf(aoRayDir.y) > 0.0f || fabsf(aoRayDir.z) > 0.0f;
                aoRayDir = nz ? aoRayDir : rayOrigin;
                aoRayDir = mul_v33(aoRayDir, po0->world);
                aoRayDir = normalize(aoRayDir);
            }
            for (uint32_t n = 0; n < N; ++n) {                                                      // :77
                const bool castShadow = shadow >= kZeroThreshold;
                if (!castShadow) { rec.castEnd = min(rec.castEnd, n); if (!cb.hasSH) break; }
                // a missed shadow ray skips the volume altogether (:95 `continue`), otherwise only the AO ray is cast
                const V3 testDir = castShadow ? lightDirW : aoRayDir;
                if (n % kLightCluster == 0 && n < nShared && ray_misses_sphere_for_sure(rayOrigin, testDir, s_clu[n / kLightCluster])) {
                    n += min(kLightCluster, nShared - n) - 1;   // nothing in the cluster can be hit: `shadow` and the ray stay as they are
                    continue;
                }
                if (n < nShared && ray_misses_sphere_for_sure(rayOrigin, testDir, s_tab[n])) continue;
                const PerObject* po = s.perObject + n;
                V3 localRayOrigin = mul_p43(rayOrigin, po->worldI);                                 // :83
                if (castShadow) {
                    const V3 rayDir = shadow_dir_local(cb, po, s_dirS, n);
                    if (ray_misses_box_for_sure(localRayOrigin, rayDir)) continue;
                    if (!compute_ray_origin(localRayOrigin, rayDir)) continue;                      // :95
                    const uint32_t texId = s.volumeDescs[n] & 0x3fffu;
                    cast_light_ray(shadow, s.volumeTex[texId], localRayOrigin, rayDir, gStep, cb.maxLightSamples, density_in_x(s, texId, kDensityOnly), samples);
                }
                if (cb.hasSH) {                                                                     // :100-108, geometry only
                    const V3 dirU = mul_v33(aoRayDir, po->worldI);
                    if (ray_misses_box_for_sure(localRayOrigin, dirU)) continue;
                    if (!compute_ray_origin(localRayOrigin, normalize(dirU))) continue;
                    const uint32_t hit = n | (castShadow ? 0x80000000u : 0u);
#pragma unroll
                    for (uint32_t j = 0; j < kLightRecHits; ++j) if (j == numItems) rec.hits[j] = hit;
                    ++numItems;
                    // per-volume ray count, one atomic per group of lanes that are at the same volume together
                    const uint32_t grp = __match_any_sync(__activemask(), n);
                    if (lane == (uint32_t)__ffs(grp) - 1u) atomicAdd(s.lightSeg + n, (uint32_t)__popc(grp));
                }
            }
            rec.shadow = shadow;
        }
        if (!cb.hasSH) {
            if (live) {
                const V3 lightColor = {cb.lightColor[0] * cb.lightColor[3], cb.lightColor[1] * cb.lightColor[3], cb.lightColor[2] * cb.lightColor[3]};
                const V3 ambient = {cb.ambient[0] * cb.ambient[3], cb.ambient[1] * cb.ambient[3], cb.ambient[2] * cb.ambient[3]};
                store_light_voxel(s, tgt, volumeId, L, x, y, z,
                                  V3{rec.shadow * lightColor.x + ambient.x, rec.shadow * lightColor.y + ambient.y, rec.shadow * lightColor.z + ambient.z});
            }
            continue;
        }
        // reserve the warp's AO factors in one atomic; a voxel's factors are contiguous, in ascending volume order
        uint32_t incl = numItems;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += t; }
        const uint32_t warpTotal = __shfl_sync(kFull, incl, 31);
        uint32_t warpBase = 0;
        if (lane == 0 && warpTotal) warpBase = atomicAdd(&s.lists->lightResultCount, warpTotal);
        warpBase = __shfl_sync(kFull, warpBase, 0);
        if (live) {
            rec.itemBase = warpBase + incl - numItems; rec.itemCount = numItems;
            rec.aoDir[0] = aoRayDir.x; rec.aoDir[1] = aoRayDir.y; rec.aoDir[2] = aoRayDir.z; rec.pad = 0;
            store_rec(s.lightRecs + base + lane, rec);
        }
    }
    if (s.stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) samples += __shfl_xor_sync(kFull, samples, d);
        if (lane == 0 && samples) atomicAdd(&s.stats->light_samples, (unsigned long long)samples);
    }
}

constexpr uint32_t kItemSentinel = 0xffffffffu;

// Pass 3, one CTA: exclusive scan of the per-volume AO-ray counts, each rounded up to a multiple of 32,
// into segment offsets (so that a warp of pass 5 never straddles two volumes); the padding slots are
// marked, the counts reset for the next frame, and the frame falls back to inline marching when the
// rays do not fit the buffers.
// The segment of the light volume itself goes last: its rays are the bulk and L2-warm (short steps), whereas the
// rays through other volumes miss to HBM on nearly every step; started first, their long chains overlap the bulk
// instead of forming the tail of pass 5.
__global__ void __launch_bounds__(1024) k_light_scan(DeviceScene s, FrameCB cb, int volumeOverride)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running;
    const uint32_t N = cb.numVolumes, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ownVolume = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;
    if (light_volume_elsewhere(s, ownVolume)) return;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (uint32_t tile = 0; tile < N; tile += 1024) {
        const uint32_t n = tile + threadIdx.x;
        const uint32_t c = (n < N && n != ownVolume) ? s.lightSeg[n] : 0u;
        const uint32_t padded = (c + 31u) & ~31u;
        uint32_t incl = padded;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = s_running;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        const uint32_t start = before + incl - padded;
        if (n < N && n != ownVolume) {
            s.lightSeg[N + n] = start;
            s.lightSeg[n] = 0;
            for (uint32_t q = start + c; q < start + padded; ++q)
                if (q < s.lightItemCapacity) s.lightItems[q] = make_uint4(kItemSentinel, 0u, 0u, 0u);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_running = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && ownVolume < N) {
        const uint32_t c = s.lightSeg[ownVolume], padded = (c + 31u) & ~31u, start = s_running;
        s.lightSeg[N + ownVolume] = start;
        s.lightSeg[ownVolume] = 0;
        for (uint32_t q = start + c; q < start + padded; ++q)
            if (q < s.lightItemCapacity) s.lightItems[q] = make_uint4(kItemSentinel, 0u, 0u, 0u);
        s_running = start + padded;
    }
    if (threadIdx.x == 0) {
        s.lists->lightItemCount = s_running;
        s.lists->lightOverflow = (s_running > s.lightItemCapacity || s.lists->lightResultCount > s.lightItemCapacity) ? 1u : 0u;
    }
}

// Pass 4, persistent over the dense voxels: scatter every deferred AO ray into its volume's segment of
// the item list. (Overflow frame: march the voxel's AO rays inline instead, CSRayMarchL.hlsl:100-108.)
template <bool kDensityOnly>
__global__ void __launch_bounds__(kLightThreads) k_light_emit(DeviceScene s, FrameCB cb, int volumeOverride)
{
    extern __shared__ float4 s_tab[];
    if (light_volume_elsewhere(s, volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume)) return;
    const uint32_t N = cb.numVolumes, L = cb.lightGridSize;
    const uint32_t nShared = min(N, kMaxSharedDirs);
    float* s_dirS = reinterpret_cast<float*>(s_tab + nShared);
    stage_volume_tables(s, cb, s_tab, s_dirS, nShared);
    const V3 lightDirW = normalize(V3{cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]});
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;
    const uint32_t count = s.lists->lightDenseCount;
    const bool overflow = s.lists->lightOverflow != 0;
    const PerObject* po0 = s.perObject + volumeId;
    const float gStep = kMaxDist / (float)cb.maxLightSamples;
    uint32_t samples = 0;

    auto emit = [&](uint32_t recIdx, uint32_t hit, uint32_t resultIdx) {
        const uint32_t n = hit & 0x7fffffffu;
        const uint32_t grp = __match_any_sync(__activemask(), n);
        const uint32_t leader = (uint32_t)__ffs(grp) - 1u;
        uint32_t slot = 0;
        if (lane == leader) slot = atomicAdd(s.lightSeg + N + n, (uint32_t)__popc(grp));
        slot = __shfl_sync(grp, slot, leader) + __popc(grp & ((1u << lane) - 1u));
        s.lightItems[slot] = make_uint4(recIdx, hit, resultIdx, 0u);
    };

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s.lists->lightEmitCursor, 32u);
        base = __shfl_sync(kFull, base, 0);
        if (base >= count) break;
        if (base + lane >= count) continue;
        const uint32_t recIdx = base + lane;
        const LightRec rec = load_rec(s.lightRecs + recIdx);
        if (!overflow && rec.itemCount <= kLightRecHits) {
            for (uint32_t j = 0; j < rec.itemCount; ++j) {
                uint32_t hit = 0;
#pragma unroll
                for (uint32_t k = 0; k < kLightRecHits; ++k) if (k == j) hit = rec.hits[k];
                emit(recIdx, hit, rec.itemBase + j);
            }
            continue;
        }
        uint32_t x, y, z;
        LightVoxel v;
        v.rayOrigin = mul_p43(light_voxel_center(rec.voxel, L, x, y, z), po0->world);               // :36, :48
        v.aoRayDir = {rec.aoDir[0], rec.aoDir[1], rec.aoDir[2]};
        if (!overflow) {
            uint32_t j = 0;
            for_each_ao_ray(s, cb, v, rec.castEnd, rec.itemCount, s_tab, s_dirS, nShared, lightDirW,
                            [&](uint32_t n, bool castShadow, V3, V3) { emit(recIdx, n | (castShadow ? 0x80000000u : 0u), rec.itemBase + j); ++j; });
        } else {
            float ao = 1.0f;
            for_each_ao_ray(s, cb, v, rec.castEnd, rec.itemCount, s_tab, s_dirS, nShared, lightDirW,
                            [&](uint32_t n, bool, V3 localRayOrigin, V3 rayDir) {
                                float transm = 1.0f;
                                const uint32_t texId = s.volumeDescs[n] & 0x3fffu;
                                cast_light_ray(transm, s.volumeTex[texId], localRayOrigin, rayDir, gStep, cb.maxLightSamples, density_in_x(s, texId, kDensityOnly), samples);
                                ao *= (n == volumeId) ? transm : pow025(saturate(transm + 0.5f));   // :107
                            });
            s.lightRecs[recIdx].ao = ao;
            s.lightRecs[recIdx].itemCount = 0;
        }
    }
    if (s.stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) samples += __shfl_xor_sync(kFull, samples, d);
        if (lane == 0 && samples) atomicAdd(&s.stats->light_samples, (unsigned long long)samples);
    }
}

// Pass 5, persistent: one thread per deferred ambient-occlusion ray (CSRayMarchL.hlsl:100-108), a warp's
// 32 rays all through the same volume. The ray is rebuilt from the voxel, the volume and the
// cast-shadow flag with the operations of pass 2.
#ifndef MV_LIGHT_AO_MIN_BLOCKS
#define MV_LIGHT_AO_MIN_BLOCKS 8
#endif
template <bool kDensityOnly>
__global__ void __launch_bounds__(kLightThreads, MV_LIGHT_AO_MIN_BLOCKS) k_light_ao(DeviceScene s, FrameCB cb, int volumeOverride)
{
    const uint32_t L = cb.lightGridSize;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;
    if (light_volume_elsewhere(s, volumeId)) return;
    const uint32_t count = s.lists->lightOverflow ? 0u : s.lists->lightItemCount;   // a multiple of 32
    const PerObject* po0 = s.perObject + volumeId;
    const float gStep = kMaxDist / (float)cb.maxLightSamples;
    uint32_t samples = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s.lists->lightItemCursor, 32u);
        base = __shfl_sync(kFull, base, 0);
        if (base >= count) break;
        const uint4 item = __ldg(s.lightItems + base + lane);
        const bool valid = item.x != kItemSentinel;
        const uint32_t validMask = __ballot_sync(kFull, valid);
        if (!validMask) continue;
        const uint32_t n = __shfl_sync(kFull, item.y & 0x7fffffffu, __ffs(validMask) - 1);          // the segment's volume
        if (!valid) continue;
        const bool castShadow = (item.y >> 31) != 0;
        const LightRec* rec = s.lightRecs + item.x;
        uint32_t x, y, z;
        V3 rayOrigin = light_voxel_center(__ldg(&rec->voxel), L, x, y, z);
        rayOrigin = mul_p43(rayOrigin, po0->world);                                                 // :48
        const V3 aoRayDir = {__ldg(&rec->aoDir[0]), __ldg(&rec->aoDir[1]), __ldg(&rec->aoDir[2])};
        const PerObject* po = s.perObject + n;
        V3 localRayOrigin = mul_p43(rayOrigin, po->worldI);                                         // :83
        if (castShadow) {                                                                           // the shadow ray's entry point moved the origin (:95)
            const V3 shadowDir = normalize(mul_v33(V3{cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]}, po->worldI));
            compute_ray_origin(localRayOrigin, shadowDir);
        }
        const V3 rayDir = normalize(mul_v33(aoRayDir, po->worldI));
        compute_ray_origin(localRayOrigin, rayDir);
        float transm = 1.0f;
        const uint32_t texId = s.volumeDescs[n] & 0x3fffu;
        cast_light_ray(transm, s.volumeTex[texId], localRayOrigin, rayDir, gStep, cb.maxLightSamples, density_in_x(s, texId, kDensityOnly), samples);
        s.lightItemResults[item.z] = (n == volumeId) ? transm : pow025(saturate(transm + 0.5f));    // :107
    }
    if (s.stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) samples += __shfl_xor_sync(kFull, samples, d);
        if (lane == 0 && samples) atomicAdd(&s.stats->light_samples, (unsigned long long)samples);
    }
}

// Pass 6: fold each dense voxel's AO factors in ascending volume order and write its texel (:112-120).
__global__ void __launch_bounds__(256) k_light_finalize(DeviceScene s, FrameCB cb, int volumeOverride, LightTarget tgt)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= s.lists->lightDenseCount) return;
    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;
    if (light_volume_elsewhere(s, volumeId)) return;
    // the first two 16-byte words of the record: {voxel, itemBase, itemCount, shadow}, {aoDir, ao}
    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(s.lightRecs + i));
    const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(s.lightRecs + i) + 1);
    float ao = __uint_as_float(w1.w);
    const uint32_t itemBase = w0.y, itemCount = w0.z;
    for (uint32_t j = 0; j < itemCount; ++j) ao *= __ldg(s.lightItemResults + itemBase + j);
    const V3 aoRayDir = {__uint_as_float(w1.x), __uint_as_float(w1.y), __uint_as_float(w1.z)};
    const V3 irradiance = evaluate_sh_irradiance(cb.sh, normalize(aoRayDir));                       // GetIrradiance
    const uint32_t L = cb.lightGridSize;
    uint32_t x, y, z;
    light_voxel_center(w0.x, L, x, y, z);
    const float shadow = __uint_as_float(w0.w);
    const V3 lightColor = {cb.lightColor[0] * cb.lightColor[3], cb.lightColor[1] * cb.lightColor[3], cb.lightColor[2] * cb.lightColor[3]};
    const V3 ambient = {ao * irradiance.x, ao * irradiance.y, ao * irradiance.z};                   // :117
    store_light_voxel(s, tgt, volumeId, L, x, y, z,
                      V3{shadow * lightColor.x + ambient.x, shadow * lightColor.y + ambient.y, shadow * lightColor.z + ambient.z});
}

} // namespace

void launch_ray_march_light(Caster& c, int volumeOverride)
{
    const uint32_t L = c.d.light_grid_size;
    LightTarget tgt{};
    tgt.z0 = 0; tgt.z1 = L;
    if (c.shardWorld > 1 && !c.shardVolumes) {       // (volume-sharded storage: the whole light map, by the volume's owner)
        const uint32_t slab = (L + c.shardWorld - 1) / c.shardWorld;
        tgt.z0 = min(L, c.shardRank * slab); tgt.z1 = min(L, tgt.z0 + slab);
        tgt.staging = c.dLightStaging;
        tgt.numPeers = c.peersMapped ? c.shardWorld : 0;
        for (uint32_t p = 0; p < tgt.numPeers; ++p)
            tgt.peerStaging[p] = (p == c.shardRank) ? nullptr
                : reinterpret_cast<uint2*>(c.peerBlock[p] + (c.stagingParity ? c.layout.light_staging2_offset : c.layout.light_staging_offset));
    }
    if (c.lightToStaging) tgt.staging = c.dLightStaging;   // pipelined frame: the main stream commits it (mv_api.cu)
    if (tgt.z1 <= tgt.z0) return;
    const uint32_t voxels = L * L * (tgt.z1 - tgt.z0);
    const uint32_t bricks = ((L + 7) / 8) * ((L + 3) / 4) * ((tgt.z1 - tgt.z0 + 3) / 4);
    const uint32_t nShared = min(c.d.num_volumes, kMaxSharedDirs);
    const size_t smem = (size_t)nShared * (sizeof(float4) + 3 * sizeof(float));
    const size_t smemMarch = smem + (size_t)((nShared + kLightCluster - 1) / kLightCluster) * sizeof(float4);   // + cluster spheres (k_ray_march_l)
    static int perSM = 0, perSMAo = 0, perSMEmit = 0;
    if (!perSM) {
        const size_t smemMax = (size_t)kMaxSharedDirs * (sizeof(float4) + 3 * sizeof(float));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_ray_march_l<false>, kLightThreads, smemMax + (kMaxSharedDirs / kLightCluster) * sizeof(float4));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMEmit, k_light_emit<false>, kLightThreads, smemMax);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMAo, k_light_ao<false>, kLightThreads, 0);
        perSM = max(perSM, 1); perSMEmit = max(perSMEmit, 1); perSMAo = max(perSMAo, 1);
    }
    // MV_LIGHT_TIMING=1 (tuning aid): CUDA events around each pass, averages printed every 64 launches
    static const bool timing = getenv("MV_LIGHT_TIMING") != nullptr;
    static cudaEvent_t tev[7]; static float tacc[6]; static int tcount = 0;
    if (timing && !tev[0]) for (auto& e : tev) cudaEventCreate(&e);
    auto mark = [&](int i) { if (timing) cudaEventRecord(tev[i], c.stream); };
    mark(0);
    const bool densityOnly = (c.d.flags & MV_FLAG_DENSITY_ONLY) != 0;   // same register budgets in both instantiations
    if (densityOnly) k_light_classify<true><<<bricks, kLightThreads, 0, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
    else k_light_classify<false><<<bricks, kLightThreads, 0, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
    mark(1);
    if (densityOnly) k_ray_march_l<true><<<c.smCount * perSM, kLightThreads, smemMarch, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
    else k_ray_march_l<false><<<c.smCount * perSM, kLightThreads, smemMarch, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
    mark(2);
    if (!c.cb.hasSH) return;
    k_light_scan<<<1, 1024, 0, c.stream>>>(c.scene(), c.cb, volumeOverride);
    mark(3);
    if (densityOnly) k_light_emit<true><<<c.smCount * perSMEmit, kLightThreads, smem, c.stream>>>(c.scene(), c.cb, volumeOverride);
    else k_light_emit<false><<<c.smCount * perSMEmit, kLightThreads, smem, c.stream>>>(c.scene(), c.cb, volumeOverride);
    mark(4);
    if (densityOnly) k_light_ao<true><<<c.smCount * perSMAo, kLightThreads, 0, c.stream>>>(c.scene(), c.cb, volumeOverride);
    else k_light_ao<false><<<c.smCount * perSMAo, kLightThreads, 0, c.stream>>>(c.scene(), c.cb, volumeOverride);
    mark(5);
    k_light_finalize<<<(voxels + 255) / 256, 256, 0, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
    mark(6);
    if (timing) {
        cudaEventSynchronize(tev[6]);
        for (int i = 0; i < 6; ++i) { float ms = 0; cudaEventElapsedTime(&ms, tev[i], tev[i + 1]); tacc[i] += ms; }
        if (++tcount % 64 == 0) {
            fprintf(stderr, "[light] classify %.4f shadow %.4f scan %.4f emit %.4f ao %.4f finalize %.4f ms\n",
                    tacc[0] / 64, tacc[1] / 64, tacc[2] / 64, tacc[3] / 64, tacc[4] / 64, tacc[5] / 64);
            for (auto& t : tacc) t = 0;
        }
    }
}


namespace {
// Work-graph order (MultiRayCaster.cpp:358-362: rayMarchL runs before the graph that culls): the light march of a frame
// picks its volume from the visible list the PREVIOUS frame's cull left behind (CSRayMarchL.hlsl:29-33 reads
// g_roVisibleVolumes / its counter as they are) and needs its work counters reset, which the cull otherwise does.
__global__ void k_pick_light_volume(FrameLists* cur, const FrameLists* prev, const uint32_t* prevVisible, uint32_t frameIdx, uint32_t N)
{
    const uint32_t count = prev->visibleCount;
    cur->lightVolume = count ? prevVisible[frameIdx % count] : frameIdx % N;
    cur->lightDenseCount = 0; cur->lightDenseCursor = 0;
    cur->lightItemCount = 0; cur->lightItemCursor = 0;
    cur->lightResultCount = 0; cur->lightEmitCursor = 0;
    cur->lightOverflow = 0;
}

} // namespace

void launch_pick_light_volume(Caster& c)
{
    FrameLists* cur = reinterpret_cast<FrameLists*>(c.dLists);
    const unsigned char* prevBase = c.dLists2[c.listParity ^ 1u];
    k_pick_light_volume<<<1, 1, 0, c.stream>>>(cur, reinterpret_cast<const FrameLists*>(prevBase),
                                               reinterpret_cast<const uint32_t*>(prevBase + sizeof(FrameLists)), c.cb.frameIdx, c.d.num_volumes);
}

} // namespace mv
