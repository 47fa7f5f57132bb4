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
MV_D V3 density_gradient(cudaTextureObject_t grid, V3 uvw, float invGrid)
{
    const float q0 = tex3D<float4>(grid, uvw.x + -1.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 0.0f * invGrid).w;
    const float q1 = tex3D<float4>(grid, uvw.x + 1.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 0.0f * invGrid).w;
    const float q2 = tex3D<float4>(grid, uvw.x + 0.0f * invGrid, uvw.y + -1.0f * invGrid, uvw.z + 0.0f * invGrid).w;
    const float q3 = tex3D<float4>(grid, uvw.x + 0.0f * invGrid, uvw.y + 1.0f * invGrid, uvw.z + 0.0f * invGrid).w;
    const float q4 = tex3D<float4>(grid, uvw.x + 0.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + -1.0f * invGrid).w;
    const float q5 = tex3D<float4>(grid, uvw.x + 0.0f * invGrid, uvw.y + 0.0f * invGrid, uvw.z + 1.0f * invGrid).w;
    return {q1 - q0, q3 - q2, q5 - q4};
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

constexpr uint32_t kMaxSharedDirs = 1024;   // volumes whose light direction is staged in shared memory

__global__ void __launch_bounds__(kLightThreads) k_ray_march_l(DeviceScene s, FrameCB cb, int volumeOverride, LightTarget tgt)
{
    // The light is directional (CSRayMarchL.hlsl:91-92): normalize(mul(g_lightPos.xyz, (float3x3)WorldI)) depends on
    // the volume only. It is evaluated once per CTA and volume here instead of once per voxel and volume.
    extern __shared__ float s_dirS[];
    {
        const V3 lightPos = {cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]};
        const uint32_t nShared = min(cb.numVolumes, kMaxSharedDirs);
        for (uint32_t n = threadIdx.x; n < nShared; n += kLightThreads) {
            const V3 d = normalize(mul_v33(lightPos, s.perObject[n].worldI));
            s_dirS[3 * n] = d.x; s_dirS[3 * n + 1] = d.y; s_dirS[3 * n + 2] = d.z;
        }
        __syncthreads();
    }
    const uint32_t z0 = tgt.z0, z1 = tgt.z1;
    const uint32_t L = cb.lightGridSize, N = cb.numVolumes;
    // 8x4x4 voxel bricks: a warp is an 8x4 slice, neighbouring rays stay coherent in the texture cache
    const uint32_t bricksX = (L + 7) / 8, bricksY = (L + 3) / 4;
    const uint32_t brick = blockIdx.x;
    const uint32_t bz = brick / (bricksX * bricksY), rem = brick - bz * bricksX * bricksY;
    const uint32_t by = rem / bricksX, bx = rem - by * bricksX;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = bx * 8 + (lane & 7), y = by * 4 + (lane >> 3), z = z0 + bz * 4 + warp;
    const bool active = x < L && y < L && z < z1;

    const uint32_t volumeId = volumeOverride >= 0 ? (uint32_t)volumeOverride : s.lists->lightVolume;   // :29-33
    uint32_t dense = 0, samples = 0;
    if (active) {
        const float gridSize = (float)L;
        V3 rayOrigin = {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f,
                        ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};                               // :36
        const uint32_t volTexId0 = s.volumeDescs[volumeId] & 0x3fffu;
        const cudaTextureObject_t grid0 = s.volumeTex[volTexId0];
        const V3 uvw = local_to_tex3d(rayOrigin);                                                   // :41
        const PerObject* po0 = s.perObject + volumeId;
        const float density = tex3D<float4>(grid0, uvw.x, uvw.y, uvw.z).w;                          // :45
        const bool hasDensity = density >= kZeroThreshold;                                          // :46
        rayOrigin = mul_p43(rayOrigin, po0->world);                                                 // :48
        float shadow = shadow_test(s, cb, rayOrigin);                                               // :51
        float ao = 1.0f;
        V3 irradiance = {0.0f, 0.0f, 0.0f};
        if (hasDensity) {
            ++dense;
            const float maxDist = 2.0f * sqrtf(3.0f);
            const float gStep = maxDist / (float)cb.maxLightSamples;                                // RayMarch.hlsli:18
            V3 aoRayDir = {0.0f, 0.0f, 0.0f};
            if (cb.hasSH) {                                                                         // :65-75
                aoRayDir = -density_gradient(grid0, uvw, 1.0f / (float)cb.gridSize);
                const bool nz = fabsf(aoRayDir.x) > 0.0f || fabsf(aoRayDir.y) > 0.0f || fabsf(aoRayDir.z) > 0.0f;
                aoRayDir = nz ? aoRayDir : rayOrigin;
                aoRayDir = mul_v33(aoRayDir, po0->world);
                aoRayDir = normalize(aoRayDir);
                irradiance = evaluate_sh_irradiance(cb.sh, normalize(aoRayDir));                    // GetIrradiance
            }
            const V3 lightPos = {cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]};
            for (uint32_t n = 0; n < N; ++n) {                                                      // :77
                const PerObject* po = s.perObject + n;
                const cudaTextureObject_t grid = s.volumeTex[s.volumeDescs[n] & 0x3fffu];
                V3 localRayOrigin = mul_p43(rayOrigin, po->worldI);                                 // :83
                if (shadow >= kZeroThreshold) {
                    const V3 rayDir = n < kMaxSharedDirs ? V3{s_dirS[3 * n], s_dirS[3 * n + 1], s_dirS[3 * n + 2]}
                                                         : normalize(mul_v33(lightPos, po->worldI));   // :91-92 (directional)
                    if (ray_misses_box_for_sure(localRayOrigin, rayDir)) continue;                  // most volumes: far off the ray
                    if (!compute_ray_origin(localRayOrigin, rayDir)) continue;                      // :95
                    cast_light_ray(shadow, grid, localRayOrigin, rayDir, gStep, cb.maxLightSamples, samples);
                }
                if (cb.hasSH) {                                                                     // :100-108
                    const V3 dirU = mul_v33(aoRayDir, po->worldI);
                    if (ray_misses_box_for_sure(localRayOrigin, dirU)) continue;
                    const V3 rayDir = normalize(dirU);
                    if (!compute_ray_origin(localRayOrigin, rayDir)) continue;
                    float transm = 1.0f;
                    cast_light_ray(transm, grid, localRayOrigin, rayDir, gStep, cb.maxLightSamples, samples);
                    ao *= (n == volumeId) ? transm : pow025(saturate(transm + 0.5f));
                }
            }
        }
        const V3 lightColor = {cb.lightColor[0] * cb.lightColor[3], cb.lightColor[1] * cb.lightColor[3], cb.lightColor[2] * cb.lightColor[3]};
        V3 ambient = {cb.ambient[0] * cb.ambient[3], cb.ambient[1] * cb.ambient[3], cb.ambient[2] * cb.ambient[3]};
        if (cb.hasSH) ambient = {ao * irradiance.x, ao * irradiance.y, ao * irradiance.z};         // :117
        const V4 out = {quantize_ufloat(shadow * lightColor.x + ambient.x, 6), quantize_ufloat(shadow * lightColor.y + ambient.y, 6),
                        quantize_ufloat(shadow * lightColor.z + ambient.z, 5), 0.0f};
        const uint2 packed = pack_half4(out);
        if (tgt.staging) {
            const size_t idx = ((size_t)z * L + y) * L + x;
            tgt.staging[idx] = packed;
            for (uint32_t p = 0; p < tgt.numPeers; ++p) if (tgt.peerStaging[p]) tgt.peerStaging[p][idx] = packed;
        } else surf3Dwrite(packed, s.lightSurf[volumeId], (int)(x * 8), (int)y, (int)z);           // :120
    }
    if (s.stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            dense += __shfl_xor_sync(kFull, dense, d);
            samples += __shfl_xor_sync(kFull, samples, d);
        }
        if (lane == 0 && (dense | samples)) {
            atomicAdd(&s.stats->light_dense_voxels, (unsigned long long)dense);
            atomicAdd(&s.stats->light_samples, (unsigned long long)samples);
        }
    }
}

} // namespace

void launch_ray_march_light(Caster& c, int volumeOverride)
{
    const uint32_t L = c.d.light_grid_size;
    LightTarget tgt{};
    tgt.z0 = 0; tgt.z1 = L;
    if (c.shardWorld > 1) {
        const uint32_t slab = (L + c.shardWorld - 1) / c.shardWorld;
        tgt.z0 = min(L, c.shardRank * slab); tgt.z1 = min(L, tgt.z0 + slab);
        tgt.staging = c.dLightStaging;
        tgt.numPeers = c.peersMapped ? c.shardWorld : 0;
        for (uint32_t p = 0; p < tgt.numPeers; ++p)
            tgt.peerStaging[p] = (p == c.shardRank) ? nullptr : reinterpret_cast<uint2*>(c.peerBlock[p] + c.layout.light_staging_offset);
    }
    if (tgt.z1 <= tgt.z0) return;
    const uint32_t bricks = ((L + 7) / 8) * ((L + 3) / 4) * ((tgt.z1 - tgt.z0 + 3) / 4);
    const size_t smem = (size_t)min(c.d.num_volumes, kMaxSharedDirs) * 3 * sizeof(float);
    k_ray_march_l<<<bricks, kLightThreads, smem, c.stream>>>(c.scene(), c.cb, volumeOverride, tgt);
}

} // namespace mv
