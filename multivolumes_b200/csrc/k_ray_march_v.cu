// k_ray_march_v.cu — cube-map-space view march on the interior faces of every cube-map volume.
//
// Replaces CSRayMarchV (MultiVolumes/Content/Shaders/CSRayMarch.hlsl:77-158) and the ExecuteIndirect
// that launches it (MultiRayCaster.cpp:1329-1368). Differences in *shape*, not in arithmetic:
//   * exactly (G >> mip)^2 texels per visible face are marched (the reference's indirect dispatch
//     covers the mip-0 size and drops the surplus threads' writes; its work-graph variant
//     LibRayMarch.hlsl:120-121 launches the exact grid, as here);
//   * one persistent launch: warps pull 8x4-texel tiles from a device-side cursor that the cull kernel
//     reset (no host round trip for the indirect arguments), tiles of one face are consecutive work
//     items so that co-resident warps walk neighbouring rays through the same texture-cache lines;
//   * the per-volume constants (PerObject record, attributes, output addresses) are staged once per
//     tile in shared memory;
//   * with peers mapped (multi-GPU), each texel is stored straight into every peer's cube-map arena
//     over NVLink from this kernel — the all-gather of the per-volume cube maps is fused into the march;
//   * kFusedCull (mv_render_work_graph): the cull runs inside this launch, on CTA 0, and the other CTAs pick up the
//     lists it wrote once it publishes the frame's serial number — the work-graph path of the reference
//     (LibRayMarch.hlsl:39-134: VolumeCull node -> RayMarch node in one DispatchGraph, MultiRayCaster.cpp:1370-1438).
#include "k_march.cuh"
#include "k_cull.cuh"
#include <cstdlib>

namespace mv {

namespace {

// 6 CTAs x 256 threads / SM = 48 resident warps (<= 40 registers): the loop is latency-bound on a dependent texture round
// trip per step, so resident warps are what buys throughput. Round 1 measured no gain beyond 40 warps on cfg 2 with the
// kernel of that time; on the round-2 kernel (empty-space bricks, light fetch issued with the density fetch) and cfg 4:
// 4 / 5 / 6 / 7 / 8 CTAs = 0.947 / 0.837 / 0.780 / 0.863 / 0.863 ms (same-box A/B, profiles/r02_notes.md section 9).
#ifndef MV_MARCH_MIN_BLOCKS
#define MV_MARCH_MIN_BLOCKS 6
#endif
constexpr int kMarchThreads = 256;
constexpr int kMarchWarps = kMarchThreads / 32;
constexpr uint32_t kFull = 0xffffffffu;

struct TileConst {                 // per-warp shared-memory staging
    float po[56];                  // PerObject: wvp, wvpi, worldI, world
    float eyeL[3];
    uint32_t pad;
};

// GetLocalPos, CSRayMarch.hlsl:28-53
MV_D V3 get_local_pos(float px, float py, uint32_t slice, float gridSize)
{
    const float x = (px + 0.5f) / gridSize * 2.0f - 1.0f;
    float y = (py + 0.5f) / gridSize * 2.0f - 1.0f;
    y = -y;
    switch (slice) {
    case 0: return {1.0f, y, -x};
    case 1: return {-1.0f, y, x};
    case 2: return {x, 1.0f, -y};
    case 3: return {x, -1.0f, y};
    case 4: return {x, y, 1.0f};
    default: return {-x, y, -1.0f};
    }
}

MV_D uint32_t nth_set_bit(uint32_t mask, uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i) mask &= mask - 1;
    return __ffs(mask) - 1;
}

// Loads of the lists the cull wrote: through the read-only path when an earlier launch wrote them, from L2 (ld.cg) when
// CTA 0 of this very launch did.
template <bool kFusedCull, class T>
MV_D T ld_list(const T* p) { return kFusedCull ? __ldcg(p) : __ldg(p); }

template <bool kStats, bool kDensityOnly, bool kFusedCull>
__global__ void __launch_bounds__(kMarchThreads, MV_MARCH_MIN_BLOCKS) k_ray_march_v(DeviceScene s, FrameCB cb, uint32_t serial, uint32_t phase)
{
    __shared__ TileConst s_tc[kMarchWarps];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TileConst& tc = s_tc[warp];
    if (kFusedCull) {
        // CTA 0 is the cull node; it releases the frame's lists by storing the serial number after a fence. The launch is
        // COOPERATIVE (cudaLaunchCooperativeKernel: the runtime refuses a grid that is not co-resident), so the spinning
        // CTAs cannot starve CTA 0 whatever the dispatch order; the spin is bounded all the same (about 2 s).
        volatile uint32_t* ready = &s.lists->cullSerial;
        if (blockIdx.x == 0) {
            cull_body<kMarchThreads>(s, cb, false);
            __threadfence();                                // every writer: its list entries are visible device-wide ...
            __syncthreads();                                // ... before thread 0 learns that they were written
            if (threadIdx.x == 0) { __threadfence(); *ready = serial; }   // release: fence, then the flag
        } else {
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                while (*ready != serial && clock64() - t0 < 4000000000ll) __nanosleep(64);
            }
            __syncthreads();
            if (*ready != serial) return;               // the cull never published: leave the frame unmarched rather than hang
        }
        __threadfence();
    }
    const uint32_t total = ld_list<kFusedCull>(&s.lists->marchTileTotal);
    const uint32_t cubeCount = ld_list<kFusedCull>(&s.lists->cubeCount);
    uint32_t nRays = 0, nSamples = 0, nLight = 0, nSkipped = 0;
    uint32_t stagedVolume = 0xffffffffu;

    // phase (sharded frame, mv_api.cu): 1 = every volume but the frame's light volume, whose light map is still being
    // marched on the light stream; 2 = the light volume alone, once its light map is committed; 0 = all
    uint32_t* cursor = phase == 2 ? &s.lists->marchTileCursor2 : &s.lists->marchTileCursor;
    const uint32_t lightVolume = phase ? s.lists->lightVolume : 0xffffffffu;

    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(cursor, 1u);
        w = __shfl_sync(kFull, w, 0);
        if (w >= total) break;

        // work item -> (cube-map volume, face, tile): upper bound in the tile prefix
        uint32_t lo = 0, hi = cubeCount;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ld_list<kFusedCull>(s.cubeTilePrefix + mid) <= w) lo = mid; else hi = mid;
        }
        const uint32_t volumeId = ld_list<kFusedCull>(s.cubeVolumes + ld_list<kFusedCull>(s.marchOrder + lo));
        if (phase && (volumeId == lightVolume) != (phase == 2)) continue;
        const uint32_t local = w - ld_list<kFusedCull>(s.cubeTilePrefix + lo) + ld_list<kFusedCull>(s.cubeTileBegin + lo);
        const ushort4 a = kFusedCull ? __ldcg(s.attribs + volumeId) : s.attribs[volumeId];
        const uint32_t mip = a.x, smpCount = a.y, maskBits = a.z, volTexId = a.w;
        const uint32_t size = cb.gridSize >> mip;
        const uint32_t tilesX = (size + 7) >> 3, tilesY = (size + 3) >> 2;
        const uint32_t tilesPerFace = tilesX * tilesY;
        const uint32_t faceOrd = local / tilesPerFace, tile = local - faceOrd * tilesPerFace;
        const uint32_t face = nth_set_bit(maskBits & 0x3fu, faceOrd);
        const uint32_t ty = tile / tilesX, tx = tile - ty * tilesX;
        const uint32_t x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);

        // stage the volume's constants (once per run of tiles of the same volume)
        if (stagedVolume != volumeId) {
            __syncwarp();
            const float* src = reinterpret_cast<const float*>(s.perObject + volumeId);
            tc.po[lane] = __ldg(src + lane);
            if (lane < 24) tc.po[32 + lane] = __ldg(src + 32 + lane);
            __syncwarp();
            if (lane == 0) {
                const V3 eye = {cb.eye[0], cb.eye[1], cb.eye[2]};
                const V3 e = mul_p43(eye, tc.po + 32);               // CSRayMarch.hlsl:88
                tc.eyeL[0] = e.x; tc.eyeL[1] = e.y; tc.eyeL[2] = e.z;
            }
            __syncwarp();
            stagedVolume = volumeId;
        }

        if (x < size && y < size) {
            V3 rayOrigin = {tc.eyeL[0], tc.eyeL[1], tc.eyeL[2]};
            const V3 target = get_local_pos((float)x, (float)y, face, (float)size);   // :93
            const V3 rayDir = normalize(target - rayOrigin);                           // :94
            if (compute_ray_origin(rayOrigin, rayDir)) {                               // :95
                const V3 u = (target - rayOrigin) / rayDir;                            // ComputeTargetHit
                float tMax = max3(u.x, u.y, u.z);
                // GetClipPos, :59-71 (scene depth point-sampled at the projection of origin + 0.01 dir)
                const V3 p01 = {rayOrigin.x + 0.01f * rayDir.x, rayOrigin.y + 0.01f * rayDir.y, rayOrigin.z + 0.01f * rayDir.z};
                const V4 hPos = mul_p44(p01, tc.po);
                const float cx = hPos.x / hPos.w, cy = hPos.y / hPos.w;
                const float uvx = cx * 0.5f + 0.5f;
                const float uvy = 1.0f - (cy * 0.5f + 0.5f);
                int ix = (int)floorf(uvx * (float)cb.width), iy = (int)floorf(uvy * (float)cb.height);
                if (!(uvx == uvx)) ix = 0;
                if (!(uvy == uvy)) iy = 0;
                ix = min(max(ix, 0), (int)cb.width - 1); iy = min(max(iy, 0), (int)cb.height - 1);
                const float z = __ldg(s.depth + (size_t)iy * cb.width + ix);
                tMax = fminf(get_tmax(V3{cx, cy, z}, rayOrigin, rayDir, tc.po + 16), tMax);   // :106

                MarchCount mc = {0, 0, 0};
                const uint32_t* emptyBits = s.occ.bits ? s.occ.bits + (size_t)volTexId * s.occ.wordsPerVolume : nullptr;
                const V4 scatter = march_ray(s.volumeTex[volTexId], s.lightTex[volumeId], smpCount, rayOrigin, rayDir, tMax, kDensityOnly, mc, emptyBits, s.occ);

                const size_t idx = ((size_t)face * size + y) * size + x;
                const unsigned long long cOfs = arena_color_offset(s.arena, volumeId, mip) + idx * 8ull;
                const unsigned long long dOfs = arena_depth_offset(s.arena, volumeId, mip) + idx * 4ull;
                const uint2 packed = pack_half4(scatter);
                *reinterpret_cast<uint2*>(s.arena.base + cOfs) = packed;        // g_rwCubeMaps[uavIdx][index], :157
                *reinterpret_cast<float*>(s.arena.base + dOfs) = z;             // g_rwCubeDepths[uavIdx][index], :105
                for (uint32_t p = 0; p < s.arena.numPeers; ++p) {
                    unsigned char* pb = s.arena.peer[p];
                    if (pb) {
                        *reinterpret_cast<uint2*>(pb + cOfs) = packed;
                        *reinterpret_cast<float*>(pb + dOfs) = z;
                    }
                }
                if (kStats) { ++nRays; nSamples += mc.samples; nLight += mc.lightFetches; nSkipped += mc.skipped; }
            }
        }
    }

    if (kStats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            nRays += __shfl_xor_sync(kFull, nRays, d);
            nSamples += __shfl_xor_sync(kFull, nSamples, d);
            nLight += __shfl_xor_sync(kFull, nLight, d);
            nSkipped += __shfl_xor_sync(kFull, nSkipped, d);
        }
        if (lane == 0 && nRays) {
            atomicAdd(&s.stats->view_rays, (unsigned long long)nRays);
            atomicAdd(&s.stats->view_samples, (unsigned long long)nSamples);
            atomicAdd(&s.stats->view_light_fetches, (unsigned long long)nLight);
            if (nSkipped) atomicAdd(&s.stats->view_skipped, (unsigned long long)nSkipped);
        }
    }
}

} // namespace

static void launch_view(Caster& c, bool fusedCull, uint32_t phase, int blocksPerSM)
{
    const bool stats = (c.d.flags & MV_FLAG_COUNT_SAMPLES) != 0, densityOnly = (c.d.flags & MV_FLAG_DENSITY_ONLY) != 0;
    using Kernel = void (*)(DeviceScene, FrameCB, uint32_t, uint32_t);
    static const Kernel kernels[8] = {k_ray_march_v<false, false, false>, k_ray_march_v<true, false, false>, k_ray_march_v<false, true, false>, k_ray_march_v<true, true, false>,
                                      k_ray_march_v<false, false, true>,  k_ray_march_v<true, false, true>,  k_ray_march_v<false, true, true>,  k_ray_march_v<true, true, true>};
    static int perSM[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int v = (stats ? 1 : 0) | (densityOnly ? 2 : 0) | (fusedCull ? 4 : 0);
    if (!perSM[v]) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM[v], kernels[v], kMarchThreads, 0);
        if (perSM[v] < 1) perSM[v] = 1;
    }
    if (fusedCull) ++c.cullSerial;
    static const int capEnv = getenv("MV_VIEW_BLOCKS") ? atoi(getenv("MV_VIEW_BLOCKS")) : 0;     // tuning aid: CTAs per SM of the view march
    if (blocksPerSM <= 0 && capEnv > 0) blocksPerSM = capEnv;
    const int blocks = blocksPerSM > 0 ? min(blocksPerSM, perSM[v]) : perSM[v];
    if (fusedCull) {
        DeviceScene scene = c.scene(); FrameCB cb = c.cb; uint32_t serial = c.cullSerial, ph = phase;
        void* args[] = {&scene, &cb, &serial, &ph};
        cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernels[v]), dim3(c.smCount * blocks), dim3(kMarchThreads), args, 0, c.stream);
    } else kernels[v]<<<c.smCount * blocks, kMarchThreads, 0, c.stream>>>(c.scene(), c.cb, c.cullSerial, phase);
}

void launch_ray_march_view(Caster& c, uint32_t phase, int blocksPerSM) { launch_view(c, false, phase, blocksPerSM); }

// cull + view march in one launch (the work-graph path)
void launch_cull_and_ray_march_view(Caster& c) { launch_view(c, true, 0, 0); }

} // namespace mv
