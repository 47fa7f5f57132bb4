// k_cull.cuh — the body of the volume cull (CSVolumeCull.hlsl:13-78, VolumeCull.hlsli:27-334) as a device function of one
// CTA, shared by the stand-alone cull kernel (k_cull.cu, 1024 threads) and by the fused cull -> view-march launch
// (k_ray_march_v.cu, CTA 0 of the persistent march kernel: the work-graph path of the reference,
// LibRayMarch.hlsl:39-134, where the cull node feeds the march node inside one DispatchGraph).
#pragma once
#include "mv_internal.h"

namespace mv {

namespace {

constexpr uint32_t kCullFull = 0xffffffffu;

// VolumeCull.hlsli:119-138 — unique edge id -> (corner, corner)
static __constant__ unsigned char c_edgeLanes[12][2] = {{0, 1}, {3, 2}, {1, 3}, {2, 0}, {6, 7}, {5, 4},
                                                 {4, 6}, {7, 5}, {4, 0}, {2, 6}, {7, 3}, {1, 5}};
// VolumeCull.hlsli:213-223 — face (by mask bit) -> 4 unique edge ids
static __constant__ unsigned char c_faceEdges[6][4] = {{8, 3, 9, 6}, {10, 2, 11, 7}, {0, 8, 5, 11},
                                                {1, 10, 4, 9}, {0, 2, 1, 3}, {4, 6, 5, 7}};

MV_D V3 project_to_viewport(uint32_t i, const float* wvp, float vw, float vh, float& w)   // VolumeCull.hlsli:27-41
{
    const V3 p3 = {(i & 1) ? 1.0f : -1.0f, ((i >> 1) & 1) ? 1.0f : -1.0f, (i >> 2) ? 1.0f : -1.0f};
    V4 p = mul_p44(p3, wvp);
    w = p.w;
    p.x /= p.w; p.y /= p.w; p.z /= p.w;
    p.x = p.x * 0.5f + 0.5f; p.y = p.y * 0.5f + 0.5f;
    p.y = 1.0f - p.y;
    return {p.x * vw, p.y * vh, p.z};
}

// One CTA of kThreads threads (a multiple of 32, at least 64) culls all N volumes. pickLightVolume: also choose the
// frame's light-march volume from the list just built (CSRayMarchL.hlsl:29-33); the work-graph order marches the light
// map before the cull, from the previous frame's list, and leaves it alone.
template <int kThreads>
MV_D void cull_body(const DeviceScene& s, const FrameCB& cb, bool pickLightVolume)
{
    constexpr int kWarps = kThreads / 32;
    constexpr uint32_t kFull = kCullFull;

    __shared__ uint32_t s_warpVis[kWarps], s_warpCube[kWarps];
    __shared__ uint32_t s_baseVis, s_baseCube;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t grp = lane >> 3, corner = lane & 7, baseLane = grp * 8;
    const uint32_t N = cb.numVolumes;
    if (threadIdx.x == 0) { s_baseVis = 0; s_baseCube = 0; }
    __syncthreads();

    for (uint32_t chunk = 0; chunk < N; chunk += kWarps * kGroupVolumeCount) {
        const uint32_t volumeId = chunk + warp * kGroupVolumeCount + grp;
        const bool valid = volumeId < N;
        const PerObject* po = s.perObject + (valid ? volumeId : 0);

        // CSVolumeCull.hlsl:29-38 — one corner per lane
        V3 v = {0.0f, 0.0f, 0.0f};
        float clipW = 1.0f;
        bool isInView = false;
        if (valid) {
            v = project_to_viewport(corner, po->wvp, cb.viewport[0], cb.viewport[1], clipW);
            isInView = (v.x <= cb.viewport[0] && v.y <= cb.viewport[1] && v.x >= 0.0f && v.y >= 0.0f) && v.z > 0.0f && v.z < 1.0f;
        }
        const uint32_t volumeVis = (__ballot_sync(kFull, isInView) >> baseLane) & 0xffu;
        const bool visible = valid && volumeVis != 0;

        // GenVisibilityMask, VolumeCull.hlsli:46-66 — one face per lane
        bool faceVis = false;
        if (visible && corner < 6) {
            const V3 eye = {cb.eye[0], cb.eye[1], cb.eye[2]};
            const V3 localEye = mul_p43(eye, po->worldI);
            const float viewComp = comp(localEye, (int)(corner >> 1));
            faceVis = (corner & 1) ? viewComp > -1.0f : viewComp < 1.0f;
        }
        const uint32_t faceMask = (__ballot_sync(kFull, faceVis) >> baseLane) & 0xffu;

        // GetCubeEdgePairPerLane, :156-181 — lanes 0..5 hold unique edges 2c and 2c + 1
        const uint32_t ec = corner < 6 ? corner : 0;
        float ex[2], ey[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t a = baseLane + c_edgeLanes[2 * ec + k][0], b = baseLane + c_edgeLanes[2 * ec + k][1];
            const float ax = __shfl_sync(kFull, v.x, a), ay = __shfl_sync(kFull, v.y, a);
            const float bx = __shfl_sync(kFull, v.x, b), by = __shfl_sync(kFull, v.y, b);
            ex[k] = bx - ax; ey[k] = by - ay;
        }
        // EstimateCubeMaxEdgeLength, :248-262
        const float ms = corner < 6 ? fmaxf(length(V2{ex[0], ey[0]}), length(V2{ex[1], ey[1]})) : 0.0f;
        float maxEdge = __shfl_sync(kFull, ms, baseLane);
#pragma unroll
        for (int k = 1; k < 6; ++k) maxEdge = fmaxf(maxEdge, __shfl_sync(kFull, ms, baseLane + k));

        // EstimateProjCoverage, :299-322 — one face per lane, area of the quad spanned by its 4 edges
        float fe[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t id = c_faceEdges[ec][k];
            const uint32_t src = baseLane + (id >> 1);
            const float x0 = __shfl_sync(kFull, ex[0], src), y0 = __shfl_sync(kFull, ey[0], src);
            const float x1 = __shfl_sync(kFull, ex[1], src), y1 = __shfl_sync(kFull, ey[1], src);
            fe[k][0] = (id & 1) ? x1 : x0; fe[k][1] = (id & 1) ? y1 : y0;
        }
        float faceArea = 0.0f;
        if (corner < 6 && (faceMask & (1u << corner))) {
            const float t0 = 0.5f * fabsf(fe[0][0] * fe[1][1] - fe[0][1] * fe[1][0]);   // CalcTriangleArea :71-74
            const float t1 = 0.5f * fabsf(fe[2][0] * fe[3][1] - fe[2][1] * fe[3][0]);
            faceArea = t0 + t1;
        }
        // WaveActiveSum pinned to a lane-ascending sequential sum
        float projCov = __shfl_sync(kFull, faceArea, baseLane);
#pragma unroll
        for (int k = 1; k < 6; ++k) projCov = projCov + __shfl_sync(kFull, faceArea, baseLane + k);

        bool useCubeMap = false;
        if (visible && corner == 0) {
            const uint32_t volumeIn = s.volumeDescs[volumeId];
            const uint32_t cubeMapSize = volumeIn >> 18, numMips = (volumeIn >> 14) & 0xfu;
            // EstimateCubeMapLOD, :267-294 (upscale 2, raySampleCountScale 2)
            const float sqrt3 = sqrtf(3.0f);
            float sz = maxEdge / 2.0f;
            float raySampleAmt = 2.0f * sz / sqrt3;
            const uint32_t raySampleCnt = float_to_uint_sat(ceilf(raySampleAmt));
            const uint32_t raySampleCount = min(raySampleCnt, cb.maxRaySamples);
            raySampleAmt = fminf(raySampleAmt, (float)raySampleCount);
            sz = raySampleAmt / 2.0f * sqrt3;
            const uint32_t level = floor_log2_clamped((float)cubeMapSize / sz);
            const uint32_t mipLevel = min(level, numMips - 1);
            // EstimateCubeMapVisiblePixels, :327-334; CSVolumeCull.hlsl:66-67
            const uint32_t edgeLength = cubeMapSize >> mipLevel;
            const float cubeMapPix = (float)(edgeLength * edgeLength) * (float)__popc(faceMask);
            useCubeMap = cubeMapPix <= projCov;
            const uint32_t maskBits = useCubeMap ? (faceMask | kCubeMapRayMarchBit) : faceMask;
            s.attribs[volumeId] = make_ushort4((unsigned short)mipLevel, (unsigned short)raySampleCount,
                                               (unsigned short)maskBits, (unsigned short)(volumeIn & 0x3fffu));
        }

        // conservative screen rectangle of the projected box for the OIT resolve (not a reference output):
        // min / max of the eight corners, two pixels of slack; any corner on or behind the eye plane
        // (or a non-finite projection) makes it the whole screen
        const bool badCorner = !(clipW > 0.0f) || !(fabsf(v.x) <= 3.0e8f) || !(fabsf(v.y) <= 3.0e8f);
        const uint32_t badBits = (__ballot_sync(kFull, badCorner) >> baseLane) & 0xffu;
        float bx0 = v.x, bx1 = v.x, by0 = v.y, by1 = v.y;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(kFull, bx0, d)); bx1 = fmaxf(bx1, __shfl_xor_sync(kFull, bx1, d));
            by0 = fminf(by0, __shfl_xor_sync(kFull, by0, d)); by1 = fmaxf(by1, __shfl_xor_sync(kFull, by1, d));
        }

        // ordered compaction: ballot inside the warp, prefix over the 32 warps through shared memory
        const uint32_t visBits = __ballot_sync(kFull, visible && corner == 0);
        const uint32_t cubeBits = __ballot_sync(kFull, useCubeMap);
        if (lane == 0) { s_warpVis[warp] = __popc(visBits); s_warpCube[warp] = __popc(cubeBits); }
        __syncthreads();
        uint32_t offVis = s_baseVis, offCube = s_baseCube;
        for (uint32_t w = 0; w < warp; ++w) { offVis += s_warpVis[w]; offCube += s_warpCube[w]; }
        const uint32_t below = (1u << lane) - 1u;
        if (visible && corner == 0) {
            const uint32_t slot = offVis + __popc(visBits & below);
            s.visible[slot] = volumeId;
            VisInfo vi;
            const V3 eye = {cb.eye[0], cb.eye[1], cb.eye[2]};
            const V3 e = mul_p43(eye, po->worldI);
            vi.eyeL[0] = e.x; vi.eyeL[1] = e.y; vi.eyeL[2] = e.z;
            vi.volumeId = volumeId;
            const int W = (int)cb.width, H = (int)cb.height;
            if (badBits) { vi.x0 = 0; vi.y0 = 0; vi.x1 = W - 1; vi.y1 = H - 1; }
            else {
                vi.x0 = max((int)floorf(bx0) - 2, 0); vi.y0 = max((int)floorf(by0) - 2, 0);
                vi.x1 = min((int)ceilf(bx1) + 2, W - 1); vi.y1 = min((int)ceilf(by1) + 2, H - 1);
            }
            s.visInfo[slot] = vi;
        }
        if (useCubeMap) s.cubeVolumes[offCube + __popc(cubeBits & below)] = volumeId;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tv = 0, tc = 0;
            for (int w = 0; w < kWarps; ++w) { tv += s_warpVis[w]; tc += s_warpCube[w]; }
            s_baseVis += tv; s_baseCube += tc;
        }
        __syncthreads();
    }

    const uint32_t visibleCount = s_baseVis, cubeCount = s_baseCube;
    // Order in which the persistent view march walks the cube-map volumes: longest rays (largest sample count) first, so
    // that the kernel ends on short rays instead of draining the SMs behind a few 256-step chains. Rank sort over the
    // whole CTA (cubeCount^2 / 1024 comparisons per thread); the cube-map volume LIST itself stays in ascending order.
    for (uint32_t k = threadIdx.x; k < cubeCount; k += kThreads) {
        const uint32_t key = s.attribs[s.cubeVolumes[k]].y;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < cubeCount; ++j) {
            const uint32_t kj = s.attribs[s.cubeVolumes[j]].y;
            rank += (kj > key || (kj == key && j < k)) ? 1u : 0u;
        }
        s.marchOrder[rank] = k;
    }
    __syncthreads();
    // Screen-space marches (RayCast, volumes on the direct scheme: VSCube.hlsl:73): every such visible volume gets its
    // screen rectangle as a slice of the result buffer and a run of 8x4-pixel tiles (warp 1, two shuffle scans). A
    // rectangle that does not fit the buffer gets no slice: the resolve kernel marches that volume itself.
    if (warp == 1) {
        uint32_t runningPix = 0, runningTiles = 0;
        for (uint32_t base = 0; base < visibleCount; base += 32) {
            const uint32_t k = base + lane;
            uint32_t pix = 0, tiles = 0;
            if (k < visibleCount) {
                const ushort4 a = s.attribs[s.visible[k]];
                if (!(a.z & kCubeMapRayMarchBit) && a.y > 0) {
                    const VisInfo vi = s.visInfo[k];
                    const int w = vi.x1 - vi.x0 + 1, h = vi.y1 - vi.y0 + 1;
                    if (w > 0 && h > 0) { pix = (uint32_t)w * (uint32_t)h; tiles = (uint32_t)((w + 7) / 8) * (uint32_t)((h + 3) / 4); }
                }
            }
            uint32_t incl = pix;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, d); if (lane >= (uint32_t)d) incl += t; }
            const uint32_t offset = runningPix + incl - pix;
            const bool fits = pix != 0 && offset <= s.directCapacity && pix <= s.directCapacity - offset;
            runningPix = min(runningPix + __shfl_sync(kFull, incl, 31), 0x80000000u);   // saturate: everything after an overflow does not fit either
            if (!fits) tiles = 0;
            uint32_t inclT = tiles;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, inclT, d); if (lane >= (uint32_t)d) inclT += t; }
            if (k < visibleCount) {
                s.directOffset[k] = fits ? offset : kNoDirect;
                s.directTilePrefix[k] = runningTiles + inclT - tiles;
            }
            runningTiles += __shfl_sync(kFull, inclT, 31);
        }
        if (lane == 0) { s.directTilePrefix[visibleCount] = runningTiles; s.lists->directTileTotal = runningTiles; }
    }
    // Tile prefix of the view march over the cube-map volumes (warp 0, shuffle scan), in march order. One GPU: every tile.
    // Sharded, peers unmapped (collective exchange: whole cube maps are broadcast by their owner): the tiles of the
    // volumes v % world == rank. Sharded, peers mapped (every texel is stored into all arenas by whoever marches it, so
    // ownership can be as fine as a tile): every volume's tile list (face-major, row-major) is cut into `world` equal
    // contiguous parts and rank r marches part (r - k) mod world of the k-th volume. Each rank then carries 1 / world
    // of every volume's rays — whatever the content makes them cost (sample counts per ray differ 2x between volumes:
    // profiles/r01_notes.md) — while still reading only the wedge of each volume its own rays cross, and the rotation
    // keeps one rank from always getting the same face. Every rank evaluates the same integers: the parts tile the list.
    if (warp == 0) {
        const bool balanced = s.shardWorld > 1 && s.arena.numPeers != 0 && !s.shardVolumes;   // volume-sharded storage: a volume is marched by the rank that holds it
        uint32_t running = 0;
        for (uint32_t base = 0; base < cubeCount; base += 32) {
            const uint32_t k = base + lane;
            uint32_t tiles = 0, begin = 0;
            if (k < cubeCount) {
                const uint32_t vol = s.cubeVolumes[s.marchOrder[k]];
                const ushort4 a = s.attribs[vol];
                const uint32_t size = cb.gridSize >> a.x;
                const uint32_t all = ((size + 7) / 8) * ((size + 3) / 4) * __popc(a.z & 0x3fu);
                if (balanced) {
                    const uint32_t part = (s.shardRank + s.shardWorld - k % s.shardWorld) % s.shardWorld;
                    begin = (uint32_t)((unsigned long long)all * part / s.shardWorld);
                    tiles = (uint32_t)((unsigned long long)all * (part + 1) / s.shardWorld) - begin;
                } else if ((s.shardVolumes ? (uint32_t)a.w : vol) % s.shardWorld == s.shardRank) tiles = all;
            }
            uint32_t incl = tiles;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            if (k < cubeCount) { s.cubeTilePrefix[k] = running + incl - tiles; s.cubeTileBegin[k] = begin; }
            running += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) {
            s.cubeTilePrefix[cubeCount] = running;
            FrameLists* L = s.lists;
            L->visibleCount = visibleCount;
            L->cubeCount = cubeCount;
            L->marchTileTotal = running;
            L->marchTileCursor = 0;
            L->marchTileCursor2 = 0;
            L->oitTileCursor = 0;
            L->directTileCursor = 0;
            L->lightDenseCount = 0;
            L->lightDenseCursor = 0;
            L->lightItemCount = 0;
            L->lightItemCursor = 0;
            L->lightResultCount = 0;
            L->lightEmitCursor = 0;
            L->lightOverflow = 0;
            // CSRayMarchL.hlsl:29-33 (visible[] was written by other threads of this CTA before the last barrier)
            if (pickLightVolume) L->lightVolume = visibleCount ? s.visible[cb.frameIdx % visibleCount] : cb.frameIdx % N;
        }
    }
}

} // namespace

} // namespace mv
