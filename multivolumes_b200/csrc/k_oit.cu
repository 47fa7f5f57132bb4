// k_oit.cu — order-independent-transparency resolve of the visible volumes, one thread per pixel.
//
// Replaces the reference's K-buffer path — VSCubeDP + PSDepthPeel (PSDepthPeel.hlsl:12-24), VSCube +
// PSCube (PSCube.hlsl:30-60, VSCube.hlsl:51-76), CubeCast (PSCube.hlsli:51-108), RayCast
// (RayCast.hlsli:42-107) and PSResolveOIT (PSResolveOIT.hlsl:12-26), host side
// MultiRayCaster.cpp:1440-1633 — with one fused kernel. The hardware rasteriser of the back faces is
// replaced by the analytic exit point of the pixel-centre ray on every visible volume's box (the
// formulation of the reference's own ray-traced variants, RTCube.hlsl:72-98 / PSCubeRT.hlsl:63-142);
// the 8 nearest exits are kept sorted in registers, so the 96 B/pixel K-buffers (8 x R32_UINT +
// 8 x RGBA16F, MultiRayCaster.cpp:230-236) and their atomics never touch memory.
// Per layer the colour comes from the volume's cube map through the depth-aware 4-tap reconstruction
// (CubeCast) or, for volumes the cull put on the direct scheme, from a screen-space march (RayCast).
#include "k_march.cuh"
#include "k_cube.cuh"

namespace mv {

namespace {

// Stated evaluation order of the OIT passes (the test oracle states the same one): divisions as multiplications by the
// correctly rounded reciprocal (rcp, mv_math.cuh), fused multiply-adds (fma1) exactly where written.
MV_D float unproject_z(float depth)   // UnprojectZ, PSCube.hlsli:21-26
{
    return (kZNear * kZFar) * rcp(fma1(depth, kZNear - kZFar, kZFar));
}
MV_D V3 mul_v33_f(V3 v, const float* M)   // mul(v, (float3x3)M), fused
{
    return {fma1(v.z, M[6], fma1(v.y, M[3], v.x * M[0])), fma1(v.z, M[7], fma1(v.y, M[4], v.x * M[1])), fma1(v.z, M[8], fma1(v.y, M[5], v.x * M[2]))};
}

// CubeCast, PSCube.hlsli:51-108
// `depth` = UnprojectZ of the pixel's scene depth (the same for every layer of the pixel, :83)
MV_D V4 cube_cast(const DeviceScene& s, const FrameCB& cb, uint32_t volumeId, uint32_t mip, float depth, int face, V3 pos, V3 rayDir)
{
    const int S = (int)(cb.gridSize >> mip);
    const float gridSize = (float)S;
    const uint2* colors = reinterpret_cast<const uint2*>(s.arena.base + arena_color_offset(s.arena, volumeId, mip));
    const float* depths = reinterpret_cast<const float*>(s.arena.base + arena_depth_offset(s.arena, volumeId, mip));
    float u, v;
    cube_face_uv(pos, face, u, v);
    const float fx = fma1(u, gridSize, -0.5f), fy = fma1(v, gridSize, -0.5f);
    const float flx = floorf(fx), fly = floorf(fy);
    const int i0 = (int)flx, j0 = (int)fly;
    V4 smp[4]; float zs[4];
    if (i0 >= 0 && j0 >= 0 && i0 + 1 < S && j0 + 1 < S) {   // all four taps on this face (the common case): one base index
        const uint32_t base = ((uint32_t)face * (uint32_t)S + (uint32_t)j0) * (uint32_t)S + (uint32_t)i0;
        const uint32_t idx[4] = {base + (uint32_t)S, base + (uint32_t)S + 1u, base + 1u, base};   // Gather order (-,+), (+,+), (+,-), (-,-)
#pragma unroll
        for (int k = 0; k < 4; ++k) { smp[k] = unpack_half4(__ldg(colors + idx[k])); zs[k] = __ldg(depths + idx[k]); }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ti = (k == 1 || k == 2) ? i0 + 1 : i0, tj = (k < 2) ? j0 + 1 : j0;
            int f, i, j;
            cube_resolve_texel(S, face, ti, tj, f, i, j);
            const uint32_t idx = ((uint32_t)f * (uint32_t)S + (uint32_t)j) * (uint32_t)S + (uint32_t)i;
            smp[k] = unpack_half4(__ldg(colors + idx));
            zs[k] = __ldg(depths + idx);
        }
    }
    // GetDomain, :31-46
    float uvx = u * gridSize, uvy = v * gridSize;
    float domx = frac(fma1(u, gridSize, 0.5f)), domy = frac(fma1(v, gridSize, 0.5f));
    const float bound = gridSize - 1.0f;
    const V3 axes = pos * gridSize;
    const bool edge = (fabsf(axes.x) > bound && axes.x * rayDir.x < 0.0f) || (fabsf(axes.y) > bound && axes.y * rayDir.y < 0.0f) ||
                      (fabsf(axes.z) > bound && axes.z * rayDir.z < 0.0f);
    if (edge) {   // clamp the exterior edge
        uvx = fminf(uvx, gridSize - 0.5f); uvy = fminf(uvy, gridSize - 0.5f);
        domx = uvx < 0.5f ? 1.0f : 0.0f; domy = uvy < 0.5f ? 1.0f : 0.0f;
    }
    const float dix = 1.0f - domx, diy = 1.0f - domy;
    const float wb[4] = {dix * domy, domx * domy, domx * diy, dix * diy};
    V4 result = {0.0f, 0.0f, 0.0f, 0.0f};
    float ws = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float zi = unproject_z(zs[k]);
        float w = fmaxf(fma1(-0.5f, fabsf(depth - zi), 1.0f), 0.0f);
        w *= wb[k];
        result = {fma1(smp[k].x, w, result.x), fma1(smp[k].y, w, result.y), fma1(smp[k].z, w, result.z), fma1(smp[k].w, w, result.w)};
        ws += w;
    }
    if (ws > 0.0f) { const float iw = rcp(ws); return {result.x * iw, result.y * iw, result.z * iw, result.w * iw}; }
    // all taps rejected by depth: plain bilinear SampleLevel of the same footprint (:57, :105)
    const float bx = fx - flx, by = fy - fly;
    const float bw[4] = {(1.0f - bx) * by, bx * by, bx * (1.0f - by), (1.0f - bx) * (1.0f - by)};
    V4 col = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 4; ++k) col = {fma1(smp[k].x, bw[k], col.x), fma1(smp[k].y, bw[k], col.y), fma1(smp[k].z, bw[k], col.z), fma1(smp[k].w, bw[k], col.w)};
    return col;
}

// Pixel-centre ray: unproject z = 0 through screenToWorld (RTCube.hlsl:54-70; PSCube.hlsl:38-40). Returns the world-space
// direction from the eye (not normalised) and the pixel's clip-space x, y.
MV_D V3 pixel_ray(const FrameCB& cb, int px, int py, float& sx, float& sy)
{
    sx = fma1((float)px + 0.5f, cb.inv2Viewport[0], -1.0f); sy = fma1((float)py + 0.5f, -cb.inv2Viewport[1], 1.0f);
    const float* S = cb.screenToWorld;
    const float whx = fma1(sx, S[0], fma1(sy, S[4], S[12])), why = fma1(sx, S[1], fma1(sy, S[5], S[13]));
    const float whz = fma1(sx, S[2], fma1(sy, S[6], S[14])), whw = fma1(sx, S[3], fma1(sy, S[7], S[15]));
    const float iw = rcp(whw);
    return V3{whx * iw, why * iw, whz * iw} - V3{cb.eye[0], cb.eye[1], cb.eye[2]};
}

// The back-face fragment of a volume under a pixel: exit of the local-space ray o + t d from the unit box, with the
// rasteriser's depth clip. false = no fragment. Shared by the resolve kernel and the screen-space march so that both
// see exactly the same fragments.
struct BackFace { int axis; float sgn; float z; };
MV_D bool back_face_fragment(V3 o, V3 d, const float* wvp, BackFace& f)
{
    if (ray_misses_box_for_sure(o, d)) return false;    // the tile overlaps the volume's rectangle, this pixel's ray does not come near
    float tmin = -kFltMax, tmax = kFltMax; int exitAxis = -1; bool miss = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float da = comp(d, a), oa = comp(o, a);
        if (da == 0.0f) { if (fabsf(oa) > 1.0f) miss = true; continue; }
        const float inv = rcp(da);
        const float t1 = (-1.0f - oa) * inv, t2 = (1.0f - oa) * inv;
        const float tn = fminf(t1, t2), tf = fmaxf(t1, t2);
        if (tn > tmin) tmin = tn;
        if (tf < tmax) { tmax = tf; exitAxis = a; }
    }
    if (miss || exitAxis < 0 || !(tmax > 0.0f) || !(tmin < tmax)) return false;
    V3 lpt = {clamp1(fma1(d.x, tmax, o.x)), clamp1(fma1(d.y, tmax, o.y)), clamp1(fma1(d.z, tmax, o.z))};
    const float sgn = comp(d, exitAxis) > 0.0f ? 1.0f : -1.0f;
    if (exitAxis == 0) lpt.x = sgn; else if (exitAxis == 1) lpt.y = sgn; else lpt.z = sgn;
    const float clipZ = fma1(lpt.z, wvp[10], fma1(lpt.y, wvp[6], fma1(lpt.x, wvp[2], wvp[14])));
    const float clipW = fma1(lpt.z, wvp[11], fma1(lpt.y, wvp[7], fma1(lpt.x, wvp[3], wvp[15])));
    if (!(clipW > 0.0f)) return false;
    const float z = clipZ * rcp(clipW);
    if (!(z >= 0.0f && z <= 1.0f)) return false;        // rasteriser depth clip
    f.axis = exitAxis; f.sgn = sgn; f.z = z;
    return true;
}

// The fragment's local-space position: exit point of the pixel ray on the back face (axis, sgn)
MV_D V3 back_face_point(V3 localEye, V3 d, int axis, float sgn)
{
    const float tmax = (sgn - comp(localEye, axis)) * rcp(comp(d, axis));
    V3 lpt = {clamp1(fma1(d.x, tmax, localEye.x)), clamp1(fma1(d.y, tmax, localEye.y)), clamp1(fma1(d.z, tmax, localEye.z))};
    if (axis == 0) lpt.x = sgn; else if (axis == 1) lpt.y = sgn; else lpt.z = sgn;
    return lpt;
}

// What the screen-space march of one volume needs, gathered by the caller so that the out-of-line fallback below does not
// take the whole scene structure by reference (it would be copied to the stack)
struct RayCastArgs {
    cudaTextureObject_t grid, light;
    const uint32_t* emptyBits;
    Occupancy occ;
    const float* wvpi;
    uint32_t smpCnt;
};
MV_D RayCastArgs ray_cast_args(const DeviceScene& s, const PerObject* po, uint32_t volumeId, uint32_t volTexId, uint32_t smpCnt)
{
    RayCastArgs a;
    a.grid = s.volumeTex[volTexId]; a.light = s.lightTex[volumeId];
    a.emptyBits = s.occ.bits ? s.occ.bits + (size_t)volTexId * s.occ.wordsPerVolume : nullptr;
    a.occ = s.occ; a.wvpi = po->wvpi; a.smpCnt = smpCnt;
    return a;
}

// RayCast, RayCast.hlsli:42-107: screen-space march of one fragment of a direct-scheme volume
MV_D V4 ray_cast(const RayCastArgs& a, V3 localEye, V3 rayDir, float sx, float sy, float sceneDepth, bool densityOnly, bool& marched, MarchCount& mc)
{
    V3 ro = localEye; const V3 rd = normalize(rayDir);
    marched = compute_ray_origin(ro, rd);
    if (!marched) return {0.0f, 0.0f, 0.0f, 0.0f};
    const float tMax = get_tmax(V3{sx, sy, sceneDepth}, ro, rd, a.wvpi);
    return march_ray(a.grid, a.light, a.smpCnt, ro, rd, tMax, densityOnly, mc, a.emptyBits, a.occ);
}

// the same, kept out of line: the resolve kernel only marches volumes whose rectangle did not fit the result buffer
__device__ __noinline__ V4 ray_cast_fallback(RayCastArgs a, V3 localEye, V3 rayDir, float sx, float sy, float sceneDepth, bool densityOnly,
                                             uint32_t& rays, uint32_t& samples, uint32_t& light)
{
    bool marched; MarchCount mc = {0, 0, 0};
    const V4 c = ray_cast(a, localEye, rayDir, sx, sy, sceneDepth, densityOnly, marched, mc);
    if (marched) { ++rays; samples += mc.samples; light += mc.lightFetches; }
    return c;
}

// Screen-space march of every direct-scheme volume over its screen rectangle, persistent: warps pull 8x4-pixel tiles of
// one volume (one texture per warp, coherent neighbouring rays) from a cursor. Replaces the RayCast branch of PSCube
// (PSCube.hlsl:44-52) as a pass of its own: inside the per-pixel resolve the marches of different volumes shared warps
// (a TEX per distinct texture, serialised) at 24 warps per SM. The result is stored as the K-buffer would hold it
// (RGBA16F, zero unless 0 < alpha <= 1, PSCube.hlsl:57).
#ifndef MV_DIRECT_MIN_BLOCKS
#define MV_DIRECT_MIN_BLOCKS 6   // same loop as the view march: 5 / 6 / 7 CTAs = 0.930 / 0.912 / 0.942 ms for this pass + the resolve on cfg 4
#endif
template <bool kStats, bool kDensityOnly>
__global__ void __launch_bounds__(256, MV_DIRECT_MIN_BLOCKS) k_ray_cast_direct(DeviceScene s, FrameCB cb)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t total = s.lists->directTileTotal, nvis = s.lists->visibleCount;
    uint32_t dSkipped = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&s.lists->directTileCursor, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        uint32_t lo = 0, hi = nvis;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(s.directTilePrefix + mid) <= w) lo = mid; else hi = mid;
        }
        const VisInfo vi = s.visInfo[lo];
        const uint32_t local = w - __ldg(s.directTilePrefix + lo);
        const int rectW = vi.x1 - vi.x0 + 1;
        const uint32_t tilesX = (uint32_t)(rectW + 7) >> 3;
        const uint32_t ty = local / tilesX, tx = local - ty * tilesX;
        const int px = vi.x0 + (int)(tx * 8 + (lane & 7)), py = vi.y0 + (int)(ty * 4 + (lane >> 3));
        const uint32_t volumeId = vi.volumeId;
        const ushort4 a = s.attribs[volumeId];
        // which pixels of the rectangle this rank marches: the rows it resolves — or, under volume-sharded storage, all of
        // them if it holds the volume (the results then go to every rank, which resolves its own rows from them)
        if (px > vi.x1 || py > vi.y1) continue;
        if (s.shardVolumes ? ((uint32_t)a.w % s.shardWorld != s.shardRank) : !row_is_resolved_here(s, cb, py)) continue;
        const PerObject* po = s.perObject + volumeId;
        float sx, sy;
        const V3 dirW = pixel_ray(cb, px, py, sx, sy);
        const V3 localEye = {vi.eyeL[0], vi.eyeL[1], vi.eyeL[2]};
        const V3 d = mul_v33_f(dirW, po->worldI);
        uint2 stored = make_uint2(0u, 0u), st = make_uint2(0u, 0u);
        BackFace f;
        if (back_face_fragment(localEye, d, po->wvp, f)) {
            const V3 lpt = back_face_point(localEye, d, f.axis, f.sgn);
            const V3 rayDir = lpt - localEye;                                    // PSCube.hlsl:34
            const float sceneDepth = __ldg(s.depth + (size_t)py * cb.width + px);
            bool marched; MarchCount mc = {0, 0, 0};
            const V4 color = ray_cast(ray_cast_args(s, po, volumeId, a.w, a.y), localEye, rayDir, sx, sy, sceneDepth, kDensityOnly, marched, mc);
            if (color.w > 0.0f && color.w <= 1.0f) stored = pack_half4(color);
            if (kStats && marched) { st = make_uint2(mc.samples | 0x80000000u, mc.lightFetches); dSkipped += mc.skipped; }
        }
        const size_t slot = (size_t)__ldg(s.directOffset + lo) + (size_t)(py - vi.y0) * rectW + (px - vi.x0);
        s.directColor[slot] = stored;
        if (s.shardVolumes) for (uint32_t p = 0; p < s.shardWorld; ++p) if (s.directPeer[p]) s.directPeer[p][slot] = stored;
        if (kStats && s.directStats) s.directStats[slot] = st;
    }
    if (kStats) {   // diagnostic: samples of every marched pixel (also of fragments that end up beyond the eighth layer)
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) dSkipped += __shfl_xor_sync(0xffffffffu, dSkipped, dd);
        if (lane == 0 && dSkipped) atomicAdd(&s.stats->direct_skipped, (unsigned long long)dSkipped);
    }
}

constexpr int kOitChunk = 256;   // visible volumes binned per pass over the CTA's 16x16-pixel tile

#ifndef MV_OIT_MIN_BLOCKS
#define MV_OIT_MIN_BLOCKS 4   // 64 registers, 32 warps / SM: the resolve is issue-bound, measured -8 % against 3 CTAs (75 registers)
#endif
__global__ void __launch_bounds__(256, MV_OIT_MIN_BLOCKS) k_resolve_oit(DeviceScene s, FrameCB cb, bool densityOnly)
{
    __shared__ VisInfo s_cand[kOitChunk];       // volumes whose screen rectangle overlaps this tile, list order kept
    __shared__ uint32_t s_candSlot[kOitChunk];  // their index in the visible list
    __shared__ uint32_t s_warpCount[8];
    const int W = (int)cb.width;
    // 16x16-pixel CTA made of 8 warps of 8x4 pixels
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rows [row0, row1) plus a one-row halo on each side (clipped): the TAA's 3x3 neighbourhood of the
    // band's border rows reads them
    int rowBegin, rowEnd, tileRow = (int)blockIdx.y;
    if (s.stripeH) {   // this rank's k-th stripe, + halo
        const int tilesPerStripe = ((int)s.stripeH + 2 + 15) / 16;
        const int k = tileRow / tilesPerStripe;
        tileRow -= k * tilesPerStripe;
        const int r0 = (k * (int)s.shardWorld + (int)s.shardRank) * (int)s.stripeH;
        rowBegin = max(r0 - 1, 0); rowEnd = min(r0 + (int)s.stripeH + 1, (int)cb.height);
    } else { rowBegin = max((int)s.row0 - 1, 0); rowEnd = min((int)s.row1 + 1, (int)cb.height); }
    const int tileX0 = (int)blockIdx.x * 16, tileY0 = rowBegin + tileRow * 16;
    const int px = tileX0 + (int)((warp & 1) * 8 + (lane & 7));
    const int py = tileY0 + (int)((warp >> 1) * 4 + (lane >> 3));
    const bool valid = px < W && py < rowEnd;

    const uint32_t nvis = s.lists->visibleCount;
    float sx, sy;
    const V3 dirW = pixel_ray(cb, px, py, sx, sy);

    // depth peel: the kNumOitLayers nearest back-face exits (PSDepthPeel.hlsl:12-24), sorted, stable
    uint32_t keys[kNumOitLayers], ids[kNumOitLayers];
#pragma unroll
    for (int l = 0; l < (int)kNumOitLayers; ++l) { keys[l] = 0xffffffffu; ids[l] = 0xffffffffu; }
    uint32_t frags = 0;
    for (uint32_t base = 0; base < nvis; base += kOitChunk) {
        // bin: which of the next kOitChunk visible volumes can touch this tile (ordered compaction)
        const uint32_t k = base + threadIdx.x;
        bool overlap = false;
        VisInfo vi;
        if (k < nvis) {
            vi = s.visInfo[k];
            overlap = vi.x0 <= tileX0 + 15 && vi.x1 >= tileX0 && vi.y0 <= tileY0 + 15 && vi.y1 >= tileY0;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, overlap);
        if (lane == 0) s_warpCount[warp] = __popc(bits);
        __syncthreads();
        uint32_t off = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < 8; ++w) { const uint32_t c = s_warpCount[w]; if (w < warp) off += c; total += c; }
        if (overlap) { const uint32_t slot = off + __popc(bits & ((1u << lane) - 1u)); s_cand[slot] = vi; s_candSlot[slot] = k; }
        __syncthreads();
        if (valid) {
            const int wx0 = tileX0 + (int)((warp & 1) * 8), wy0 = tileY0 + (int)((warp >> 1) * 4);   // this warp's 8x4 pixels
            for (uint32_t ci = 0; ci < total; ++ci) {
                // the volume's rectangle overlaps the CTA's tile; does it reach this warp's pixels? (uniform over the warp)
                if (s_cand[ci].x0 > wx0 + 7 || s_cand[ci].x1 < wx0 || s_cand[ci].y0 > wy0 + 3 || s_cand[ci].y1 < wy0) continue;
                const uint32_t volumeId = s_cand[ci].volumeId;
                const PerObject* po = s.perObject + volumeId;
                const V3 o = {s_cand[ci].eyeL[0], s_cand[ci].eyeL[1], s_cand[ci].eyeL[2]};   // mul(float4(g_eyePt, 1), WorldI)
                const V3 d = mul_v33_f(dirW, po->worldI);
                BackFace f;
                if (!back_face_fragment(o, d, po->wvp, f)) continue;
                ++frags;
                // insert (key, visible-list index) keeping ascending keys; equal keys keep list order
                uint32_t key = as_uint(f.z), id = s_candSlot[ci] | ((uint32_t)(f.axis * 2 + (f.sgn > 0.0f ? 0 : 1)) << 24);
#pragma unroll
                for (int l = 0; l < (int)kNumOitLayers; ++l) {
                    if (key < keys[l]) {
                        const uint32_t tk = keys[l], ti = ids[l];
                        keys[l] = key; ids[l] = id; key = tk; id = ti;
                    }
                }
            }
        }
        __syncthreads();
    }

    // shade + resolve front to back (PSCube.hlsl:30-60, PSResolveOIT.hlsl:12-26)
    const float sceneDepth = valid ? __ldg(s.depth + (size_t)py * W + px) : 1.0f;
    const float sceneZ = unproject_z(sceneDepth);
    V4 result = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t dRays = 0, dSamples = 0, dLight = 0;
#pragma unroll 1
    for (int l = 0; l < (int)kNumOitLayers; ++l) {
        // (register arrays are indexed by a runtime l here; the loop is kept rolled because the body is large)
        uint32_t id = 0xffffffffu;
#pragma unroll
        for (int m = 0; m < (int)kNumOitLayers; ++m) if (m == l) id = ids[m];
        if (id == 0xffffffffu) break;
        const uint32_t volumeId = __ldg(s.visible + (id & 0xffffffu));
        const int face = (int)(id >> 24);
        const PerObject* po = s.perObject + volumeId;
        const ushort4 a = s.attribs[volumeId];
        const V3 localEye = {__ldg(&s.visInfo[id & 0xffffffu].eyeL[0]), __ldg(&s.visInfo[id & 0xffffffu].eyeL[1]), __ldg(&s.visInfo[id & 0xffffffu].eyeL[2])};
        // the fragment's local-space position: exit point of the pixel ray on the back face
        const V3 d = mul_v33_f(dirW, po->worldI);
        const int axis = face >> 1;
        const float sgn = (face & 1) ? -1.0f : 1.0f;
        const V3 lpt = back_face_point(localEye, d, axis, sgn);
        const V3 rayDir = lpt - localEye;                                        // PSCube.hlsl:34
        const uint32_t smpCnt = (a.z & kCubeMapRayMarchBit) ? 0u : (uint32_t)a.y;   // VSCube.hlsl:73
        // K-colour layers are RGBA16F; a layer is written only if 0 < alpha <= 1 (PSCube.hlsl:57)
        V4 src = {0.0f, 0.0f, 0.0f, 0.0f};
        V4 color = {0.0f, 0.0f, 0.0f, 0.0f};
        bool stored = false;
        if (smpCnt > 0) {
            const uint32_t slot = id & 0xffffffu;
            const uint32_t off = __ldg(s.directOffset + slot);
            const int x0 = __ldg(&s.visInfo[slot].x0), y0 = __ldg(&s.visInfo[slot].y0), x1 = __ldg(&s.visInfo[slot].x1), y1 = __ldg(&s.visInfo[slot].y1);
            if (off != kNoDirect && px >= x0 && px <= x1 && py >= y0 && py <= y1) {   // marched by k_ray_cast_direct, already in K-buffer form
                const size_t at = (size_t)off + (size_t)(py - y0) * (x1 - x0 + 1) + (px - x0);
                src = unpack_half4(__ldg(s.directColor + at));
                stored = true;
                if (s.directStats) {
                    const uint2 st = __ldg(s.directStats + at);
                    if (st.x >> 31) { ++dRays; dSamples += st.x & 0x7fffffffu; dLight += st.y; }
                }
            } else if (!s.shardVolumes)      // (volume-sharded storage: the resolve cannot march a volume this rank does not hold; include/mv.h)
                color = ray_cast_fallback(ray_cast_args(s, po, volumeId, a.w, smpCnt), localEye, rayDir, sx, sy, sceneDepth, densityOnly, dRays, dSamples, dLight);
        } else color = cube_cast(s, cb, volumeId, a.x, sceneZ, face, lpt, rayDir);
        if (!stored && color.w > 0.0f && color.w <= 1.0f) src = unpack_half4(pack_half4(color));
        const float k1 = 1.0f - result.w;
        result = {src.x * k1 + result.x, src.y * k1 + result.y, src.z * k1 + result.z, src.w * k1 + result.w};   // fmul, fadd: as PSResolveOIT.cso has it
    }
    result.w = fminf(result.w, kAlphaClamp);                                         // PSResolveOIT.hlsl:22
    if (valid) {
        // premultiplied-alpha blend onto the colour RT (Graphics::PREMULTIPLITED, MultiRayCaster.cpp:931)
        uint2* dst = s.color + (size_t)py * W + px;
        const V4 d4 = unpack_half4(*dst);
        const float ia = 1.0f - result.w;
        *dst = pack_half4(V4{fma1(d4.x, ia, result.x), fma1(d4.y, ia, result.y), fma1(d4.z, ia, result.z), fma1(d4.w, ia, result.w)});
    }

    if (s.stats) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            frags += __shfl_xor_sync(0xffffffffu, frags, dd);
            dRays += __shfl_xor_sync(0xffffffffu, dRays, dd);
            dSamples += __shfl_xor_sync(0xffffffffu, dSamples, dd);
            dLight += __shfl_xor_sync(0xffffffffu, dLight, dd);
        }
        if (lane == 0 && frags) {
            atomicAdd(&s.stats->oit_fragments, (unsigned long long)frags);
            if (dRays) {
                atomicAdd(&s.stats->direct_rays, (unsigned long long)dRays);
                atomicAdd(&s.stats->direct_samples, (unsigned long long)dSamples);
                atomicAdd(&s.stats->direct_light_fetches, (unsigned long long)dLight);
            }
        }
    }
}

} // namespace

void launch_ray_cast_direct(Caster& c)
{
    const bool stats = (c.d.flags & MV_FLAG_COUNT_SAMPLES) != 0, densityOnly = (c.d.flags & MV_FLAG_DENSITY_ONLY) != 0;
    using Kernel = void (*)(DeviceScene, FrameCB);
    static const Kernel kernels[4] = {k_ray_cast_direct<false, false>, k_ray_cast_direct<true, false>, k_ray_cast_direct<false, true>, k_ray_cast_direct<true, true>};
    static int perSM[4] = {0, 0, 0, 0};
    const int v = (stats ? 1 : 0) | (densityOnly ? 2 : 0);
    if (!perSM[v]) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM[v], kernels[v], 256, 0);
        perSM[v] = max(perSM[v], 1);
    }
    kernels[v]<<<c.smCount * perSM[v], 256, 0, c.stream>>>(c.scene(), c.cb);
}

void launch_resolve_oit(Caster& c)
{
    uint32_t tileRows;
    if (c.shardWorld > 1 && c.stripeH) tileRows = num_own_stripes(c.d.height, c.stripeH, c.shardRank, c.shardWorld) * ((c.stripeH + 2 + 15) / 16);
    else {
        if (c.row1 <= c.row0) return;
        const uint32_t rowBegin = c.row0 > 0 ? c.row0 - 1 : 0, rowEnd = min(c.row1 + 1, c.d.height);
        tileRows = (rowEnd - rowBegin + 15) / 16;
    }
    if (tileRows == 0) return;
    dim3 grid((c.d.width + 15) / 16, tileRows);
    k_resolve_oit<<<grid, 256, 0, c.stream>>>(c.scene(), c.cb, (c.d.flags & MV_FLAG_DENSITY_ONLY) != 0);
}

} // namespace mv
