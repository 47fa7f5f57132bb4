// mv_mesh.cu — producer of the two depth inputs of the volume path: the scene depth (D32, W x H) and the
// light's orthographic shadow map (D16, S x S) of the occluder mesh.
//
// Replaces ObjectRenderer's depth-only passes (MultiVolumes/Content/ObjectRenderer.cpp:171-190 shadow
// view-projection, :220-243 RenderShadow, :555-570 renderDepth; VSDepth.hlsl:25-28) and the OBJ import
// in front of them (XUSG/Optional/XUSGObjLoader.cpp:18-40, :166-228). The D3D12 rasteriser is fixed
// function; here it is a sm_100a kernel pair with the Direct3D rules written out, so that the result is
// independent of the order in which triangles are processed and reproducible bit for bit:
//   * clip = mul(float4(pos, 1), WVP) in fp32 (mv_math.cuh order); triangles are clipped against the
//     near plane z >= 0 only (at most two triangles come out); x, y are clamped by the bounding-box
//     scissor, z > 1 is rejected per pixel (depth clip);
//   * screen x, y are snapped to 1/256 pixel (round half up) and the three edge functions are
//     evaluated in 64-bit integers at pixel centres (x + 0.5, y + 0.5) with the top-left fill rule;
//   * depth = (E12 z0 + E20 z1 + E01 z2) / (2 area) in fp64 from the integer edge values, rounded once
//     to fp32 (z / w is affine in screen space); LESS test against a 1.0 clear = atomicMin on the bit
//     pattern of the non-negative float;
//   * no face culling (XUSG's default rasteriser state is not visible in the reference; for a closed
//     mesh the nearest surface is a front face either way); D16 = floor(z * 65535 + 0.5).
//
// Round 2 adds the BASE PASS over the same rasteriser (ObjectRenderer::Render, ObjectRenderer.cpp:532-553; VSBasePass.hlsl:39-55,
// PSBasePass.hlsl:94-153): a visibility buffer (depth bits | record index, 64-bit atomicMin: LESS, ties to the earlier
// triangle as in draw order), then one thread per pixel interpolates the vertex shader's outputs perspective-correctly —
// fp64 from the integer edge values and 1 / w, rounded once to fp32 — and runs the pixel shader: 2x2 PCF shadow test,
// SH irradiance (or the hemisphere ambient), Lambert + pow(NoH, 64) Schlick specular, velocity from the previous frame's
// world-view-projection. Declared deviations: no sub-pixel jitter (XUSG's Halton sequence is binary-only; ProjBias = 0), no
// radiance term (`SampleBias(R, 2.0)` takes its LOD from the hardware's quad derivatives; RADIANCE_BIT clear), zero velocity
// on the first frame (the reference reads an uninitialised matrix there).
// Kernels: k_mesh_setup (one thread per triangle -> up to two screen-space records) and k_mesh_raster
// (one warp per record, lanes sweep the bounding box in 8x4 blocks). Bunny-sized meshes (70 k
// triangles of a few pixels) are latency-bound: a few tens of microseconds per map.
#include "k_march.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <new>
#include <string>

using namespace mv;

struct mv_caster { Caster c; };

namespace mv {

namespace {

struct ScreenTri {
    int x[3], y[3];      // 24.8 fixed point
    float z[3];
    int valid;
};

struct RasterTarget {
    float wvp[16];
    uint32_t width, height;
    uint32_t* depthBits;   // width * height, float bit patterns
};

// Base pass: what the vertex shader hands to the pixel shader (VSBasePass.hlsl:14-22), per vertex of a screen-space record
struct ShadeVertex {
    float ws[3];       // WSPos
    float nrm[3];      // Norm = mul(Nrm, (float3x3)World)
    float ls[3];       // LSPos.xyz (the light's projection is orthographic: w = 1)
    float cs[3];       // CSPos.x, .y, .w
    float ts[3];       // TSPos.x, .y, .w
    float invW;        // 1 / Pos.w
};
struct ShadeTri { ShadeVertex v[3]; };

struct BasePassCB {       // cbPerObject + cbPerFrame of the base pass (VSBasePass.hlsl:27-34, PSBasePass.hlsl:36-49)
    float wvp[16], wvpPrev[16], world[12], shadowWVP[16];
    float eye[3], lightPos[3], lightColor[4], ambient[4];
    uint32_t hasSH;
    float sh[27];
    float clear[4];
};

MV_D ShadeVertex lerp_sv(const ShadeVertex& a, const ShadeVertex& b, float t)
{
    ShadeVertex r;
    const float* pa = reinterpret_cast<const float*>(&a); const float* pb = reinterpret_cast<const float*>(&b); float* pr = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int k = 0; k < 15; ++k) pr[k] = pa[k] + (pb[k] - pa[k]) * t;      // every output of the vertex shader is affine in the position
    r.invW = 0.0f;
    return r;
}

MV_D V4 lerp4(V4 a, V4 b, float t) { return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t}; }

MV_D void to_screen(const RasterTarget& rt, V4 c, int& x, int& y, float& z)
{
    const float ndcX = c.x / c.w, ndcY = c.y / c.w;
    z = c.z / c.w;
    const float sx = (ndcX * 0.5f + 0.5f) * (float)rt.width;
    const float sy = (1.0f - (ndcY * 0.5f + 0.5f)) * (float)rt.height;
    // snap to 1/256 pixel; clamp far outside the guard band (such a vertex only stretches the triangle,
    // the bounding-box scissor cuts it back) so that the 64-bit edge functions cannot overflow
    const float lim = 4194304.0f;
    x = (int)floorf(fminf(fmaxf(sx, -lim), lim) * 256.0f + 0.5f);
    y = (int)floorf(fminf(fmaxf(sy, -lim), lim) * 256.0f + 0.5f);
}

template <bool kShade>
__global__ void __launch_bounds__(256) k_mesh_setup(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t numTris,
                                                   RasterTarget rt, ScreenTri* __restrict__ out,
                                                   const float* __restrict__ nrm, const BasePassCB cb, ShadeTri* __restrict__ shadeOut)
{
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= numTris) return;
    V4 c[3];
    ShadeVertex sv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t v = idx[3 * t + k];
        const V3 p = {pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]};
        c[k] = mul_p44(p, rt.wvp);
        if (kShade) {      // VSBasePass.hlsl:44-53
            const V3 ws = mul_p43(p, cb.world);
            const V3 n = mul_v33(V3{nrm[3 * v], nrm[3 * v + 1], nrm[3 * v + 2]}, cb.world);
            const V4 ls = mul_p44(p, cb.shadowWVP), ts = mul_p44(p, cb.wvpPrev);
            sv[k] = {{ws.x, ws.y, ws.z}, {n.x, n.y, n.z}, {ls.x, ls.y, ls.z}, {c[k].x, c[k].y, c[k].w}, {ts.x, ts.y, ts.w}, 0.0f};
        }
    }
    // near-plane clip (z >= 0), Sutherland-Hodgman on one plane: 0, 3 or 4 vertices
    V4 poly[4]; ShadeVertex spoly[4]; int n = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const V4 a = c[k], b = c[(k + 1) % 3];
        const bool ain = a.z >= 0.0f, bin = b.z >= 0.0f;
        if (ain) { if (kShade) spoly[n] = sv[k]; poly[n++] = a; }
        if (ain != bin) {
            // intersect from the inside vertex so that both triangles sharing the edge get the same point
            const V4 p = ain ? a : b, q = ain ? b : a;
            const float tt = p.z / (p.z - q.z);
            if (kShade) spoly[n] = lerp_sv(ain ? sv[k] : sv[(k + 1) % 3], ain ? sv[(k + 1) % 3] : sv[k], tt);
            poly[n++] = lerp4(p, q, tt);
        }
    }
    ScreenTri r0, r1;
    r0.valid = 0; r1.valid = 0;
    if (n >= 3) {
        int x[4], y[4]; float z[4];
        bool ok = true;
        for (int k = 0; k < n; ++k) {
            if (!(poly[k].w > 0.0f)) ok = false;
            else to_screen(rt, poly[k], x[k], y[k], z[k]);
        }
        if (ok) {
            r0.x[0] = x[0]; r0.y[0] = y[0]; r0.z[0] = z[0];
            r0.x[1] = x[1]; r0.y[1] = y[1]; r0.z[1] = z[1];
            r0.x[2] = x[2]; r0.y[2] = y[2]; r0.z[2] = z[2];
            r0.valid = 1;
            if (n == 4) {
                r1.x[0] = x[0]; r1.y[0] = y[0]; r1.z[0] = z[0];
                r1.x[1] = x[2]; r1.y[1] = y[2]; r1.z[1] = z[2];
                r1.x[2] = x[3]; r1.y[2] = y[3]; r1.z[2] = z[3];
                r1.valid = 1;
            }
        }
    }
    out[2 * t] = r0;
    out[2 * t + 1] = r1;
    if (kShade && r0.valid) {
        ShadeTri s0, s1;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < n) spoly[k].invW = 1.0f / poly[k].w;
        s0.v[0] = spoly[0]; s0.v[1] = spoly[1]; s0.v[2] = spoly[2];
        shadeOut[2 * t] = s0;
        if (r1.valid) { s1.v[0] = spoly[0]; s1.v[1] = spoly[2]; s1.v[2] = spoly[3]; shadeOut[2 * t + 1] = s1; }
    }
}

MV_D long long edge_fn(int ax, int ay, int bx, int by, int px, int py)
{
    return (long long)(bx - ax) * (long long)(py - ay) - (long long)(by - ay) * (long long)(px - ax);
}
// top-left rule for a triangle with positive area in a y-down screen: a top edge is horizontal with the
// interior below it (dx > 0), a left edge goes up (dy < 0)
MV_D bool is_top_left(int ax, int ay, int bx, int by)
{
    const int dx = bx - ax, dy = by - ay;
    return (dy == 0 && dx > 0) || dy < 0;
}

template <bool kVisibility>
__global__ void __launch_bounds__(256) k_mesh_raster(const ScreenTri* __restrict__ tris, uint32_t numRecords, RasterTarget rt, unsigned long long* __restrict__ vis)
{
    const uint32_t rec = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (rec >= numRecords) return;
    ScreenTri t = tris[rec];
    if (!t.valid) return;
    long long area = edge_fn(t.x[0], t.y[0], t.x[1], t.y[1], t.x[2], t.y[2]);
    if (area == 0) return;
    if (area < 0) {   // make the winding positive: swap vertices 1 and 2
        int ti = t.x[1]; t.x[1] = t.x[2]; t.x[2] = ti;
        ti = t.y[1]; t.y[1] = t.y[2]; t.y[2] = ti;
        const float tz = t.z[1]; t.z[1] = t.z[2]; t.z[2] = tz;
        area = -area;
    }
    // pixels whose centre can be covered: centre (px + 0.5) * 256 inside [min, max]
    const int minX = min(t.x[0], min(t.x[1], t.x[2])), maxX = max(t.x[0], max(t.x[1], t.x[2]));
    const int minY = min(t.y[0], min(t.y[1], t.y[2])), maxY = max(t.y[0], max(t.y[1], t.y[2]));
    const int px0 = max((minX - 128 + 255) >> 8, 0), px1 = min((maxX - 128) >> 8, (int)rt.width - 1);
    const int py0 = max((minY - 128 + 255) >> 8, 0), py1 = min((maxY - 128) >> 8, (int)rt.height - 1);
    if (px0 > px1 || py0 > py1) return;
    // an edge that is not top-left excludes the pixels exactly on it
    const long long b0 = is_top_left(t.x[1], t.y[1], t.x[2], t.y[2]) ? 0 : 1;
    const long long b1 = is_top_left(t.x[2], t.y[2], t.x[0], t.y[0]) ? 0 : 1;
    const long long b2 = is_top_left(t.x[0], t.y[0], t.x[1], t.y[1]) ? 0 : 1;
    const double inv = (double)area;
    for (int by = py0; by <= py1; by += 4)
        for (int bx = px0; bx <= px1; bx += 8) {
            const int px = bx + (int)(lane & 7), py = by + (int)(lane >> 3);
            if (px > px1 || py > py1) continue;
            const int cx = px * 256 + 128, cy = py * 256 + 128;
            const long long e0 = edge_fn(t.x[1], t.y[1], t.x[2], t.y[2], cx, cy);
            const long long e1 = edge_fn(t.x[2], t.y[2], t.x[0], t.y[0], cx, cy);
            const long long e2 = edge_fn(t.x[0], t.y[0], t.x[1], t.y[1], cx, cy);
            if (e0 < b0 || e1 < b1 || e2 < b2) continue;
            const double zd = (((double)e0 * (double)t.z[0] + (double)e1 * (double)t.z[1]) + (double)e2 * (double)t.z[2]) / inv;
            const float z = (float)zd;
            if (!(z >= 0.0f && z <= 1.0f)) continue;       // depth clip
            if (kVisibility) atomicMin(vis + (size_t)py * rt.width + px, ((unsigned long long)__float_as_uint(z) << 32) | rec);   // LESS; a tie goes to the earlier triangle
            else atomicMin(rt.depthBits + (size_t)py * rt.width + px, __float_as_uint(z));   // LESS against the 1.0 clear
        }
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* p, size_t n, uint32_t v)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256) k_depth_to_d16(const uint32_t* __restrict__ bits, uint16_t* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = (uint16_t)floorf(__uint_as_float(bits[i]) * 65535.0f + 0.5f);   // D16_UNORM
}

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long* p, size_t n, unsigned long long v)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr unsigned long long kVisClear = (0x3f800000ull << 32) | 0xffffffffull;   // depth 1.0, no triangle

MV_D float shadow_pcf(const uint16_t* __restrict__ shadow, int S, V3 ls)   // ShadowMap, PSBasePass.hlsl:72-78 (LINEAR_LESS_EQUAL, clamp)
{
    const float uvx = ls.x * 0.5f + 0.5f, uvy = 1.0f - (ls.y * 0.5f + 0.5f), ref = ls.z - 0.0027f;
    const float fx = uvx * (float)S - 0.5f, fy = uvy * (float)S - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float wx = fx - flx, wy = fy - fly;
    const int ix = (int)flx, iy = (int)fly;
    auto tap = [&](int x, int y) {
        x = min(max(x, 0), S - 1); y = min(max(y, 0), S - 1);
        return ref <= (float)__ldg(shadow + (size_t)y * S + x) / 65535.0f ? 1.0f : 0.0f;
    };
    const float t00 = tap(ix, iy), t10 = tap(ix + 1, iy), t01 = tap(ix, iy + 1), t11 = tap(ix + 1, iy + 1);
    return lerp(lerp(t00, t10, wx), lerp(t01, t11, wx), wy);
}

// The pixel shader of the base pass over the visibility buffer: one thread per pixel.
__global__ void __launch_bounds__(256) k_mesh_shade(const unsigned long long* __restrict__ vis, const ScreenTri* __restrict__ tris, const ShadeTri* __restrict__ shade,
                                                   const BasePassCB cb, const uint16_t* __restrict__ shadow, int shadowSize,
                                                   uint32_t width, uint32_t height, float* __restrict__ depthOut, uint2* __restrict__ colorOut, uint32_t* __restrict__ velocityOut)
{
    const uint32_t px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= width || py >= height) return;
    const size_t pix = (size_t)py * width + px;
    const unsigned long long v = vis[pix];
    if ((uint32_t)v == 0xffffffffu) {      // ClearRenderTargetView / ClearDepthStencilView (MultiVolumes.cpp:661-663)
        depthOut[pix] = 1.0f;
        colorOut[pix] = pack_half4(V4{cb.clear[0], cb.clear[1], cb.clear[2], cb.clear[3]});
        velocityOut[pix] = 0u;
        return;
    }
    const uint32_t rec = (uint32_t)v;
    ScreenTri t = tris[rec];
    int order[3] = {0, 1, 2};
    if (edge_fn(t.x[0], t.y[0], t.x[1], t.y[1], t.x[2], t.y[2]) < 0) {   // the rasteriser's winding fix
        int ti = t.x[1]; t.x[1] = t.x[2]; t.x[2] = ti;
        ti = t.y[1]; t.y[1] = t.y[2]; t.y[2] = ti;
        order[1] = 2; order[2] = 1;
    }
    const int cx = (int)px * 256 + 128, cy = (int)py * 256 + 128;
    const double e[3] = {(double)edge_fn(t.x[1], t.y[1], t.x[2], t.y[2], cx, cy), (double)edge_fn(t.x[2], t.y[2], t.x[0], t.y[0], cx, cy),
                         (double)edge_fn(t.x[0], t.y[0], t.x[1], t.y[1], cx, cy)};
    const ShadeTri& st = shade[rec];
    const ShadeVertex& v0 = st.v[order[0]]; const ShadeVertex& v1 = st.v[order[1]]; const ShadeVertex& v2 = st.v[order[2]];
    // perspective-correct weights e_i / w_i, normalised; every attribute rounded once to fp32
    const double b0 = e[0] * (double)v0.invW, b1 = e[1] * (double)v1.invW, b2 = e[2] * (double)v2.invW, den = (b0 + b1) + b2;
    float a[15];
    const float* f0 = reinterpret_cast<const float*>(&v0); const float* f1 = reinterpret_cast<const float*>(&v1); const float* f2 = reinterpret_cast<const float*>(&v2);
#pragma unroll
    for (int k = 0; k < 15; ++k) a[k] = (float)((((b0 * (double)f0[k]) + b1 * (double)f1[k]) + b2 * (double)f2[k]) / den);
    const V3 wsPos = {a[0], a[1], a[2]}, norm = {a[3], a[4], a[5]}, ls = {a[6], a[7], a[8]};
    // PSBasePass.hlsl:94-153
    const float shadowT = shadowSize ? shadow_pcf(shadow, shadowSize, ls) : 1.0f;
    const V3 N = normalize(norm);
    const V2 csPos = {a[9] / a[11], a[10] / a[11]}, tsPos = {a[12] / a[14], a[13] / a[14]};
    const V2 velocity = {(csPos.x - tsPos.x) * 0.5f, (csPos.y - tsPos.y) * -0.5f};
    const V3 L = normalize(V3{cb.lightPos[0], cb.lightPos[1], cb.lightPos[2]});
    const float NoL = saturate(dot(N, L));
    const V3 V = normalize(V3{cb.eye[0], cb.eye[1], cb.eye[2]} - wsPos);
    const V3 H = normalize(V + L);
    const float NoH = saturate(dot(N, H)), NoV = saturate(dot(N, V));
    const V3 lightColor = {cb.lightColor[0] * cb.lightColor[3], cb.lightColor[1] * cb.lightColor[3], cb.lightColor[2] * cb.lightColor[3]};
    V3 ambient = {cb.ambient[0] * cb.ambient[3], cb.ambient[1] * cb.ambient[3], cb.ambient[2] * cb.ambient[3]};
    ambient = ambient * (N.y * 0.25f + 0.75f);                                       // lerp(0.5, 1.0, N.y * 0.5 + 0.5) as PSBasePass.cso folds it
    if (cb.hasSH) ambient = evaluate_sh_irradiance(cb.sh, N);
    const V3 diffuseBRDF = {0.318359375f, 0.191040039062f, 0.0636596679688f};          // g_baseColor / PI as the shipped DXIL holds it (0xH3518, 0xH321D, 0xH2C13)
    float p64 = NoH;                                                                 // pow(NoH, 64): six squarings
#pragma unroll
    for (int k = 0; k < 6; ++k) p64 = p64 * p64;
    const float om = 1.0f - NoV, om2 = om * om, fres5 = (om2 * om2) * om;             // pow(1 - NoV, 5)
    const float fresnel = (1.0f - fres5) * 0.0800170898438f + fres5;                  // Fresnel(NoV, 0.08): lerp as compiled, 0.08 -> 0xH2D1F
    const float spec = p64 * fresnel;
    V3 result = {diffuseBRDF.x * NoL + spec, diffuseBRDF.y * NoL + spec, diffuseBRDF.z * NoL + spec};
    result = {result.x * (lightColor.x * shadowT), result.y * (lightColor.y * shadowT), result.z * (lightColor.z * shadowT)};
    result = {result.x + diffuseBRDF.x * ambient.x, result.y + diffuseBRDF.y * ambient.y, result.z + diffuseBRDF.z * ambient.z};
    depthOut[pix] = __uint_as_float((uint32_t)(v >> 32));
    colorOut[pix] = pack_half4(V4{result.x, result.y, result.z, 1.0f});
    velocityOut[pix] = (uint32_t)f32_to_f16(velocity.x) | ((uint32_t)f32_to_f16(velocity.y) << 16);
}

// ---- host matrices (DirectXMath call sites ObjectRenderer.cpp:182-186), evaluated in double, rounded once ----
void look_at_lh(const float eye[3], const float at[3], const float up[3], double M[16])
{
    double z[3] = {(double)at[0] - eye[0], (double)at[1] - eye[1], (double)at[2] - eye[2]};
    double l = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (double& v : z) v /= l;
    double x[3] = {up[1] * z[2] - up[2] * z[1], up[2] * z[0] - up[0] * z[2], up[0] * z[1] - up[1] * z[0]};
    l = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (double& v : x) v /= l;
    const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    const double e[3] = {eye[0], eye[1], eye[2]};
    auto d3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const double m[16] = {x[0], y[0], z[0], 0, x[1], y[1], z[1], 0, x[2], y[2], z[2], 0, -d3(x, e), -d3(y, e), -d3(z, e), 1};
    memcpy(M, m, sizeof m);
}

void orthographic_lh(double w, double h, double zn, double zf, double M[16])
{
    const double m[16] = {2.0 / w, 0, 0, 0, 0, 2.0 / h, 0, 0, 0, 0, 1.0 / (zf - zn), 0, 0, 0, -zn / (zf - zn), 1};
    memcpy(M, m, sizeof m);
}

void mul44d(const double* A, const double* B, double* R)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            R[i * 4 + j] = s;
        }
}

int fail(int code, const char* fmt, const char* a = "")
{
    set_error(fmt, a);
    return code;
}

int raster_pass(Caster& c, const float wvp[16], uint32_t width, uint32_t height, uint32_t* bits)
{
    RasterTarget rt;
    memcpy(rt.wvp, wvp, sizeof rt.wvp);
    rt.width = width; rt.height = height; rt.depthBits = bits;
    const size_t n = (size_t)width * height;
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(bits, n, 0x3f800000u);   // ClearDepthStencilView(1.0)
    const uint32_t numTris = c.meshNumIndices / 3;
    if (numTris) {
        k_mesh_setup<false><<<(numTris + 255) / 256, 256, 0, c.stream>>>(c.dMeshPos, c.dMeshIdx, numTris, rt, static_cast<ScreenTri*>(c.dMeshTris), nullptr, BasePassCB{}, nullptr);
        const uint32_t records = 2 * numTris;
        k_mesh_raster<false><<<(records + 7) / 8, 256, 0, c.stream>>>(static_cast<const ScreenTri*>(c.dMeshTris), records, rt, nullptr);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MV_ERR_CUDA, "mesh raster launch failed: %s", cudaGetErrorString(e));
    return MV_OK;
}

} // namespace

} // namespace mv

#define MV_CUDA(expr)                                                                                     \
    do {                                                                                                  \
        const cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return MV_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)
#define MV_REQUIRE(cond) do { if (!(cond)) { set_error("invalid argument: %s", #cond); return MV_ERR_INVALID; } } while (0)
#define MV_ENTER(h)            \
    MV_REQUIRE(h != nullptr);  \
    Caster& c = h->c;          \
    MV_CUDA(cudaSetDevice(c.device)); \
    flush_deferred(c)

extern "C" {

// XUSGObjLoader.cpp:18-40, :166-228 with forDX = true, swapYZ = false: positions with z negated, the index
// list reversed (winding flipped for the left-handed frame). Faces with more than three corners are
// fanned; negative (relative) indices are resolved; texture / normal references are skipped.
int mv_obj_parse(const char* path, float** positions, uint32_t* numVertices, uint32_t** indices, uint32_t* numIndices)
try {
    MV_REQUIRE(path && positions && numVertices && indices && numIndices);
    *positions = nullptr; *indices = nullptr; *numVertices = 0; *numIndices = 0;
    FILE* f = fopen(path, "r");
    if (!f) { set_error("cannot open %s", path); return MV_ERR_INVALID; }
    std::vector<float> pos;
    std::vector<uint32_t> idx;
    std::vector<char> line(1 << 16);
    bool bad = false;
    while (fgets(line.data(), (int)line.size(), f)) {
        const char* s = line.data();
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            float x, y, z;
            if (sscanf(s + 2, "%f %f %f", &x, &y, &z) != 3) { bad = true; break; }
            pos.push_back(x); pos.push_back(y); pos.push_back(-z);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            std::vector<uint32_t> corner;
            const char* p = s + 2;
            for (;;) {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == '\0' || *p == '\n' || *p == '\r' || *p == '#') break;
                char* end = nullptr;
                const long v = strtol(p, &end, 10);
                if (end == p) { bad = true; break; }
                const long nv = (long)(pos.size() / 3);
                const long vi = v > 0 ? v - 1 : nv + v;
                if (vi < 0 || vi >= nv) { bad = true; break; }
                corner.push_back((uint32_t)vi);
                p = end;
                while (*p != '\0' && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;   // skip /vt/vn
            }
            if (bad || corner.size() < 3) { bad = true; break; }
            for (size_t k = 1; k + 1 < corner.size(); ++k) { idx.push_back(corner[0]); idx.push_back(corner[k]); idx.push_back(corner[k + 1]); }
        }
    }
    fclose(f);
    if (bad) { set_error("malformed OBJ: %s", path); return MV_ERR_INVALID; }
    for (size_t a = 0, b = idx.size(); a + 1 < b; ++a) { --b; const uint32_t t = idx[a]; idx[a] = idx[b]; idx[b] = t; }   // reverse(m_indices)
    float* P = (float*)malloc(std::max<size_t>(pos.size(), 1) * sizeof(float));
    uint32_t* I = (uint32_t*)malloc(std::max<size_t>(idx.size(), 1) * sizeof(uint32_t));
    if (!P || !I) { free(P); free(I); set_error("out of host memory"); return MV_ERR_NOMEM; }
    memcpy(P, pos.data(), pos.size() * sizeof(float));
    memcpy(I, idx.data(), idx.size() * sizeof(uint32_t));
    *positions = P; *indices = I;
    *numVertices = (uint32_t)(pos.size() / 3); *numIndices = (uint32_t)idx.size();
    return MV_OK;
} catch (const std::bad_alloc&) { set_error("%s: out of host memory while parsing", path); return MV_ERR_NOMEM; }

void mv_obj_free(float* positions, uint32_t* indices) { free(positions); free(indices); }

// createVB / createIB + the AABB extent that sizes the shadow frustum (ObjectRenderer.cpp:68-77)
int mv_mesh_set(mv_caster* h, const float* positions, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices)
{
    MV_ENTER(h);
    MV_REQUIRE(numIndices % 3 == 0);
    MV_REQUIRE((positions && indices) || numIndices == 0);
    for (uint32_t i = 0; i < numIndices; ++i) MV_REQUIRE(indices[i] < numVertices);
    MV_CUDA(cudaStreamSynchronize(c.stream));
    if (c.dMeshPos) { cudaFree(c.dMeshPos); c.dMeshPos = nullptr; }
    if (c.dMeshIdx) { cudaFree(c.dMeshIdx); c.dMeshIdx = nullptr; }
    if (c.dMeshTris) { cudaFree(c.dMeshTris); c.dMeshTris = nullptr; }
    if (c.dMeshNrm) { cudaFree(c.dMeshNrm); c.dMeshNrm = nullptr; }
    if (c.dMeshShade) { cudaFree(c.dMeshShade); c.dMeshShade = nullptr; }
    c.meshNumIndices = 0; c.meshExtent = 1.0f; c.meshHavePrev = false;
    if (numIndices == 0) return MV_OK;
    MV_CUDA(cudaMalloc(&c.dMeshPos, (size_t)numVertices * 3 * sizeof(float)));
    MV_CUDA(cudaMalloc(&c.dMeshIdx, (size_t)numIndices * sizeof(uint32_t)));
    MV_CUDA(cudaMalloc(&c.dMeshTris, (size_t)(numIndices / 3) * 2 * sizeof(ScreenTri)));
    MV_CUDA(cudaMemcpyAsync(c.dMeshPos, positions, (size_t)numVertices * 3 * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    MV_CUDA(cudaMemcpyAsync(c.dMeshIdx, indices, (size_t)numIndices * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    // ObjLoader::recomputeNormals (XUSGObjLoader.cpp:337-384): unit face normals accumulated per vertex in index order, renormalised.
    // Zero-area faces and vertices without a face are skipped (the reference divides by zero there and shades NaN).
    std::vector<float> nrm((size_t)numVertices * 3, 0.0f);
    for (uint32_t t = 0; t < numIndices / 3; ++t) {
        const float* p0 = positions + 3 * (size_t)indices[3 * t]; const float* p1 = positions + 3 * (size_t)indices[3 * t + 1]; const float* p2 = positions + 3 * (size_t)indices[3 * t + 2];
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (!(l > 0.0f)) continue;
        for (int k = 0; k < 3; ++k) n[k] /= l;
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) nrm[3 * (size_t)indices[3 * t + v] + k] += n[k];
    }
    for (uint32_t v = 0; v < numVertices; ++v) {
        float* n = nrm.data() + 3 * (size_t)v;
        const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (!(l > 0.0f)) { n[0] = 0.0f; n[1] = 1.0f; n[2] = 0.0f; continue; }
        for (int k = 0; k < 3; ++k) n[k] /= l;
    }
    MV_CUDA(cudaMalloc(&c.dMeshNrm, (size_t)numVertices * 3 * sizeof(float)));
    MV_CUDA(cudaMalloc(&c.dMeshShade, (size_t)(numIndices / 3) * 2 * sizeof(ShadeTri)));
    MV_CUDA(cudaMemcpyAsync(c.dMeshNrm, nrm.data(), nrm.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    float mn[3] = {kFltMax, kFltMax, kFltMax}, mx[3] = {-kFltMax, -kFltMax, -kFltMax};
    for (uint32_t v = 0; v < numVertices; ++v)
        for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], positions[3 * v + k]); mx[k] = fmaxf(mx[k], positions[3 * v + k]); }
    c.meshExtent = fmaxf(mx[0] - mn[0], fmaxf(mx[1] - mn[1], mx[2] - mn[2]));
    c.meshNumIndices = numIndices;
    return MV_OK;
}

int mv_mesh_load_obj(mv_caster* h, const char* path)
{
    MV_REQUIRE(h != nullptr);
    float* P; uint32_t* I; uint32_t nv, ni;
    int rc = mv_obj_parse(path, &P, &nv, &I, &ni);
    if (rc != MV_OK) return rc;
    rc = mv_mesh_set(h, P, nv, I, ni);
    mv_obj_free(P, I);
    return rc;
}

int mv_mesh_set_world(mv_caster* h, float scale, const float pos[3])   // ObjectRenderer::SetWorld, :147-153 (no rotation)
{
    MV_ENTER(h);
    MV_REQUIRE(pos);
    c.meshScale = scale;
    memcpy(c.meshPos, pos, 3 * sizeof(float));
    return MV_OK;
}

static int mesh_render_impl(mv_caster* h, const float viewProj[16], float shadowVpOut[16], bool basePass, const float* eye, const float* clear)
{
    MV_ENTER(h);
    MV_REQUIRE(viewProj && (eye || !basePass));
    c.inputsDirty = true;
    const uint32_t S = 1024;                                  // m_shadowMapSize, ObjectRenderer.cpp:42
    if (c.shadowSize != S) {
        if (c.dShadow) { MV_CUDA(cudaStreamSynchronize(c.stream)); MV_CUDA(cudaFree(c.dShadow)); c.dShadow = nullptr; }
        MV_CUDA(cudaMalloc(&c.dShadow, (size_t)S * S * sizeof(uint16_t)));
        c.shadowSize = S;
    }
    if (!c.dShadowBits) MV_CUDA(cudaMalloc(&c.dShadowBits, (size_t)S * S * sizeof(uint32_t)));
    c.cb.shadowSize = S;
    // world = scaling * translation (:149-152)
    const double s = c.meshScale;
    const double world[16] = {s, 0, 0, 0, 0, s, 0, 0, 0, 0, s, 0, c.meshPos[0], c.meshPos[1], c.meshPos[2], 1};
    double vp[16], wvp[16], lv[16], lp[16], lvp[16], swvp[16];
    for (int i = 0; i < 16; ++i) vp[i] = viewProj[i];
    mul44d(world, vp, wvp);
    const float origin[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    const double size = (double)(c.meshExtent * c.meshScale) * 1.5;   // m_sceneSize * 1.5, :76, :180
    look_at_lh(c.lightPt, origin, up, lv);
    orthographic_lh(size, size, 1.0, 200.0, lp);
    mul44d(lv, lp, lvp);
    mul44d(world, lvp, swvp);
    float wvpF[16], swvpF[16];
    for (int i = 0; i < 16; ++i) { wvpF[i] = (float)wvp[i]; swvpF[i] = (float)swvp[i]; if (shadowVpOut) shadowVpOut[i] = (float)lvp[i]; }
    int rc = raster_pass(c, swvpF, S, S, c.dShadowBits);
    if (rc != MV_OK) return rc;
    k_depth_to_d16<<<(unsigned)(((size_t)S * S + 255) / 256), 256, 0, c.stream>>>(c.dShadowBits, c.dShadow, (size_t)S * S);
    if (!basePass) return raster_pass(c, wvpF, c.d.width, c.d.height, reinterpret_cast<uint32_t*>(c.dDepth));

    // base pass (ObjectRenderer::Render, :532-553): visibility buffer, then the pixel shader per pixel
    const size_t px = (size_t)c.d.width * c.d.height;
    if (!c.dMeshVis) MV_CUDA(cudaMalloc(&c.dMeshVis, px * sizeof(unsigned long long)));
    BasePassCB cb{};
    memcpy(cb.wvp, wvpF, sizeof cb.wvp);
    memcpy(cb.wvpPrev, c.meshHavePrev ? c.meshWvpPrev : wvpF, sizeof cb.wvpPrev);      // WorldViewProjPrev = m_worldViewProj (:203-204)
    memcpy(c.meshWvpPrev, wvpF, sizeof c.meshWvpPrev); c.meshHavePrev = true;
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 3; ++k) cb.world[r * 3 + k] = (float)world[r * 4 + k];
    memcpy(cb.shadowWVP, swvpF, sizeof cb.shadowWVP);
    memcpy(cb.eye, eye, sizeof cb.eye);
    memcpy(cb.lightPos, c.lightPt, sizeof cb.lightPos);
    memcpy(cb.lightColor, c.lightColor, sizeof cb.lightColor);
    memcpy(cb.ambient, c.ambient, sizeof cb.ambient);
    cb.hasSH = c.cb.hasSH;
    memcpy(cb.sh, c.cb.sh, sizeof cb.sh);
    for (int k = 0; k < 4; ++k) cb.clear[k] = clear ? clear[k] : 0.0f;
    RasterTarget rt;
    memcpy(rt.wvp, wvpF, sizeof rt.wvp);
    rt.width = c.d.width; rt.height = c.d.height; rt.depthBits = nullptr;
    k_fill_u64<<<(unsigned)((px + 255) / 256), 256, 0, c.stream>>>(c.dMeshVis, px, kVisClear);
    const uint32_t numTris = c.meshNumIndices / 3;
    if (numTris) {
        k_mesh_setup<true><<<(numTris + 255) / 256, 256, 0, c.stream>>>(c.dMeshPos, c.dMeshIdx, numTris, rt, static_cast<ScreenTri*>(c.dMeshTris), c.dMeshNrm, cb,
                                                                        static_cast<ShadeTri*>(c.dMeshShade));
        k_mesh_raster<true><<<(2 * numTris + 7) / 8, 256, 0, c.stream>>>(static_cast<const ScreenTri*>(c.dMeshTris), 2 * numTris, rt, c.dMeshVis);
    }
    k_mesh_shade<<<dim3((c.d.width + 31) / 32, (c.d.height + 7) / 8), 256, 0, c.stream>>>(
        c.dMeshVis, static_cast<const ScreenTri*>(c.dMeshTris), static_cast<const ShadeTri*>(c.dMeshShade), cb, c.dShadow, (int)S, c.d.width, c.d.height, c.dDepth,
        reinterpret_cast<uint2*>(c.dBackground), reinterpret_cast<uint32_t*>(c.dVelocity));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MV_ERR_CUDA, "mesh base pass launch failed: %s", cudaGetErrorString(e));
    MV_CUDA(cudaMemcpyAsync(c.dColor, c.dBackground, px * 8, cudaMemcpyDeviceToDevice, c.stream));
    c.velocityGiven = true;
    return MV_OK;
}

// ObjectRenderer::UpdateFrame (:171-190) + RenderShadow (:220-243) + the depth pre-pass (:555-570): fills the
// caster's scene depth and shadow map from the mesh under view_proj, and returns the light's
// view-projection (row-vector convention, as mv_update_frame takes it).
int mv_mesh_render_depth(mv_caster* h, const float viewProj[16], float shadowVpOut[16]) { return mesh_render_impl(h, viewProj, shadowVpOut, false, nullptr, nullptr); }

// ... + ObjectRenderer::Render (:532-553): the shaded base pass. Fills scene depth, the shadow map, the background
// colour the resolve composites over (clear_rgba where no triangle covers the pixel) and the TAA velocity field.
// view_proj and eye are ObjectRenderer::UpdateFrame's arguments (:171); the light, ambient and SH coefficients are the
// ones the caster holds (mv_set_light / mv_set_ambient / mv_set_sh).
int mv_mesh_render(mv_caster* h, const float viewProj[16], const float eye[3], const float clearRgba[4], float shadowVpOut[16])
{
    return mesh_render_impl(h, viewProj, shadowVpOut, true, eye, clearRgba);
}

int mv_read_velocity(mv_caster* h, uint16_t* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    MV_CUDA(cudaMemcpyAsync(out, c.dVelocity, (size_t)c.d.width * c.d.height * 4, cudaMemcpyDeviceToHost, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return MV_OK;
}

int mv_read_depth(mv_caster* h, float* depth, uint16_t* shadow, uint32_t* shadowSize)
{
    MV_ENTER(h);
    if (depth) MV_CUDA(cudaMemcpyAsync(depth, c.dDepth, (size_t)c.d.width * c.d.height * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    if (shadow && c.shadowSize) MV_CUDA(cudaMemcpyAsync(shadow, c.dShadow, (size_t)c.shadowSize * c.shadowSize * sizeof(uint16_t), cudaMemcpyDeviceToHost, c.stream));
    if (shadowSize) *shadowSize = c.shadowSize;
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return MV_OK;
}

} // extern "C"
