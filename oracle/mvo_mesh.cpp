// mvo_mesh.cpp — ORACLE (test infrastructure, not product code): scalar restatement of the depth-only
// passes that produce the volume path's scene depth and shadow map (reference: ObjectRenderer.cpp:171-190
// light view-projection, :220-243 RenderShadow, :555-570 renderDepth, VSDepth.hlsl:25-28).
//
// The reference's rasteriser is fixed-function hardware, so there is no source to follow for coverage and
// depth interpolation; what is restated here are the Direct3D rasterisation rules as written in the
// header of multivolumes_b200/csrc/mv_mesh.cu (near-plane clip, 1/256-pixel snapping, integer edge
// functions with the top-left rule, fp64 barycentric depth, LESS test against a 1.0 clear, no culling,
// D16 = floor(z * 65535 + 0.5)). Triangles are processed one after the other in index order; the rules
// make the result independent of that order. "Parity unpinned" against the D3D12 hardware rasteriser.
//
// Round 2: the shaded base pass over the same rasteriser (ObjectRenderer::Render, ObjectRenderer.cpp:532-553;
// VSBasePass.hlsl:39-55; PSBasePass.hlsl:94-153). Interpolants are perspective-correct, evaluated in fp64 from the
// integer edge values and 1 / w and rounded once. Declared deviations (same as the product): no sub-pixel jitter,
// no radiance term, zero velocity on the first frame. The normals follow ObjLoader::recomputeNormals
// (XUSGObjLoader.cpp:337-384).
#include "mvo_core.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace mvo {

namespace {

struct Vtx { int x, y; float z; };

Vtx to_screen(f4 c, uint32_t width, uint32_t height)
{
    const float ndcX = c.x / c.w, ndcY = c.y / c.w;
    Vtx v;
    v.z = c.z / c.w;
    const float sx = (ndcX * 0.5f + 0.5f) * (float)width;
    const float sy = (1.0f - (ndcY * 0.5f + 0.5f)) * (float)height;
    const float lim = 4194304.0f;
    v.x = (int)floorf(fminf(fmaxf(sx, -lim), lim) * 256.0f + 0.5f);
    v.y = (int)floorf(fminf(fmaxf(sy, -lim), lim) * 256.0f + 0.5f);
    return v;
}

int64_t edge_fn(const Vtx& a, const Vtx& b, int px, int py)
{
    return (int64_t)(b.x - a.x) * (int64_t)(py - a.y) - (int64_t)(b.y - a.y) * (int64_t)(px - a.x);
}

bool top_left(const Vtx& a, const Vtx& b)
{
    const int dx = b.x - a.x, dy = b.y - a.y;
    return (dy == 0 && dx > 0) || dy < 0;
}

void raster_triangle(Vtx v0, Vtx v1, Vtx v2, uint32_t width, uint32_t height, float* depth)
{
    int64_t area = edge_fn(v0, v1, v2.x, v2.y);
    if (area == 0) return;
    if (area < 0) { std::swap(v1, v2); area = -area; }
    const int minX = std::min(v0.x, std::min(v1.x, v2.x)), maxX = std::max(v0.x, std::max(v1.x, v2.x));
    const int minY = std::min(v0.y, std::min(v1.y, v2.y)), maxY = std::max(v0.y, std::max(v1.y, v2.y));
    const int px0 = std::max((minX - 128 + 255) >> 8, 0), px1 = std::min((maxX - 128) >> 8, (int)width - 1);
    const int py0 = std::max((minY - 128 + 255) >> 8, 0), py1 = std::min((maxY - 128) >> 8, (int)height - 1);
    const int64_t b0 = top_left(v1, v2) ? 0 : 1, b1 = top_left(v2, v0) ? 0 : 1, b2 = top_left(v0, v1) ? 0 : 1;
    for (int py = py0; py <= py1; ++py)
        for (int px = px0; px <= px1; ++px) {
            const int cx = px * 256 + 128, cy = py * 256 + 128;
            const int64_t e0 = edge_fn(v1, v2, cx, cy), e1 = edge_fn(v2, v0, cx, cy), e2 = edge_fn(v0, v1, cx, cy);
            if (e0 < b0 || e1 < b1 || e2 < b2) continue;
            const double zd = (((double)e0 * (double)v0.z + (double)e1 * (double)v1.z) + (double)e2 * (double)v2.z) / (double)area;
            const float z = (float)zd;
            if (!(z >= 0.0f && z <= 1.0f)) continue;
            float& d = depth[(size_t)py * width + px];
            if (z < d) d = z;
        }
}

f4 lerp4(f4 a, f4 b, float t) { return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t}; }

// what VSBasePass hands to PSBasePass: WSPos, Norm, LSPos.xyz, CSPos.xyw, TSPos.xyw (15 floats) + 1 / Pos.w
struct ShadeVtx { float a[15]; float invW; };

ShadeVtx lerp_sv(const ShadeVtx& p, const ShadeVtx& q, float t)
{
    ShadeVtx r;
    for (int k = 0; k < 15; ++k) r.a[k] = p.a[k] + (q.a[k] - p.a[k]) * t;
    r.invW = 0.0f;
    return r;
}

struct VisPixel { float z; uint32_t rec; };

// the depth rasteriser again, recording which (triangle, half) covers the pixel: LESS, so a tie stays with the earlier record
void raster_visibility(Vtx v0, Vtx v1, Vtx v2, uint32_t rec, uint32_t width, uint32_t height, VisPixel* vis)
{
    int64_t area = edge_fn(v0, v1, v2.x, v2.y);
    if (area == 0) return;
    if (area < 0) { std::swap(v1, v2); area = -area; }
    const int minX = std::min(v0.x, std::min(v1.x, v2.x)), maxX = std::max(v0.x, std::max(v1.x, v2.x));
    const int minY = std::min(v0.y, std::min(v1.y, v2.y)), maxY = std::max(v0.y, std::max(v1.y, v2.y));
    const int px0 = std::max((minX - 128 + 255) >> 8, 0), px1 = std::min((maxX - 128) >> 8, (int)width - 1);
    const int py0 = std::max((minY - 128 + 255) >> 8, 0), py1 = std::min((maxY - 128) >> 8, (int)height - 1);
    const int64_t b0 = top_left(v1, v2) ? 0 : 1, b1 = top_left(v2, v0) ? 0 : 1, b2 = top_left(v0, v1) ? 0 : 1;
    for (int py = py0; py <= py1; ++py)
        for (int px = px0; px <= px1; ++px) {
            const int cx = px * 256 + 128, cy = py * 256 + 128;
            const int64_t e0 = edge_fn(v1, v2, cx, cy), e1 = edge_fn(v2, v0, cx, cy), e2 = edge_fn(v0, v1, cx, cy);
            if (e0 < b0 || e1 < b1 || e2 < b2) continue;
            const double zd = (((double)e0 * (double)v0.z + (double)e1 * (double)v1.z) + (double)e2 * (double)v2.z) / (double)area;
            const float z = (float)zd;
            if (!(z >= 0.0f && z <= 1.0f)) continue;
            VisPixel& d = vis[(size_t)py * width + px];
            if (z < d.z || (z == d.z && rec < d.rec)) { d.z = z; d.rec = rec; }
        }
}

float shadow_pcf(const Caster& c, f3 ls)   // ShadowMap, PSBasePass.hlsl:72-78: 2x2 comparison taps, bilinear weights, clamp addressing
{
    const int S = (int)c.shadowSize;
    const float uvx = ls.x * 0.5f + 0.5f, uvy = 1.0f - (ls.y * 0.5f + 0.5f), ref = ls.z - 0.0027f;
    const float fx = uvx * (float)S - 0.5f, fy = uvy * (float)S - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float wx = fx - flx, wy = fy - fly;
    const int ix = (int)flx, iy = (int)fly;
    auto tap = [&](int x, int y) {
        x = std::min(std::max(x, 0), S - 1); y = std::min(std::max(y, 0), S - 1);
        return ref <= (float)c.shadow[(size_t)y * S + x] / 65535.0f ? 1.0f : 0.0f;
    };
    const float t00 = tap(ix, iy), t10 = tap(ix + 1, iy), t01 = tap(ix, iy + 1), t11 = tap(ix + 1, iy + 1);
    return lerp1(lerp1(t00, t10, wx), lerp1(t01, t11, wx), wy);
}

} // namespace

void recompute_normals(const std::vector<float>& pos, const std::vector<uint32_t>& idx, std::vector<float>& nrm)
{
    nrm.assign(pos.size(), 0.0f);
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        const float* p0 = pos.data() + 3 * (size_t)idx[t]; const float* p1 = pos.data() + 3 * (size_t)idx[t + 1]; const float* p2 = pos.data() + 3 * (size_t)idx[t + 2];
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (!(l > 0.0f)) continue;                       // zero-area face: the reference divides by zero here (NaN normals)
        for (int k = 0; k < 3; ++k) n[k] /= l;
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) nrm[3 * (size_t)idx[t + v] + k] += n[k];
    }
    for (size_t v = 0; v + 2 < nrm.size(); v += 3) {
        float* n = nrm.data() + v;
        const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (!(l > 0.0f)) { n[0] = 0.0f; n[1] = 1.0f; n[2] = 0.0f; continue; }
        for (int k = 0; k < 3; ++k) n[k] /= l;
    }
}

void render_base_pass(Caster& c, const m44& wvp, const m44& wvpPrev, const m43& world, const m44& shadowWVP, f3 eye, const float clear[4])
{
    const uint32_t width = c.d.width, height = c.d.height;
    const size_t px = (size_t)width * height;
    std::vector<VisPixel> vis(px, VisPixel{1.0f, 0xffffffffu});
    const size_t numTris = c.meshIdx.size() / 3;
    std::vector<Vtx> screen(numTris * 2 * 3);
    std::vector<ShadeVtx> shade(numTris * 2 * 3);
    for (size_t t = 0; t < numTris; ++t) {
        f4 cpos[3]; ShadeVtx sv[3];
        for (int k = 0; k < 3; ++k) {      // VSBasePass.hlsl:44-53
            const uint32_t v = c.meshIdx[3 * t + k];
            const f3 p = {c.meshPos[3 * v], c.meshPos[3 * v + 1], c.meshPos[3 * v + 2]};
            cpos[k] = mul_p44(p, wvp);
            const f3 ws = mul_p43(p, world);
            const f3 n = mul_v33(f3{c.meshNrm[3 * v], c.meshNrm[3 * v + 1], c.meshNrm[3 * v + 2]}, world);
            const f4 ls = mul_p44(p, shadowWVP), ts = mul_p44(p, wvpPrev);
            const float a[15] = {ws.x, ws.y, ws.z, n.x, n.y, n.z, ls.x, ls.y, ls.z, cpos[k].x, cpos[k].y, cpos[k].w, ts.x, ts.y, ts.w};
            memcpy(sv[k].a, a, sizeof a); sv[k].invW = 0.0f;
        }
        f4 poly[4]; ShadeVtx spoly[4]; int n = 0;
        for (int k = 0; k < 3; ++k) {
            const f4 a = cpos[k], b = cpos[(k + 1) % 3];
            const bool ain = a.z >= 0.0f, bin = b.z >= 0.0f;
            if (ain) { spoly[n] = sv[k]; poly[n++] = a; }
            if (ain != bin) {
                const f4 p = ain ? a : b, q = ain ? b : a;
                const float tt = p.z / (p.z - q.z);
                spoly[n] = lerp_sv(ain ? sv[k] : sv[(k + 1) % 3], ain ? sv[(k + 1) % 3] : sv[k], tt);
                poly[n++] = lerp4(p, q, tt);
            }
        }
        if (n < 3) continue;
        bool ok = true;
        Vtx v[4];
        for (int k = 0; k < n; ++k) {
            if (!(poly[k].w > 0.0f)) { ok = false; break; }
            v[k] = to_screen(poly[k], width, height);
            spoly[k].invW = 1.0f / poly[k].w;
        }
        if (!ok) continue;
        const int order[2][3] = {{0, 1, 2}, {0, 2, 3}};
        for (int half = 0; half < (n == 4 ? 2 : 1); ++half) {
            const uint32_t rec = (uint32_t)(2 * t + half);
            for (int k = 0; k < 3; ++k) { screen[3 * (size_t)rec + k] = v[order[half][k]]; shade[3 * (size_t)rec + k] = spoly[order[half][k]]; }
            raster_visibility(v[order[half][0]], v[order[half][1]], v[order[half][2]], rec, width, height, vis.data());
        }
    }

    c.depth.resize(px); c.background.resize(px * 4); c.color.resize(px * 4); c.velocity.resize(px * 2);
    const f3 L = normalize3(c.lightPt);
    for (uint32_t py = 0; py < height; ++py)
        for (uint32_t pxl = 0; pxl < width; ++pxl) {
            const size_t pix = (size_t)py * width + pxl;
            const VisPixel vp = vis[pix];
            if (vp.rec == 0xffffffffu) {
                c.depth[pix] = 1.0f;
                for (int k = 0; k < 4; ++k) c.background[4 * pix + k] = f32_to_f16(clear[k]);
                c.velocity[2 * pix] = c.velocity[2 * pix + 1] = 0;
                continue;
            }
            Vtx v0 = screen[3 * (size_t)vp.rec], v1 = screen[3 * (size_t)vp.rec + 1], v2 = screen[3 * (size_t)vp.rec + 2];
            const ShadeVtx* s0 = &shade[3 * (size_t)vp.rec]; const ShadeVtx* s1 = s0 + 1; const ShadeVtx* s2 = s0 + 2;
            if (edge_fn(v0, v1, v2.x, v2.y) < 0) { std::swap(v1, v2); std::swap(s1, s2); }
            const int cx = (int)pxl * 256 + 128, cy = (int)py * 256 + 128;
            const double e0 = (double)edge_fn(v1, v2, cx, cy), e1 = (double)edge_fn(v2, v0, cx, cy), e2 = (double)edge_fn(v0, v1, cx, cy);
            const double b0 = e0 * (double)s0->invW, b1 = e1 * (double)s1->invW, b2 = e2 * (double)s2->invW, den = (b0 + b1) + b2;
            float a[15];
            for (int k = 0; k < 15; ++k) a[k] = (float)((((b0 * (double)s0->a[k]) + b1 * (double)s1->a[k]) + b2 * (double)s2->a[k]) / den);
            const f3 wsPos = {a[0], a[1], a[2]}, norm = {a[3], a[4], a[5]}, ls = {a[6], a[7], a[8]};
            // PSBasePass.hlsl:94-153
            const float shadowT = c.shadowSize ? shadow_pcf(c, ls) : 1.0f;
            const f3 N = normalize3(norm);
            const f2 csPos = {a[9] / a[11], a[10] / a[11]}, tsPos = {a[12] / a[14], a[13] / a[14]};
            const f2 velocity = {(csPos.x - tsPos.x) * 0.5f, (csPos.y - tsPos.y) * -0.5f};
            const float NoL = saturate(dot3(N, L));
            const f3 V = normalize3(eye - wsPos);
            const f3 H = normalize3(V + L);
            const float NoH = saturate(dot3(N, H)), NoV = saturate(dot3(N, V));
            const f3 lightColor = {c.lightColor.x * c.lightColor.w, c.lightColor.y * c.lightColor.w, c.lightColor.z * c.lightColor.w};
            f3 ambient = {c.ambient.x * c.ambient.w, c.ambient.y * c.ambient.w, c.ambient.z * c.ambient.w};
            ambient = ambient * (N.y * 0.25f + 0.75f);        // lerp(0.5, 1.0, N.y * 0.5 + 0.5) as PSBasePass.cso folds it
            if (c.hasSH) { const f4 irr = evaluate_sh_irradiance(c.sh, N); ambient = {irr.x, irr.y, irr.z}; }
            const f3 diffuseBRDF = {0.318359375f, 0.191040039062f, 0.0636596679688f};   // g_baseColor / PI in the shipped DXIL: 0xH3518, 0xH321D, 0xH2C13
            float p64 = NoH;
            for (int k = 0; k < 6; ++k) p64 = p64 * p64;
            const float om = 1.0f - NoV, om2 = om * om, fres5 = (om2 * om2) * om;
            const float fresnel = (1.0f - fres5) * 0.0800170898438f + fres5;           // lerp as compiled; 0.08 -> 0xH2D1F
            const float spec = p64 * fresnel;
            f3 result = {diffuseBRDF.x * NoL + spec, diffuseBRDF.y * NoL + spec, diffuseBRDF.z * NoL + spec};
            result = {result.x * (lightColor.x * shadowT), result.y * (lightColor.y * shadowT), result.z * (lightColor.z * shadowT)};
            result = {result.x + diffuseBRDF.x * ambient.x, result.y + diffuseBRDF.y * ambient.y, result.z + diffuseBRDF.z * ambient.z};
            c.depth[pix] = vp.z;
            c.background[4 * pix] = f32_to_f16(result.x); c.background[4 * pix + 1] = f32_to_f16(result.y);
            c.background[4 * pix + 2] = f32_to_f16(result.z); c.background[4 * pix + 3] = f32_to_f16(1.0f);
            c.velocity[2 * pix] = f32_to_f16(velocity.x); c.velocity[2 * pix + 1] = f32_to_f16(velocity.y);
        }
    c.color = c.background;
}

void raster_depth(const std::vector<float>& pos, const std::vector<uint32_t>& idx, const m44& wvp, uint32_t width, uint32_t height, float* depth)
{
    std::fill(depth, depth + (size_t)width * height, 1.0f);
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        f4 c[3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = idx[t + k];
            c[k] = mul_p44(f3{pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]}, wvp);
        }
        f4 poly[4]; int n = 0;
        for (int k = 0; k < 3; ++k) {
            const f4 a = c[k], b = c[(k + 1) % 3];
            const bool ain = a.z >= 0.0f, bin = b.z >= 0.0f;
            if (ain) poly[n++] = a;
            if (ain != bin) {
                const f4 p = ain ? a : b, q = ain ? b : a;
                poly[n++] = lerp4(p, q, p.z / (p.z - q.z));
            }
        }
        if (n < 3) continue;
        bool ok = true;
        Vtx v[4];
        for (int k = 0; k < n; ++k) {
            if (!(poly[k].w > 0.0f)) { ok = false; break; }
            v[k] = to_screen(poly[k], width, height);
        }
        if (!ok) continue;
        raster_triangle(v[0], v[1], v[2], width, height, depth);
        if (n == 4) raster_triangle(v[0], v[2], v[3], width, height, depth);
    }
}

} // namespace mvo
