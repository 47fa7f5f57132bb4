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

} // namespace

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
