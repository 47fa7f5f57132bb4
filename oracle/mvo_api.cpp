// mvo_api.cpp — ORACLE (test infrastructure, not product code): C API over the restated passes.
// Host-side scene maths follows MultiRayCaster.cpp:266-353 (SetVolumesWorld / SetVolumeWorld /
// UpdateFrame) of the reference.
#include "mvo_core.h"
#include <omp.h>
#include <new>

using namespace mvo;

struct mvo_caster { Caster c; };

static int ok_src(Caster& c, uint32_t s) { return s < c.d.num_volume_srcs; }
static int ok_vol(Caster& c, uint32_t v) { return v < c.d.num_volumes; }

static void set_volume_world(Caster& c, uint32_t i, float size, const float pos[3])   // MultiRayCaster.cpp:297-303
{
    size *= 0.5f;
    m43& w = c.volumeWorlds[i];
    w = {{{size, 0, 0}, {0, size, 0}, {0, 0, size}, {pos[0], pos[1], pos[2]}}};
}

static void set_volumes_world(Caster& c, float size, const float center[3])   // MultiRayCaster.cpp:277-295
{
    const uint32_t numVolumes = c.d.num_volumes;
    const uint32_t rowLength = (uint32_t)ceilf(sqrtf((float)numVolumes));
    const uint32_t colLength = (uint32_t)ceilf((float)(numVolumes / rowLength));   // integer division first, as in the reference
    float pos[3] = {center[0], center[1], center[2]};
    pos[2] -= ((float)colLength / 2.0f - 0.5f) * size * 1.5f;
    for (uint32_t m = 0; m < colLength; ++m) {
        pos[0] = center[0] - ((float)rowLength / 2.0f - 0.5f) * size * 1.5f;
        for (uint32_t n = 0; n < rowLength; ++n) {
            set_volume_world(c, rowLength * m + n, size, pos);
            pos[0] += size * 1.5f;
        }
        pos[2] += size * 1.5f;
    }
}

extern "C" {

int mvo_create(const mvo_desc* d, mvo_caster** out)
{
    if (!d || !out || d->grid_size == 0 || d->num_volumes == 0 || d->num_volume_srcs == 0 || d->width == 0 || d->height == 0) return -1;
    if (d->grid_size >= (1u << 14) || d->num_volume_srcs >= (1u << 14) || (d->grid_size >> (kNumCubeMip - 1)) == 0) return -1;
    mvo_caster* h = new (std::nothrow) mvo_caster();
    if (!h) return -2;
    Caster& c = h->c;
    c.d = *d;
    if (c.d.light_grid_size == 0) c.d.light_grid_size = 96;
    if (c.d.max_ray_samples == 0) c.d.max_ray_samples = 256;
    if (c.d.max_light_samples == 0) c.d.max_light_samples = 96;
    c.filterModel = d->tex_filter_model ? MODEL_SM100 : MODEL_EXACT;
    if (d->num_threads) omp_set_num_threads((int)d->num_threads);
    const uint32_t G = c.d.grid_size, L = c.d.light_grid_size, N = c.d.num_volumes;
    c.volumes.resize(c.d.num_volume_srcs);
    for (auto& t : c.volumes) { t.n = G; t.texels.assign((size_t)G * G * G * 4, 0); }
    c.lightMaps.resize(N);
    for (auto& t : c.lightMaps) { t.n = L; t.texels.assign((size_t)L * L * L * 4, 0); }
    c.cubeMaps.resize(N);
    for (auto& cm : c.cubeMaps)
        for (uint32_t m = 0; m < kNumCubeMip; ++m) {
            const size_t s = G >> m;
            cm.color[m].assign(6 * s * s * 4, 0);
            cm.depth[m].assign(6 * s * s, 0.0f);
        }
    c.volumeWorlds.assign(N, m43{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}}});
    c.perObject.resize(N);
    c.volumeDescs.resize(N);
    for (uint32_t i = 0; i < N; ++i)   // MultiRayCaster.cpp:470-479
        c.volumeDescs[i] = (i % c.d.num_volume_srcs) | (kNumCubeMip << 14) | (G << 18);
    c.attribs.assign((size_t)N * 4, 0);
    const size_t px = (size_t)c.d.width * c.d.height;
    c.depth.assign(px, 1.0f);
    c.shadowSize = 0;
    c.color.assign(px * 4, 0);
    c.velocity.assign(px * 2, 0);
    c.taaHistory[0].assign(px * 4, 0); c.taaHistory[1].assign(px * 4, 0);
    c.backBuffer.assign(px * 4, 0);
    // MultiRayCaster.cpp:62-64 defaults
    c.lightPt = {75.0f, 75.0f, -75.0f};
    c.lightColor = {1.0f, 0.7f, 0.3f, 1.0f};
    c.ambient = {0.0f, 0.3f, 1.0f, 0.4f};
    c.stats = mvo_stats();
    c.stats.threads = (uint32_t)omp_get_max_threads();
    c.row0 = 0; c.row1 = c.d.height;
    const float center[3] = {0, 0, 0};
    set_volumes_world(c, 20.0f, center);
    *out = h;
    return 0;
}

void mvo_destroy(mvo_caster* h) { delete h; }

int mvo_volume_init_procedural(mvo_caster* h, uint32_t src, uint32_t mode, uint32_t seed)
{
    if (!h || !ok_src(h->c, src)) return -1;
    init_grid_data(h->c, src, mode, seed);
    return 0;
}
int mvo_volume_upload_rgba16f(mvo_caster* h, uint32_t src, const uint16_t* texels)
{
    if (!h || !texels || !ok_src(h->c, src)) return -1;
    auto& t = h->c.volumes[src];
    std::copy(texels, texels + t.texels.size(), t.texels.begin());
    return 0;
}
int mvo_volume_upload_r32f(mvo_caster* h, uint32_t src, const float* density)   // CSR32FToRGBA16F.hlsl:16-26
{
    if (!h || !density || !ok_src(h->c, src)) return -1;
    auto& t = h->c.volumes[src];
    const size_t n = (size_t)t.n * t.n * t.n;
    const uint16_t one = f32_to_f16(1.0f);
    for (size_t i = 0; i < n; ++i) {
        t.texels[i * 4 + 0] = one; t.texels[i * 4 + 1] = one; t.texels[i * 4 + 2] = one;
        t.texels[i * 4 + 3] = f32_to_f16(density[i] * 0.25f);
    }
    return 0;
}
// CSR32FToRGBA16F.hlsl:16-26 on a source of any resolution: LINEAR / clamp fetch at the voxel centres. Texel
// addressing and the 8-bit weights follow the texture-unit model of mvo_sampler.h; the fp32 texels are blended
// in double (the unit's internal precision for 32-bit float texels is not modelled: the parity test allows
// one fp16 ulp on resampled volumes and is exact when the source already has the grid's resolution).
int mvo_volume_upload_r32f_sized(mvo_caster* h, uint32_t src, const float* density, uint32_t w, uint32_t hh, uint32_t d)
{
    if (!h || !density || !ok_src(h->c, src) || !w || !hh || !d) return -1;
    auto& t = h->c.volumes[src];
    const int n = (int)t.n;
    const uint16_t one = f32_to_f16(1.0f);
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const float gs = (float)n;
                const AxisFix ax = axis_sm100(((float)x + 0.5f) / gs, (int)w), ay = axis_sm100(((float)y + 0.5f) / gs, (int)hh),
                              az = axis_sm100(((float)z + 0.5f) / gs, (int)d);
                int wt[8];
                weights_sm100(ax.frac, ay.frac, az.frac, wt);
                double acc = 0.0;
                for (int k = 0; k < 8; ++k) {
                    const int xi = (k & 1) ? ax.i1 : ax.i0, yi = (k & 2) ? ay.i1 : ay.i0, zi = (k & 4) ? az.i1 : az.i0;
                    acc += (double)wt[k] * (double)density[((size_t)zi * hh + yi) * w + xi];
                }
                const float a = (float)(acc / 256.0);
                const size_t i = ((size_t)z * n + y) * n + x;
                t.texels[i * 4 + 0] = one; t.texels[i * 4 + 1] = one; t.texels[i * 4 + 2] = one;
                t.texels[i * 4 + 3] = f32_to_f16(a * 0.25f);
            }
    return 0;
}
int mvo_volume_read(mvo_caster* h, uint32_t src, uint16_t* out)
{
    if (!h || !out || !ok_src(h->c, src)) return -1;
    auto& t = h->c.volumes[src];
    std::copy(t.texels.begin(), t.texels.end(), out);
    return 0;
}

int mvo_set_targets(mvo_caster* h, const float* depth, const uint16_t* shadow, uint32_t shadowSize, const uint16_t* color, const uint16_t* velocity)
{
    if (!h) return -1;
    Caster& c = h->c;
    const size_t px = (size_t)c.d.width * c.d.height;
    if (depth) std::copy(depth, depth + px, c.depth.begin()); else c.depth.assign(px, 1.0f);
    if (shadow && shadowSize) { c.shadowSize = shadowSize; c.shadow.assign(shadow, shadow + (size_t)shadowSize * shadowSize); }
    else { c.shadowSize = 0; c.shadow.clear(); }
    if (color) std::copy(color, color + px * 4, c.color.begin()); else c.color.assign(px * 4, 0);
    c.background = c.color;
    if (velocity) std::copy(velocity, velocity + px * 2, c.velocity.begin()); else c.velocity.assign(px * 2, 0);
    return 0;
}
int mvo_reset_color(mvo_caster* h)
{
    if (!h) return -1;
    if (h->c.background.size() == h->c.color.size()) h->c.color = h->c.background; else std::fill(h->c.color.begin(), h->c.color.end(), 0);
    return 0;
}
int mvo_set_sh(mvo_caster* h, const float* k)
{
    if (!h) return -1;
    h->c.hasSH = k != nullptr;
    if (k) for (int i = 0; i < 9; ++i) h->c.sh[i] = {k[i * 3], k[i * 3 + 1], k[i * 3 + 2]};
    return 0;
}
int mvo_set_max_samples(mvo_caster* h, uint32_t ray, uint32_t light)
{
    if (!h || !ray || !light) return -1;
    h->c.d.max_ray_samples = ray; h->c.d.max_light_samples = light;
    return 0;
}
int mvo_set_volumes_world(mvo_caster* h, float size, const float center[3]) { if (!h) return -1; set_volumes_world(h->c, size, center); return 0; }
int mvo_set_volume_world(mvo_caster* h, uint32_t i, float size, const float pos[3])
{
    if (!h || !ok_vol(h->c, i)) return -1;
    set_volume_world(h->c, i, size, pos);
    return 0;
}
int mvo_set_volume_world_matrix(mvo_caster* h, uint32_t i, const float w[12])
{
    if (!h || !ok_vol(h->c, i)) return -1;
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 3; ++k) h->c.volumeWorlds[i].m[r][k] = w[r * 3 + k];
    return 0;
}
int mvo_set_light(mvo_caster* h, const float p[3], const float col[3], float intensity)
{
    if (!h) return -1;
    h->c.lightPt = {p[0], p[1], p[2]}; h->c.lightColor = {col[0], col[1], col[2], intensity};
    return 0;
}
// ---- occluder mesh: ObjectRenderer.cpp:68-77 (AABB), :147-153 (SetWorld), :171-190 (light view-projection) ----
int mvo_mesh_set(mvo_caster* h, const float* pos, uint32_t nv, const uint32_t* idx, uint32_t ni)
{
    if (!h || ni % 3 != 0 || (ni && (!pos || !idx))) return -1;
    for (uint32_t i = 0; i < ni; ++i) if (idx[i] >= nv) return -1;
    Caster& c = h->c;
    c.meshPos.assign(pos, pos + (size_t)nv * 3);
    c.meshIdx.assign(idx, idx + ni);
    recompute_normals(c.meshPos, c.meshIdx, c.meshNrm);
    c.meshHavePrev = false;
    c.meshExtent = 1.0f;
    if (ni) {
        float mn[3] = {kFltMax, kFltMax, kFltMax}, mx[3] = {-kFltMax, -kFltMax, -kFltMax};
        for (uint32_t v = 0; v < nv; ++v)
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], pos[3 * v + k]); mx[k] = fmaxf(mx[k], pos[3 * v + k]); }
        c.meshExtent = fmaxf(mx[0] - mn[0], fmaxf(mx[1] - mn[1], mx[2] - mn[2]));
    }
    return 0;
}
int mvo_mesh_set_world(mvo_caster* h, float scale, const float pos[3])
{
    if (!h || !pos) return -1;
    h->c.meshScale = scale; h->c.meshPosition = {pos[0], pos[1], pos[2]};
    return 0;
}
namespace {
// DirectXMath call sites of ObjectRenderer.cpp:182-186, evaluated in double and rounded once to fp32
void mul44d(const double* A, const double* B, double* R)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            R[i * 4 + j] = s;
        }
}
void look_at_lh(const double e[3], double M[16])   // XMMatrixLookAtLH(eye, 0, (0, 1, 0))
{
    const double up[3] = {0.0, 1.0, 0.0};
    double z[3] = {0.0 - e[0], 0.0 - e[1], 0.0 - e[2]};
    double l = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (double& v : z) v /= l;
    double x[3] = {up[1] * z[2] - up[2] * z[1], up[2] * z[0] - up[0] * z[2], up[0] * z[1] - up[1] * z[0]};
    l = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (double& v : x) v /= l;
    const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    auto d3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const double m[16] = {x[0], y[0], z[0], 0, x[1], y[1], z[1], 0, x[2], y[2], z[2], 0, -d3(x, e), -d3(y, e), -d3(z, e), 1};
    memcpy(M, m, sizeof m);
}
}
static int mesh_render_impl(mvo_caster* h, const float viewProj[16], float shadowVpOut[16], bool basePass, const float* eye, const float* clear)
{
    if (!h || !viewProj || (basePass && !eye)) return -1;
    Caster& c = h->c;
    const uint32_t S = 1024;   // m_shadowMapSize
    const double s = c.meshScale;
    const double world[16] = {s, 0, 0, 0, 0, s, 0, 0, 0, 0, s, 0, c.meshPosition.x, c.meshPosition.y, c.meshPosition.z, 1};
    double vp[16], wvp[16], lv[16], lvp[16], swvp[16];
    for (int i = 0; i < 16; ++i) vp[i] = viewProj[i];
    mul44d(world, vp, wvp);
    const double size = (double)(c.meshExtent * c.meshScale) * 1.5, zn = 1.0, zf = 200.0;
    const double lightEye[3] = {c.lightPt.x, c.lightPt.y, c.lightPt.z};
    look_at_lh(lightEye, lv);
    const double lp[16] = {2.0 / size, 0, 0, 0, 0, 2.0 / size, 0, 0, 0, 0, 1.0 / (zf - zn), 0, 0, 0, -zn / (zf - zn), 1};
    mul44d(lv, lp, lvp);
    mul44d(world, lvp, swvp);
    m44 W, SW;
    for (int i = 0; i < 16; ++i) { W.m[i / 4][i % 4] = (float)wvp[i]; SW.m[i / 4][i % 4] = (float)swvp[i]; if (shadowVpOut) shadowVpOut[i] = (float)lvp[i]; }
    std::vector<float> sd((size_t)S * S);
    raster_depth(c.meshPos, c.meshIdx, SW, S, S, sd.data());
    c.shadow.resize((size_t)S * S); c.shadowSize = S;
    for (size_t i = 0; i < sd.size(); ++i) c.shadow[i] = (uint16_t)floorf(sd[i] * 65535.0f + 0.5f);
    if (!basePass) {
        c.depth.resize((size_t)c.d.width * c.d.height);
        raster_depth(c.meshPos, c.meshIdx, W, c.d.width, c.d.height, c.depth.data());
        return 0;
    }
    m43 world43;
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 3; ++k) world43.m[r][k] = (float)world[r * 4 + k];
    const m44 prev = c.meshHavePrev ? c.meshWvpPrev : W;
    c.meshWvpPrev = W; c.meshHavePrev = true;
    const float zero[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    render_base_pass(c, W, prev, world43, SW, f3{eye[0], eye[1], eye[2]}, clear ? clear : zero);
    return 0;
}
int mvo_mesh_render_depth(mvo_caster* h, const float viewProj[16], float shadowVpOut[16]) { return mesh_render_impl(h, viewProj, shadowVpOut, false, nullptr, nullptr); }
int mvo_mesh_render(mvo_caster* h, const float viewProj[16], const float eye[3], const float clearRgba[4], float shadowVpOut[16])
{
    return mesh_render_impl(h, viewProj, shadowVpOut, true, eye, clearRgba);
}
int mvo_read_velocity(mvo_caster* h, uint16_t* out)
{
    if (!h || !out) return -1;
    Caster& c = h->c;
    const size_t n = (size_t)c.d.width * c.d.height * 2;
    if (c.velocity.size() != n) memset(out, 0, n * sizeof(uint16_t));
    else memcpy(out, c.velocity.data(), n * sizeof(uint16_t));
    return 0;
}
int mvo_read_depth(mvo_caster* h, float* depth, uint16_t* shadow, uint32_t* shadowSize)
{
    if (!h) return -1;
    Caster& c = h->c;
    if (depth) memcpy(depth, c.depth.data(), c.depth.size() * sizeof(float));
    if (shadow && c.shadowSize) memcpy(shadow, c.shadow.data(), c.shadow.size() * sizeof(uint16_t));
    if (shadowSize) *shadowSize = c.shadowSize;
    return 0;
}
int mvo_set_ambient(mvo_caster* h, const float col[3], float intensity)
{
    if (!h) return -1;
    h->c.ambient = {col[0], col[1], col[2], intensity};
    return 0;
}

int mvo_update_frame(mvo_caster* h, const float viewProj[16], const float shadowVP[16], const float eye[3])   // MultiRayCaster.cpp:316-353
{
    if (!h || !viewProj || !eye) return -1;
    Caster& c = h->c;
    m44 vp, svp;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { vp.m[i][j] = viewProj[i * 4 + j]; svp.m[i][j] = shadowVP ? shadowVP[i * 4 + j] : (i == j ? 1.0f : 0.0f); }
    c.cb.eyePt = {eye[0], eye[1], eye[2]};
    c.cb.viewport = {(float)c.d.width, (float)c.d.height};
    c.cb.screenToWorld = inverse44(vp);
    c.cb.shadowViewProj = svp;
    c.cb.lightPos = {c.lightPt.x, c.lightPt.y, c.lightPt.z, 1.0f};
    c.cb.lightColor = c.lightColor;
    c.cb.ambient = c.ambient;
    c.cb.frameIdx = c.frameIdx;
    for (uint32_t i = 0; i < c.d.num_volumes; ++i) {
        const m44 world = from43(c.volumeWorlds[i]);
        const m44 worldI = inverse44(world);
        const m44 wvp = mul44(world, vp);
        PerObject& po = c.perObject[i];
        po.WorldViewProj = wvp;
        po.WorldViewProjI = inverse44(wvp);
        po.WorldI = to43(worldI);
        po.World = to43(world);
    }
    return 0;
}

int mvo_cull(mvo_caster* h) { if (!h) return -1; cull_volumes(h->c); return 0; }
int mvo_ray_march_light(mvo_caster* h, int32_t v)
{
    if (!h || v >= (int32_t)h->c.d.num_volumes) return -1;
    ray_march_light(h->c, v);
    return 0;
}
int mvo_ray_march_view(mvo_caster* h) { if (!h) return -1; ray_march_view(h->c); return 0; }
int mvo_resolve_oit(mvo_caster* h) { if (!h) return -1; resolve_oit(h->c); return 0; }

int mvo_render(mvo_caster* h, uint32_t oit)   // MultiRayCaster.cpp:355-385
{
    if (!h) return -1;
    (void)oit;   // only the K-buffer semantics (default branch, :377-381) are restated
    Caster& c = h->c;
    cull_volumes(c);
    ray_march_light(c, -1);
    ray_march_view(c);
    resolve_oit(c);
    if (c.frameIdx != 0xffffffffu) ++c.frameIdx;
    return 0;
}
int mvo_render_work_graph(mvo_caster* h, uint32_t oit)   // MultiRayCaster.cpp:358-362 (useWorkGraph), rayMarchWG :1370-1438
{
    if (!h) return -1;
    (void)oit;
    Caster& c = h->c;
    // rayMarchL first: CSRayMarchL.hlsl:29-33 reads the visible list and its counter as the previous frame's graph left
    // them (c.visible is only rewritten by cull_volumes); then the graph: VolumeCull node -> RayMarch node
    ray_march_light(c, -1);
    cull_volumes(c);
    ray_march_view(c);
    resolve_oit(c);
    if (c.frameIdx != 0xffffffffu) ++c.frameIdx;
    return 0;
}
int mvo_postprocess(mvo_caster* h, uint32_t taa) { if (!h) return -1; temporal_aa(h->c, taa != 0); tone_map(h->c); return 0; }
int mvo_sh_project(mvo_caster* h, const float* cube, uint32_t size, float* out27)
{
    (void)h;
    if (!cube || !out27 || !size) return -1;
    sh_project(cube, size, out27);
    return 0;
}

int mvo_read_per_object(mvo_caster* h, float* out)
{
    if (!h || !out) return -1;
    for (const PerObject& po : h->c.perObject) {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) *out++ = po.WorldViewProj.m[i][j];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) *out++ = po.WorldViewProjI.m[i][j];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) *out++ = po.WorldI.m[i][j];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) *out++ = po.World.m[i][j];
    }
    return 0;
}
int mvo_read_visible(mvo_caster* h, uint32_t* ids, uint32_t* count)
{
    if (!h || !count) return -1;
    *count = (uint32_t)h->c.visible.size();
    if (ids) std::copy(h->c.visible.begin(), h->c.visible.end(), ids);
    return 0;
}
int mvo_read_cube_volumes(mvo_caster* h, uint32_t* ids, uint32_t* count)
{
    if (!h || !count) return -1;
    *count = (uint32_t)h->c.cubeVolumes.size();
    if (ids) std::copy(h->c.cubeVolumes.begin(), h->c.cubeVolumes.end(), ids);
    return 0;
}
int mvo_read_attribs(mvo_caster* h, uint16_t* out) { if (!h || !out) return -1; std::copy(h->c.attribs.begin(), h->c.attribs.end(), out); return 0; }
int mvo_read_cubemap(mvo_caster* h, uint32_t v, uint32_t mip, uint16_t* rgba, float* depth)
{
    if (!h || !ok_vol(h->c, v) || mip >= kNumCubeMip) return -1;
    const CubeMap& cm = h->c.cubeMaps[v];
    if (rgba) std::copy(cm.color[mip].begin(), cm.color[mip].end(), rgba);
    if (depth) std::copy(cm.depth[mip].begin(), cm.depth[mip].end(), depth);
    return 0;
}
int mvo_read_lightmap(mvo_caster* h, uint32_t v, uint16_t* out)
{
    if (!h || !out || !ok_vol(h->c, v)) return -1;
    const auto& t = h->c.lightMaps[v].texels;
    std::copy(t.begin(), t.end(), out);
    return 0;
}
int mvo_read_frame(mvo_caster* h, uint16_t* out) { if (!h || !out) return -1; std::copy(h->c.color.begin(), h->c.color.end(), out); return 0; }
int mvo_read_post(mvo_caster* h, uint16_t* taa, uint8_t* rgba8)
{
    if (!h) return -1;
    const auto& t = h->c.taaHistory[h->c.frameParity];
    if (taa) std::copy(t.begin(), t.end(), taa);
    if (rgba8) std::copy(h->c.backBuffer.begin(), h->c.backBuffer.end(), rgba8);
    return 0;
}
int mvo_get_stats(mvo_caster* h, mvo_stats* out) { if (!h || !out) return -1; *out = h->c.stats; return 0; }
int mvo_set_frame_index(mvo_caster* h, uint32_t f) { if (!h) return -1; h->c.frameIdx = f; h->c.cb.frameIdx = f; return 0; }

int mvo_set_shard(mvo_caster* h, uint32_t rank, uint32_t world)
{
    if (!h || world == 0 || rank >= world) return -1;
    h->c.shardRank = rank; h->c.shardWorld = world;
    return 0;
}
// Density proxy of a source volume: the mean density of every (G / P)^3 block (fp32, summed x fastest then y then z, times
// the reciprocal of the block size — a power of two), rounded to binary16. Same arithmetic as k_build_proxy of the product.
static void build_proxy(const Tex3D& full, uint32_t P, Tex3D& out)
{
    const uint32_t G = full.n, f = G / P;
    out.n = P;
    out.texels.assign((size_t)P * P * P * 4, f32_to_f16(1.0f));
    const float inv = 1.0f / (float)(f * f * f);
    for (uint32_t z = 0; z < P; ++z)
        for (uint32_t y = 0; y < P; ++y)
            for (uint32_t x = 0; x < P; ++x) {
                float sum = 0.0f;
                for (uint32_t k = 0; k < f; ++k)
                    for (uint32_t j = 0; j < f; ++j)
                        for (uint32_t i = 0; i < f; ++i) sum += f16_to_f32(full.at((int)(x * f + i), (int)(y * f + j), (int)(z * f + k))[3]);
                out.texels[(((size_t)z * P + y) * P + x) * 4 + 3] = f32_to_f16(sum * inv);
            }
}

int mvo_set_shard_volumes(mvo_caster* h, uint32_t rank, uint32_t world, uint32_t proxyGrid)
{
    if (!h || world == 0 || rank >= world || proxyGrid == 0 || h->c.d.grid_size % proxyGrid != 0) return -1;
    Caster& c = h->c;
    c.shardRank = rank; c.shardWorld = world; c.shardVolumes = true; c.proxyGrid = proxyGrid;
    c.proxies.assign(c.volumes.size(), Tex3D());
    for (uint32_t s = 0; s < c.volumes.size(); ++s)
        if (!c.owns_source(s)) build_proxy(c.volumes[s], proxyGrid, c.proxies[s]);
    return 0;
}

int mvo_set_environment(mvo_caster* h, const float* cubeRGB, uint32_t size)
{
    if (!h || ((cubeRGB == nullptr) != (size == 0))) return -1;
    Caster& c = h->c;
    c.envSize = size;
    c.envCube.assign((size_t)6 * size * size * 4, 0);
    for (size_t i = 0; i < (size_t)6 * size * size; ++i)
        for (int k = 0; k < 3; ++k) c.envCube[4 * i + k] = f32_to_f16(cubeRGB[3 * i + k]);
    return 0;
}

int mvo_render_environment(mvo_caster* h) { if (!h) return -1; render_environment(h->c); return 0; }

// process-wide: the `min16float` literals as the shipped DXIL holds them (binary16-rounded), SURVEY.md App. B.2
void mvo_set_min16_consts_as_half(int on)
{
    Min16Consts k;                               // on (the default): the shipped DXIL's binary16 literals
    if (!on) {                                   // off: the decimal literals of the HLSL text, as fp32
        k.absorption = 0.8f;
        k.zeroThreshold = 0.01f;
        k.maxDist = 3.4641016151377544f;         // 2 sqrt(3)
        k.invTwoPi = 0.0f;                       // divide by 2 pi
        k.alphaClamp = 0.9997f;
        k.ninth = 1.0f / 9.0f;
        k.toneScale = 1.05f; k.toneBias = 0.7f;
    }
    g_min16 = k;
}

int mvo_set_row_band(mvo_caster* h, uint32_t row0, uint32_t row1)
{
    if (!h || row0 > row1 || row1 > h->c.d.height) return -1;
    h->c.row0 = row0; h->c.row1 = row1;
    return 0;
}
int mvo_write_cubemap(mvo_caster* h, uint32_t v, uint32_t mip, const uint16_t* rgba, const float* depth)
{
    if (!h || !ok_vol(h->c, v) || mip >= kNumCubeMip) return -1;
    CubeMap& cm = h->c.cubeMaps[v];
    if (rgba) std::copy(rgba, rgba + cm.color[mip].size(), cm.color[mip].begin());
    if (depth) std::copy(depth, depth + cm.depth[mip].size(), cm.depth[mip].begin());
    return 0;
}
int mvo_write_lightmap_slab(mvo_caster* h, uint32_t v, uint32_t z0, uint32_t z1, const uint16_t* slab)
{
    if (!h || !slab || !ok_vol(h->c, v) || z0 > z1 || z1 > h->c.d.light_grid_size) return -1;
    const size_t L = h->c.d.light_grid_size;
    std::copy(slab, slab + (size_t)(z1 - z0) * L * L * 4, h->c.lightMaps[v].texels.begin() + (size_t)z0 * L * L * 4);
    return 0;
}
int mvo_write_rows(mvo_caster* h, uint32_t what, uint32_t row0, uint32_t row1, const void* rows)
{
    if (!h || !rows || row0 > row1 || row1 > h->c.d.height) return -1;
    Caster& c = h->c;
    const size_t W = c.d.width, n = (size_t)(row1 - row0) * W;
    if (what == 0) std::copy((const uint16_t*)rows, (const uint16_t*)rows + n * 4, c.color.begin() + (size_t)row0 * W * 4);
    else if (what == 1) std::copy((const uint16_t*)rows, (const uint16_t*)rows + n * 4, c.taaHistory[c.frameParity].begin() + (size_t)row0 * W * 4);
    else if (what == 2) std::copy((const uint8_t*)rows, (const uint8_t*)rows + n * 4, c.backBuffer.begin() + (size_t)row0 * W * 4);
    else return -1;
    return 0;
}

void mvo_sample_volume(mvo_caster* h, uint32_t src, const float uvw[3], float out[4])
{
    const f4 r = sample3d(h->c.volumes[src], {uvw[0], uvw[1], uvw[2]}, h->c.filterModel);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
/* oracle/dxil: resolve the frame once more with a per-fragment record (the colour target is restored afterwards) */
int mvo_debug_oit(mvo_caster* h, uint32_t* count, uint32_t* info, float* data, float* result, uint32_t* allKeys)
{
    if (!h) return -1;
    Caster& c = h->c;
    const std::vector<uint16_t> keep = c.color;
    const mvo_stats st = c.stats;
    c.debugOIT = true;
    resolve_oit(c);
    c.debugOIT = false;
    c.color = keep; c.stats = st;
    if (count) std::copy(c.dbgCount.begin(), c.dbgCount.end(), count);
    if (info) std::copy(c.dbgInfo.begin(), c.dbgInfo.end(), info);
    if (data) std::copy(c.dbgData.begin(), c.dbgData.end(), data);
    if (result) std::copy(c.dbgResult.begin(), c.dbgResult.end(), result);
    if (allKeys) std::copy(c.dbgAllKeys.begin(), c.dbgAllKeys.end(), allKeys);
    return 0;
}
/* oracle/dxil: keep fp32 copies of what the marches computed (what = 0 off, 1 on); read them back with out != NULL */
int mvo_debug_f32(mvo_caster* h, int on, float* cubeOut, float* lightOut)
{
    if (!h) return -1;
    Caster& c = h->c;
    const size_t G = c.d.grid_size, L = c.d.light_grid_size;
    if (on && !c.debugF32) { c.dbgCubeF32.assign((size_t)c.d.num_volumes * 6 * G * G * 4, 0.0f); c.dbgLightF32.assign(L * L * L * 3, 0.0f); }
    c.debugF32 = on != 0;
    if (cubeOut && !c.dbgCubeF32.empty()) std::copy(c.dbgCubeF32.begin(), c.dbgCubeF32.end(), cubeOut);
    if (lightOut && !c.dbgLightF32.empty()) std::copy(c.dbgLightF32.begin(), c.dbgLightF32.end(), lightOut);
    return 0;
}
/* oracle/dxil: cbPerFrame as the shaders see it (row-vector matrices, un-transposed): eye 3, viewport 2, screenToWorld 16, shadowViewProj 16 */
void mvo_read_per_frame(mvo_caster* h, float* out37)
{
    const PerFrame& f = h->c.cb;
    out37[0] = f.eyePt.x; out37[1] = f.eyePt.y; out37[2] = f.eyePt.z; out37[3] = f.viewport.x; out37[4] = f.viewport.y;
    memcpy(out37 + 5, f.screenToWorld.m, 64); memcpy(out37 + 21, f.shadowViewProj.m, 64);
}
void mvo_sample_lightmap(mvo_caster* h, uint32_t volume, const float uvw[3], float out[4])
{
    const f4 r = sample3d(h->c.lightMaps[volume], {uvw[0], uvw[1], uvw[2]}, h->c.filterModel);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
float mvo_quantize_r11(float v) { return quantize_ufloat(v, 6); }
float mvo_quantize_b10(float v) { return quantize_ufloat(v, 5); }
uint16_t mvo_f32_to_f16(float v) { return f32_to_f16(v); }
float mvo_f16_to_f32(uint16_t hbits) { return f16_to_f32(hbits); }
void mvo_eval_sh_irradiance(const float* k, const float n[3], float out4[4])
{
    f3 sh[9];
    for (int i = 0; i < 9; ++i) sh[i] = {k[i * 3], k[i * 3 + 1], k[i * 3 + 2]};
    const f4 r = evaluate_sh_irradiance(sh, {n[0], n[1], n[2]});
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

} // extern "C"
