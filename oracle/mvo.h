/* mvo.h — ORACLE C API (test infrastructure, NOT product code).
 *
 * CPU restatement (scalar C++17 + OpenMP) of the reference's cube-map-space volume rendering path:
 * CSVolumeCull, CSRayMarchL, CSRayMarchV, RayCast, CubeCast, depth-peel / PSResolveOIT, CSTemporalAA,
 * PSToneMap, SH evaluation / projection and CSInitGridData (reference: MultiVolumes/Content/Shaders).
 * It mirrors the product C-ABI (include/mv.h) call for call so the parity tests drive both the same
 * way. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libmv_b200.so) never does.
 *
 * Parity pinning status: PINNED against outputs of the reference itself. The reference ships no
 * golden vectors or CPU path, but it ships its shaders compiled (Bin/ *.cso, DXIL); oracle/dxil
 * disassembles and executes them here, and tests/test_dxil_golden.py holds this library to their
 * outputs (cull exact; light march, view march, CubeCast / RayCast / resolve, volume init, base pass
 * bit-exact; TAA within one binary16 step; SH projection — DXIL-only in the reference — to 4e-7).
 * Unpinned remain the D3D12 rasteriser and the texture unit of the GPU the reference runs on
 * (DESIGN.md section 2). Further pins: analytic known-answer tests (tests/test_oracle_kat.py).
 */
#ifndef MVO_H
#define MVO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvo_caster mvo_caster;

typedef struct mvo_desc {
    uint32_t grid_size;         /* G: volume and cube-map edge (reference default 128) */
    uint32_t light_grid_size;   /* L: light-map edge (default 96) */
    uint32_t num_volumes;       /* N instances */
    uint32_t num_volume_srcs;   /* distinct density textures; VolTexId = i % srcs */
    uint32_t width, height;     /* viewport */
    uint32_t max_ray_samples;   /* default 256 */
    uint32_t max_light_samples; /* default 96 */
    uint32_t tex_filter_model;  /* 0 = exact fp32 trilinear weights, 1 = sm_100a texture-unit model */
    uint32_t num_threads;       /* 0 = OpenMP default */
} mvo_desc;

typedef struct mvo_stats {
    uint64_t view_rays, view_samples, view_light_fetches;
    uint64_t light_voxels, light_dense_voxels, light_samples;
    uint64_t direct_rays, direct_samples, direct_light_fetches;
    uint64_t oit_fragments;
    uint32_t visible_count, cubemap_count;
    uint32_t light_volume;      /* volume whose light map the last render filled */
    uint32_t threads;
    uint64_t view_skipped, direct_skipped;   /* layout of mv_stats; the oracle fetches every sample: always 0 */
} mvo_stats;

int  mvo_create(const mvo_desc* desc, mvo_caster** out);
void mvo_destroy(mvo_caster* c);

/* MultiRayCaster::InitVolumeData / LoadVolumeData */
int mvo_volume_init_procedural(mvo_caster* c, uint32_t src, uint32_t mode, uint32_t seed);
int mvo_volume_upload_rgba16f(mvo_caster* c, uint32_t src, const uint16_t* texels);
int mvo_volume_upload_r32f(mvo_caster* c, uint32_t src, const float* density);
int mvo_volume_upload_r32f_sized(mvo_caster* c, uint32_t src, const float* density, uint32_t width, uint32_t height, uint32_t depth);
int mvo_volume_read(mvo_caster* c, uint32_t src, uint16_t* texels_out);

/* MultiRayCaster::SetRenderTargets / SetViewport: borrowed scene depth, shadow map, colour RT */
int mvo_set_targets(mvo_caster* c, const float* depth, const uint16_t* shadow_d16, uint32_t shadow_size,
                    const uint16_t* color_rgba16f, const uint16_t* velocity_rg16f);
int mvo_reset_color(mvo_caster* c);
int mvo_set_sh(mvo_caster* c, const float* coeffs27);
int mvo_set_max_samples(mvo_caster* c, uint32_t ray, uint32_t light);
int mvo_set_volumes_world(mvo_caster* c, float size, const float center[3]);
int mvo_set_volume_world(mvo_caster* c, uint32_t i, float size, const float pos[3]);
int mvo_set_volume_world_matrix(mvo_caster* c, uint32_t i, const float world43[12]);
int mvo_set_light(mvo_caster* c, const float pos[3], const float color[3], float intensity);
int mvo_set_ambient(mvo_caster* c, const float color[3], float intensity);
int mvo_update_frame(mvo_caster* c, const float view_proj[16], const float shadow_vp[16], const float eye[3]);

/* passes; mvo_render = cull -> light march (one volume, round-robin) -> view march -> OIT, frame++ */
int mvo_render(mvo_caster* c, uint32_t oit_method);
/* Render(..., useWorkGraph = true), MultiRayCaster.cpp:358-362: light march (volume from the previous frame's visible list) -> cull -> view march -> OIT */
int mvo_render_work_graph(mvo_caster* c, uint32_t oit_method);
int mvo_cull(mvo_caster* c);
int mvo_ray_march_light(mvo_caster* c, int32_t volume_override);   /* -1: reference round-robin */
int mvo_ray_march_view(mvo_caster* c);
int mvo_resolve_oit(mvo_caster* c);
int mvo_postprocess(mvo_caster* c, uint32_t taa_on);
/* LightProbe: radiance cube map + RenderEnvironment (same semantics as mv_set_environment / mv_render_environment) */
int mvo_set_environment(mvo_caster* c, const float* cube_rgb_f32, uint32_t size);
int mvo_render_environment(mvo_caster* c);
int mvo_sh_project(mvo_caster* c, const float* cube_rgb_f32, uint32_t size, float* coeffs27_out);

/* ObjectRenderer's depth-only passes (same semantics as mv_mesh_* of the product) */
int mvo_mesh_set(mvo_caster* c, const float* positions_xyz, uint32_t num_vertices, const uint32_t* indices, uint32_t num_indices);
int mvo_mesh_set_world(mvo_caster* c, float scale, const float pos[3]);
int mvo_mesh_render_depth(mvo_caster* c, const float view_proj[16], float shadow_vp_out[16]);
int mvo_mesh_render(mvo_caster* c, const float view_proj[16], const float eye[3], const float clear_rgba[4], float shadow_vp_out[16]);
int mvo_read_velocity(mvo_caster* c, uint16_t* rg16f);
int mvo_read_depth(mvo_caster* c, float* depth, uint16_t* shadow_d16, uint32_t* shadow_size);

/* read-backs */
int mvo_read_per_object(mvo_caster* c, float* out56xN);
int mvo_read_visible(mvo_caster* c, uint32_t* ids, uint32_t* count);
int mvo_read_cube_volumes(mvo_caster* c, uint32_t* ids, uint32_t* count);
int mvo_read_attribs(mvo_caster* c, uint16_t* out4xN);
int mvo_read_cubemap(mvo_caster* c, uint32_t volume, uint32_t mip, uint16_t* rgba16f, float* depth);
int mvo_read_lightmap(mvo_caster* c, uint32_t volume, uint16_t* rgba16f);
int mvo_read_frame(mvo_caster* c, uint16_t* rgba16f);
int mvo_read_post(mvo_caster* c, uint16_t* taa_rgba16f, uint8_t* rgba8);
int mvo_get_stats(mvo_caster* c, mvo_stats* out);
int mvo_set_frame_index(mvo_caster* c, uint32_t frame_idx);
/* sharding, same semantics as mv_set_shard / mv_set_row_band */
int mvo_set_shard(mvo_caster* c, uint32_t rank, uint32_t world);
int mvo_set_row_band(mvo_caster* c, uint32_t row0, uint32_t row1);
/* volume-sharded storage, same semantics as mv_create_sharded: call after the volumes are loaded (builds the proxies) */
int mvo_set_shard_volumes(mvo_caster* c, uint32_t rank, uint32_t world, uint32_t proxy_grid);
/* raw access for the exchange steps of the multi-rank host logic (tests): cube map / light map of a volume */
int mvo_write_cubemap(mvo_caster* c, uint32_t volume, uint32_t mip, const uint16_t* rgba16f, const float* depth);
int mvo_write_lightmap_slab(mvo_caster* c, uint32_t volume, uint32_t z0, uint32_t z1, const uint16_t* rgba16f_slab);
int mvo_write_rows(mvo_caster* c, uint32_t what, uint32_t row0, uint32_t row1, const void* rows);

/* MV_MIN16_CONSTS_AS_HALF: 1 (the default since the DXIL was disassembled and executed, oracle/dxil) =
 * the binary16 `min16float` literals the shipped shaders hold (g_maxDist 3.4648, ABSORPTION 0.7998, ZERO_THRESHOLD 0.010002,
 * 1/(2 pi) 0.15918, alpha clamp 0.99951, 1/9 0.11108, tone map 1.0498 / 0.70020); 0 = the decimal literals of the HLSL text
 * as fp32. Process-wide. For reporting how far the two readings sit apart (profiles/r02_min16_delta.json). */
void mvo_set_min16_consts_as_half(int on);

/* stand-alone helpers used by the known-answer tests */
void  mvo_cube_resolve_texel(int size, int face, int i, int j, int out_face_i_j[3]);   /* seamless Gather addressing of CubeCast */
int   mvo_debug_oit(mvo_caster* c, uint32_t* count_wh, uint32_t* info_wh8x4, float* data_wh8x9, float* result_wh4, uint32_t* all_keys_whn);   /* per-fragment record of resolve_oit */
int   mvo_debug_f32(mvo_caster* c, int on, float* cube_n6ggx4, float* light_lllx3);   /* fp32 outputs of the marches before their format conversion */
void  mvo_read_per_frame(mvo_caster* c, float* out37);   /* eye 3, viewport 2, screenToWorld 16, shadowViewProj 16 */
void  mvo_sample_volume(mvo_caster* c, uint32_t src, const float uvw[3], float rgba_out[4]);
void  mvo_sample_lightmap(mvo_caster* c, uint32_t volume, const float uvw[3], float rgba_out[4]);   /* the texture filter of the caster's model */
float mvo_quantize_r11(float v);
float mvo_quantize_b10(float v);
uint16_t mvo_f32_to_f16(float v);
float mvo_f16_to_f32(uint16_t h);
void  mvo_eval_sh_irradiance(const float* coeffs27, const float normal[3], float out4[4]);

#ifdef __cplusplus
}
#endif
#endif
