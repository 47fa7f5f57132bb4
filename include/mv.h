/* mv.h — C-ABI of the B200-native cube-map-space multi-volume renderer (libmv_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of StarsX/MultiVolumes: the public section of
 * `class MultiRayCaster` (reference: MultiVolumes/Content/MultiRayCaster.h:28-50) and the post-process
 * entry of ObjectRenderer (MultiVolumes/Content/ObjectRenderer.h:46-48). Each export below cites the
 * reference interface it replaces. The XUSG/D3D12 plumbing (command lists, descriptor tables,
 * barriers, ExecuteIndirect) collapses into one CUDA stream owned by the handle.
 *
 * Conventions
 *   - every call returns 0 on success, a negative mv_status otherwise; mv_last_error() gives text;
 *   - one handle = one owner thread = one CUDA device + stream; calls are asynchronous on that
 *     stream until a mv_read_* or mv_sync;
 *   - matrices are row-major float arrays in the reference's row-vector convention
 *     (v' = v * M, DirectXMath layout before the reference's XMMatrixTranspose for upload);
 *   - images are tightly packed, row-major, top row first; RGBA16F = 4 x IEEE binary16;
 *   - volumes are RGBA16F (R16F density with MV_FLAG_DENSITY_ONLY), x fastest then y then z; cube maps are [face][y][x] with the D3D face
 *     order +X,-X,+Y,-Y,+Z,-Z;
 *   - there is NO CPU fallback: mv_create fails with MV_ERR_NO_DEVICE when no sm_100 device exists;
 *   - numerics: the shaders' arithmetic in fp32, in a stated evaluation order, with the `min16float`
 *     literals the reference's SHIPPED shaders hold (Bin/ *.cso: dxc folds them to binary16 — g_maxDist 3.46484375,
 *     ABSORPTION 0.7998046875, ZERO_THRESHOLD 0.010002136 ...), not the decimal text of the HLSL: results reproduce those
 *     compiled shaders (cull and view march exactly; light march, CubeCast / RayCast and resolve in every stored value of
 *     the test vectors, one format step off at a rate below 2e-4 in a random sweep; tests/test_dxil_golden.py).
 */
#ifndef MV_H
#define MV_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mv_caster mv_caster;

enum mv_status {
    MV_OK = 0,
    MV_ERR_INVALID = -1,     /* bad argument */
    MV_ERR_NOMEM = -2,       /* host or device allocation failed */
    MV_ERR_CUDA = -3,        /* a CUDA call failed; see mv_last_error */
    MV_ERR_NO_DEVICE = -4,   /* no usable CUDA device (the product has no CPU path) */
    MV_ERR_PEER_TIMEOUT = -5 /* multi-GPU: a peer did not reach a device-side barrier within ~2 s; the frames rendered since the
                                last successful mv_sync / mv_present_wait / mv_peer_barrier are not valid */
};

/* mv_desc.flags */
#define MV_FLAG_COUNT_SAMPLES   1u   /* keep exact ray/sample counters (mv_get_stats); small cost */
#define MV_FLAG_TIME_PASSES     2u   /* bracket every pass with CUDA events (mv_get_timings) */
/* creation-time only: the volumes are stored as R16F 3-D textures holding the density alone, 2 B / voxel instead of 8.
 * The colour of such a volume is (1, 1, 1) — what the reference's own file ingest produces for every asset it ships
 * (CSR32FToRGBA16F.hlsl:26: rgb = 1, a = 0.25 src) — so a frame equals, bit for bit, the frame of the RGBA16F storage
 * holding (1, 1, 1, a). Uploads keep the alpha channel only; mv_volume_read expands to (1, 1, 1, a).
 * SURVEY.md 8(d) cfg 5: 512 x 512^3 = 137 GB instead of 550 GB. */
#define MV_FLAG_DENSITY_ONLY    4u

/* set by mv_create_sharded (do not pass it to mv_create): volume-sharded storage, see the multi-GPU section */
#define MV_FLAG_SHARD_VOLUMES   8u

/* MultiRayCaster::Init arguments (MultiRayCaster.h:31-34) + viewport (SetViewport, :38) +
 * SetMaxSamples defaults (MultiVolumes.cpp:27-68). */
typedef struct mv_desc {
    uint32_t grid_size;         /* G: volume and cube-map edge (reference default 128) */
    uint32_t light_grid_size;   /* L: light-map edge (default 96) */
    uint32_t num_volumes;       /* N instances */
    uint32_t num_volume_srcs;   /* distinct density textures; VolTexId = i % srcs (MultiRayCaster.cpp:476) */
    uint32_t width, height;     /* viewport */
    uint32_t max_ray_samples;   /* default 256 */
    uint32_t max_light_samples; /* default 96 */
    uint32_t device;            /* CUDA device ordinal */
    uint32_t flags;             /* MV_FLAG_* */
} mv_desc;

typedef struct mv_stats {
    uint64_t view_rays, view_samples, view_light_fetches;        /* CSRayMarchV */
    uint64_t light_voxels, light_dense_voxels, light_samples;    /* CSRayMarchL */
    uint64_t direct_rays, direct_samples, direct_light_fetches;  /* RayCast inside the OIT pass */
    uint64_t oit_fragments;
    uint32_t visible_count, cubemap_count;
    uint32_t light_volume;      /* volume whose light map the last render filled */
    uint32_t threads;           /* CUDA: SM count of the device */
    /* of view_samples / direct_samples: samples that fell into a brick known to be empty and needed no texture fetch
     * (csrc/mv_internal.h, Occupancy); they are samples of the algorithm all the same, so the counters above include them */
    uint64_t view_skipped, direct_skipped;
} mv_stats;

/* per-pass device time of the last frame, milliseconds (CUDA events on the caster's stream) */
typedef struct mv_timings {
    float cull, ray_march_light, ray_march_view, resolve_oit, postprocess, total;
} mv_timings;

const char* mv_last_error(void);
uint32_t mv_abi_version(void);

/* Init (MultiRayCaster.h:31-34, MultiRayCaster.cpp:79-166) / destructor */
int  mv_create(const mv_desc* desc, mv_caster** out);
void mv_destroy(mv_caster* c);

/* InitVolumeData (MultiRayCaster.h:40, CSInitGridData.hlsl:10-27); mode 1 adds seeded value noise.
 * LoadVolumeData (MultiRayCaster.h:35-36): RGBA16F texels, or R32F density through the
 * CSR32FToRGBA16F conversion (rgb = 1, a = 0.25 * src). Host pointers. */
int mv_volume_init_procedural(mv_caster* c, uint32_t src, uint32_t mode, uint32_t seed);
int mv_volume_upload_rgba16f(mv_caster* c, uint32_t src, const uint16_t* texels);
int mv_volume_upload_r32f(mv_caster* c, uint32_t src, const float* density);
int mv_volume_read(mv_caster* c, uint32_t src, uint16_t* texels_out);
/* LoadVolumeData from a file (MultiRayCaster.cpp:168-209; DDS::Loader, XUSG/Advanced/XUSGDDSLoader.h:21-37): a 3-D DDS
 * with one scalar channel, of any resolution, resampled to the G^3 grid by CSR32FToRGBA16F's LINEAR fetch at the voxel
 * centres (CSR32FToRGBA16F.hlsl:16-26). mv_dds_parse is host-only (no device needed). */
enum { MV_DDS_R32_FLOAT = 1, MV_DDS_R16_FLOAT = 2, MV_DDS_R16_UNORM = 3, MV_DDS_R8_UNORM = 4 };
typedef struct mv_dds_info {
    uint32_t width, height, depth;
    uint32_t format;            /* MV_DDS_* */
    uint32_t bytes_per_texel;
    uint32_t data_offset;       /* of the top mip level in the file */
} mv_dds_info;
int mv_dds_parse(const char* path, mv_dds_info* out);
int mv_volume_upload_r32f_sized(mv_caster* c, uint32_t src, const float* density, uint32_t width, uint32_t height, uint32_t depth);
int mv_volume_load_dds(mv_caster* c, uint32_t src, const char* path);

/* SetRenderTargets / SetViewport (MultiRayCaster.h:37-38): scene depth D32 WxH (NULL = 1.0),
 * shadow map D16 SxS (NULL = none -> unshadowed), colour RT RGBA16F WxH that the volumes are
 * composited over (NULL = zeros), velocity RG16F for the TAA (NULL = zeros). HOST pointers, copied. */
int mv_set_targets(mv_caster* c, const float* depth, const uint16_t* shadow_d16, uint32_t shadow_size,
                   const uint16_t* color_rgba16f, const uint16_t* velocity_rg16f);
/* same, DEVICE pointers on the caster's device; copied device-to-device on the caster's stream */
int mv_set_targets_device(mv_caster* c, const float* depth, const uint16_t* shadow_d16, uint32_t shadow_size,
                          const uint16_t* color_rgba16f, const uint16_t* velocity_rg16f);
/* restores the colour RT to the background given to the last mv_set_targets* (what the mesh and
 * environment passes do every frame in the reference, MultiVolumes.cpp:645-673) */
int mv_reset_color(mv_caster* c);

int mv_set_sh(mv_caster* c, const float* coeffs27);                          /* SetSH, :41 (NULL = no probe) */
int mv_set_max_samples(mv_caster* c, uint32_t ray, uint32_t light);          /* SetMaxSamples, :42 */
int mv_set_volumes_world(mv_caster* c, float size, const float center[3]);   /* SetVolumesWorld, :43 */
int mv_set_volume_world(mv_caster* c, uint32_t i, float size, const float pos[3]); /* SetVolumeWorld, :44 */
int mv_set_volume_world_matrix(mv_caster* c, uint32_t i, const float world43[12]); /* animated transforms */
/* the same for `count` consecutive volumes starting at `first` in one call (count x 12 floats): what a caller that animates
 * every volume each frame does with N SetVolumeWorld calls (MultiRayCaster.h:44) */
int mv_set_volume_world_matrices(mv_caster* c, uint32_t first, uint32_t count, const float* world43);
int mv_set_light(mv_caster* c, const float pos[3], const float color[3], float intensity);   /* SetLight, :45 */
int mv_set_ambient(mv_caster* c, const float color[3], float intensity);                     /* SetAmbient, :46 */
/* UpdateFrame (:47-48, MultiRayCaster.cpp:316-353): builds CBPerFrame and the N PerObject records */
int mv_update_frame(mv_caster* c, const float view_proj[16], const float shadow_vp[16], const float eye[3]);

/* Render (:49-50, MultiRayCaster.cpp:355-385) = cull -> light march (one volume, round-robin) ->
 * view march -> OIT resolve into the colour RT; frame index++. The individual passes are exported
 * for the parity tests. `oit_method` is accepted for signature compatibility and ignored: the library has ONE OIT
 * implementation, the fused per-pixel resolve with the K-buffer semantics of the reference's default branch
 * (MultiRayCaster.cpp:377-381); its DXR / ray-query variants produce the same layers by other means. */
int mv_render(mv_caster* c, uint32_t oit_method);
/* Render with useWorkGraph = true (MultiRayCaster.h:49-50, MultiRayCaster.cpp:358-362, rayMarchWG :1370-1438,
 * LibRayMarch.hlsl:39-134): the light march runs first and takes its volume from the PREVIOUS frame's visible list,
 * then cull and view march run as ONE launch (the cull on CTA 0 of the persistent march kernel, which releases the
 * lists to the other CTAs), then the OIT passes. One GPU; not pipelined across frames. */
int mv_render_work_graph(mv_caster* c, uint32_t oit_method);
int mv_cull(mv_caster* c);                                   /* cullVolumes, MultiRayCaster.cpp:1249-1285 */
int mv_ray_march_light(mv_caster* c, int32_t volume_override); /* rayMarchL, :1299-1327; -1 = round-robin */
int mv_ray_march_view(mv_caster* c);                         /* rayMarchV, :1329-1368 */
int mv_resolve_oit(mv_caster* c);                            /* cubeDepthPeel+renderCube+resolveOIT, :1440-1633 */
/* ObjectRenderer::Postprocess (ObjectRenderer.h:46-48): CSTemporalAA (or copy when taa_on = 0) + PSToneMap */
int mv_postprocess(mv_caster* c, uint32_t taa_on);
/* SphericalHarmonics::Transform (XUSGSphericalHarmonics.h:25-26), order 3; cube = 6 x size x size x RGB f32 (host) */
int mv_sh_project(mv_caster* c, const float* cube_rgb_f32, uint32_t size, float* coeffs27_out);

/* ---- producer of the depth inputs: the occluder mesh ----
 * ObjectRenderer's depth-only passes: Init's OBJ import and AABB (ObjectRenderer.cpp:68-77,
 * XUSG/Optional/XUSGObjLoader.cpp:18-40, :166-228), SetWorld (:147-153), the light's orthographic
 * view-projection of UpdateFrame (:171-190), RenderShadow (:220-243, D16 1024^2) and the depth pre-pass
 * (:555-570, VSDepth.hlsl:25-28, D32 W x H). mv_mesh_render_depth rasterises both maps straight into the
 * caster's scene depth and shadow map (what SetRenderTargets borrows in the reference) and returns the
 * light's view-projection for mv_update_frame. Rasterisation rules: mv_mesh.cu. */
int  mv_obj_parse(const char* path, float** positions_xyz, uint32_t* num_vertices, uint32_t** indices, uint32_t* num_indices); /* host only */
void mv_obj_free(float* positions_xyz, uint32_t* indices);
int  mv_mesh_set(mv_caster* c, const float* positions_xyz, uint32_t num_vertices, const uint32_t* indices, uint32_t num_indices);
int  mv_mesh_load_obj(mv_caster* c, const char* path);
int  mv_mesh_set_world(mv_caster* c, float scale, const float pos[3]);
int  mv_mesh_render_depth(mv_caster* c, const float view_proj[16], float shadow_vp_out[16]);
/* ObjectRenderer::Render (ObjectRenderer.cpp:532-553; VSBasePass.hlsl:39-55, PSBasePass.hlsl:94-153): shadow pass + the
 * shaded base pass. Fills scene depth, the shadow map, the colour target the resolve composites over (clear_rgba where no
 * triangle covers the pixel; NULL = zeros) and the TAA velocity field from the mesh, with the caster's current light,
 * ambient and SH coefficients; view_proj and eye are ObjectRenderer::UpdateFrame's arguments (:171). Deviations (mv_mesh.cu header): no sub-pixel
 * jitter, no radiance cube term, zero velocity on the first frame after mv_mesh_set. */
int  mv_mesh_render(mv_caster* c, const float view_proj[16], const float eye[3], const float clear_rgba[4], float shadow_vp_out[16]);
int  mv_read_velocity(mv_caster* c, uint16_t* rg16f);
int  mv_read_depth(mv_caster* c, float* depth, uint16_t* shadow_d16, uint32_t* shadow_size);

/* ---- the environment under the volumes and the screenshot ----
 * LightProbe (MultiVolumes/Content/LightProbe.cpp:29-61): the radiance cube map, 6 x size x size x RGB f32 (host), D3D face
 * order; NULL / 0 = none. mv_render_environment = what the frame loop does to the colour RT before MultiRayCaster::Render
 * (MultiVolumes.cpp:645-673): the background given to mv_set_targets (the mesh pass's output), then RenderEnvironment
 * (LightProbe.cpp:85-97, PSEnvironment.hlsl:46-69) wherever the scene depth is 1 — the radiance along the pixel's ray,
 * alpha 0. Use it in place of mv_reset_color. mv_screenshot = MultiVolumes::SaveImage (MultiVolumes.cpp:744-764): the RGBA8
 * back buffer as a PNG; mv_write_png is host-only. */
int mv_set_environment(mv_caster* c, const float* cube_rgb_f32, uint32_t size);
int mv_render_environment(mv_caster* c);
int mv_write_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height);
int mv_screenshot(mv_caster* c, const char* path);

/* read-backs (host pointers; each synchronises the stream). The reference has no read-back API. */
int mv_read_per_object(mv_caster* c, float* out56xN);
int mv_read_visible(mv_caster* c, uint32_t* ids, uint32_t* count);
int mv_read_cube_volumes(mv_caster* c, uint32_t* ids, uint32_t* count);
int mv_read_attribs(mv_caster* c, uint16_t* out4xN);
int mv_read_cubemap(mv_caster* c, uint32_t volume, uint32_t mip, uint16_t* rgba16f, float* depth);
int mv_read_lightmap(mv_caster* c, uint32_t volume, uint16_t* rgba16f);
int mv_read_frame(mv_caster* c, uint16_t* rgba16f);
int mv_read_post(mv_caster* c, uint16_t* taa_rgba16f, uint8_t* rgba8);
/* Present (the swap-chain Present of the reference's frame loop, MultiVolumes.cpp:OnRender, with
 * FrameCount = 3 frames in flight): asynchronous read-back of the RGBA8 back buffer of the frame
 * rendered so far into PINNED host memory, on a copy stream, overlapping the next frame's passes.
 * slot < MV_PRESENT_SLOTS names the in-flight copy; mv_present_wait(slot) blocks until that copy has
 * landed. The next frame's post-process (or, sharded, its first peer barrier) waits for the copy on
 * the device, so the back buffer is never overwritten under it. host_rgba8 = NULL only marks the
 * frame's end (ranks that hold no frame use it to bound their queue depth). */
#define MV_PRESENT_SLOTS 3
int mv_present_async(mv_caster* c, uint8_t* host_rgba8_pinned, uint32_t slot);
int mv_present_wait(mv_caster* c, uint32_t slot);
int mv_get_stats(mv_caster* c, mv_stats* out);
int mv_get_timings(mv_caster* c, mv_timings* out);
int mv_set_frame_index(mv_caster* c, uint32_t frame_idx);
int mv_set_flags(mv_caster* c, uint32_t flags);            /* MV_FLAG_*: switch the instrumentation on or off between frames */
int mv_sync(mv_caster* c);

/* pinned host memory for the read-backs / uploads of a frame loop */
void* mv_host_alloc(size_t bytes);
void  mv_host_free(void* p);
/* page-lock memory the caller owns (e.g. a POSIX shared-memory frame that the ranks of a multi-GPU run all map) */
int   mv_host_register(void* p, size_t bytes);
int   mv_host_unregister(void* p);
/* Present of a sharded frame: like mv_present_async, but copies only the rows THIS rank resolved (its band or stripes) to
 * their place in `host_frame_rgba8` (the whole H x W frame, e.g. shared by all ranks' processes), over this rank's own
 * PCIe link — no rank has to carry the whole frame. */
int   mv_present_rows_async(mv_caster* c, uint8_t* host_frame_rgba8_pinned, uint32_t slot);

/* ---- multi-GPU (one process per GPU; no counterpart in the single-adapter reference) ----
 * Partition (BASELINE.json north_star): the cull is replicated; volume v's cube map is marched by
 * rank v % world; the light map of the frame's light volume is filled in z-slabs, slab r by rank r;
 * the OIT resolve + post-process run on a band of rows per rank. Every buffer that crosses GPUs lives
 * in ONE device allocation per caster, the exchange block:
 *     [cube-map arena | light-map staging | back buffer RGBA8 | barrier flags]
 * with identical layout on every rank. Two ways to move the data:
 *   (a) fused: map every peer's block with mv_ipc_export / mv_ipc_import; the march, light and
 *       post-process kernels then store their results straight into all peers' blocks over NVLink and
 *       mv_peer_barrier orders the phases with device-side flags — no collective call at all;
 *   (b) collective: leave the peers unmapped and run NCCL (all-gather / broadcast / gather) on
 *       regions of the block between the passes (mv_exchange_block gives the pointer, mv_cube_region
 *       and mv_exchange_layout the offsets). */
typedef struct mv_exchange_layout {
    uint64_t block_bytes;
    uint64_t arena_offset, arena_bytes;                 /* cube maps of all volumes, all mips */
    uint64_t light_staging_offset, light_staging_bytes; /* world x slab: [z][y][x] RGBA16F, slabs padded to ceil(L / world) */
    uint64_t back_buffer_offset, back_buffer_bytes;     /* H x W RGBA8 */
    uint64_t flags_offset, flags_bytes;
    uint32_t light_slab_depth;                          /* ceil(L / world) */
    uint32_t reserved;
    uint64_t light_staging2_offset;                     /* second staging buffer (frames pipelined across ranks alternate) */
    uint64_t direct_offset, direct_bytes;               /* results of the screen-space marches (RGBA16F per pixel of every direct-scheme
                                                           volume's rectangle): stored into every peer's block under volume-sharded storage */
    uint64_t history_offset[2], history_bytes;          /* the two TAA history images, H x W RGBA16F: a rank's TAA output rows
                                                           are stored into every peer's image too, because the next frame's
                                                           history fetch (uv - velocity, bilinear) may land on any row */
} mv_exchange_layout;

/* Volume-sharded storage (BASELINE.json configs[4]: 512 x 512^3 RGBA16F = 550 GB, more than one GPU holds): rank `rank` of
 * `world` keeps the full-resolution texture of the sources s with s % world == rank only, the light maps of the instances
 * that use them, and — for every other source — a DENSITY PROXY: the volume's density box-filtered to proxy_grid^3 (R16F;
 * grid_size must be a multiple of proxy_grid). Every rank is given the same ingest calls (mv_volume_init_procedural,
 * mv_volume_upload_*, mv_volume_load_dds) and keeps what it owns: no volume data crosses GPUs.
 *   - the cull is replicated; light maps, cube maps and screen-space marches of a volume are produced by its owner alone
 *     (cube-map texels and screen-space march results are stored into every peer's exchange block, as in the fused mode);
 *   - the light march reads the OTHER ranks' volumes through their proxies: inter-volume shadows and ambient occlusion
 *     cast by a remote volume are those of its box-filtered density. This is the one deviation from the replicated
 *     storage (stated in DESIGN.md); shadows within a volume and between volumes of one rank are exact;
 *   - a direct-scheme volume whose screen rectangle does not fit the march-result buffer is left out of the frame (the
 *     resolve cannot march a volume it does not hold); the buffer holds four full-screen rectangles.
 * Peers must be mapped (mv_ipc_import, or mv_set_peer_block for casters of one process) before mv_render when world > 1. */
int mv_create_sharded(const mv_desc* desc, uint32_t rank, uint32_t world, uint32_t proxy_grid, mv_caster** out);
/* a peer whose exchange block is directly addressable (another caster of this process on the same or a peer-enabled device) */
int mv_set_peer_block(mv_caster* c, uint32_t peer, void* peer_exchange_block);
int mv_set_shard(mv_caster* c, uint32_t rank, uint32_t world);
int mv_set_row_band(mv_caster* c, uint32_t row0, uint32_t row1);     /* rows this rank resolves and post-processes */
/* interleaved alternative to a contiguous band (better balance when the expensive pixels cluster): with
 * stripe_height > 0 and world > 1 the rank owns the rows r with (r / stripe_height) % world == rank; 0 = band */
int mv_set_row_stripes(mv_caster* c, uint32_t stripe_height);
int mv_exchange_block(mv_caster* c, void** dev_ptr, uint64_t* bytes);
int mv_exchange_layout_get(mv_caster* c, mv_exchange_layout* out);
/* byte offsets (from the block base) of volume `volume`'s cube map at `mip`: [face][y][x] colour, then depth */
int mv_cube_region(mv_caster* c, uint32_t volume, uint32_t mip, uint64_t* color_offset, uint64_t* color_bytes,
                   uint64_t* depth_offset, uint64_t* depth_bytes);
int mv_ipc_export(mv_caster* c, void* handle64);                     /* 64-byte cudaIpcMemHandle_t of the block */
int mv_ipc_import(mv_caster* c, uint32_t peer_rank, const void* handle64);
int mv_peer_barrier(mv_caster* c);                                   /* device-side all-ranks barrier on the stream (fused mode) */
int mv_light_commit(mv_caster* c);                                   /* light-map staging -> the light volume's 3-D array */
/* enqueue the caster's work on an external stream (e.g. the one NCCL collectives are ordered on); NULL = own stream */
int mv_set_stream(mv_caster* c, void* cuda_stream);
int mv_get_stream(mv_caster* c, void** cuda_stream);
/* device pointers of the frame images: colour RT RGBA16F, TAA output RGBA16F (current), back buffer RGBA8 */
int mv_frame_buffers(mv_caster* c, void** color_rgba16f, void** post_rgba16f, void** back_rgba8);

#ifdef __cplusplus
}
#endif
#endif
