// MultiRayCaster.hpp — header-only C++ mirror of the reference's operator surface over the C-ABI.
//
// Same method names, argument meaning and bool-returning error behaviour as
// `class MultiRayCaster` (reference: MultiVolumes/Content/MultiRayCaster.h:28-50) so that
// MultiVolumes.cpp (its only caller: :217-235 init, :310-311 targets, :345-355 per-frame update,
// :640 SH, :674 render, :677 post-process) can switch by changing one include. The XUSG arguments
// (command lists, descriptor-table library, upload buffers) disappear: libmv_b200.so owns one CUDA
// stream per caster. Matrices are row-major float[16] in the DirectXMath row-vector convention, i.e.
// exactly the XMFLOAT4X4 contents the reference passes.
#pragma once
#include "mv.h"
#include <cstdint>
#include <string>

namespace mvb200 {

enum OITMethod : uint32_t { OIT_K_BUFFER, OIT_RAY_TRACING, OIT_RAY_QUERY };   // MultiRayCaster.h:20-26; all map to the fused resolve

struct Float3 { float x, y, z; };

class MultiRayCaster {
public:
    MultiRayCaster() = default;
    MultiRayCaster(const MultiRayCaster&) = delete;
    MultiRayCaster& operator=(const MultiRayCaster&) = delete;
    ~MultiRayCaster() { mv_destroy(m_h); }

    // Init (MultiRayCaster.h:31-34). width/height come from SetViewport in the reference.
    bool Init(uint32_t width, uint32_t height, uint32_t gridSize, uint32_t lightGridSize, uint32_t numVolumes,
              uint32_t numVolumeSrcs, uint32_t device = 0, uint32_t flags = 0)
    {
        mv_desc d{};
        d.grid_size = gridSize; d.light_grid_size = lightGridSize; d.num_volumes = numVolumes; d.num_volume_srcs = numVolumeSrcs;
        d.width = width; d.height = height; d.max_ray_samples = 256; d.max_light_samples = 96; d.device = device; d.flags = flags;
        mv_destroy(m_h); m_h = nullptr;
        return ok(mv_create(&d, &m_h));
    }
    // The same with volume-sharded storage (one process per GPU; no counterpart in the single-adapter reference): this rank
    // keeps the sources s % world == rank at full resolution and a proxyGrid^3 density proxy of the others (mv.h)
    bool InitSharded(uint32_t width, uint32_t height, uint32_t gridSize, uint32_t lightGridSize, uint32_t numVolumes,
                     uint32_t numVolumeSrcs, uint32_t rank, uint32_t world, uint32_t proxyGrid, uint32_t device = 0, uint32_t flags = 0)
    {
        mv_desc d{};
        d.grid_size = gridSize; d.light_grid_size = lightGridSize; d.num_volumes = numVolumes; d.num_volume_srcs = numVolumeSrcs;
        d.width = width; d.height = height; d.max_ray_samples = 256; d.max_light_samples = 96; d.device = device; d.flags = flags;
        mv_destroy(m_h); m_h = nullptr;
        return ok(mv_create_sharded(&d, rank, world, proxyGrid, &m_h));
    }
    // LoadVolumeData (:35-36): R32F density (the DDS payload) through the CSR32FToRGBA16F conversion, or RGBA16F texels
    bool LoadVolumeData(uint32_t i, const float* density) { return ok(mv_volume_upload_r32f(m_h, i, density)); }
    bool LoadVolumeData(uint32_t i, const uint16_t* rgba16f) { return ok(mv_volume_upload_rgba16f(m_h, i, rgba16f)); }
    // LoadVolumeData(cmdList, i, fileName, uploaders), MultiRayCaster.h:35-36: a 3-D scalar DDS of any resolution
    bool LoadVolumeData(uint32_t i, const char* ddsFileName) { return ok(mv_volume_load_dds(m_h, i, ddsFileName)); }
    // InitVolumeData (:40)
    bool InitVolumeData(uint32_t i, uint32_t mode = 0, uint32_t seed = 0) { return ok(mv_volume_init_procedural(m_h, i, mode, seed)); }
    // SetRenderTargets + SetViewport (:37-38): scene depth, shadow map, colour RT (device pointers, as the reference borrows GPU resources)
    bool SetRenderTargets(const float* depth, const uint16_t* shadowD16, uint32_t shadowSize, const uint16_t* colorRGBA16F,
                          const uint16_t* velocityRG16F = nullptr)
    { return ok(mv_set_targets_device(m_h, depth, shadowD16, shadowSize, colorRGBA16F, velocityRG16F)); }
    bool SetRenderTargetsHost(const float* depth, const uint16_t* shadowD16, uint32_t shadowSize, const uint16_t* colorRGBA16F,
                              const uint16_t* velocityRG16F = nullptr)
    { return ok(mv_set_targets(m_h, depth, shadowD16, shadowSize, colorRGBA16F, velocityRG16F)); }
    void SetSH(const float* coeffs27) { mv_set_sh(m_h, coeffs27); }                                      // :41
    void SetMaxSamples(uint32_t maxRaySamples, uint32_t maxLightSamples) { mv_set_max_samples(m_h, maxRaySamples, maxLightSamples); }   // :42
    void SetVolumesWorld(float size, const Float3& center) { const float c[3] = {center.x, center.y, center.z}; mv_set_volumes_world(m_h, size, c); }   // :43
    void SetVolumeWorld(uint32_t i, float size, const Float3& pos) { const float p[3] = {pos.x, pos.y, pos.z}; mv_set_volume_world(m_h, i, size, p); } // :44
    void SetVolumeWorld(uint32_t i, const float world4x3[12]) { mv_set_volume_world_matrix(m_h, i, world4x3); }
    void SetVolumeWorlds(uint32_t first, uint32_t count, const float* world4x3) { mv_set_volume_world_matrices(m_h, first, count, world4x3); }
    void SetLight(const Float3& pos, const Float3& color, float intensity)                              // :45
    { const float p[3] = {pos.x, pos.y, pos.z}, c[3] = {color.x, color.y, color.z}; mv_set_light(m_h, p, c, intensity); }
    void SetAmbient(const Float3& color, float intensity) { const float c[3] = {color.x, color.y, color.z}; mv_set_ambient(m_h, c, intensity); }       // :46
    // UpdateFrame (:47-48): shadowVP arrives already transposed in the reference (ObjectRenderer.cpp:185); pass the un-transposed matrix here
    void UpdateFrame(const float viewProj[16], const float shadowVP[16], const Float3& eyePt)
    { const float e[3] = {eyePt.x, eyePt.y, eyePt.z}; mv_update_frame(m_h, viewProj, shadowVP, e); }
    // Render (:49-50)
    bool Render(OITMethod oitMethod = OIT_K_BUFFER, bool useWorkGraph = false)
    {
        return ok(useWorkGraph ? mv_render_work_graph(m_h, oitMethod) : mv_render(m_h, oitMethod));
    }
    // ObjectRenderer::Postprocess (ObjectRenderer.h:46-48)
    bool Postprocess(bool taa = true) { return ok(mv_postprocess(m_h, taa ? 1u : 0u)); }
    // XUSG SphericalHarmonics::Transform (XUSGSphericalHarmonics.h:25-26)
    // ObjectRenderer's depth-only passes (ObjectRenderer.h: Init's mesh import, SetWorld, RenderShadow + depth pre-pass):
    // RenderMeshDepth fills the scene depth and the shadow map this caster reads, and returns UpdateFrame's shadowVP
    bool LoadMesh(const char* objFileName) { return ok(mv_mesh_load_obj(m_h, objFileName)); }
    bool SetMesh(const float* positions, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices)
    { return ok(mv_mesh_set(m_h, positions, numVertices, indices, numIndices)); }
    bool SetMeshWorld(float scale, const float pos[3]) { return ok(mv_mesh_set_world(m_h, scale, pos)); }
    bool RenderMeshDepth(const float viewProj[16], float shadowVP[16]) { return ok(mv_mesh_render_depth(m_h, viewProj, shadowVP)); }
    // ObjectRenderer::UpdateFrame + RenderShadow + Render: shadow pass + shaded base pass (colour, depth, velocity)
    bool RenderMesh(const float viewProj[16], const float eyePt[3], const float clearRGBA[4], float shadowVP[16])
    { return ok(mv_mesh_render(m_h, viewProj, eyePt, clearRGBA, shadowVP)); }
    // LightProbe (LightProbe.h): the radiance cube map; RenderEnvironment prepares the colour RT of a frame (the background
    // given to SetRenderTargets, the environment where the scene depth is 1) — call it before Render, as MultiVolumes.cpp:673-674
    bool SetEnvironment(const float* cubeRGB, uint32_t size) { return ok(mv_set_environment(m_h, cubeRGB, size)); }
    bool RenderEnvironment() { return ok(mv_render_environment(m_h)); }
    // MultiVolumes::SaveImage (MultiVolumes.cpp:744-764): the back buffer as a PNG
    bool Screenshot(const char* fileName) { return ok(mv_screenshot(m_h, fileName)); }
    // Present with FrameCount = 3 frames in flight: asynchronous read-back of the RGBA8 back buffer into pinned memory
    bool Present(uint8_t* pinnedRGBA8, uint32_t slot) { return ok(mv_present_async(m_h, pinnedRGBA8, slot)); }
    bool WaitPresent(uint32_t slot) { return ok(mv_present_wait(m_h, slot)); }

    bool TransformSH(const float* cubeRGB, uint32_t size, float coeffs27[27]) { return ok(mv_sh_project(m_h, cubeRGB, size, coeffs27)); }

    bool ReadFrame(uint16_t* rgba16f) { return ok(mv_read_frame(m_h, rgba16f)); }
    bool ReadBackBuffer(uint8_t* rgba8) { return ok(mv_read_post(m_h, nullptr, rgba8)); }
    bool GetStats(mv_stats& s) { return ok(mv_get_stats(m_h, &s)); }
    mv_caster* Handle() const { return m_h; }
    static std::string LastError() { return mv_last_error(); }

    static const uint8_t FrameCount = 3;   // MultiRayCaster.h:52

private:
    static bool ok(int rc) { return rc == MV_OK; }
    mv_caster* m_h = nullptr;
};

} // namespace mvb200
