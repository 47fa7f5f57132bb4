// mvo_core.h — ORACLE (test infrastructure, not product code): scene state shared by the passes.
#pragma once
#include "mvo.h"
#include "mvo_math.h"
#include "mvo_sampler.h"
#include <vector>
#include <atomic>

namespace mvo {

// SharedConsts.h:5-10
constexpr uint32_t kGroupVolumeCount = 4;
constexpr uint32_t kNumCubeMip = 5;
constexpr uint32_t kNumOitLayers = 8;
constexpr float kZNear = 1.0f, kZFar = 1000.0f;
// Common.hlsli:12, RayMarch.hlsli:11-12,17
constexpr uint32_t kCubeMapRayMarchBit = 1u << 15;
// RayMarch.hlsli:11-12, :17; CSRayMarch.hlsl:155; PSResolveOIT.hlsl:22; CSTemporalAA.hlsl (1 / 9). These are `min16float`
// literals in the HLSL. The reference is compiled WITHOUT -enable-16bit-types, so min-precision is a hint: arithmetic is fp32
// here; the LITERALS, however, are folded to binary16 in the shipped DXIL, which is what every D3D12
// driver feeds its ALUs, and with them — only with them — this oracle reproduces the compiled marches bit for bit
// (tests/test_dxil_golden.py). mvo_set_min16_consts_as_half(0) switches to the decimal text of the HLSL so that the difference
// can be measured (tests/test_oracle_kat.py::test_min16_consts_as_half_delta, tools/min16_delta.py).
struct Min16Consts {
    // defaults: the binary16 values the reference's shipped DXIL holds (Bin/*.cso; oracle/dxil). mvo_set_min16_consts_as_half(0)
    // switches to the decimal literals of the HLSL text, to report how far the two readings sit apart.
    float absorption = 0.7998046875f;          // 0xH3A66
    float zeroThreshold = 0.01000213623046875f; // 0xH211F
    float maxDist = 3.46484375f;               // 0xH42EE
    float invTwoPi = 0.1591796875f;            // 0xH3118; 0 = divide by 2 pi (source form)
    float alphaClamp = 0.99951171875f;         // 0xH3BFF
    float ninth = 0.111083984375f;             // 0xH2F1C
    float toneScale = 1.0498046875f, toneBias = 0.7001953125f;   // 0xH3C33, 0xH399A (PSToneMap.cso)
};
extern Min16Consts g_min16;
#define kAbsorption (mvo::g_min16.absorption)
#define kZeroThreshold (mvo::g_min16.zeroThreshold)
constexpr float kFltMax = 3.402823466e+38f;
constexpr float kPi = 3.1415926535897f;           // SHIrradiance.hlsli:6

// Common.hlsli:28-34 / MultiRayCaster.cpp:35-41. Stored un-transposed, row-vector convention.
struct PerObject {
    m44 WorldViewProj;
    m44 WorldViewProjI;
    m43 WorldI;
    m43 World;
};

// Common.hlsli:39-49
struct PerFrame {
    f3 eyePt;
    f2 viewport;
    m44 screenToWorld;
    m44 shadowViewProj;
    f4 lightPos, lightColor, ambient;
    uint32_t frameIdx;
};

struct CubeMap {               // RGBA16F colour + R32F depth, 6 faces x kNumCubeMip mips
    std::vector<uint16_t> color[kNumCubeMip];
    std::vector<float> depth[kNumCubeMip];
};

struct Caster {
    mvo_desc d;
    std::vector<Tex3D> volumes;           // per source
    std::vector<Tex3D> lightMaps;         // per instance; RGBA16F holding R11G11B10F-quantised rgb
    std::vector<CubeMap> cubeMaps;        // per instance
    std::vector<m43> volumeWorlds;        // m_volumeWorlds
    std::vector<PerObject> perObject;
    std::vector<uint32_t> volumeDescs;    // VolTexId:14 | NumMips:4 | CubeMapSize:14
    std::vector<uint16_t> attribs;        // N x {MipLevel, SmpCount, MaskBits, VolTexId}
    std::vector<uint32_t> visible, cubeVolumes;
    PerFrame cb;
    f3 lightPt; f4 lightColor, ambient;
    bool hasSH = false; f3 sh[9];
    // borrowed targets (copied in)
    std::vector<float> depth;             // W*H, D32
    std::vector<uint16_t> shadow;         // S*S, D16 unorm
    uint32_t shadowSize = 0;
    std::vector<uint16_t> color;          // W*H RGBA16F: in = background, out = composited frame
    std::vector<uint16_t> background;     // the colour RT as given to set_targets (mvo_reset_color)
    std::vector<uint16_t> velocity;       // W*H RG16F
    std::vector<uint16_t> taaHistory[2];  // RGBA16F ping-pong
    std::vector<uint8_t> backBuffer;      // RGBA8
    uint32_t frameParity = 0;
    uint32_t frameIdx = 0;
    mvo_stats stats;
    int filterModel = MODEL_SM100;
    // sharding (mirrors mv_set_shard / mv_set_row_band of the product): volume v is marched by rank v % world,
    // the light map is filled in z-slabs of ceil(L / world), OIT and post-process cover rows [row0, row1)
    uint32_t shardRank = 0, shardWorld = 1;
    uint32_t row0 = 0, row1 = 0;
    // volume-sharded storage (mirrors mv_create_sharded of the product): rank r holds the full-resolution texture of the
    // sources s with s % world == r only; the light march reads every other volume through a density proxy (box-filtered
    // to proxyGrid^3), marches light maps and cube maps of its own volumes only, and is not split in z-slabs
    bool shardVolumes = false;
    uint32_t proxyGrid = 0;
    std::vector<Tex3D> proxies;           // per source (empty for the rank's own sources)
    bool owns_source(uint32_t src) const { return !shardVolumes || src % shardWorld == shardRank; }
    const Tex3D& density_source(uint32_t src) const { return owns_source(src) ? volumes[src] : proxies[src]; }
    // radiance cube map of the environment pass (LightProbe), RGBA16F [face][y][x]
    std::vector<uint16_t> envCube;
    uint32_t envSize = 0;
    // per-fragment record of the last resolve_oit (mvo_debug_oit; for oracle/dxil: inputs of PSCube / PSResolveOIT per pixel)
    bool debugOIT = false;
    std::vector<uint32_t> dbgCount;       // W*H layers per pixel
    std::vector<uint32_t> dbgInfo;        // W*H*8 x {depth key, volume, cube face (+x -x +y -y +z -z), stored}
    std::vector<float> dbgData;           // W*H*8 x {lpt xyz, face uv, colour rgba}
    std::vector<float> dbgResult;         // W*H x rgba: the blended layers before the render-target blend
    bool debugF32 = false;                // keep the last marches' outputs before their format conversion (mvo_debug_f32)
    std::vector<float> dbgCubeF32;        // N x 6 x G x G x 4: scatter of the view march at the volume's mip (top-left corner of the slab)
    std::vector<float> dbgLightF32;       // L^3 x 3: the light march's value before the R11G11B10 store
    std::vector<uint32_t> dbgAllKeys;     // W*H x N: depth key of EVERY fragment of the pixel, in draw (visible-list) order; 0xffffffff = none
    // occluder mesh (mvo_mesh.cpp)
    std::vector<float> meshPos;           // V x 3
    std::vector<uint32_t> meshIdx;        // 3 T
    std::vector<float> meshNrm;           // V x 3, recomputed
    m44 meshWvpPrev; bool meshHavePrev = false;
    float meshExtent = 1.0f, meshScale = 1.0f;
    f3 meshPosition = {0.0f, 0.0f, 0.0f};
};

// passes (mvo_passes.cpp)
void cull_volumes(Caster& c);
void ray_march_light(Caster& c, int volumeOverride);
void ray_march_view(Caster& c);
void resolve_oit(Caster& c);
void temporal_aa(Caster& c, bool taaOn);
void render_environment(Caster& c);
void tone_map(Caster& c);
void init_grid_data(Caster& c, uint32_t src, uint32_t mode, uint32_t seed);
void sh_project(const float* cubeRGB, uint32_t size, float out27[27]);
f4 evaluate_sh_irradiance(const f3 sh[9], f3 norm);
// mvo_mesh.cpp: depth-only rasterisation of an indexed triangle list under wvp into depth[width * height]
void raster_depth(const std::vector<float>& pos, const std::vector<uint32_t>& idx, const m44& wvp, uint32_t width, uint32_t height, float* depth);
void recompute_normals(const std::vector<float>& pos, const std::vector<uint32_t>& idx, std::vector<float>& nrm);
void render_base_pass(Caster& c, const m44& wvp, const m44& wvpPrev, const m43& world, const m44& shadowWVP, f3 eye, const float clear[4]);

} // namespace mvo
