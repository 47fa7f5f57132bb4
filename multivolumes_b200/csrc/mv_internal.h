// mv_internal.h — device-visible records and the host-side caster state of libmv_b200.so.
#pragma once
#include "mv_math.cuh"
#include "../../include/mv.h"
#include <vector>
#include <string>

namespace mv {

constexpr int kMaxPeers = 8;
constexpr int kUploadRing = 3;   // MultiRayCaster.h:52 FrameCount

// Common.hlsli:28-34 / MultiRayCaster.cpp:35-41 — 56 floats, matrices NOT transposed (row-vector use).
struct PerObject {
    float wvp[16];
    float wvpi[16];
    float worldI[12];
    float world[12];
};

// Common.hlsli:39-56 (cbPerFrame + cbSampleRes) plus the SH coefficients (g_roSHCoeffs, 9 x float3).
// Passed to every kernel by value (constant bank).
struct FrameCB {
    float eye[3];
    float viewport[2];
    float inv2Viewport[2];     // 2 / viewport (the same fp32 quotient the shader's `/ g_viewport * 2` restatement divides out per pixel)
    float screenToWorld[16];
    float shadowViewProj[16];
    float lightPos[4];
    float lightColor[4];
    float ambient[4];
    uint32_t frameIdx;
    uint32_t hasSH;
    float sh[27];
    uint32_t numVolumes;
    uint32_t gridSize;          // G
    uint32_t lightGridSize;     // L
    uint32_t width, height;
    uint32_t maxRaySamples, maxLightSamples;
    uint32_t shadowSize;        // 0 = no shadow map bound
};

// Per-frame lists produced by the cull kernel (device memory, one allocation).
struct FrameLists {
    uint32_t visibleCount;
    uint32_t cubeCount;
    uint32_t marchTileTotal;    // total 8x4 texel tiles of the view march
    uint32_t marchTileCursor;   // work-stealing cursor of the persistent view-march kernel
    uint32_t lightVolume;       // CSRayMarchL.hlsl:29-33
    uint32_t oitTileCursor;
    uint32_t lightDenseCount;   // light march: voxels with density >= 0.01 appended by the classify pass
    uint32_t lightDenseCursor;  // work cursor of the persistent shadow-march kernel
    uint32_t lightItemCount;    // deferred ambient-occlusion rays: slots of the volume-sorted item list (segments padded to 32)
    uint32_t lightItemCursor;   // work cursor of the persistent AO-march kernel
    uint32_t lightResultCount;  // AO factors reserved, voxel-major (a voxel's factors are contiguous, ascending volume index)
    uint32_t lightEmitCursor;   // work cursor of the item-emission kernel
    uint32_t lightOverflow;     // the frame's AO rays do not fit the item buffers: they are marched inline instead
    uint32_t directTileTotal;   // 8x4-pixel tiles of the screen-space marches (direct-scheme volumes)
    uint32_t directTileCursor;  // work cursor of the persistent direct-march kernel
    uint32_t cullSerial;        // fused cull -> view march: the launch whose cull results these lists hold (k_ray_march_v.cu)
    uint32_t marchTileCursor2;  // cursor of the second view-march launch of a sharded frame (the light volume's own tiles)
    uint32_t pad0[1];
    // followed in memory by: visible[N], cubeVolumes[N], cubeTilePrefix[N + 1], directTilePrefix[N + 1], directOffset[N], marchOrder[N], cubeTileBegin[N]
};

// Per visible volume (same order as the visible list), written by the cull for the OIT resolve: the
// eye in the volume's local space and a conservative screen rectangle of its projected box.
struct VisInfo {
    float eyeL[3];
    uint32_t volumeId;
    int x0, y0, x1, y1;   // inclusive pixel bounds; full screen when a corner is behind the eye plane
};

// Light march, per dense voxel (k_ray_march_l.cu): what the finalize pass needs once the voxel's
// deferred ambient-occlusion rays have been marched.
constexpr uint32_t kLightRecHits = 6;
struct alignas(16) LightRec {   // moved as four 16-byte words
    uint32_t voxel;        // (z L + y) L + x
    uint32_t itemBase;     // first of the voxel's AO factors (ascending volume index)
    uint32_t itemCount;
    float shadow;          // transmittance toward the light through all volumes
    float aoDir[3];        // world-space AO direction (normalised negative density gradient)
    float ao;              // AO product of the rays marched inline (1 when every ray was deferred)
    uint32_t castEnd;      // first volume at which the shadow ray was no longer cast
    uint32_t hits[kLightRecHits];   // the first AO rays of the voxel: volume | castShadow << 31
    uint32_t pad;
};
static_assert(sizeof(LightRec) == 64, "LightRec layout");

constexpr uint32_t kNoDirect = 0xffffffffu;
// bytes of the per-frame lists in front of the VisInfo records (mv_api.cu lays the block out)
MV_HD size_t frame_lists_header_bytes(size_t N) { return (sizeof(FrameLists) + (7 * N + 2) * sizeof(uint32_t) + 31) & ~(size_t)31; }

struct StatsDev {
    unsigned long long view_rays, view_samples, view_light_fetches;
    unsigned long long light_voxels, light_dense_voxels, light_samples;
    unsigned long long direct_rays, direct_samples, direct_light_fetches;
    unsigned long long oit_fragments;
    unsigned long long view_skipped, direct_skipped;   // samples whose texture fetch the occupancy bricks made unnecessary
};

// Cube-map arena: one device allocation holding, for every volume, the RGBA16F colour and R32F depth
// of all kNumCubeMip mips ([mip][face][y][x]). The same layout on every rank of a multi-GPU run, so
// a peer's texel lives at the same byte offset in the peer's arena.
struct CubeArena {
    unsigned char* base;                 // this device
    unsigned char* peer[kMaxPeers];      // peer-mapped arenas (nullptr = not mapped); peer[rank] = nullptr
    uint32_t numPeers;                   // world size when peers are mapped, else 0
    unsigned long long colorStride;      // bytes per volume, colour
    unsigned long long depthBase;        // byte offset of the depth region
    unsigned long long depthStride;      // bytes per volume, depth
    uint32_t mipTexelOffset[kNumCubeMip + 1];   // texel offset of each mip inside a volume slot
};

MV_HD unsigned long long arena_color_offset(const CubeArena& a, uint32_t volume, uint32_t mip)
{
    return (unsigned long long)volume * a.colorStride + (unsigned long long)a.mipTexelOffset[mip] * 8ull;
}
MV_HD unsigned long long arena_depth_offset(const CubeArena& a, uint32_t volume, uint32_t mip)
{
    return a.depthBase + (unsigned long long)volume * a.depthStride + (unsigned long long)a.mipTexelOffset[mip] * 4ull;
}

// Empty-space bricks of the source volumes (k_init.cu, k_build_occupancy): one bit per brick of 2^shift texels per
// side; a set bit means that every texel a trilinear fetch from inside the brick can touch (the brick and one texel
// around it) holds a density <= kZeroThreshold, so the filtered density — a convex combination — cannot exceed the
// threshold either and the march treats the sample as the empty sample it is without fetching it
// (CSRayMarch.hlsl:128 `if (color.w > ZERO_THRESHOLD)`: an empty sample changes nothing but t). Results are unchanged.
struct Occupancy {
    const uint32_t* bits;          // [srcs][wordsPerVolume]; nullptr = no skipping
    uint32_t wordsPerVolume;
    uint32_t shift;                // log2 of the brick edge in texels
    uint32_t bricks;               // bricks per axis
    float gridSize;                // G as float
    float halfBricks;              // bricks / 2: a local-space coordinate in [-1, 1] times this, plus this, is its brick coordinate
};

// Everything the kernels need, passed by value.
struct DeviceScene {
    const PerObject* perObject;          // [N]
    const uint32_t* volumeDescs;         // [N] VolTexId:14 | NumMips:4 | CubeMapSize:14
    ushort4* attribs;                    // [N] {MipLevel, SmpCount, MaskBits, VolTexId}
    FrameLists* lists;
    uint32_t* visible;                   // [N]
    uint32_t* cubeVolumes;               // [N]
    uint32_t* cubeTilePrefix;            // [N + 1] over marchOrder
    uint32_t* marchOrder;                // [N] indices into cubeVolumes, longest rays first
    uint32_t* cubeTileBegin;             // [N] over marchOrder: first tile of the volume this rank marches (tile-balanced sharding)
    VisInfo* visInfo;                    // [N]
    uint32_t* directTilePrefix;          // [N + 1] over the visible list: tiles of the screen-space march of each direct-scheme volume
    uint32_t* directOffset;              // [N] over the visible list: first pixel of the volume's rectangle in directColor (kNoDirect = none)
    uint2* directColor;                  // RayCast results (RGBA16F as stored in the K-buffer), rectangle by rectangle
    uint2* directStats;                  // per result {samples | marched << 31, light fetches}; nullptr when counters are off
    uint32_t directCapacity;             // pixels
    const cudaTextureObject_t* volumeTex;   // [srcs]
    Occupancy occ;                       // empty-space bricks of the source volumes
    const cudaTextureObject_t* lightTex;    // [N]
    const cudaSurfaceObject_t* lightSurf;   // [N]
    const float* depth;                  // W*H D32
    const uint16_t* shadow;              // S*S D16
    uint2* lightDense;                   // [L^3] {voxel index, shadow-test result} of the dense voxels of the frame's light volume
    LightRec* lightRecs;                 // [L^3] per dense voxel, same order as lightDense
    uint4* lightItems;                   // [lightItemCapacity] {record index, volume | castShadow << 31, result index, -}, sorted by volume
    float* lightItemResults;             // [lightItemCapacity] AO factor of each deferred ray, voxel-major
    uint32_t* lightSeg;                  // [2 N] per volume: AO-ray count of the frame, then the append cursor of its segment
    uint32_t lightItemCapacity;
    uint2* color;                        // W*H RGBA16F (half4 as uint2)
    StatsDev* stats;                     // nullptr when counters are off
    CubeArena arena;
    uint32_t shardVolumes;               // volume-sharded storage (mv_create_sharded): source s lives on rank s % world
    const unsigned char* srcIsProxy;     // [srcs] 1 = this rank holds the source as an R16F density proxy only; nullptr = none
    uint2* directPeer[kMaxPeers];        // the peers' directColor buffers (volume-sharded storage), nullptr otherwise
    uint32_t shardRank, shardWorld;      // volume v is marched by rank v % world
    uint32_t row0, row1;                 // rows of the frame this rank resolves
    uint32_t stripeH;                    // > 0: interleaved stripes (r / stripeH) % shardWorld == shardRank instead of the band
};

// Row ownership. Band: rows [row0, row1). Stripes (multi-GPU, better balanced): stripe g = rows
// [g stripeH, (g + 1) stripeH) belongs to rank g % world; the rank's k-th stripe is g = k world + rank.
MV_HD uint32_t num_own_stripes(uint32_t height, uint32_t stripeH, uint32_t rank, uint32_t world)
{
    const uint32_t total = (height + stripeH - 1) / stripeH;
    return total > rank ? (total - rank + world - 1) / world : 0;
}

struct Caster;

// kernel launchers (one translation unit per pass)
struct Volume3D;
void launch_init_grid(Caster& c, Volume3D& target, uint32_t mode, uint32_t seed);
void launch_r32f_to_rgba16f(Caster& c, Volume3D& target, const float* devDensity);
void launch_build_occupancy(Caster& c, uint32_t src);   // after every write to a source volume
void launch_build_proxy(Caster& c, const Volume3D& full, Volume3D& proxy);
// Every write to a source volume goes through these two: begin_ingest hands out the texture to write (the source's own, or —
// when this rank keeps only a proxy of it — a temporary full-resolution one), end_ingest derives what the rank keeps from it
// (empty-space bricks, or the proxy) and releases the temporary.
struct IngestTarget { Volume3D* vol = nullptr; bool temporary = false; };
int begin_ingest(Caster& c, uint32_t src, IngestTarget& t);
int end_ingest(Caster& c, uint32_t src, IngestTarget& t);
void launch_cull(Caster& c);
void launch_pick_light_volume(Caster& c);
void launch_sh_project(Caster& c, const float* devCube, uint32_t size, float* devOut27);
void launch_light_commit(Caster& c);
constexpr uint32_t kBarrierMain = 0, kBarrierLight = 1;
void launch_peer_barrier(Caster& c);                       // signal + wait on the main channel
void launch_peer_signal(Caster& c, uint32_t channel);
void launch_peer_wait(Caster& c, uint32_t channel);
int check_peer_timeout(Caster& c);                          // MV_ERR_PEER_TIMEOUT once after a k_peer_wait gave up

void launch_ray_march_light(Caster& c, int volumeOverride);
// phase 0: every cube-map volume; 1: all but the frame's light volume (at most blocksPerSM CTAs per SM, 0 = all that fit);
// 2: the light volume alone
void launch_ray_march_view(Caster& c, uint32_t phase = 0, int blocksPerSM = 0);
void launch_cull_and_ray_march_view(Caster& c);
void launch_ray_cast_direct(Caster& c);
void launch_resolve_oit(Caster& c);
void launch_postprocess(Caster& c, bool taaOn);
void build_tone_lut(Caster& c);
void launch_environment(Caster& c, bool copyBackground);
// mv_render_environment may leave its pass to the next mv_render (Caster::envDeferred), which runs it beside the view march;
// every other entry point runs it first (MV_ENTER)
void flush_deferred(Caster& c);
bool frame_is_pipelined_on_one_gpu(const Caster& c);

struct Volume3D {
    bool proxy = false;                  // volume-sharded storage: an R16F density proxy of another rank's source
    uint32_t edge = 0;                   // texels per side
    uint32_t channels = 4;               // 4 = RGBA16F, 1 = R16F (density-only storage of the source volumes)
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    cudaSurfaceObject_t surf = 0;
};

struct Caster {
    mv_desc d{};
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    // host scene state (MultiRayCaster.h:189-215)
    std::vector<float> volumeWorlds;     // N x 12 (float4x3)
    float lightPt[3] = {75.0f, 75.0f, -75.0f};
    float lightColor[4] = {1.0f, 0.7f, 0.3f, 1.0f};
    float ambient[4] = {0.0f, 0.3f, 1.0f, 0.4f};
    FrameCB cb{};
    uint32_t frameIdx = 0;
    uint32_t frameParity = 0;
    // device resources
    std::vector<Volume3D> volumes;       // per source, RGBA16F G^3
    std::vector<Volume3D> lightMaps;     // per instance, RGBA16F L^3 holding R11G11B10F-quantised rgb
    uint32_t* dOcc = nullptr;            // empty-space bricks of every source volume (Occupancy)
    uint32_t occWords = 0, occShift = 0, occBricks = 0;
    cudaTextureObject_t* dVolumeTex = nullptr;
    cudaTextureObject_t* dLightTex = nullptr;
    cudaSurfaceObject_t* dLightSurf = nullptr;
    PerObject* dPerObject = nullptr;
    PerObject* hPerObjectPinned = nullptr;  // ring of kUploadRing x N records (the reference keeps FrameCount = 3 upload slots)
    uint32_t uploadSlot = 0;
    uint32_t* dVolumeDescs = nullptr;
    ushort4* dAttribs = nullptr;
    unsigned char* dLists = nullptr;     // FrameLists + visible + cubeVolumes + cubeTilePrefix
    StatsDev* dStats = nullptr;
    // occluder mesh (mv_mesh.cu): the producer of the scene depth and the shadow map
    float* dMeshPos = nullptr;           // V x 3
    uint32_t* dMeshIdx = nullptr;        // 3 T
    void* dMeshTris = nullptr;           // 2 T screen-space records of the pass being rasterised
    bool envDeferred = false;            // an environment pass is owed to the colour target (see flush_deferred)
    float* dMeshNrm = nullptr;           // V x 3 (recomputed vertex normals)
    void* dMeshShade = nullptr;          // 2 T records of interpolants for the base pass
    unsigned long long* dMeshVis = nullptr;   // W x H visibility buffer: depth bits << 32 | record
    float meshWvpPrev[16] = {};          // last frame's world-view-projection (velocity)
    bool meshHavePrev = false;
    uint32_t* dShadowBits = nullptr;     // S x S float bit patterns (depth test target of the shadow pass)
    uint32_t meshNumIndices = 0;
    float meshExtent = 1.0f;             // largest AABB extent (ObjectRenderer.cpp:74-76)
    float meshScale = 1.0f, meshPos[3] = {0.0f, 0.0f, 0.0f};
    uint2* dDirectColor = nullptr;       // screen-space march results, directCapacity pixels
    uint2* dDirectStats = nullptr;       // allocated on first use with counters on
    uint32_t directCapacity = 0;
    uint2* dLightDense = nullptr;        // L^3 entries
    LightRec* dLightRecs = nullptr;      // L^3 entries
    uint4* dLightItems = nullptr;
    uint32_t* dLightSeg = nullptr;
    float* dLightItemResults = nullptr;
    uint32_t lightItemCapacity = 0;
    unsigned char* dBlock = nullptr;     // exchange block: arena | light staging | back buffer | flags
    mv_exchange_layout layout{};
    unsigned char* dArena = nullptr;     // = dBlock + layout.arena_offset
    size_t arenaBytes = 0;
    CubeArena arena{};
    uint2* dLightStaging = nullptr;      // inside the block: the staging buffer the current light march fills
    uint2* dLightStaging2[2] = {};       // the two staging buffers (frames pipelined across ranks alternate)
    uint32_t stagingParity = 0;
    uint32_t* dFlags = nullptr;          // inside the block: arrival counters, [channel][kMaxPeers] (kBarrierMain, kBarrierLight)
    uint32_t* hTimeout = nullptr;        // mapped pinned word a timed-out k_peer_wait sets (read by the host without a copy)
    uint32_t* dTimeout = nullptr;        // its device address
    uint32_t barrierSeqCh[2] = {0, 0};
    unsigned char* peerBlock[kMaxPeers] = {};
    uint32_t barrierSeq = 0;
    uint32_t** dPeerFlagPtrs = nullptr;  // [kMaxPeers] flag arrays of every rank (peer-mapped)
    bool peersMapped = false;
    cudaStream_t ownStream = nullptr;
    cudaEvent_t uploadDone[3] = {};
    bool uploadPending[3] = {};
    // Frame pipelining (one GPU, uninstrumented): cull + light march of a frame run on lightStream and write the light map
    // into the staging buffer, so they depend on nothing the previous frame's view march / resolve / post-process (main
    // stream) still use; the main stream commits the staging buffer into the light volume's array before its view march.
    // Per-frame device state is double-buffered: PerObject records by upload, lists + attributes by render.
    cudaStream_t lightStream = nullptr;
    cudaEvent_t lightDone = nullptr, commitDone = nullptr, inputsReady = nullptr, frameEnd[2] = {};
    bool lightDoneValid = false, commitValid = false, frameEndValid[2] = {false, false}, inputsDirty = true, lightToStaging = false;
    uint32_t cullSerial = 0;             // serial number of the last fused cull + view-march launch
    // Sharded frame: view-march CTAs per SM while the light march of the frame's light volume runs beside it on the light
    // stream; 0 = the passes run one after the other. -1 = by world size: 4 from 8 ranks on (1 / 8 of a frame per GPU is
    // latency-bound: cfg4 1181 -> 1274 frames/s), 0 below (at N = 2 the GPUs are still throughput-bound and the overlap
    // costs 0 - 14 %). MV_SHARD_V_BLOCKS overrides. profiles/r01_scaling.md
    int shardViewBlocks = -1;
    cudaEvent_t cullDone = nullptr;      // sharded frame: main stream -> light stream
    int overlapLight = 1;                // MV_OVERLAP=0: every pass on the main stream
    int shardPipeline = 1;               // MV_SHARD_PIPELINE=0: sharded frames are not pipelined across frames (mv_api.cu)
    PerObject* dPerObject2[2] = {};      // dPerObject points at the one of the last mv_update_frame
    ushort4* dAttribs2[2] = {};
    unsigned char* dLists2[2] = {};
    uint32_t poParity = 0, listParity = 0;
    int poLastUse[2] = {-1, -1};         // frameEnd slot of the last render that read dPerObject2[q]
    int lastUpload = -1;                 // upload-ring slot of the last PerObject upload
    cudaStream_t copyStream = nullptr;   // mv_present_async: back-buffer read-back overlapping the next frame
    cudaStream_t directStream = nullptr; // screen-space march beside the view march (directOverlap)
    cudaEvent_t directFork = nullptr, directJoin = nullptr;
    int directOverlap = 1;               // the two marches of a frame run on two streams (uninstrumented frames; MV_DIRECT_OVERLAP=0: one after the other)
    cudaEvent_t frameDone = nullptr;     // main stream -> copy stream
    cudaEvent_t presentDone[MV_PRESENT_SLOTS] = {};
    bool presentPending[MV_PRESENT_SLOTS] = {};
    bool backBufferBusyOwnRows = false;  // that copy reads only this rank's own rows (mv_present_rows_async)
    int backBufferBusy = -1;             // slot whose copy still reads the back buffer (device-side wait before it is rewritten)
    float* dDepth = nullptr;
    uint16_t* dShadow = nullptr;
    uint32_t shadowSize = 0;
    uint2* dColor = nullptr;             // colour RT
    uint2* dBackground = nullptr;        // copy of the colour RT given to set_targets
    uint32_t* dVelocity = nullptr;       // RG16F
    bool velocityGiven = false;          // a velocity field was passed to mv_set_targets (else it is all zero)
    uint2* dHistory[2] = {nullptr, nullptr};
    uint2* dEnvCube = nullptr;           // radiance cube map of the environment pass, RGBA16F [face][y][x] (k_env.cu)
    uint32_t envSize = 0;
    unsigned char* dToneLut = nullptr;   // [65536] PSToneMap + RGBA8 write of every half pattern (k_post.cu)
    uchar4* dBackBuffer = nullptr;
    uchar4* dPeerBackBuffer = nullptr;   // rank 0's back buffer (peer-mapped) in a multi-GPU run
    uint2* peerHistory[kMaxPeers][2] = {};   // the peers' TAA history images (peer-mapped; nullptr for this rank)
    float* dScratch = nullptr;           // SH projection partial sums
    size_t scratchBytes = 0;
    // sharding
    uint32_t shardRank = 0, shardWorld = 1;
    bool shardVolumes = false;           // mv_create_sharded
    uint32_t proxyGrid = 0;
    unsigned char* dSrcIsProxy = nullptr;
    bool owns_source(uint32_t src) const { return !shardVolumes || src % shardWorld == shardRank; }
    uint32_t row0 = 0, row1 = 0;
    uint32_t stripeH = 0;
    std::vector<void*> openedIpc;
    // timing
    cudaEvent_t ev[8] = {};
    bool evValid[8] = {};
    mv_timings lastTimings{};
    DeviceScene scene() const;
};

void set_error(const char* fmt, ...);

} // namespace mv
