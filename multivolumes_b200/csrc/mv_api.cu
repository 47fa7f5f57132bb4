// mv_api.cu — C-ABI of libmv_b200.so (include/mv.h): resource ownership, host-side scene maths and
// pass ordering of the reference's MultiRayCaster (MultiVolumes/Content/MultiRayCaster.cpp) over one
// CUDA stream. No CPU path: every entry point that computes launches sm_100a kernels.
#include "mv_internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

using namespace mv;

struct mv_caster { Caster c; };

namespace mv {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

#define MV_CUDA(expr)                                                                                     \
    do {                                                                                                  \
        const cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return MV_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

#define MV_REQUIRE(cond)                                                          \
    do {                                                                          \
        if (!(cond)) { set_error("invalid argument: %s", #cond); return MV_ERR_INVALID; } \
    } while (0)

DeviceScene Caster::scene() const
{
    DeviceScene s{};
    s.perObject = dPerObject;
    s.volumeDescs = dVolumeDescs;
    s.attribs = dAttribs;
    s.lists = reinterpret_cast<FrameLists*>(dLists);
    const uint32_t N = d.num_volumes;
    uint32_t* tail = reinterpret_cast<uint32_t*>(dLists + sizeof(FrameLists));
    s.visible = tail;
    s.cubeVolumes = tail + N;
    s.cubeTilePrefix = tail + 2 * N;
    s.directTilePrefix = tail + 3 * N + 1;
    s.directOffset = tail + 4 * N + 2;
    s.marchOrder = tail + 5 * N + 2;
    s.cubeTileBegin = tail + 6 * N + 2;
    s.visInfo = reinterpret_cast<VisInfo*>(dLists + frame_lists_header_bytes(N));
    s.directColor = dDirectColor;
    // (volume-sharded storage: a pixel's march may have run on another rank; the per-result counters stay where they were counted)
    s.directStats = ((d.flags & MV_FLAG_COUNT_SAMPLES) && !(shardVolumes && shardWorld > 1)) ? dDirectStats : nullptr;
    s.directCapacity = directCapacity;
    s.volumeTex = dVolumeTex;
    s.occ.bits = dOcc; s.occ.wordsPerVolume = occWords; s.occ.shift = occShift; s.occ.bricks = occBricks; s.occ.gridSize = (float)d.grid_size; s.occ.halfBricks = 0.5f * (float)d.grid_size / (float)(1u << occShift);
    s.lightTex = dLightTex;
    s.lightSurf = dLightSurf;
    s.depth = dDepth;
    s.shadow = dShadow;
    s.color = dColor;
    s.lightDense = dLightDense;
    s.lightRecs = dLightRecs;
    s.lightItems = dLightItems;
    s.lightItemResults = dLightItemResults;
    s.lightSeg = dLightSeg;
    s.lightItemCapacity = lightItemCapacity;
    s.stats = (d.flags & MV_FLAG_COUNT_SAMPLES) ? dStats : nullptr;
    s.arena = arena;
    s.shardRank = shardRank; s.shardWorld = shardWorld;
    s.shardVolumes = shardVolumes ? 1u : 0u;
    s.srcIsProxy = shardVolumes ? dSrcIsProxy : nullptr;
    for (int p = 0; p < kMaxPeers; ++p)
        s.directPeer[p] = (shardVolumes && peersMapped && (uint32_t)p < shardWorld && (uint32_t)p != shardRank && peerBlock[p])
                              ? reinterpret_cast<uint2*>(peerBlock[p] + layout.direct_offset) : nullptr;
    s.row0 = row0; s.row1 = row1;
    s.stripeH = (shardWorld > 1) ? stripeH : 0;
    return s;
}

// ---- host matrix algebra (DirectXMath call sites: MultiRayCaster.cpp:325-350) ----
// Products and inverses are evaluated in double and rounded once to fp32.
static void mul44(const float* A, const float* B, float* R)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += (double)A[i * 4 + k] * (double)B[k * 4 + j];
            R[i * 4 + j] = (float)s;
        }
}

static void inverse44(const float* A, float* R)   // adjugate / determinant
{
    double a[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = A[i * 4 + j];
    auto minor3 = [&](const int r[3], const int c[3]) {
        return a[r[0]][c[0]] * (a[r[1]][c[1]] * a[r[2]][c[2]] - a[r[1]][c[2]] * a[r[2]][c[1]])
             - a[r[0]][c[1]] * (a[r[1]][c[0]] * a[r[2]][c[2]] - a[r[1]][c[2]] * a[r[2]][c[0]])
             + a[r[0]][c[2]] * (a[r[1]][c[0]] * a[r[2]][c[1]] - a[r[1]][c[1]] * a[r[2]][c[0]]);
    };
    double cof[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int r[3], c[3], ri = 0, ci = 0;
            for (int k = 0; k < 4; ++k) { if (k != i) r[ri++] = k; if (k != j) c[ci++] = k; }
            const double m = minor3(r, c);
            cof[i][j] = ((i + j) & 1) ? -m : m;
        }
    const double det = a[0][0] * cof[0][0] + a[0][1] * cof[0][1] + a[0][2] * cof[0][2] + a[0][3] * cof[0][3];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R[i * 4 + j] = (float)(cof[j][i] / det);
}

static void world_from43(const float* W, float* M)
{
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 3; ++j) M[i * 4 + j] = W[i * 3 + j]; M[i * 4 + 3] = (i == 3) ? 1.0f : 0.0f; }
}
static void to43(const float* M, float* W)
{
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) W[i * 3 + j] = M[i * 4 + j];
}

static void set_volume_world(Caster& c, uint32_t i, float size, const float pos[3])   // MultiRayCaster.cpp:297-303
{
    size *= 0.5f;
    float* w = &c.volumeWorlds[(size_t)i * 12];
    const float m[12] = {size, 0, 0, 0, size, 0, 0, 0, size, pos[0], pos[1], pos[2]};
    memcpy(w, m, sizeof m);
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

static int make_volume3d(Volume3D& v, uint32_t n, uint32_t channels = 4)
{
    v.channels = channels;
    v.edge = n;
    const cudaChannelFormatDesc cd = channels == 1 ? cudaCreateChannelDescHalf() : cudaCreateChannelDescHalf4();
    MV_CUDA(cudaMalloc3DArray(&v.array, &cd, make_cudaExtent(n, n, n), cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = v.array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;   // LINEAR_CLAMP
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    MV_CUDA(cudaCreateTextureObject(&v.tex, &rd, &td, nullptr));
    MV_CUDA(cudaCreateSurfaceObject(&v.surf, &rd));
    return MV_OK;
}

static int clear_volume3d(Caster& c, Volume3D& v, uint32_t n)
{
    // zero-fill through a device staging row-block (cudaMemset3D does not take arrays)
    const size_t texel = 2 * (size_t)v.channels;
    const size_t bytes = (size_t)n * n * n * texel;
    void* z = nullptr;
    MV_CUDA(cudaMalloc(&z, bytes));
    MV_CUDA(cudaMemsetAsync(z, 0, bytes, c.stream));
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(z, (size_t)n * texel, n, n);
    p.dstArray = v.array;
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyDeviceToDevice;
    MV_CUDA(cudaMemcpy3DAsync(&p, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    MV_CUDA(cudaFree(z));
    return MV_OK;
}

static void record(Caster& c, int i)
{
    if (c.d.flags & MV_FLAG_TIME_PASSES) { cudaEventRecord(c.ev[i], c.stream); c.evValid[i] = true; }
}

// The back buffer may still be read by an mv_present_async copy: order the next writer after it.
// `beforePeersWrite`: the call sits in front of the barrier after which PEERS may store rows into this rank's back buffer. A copy
// that reads only this rank's own rows (mv_present_rows_async) cannot be overtaken by those — only by this rank's own next
// post-process, whose call (beforePeersWrite = false) still waits — so the early wait, which would stall the frame's marches
// behind the previous frame's read-back, is skipped for it.
static void wait_back_buffer_free(Caster& c, bool beforePeersWrite = false)
{
    if (c.backBufferBusy < 0) return;
    if (beforePeersWrite && c.backBufferBusyOwnRows) return;
    cudaStreamWaitEvent(c.stream, c.presentDone[c.backBufferBusy], 0);
    c.backBufferBusy = -1;
}

// the per-result counters of the screen-space marches exist only while the sample counters are on
#define MV_TRY_DIRECT_STATS(c)                                                                                   \
    do {                                                                                                        \
        if (((c).d.flags & MV_FLAG_COUNT_SAMPLES) && !(c).dDirectStats)                                         \
            MV_CUDA(cudaMalloc(&(c).dDirectStats, std::max<size_t>((c).directCapacity, 1) * sizeof(uint2)));    \
    } while (0)

// Frames are pipelined (cull + light march of frame i + 1 on the light stream beside frame i's view march / resolve /
// post-process on the main stream) on one GPU and, with the peers mapped, across the ranks of a sharded frame.
static bool pipelined_sharded(const Caster& c)
{
    return c.overlapLight && c.shardPipeline && c.shardWorld > 1 && !c.shardVolumes && c.peersMapped && !(c.d.flags & (MV_FLAG_COUNT_SAMPLES | MV_FLAG_TIME_PASSES));
}
bool frame_is_pipelined_on_one_gpu(const Caster& c);
static bool pipelined(const Caster& c)
{
    return (c.overlapLight && (c.shardWorld == 1 || c.shardVolumes) && !(c.d.flags & (MV_FLAG_COUNT_SAMPLES | MV_FLAG_TIME_PASSES))) || pipelined_sharded(c);
}

bool frame_is_pipelined_on_one_gpu(const Caster& c) { return c.shardWorld == 1 && !c.shardVolumes && c.directOverlap && pipelined(c); }

// A new frame's lists and attributes go into the other buffer (the previous frame's resolve may still read its own)
static void flip_frame_lists(Caster& c)
{
    c.listParity ^= 1u;
    c.dLists = c.dLists2[c.listParity];
    c.dAttribs = c.dAttribs2[c.listParity];
}

// The PerObject records may have been uploaded on the other stream
static void wait_upload(Caster& c, cudaStream_t s)
{
    if (c.lastUpload >= 0) cudaStreamWaitEvent(s, c.uploadDone[c.lastUpload], 0);
}

// View march and screen-space march of a frame are independent of each other (both read the light maps; one writes the cube
// maps, the other its result buffer). Sharded over many ranks both are short, latency-bound launches: with directOverlap the
// screen-space march runs on its own stream beside the view march (uninstrumented frames only).
static void launch_view_and_direct(Caster& c)
{
    if (!c.directOverlap) { flush_deferred(c); launch_ray_march_view(c); launch_ray_cast_direct(c); return; }
    cudaStream_t mainStream = c.stream;
    cudaEventRecord(c.directFork, mainStream);
    cudaStreamWaitEvent(c.directStream, c.directFork, 0);
    c.stream = c.directStream;
    // an owed environment pass (ALU-bound, writes the colour target the resolve reads after the join) runs here, beside the
    // texture-bound view march
    if (c.envDeferred) { c.envDeferred = false; launch_environment(c, true); }
    launch_ray_cast_direct(c);
    c.stream = mainStream;
    cudaEventRecord(c.directJoin, c.directStream);
    launch_ray_march_view(c);
    cudaStreamWaitEvent(mainStream, c.directJoin, 0);
}

static int check_launch(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s launch failed: %s", what, cudaGetErrorString(e)); return MV_ERR_CUDA; }
    return MV_OK;
}

static void kill_volume3d(Volume3D& v)
{
    if (v.tex) cudaDestroyTextureObject(v.tex);
    if (v.surf) cudaDestroySurfaceObject(v.surf);
    if (v.array) cudaFreeArray(v.array);
    v.tex = 0; v.surf = 0; v.array = nullptr;
}

int begin_ingest(Caster& c, uint32_t src, IngestTarget& t)
{
    Volume3D& own = c.volumes[src];
    if (!own.proxy) { t.vol = &own; t.temporary = false; return MV_OK; }
    // this rank keeps only the proxy of the source: the full-resolution data passes through a temporary texture
    t.vol = new (std::nothrow) Volume3D();
    if (!t.vol) { set_error("out of host memory"); return MV_ERR_NOMEM; }
    t.temporary = true;
    const int rc = make_volume3d(*t.vol, c.d.grid_size, (c.d.flags & MV_FLAG_DENSITY_ONLY) ? 1u : 4u);
    if (rc != MV_OK) { kill_volume3d(*t.vol); delete t.vol; t.vol = nullptr; }
    return rc;
}

int end_ingest(Caster& c, uint32_t src, IngestTarget& t)
{
    if (!t.vol) return MV_OK;
    if (!t.temporary) { launch_build_occupancy(c, src); return check_launch("k_build_occupancy"); }
    launch_build_proxy(c, *t.vol, c.volumes[src]);
    int rc = check_launch("k_build_proxy");
    if (cudaStreamSynchronize(c.stream) != cudaSuccess && rc == MV_OK) { set_error("proxy build failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MV_ERR_CUDA; }
    kill_volume3d(*t.vol);
    delete t.vol; t.vol = nullptr;
    return rc;
}

static void destroy_caster(Caster& c)
{
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    for (void* p : c.openedIpc) cudaIpcCloseMemHandle(p);
    for (auto& v : c.volumes) kill_volume3d(v);
    for (auto& v : c.lightMaps) kill_volume3d(v);
    void* frees[] = {c.dVolumeTex, c.dLightTex, c.dLightSurf, c.dPerObject2[0], c.dPerObject2[1], c.dVolumeDescs, c.dAttribs2[0], c.dAttribs2[1],
                     c.dLists2[0], c.dLists2[1], c.dStats, c.dMeshPos, c.dMeshIdx, c.dMeshTris, c.dMeshNrm, c.dMeshShade, c.dMeshVis, c.dShadowBits, c.dDirectStats, c.dLightDense, c.dLightRecs, c.dLightItems, c.dLightItemResults, c.dLightSeg, c.dBlock,
                     c.dOcc, c.dDepth, c.dShadow, c.dColor, c.dBackground, c.dVelocity, c.dScratch, c.dPeerFlagPtrs, c.dToneLut, c.dSrcIsProxy, c.dEnvCube};
    for (void* p : frees) if (p) cudaFree(p);
    if (c.hPerObjectPinned) cudaFreeHost(c.hPerObjectPinned);
    if (c.hTimeout) cudaFreeHost(c.hTimeout);
    for (auto& e : c.ev) if (e) cudaEventDestroy(e);
    for (auto& e : c.uploadDone) if (e) cudaEventDestroy(e);
    for (auto& e : c.presentDone) if (e) cudaEventDestroy(e);
    if (c.frameDone) cudaEventDestroy(c.frameDone);
    if (c.copyStream) { cudaStreamSynchronize(c.copyStream); cudaStreamDestroy(c.copyStream); }
    if (c.directStream) { cudaStreamSynchronize(c.directStream); cudaStreamDestroy(c.directStream); }
    for (cudaEvent_t e : {c.directFork, c.directJoin}) if (e) cudaEventDestroy(e);
    if (c.lightStream) { cudaStreamSynchronize(c.lightStream); cudaStreamDestroy(c.lightStream); }
    for (cudaEvent_t e : {c.lightDone, c.commitDone, c.inputsReady, c.cullDone, c.frameEnd[0], c.frameEnd[1]}) if (e) cudaEventDestroy(e);
    c.dPerObject = nullptr; c.dAttribs = nullptr; c.dLists = nullptr;
    if (c.ownStream) cudaStreamDestroy(c.ownStream);
}

} // namespace mv

extern "C" {

const char* mv_last_error(void) { return g_error; }
uint32_t mv_abi_version(void) { return 1; }

static int create_impl(const mv_desc* d, bool shardVolumes, uint32_t shardRank, uint32_t shardWorld, uint32_t proxyGrid, mv_caster** out)
{
    if (out) *out = nullptr;
    MV_REQUIRE(d && out);
    MV_REQUIRE(shardVolumes || !(d->flags & MV_FLAG_SHARD_VOLUMES));
    if (shardVolumes) MV_REQUIRE(shardWorld >= 1 && shardWorld <= (uint32_t)kMaxPeers && shardRank < shardWorld && proxyGrid >= 2 && d->grid_size % proxyGrid == 0);
    MV_REQUIRE(d->grid_size != 0 && d->num_volumes != 0 && d->num_volume_srcs != 0 && d->width != 0 && d->height != 0);
    MV_REQUIRE(d->grid_size < (1u << 14) && d->num_volume_srcs < (1u << 14) && (d->grid_size >> (kNumCubeMip - 1)) != 0);
    MV_REQUIRE(d->num_volumes < (1u << 24));
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libmv_b200 has no CPU path");
        return MV_ERR_NO_DEVICE;
    }
    MV_REQUIRE((int)d->device < ndev);
    cudaDeviceProp prop{};
    MV_CUDA(cudaGetDeviceProperties(&prop, (int)d->device));
    if (prop.major != 10) {
        set_error("device %u is sm_%d%d; libmv_b200 is built for sm_100a only", d->device, prop.major, prop.minor);
        return MV_ERR_NO_DEVICE;
    }
    mv_caster* h = new (std::nothrow) mv_caster();
    if (!h) { set_error("out of host memory"); return MV_ERR_NOMEM; }
    Caster& c = h->c;
    c.d = *d;
    if (c.d.light_grid_size == 0) c.d.light_grid_size = 96;
    if (c.d.max_ray_samples == 0) c.d.max_ray_samples = 256;
    if (c.d.max_light_samples == 0) c.d.max_light_samples = 96;
    c.device = (int)d->device;
    c.smCount = prop.multiProcessorCount;
    if (shardVolumes) { c.shardVolumes = true; c.shardRank = shardRank; c.shardWorld = shardWorld; c.proxyGrid = proxyGrid; c.d.flags |= MV_FLAG_SHARD_VOLUMES; }
    const uint32_t G = c.d.grid_size, L = c.d.light_grid_size, N = c.d.num_volumes, S = c.d.num_volume_srcs;
    const size_t px = (size_t)c.d.width * c.d.height;
    c.row0 = 0; c.row1 = c.d.height;

    auto fail = [&](int rc) { destroy_caster(c); delete h; return rc; };
#define MV_TRY(expr) do { const int rc_ = (expr); if (rc_ != MV_OK) return fail(rc_); } while (0)
#define MV_CUDA_C(expr)                                                                                   \
    do {                                                                                                  \
        const cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return fail(e_ == cudaErrorMemoryAllocation ? MV_ERR_NOMEM : MV_ERR_CUDA);                    \
        }                                                                                                 \
    } while (0)

    MV_CUDA_C(cudaSetDevice(c.device));
    MV_CUDA_C(cudaStreamCreateWithFlags(&c.ownStream, cudaStreamNonBlocking));
    c.stream = c.ownStream;
    for (auto& e : c.ev) MV_CUDA_C(cudaEventCreate(&e));
    MV_CUDA_C(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        MV_CUDA_C(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MV_CUDA_C(cudaStreamCreateWithPriority(&c.lightStream, cudaStreamNonBlocking, hi));
        for (cudaEvent_t* e : {&c.lightDone, &c.commitDone, &c.inputsReady, &c.frameEnd[0], &c.frameEnd[1]})
            MV_CUDA_C(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        if (const char* e = getenv("MV_OVERLAP")) c.overlapLight = atoi(e);
        if (const char* e = getenv("MV_SHARD_V_BLOCKS")) c.shardViewBlocks = atoi(e);
        if (const char* e = getenv("MV_SHARD_PIPELINE")) c.shardPipeline = atoi(e);
        MV_CUDA_C(cudaEventCreateWithFlags(&c.cullDone, cudaEventDisableTiming));
    }
    MV_CUDA_C(cudaEventCreateWithFlags(&c.frameDone, cudaEventDisableTiming));
    MV_CUDA_C(cudaStreamCreateWithFlags(&c.directStream, cudaStreamNonBlocking));
    MV_CUDA_C(cudaEventCreateWithFlags(&c.directFork, cudaEventDisableTiming));
    MV_CUDA_C(cudaEventCreateWithFlags(&c.directJoin, cudaEventDisableTiming));
    if (const char* e = getenv("MV_DIRECT_OVERLAP")) c.directOverlap = atoi(e);
    for (auto& e : c.presentDone) MV_CUDA_C(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    // MultiRayCaster.cpp:99-126 — per-source volumes, per-instance light maps and cube maps
    c.volumes.resize(S);
    {
        std::vector<unsigned char> isProxy(S, 0);
        for (uint32_t i = 0; i < S; ++i) {
            Volume3D& v = c.volumes[i];
            if (c.owns_source(i)) { MV_TRY(make_volume3d(v, G, (c.d.flags & MV_FLAG_DENSITY_ONLY) ? 1u : 4u)); MV_TRY(clear_volume3d(c, v, G)); }
            else { MV_TRY(make_volume3d(v, c.proxyGrid, 1u)); v.proxy = true; isProxy[i] = 1; MV_TRY(clear_volume3d(c, v, c.proxyGrid)); }
        }
        if (c.shardVolumes) {
            MV_CUDA_C(cudaMalloc(&c.dSrcIsProxy, S));
            MV_CUDA_C(cudaMemcpy(c.dSrcIsProxy, isProxy.data(), S, cudaMemcpyHostToDevice));
        }
    }
    {
        // empty-space bricks: 64 per axis (32 KB of bits per source volume) for volumes of 256^3 and up; measured on B200
        // (profiles/r02_notes.md): 43 % of cfg4's march samples need no fetch, view march -12 %, screen-space march -10 %.
        // Below 256^3 the volumes are cache-friendly enough that the lookup costs more than the fetch it saves (cfg2: +5 %),
        // so the bricks are off there. MV_OCC_BRICKS overrides (0 = off).
        uint32_t perAxis = G >= 256 ? 64 : 0;
        if (const char* e = getenv("MV_OCC_BRICKS")) perAxis = (uint32_t)strtoul(e, nullptr, 10);
        if (perAxis) {
            uint32_t shift = 1;
            while (((G - 1) >> shift) + 1 > perAxis) ++shift;
            c.occShift = shift; c.occBricks = ((G - 1) >> shift) + 1;
            c.occWords = (c.occBricks * c.occBricks * c.occBricks + 31) / 32;
            MV_CUDA_C(cudaMalloc(&c.dOcc, (size_t)S * c.occWords * sizeof(uint32_t)));
            MV_CUDA_C(cudaMemsetAsync(c.dOcc, 0, (size_t)S * c.occWords * sizeof(uint32_t), c.stream));   // nothing known to be empty yet
        }
    }
    c.lightMaps.resize(N);     // volume-sharded storage: only of the instances whose source this rank holds (only their marches read them)
    for (uint32_t i = 0; i < N; ++i) if (c.owns_source(i % S)) { MV_TRY(make_volume3d(c.lightMaps[i], L)); MV_TRY(clear_volume3d(c, c.lightMaps[i], L)); }
    std::vector<cudaTextureObject_t> vt(S), lt(N);
    std::vector<cudaSurfaceObject_t> ls(N);
    for (uint32_t i = 0; i < S; ++i) vt[i] = c.volumes[i].tex;
    for (uint32_t i = 0; i < N; ++i) { lt[i] = c.lightMaps[i].tex; ls[i] = c.lightMaps[i].surf; }
    MV_CUDA_C(cudaMalloc(&c.dVolumeTex, S * sizeof(cudaTextureObject_t)));
    MV_CUDA_C(cudaMalloc(&c.dLightTex, N * sizeof(cudaTextureObject_t)));
    MV_CUDA_C(cudaMalloc(&c.dLightSurf, N * sizeof(cudaSurfaceObject_t)));
    MV_CUDA_C(cudaMemcpy(c.dVolumeTex, vt.data(), S * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));
    MV_CUDA_C(cudaMemcpy(c.dLightTex, lt.data(), N * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));
    MV_CUDA_C(cudaMemcpy(c.dLightSurf, ls.data(), N * sizeof(cudaSurfaceObject_t), cudaMemcpyHostToDevice));

    // cube-map arena: colour region then depth region, volume-major, [mip][face][y][x] inside a slot
    uint32_t texels = 0;
    for (uint32_t m = 0; m < kNumCubeMip; ++m) { c.arena.mipTexelOffset[m] = texels; texels += 6u * (G >> m) * (G >> m); }
    c.arena.mipTexelOffset[kNumCubeMip] = texels;
    c.arena.colorStride = (unsigned long long)texels * 8ull;
    c.arena.depthStride = (unsigned long long)texels * 4ull;
    c.arena.depthBase = c.arena.colorStride * N;
    c.arenaBytes = (size_t)(c.arena.depthBase + c.arena.depthStride * N);
    // exchange block (include/mv.h): arena | light staging | back buffer | flags, 256-B aligned parts
    auto align256 = [](uint64_t v) { return (v + 255ull) & ~255ull; };
    mv_exchange_layout& lay = c.layout;
    lay.arena_offset = 0; lay.arena_bytes = c.arenaBytes;
    lay.light_staging_offset = align256(lay.arena_offset + lay.arena_bytes);
    lay.light_staging_bytes = (uint64_t)(L + kMaxPeers) * L * L * 8ull;
    lay.back_buffer_offset = align256(lay.light_staging_offset + lay.light_staging_bytes);
    lay.back_buffer_bytes = (uint64_t)px * 4ull;
    lay.flags_offset = align256(lay.back_buffer_offset + lay.back_buffer_bytes);
    lay.flags_bytes = 256;
    lay.light_staging2_offset = align256(lay.flags_offset + lay.flags_bytes);
    // results of the screen-space marches (RayCast of the direct-scheme volumes), rectangle by rectangle: room for four
    // full-screen rectangles; volumes beyond that are marched inside the resolve kernel
    c.directCapacity = (uint32_t)std::min<size_t>(4 * px, 0x7fffffffu);
    if (const char* cap = getenv("MV_DIRECT_CAPACITY")) c.directCapacity = (uint32_t)strtoul(cap, nullptr, 10);   // tests: force the fallback
    lay.direct_offset = align256(lay.light_staging2_offset + lay.light_staging_bytes);
    lay.direct_bytes = std::max<uint64_t>(c.directCapacity, 1) * 8ull;
    lay.history_bytes = (uint64_t)px * 8ull;
    lay.history_offset[0] = align256(lay.direct_offset + lay.direct_bytes);
    lay.history_offset[1] = align256(lay.history_offset[0] + lay.history_bytes);
    lay.block_bytes = lay.history_offset[1] + lay.history_bytes;
    lay.light_slab_depth = L;
    MV_CUDA_C(cudaMalloc(&c.dBlock, lay.block_bytes));
    MV_CUDA_C(cudaMemsetAsync(c.dBlock, 0, lay.block_bytes, c.stream));
    c.dArena = c.dBlock + lay.arena_offset;
    c.dLightStaging2[0] = reinterpret_cast<uint2*>(c.dBlock + lay.light_staging_offset);
    c.dLightStaging2[1] = reinterpret_cast<uint2*>(c.dBlock + lay.light_staging2_offset);
    c.dLightStaging = c.dLightStaging2[0];
    c.dDirectColor = reinterpret_cast<uint2*>(c.dBlock + lay.direct_offset);
    c.dHistory[0] = reinterpret_cast<uint2*>(c.dBlock + lay.history_offset[0]);
    c.dHistory[1] = reinterpret_cast<uint2*>(c.dBlock + lay.history_offset[1]);
    MV_CUDA_C(cudaHostAlloc(&c.hTimeout, sizeof(uint32_t), cudaHostAllocMapped));
    *c.hTimeout = 0;
    MV_CUDA_C(cudaHostGetDevicePointer(&c.dTimeout, c.hTimeout, 0));
    c.dBackBuffer = reinterpret_cast<uchar4*>(c.dBlock + lay.back_buffer_offset);
    c.dFlags = reinterpret_cast<uint32_t*>(c.dBlock + lay.flags_offset);
    c.arena.base = c.dArena;
    c.arena.numPeers = 0;

    // createVolumeInfoBuffers, MultiRayCaster.cpp:455-549
    for (int q = 0; q < 2; ++q) {
        MV_CUDA_C(cudaMalloc(&c.dPerObject2[q], N * sizeof(PerObject)));
        MV_CUDA_C(cudaMemsetAsync(c.dPerObject2[q], 0, N * sizeof(PerObject), c.stream));
    }
    c.dPerObject = c.dPerObject2[0];
    MV_CUDA_C(cudaMallocHost(&c.hPerObjectPinned, (size_t)kUploadRing * N * sizeof(PerObject)));
    for (auto& e : c.uploadDone) MV_CUDA_C(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    std::vector<uint32_t> descs(N);
    for (uint32_t i = 0; i < N; ++i) descs[i] = (i % S) | (kNumCubeMip << 14) | (G << 18);   // :470-479
    MV_CUDA_C(cudaMalloc(&c.dVolumeDescs, N * sizeof(uint32_t)));
    MV_CUDA_C(cudaMemcpy(c.dVolumeDescs, descs.data(), N * sizeof(uint32_t), cudaMemcpyHostToDevice));
    for (int q = 0; q < 2; ++q) {
        MV_CUDA_C(cudaMalloc(&c.dAttribs2[q], N * sizeof(ushort4)));
        MV_CUDA_C(cudaMemsetAsync(c.dAttribs2[q], 0, N * sizeof(ushort4), c.stream));
    }
    c.dAttribs = c.dAttribs2[0];
    const size_t listBytes = frame_lists_header_bytes(N) + (size_t)N * sizeof(VisInfo);
    for (int q = 0; q < 2; ++q) {
        MV_CUDA_C(cudaMalloc(&c.dLists2[q], listBytes));
        MV_CUDA_C(cudaMemsetAsync(c.dLists2[q], 0, listBytes, c.stream));
    }
    c.dLists = c.dLists2[0];
    if (c.d.flags & MV_FLAG_COUNT_SAMPLES) MV_CUDA_C(cudaMalloc(&c.dDirectStats, std::max<size_t>(c.directCapacity, 1) * sizeof(uint2)));   // else on first use
    MV_CUDA_C(cudaMalloc(&c.dLightDense, (size_t)L * L * L * sizeof(uint2)));
    MV_CUDA_C(cudaMalloc(&c.dLightRecs, (size_t)L * L * L * sizeof(LightRec)));
    c.lightItemCapacity = 4u * L * L * L + 32u * N;   // deferred AO rays (4 per voxel + segment padding); a frame that needs more marches them inline
    if (const char* cap = getenv("MV_LIGHT_ITEM_CAPACITY")) c.lightItemCapacity = (uint32_t)strtoul(cap, nullptr, 10) + 32u * N;   // tests: force the inline fallback
    MV_CUDA_C(cudaMalloc(&c.dLightItems, (size_t)c.lightItemCapacity * sizeof(uint4)));
    MV_CUDA_C(cudaMalloc(&c.dLightSeg, 2 * (size_t)N * sizeof(uint32_t)));
    MV_CUDA_C(cudaMemsetAsync(c.dLightSeg, 0, 2 * (size_t)N * sizeof(uint32_t), c.stream));
    MV_CUDA_C(cudaMalloc(&c.dLightItemResults, (size_t)c.lightItemCapacity * sizeof(float)));
    MV_CUDA_C(cudaMalloc(&c.dStats, sizeof(StatsDev)));
    MV_CUDA_C(cudaMemsetAsync(c.dStats, 0, sizeof(StatsDev), c.stream));

    // borrowed targets are copied in (SetRenderTargets), plus the post-process images
    MV_CUDA_C(cudaMalloc(&c.dDepth, px * sizeof(float)));
    MV_CUDA_C(cudaMalloc(&c.dColor, px * 8));
    MV_CUDA_C(cudaMalloc(&c.dBackground, px * 8));
    MV_CUDA_C(cudaMalloc(&c.dVelocity, px * 4));
    {
        std::vector<float> ones(px, 1.0f);
        MV_CUDA_C(cudaMemcpy(c.dDepth, ones.data(), px * sizeof(float), cudaMemcpyHostToDevice));
    }
    MV_CUDA_C(cudaMemsetAsync(c.dColor, 0, px * 8, c.stream));
    MV_CUDA_C(cudaMemsetAsync(c.dBackground, 0, px * 8, c.stream));
    MV_CUDA_C(cudaMemsetAsync(c.dVelocity, 0, px * 4, c.stream));
    MV_CUDA_C(cudaMalloc(&c.dToneLut, 65536));
    build_tone_lut(c);
    c.scratchBytes = (size_t)c.smCount * 4 * 8 * 28 * sizeof(float);
    MV_CUDA_C(cudaMalloc(&c.dScratch, c.scratchBytes));

    c.volumeWorlds.assign((size_t)N * 12, 0.0f);
    for (uint32_t i = 0; i < N; ++i) { float* w = &c.volumeWorlds[(size_t)i * 12]; w[0] = w[4] = w[8] = 1.0f; }
    const float center[3] = {0, 0, 0};
    set_volumes_world(c, 20.0f, center);                    // MultiRayCaster.cpp:143-144

    memset(&c.cb, 0, sizeof c.cb);
    c.cb.numVolumes = N; c.cb.gridSize = G; c.cb.lightGridSize = L;
    c.cb.width = c.d.width; c.cb.height = c.d.height;
    c.cb.maxRaySamples = c.d.max_ray_samples; c.cb.maxLightSamples = c.d.max_light_samples;
    c.cb.viewport[0] = (float)c.d.width; c.cb.viewport[1] = (float)c.d.height;
    c.cb.inv2Viewport[0] = 2.0f / c.cb.viewport[0]; c.cb.inv2Viewport[1] = 2.0f / c.cb.viewport[1];
    MV_CUDA_C(cudaStreamSynchronize(c.stream));
#undef MV_TRY
#undef MV_CUDA_C
    *out = h;
    return MV_OK;
}

// (no C++ exception may cross the C boundary: the host-side containers of the entry points that size them from their
// arguments are guarded, and the failure is reported as MV_ERR_NOMEM)
int mv_create(const mv_desc* d, mv_caster** out)
try { return create_impl(d, false, 0, 1, 0, out); }
catch (const std::bad_alloc&) { set_error("out of host memory"); if (out) *out = nullptr; return MV_ERR_NOMEM; }

int mv_create_sharded(const mv_desc* d, uint32_t rank, uint32_t world, uint32_t proxyGrid, mv_caster** out)
try { return create_impl(d, true, rank, world, proxyGrid, out); }
catch (const std::bad_alloc&) { set_error("out of host memory"); if (out) *out = nullptr; return MV_ERR_NOMEM; }

void mv_destroy(mv_caster* h)
{
    if (!h) return;
    destroy_caster(h->c);
    delete h;
}

#define MV_ENTER(h)                                  \
    MV_REQUIRE(h != nullptr);                        \
    Caster& c = h->c;                                \
    MV_CUDA(cudaSetDevice(c.device));                \
    flush_deferred(c)
#define MV_ENTER_KEEP(h)                             \
    MV_REQUIRE(h != nullptr);                        \
    Caster& c = h->c;                                \
    MV_CUDA(cudaSetDevice(c.device))

int mv_volume_init_procedural(mv_caster* h, uint32_t src, uint32_t mode, uint32_t seed)
{
    MV_ENTER(h);
    c.inputsDirty = true;
    MV_REQUIRE(src < c.d.num_volume_srcs && mode <= 1);
    IngestTarget t;
    int rc = begin_ingest(c, src, t);
    if (rc != MV_OK) return rc;
    launch_init_grid(c, *t.vol, mode, seed);
    rc = check_launch("k_init_grid");
    const int rc2 = end_ingest(c, src, t);
    return rc != MV_OK ? rc : rc2;
}

int mv_volume_upload_rgba16f(mv_caster* h, uint32_t src, const uint16_t* texels)
try {
    MV_ENTER(h);
    c.inputsDirty = true;
    MV_REQUIRE(texels && src < c.d.num_volume_srcs);
    const uint32_t n = c.d.grid_size;
    IngestTarget t;
    int rc = begin_ingest(c, src, t);
    if (rc != MV_OK) return rc;
    cudaMemcpy3DParms p{};
    std::vector<uint16_t> alpha;
    if (t.vol->channels == 1) {   // density-only storage keeps the alpha channel
        const size_t count = (size_t)n * n * n;
        alpha.resize(count);
        for (size_t i = 0; i < count; ++i) alpha[i] = texels[4 * i + 3];
        p.srcPtr = make_cudaPitchedPtr(alpha.data(), (size_t)n * 2, n, n);
    } else p.srcPtr = make_cudaPitchedPtr((void*)texels, (size_t)n * 8, n, n);
    p.dstArray = t.vol->array;
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyHostToDevice;
    const cudaError_t e = cudaMemcpy3DAsync(&p, c.stream);
    rc = end_ingest(c, src, t);
    if (e != cudaSuccess) { set_error("volume upload failed: %s", cudaGetErrorString(e)); return MV_ERR_CUDA; }
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return rc;
} catch (const std::bad_alloc&) { set_error("out of host memory"); return MV_ERR_NOMEM; }

int mv_volume_upload_r32f(mv_caster* h, uint32_t src, const float* density)
{
    MV_ENTER(h);
    c.inputsDirty = true;
    MV_REQUIRE(density && src < c.d.num_volume_srcs);
    const uint32_t n = c.d.grid_size;
    const size_t bytes = (size_t)n * n * n * sizeof(float);
    float* dtmp = nullptr;
    MV_CUDA(cudaMalloc(&dtmp, bytes));
    cudaError_t e = cudaMemcpyAsync(dtmp, density, bytes, cudaMemcpyHostToDevice, c.stream);
    int rc = MV_OK;
    if (e == cudaSuccess) {
        IngestTarget t;
        rc = begin_ingest(c, src, t);
        if (rc == MV_OK) { launch_r32f_to_rgba16f(c, *t.vol, dtmp); e = cudaGetLastError(); rc = end_ingest(c, src, t); }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFree(dtmp);
    if (e != cudaSuccess) { set_error("volume_upload_r32f: %s", cudaGetErrorString(e)); return MV_ERR_CUDA; }
    return rc;
}

int mv_volume_read(mv_caster* h, uint32_t src, uint16_t* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out && src < c.d.num_volume_srcs);
    if (c.volumes[src].proxy) { set_error("source %u lives on rank %u: this rank holds its density proxy only", src, src % c.shardWorld); return MV_ERR_INVALID; }
    const uint32_t n = c.d.grid_size;
    cudaMemcpy3DParms p{};
    p.srcArray = c.volumes[src].array;
    const bool densityOnly = c.volumes[src].channels == 1;
    const size_t count = (size_t)n * n * n;
    // density-only storage: the halves land in the tail of the caller's buffer and are expanded in place, front to back
    uint16_t* packed = densityOnly ? out + 3 * count : out;
    p.dstPtr = make_cudaPitchedPtr(packed, (size_t)n * (densityOnly ? 2 : 8), n, n);
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyDeviceToHost;
    MV_CUDA(cudaMemcpy3DAsync(&p, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    if (densityOnly)
        for (size_t i = 0; i < count; ++i) {
            const uint16_t a = packed[i];
            out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = 0x3c00u;   // 1.0
            out[4 * i + 3] = a;
        }
    return MV_OK;
}

static int set_targets_impl(Caster& c, const float* depth, const uint16_t* shadow, uint32_t shadowSize, const uint16_t* color,
                            const uint16_t* velocity, cudaMemcpyKind kind)
try {
    c.inputsDirty = true;
    const size_t px = (size_t)c.d.width * c.d.height;
    if (depth) MV_CUDA(cudaMemcpyAsync(c.dDepth, depth, px * sizeof(float), kind, c.stream));
    else {
        std::vector<float> ones(px, 1.0f);
        MV_CUDA(cudaMemcpyAsync(c.dDepth, ones.data(), px * sizeof(float), cudaMemcpyHostToDevice, c.stream));
        MV_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (shadow && shadowSize) {
        if (shadowSize != c.shadowSize) {
            if (c.dShadow) { MV_CUDA(cudaStreamSynchronize(c.stream)); MV_CUDA(cudaFree(c.dShadow)); c.dShadow = nullptr; }
            MV_CUDA(cudaMalloc(&c.dShadow, (size_t)shadowSize * shadowSize * sizeof(uint16_t)));
            c.shadowSize = shadowSize;
        }
        MV_CUDA(cudaMemcpyAsync(c.dShadow, shadow, (size_t)shadowSize * shadowSize * sizeof(uint16_t), kind, c.stream));
    } else c.shadowSize = 0;
    c.cb.shadowSize = c.shadowSize;
    if (color) MV_CUDA(cudaMemcpyAsync(c.dBackground, color, px * 8, kind, c.stream));
    else MV_CUDA(cudaMemsetAsync(c.dBackground, 0, px * 8, c.stream));
    MV_CUDA(cudaMemcpyAsync(c.dColor, c.dBackground, px * 8, cudaMemcpyDeviceToDevice, c.stream));
    if (velocity) MV_CUDA(cudaMemcpyAsync(c.dVelocity, velocity, px * 4, kind, c.stream));
    else MV_CUDA(cudaMemsetAsync(c.dVelocity, 0, px * 4, c.stream));
    c.velocityGiven = velocity != nullptr;
    if (kind == cudaMemcpyHostToDevice) MV_CUDA(cudaStreamSynchronize(c.stream));   // host buffers may be pageable / freed by the caller
    return MV_OK;
} catch (const std::bad_alloc&) { set_error("out of host memory"); return MV_ERR_NOMEM; }

int mv_set_targets(mv_caster* h, const float* depth, const uint16_t* shadow, uint32_t shadowSize, const uint16_t* color, const uint16_t* velocity)
{
    MV_ENTER(h);
    return set_targets_impl(c, depth, shadow, shadowSize, color, velocity, cudaMemcpyHostToDevice);
}

int mv_set_targets_device(mv_caster* h, const float* depth, const uint16_t* shadow, uint32_t shadowSize, const uint16_t* color, const uint16_t* velocity)
{
    MV_ENTER(h);
    return set_targets_impl(c, depth, shadow, shadowSize, color, velocity, cudaMemcpyDeviceToDevice);
}

int mv_reset_color(mv_caster* h)
{
    MV_ENTER(h);
    const size_t px = (size_t)c.d.width * c.d.height;
    MV_CUDA(cudaMemcpyAsync(c.dColor, c.dBackground, px * 8, cudaMemcpyDeviceToDevice, c.stream));
    return MV_OK;
}

int mv_set_sh(mv_caster* h, const float* k)
{
    MV_ENTER(h);
    c.cb.hasSH = k ? 1u : 0u;
    if (k) memcpy(c.cb.sh, k, 27 * sizeof(float));
    return MV_OK;
}

int mv_set_max_samples(mv_caster* h, uint32_t ray, uint32_t light)
{
    MV_ENTER(h);
    MV_REQUIRE(ray != 0 && light != 0 && ray < 65536 && light < 65536);
    c.d.max_ray_samples = ray; c.d.max_light_samples = light;
    c.cb.maxRaySamples = ray; c.cb.maxLightSamples = light;
    return MV_OK;
}

int mv_set_volumes_world(mv_caster* h, float size, const float center[3])
{
    MV_ENTER(h);
    MV_REQUIRE(center);
    set_volumes_world(c, size, center);
    return MV_OK;
}

int mv_set_volume_world(mv_caster* h, uint32_t i, float size, const float pos[3])
{
    MV_ENTER(h);
    MV_REQUIRE(pos && i < c.d.num_volumes);
    set_volume_world(c, i, size, pos);
    return MV_OK;
}

int mv_set_volume_world_matrix(mv_caster* h, uint32_t i, const float w[12])
{
    MV_ENTER(h);
    MV_REQUIRE(w && i < c.d.num_volumes);
    memcpy(&c.volumeWorlds[(size_t)i * 12], w, 12 * sizeof(float));
    return MV_OK;
}

int mv_set_volume_world_matrices(mv_caster* h, uint32_t first, uint32_t count, const float* w)
{
    MV_ENTER(h);
    MV_REQUIRE(w && first <= c.d.num_volumes && count <= c.d.num_volumes - first);
    memcpy(&c.volumeWorlds[(size_t)first * 12], w, (size_t)count * 12 * sizeof(float));
    return MV_OK;
}

int mv_set_light(mv_caster* h, const float p[3], const float col[3], float intensity)
{
    MV_ENTER(h);
    MV_REQUIRE(p && col);
    memcpy(c.lightPt, p, 3 * sizeof(float));
    c.lightColor[0] = col[0]; c.lightColor[1] = col[1]; c.lightColor[2] = col[2]; c.lightColor[3] = intensity;
    return MV_OK;
}

int mv_set_ambient(mv_caster* h, const float col[3], float intensity)
{
    MV_ENTER(h);
    MV_REQUIRE(col);
    c.ambient[0] = col[0]; c.ambient[1] = col[1]; c.ambient[2] = col[2]; c.ambient[3] = intensity;
    return MV_OK;
}

int mv_update_frame(mv_caster* h, const float viewProj[16], const float shadowVP[16], const float eye[3])   // MultiRayCaster.cpp:316-353
{
    MV_ENTER(h);
    MV_REQUIRE(viewProj && eye);
    FrameCB& cb = c.cb;
    memcpy(cb.eye, eye, 3 * sizeof(float));
    cb.viewport[0] = (float)c.d.width; cb.viewport[1] = (float)c.d.height;
    cb.inv2Viewport[0] = 2.0f / cb.viewport[0]; cb.inv2Viewport[1] = 2.0f / cb.viewport[1];
    inverse44(viewProj, cb.screenToWorld);
    if (shadowVP) memcpy(cb.shadowViewProj, shadowVP, 16 * sizeof(float));
    else for (int i = 0; i < 16; ++i) cb.shadowViewProj[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    cb.lightPos[0] = c.lightPt[0]; cb.lightPos[1] = c.lightPt[1]; cb.lightPos[2] = c.lightPt[2]; cb.lightPos[3] = 1.0f;
    memcpy(cb.lightColor, c.lightColor, sizeof cb.lightColor);
    memcpy(cb.ambient, c.ambient, sizeof cb.ambient);
    cb.frameIdx = c.frameIdx;
    const uint32_t N = c.d.num_volumes;
    // upload ring: a slot is rewritten only after the copy that last read it has completed
    const uint32_t slot = c.uploadSlot;
    c.uploadSlot = (slot + 1) % kUploadRing;
    if (c.uploadPending[slot]) { MV_CUDA(cudaEventSynchronize(c.uploadDone[slot])); c.uploadPending[slot] = false; }
    PerObject* staging = c.hPerObjectPinned + (size_t)slot * N;
    for (uint32_t i = 0; i < N; ++i) {
        float world[16], worldI[16], wvp[16];
        world_from43(&c.volumeWorlds[(size_t)i * 12], world);
        inverse44(world, worldI);
        mul44(world, viewProj, wvp);
        PerObject& po = staging[i];
        memcpy(po.wvp, wvp, sizeof wvp);
        inverse44(wvp, po.wvpi);
        to43(worldI, po.worldI);
        to43(world, po.world);
    }
    // the records go into the buffer the previous frame does not use; pipelined, the copy runs on the light stream (beside
    // the previous frame's passes) once the last frame that read this buffer has finished its resolve
    const uint32_t q = c.poParity ^ 1u;
    cudaStream_t up = c.stream;
    if (pipelined(c)) {
        up = c.lightStream;
        if (c.poLastUse[q] >= 0 && c.frameEndValid[c.poLastUse[q]]) MV_CUDA(cudaStreamWaitEvent(up, c.frameEnd[c.poLastUse[q]], 0));
    }
    MV_CUDA(cudaMemcpyAsync(c.dPerObject2[q], staging, N * sizeof(PerObject), cudaMemcpyHostToDevice, up));
    MV_CUDA(cudaEventRecord(c.uploadDone[slot], up));
    c.uploadPending[slot] = true;
    c.lastUpload = (int)slot;
    c.poParity = q;
    c.dPerObject = c.dPerObject2[q];
    return MV_OK;
}

int mv_cull(mv_caster* h)
{
    MV_ENTER(h);
    wait_upload(c, c.stream);
    flip_frame_lists(c);
    launch_cull(c);
    return check_launch("k_cull");
}

int mv_ray_march_light(mv_caster* h, int32_t v)
{
    MV_ENTER(h);
    MV_REQUIRE(v < (int32_t)c.d.num_volumes);
    if (c.d.flags & MV_FLAG_COUNT_SAMPLES)
        MV_CUDA(cudaMemsetAsync(&c.dStats->light_voxels, 0, 3 * sizeof(unsigned long long), c.stream));
    // stand-alone call (the frame path has the cull kernel reset these)
    MV_CUDA(cudaMemsetAsync(&reinterpret_cast<FrameLists*>(c.dLists)->lightDenseCount, 0, 7 * sizeof(uint32_t), c.stream));
    launch_ray_march_light(c, v);
    return check_launch("k_ray_march_l");
}

int mv_ray_march_view(mv_caster* h)
{
    MV_ENTER(h);
    if (c.d.flags & MV_FLAG_COUNT_SAMPLES)
        MV_CUDA(cudaMemsetAsync(&c.dStats->view_rays, 0, 3 * sizeof(unsigned long long), c.stream));
    launch_ray_march_view(c);
    return check_launch("k_ray_march_v");
}

int mv_resolve_oit(mv_caster* h)
{
    MV_ENTER(h);
    if (c.d.flags & MV_FLAG_COUNT_SAMPLES)
        MV_CUDA(cudaMemsetAsync(&c.dStats->direct_rays, 0, 4 * sizeof(unsigned long long), c.stream));
    MV_TRY_DIRECT_STATS(c);
    launch_ray_cast_direct(c);
    launch_resolve_oit(c);
    return check_launch("k_resolve_oit");
}

int mv_render(mv_caster* h, uint32_t oit)   // MultiRayCaster.cpp:355-385
{
    MV_ENTER_KEEP(h);
    if (c.envDeferred && !frame_is_pipelined_on_one_gpu(c)) flush_deferred(c);
    (void)oit;   // one OIT implementation: the K-buffer semantics of the default branch (:377-381)
    if (c.shardWorld > 1 && !c.peersMapped) {
        set_error(c.shardVolumes ? "volume-sharded caster without mapped peers (mv_ipc_import / mv_set_peer_block)"
                                 : "sharded caster without mapped peers: run the passes and the collectives one by one (mv_cull, mv_ray_march_light, ...)");
        return MV_ERR_INVALID;
    }
    if (const int rc = check_peer_timeout(c)) return rc;
    flip_frame_lists(c);
    const uint32_t slot = c.listParity;             // frameEnd slot of this frame
    c.poLastUse[c.poParity] = (int)slot;
    if (c.shardVolumes && c.shardWorld > 1) {
        // Volume-sharded storage (mv_create_sharded): light map, cube map and screen-space march of a volume are produced by
        // the rank that holds it. The light march (a no-op on every rank but the light volume's owner) goes through the staging
        // buffer, so that — uninstrumented — it can run on the light stream beside the previous frame's passes, as on one GPU;
        // it needs no exchange. View march and screen-space march store into every peer's block; one barrier, then the resolve
        // of this rank's rows. mv_postprocess ends the frame with the closing barrier.
        cudaStream_t mainStream = c.stream, B = c.lightStream;
        const bool piped = pipelined(c);
        if (piped) {
            wait_upload(c, B);
            if (c.inputsDirty) { MV_CUDA(cudaEventRecord(c.inputsReady, mainStream)); MV_CUDA(cudaStreamWaitEvent(B, c.inputsReady, 0)); c.inputsDirty = false; }
            if (c.frameEndValid[slot]) MV_CUDA(cudaStreamWaitEvent(B, c.frameEnd[slot], 0));
            if (c.commitValid) MV_CUDA(cudaStreamWaitEvent(B, c.commitDone, 0));
            c.stream = B;
        } else {
            wait_upload(c, c.stream);
            c.inputsDirty = true;
            if (c.d.flags & MV_FLAG_COUNT_SAMPLES) MV_CUDA(cudaMemsetAsync(c.dStats, 0, sizeof(StatsDev), c.stream));
            record(c, 0);
        }
        launch_cull(c);
        if (!piped) record(c, 1);
        c.lightToStaging = true;
        launch_ray_march_light(c, -1);
        c.lightToStaging = false;
        if (piped) {
            c.stream = mainStream;
            MV_CUDA(cudaEventRecord(c.lightDone, B));
            c.lightDoneValid = true;
            MV_CUDA(cudaStreamWaitEvent(mainStream, c.lightDone, 0));
        }
        launch_light_commit(c);
        if (piped) { MV_CUDA(cudaEventRecord(c.commitDone, mainStream)); c.commitValid = true; }
        else record(c, 2);
        launch_ray_march_view(c);
        if (!piped) record(c, 3);
        MV_TRY_DIRECT_STATS(c);
        launch_ray_cast_direct(c);
        wait_back_buffer_free(c, true);
        launch_peer_barrier(c);                     // every owner's cube-map texels and screen-space march results have landed
        launch_resolve_oit(c);
        if (!piped) { record(c, 4); c.evValid[5] = false; }
    } else if (pipelined_sharded(c)) {
        // Sharded frame, pipelined across frames. Light stream: cull -> this rank's z-slab of the light map, stored into every
        // rank's staging buffer (two buffers, alternating) -> signal on the light channel. Main stream: wait for every
        // rank's slab -> commit -> view march (texels into every arena) -> barrier -> screen-space march + resolve of this
        // rank's rows; mv_postprocess ends the frame with the closing barrier. What keeps the streams and ranks apart:
        //   * lists, attributes and PerObject records are double-buffered; the light stream waits for the frame before last
        //     (frameEnd), whose buffers it overwrites;
        //   * by then every rank has committed the staging buffer this frame's slabs go into: a rank's resolve of frame
        //     i - 2 follows the barrier after that frame's view march, which every rank signals after its commit;
        //   * peers store frame i's cube-map texels into this rank's arena only after the closing barrier of frame i - 1,
        //     i.e. after this rank's resolve of frame i - 1 has read it.
        cudaStream_t mainStream = c.stream, B = c.lightStream;
        wait_upload(c, B);
        if (c.inputsDirty) { MV_CUDA(cudaEventRecord(c.inputsReady, mainStream)); MV_CUDA(cudaStreamWaitEvent(B, c.inputsReady, 0)); c.inputsDirty = false; }
        if (c.frameEndValid[slot]) MV_CUDA(cudaStreamWaitEvent(B, c.frameEnd[slot], 0));
        c.stagingParity ^= 1u;
        c.dLightStaging = c.dLightStaging2[c.stagingParity];
        c.stream = B;
        launch_cull(c);
        launch_ray_march_light(c, -1);
        launch_peer_signal(c, kBarrierLight);
        c.stream = mainStream;
        MV_CUDA(cudaEventRecord(c.lightDone, B));
        c.lightDoneValid = true;
        MV_CUDA(cudaStreamWaitEvent(mainStream, c.lightDone, 0));
        launch_peer_wait(c, kBarrierLight);
        launch_light_commit(c);
        launch_view_and_direct(c);
        wait_back_buffer_free(c, true);             // once this rank signals, peers may run ahead into their post-process and store rows into rank 0's back buffer
        launch_peer_barrier(c);                     // every rank's cube-map texels have landed
        launch_resolve_oit(c);
    } else if (pipelined(c)) {
        // light stream: cull -> light march into the staging buffer. Waits: the PerObject upload; inputs changed on the main
        // stream since the last frame (volumes, depth / shadow targets); the frame before last, whose lists this frame
        // overwrites; the previous frame's commit, which reads the staging buffer.
        cudaStream_t mainStream = c.stream, B = c.lightStream;
        wait_upload(c, B);
        if (c.inputsDirty) { MV_CUDA(cudaEventRecord(c.inputsReady, mainStream)); MV_CUDA(cudaStreamWaitEvent(B, c.inputsReady, 0)); c.inputsDirty = false; }
        if (c.frameEndValid[slot]) MV_CUDA(cudaStreamWaitEvent(B, c.frameEnd[slot], 0));
        if (c.commitValid) MV_CUDA(cudaStreamWaitEvent(B, c.commitDone, 0));
        c.stream = B;
        c.lightToStaging = true;
        launch_cull(c);
        launch_ray_march_light(c, -1);
        c.lightToStaging = false;
        c.stream = mainStream;
        MV_CUDA(cudaEventRecord(c.lightDone, B));
        c.lightDoneValid = true;
        // main stream: commit the light map, then the passes that read it
        MV_CUDA(cudaStreamWaitEvent(mainStream, c.lightDone, 0));
        launch_light_commit(c);
        MV_CUDA(cudaEventRecord(c.commitDone, mainStream));
        c.commitValid = true;
        launch_view_and_direct(c);
        launch_resolve_oit(c);
    } else {
        wait_upload(c, c.stream);
        c.inputsDirty = true;                       // the next pipelined frame orders its light stream after this one
        if (c.d.flags & MV_FLAG_COUNT_SAMPLES) MV_CUDA(cudaMemsetAsync(c.dStats, 0, sizeof(StatsDev), c.stream));
        record(c, 0);
        launch_cull(c);
        record(c, 1);
        const int viewBlocks = c.shardViewBlocks >= 0 ? c.shardViewBlocks : (c.shardWorld >= 8 ? 4 : 0);
        if (c.shardWorld > 1 && c.peersMapped && c.overlapLight && viewBlocks > 0 && !(c.d.flags & MV_FLAG_TIME_PASSES)) {
            // Sharded frame, uninstrumented: the light march fills ONE volume's light map, so the view march of every other
            // volume does not depend on it. The rank's slab is marched on the light stream while the main stream marches
            // the other volumes' tiles (a few CTAs per SM fewer, so that both are resident: with 1 / world of the frame
            // per GPU both are latency-bound, not throughput-bound); the light volume's own tiles follow the commit.
            cudaStream_t mainStream = c.stream, B = c.lightStream;
            MV_CUDA(cudaEventRecord(c.cullDone, mainStream));
            MV_CUDA(cudaStreamWaitEvent(B, c.cullDone, 0));
            c.stream = B;
            launch_ray_march_light(c, -1);
            c.stream = mainStream;
            MV_CUDA(cudaEventRecord(c.lightDone, B));
            c.lightDoneValid = true;
            launch_ray_march_view(c, 1, viewBlocks);
            MV_CUDA(cudaStreamWaitEvent(mainStream, c.lightDone, 0));
            wait_back_buffer_free(c, true); launch_peer_barrier(c); launch_light_commit(c);
            launch_ray_march_view(c, 2);
            launch_peer_barrier(c);
        } else {
            launch_ray_march_light(c, -1);
            // sharded: once this rank signals, peers may run ahead into their post-process and store rows into rank 0's back buffer
            if (c.shardWorld > 1) { wait_back_buffer_free(c, true); launch_peer_barrier(c); launch_light_commit(c); }   // slabs of all ranks -> the light volume's array
            record(c, 2);
            launch_ray_march_view(c);
            if (c.shardWorld > 1) launch_peer_barrier(c);                            // every owner's cube maps have landed
        }
        record(c, 3);
        MV_TRY_DIRECT_STATS(c);
        launch_ray_cast_direct(c);
        launch_resolve_oit(c);
        record(c, 4);
        c.evValid[5] = false;
    }
    MV_CUDA(cudaEventRecord(c.frameEnd[slot], c.stream));
    c.frameEndValid[slot] = true;
    if (c.frameIdx != 0xffffffffu) ++c.frameIdx;
    return check_launch("render");
}

// Render(..., useWorkGraph = true), MultiRayCaster.cpp:358-362: the light march runs first, on the visible list of the
// previous frame, then ONE launch culls and marches the cube maps (rayMarchWG, :1370-1438), then the OIT passes.
int mv_render_work_graph(mv_caster* h, uint32_t oit)
{
    MV_ENTER(h);
    (void)oit;
    if (c.shardWorld > 1) { set_error("mv_render_work_graph: one GPU only (the sharded frame drives the passes one by one)"); return MV_ERR_INVALID; }
    flip_frame_lists(c);
    const uint32_t slot = c.listParity;
    c.poLastUse[c.poParity] = (int)slot;
    wait_upload(c, c.stream);
    // pipelined frames of mv_render may still be in flight on the light stream: this path is serial on the main stream
    if (c.lightDoneValid) MV_CUDA(cudaStreamWaitEvent(c.stream, c.lightDone, 0));
    c.inputsDirty = true;
    if (c.d.flags & MV_FLAG_COUNT_SAMPLES) MV_CUDA(cudaMemsetAsync(c.dStats, 0, sizeof(StatsDev), c.stream));
    record(c, 0);
    record(c, 1);                       // the cull has no pass of its own here: it is timed with the view march
    launch_pick_light_volume(c);
    launch_ray_march_light(c, -1);
    record(c, 2);
    launch_cull_and_ray_march_view(c);
    record(c, 3);
    MV_TRY_DIRECT_STATS(c);
    launch_ray_cast_direct(c);
    launch_resolve_oit(c);
    record(c, 4);
    c.evValid[5] = false;
    MV_CUDA(cudaEventRecord(c.frameEnd[slot], c.stream));
    c.frameEndValid[slot] = true;
    if (c.frameIdx != 0xffffffffu) ++c.frameIdx;
    return check_launch("render_work_graph");
}

int mv_postprocess(mv_caster* h, uint32_t taa)
{
    MV_ENTER(h);
    if (!c.evValid[4]) record(c, 4);
    wait_back_buffer_free(c);
    launch_postprocess(c, taa != 0);
    // Sharded, peers mapped: the frame ends with a barrier — rank 0 holds every rank's rows, every peer's TAA history holds
    // this rank's rows, and no rank starts storing the next frame's cube-map texels into a peer's arena while that peer's
    // resolve still reads it. Part of the call, so that a C-ABI user cannot leave it out.
    if (c.shardWorld > 1 && c.peersMapped) launch_peer_barrier(c);
    record(c, 5);
    return check_launch("k_postprocess");
}

int mv_sh_project(mv_caster* h, const float* cube, uint32_t size, float* out27)
{
    MV_ENTER(h);
    MV_REQUIRE(cube && out27 && size != 0 && size <= 8192);
    const size_t bytes = (size_t)6 * size * size * 3 * sizeof(float);
    float* dcube = nullptr; float* dout = nullptr;
    MV_CUDA(cudaMalloc(&dcube, bytes));
    cudaError_t e = cudaMalloc(&dout, 27 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dcube, cube, bytes, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) { launch_sh_project(c, dcube, size, dout); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out27, dout, 27 * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFree(dcube); if (dout) cudaFree(dout);
    if (e != cudaSuccess) { set_error("sh_project: %s", cudaGetErrorString(e)); return MV_ERR_CUDA; }
    return MV_OK;
}

// ---- read-backs ----
static int d2h(Caster& c, void* dst, const void* src, size_t bytes)
{
    MV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return MV_OK;
}

int mv_read_per_object(mv_caster* h, float* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    wait_upload(c, c.stream);
    return d2h(c, out, c.dPerObject, c.d.num_volumes * sizeof(PerObject));
}

static int read_list(Caster& c, uint32_t* ids, uint32_t* count, bool cube)
{
    FrameLists fl;
    int rc = d2h(c, &fl, c.dLists, sizeof fl);
    if (rc != MV_OK) return rc;
    const uint32_t n = cube ? fl.cubeCount : fl.visibleCount;
    *count = n;
    if (ids && n) {
        const DeviceScene s = c.scene();
        return d2h(c, ids, cube ? s.cubeVolumes : s.visible, n * sizeof(uint32_t));
    }
    return MV_OK;
}

int mv_read_visible(mv_caster* h, uint32_t* ids, uint32_t* count)
{
    MV_ENTER(h);
    MV_REQUIRE(count);
    return read_list(c, ids, count, false);
}

int mv_read_cube_volumes(mv_caster* h, uint32_t* ids, uint32_t* count)
{
    MV_ENTER(h);
    MV_REQUIRE(count);
    return read_list(c, ids, count, true);
}

int mv_read_attribs(mv_caster* h, uint16_t* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    return d2h(c, out, c.dAttribs, c.d.num_volumes * sizeof(ushort4));
}

int mv_read_cubemap(mv_caster* h, uint32_t v, uint32_t mip, uint16_t* rgba, float* depth)
{
    MV_ENTER(h);
    MV_REQUIRE(v < c.d.num_volumes && mip < kNumCubeMip);
    const size_t s = c.d.grid_size >> mip, texels = 6 * s * s;
    if (rgba) { const int rc = d2h(c, rgba, c.dArena + arena_color_offset(c.arena, v, mip), texels * 8); if (rc) return rc; }
    if (depth) { const int rc = d2h(c, depth, c.dArena + arena_depth_offset(c.arena, v, mip), texels * 4); if (rc) return rc; }
    return MV_OK;
}

int mv_read_lightmap(mv_caster* h, uint32_t v, uint16_t* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out && v < c.d.num_volumes);
    if (!c.lightMaps[v].array) { set_error("the light map of volume %u lives on the rank that owns its source", v); return MV_ERR_INVALID; }
    const uint32_t n = c.d.light_grid_size;
    cudaMemcpy3DParms p{};
    p.srcArray = c.lightMaps[v].array;
    p.dstPtr = make_cudaPitchedPtr(out, (size_t)n * 8, n, n);
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyDeviceToHost;
    MV_CUDA(cudaMemcpy3DAsync(&p, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return MV_OK;
}

int mv_read_frame(mv_caster* h, uint16_t* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    return d2h(c, out, c.dColor, (size_t)c.d.width * c.d.height * 8);
}

int mv_read_post(mv_caster* h, uint16_t* taa, uint8_t* rgba8)
{
    MV_ENTER(h);
    const size_t px = (size_t)c.d.width * c.d.height;
    if (taa) MV_CUDA(cudaMemcpyAsync(taa, c.dHistory[c.frameParity], px * 8, cudaMemcpyDeviceToHost, c.stream));
    if (rgba8) MV_CUDA(cudaMemcpyAsync(rgba8, c.dBackBuffer, px * 4, cudaMemcpyDeviceToHost, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return MV_OK;
}

int mv_present_async(mv_caster* h, uint8_t* host, uint32_t slot)
{
    MV_ENTER(h);
    MV_REQUIRE(slot < MV_PRESENT_SLOTS);
    if (c.presentPending[slot]) { MV_CUDA(cudaEventSynchronize(c.presentDone[slot])); c.presentPending[slot] = false; }
    MV_CUDA(cudaEventRecord(c.frameDone, c.stream));
    MV_CUDA(cudaStreamWaitEvent(c.copyStream, c.frameDone, 0));
    if (host) MV_CUDA(cudaMemcpyAsync(host, c.dBackBuffer, (size_t)c.d.width * c.d.height * 4, cudaMemcpyDeviceToHost, c.copyStream));
    MV_CUDA(cudaEventRecord(c.presentDone[slot], c.copyStream));
    c.presentPending[slot] = true;
    if (host) { c.backBufferBusy = (int)slot; c.backBufferBusyOwnRows = false; }
    return MV_OK;
}

int mv_present_wait(mv_caster* h, uint32_t slot)
{
    MV_ENTER(h);
    MV_REQUIRE(slot < MV_PRESENT_SLOTS);
    if (c.presentPending[slot]) { MV_CUDA(cudaEventSynchronize(c.presentDone[slot])); c.presentPending[slot] = false; }
    return check_peer_timeout(c);
}

int mv_get_stats(mv_caster* h, mv_stats* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    StatsDev sd; FrameLists fl;
    int rc = d2h(c, &sd, c.dStats, sizeof sd);
    if (rc) return rc;
    rc = d2h(c, &fl, c.dLists, sizeof fl);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->view_rays = sd.view_rays; out->view_samples = sd.view_samples; out->view_light_fetches = sd.view_light_fetches;
    const uint64_t L = c.d.light_grid_size;
    out->light_voxels = L * L * L; out->light_dense_voxels = sd.light_dense_voxels; out->light_samples = sd.light_samples;
    out->direct_rays = sd.direct_rays; out->direct_samples = sd.direct_samples; out->direct_light_fetches = sd.direct_light_fetches;
    out->oit_fragments = sd.oit_fragments;
    out->view_skipped = sd.view_skipped; out->direct_skipped = sd.direct_skipped;
    out->visible_count = fl.visibleCount; out->cubemap_count = fl.cubeCount; out->light_volume = fl.lightVolume;
    out->threads = (uint32_t)c.smCount;
    return MV_OK;
}

int mv_get_timings(mv_caster* h, mv_timings* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    MV_REQUIRE(c.d.flags & MV_FLAG_TIME_PASSES);
    MV_CUDA(cudaStreamSynchronize(c.stream));
    memset(out, 0, sizeof *out);
    float* slots[5] = {&out->cull, &out->ray_march_light, &out->ray_march_view, &out->resolve_oit, &out->postprocess};
    int last = 0;
    for (int i = 0; i < 5; ++i)
        if (c.evValid[i] && c.evValid[i + 1]) { MV_CUDA(cudaEventElapsedTime(slots[i], c.ev[i], c.ev[i + 1])); last = i + 1; }
    if (c.evValid[0] && last) MV_CUDA(cudaEventElapsedTime(&out->total, c.ev[0], c.ev[last]));
    return MV_OK;
}

int mv_set_frame_index(mv_caster* h, uint32_t f)
{
    MV_ENTER(h);
    c.frameIdx = f; c.cb.frameIdx = f;
    return MV_OK;
}

int mv_set_flags(mv_caster* h, uint32_t flags)
{
    MV_ENTER(h);
    MV_REQUIRE((flags & ~(MV_FLAG_COUNT_SAMPLES | MV_FLAG_TIME_PASSES)) == 0);
    c.d.flags = flags | (c.d.flags & (MV_FLAG_DENSITY_ONLY));   // the storage mode is fixed at creation
    c.inputsDirty = true;
    for (auto& v : c.evValid) v = false;
    return MV_OK;
}

int mv_sync(mv_caster* h)
{
    MV_ENTER(h);
    MV_CUDA(cudaStreamSynchronize(c.lightStream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return check_peer_timeout(c);
}

void* mv_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
void mv_host_free(void* p) { if (p) cudaFreeHost(p); }

} // extern "C"
