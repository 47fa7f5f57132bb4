// mv_multi.cu — multi-GPU side of the C-ABI (include/mv.h, "multi-GPU" section): sharding state, the
// exchange block, CUDA-IPC peer mapping, the device-side barrier and the light-map commit.
// The single-adapter reference has no counterpart; the partition follows BASELINE.json's north_star
// (by volume for cull + march, by screen band for the OIT resolve and the post-process).
#include "mv_internal.h"
#include <cstring>

using namespace mv;

struct mv_caster { Caster c; };

namespace mv {

namespace {

// Arrival counters: flags[channel][r] on rank q is written by rank r. Each barrier of a channel bumps that channel's
// sequence number; a rank signals every peer (system-scope release after a system fence, so that the peer stores of
// the kernels before it on this stream are visible first), then spins until every peer has signalled the same
// sequence. Two channels, because a pipelined sharded frame runs its light march (slab stores into the peers' staging
// buffers, channel kBarrierLight) on the light stream beside the previous frame's passes on the main stream (cube-map
// and back-buffer stores, channel kBarrierMain). A bounded spin (about 2 s) turns a dead peer into an error — a mapped
// host word that mv_sync / mv_present_wait / mv_peer_barrier / mv_render report as MV_ERR_PEER_TIMEOUT — instead of a hang.
__global__ void k_peer_signal(uint32_t* const* peerFlags, uint32_t world, uint32_t rank, uint32_t channel, uint32_t seq)
{
    const uint32_t p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    volatile uint32_t* dst = peerFlags[p] + channel * kMaxPeers + rank;
    *dst = seq;
    __threadfence_system();
}

__global__ void k_peer_wait(volatile uint32_t* flags, uint32_t world, uint32_t seq, uint32_t* timeoutFlag)
{
    const uint32_t p = threadIdx.x;
    if (p >= world) return;
    const long long t0 = clock64();
    while ((int32_t)(flags[p] - seq) < 0) {
        if (clock64() - t0 > 4000000000ll) { *timeoutFlag = 1; __threadfence_system(); break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

// light-map staging ([z][y][x] RGBA16F, linear) -> the light volume's 3-D array
__global__ void __launch_bounds__(256) k_light_commit(DeviceScene s, const uint2* __restrict__ staging, uint32_t L)
{
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const uint32_t z = blockIdx.z;
    if (x >= L || y >= L) return;
    const uint32_t volumeId = s.lists->lightVolume;
    if (s.shardVolumes && (s.volumeDescs[volumeId] & 0x3fffu) % s.shardWorld != s.shardRank) return;   // committed by the volume's owner
    surf3Dwrite(__ldg(staging + ((size_t)z * L + y) * L + x), s.lightSurf[volumeId], (int)(x * 8), (int)y, (int)z);
}

} // namespace

void launch_light_commit(Caster& c)
{
    const uint32_t L = c.d.light_grid_size;
    dim3 grid((L + 31) / 32, (L + 7) / 8, L);
    k_light_commit<<<grid, 256, 0, c.stream>>>(c.scene(), c.dLightStaging, L);
}

void launch_peer_signal(Caster& c, uint32_t channel)
{
    ++c.barrierSeqCh[channel];
    k_peer_signal<<<1, 32, 0, c.stream>>>(c.dPeerFlagPtrs, c.shardWorld, c.shardRank, channel, c.barrierSeqCh[channel]);
}

void launch_peer_wait(Caster& c, uint32_t channel)
{
    k_peer_wait<<<1, 32, 0, c.stream>>>(c.dFlags + channel * kMaxPeers, c.shardWorld, c.barrierSeqCh[channel], c.dTimeout);
}

void launch_peer_barrier(Caster& c)
{
    launch_peer_signal(c, kBarrierMain);
    launch_peer_wait(c, kBarrierMain);
}

int check_peer_timeout(Caster& c)
{
    if (c.hTimeout && *reinterpret_cast<volatile uint32_t*>(c.hTimeout)) {
        *reinterpret_cast<volatile uint32_t*>(c.hTimeout) = 0;
        set_error("multi-GPU barrier timed out on rank %u: a peer did not arrive within ~2 s; the frames since the last successful sync are not valid", c.shardRank);
        return MV_ERR_PEER_TIMEOUT;
    }
    return MV_OK;
}

} // namespace mv

#define MV_FAIL(code, ...) do { set_error(__VA_ARGS__); return code; } while (0)
#define MV_CUDA(expr)                                                                                     \
    do {                                                                                                  \
        const cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) MV_FAIL(MV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define MV_REQUIRE(cond) do { if (!(cond)) MV_FAIL(MV_ERR_INVALID, "invalid argument: %s", #cond); } while (0)
#define MV_ENTER(h)            \
    MV_REQUIRE(h != nullptr);  \
    Caster& c = h->c;          \
    MV_CUDA(cudaSetDevice(c.device)); \
    flush_deferred(c)

static int refresh_peers(Caster& c)
{
    // peer table of the kernels: arena pointers, flag pointers, rank 0's back buffer
    bool all = c.shardWorld > 1;
    for (uint32_t p = 0; p < c.shardWorld; ++p) if (p != c.shardRank && !c.peerBlock[p]) all = false;
    c.peersMapped = all;
    c.arena.numPeers = all ? c.shardWorld : 0;
    uint32_t* flagPtrs[kMaxPeers] = {};
    for (uint32_t p = 0; p < kMaxPeers; ++p) {
        unsigned char* blk = (p == c.shardRank) ? c.dBlock : c.peerBlock[p];
        c.arena.peer[p] = (all && p != c.shardRank && p < c.shardWorld) ? blk + c.layout.arena_offset : nullptr;
        flagPtrs[p] = blk ? reinterpret_cast<uint32_t*>(blk + c.layout.flags_offset) : nullptr;
        for (int k = 0; k < 2; ++k)
            c.peerHistory[p][k] = (all && p != c.shardRank && p < c.shardWorld) ? reinterpret_cast<uint2*>(blk + c.layout.history_offset[k]) : nullptr;
    }
    if (!c.dPeerFlagPtrs) MV_CUDA(cudaMalloc(&c.dPeerFlagPtrs, sizeof flagPtrs));
    MV_CUDA(cudaMemcpyAsync(c.dPeerFlagPtrs, flagPtrs, sizeof flagPtrs, cudaMemcpyHostToDevice, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    c.dPeerBackBuffer = (all && c.shardRank != 0) ? reinterpret_cast<uchar4*>(c.peerBlock[0] + c.layout.back_buffer_offset) : nullptr;
    return MV_OK;
}

extern "C" {

int mv_set_shard(mv_caster* h, uint32_t rank, uint32_t world)
{
    MV_ENTER(h);
    MV_REQUIRE(world >= 1 && world <= (uint32_t)kMaxPeers && rank < world);
    if (c.shardVolumes && (rank != c.shardRank || world != c.shardWorld)) MV_FAIL(MV_ERR_INVALID, "a volume-sharded caster keeps the rank / world it was created with (%u / %u)", c.shardRank, c.shardWorld);
    c.shardRank = rank; c.shardWorld = world;
    c.layout.light_slab_depth = (c.d.light_grid_size + world - 1) / world;
    return refresh_peers(c);
}

int mv_set_row_band(mv_caster* h, uint32_t row0, uint32_t row1)
{
    MV_ENTER(h);
    MV_REQUIRE(row0 <= row1 && row1 <= c.d.height);
    c.row0 = row0; c.row1 = row1;
    return MV_OK;
}

int mv_set_row_stripes(mv_caster* h, uint32_t stripeHeight)
{
    MV_ENTER(h);
    MV_REQUIRE(stripeHeight <= c.d.height);
    c.stripeH = stripeHeight;
    return MV_OK;
}

int mv_exchange_block(mv_caster* h, void** p, uint64_t* bytes)
{
    MV_ENTER(h);
    MV_REQUIRE(p && bytes);
    *p = c.dBlock; *bytes = c.layout.block_bytes;
    return MV_OK;
}

int mv_exchange_layout_get(mv_caster* h, mv_exchange_layout* out)
{
    MV_ENTER(h);
    MV_REQUIRE(out);
    *out = c.layout;
    return MV_OK;
}

int mv_cube_region(mv_caster* h, uint32_t v, uint32_t mip, uint64_t* co, uint64_t* cb, uint64_t* dof, uint64_t* db)
{
    MV_ENTER(h);
    MV_REQUIRE(v < c.d.num_volumes && mip < kNumCubeMip && co && cb && dof && db);
    const uint64_t s = c.d.grid_size >> mip, texels = 6 * s * s;
    *co = c.layout.arena_offset + arena_color_offset(c.arena, v, mip); *cb = texels * 8;
    *dof = c.layout.arena_offset + arena_depth_offset(c.arena, v, mip); *db = texels * 4;
    return MV_OK;
}

int mv_ipc_export(mv_caster* h, void* handle64)
{
    MV_ENTER(h);
    MV_REQUIRE(handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t hd;
    MV_CUDA(cudaIpcGetMemHandle(&hd, c.dBlock));
    memcpy(handle64, &hd, 64);
    return MV_OK;
}

int mv_ipc_import(mv_caster* h, uint32_t peer, const void* handle64)
{
    MV_ENTER(h);
    MV_REQUIRE(handle64 && peer < (uint32_t)kMaxPeers && peer != c.shardRank);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void* p = nullptr;
    MV_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c.openedIpc.push_back(p);
    c.peerBlock[peer] = static_cast<unsigned char*>(p);
    return refresh_peers(c);
}

int mv_set_peer_block(mv_caster* h, uint32_t peer, void* block)
{
    MV_ENTER(h);
    MV_REQUIRE(block && peer < (uint32_t)kMaxPeers && peer != c.shardRank);
    c.peerBlock[peer] = static_cast<unsigned char*>(block);
    return refresh_peers(c);
}

int mv_peer_barrier(mv_caster* h)
{
    MV_ENTER(h);
    MV_REQUIRE(c.peersMapped);
    launch_peer_barrier(c);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MV_FAIL(MV_ERR_CUDA, "peer barrier launch failed: %s", cudaGetErrorString(e));
    return check_peer_timeout(c);   // of an earlier barrier (this one has only been enqueued)
}

int mv_light_commit(mv_caster* h)
{
    MV_ENTER(h);
    launch_light_commit(c);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MV_FAIL(MV_ERR_CUDA, "k_light_commit launch failed: %s", cudaGetErrorString(e));
    return MV_OK;
}

int mv_set_stream(mv_caster* h, void* stream)
{
    MV_ENTER(h);
    MV_CUDA(cudaStreamSynchronize(c.lightStream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    c.stream = stream ? static_cast<cudaStream_t>(stream) : c.ownStream;
    c.inputsDirty = true;
    return MV_OK;
}

int mv_get_stream(mv_caster* h, void** stream)
{
    MV_ENTER(h);
    MV_REQUIRE(stream);
    *stream = c.stream;
    return MV_OK;
}

int mv_host_register(void* p, size_t bytes)
{
    MV_REQUIRE(p && bytes);
    MV_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return MV_OK;
}

int mv_host_unregister(void* p)
{
    MV_REQUIRE(p);
    MV_CUDA(cudaHostUnregister(p));
    return MV_OK;
}

// Rows this rank resolved -> their place in the caller's whole-frame host buffer, on the copy stream (see mv_present_async).
int mv_present_rows_async(mv_caster* h, uint8_t* host, uint32_t slot)
{
    MV_ENTER(h);
    MV_REQUIRE(host && slot < MV_PRESENT_SLOTS);
    if (c.presentPending[slot]) { MV_CUDA(cudaEventSynchronize(c.presentDone[slot])); c.presentPending[slot] = false; }
    MV_CUDA(cudaEventRecord(c.frameDone, c.stream));
    MV_CUDA(cudaStreamWaitEvent(c.copyStream, c.frameDone, 0));
    const size_t rowBytes = (size_t)c.d.width * 4;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(c.dBackBuffer);
    if (c.shardWorld > 1 && c.stripeH) {
        // the rank's k-th stripe is rows [(k world + rank) stripeH, ... + stripeH), clipped to the image
        const uint32_t H = c.d.height, sh = c.stripeH, world = c.shardWorld, rank = c.shardRank;
        const uint32_t own = num_own_stripes(H, sh, rank, world);
        // one plain asynchronous copy per stripe (its rows are contiguous). Measured on 8 x B200 (profiles/r02_scaling.md): as ONE
        // pitched cudaMemcpy2DAsync into the registered shared-memory frame the call held the host long enough to cap the
        // frame loop at 1189 frames/s end to end; as nine 1-D copies it runs at 1585 (the loop without any read-back: 1675)
        for (uint32_t k = 0; k < own; ++k) {
            const uint32_t begin = (k * world + rank) * sh, end = begin + sh < H ? begin + sh : H;
            MV_CUDA(cudaMemcpyAsync(host + (size_t)begin * rowBytes, src + (size_t)begin * rowBytes, (size_t)(end - begin) * rowBytes, cudaMemcpyDeviceToHost, c.copyStream));
        }
    } else if (c.row1 > c.row0)
        MV_CUDA(cudaMemcpyAsync(host + (size_t)c.row0 * rowBytes, src + (size_t)c.row0 * rowBytes, (size_t)(c.row1 - c.row0) * rowBytes, cudaMemcpyDeviceToHost, c.copyStream));
    MV_CUDA(cudaEventRecord(c.presentDone[slot], c.copyStream));
    c.presentPending[slot] = true;
    c.backBufferBusy = (int)slot;
    c.backBufferBusyOwnRows = true;
    return MV_OK;
}

int mv_frame_buffers(mv_caster* h, void** color, void** post, void** back)
{
    MV_ENTER(h);
    if (color) *color = c.dColor;
    if (post) *post = c.dHistory[c.frameParity];
    if (back) *back = c.dBackBuffer;
    return MV_OK;
}

} // extern "C"
