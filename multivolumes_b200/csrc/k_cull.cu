// k_cull.cu — per-volume viewport cull, face-visibility mask, cube-map LOD / sample-count estimate,
// cube-map-vs-direct decision and visible-list compaction.
//
// Replaces CSVolumeCull (MultiVolumes/Content/Shaders/CSVolumeCull.hlsl:13-78 with
// VolumeCull.hlsli:27-334) and CSCopyVolumeDrawArg. Same lane mapping as the reference — 8 lanes per
// volume (one cube corner each), 4 volumes per warp — on __ballot_sync / __shfl_sync, but the two
// atomic AppendStructuredBuffer appends become a ballot + block prefix sum, so both lists come out in
// ascending volume order on every run. The kernel also produces what the reference needs extra
// dispatches or ExecuteIndirect arguments for: the tile prefix of the view march (exact
// (G >> mip)^2 texels per visible face, as the work-graph variant LibRayMarch.hlsl:120-121 launches),
// the work-stealing cursors, and the light-march volume of this frame (CSRayMarchL.hlsl:29-33).
// One CTA of 32 warps (128 volumes per sweep); N <= a few thousand, so this is latency-bound.
#include "k_cull.cuh"

namespace mv {

namespace {

constexpr int kCullThreads = 1024;

__global__ void __launch_bounds__(kCullThreads, 1) k_cull(DeviceScene s, FrameCB cb)
{
    cull_body<kCullThreads>(s, cb, true);
}

} // namespace

void launch_cull(Caster& c)
{
    k_cull<<<1, kCullThreads, 0, c.stream>>>(c.scene(), c.cb);
}

} // namespace mv
