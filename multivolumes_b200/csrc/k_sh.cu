// k_sh.cu — spherical-harmonics projection of a radiance cube map (order 3, 9 x RGB coefficients).
//
// Replaces XUSG's SphericalHarmonics::Transform (MultiVolumes/XUSG/Advanced/XUSGSphericalHarmonics.h:25-26,
// called from LightProbe.cpp:80-83), whose three compute shaders CSSHCubeMap / CSSHSum / CSSHNormalize
// ship only as DXIL (Bin/*.cso): numthreads(32,1,1), per-wave WaveActiveSum partials of
// radiance * Y_lm * dOmega and of dOmega, hierarchical summation, final scale 4 pi / sum(dOmega)
// (SURVEY.md App. B.1). Their HLSL is not in the reference, so this follows the published DirectXSH
// SHProjectCubeMap algorithm the DXIL constants point to — PARITY UNPINNED.
//   k_sh_cubemap   : one warp = one reference thread group; per-texel basis evaluation, warp-shuffle
//                    tree reduction, one 28-float partial per warp;
//   k_sh_sum_norm  : one warp sums the partials (strided, then shuffle tree) and normalises.
// Runs once per probe (first frame in the reference, MultiVolumes.cpp:633-643): latency, not throughput.
#include "mv_internal.h"

namespace mv {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kShTerms = 28;   // 9 coefficients x RGB + weight

MV_D V3 cube_dir(float px, float py, uint32_t slice, float gridSize)   // same face convention as the cube maps
{
    const float x = (px + 0.5f) / gridSize * 2.0f - 1.0f;
    float y = (py + 0.5f) / gridSize * 2.0f - 1.0f;
    y = -y;
    switch (slice) {
    case 0: return {1.0f, y, -x};
    case 1: return {-1.0f, y, x};
    case 2: return {x, 1.0f, -y};
    case 3: return {x, -1.0f, y};
    case 4: return {x, y, 1.0f};
    default: return {-x, y, -1.0f};
    }
}

__global__ void __launch_bounds__(256) k_sh_cubemap(const float* __restrict__ cube, uint32_t size, float* __restrict__ partials)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t numWarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t texels = 6u * size * size;
    const float fS = (float)size;
    float acc[kShTerms];
#pragma unroll
    for (int i = 0; i < kShTerms; ++i) acc[i] = 0.0f;
    for (uint32_t t = warpGlobal * 32 + lane; t < texels; t += numWarps * 32) {
        const uint32_t face = t / (size * size), r = t - face * size * size;
        const uint32_t y = r / size, x = r - y * size;
        const float u = ((float)x + 0.5f) / fS * 2.0f - 1.0f;
        const float v = ((float)y + 0.5f) / fS * 2.0f - 1.0f;
        const V3 d = normalize(cube_dir((float)x, (float)y, face, fS));
        const float tt = 1.0f + u * u + v * v;
        const float w = 4.0f / (sqrtf(tt) * tt);                       // differential solid angle
        // XMSHEvalDirection basis, order 3
        const float Y[9] = {0.282094792f, -0.488602512f * d.y, 0.488602512f * d.z, -0.488602512f * d.x,
                            1.092548431f * d.x * d.y, -1.092548431f * d.y * d.z, 0.946174695f * d.z * d.z - 0.315391565f,
                            -1.092548431f * d.x * d.z, 0.546274215f * (d.x * d.x - d.y * d.y)};
        const float* px = cube + (size_t)t * 3;
        const float rgb[3] = {__ldg(px), __ldg(px + 1), __ldg(px + 2)};
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[i * 3 + k] += rgb[k] * Y[i] * w;
        acc[27] += w;
    }
#pragma unroll
    for (int i = 0; i < kShTerms; ++i) {
        float v = acc[i];
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(kFull, v, dd);   // WaveActiveSum
        if (lane == 0) partials[(size_t)warpGlobal * kShTerms + i] = v;
    }
}

__global__ void __launch_bounds__(32) k_sh_sum_norm(const float* __restrict__ partials, uint32_t numPartials, float* __restrict__ out27)
{
    const uint32_t lane = threadIdx.x;
    float sums[kShTerms];
#pragma unroll
    for (int i = 0; i < kShTerms; ++i) sums[i] = 0.0f;
    for (uint32_t p = lane; p < numPartials; p += 32)
#pragma unroll
        for (int i = 0; i < kShTerms; ++i) sums[i] += partials[(size_t)p * kShTerms + i];
#pragma unroll
    for (int i = 0; i < kShTerms; ++i)
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) sums[i] += __shfl_xor_sync(kFull, sums[i], dd);
    const float norm = 12.566371f / sums[27];                          // CSSHNormalize: coeff * 4 pi / sum(w)
    if (lane < 27) {
        float v = 0.0f;
#pragma unroll
        for (int i = 0; i < 27; ++i) if (i == (int)lane) v = sums[i];
        out27[lane] = v * norm;
    }
}

} // namespace

void launch_sh_project(Caster& c, const float* devCube, uint32_t size, float* devOut27)
{
    const uint32_t texels = 6u * size * size;
    uint32_t blocks = (texels + 255) / 256;
    const uint32_t maxBlocks = (uint32_t)(c.scratchBytes / (8 * kShTerms * sizeof(float)));
    if (blocks > maxBlocks) blocks = maxBlocks;
    if (blocks > (uint32_t)c.smCount * 4) blocks = (uint32_t)c.smCount * 4;
    if (blocks < 1) blocks = 1;
    k_sh_cubemap<<<blocks, 256, 0, c.stream>>>(devCube, size, c.dScratch);
    k_sh_sum_norm<<<1, 32, 0, c.stream>>>(c.dScratch, blocks * 8, devOut27);
}

} // namespace mv
