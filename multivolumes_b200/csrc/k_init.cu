// k_init.cu — volume ingest kernels.
//
// k_init_grid replaces CSInitGridData (MultiVolumes/Content/Shaders/CSInitGridData.hlsl:10-27; host
// MultiRayCaster.cpp:243-264): procedural RGBA16F density, a = saturate(2 (1 - r^2)^4), colour lerped
// by height. Mode 1 multiplies the same envelope by three octaves of seeded value noise so that the
// sources of a synthetic scene differ. k_r32f_to_rgba16f replaces CSR32FToRGBA16F
// (CSR32FToRGBA16F.hlsl:16-26; host MultiRayCaster.cpp:168-209): rgb = 1, a = 0.25 * density.
// Both write the 3-D CUDA array through a surface object, 8 B per voxel, x fastest.
#include "k_march.cuh"

namespace mv {

namespace {

MV_D uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    uint32_t h = seed ^ (x * 0x8da6b343u) ^ (y * 0xd8163841u) ^ (z * 0xcb1ab31fu);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; h *= 0x27d4eb2fu; h ^= h >> 16;
    return h;
}
MV_D float lattice(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    return (float)(hash3(x, y, z, seed) >> 8) * (1.0f / 16777216.0f);
}
MV_D float value_noise(V3 p, uint32_t seed)   // p in lattice units, p >= 0; smoothstep-faded trilinear
{
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    float tx = p.x - fx, ty = p.y - fy, tz = p.z - fz;
    tx = tx * tx * (3.0f - 2.0f * tx); ty = ty * ty * (3.0f - 2.0f * ty); tz = tz * tz * (3.0f - 2.0f * tz);
    const float c000 = lattice(ix, iy, iz, seed), c001 = lattice(ix + 1, iy, iz, seed);
    const float c010 = lattice(ix, iy + 1, iz, seed), c011 = lattice(ix + 1, iy + 1, iz, seed);
    const float c100 = lattice(ix, iy, iz + 1, seed), c101 = lattice(ix + 1, iy, iz + 1, seed);
    const float c110 = lattice(ix, iy + 1, iz + 1, seed), c111 = lattice(ix + 1, iy + 1, iz + 1, seed);
    const float x00 = lerp(c000, c001, tx), x10 = lerp(c010, c011, tx);
    const float x01 = lerp(c100, c101, tx), x11 = lerp(c110, c111, tx);
    return lerp(lerp(x00, x10, ty), lerp(x01, x11, ty), tz);
}

__global__ void __launch_bounds__(256) k_init_grid(cudaSurfaceObject_t surf, uint32_t n, uint32_t mode, uint32_t seed, bool densityOnly)
{
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const uint32_t z = blockIdx.z;
    if (x >= n || y >= n) return;
    const float gridSize = (float)n;
    const V3 pos = {((float)x + 0.5f) / gridSize * 2.0f - 1.0f, ((float)y + 0.5f) / gridSize * 2.0f - 1.0f,
                    ((float)z + 0.5f) / gridSize * 2.0f - 1.0f};        // :17
    const float r_sq = dot(pos, pos);                                   // :18
    float a = 1.0f - r_sq;                                              // :19
    a *= a;                                                             // :20
    a = saturate(a * a * 2.0f);                                         // :21
    if (mode == 1) {
        const V3 q = {(pos.x + 1.0f) * 2.0f, (pos.y + 1.0f) * 2.0f, (pos.z + 1.0f) * 2.0f};
        float nz = 0.5f * value_noise(q, seed);
        nz += 0.3f * value_noise(q * 2.0f, seed ^ 0x68bc21ebu);
        nz += 0.2f * value_noise(q * 4.0f, seed ^ 0x02e5be93u);
        a = saturate(a * (0.25f + 1.5f * nz));
    }
    const V3 colorU = {1.0f, 0.6f, 0.0f}, colorD = {0.5f, 0.8f, 1.0f};  // :23-24
    const float t = saturate(pos.y * 0.5f + 0.2f);                      // :25
    const V4 out = {lerp(colorD.x, colorU.x, t), lerp(colorD.y, colorU.y, t), lerp(colorD.z, colorU.z, t), a};
    store_volume_texel(surf, x, y, z, out, densityOnly);
}

__global__ void __launch_bounds__(256) k_r32f_to_rgba16f(cudaSurfaceObject_t surf, const float* __restrict__ density, uint32_t n, bool densityOnly)
{
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const uint32_t z = blockIdx.z;
    if (x >= n || y >= n) return;
    const float d = density[((size_t)z * n + y) * n + x];
    store_volume_texel(surf, x, y, z, V4{1.0f, 1.0f, 1.0f, d * 0.25f}, densityOnly);
}

// One thread per brick: the brick is empty when no texel of the brick, or of the one-texel border a trilinear footprint
// can reach from inside it, holds a density above the empty-sample threshold of the march (NaN counts as dense).
__global__ void __launch_bounds__(128) k_build_occupancy(cudaSurfaceObject_t surf, uint32_t n, uint32_t shift, uint32_t bricks, bool densityOnly, uint32_t* bits)
{
    const uint32_t b = blockIdx.x * 128 + threadIdx.x;
    if (b >= bricks * bricks * bricks) return;
    const uint32_t bx = b % bricks, by = (b / bricks) % bricks, bz = b / (bricks * bricks);
    const int edge = 1 << shift;
    const int x0 = max((int)(bx << shift) - 1, 0), x1 = min((int)(bx << shift) + edge, (int)n - 1);
    const int y0 = max((int)(by << shift) - 1, 0), y1 = min((int)(by << shift) + edge, (int)n - 1);
    const int z0 = max((int)(bz << shift) - 1, 0), z1 = min((int)(bz << shift) + edge, (int)n - 1);
    bool empty = true;
    for (int z = z0; z <= z1 && empty; ++z)
        for (int y = y0; y <= y1 && empty; ++y)
            for (int x = x0; x <= x1; ++x) {
                uint16_t h;
                if (densityOnly) h = surf3Dread<unsigned short>(surf, x * 2, y, z);
                else h = (uint16_t)(surf3Dread<uint2>(surf, x * 8, y, z).y >> 16);
                if (!(f16_to_f32(h) <= kZeroThreshold)) { empty = false; break; }
            }
    if (empty) atomicOr(bits + (b >> 5), 1u << (b & 31));
}

// Density proxy (volume-sharded storage, include/mv.h): one thread per proxy voxel, the mean density of its f^3 block of the
// full-resolution volume — fp32, summed x fastest then y then z, times 1 / f^3 (a power of two) — rounded to binary16.
__global__ void __launch_bounds__(128) k_build_proxy(cudaSurfaceObject_t full, bool fullDensityOnly, cudaSurfaceObject_t proxy, uint32_t P, uint32_t f)
{
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= P * P * P) return;
    const uint32_t x = i % P, y = (i / P) % P, z = i / (P * P);
    float sum = 0.0f;
    for (uint32_t k = 0; k < f; ++k)
        for (uint32_t j = 0; j < f; ++j)
            for (uint32_t q = 0; q < f; ++q) {
                const int X = (int)(x * f + q), Y = (int)(y * f + j), Z = (int)(z * f + k);
                uint16_t h;
                if (fullDensityOnly) h = surf3Dread<unsigned short>(full, X * 2, Y, Z);
                else h = (uint16_t)(surf3Dread<uint2>(full, X * 8, Y, Z).y >> 16);
                sum += f16_to_f32(h);
            }
    surf3Dwrite((unsigned short)f32_to_f16(sum * (1.0f / (float)(f * f * f))), proxy, (int)(x * 2), (int)y, (int)z);
}

} // namespace

void launch_build_proxy(Caster& c, const Volume3D& full, Volume3D& proxy)
{
    const uint32_t P = proxy.edge, f = full.edge / P;
    k_build_proxy<<<(P * P * P + 127) / 128, 128, 0, c.stream>>>(full.surf, full.channels == 1, proxy.surf, P, f);
}

void launch_build_occupancy(Caster& c, uint32_t src)
{
    if (!c.dOcc) return;
    uint32_t* bits = c.dOcc + (size_t)src * c.occWords;
    cudaMemsetAsync(bits, 0, (size_t)c.occWords * sizeof(uint32_t), c.stream);
    const uint32_t total = c.occBricks * c.occBricks * c.occBricks;
    k_build_occupancy<<<(total + 127) / 128, 128, 0, c.stream>>>(c.volumes[src].surf, c.d.grid_size, c.occShift, c.occBricks, c.volumes[src].channels == 1, bits);
}

void launch_init_grid(Caster& c, Volume3D& target, uint32_t mode, uint32_t seed)
{
    const uint32_t n = c.d.grid_size;
    dim3 grid((n + 31) / 32, (n + 7) / 8, n);
    k_init_grid<<<grid, 256, 0, c.stream>>>(target.surf, n, mode, seed, target.channels == 1);
}

void launch_r32f_to_rgba16f(Caster& c, Volume3D& target, const float* devDensity)
{
    const uint32_t n = c.d.grid_size;
    dim3 grid((n + 31) / 32, (n + 7) / 8, n);
    k_r32f_to_rgba16f<<<grid, 256, 0, c.stream>>>(target.surf, devDensity, n, target.channels == 1);
}

} // namespace mv
