// mv_dds.cu — volume ingest from DDS files.
//
// Replaces MultiRayCaster::LoadVolumeData (MultiVolumes/Content/MultiRayCaster.cpp:168-209): the DDS
// import of XUSG's DDS::Loader (XUSG/Advanced/XUSGDDSLoader.h:21-37) followed by CSR32FToRGBA16F
// (CSR32FToRGBA16F.hlsl:16-26), which resamples the scalar source texture — whatever its resolution — at
// the centres of the G^3 grid through the LINEAR sampler and stores float4(1, 1, 1, 0.25 a).
//
// The parser is host code (no device needed): "DDS " magic, DDS_HEADER, optional DDS_HEADER_DXT10; 3-D
// (volume) textures with one scalar channel — DXGI R32_FLOAT / R16_FLOAT / R16_UNORM / R8_UNORM, or the
// legacy D3DFMT codes 114 (R32F), 111 (R16F) and 8-bit luminance; only the top mip level is read.
// The resampling runs on the texture unit: the source goes into an R32F CUDA 3-D array with a
// linear / clamp texture object and k_resample_r32f fetches it at (id + 0.5) / G. When the source already
// has the grid's resolution every fetch lands on a texel centre, and the result is the texel itself.
#include "k_march.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
#include <new>

using namespace mv;

struct mv_caster { Caster c; };

namespace mv {

namespace {

__global__ void __launch_bounds__(256) k_resample_r32f(cudaSurfaceObject_t surf, cudaTextureObject_t src, uint32_t n, bool densityOnly)
{
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const uint32_t z = blockIdx.z;
    if (x >= n || y >= n) return;
    const float gridSize = (float)n;
    const float a = tex3D<float>(src, ((float)x + 0.5f) / gridSize, ((float)y + 0.5f) / gridSize, ((float)z + 0.5f) / gridSize);   // :23-24
    store_volume_texel(surf, x, y, z, V4{1.0f, 1.0f, 1.0f, a * 0.25f}, densityOnly);                                               // :26
}

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

} // namespace

} // namespace mv

#define MV_CUDA(expr)                                                                                     \
    do {                                                                                                  \
        const cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return MV_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)
#define MV_REQUIRE(cond) do { if (!(cond)) { set_error("invalid argument: %s", #cond); return MV_ERR_INVALID; } } while (0)
#define MV_ENTER(h)            \
    MV_REQUIRE(h != nullptr);  \
    Caster& c = h->c;          \
    MV_CUDA(cudaSetDevice(c.device)); \
    flush_deferred(c)

extern "C" {

int mv_dds_parse(const char* path, mv_dds_info* out)
{
    MV_REQUIRE(path && out);
    memset(out, 0, sizeof *out);
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return MV_ERR_INVALID; }
    unsigned char h[148];
    const size_t got = fread(h, 1, sizeof h, f);
    fseek(f, 0, SEEK_END);
    const long fileSize = ftell(f);
    fclose(f);
    if (got < 128 || memcmp(h, "DDS ", 4) != 0 || rd32(h + 4) != 124 || rd32(h + 76) != 32) { set_error("%s is not a DDS file", path); return MV_ERR_INVALID; }
    const uint32_t flags = rd32(h + 8), height = rd32(h + 12), width = rd32(h + 16), depth = rd32(h + 24);
    const uint32_t pfFlags = rd32(h + 80), fourCC = rd32(h + 84), rgbBits = rd32(h + 88), caps2 = rd32(h + 112);
    uint32_t offset = 128, fmt = 0;
    bool volume = (flags & 0x800000u) != 0 || (caps2 & 0x200000u) != 0;      // DDSD_DEPTH / DDSCAPS2_VOLUME
    if ((pfFlags & 0x4u) && fourCC == 0x30315844u) {                         // "DX10"
        if (got < 148) { set_error("%s: truncated DX10 header", path); return MV_ERR_INVALID; }
        const uint32_t dxgi = rd32(h + 128), dim = rd32(h + 132);
        offset = 148;
        volume = dim == 4;                                                   // D3D10_RESOURCE_DIMENSION_TEXTURE3D
        if (dxgi == 41) fmt = MV_DDS_R32_FLOAT; else if (dxgi == 54) fmt = MV_DDS_R16_FLOAT;
        else if (dxgi == 56) fmt = MV_DDS_R16_UNORM; else if (dxgi == 61) fmt = MV_DDS_R8_UNORM;
    } else if (pfFlags & 0x4u) {
        if (fourCC == 114) fmt = MV_DDS_R32_FLOAT; else if (fourCC == 111) fmt = MV_DDS_R16_FLOAT;
    } else if ((pfFlags & 0x20000u) && rgbBits == 8) fmt = MV_DDS_R8_UNORM;  // DDPF_LUMINANCE, 8 bits
    else if ((pfFlags & 0x20000u) && rgbBits == 16) fmt = MV_DDS_R16_UNORM;
    if (!fmt) { set_error("%s: unsupported DDS pixel format (one scalar channel expected)", path); return MV_ERR_INVALID; }
    if (!volume || !width || !height || !depth) { set_error("%s is not a 3-D (volume) texture", path); return MV_ERR_INVALID; }
    // header fields come from the file: bound them before multiplying (CUDA 3-D arrays end at 16384 per side anyway)
    if (width > 16384u || height > 16384u || depth > 16384u) { set_error("%s: %u x %u x %u is beyond the 16384^3 a 3-D texture can hold", path, width, height, depth); return MV_ERR_INVALID; }
    const uint32_t bpt = fmt == MV_DDS_R32_FLOAT ? 4 : (fmt == MV_DDS_R8_UNORM ? 1 : 2);
    const uint64_t need = (uint64_t)offset + (uint64_t)width * height * depth * bpt;
    if (need > (uint64_t)fileSize) { set_error("%s: file shorter than its top mip level", path); return MV_ERR_INVALID; }
    out->width = width; out->height = height; out->depth = depth; out->format = fmt; out->bytes_per_texel = bpt; out->data_offset = offset;
    return MV_OK;
}

// CSR32FToRGBA16F on a host array of any resolution (x fastest)
int mv_volume_upload_r32f_sized(mv_caster* h, uint32_t src, const float* density, uint32_t w, uint32_t hgt, uint32_t d)
{
    MV_ENTER(h);
    MV_REQUIRE(density && src < c.d.num_volume_srcs && w && hgt && d);
    c.inputsDirty = true;
    cudaArray_t arr = nullptr;
    const cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    MV_CUDA(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(w, hgt, d)));
    cudaTextureObject_t tex = 0;
    cudaError_t e;
    {
        cudaMemcpy3DParms p{};
        p.srcPtr = make_cudaPitchedPtr((void*)density, (size_t)w * sizeof(float), w, hgt);
        p.dstArray = arr;
        p.extent = make_cudaExtent(w, hgt, d);
        p.kind = cudaMemcpyHostToDevice;
        e = cudaMemcpy3DAsync(&p, c.stream);
    }
    if (e == cudaSuccess) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 1;
        e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    }
    int rc = MV_OK;
    if (e == cudaSuccess) {
        IngestTarget t;
        rc = begin_ingest(c, src, t);
        if (rc == MV_OK) {
            const uint32_t n = c.d.grid_size;
            dim3 grid((n + 31) / 32, (n + 7) / 8, n);
            k_resample_r32f<<<grid, 256, 0, c.stream>>>(t.vol->surf, tex, n, t.vol->channels == 1);
            e = cudaGetLastError();
            rc = end_ingest(c, src, t);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (tex) cudaDestroyTextureObject(tex);
    cudaFreeArray(arr);
    if (e != cudaSuccess) { set_error("volume_upload_r32f_sized: %s", cudaGetErrorString(e)); return MV_ERR_CUDA; }
    return rc;
}

// LoadVolumeData (MultiRayCaster.h:35-36)
int mv_volume_load_dds(mv_caster* h, uint32_t src, const char* path)
try {
    MV_REQUIRE(h != nullptr);
    mv_dds_info info;
    int rc = mv_dds_parse(path, &info);
    if (rc != MV_OK) return rc;
    const size_t n = (size_t)info.width * info.height * info.depth;
    std::vector<unsigned char> raw(n * info.bytes_per_texel);
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return MV_ERR_INVALID; }
    fseek(f, (long)info.data_offset, SEEK_SET);
    const size_t got = fread(raw.data(), 1, raw.size(), f);
    fclose(f);
    if (got != raw.size()) { set_error("%s: short read", path); return MV_ERR_INVALID; }
    std::vector<float> density(n);
    for (size_t i = 0; i < n; ++i) {
        switch (info.format) {
        case MV_DDS_R32_FLOAT: memcpy(&density[i], &raw[4 * i], 4); break;
        case MV_DDS_R16_FLOAT: density[i] = f16_to_f32((uint16_t)(raw[2 * i] | (raw[2 * i + 1] << 8))); break;
        case MV_DDS_R16_UNORM: density[i] = (float)(raw[2 * i] | (raw[2 * i + 1] << 8)) / 65535.0f; break;
        default: density[i] = (float)raw[i] / 255.0f; break;
        }
    }
    return mv_volume_upload_r32f_sized(h, src, density.data(), info.width, info.height, info.depth);
} catch (const std::bad_alloc&) { set_error("%s: out of host memory while reading the file", path); return MV_ERR_NOMEM; }

} // extern "C"
