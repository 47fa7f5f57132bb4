// k_env.cu — the environment pass under the volumes, and the PNG screenshot.
//
// Replaces LightProbe::RenderEnvironment (MultiVolumes/Content/LightProbe.cpp:85-97: full-screen triangle at z = 1, depth test
// LESS_EQUAL, read-only) with PSEnvironment (Content/Shaders/PSEnvironment.hlsl:46-69, infinite-size branch :66-68:
// g_txEnv.SampleLevel(g_smpLinear, rayDir, 0), alpha 0), and MultiVolumes::SaveImage (MultiVolumes.cpp:744-764, stb's PNG
// writer). The radiance cube map (LA_Radiance.dds in the reference; any 6 x S x S RGB image here) is kept as RGBA16F in linear
// memory, [face][y][x] in the D3D face order, and filtered in the kernel: bilinear over mip 0 with fp32 weights and the four
// taps resolved across cube edges (k_cube.cuh) — the sampler's own >= 8-bit fixed-point weights are a hardware detail no
// restatement can share, so both this kernel and the test oracle state the fp32 form.
#include "k_cube.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
#include <new>

using namespace mv;
struct mv_caster { Caster c; };

namespace mv {

namespace {

MV_D float lerpf(float a, float b, float t) { return fma1(b - a, t, a); }

// one thread per pixel of the rows this rank resolves (+ the halo rows its TAA reads)
__global__ void __launch_bounds__(256) k_environment(DeviceScene s, FrameCB cb, const uint2* __restrict__ cube, int S, const uint2* __restrict__ background)
{
    const int px = (int)(blockIdx.x * 32 + (threadIdx.x & 31)), py = (int)(blockIdx.y * 8 + (threadIdx.x >> 5));
    if (px >= (int)cb.width || py >= (int)cb.height) return;
    if (s.shardWorld > 1 && !row_is_resolved_here(s, cb, py)) return;
    const size_t pix = (size_t)py * cb.width + px;
    // DEPTH_READ_LESS_EQUAL against the quad's z = 1. `background` (may be null): the copy of the mesh pass's colour into the
    // colour target folded into this pass — an occluded pixel takes it, a sky pixel is overwritten anyway
    if (!(1.0f <= __ldg(s.depth + pix))) { if (background) s.color[pix] = __ldg(background + pix); return; }
    // PSEnvironment.hlsl:48-56: the pixel centre unprojected at z = 1, ray from the eye through it
    const float sx = fma1((float)px + 0.5f, cb.inv2Viewport[0], -1.0f), sy = fma1((float)py + 0.5f, -cb.inv2Viewport[1], 1.0f);
    const float* M = cb.screenToWorld;
    const float whx = fma1(sx, M[0], fma1(sy, M[4], M[8] + M[12])), why = fma1(sx, M[1], fma1(sy, M[5], M[9] + M[13]));
    const float whz = fma1(sx, M[2], fma1(sy, M[6], M[10] + M[14])), whw = fma1(sx, M[3], fma1(sy, M[7], M[11] + M[15]));
    const float iw = rcp(whw);
    const V3 viewDir = normalize(V3{cb.eye[0], cb.eye[1], cb.eye[2]} - V3{whx * iw, why * iw, whz * iw});
    const V3 d = -viewDir;
    // TextureCube lookup: face of the major axis, (u, v) on it (D3D convention), bilinear footprint on mip 0
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face; float ma;
    if (ax >= ay && ax >= az) { face = d.x > 0.0f ? 0 : 1; ma = ax; }
    else if (ay >= az) { face = d.y > 0.0f ? 2 : 3; ma = ay; }
    else { face = d.z > 0.0f ? 4 : 5; ma = az; }
    const float im = rcp(ma);
    float u, v;
    cube_face_uv(V3{d.x * im, d.y * im, d.z * im}, face, u, v);
    const float fx = fma1(u, (float)S, -0.5f), fy = fma1(v, (float)S, -0.5f);
    const float flx = floorf(fx), fly = floorf(fy);
    const float wx = fx - flx, wy = fy - fly;
    const int i0 = (int)flx, j0 = (int)fly;
    V4 t[4];                                                     // (0,0) (1,0) (0,1) (1,1)
    if (i0 >= 0 && j0 >= 0 && i0 + 1 < S && j0 + 1 < S) {        // the whole footprint on this face (nearly always)
        const uint2* p = cube + ((uint32_t)face * (uint32_t)S + (uint32_t)j0) * (uint32_t)S + (uint32_t)i0;
        t[0] = unpack_half4(__ldg(p)); t[1] = unpack_half4(__ldg(p + 1)); t[2] = unpack_half4(__ldg(p + S)); t[3] = unpack_half4(__ldg(p + S + 1));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int f, i, j;
            cube_resolve_texel(S, face, i0 + (k & 1), j0 + (k >> 1), f, i, j);
            t[k] = unpack_half4(__ldg(cube + ((size_t)f * S + j) * S + i));
        }
    }
    const V4 c = {lerpf(lerpf(t[0].x, t[1].x, wx), lerpf(t[2].x, t[3].x, wx), wy), lerpf(lerpf(t[0].y, t[1].y, wx), lerpf(t[2].y, t[3].y, wx), wy),
                  lerpf(lerpf(t[0].z, t[1].z, wx), lerpf(t[2].z, t[3].z, wx), wy), 0.0f};      // :68 alpha 0
    s.color[pix] = pack_half4(c);
}

// ---- PNG (stored deflate blocks: the image is written as it is, no compressor to get wrong) ----
uint32_t crc32_update(uint32_t crc, const unsigned char* p, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void put32(std::vector<unsigned char>& v, uint32_t x) { v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16)); v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x); }
void chunk(std::vector<unsigned char>& out, const char* type, const std::vector<unsigned char>& data)
{
    put32(out, (uint32_t)data.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, crc32_update(0xffffffffu, out.data() + at, 4 + data.size()) ^ 0xffffffffu);
}

} // namespace

void flush_deferred(Caster& c)
{
    if (!c.envDeferred) return;
    c.envDeferred = false;
    launch_environment(c, true);          // deferral implies the fold conditions (mv_render_environment)
}

void launch_environment(Caster& c, bool copyBackground)
{
    dim3 grid((c.d.width + 31) / 32, (c.d.height + 7) / 8);
    k_environment<<<grid, 256, 0, c.stream>>>(c.scene(), c.cb, c.dEnvCube, (int)c.envSize, copyBackground ? c.dBackground : nullptr);
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

extern "C" {

int mv_set_environment(mv_caster* h, const float* cubeRGB, uint32_t size)
try {
    MV_ENTER(h);
    MV_REQUIRE((cubeRGB == nullptr) == (size == 0) && size <= 8192);
    MV_CUDA(cudaStreamSynchronize(c.stream));
    if (c.dEnvCube) { MV_CUDA(cudaFree(c.dEnvCube)); c.dEnvCube = nullptr; }
    c.envSize = size;
    c.inputsDirty = true;
    if (!size) return MV_OK;
    const size_t texels = (size_t)6 * size * size;
    std::vector<uint2> half(texels);
    for (size_t i = 0; i < texels; ++i) half[i] = pack_half4(V4{cubeRGB[3 * i], cubeRGB[3 * i + 1], cubeRGB[3 * i + 2], 0.0f});
    MV_CUDA(cudaMalloc(&c.dEnvCube, texels * sizeof(uint2)));
    MV_CUDA(cudaMemcpy(c.dEnvCube, half.data(), texels * sizeof(uint2), cudaMemcpyHostToDevice));
    return MV_OK;
} catch (const std::bad_alloc&) { set_error("out of host memory"); return MV_ERR_NOMEM; }

int mv_render_environment(mv_caster* h)
{
    MV_ENTER(h);
    const size_t px = (size_t)c.d.width * c.d.height;
    // what the mesh pass left (mv_reset_color): a copy of its own, or — one GPU, whole frame — done by the environment kernel
    const bool fold = c.dEnvCube && c.shardWorld == 1 && c.row0 == 0 && c.row1 == c.d.height;
    // one GPU, frames pipelined: the pass is owed to the next mv_render, which runs it on the screen-space march's stream beside the
    // view march; any other call on the handle runs it first (flush_deferred in MV_ENTER)
    if (fold && frame_is_pipelined_on_one_gpu(c)) { c.envDeferred = true; return MV_OK; }
    if (!fold) MV_CUDA(cudaMemcpyAsync(c.dColor, c.dBackground, px * 8, cudaMemcpyDeviceToDevice, c.stream));
    if (c.dEnvCube) launch_environment(c, fold);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MV_FAIL(MV_ERR_CUDA, "k_environment launch failed: %s", cudaGetErrorString(e));
    return MV_OK;
}

int mv_write_png(const char* path, const uint8_t* rgba8, uint32_t w, uint32_t hgt)
try {
    MV_REQUIRE(path && rgba8 && w && hgt && w <= 65535u && hgt <= 65535u);
    std::vector<unsigned char> raw;                                  // filter byte 0 + the row
    raw.reserve((size_t)hgt * (1 + (size_t)w * 4));
    for (uint32_t y = 0; y < hgt; ++y) { raw.push_back(0); raw.insert(raw.end(), rgba8 + (size_t)y * w * 4, rgba8 + (size_t)(y + 1) * w * 4); }
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;                                           // Adler-32 of the raw stream
    for (size_t at = 0; at < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - at);
        z.push_back(at + n == raw.size() ? 1 : 0);
        z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
        for (size_t i = 0; i < n; ++i) { a = (a + raw[at + i]) % 65521u; b = (b + a) % 65521u; }
        z.insert(z.end(), raw.begin() + at, raw.begin() + at + n);
        at += n;
    }
    put32(z, (b << 16) | a);
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<unsigned char> ihdr;
    put32(ihdr, w); put32(ihdr, hgt);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bits, RGBA, deflate, no filter set, no interlace
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) MV_FAIL(MV_ERR_INVALID, "cannot write %s", path);
    const size_t wrote = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    if (wrote != out.size()) MV_FAIL(MV_ERR_INVALID, "short write to %s", path);
    return MV_OK;
} catch (const std::bad_alloc&) { set_error("out of host memory"); return MV_ERR_NOMEM; }

int mv_screenshot(mv_caster* h, const char* path)
try {
    MV_ENTER(h);
    MV_REQUIRE(path);
    std::vector<uint8_t> img((size_t)c.d.width * c.d.height * 4);
    MV_CUDA(cudaMemcpyAsync(img.data(), c.dBackBuffer, img.size(), cudaMemcpyDeviceToHost, c.stream));
    MV_CUDA(cudaStreamSynchronize(c.stream));
    return mv_write_png(path, img.data(), c.d.width, c.d.height);
} catch (const std::bad_alloc&) { set_error("out of host memory"); return MV_ERR_NOMEM; }

} // extern "C"
