// tex_probe.cu — pins down the sm_100a texture unit's trilinear filter arithmetic and measures the
// texture-fetch / L2 roofline denominators that MEASURED_PEAKS.json lacks.
// Not product code: a one-off measurement tool. Outputs go to gpurun_out/.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

struct Tex3D {
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    int n = 0;
};

static Tex3D make_tex_half4(const std::vector<__half>& data, int n, bool linear, bool normalized)
{
    Tex3D t; t.n = n;
    cudaChannelFormatDesc cd = cudaCreateChannelDescHalf4();
    CK(cudaMalloc3DArray(&t.arr, &cd, make_cudaExtent(n, n, n)));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)data.data(), n * 8, n, n);
    p.dstArray = t.arr;
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = t.arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = normalized ? 1 : 0;
    CK(cudaCreateTextureObject(&t.tex, &rd, &td, nullptr));
    return t;
}

static Tex3D make_tex_half1(const std::vector<__half>& data, int n)
{
    Tex3D t; t.n = n;
    cudaChannelFormatDesc cd = cudaCreateChannelDescHalf();
    CK(cudaMalloc3DArray(&t.arr, &cd, make_cudaExtent(n, n, n)));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)data.data(), n * 2, n, n);
    p.dstArray = t.arr;
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = t.arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    CK(cudaCreateTextureObject(&t.tex, &rd, &td, nullptr));
    return t;
}

__global__ void k_sample(cudaTextureObject_t tex, const float3* __restrict__ uvw, float4* __restrict__ out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float3 c = uvw[i]; out[i] = tex3D<float4>(tex, c.x, c.y, c.z); }
}

static void dump(const char* name, const void* p, size_t bytes)
{
    char path[256]; snprintf(path, sizeof path, "gpurun_out/%s", name);
    FILE* f = fopen(path, "wb"); if (!f) { perror(path); exit(1); }
    fwrite(p, 1, bytes, f); fclose(f);
}

// ---------------- throughput kernels ----------------
// Coherent march: a warp is an 8x4 tile of neighbouring rays, spaced `pitch` texels, stepping `dz`
// texels along z per fetch; fetches are independent (coordinates are arithmetic), K per thread.
template <typename T, int K>
__global__ void __launch_bounds__(256) k_tex_rate(const cudaTextureObject_t* __restrict__ texs, int ntex, float invN,
                                                   float pitch, float dz, float* __restrict__ sink)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = lane & 7, ty = lane >> 3;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    const cudaTextureObject_t tex = texs[tile % ntex];
    // tile origin scattered deterministically over the xy face
    const float ox = (float)((tile * 37) & 31) * 8.0f, oy = (float)((tile * 11) & 63) * 4.0f;
    const float u = (ox + tx * pitch + 0.37f) * invN, v = (oy + ty * pitch + 0.61f) * invN;
    float acc = 0.f;
    float w = 0.5f * invN;
    #pragma unroll 8
    for (int k = 0; k < K; ++k) {
        if constexpr (sizeof(T) == 16) { float4 c = tex3D<float4>(tex, u, v, w); acc += c.x + c.w; }
        else { acc += tex3D<float>(tex, u, v, w); }
        w += dz * invN;
    }
    if (acc == -1.f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) k_l2_read(const float4* __restrict__ buf, size_t n4, int reps, float* sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = __ldcg(buf + i); acc += v.x + v.y + v.z + v.w;
        }
    if (acc == -1.f) sink[0] = acc;
}

template <typename F> static float time_ms(F f, int warm = 2, int iters = 5)
{
    for (int i = 0; i < warm; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    FILE* js = fopen("gpurun_out/tex_probe.json", "w");
    fprintf(js, "{\"device\": \"%s\", \"sms\": %d", prop.name, prop.multiProcessorCount);

    // ---------- Test 1: weight quantisation staircase ----------
    for (int n : {8, 128, 256}) {
        std::vector<__half> data((size_t)n * n * n * 4);
        for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) {
            size_t o = (((size_t)z * n + y) * n + x) * 4;
            data[o + 0] = __float2half((float)(x & 1));      // alternates 0,1 -> lerp weight directly visible
            data[o + 1] = __float2half((float)(y & 1));
            data[o + 2] = __float2half((float)(z & 1));
            data[o + 3] = __float2half((float)x);
        }
        Tex3D t = make_tex_half4(data, n, true, true);
        const int base = (n == 8) ? 2 : (n == 128 ? 100 : 201);
        const int S = 8192;   // 32 sub-steps per 1/256
        std::vector<float3> c(3 * S);
        for (int k = 0; k < S; ++k) {
            float f = (float)k / (float)S;
            float cen = ((float)base + 0.5f);
            // axis x sweep
            c[k] = make_float3((cen + f) / n, (base + 0.5f) / n, (base + 0.5f) / n);
            c[S + k] = make_float3((base + 0.5f) / n, (cen + f) / n, (base + 0.5f) / n);
            c[2 * S + k] = make_float3((base + 0.5f) / n, (base + 0.5f) / n, (cen + f) / n);
        }
        float3* dc; float4* dout; CK(cudaMalloc(&dc, c.size() * sizeof(float3))); CK(cudaMalloc(&dout, c.size() * sizeof(float4)));
        CK(cudaMemcpy(dc, c.data(), c.size() * sizeof(float3), cudaMemcpyHostToDevice));
        k_sample<<<(int)(c.size() + 255) / 256, 256>>>(t.tex, dc, dout, (int)c.size());
        std::vector<float4> out(c.size());
        CK(cudaMemcpy(out.data(), dout, out.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        char nm[64];
        snprintf(nm, sizeof nm, "stair_%d_coords.bin", n); dump(nm, c.data(), c.size() * sizeof(float3));
        snprintf(nm, sizeof nm, "stair_%d_out.bin", n); dump(nm, out.data(), out.size() * sizeof(float4));
        cudaFree(dc); cudaFree(dout); cudaDestroyTextureObject(t.tex); cudaFreeArray(t.arr);
    }

    // ---------- Test 2: random trilinear on random fp16 texels ----------
    {
        const int n = 32; const int M = 1 << 17;
        std::vector<__half> data((size_t)n * n * n * 4);
        uint32_t s = 12345u;
        for (auto& h : data) { float v = (float)(lcg(s) >> 8) / 16777216.0f; h = __float2half(v * v * 4.0f); }
        Tex3D t = make_tex_half4(data, n, true, true);
        std::vector<float3> c(M);
        for (int i = 0; i < M; ++i) {
            float a = (float)(lcg(s) >> 8) / 16777216.0f, b = (float)(lcg(s) >> 8) / 16777216.0f, d = (float)(lcg(s) >> 8) / 16777216.0f;
            // include out-of-range coordinates to exercise clamp addressing
            c[i] = make_float3(a * 1.1f - 0.05f, b * 1.1f - 0.05f, d * 1.1f - 0.05f);
        }
        float3* dc; float4* dout; CK(cudaMalloc(&dc, M * sizeof(float3))); CK(cudaMalloc(&dout, M * sizeof(float4)));
        CK(cudaMemcpy(dc, c.data(), M * sizeof(float3), cudaMemcpyHostToDevice));
        k_sample<<<(M + 255) / 256, 256>>>(t.tex, dc, dout, M);
        std::vector<float4> out(M);
        CK(cudaMemcpy(out.data(), dout, M * sizeof(float4), cudaMemcpyDeviceToHost));
        dump("rand_tex.bin", data.data(), data.size() * sizeof(__half));
        dump("rand_coords.bin", c.data(), c.size() * sizeof(float3));
        dump("rand_out.bin", out.data(), out.size() * sizeof(float4));
        cudaFree(dc); cudaFree(dout); cudaDestroyTextureObject(t.tex); cudaFreeArray(t.arr);
    }

    // ---------- Throughput: trilinear fetch rate ----------
    float* sink; CK(cudaMalloc(&sink, 4));
    {
        struct Cfg { const char* name; int n; int ntex; bool h4; float pitch, dz; };
        const Cfg cfgs[] = {
            {"rgba16f_32_l1", 32, 1, true, 1.0f, 0.05f},
            {"rgba16f_256x8_stream_1vox", 256, 8, true, 1.0f, 1.0f},
            {"rgba16f_256x8_stream_halfvox", 256, 8, true, 1.0f, 0.5f},
            {"rgba16f_256x8_pitch2", 256, 8, true, 2.0f, 1.0f},
            {"r16f_32_l1", 32, 1, false, 1.0f, 0.05f},
            {"r16f_256x8_stream_1vox", 256, 8, false, 1.0f, 1.0f},
        };
        fprintf(js, ", \"tex_rate\": {");
        bool first = true;
        for (const Cfg& cf : cfgs) {
            std::vector<cudaTextureObject_t> texs; std::vector<cudaArray_t> arrs;
            const int n = cf.n;
            std::vector<__half> data((size_t)n * n * n * (cf.h4 ? 4 : 1));
            uint32_t s = 99u; for (auto& h : data) h = __float2half((float)(lcg(s) >> 8) / 16777216.0f);
            for (int i = 0; i < cf.ntex; ++i) {
                Tex3D t = cf.h4 ? make_tex_half4(data, n, true, true) : make_tex_half1(data, n);
                texs.push_back(t.tex); arrs.push_back(t.arr);
            }
            cudaTextureObject_t* dt; CK(cudaMalloc(&dt, texs.size() * sizeof(cudaTextureObject_t)));
            CK(cudaMemcpy(dt, texs.data(), texs.size() * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));
            constexpr int K = 240;
            const int blocks = 148 * 8 * 8;   // 8 resident CTAs/SM x 8 waves
            auto run = [&]() {
                if (cf.h4) k_tex_rate<float4, K><<<blocks, 256>>>(dt, cf.ntex, 1.0f / n, cf.pitch, cf.dz, sink);
                else k_tex_rate<float, K><<<blocks, 256>>>(dt, cf.ntex, 1.0f / n, cf.pitch, cf.dz, sink);
            };
            float ms = time_ms(run);
            double fetches = (double)blocks * 256 * K;
            double rate = fetches / (ms * 1e-3);
            printf("tex_rate %-34s %8.3f ms  %.3f Gfetch/s\n", cf.name, ms, rate * 1e-9);
            fprintf(js, "%s\"%s\": %.6e", first ? "" : ", ", cf.name, rate); first = false;
            for (auto t : texs) cudaDestroyTextureObject(t);
            for (auto a : arrs) cudaFreeArray(a);
            cudaFree(dt);
        }
        fprintf(js, "}");
    }
    // ---------- L2 read bandwidth ----------
    {
        fprintf(js, ", \"l2_read_gbs\": {");
        bool first = true;
        for (size_t mb : {16, 32, 64, 96}) {
            size_t bytes = mb << 20; float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
            const int reps = 20;
            auto run = [&]() { k_l2_read<<<148 * 8, 256>>>(buf, bytes / 16, reps, sink); };
            float ms = time_ms(run);
            double gbs = (double)bytes * reps / (ms * 1e-3) * 1e-9;
            printf("l2_read %3zu MB  %8.3f ms  %.1f GB/s\n", mb, ms, gbs);
            fprintf(js, "%s\"%zuMB\": %.1f", first ? "" : ", ", mb, gbs); first = false;
            cudaFree(buf);
        }
        fprintf(js, "}");
        // HBM streaming read for comparison (4 GB)
        size_t bytes = (size_t)4 << 30; float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        auto run = [&]() { k_l2_read<<<148 * 8, 256>>>(buf, bytes / 16, 1, sink); };
        float ms = time_ms(run);
        double gbs = (double)bytes / (ms * 1e-3) * 1e-9;
        printf("hbm_read 4096 MB %8.3f ms  %.1f GB/s\n", ms, gbs);
        fprintf(js, ", \"hbm_read_gbs\": %.1f", gbs);
        cudaFree(buf);
    }
    fprintf(js, "}\n"); fclose(js);
    return 0;
}
