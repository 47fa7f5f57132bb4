// tex_probe2.cu — one-hot probes that expose the texture unit's per-corner trilinear weights.
// Measurement tool only (see tools/tex_probe.cu).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

template <typename T> static cudaTextureObject_t make_tex(const std::vector<T>& data, int n, cudaChannelFormatDesc cd, cudaArray_t* out_arr)
{
    cudaArray_t arr;
    CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n)));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)data.data(), n * sizeof(T) * 4, n, n);
    p.dstArray = arr; p.extent = make_cudaExtent(n, n, n); p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    *out_arr = arr; return tex;
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
static void sample_dump(cudaTextureObject_t tex, const std::vector<float3>& c, const char* name)
{
    float3* dc; float4* dout; size_t M = c.size();
    CK(cudaMalloc(&dc, M * sizeof(float3))); CK(cudaMalloc(&dout, M * sizeof(float4)));
    CK(cudaMemcpy(dc, c.data(), M * sizeof(float3), cudaMemcpyHostToDevice));
    k_sample<<<(int)((M + 255) / 256), 256>>>(tex, dc, dout, (int)M);
    std::vector<float4> out(M);
    CK(cudaMemcpy(out.data(), dout, M * sizeof(float4), cudaMemcpyDeviceToHost));
    dump(name, out.data(), M * sizeof(float4));
    cudaFree(dc); cudaFree(dout);
}

int main()
{
    const int n = 4;
    // coordinates: (a) full 2-D sweep kx,ky in 0..256 at kz=0; (b) kz sweep at a few (kx,ky);
    // (c) random (kx,ky,kz) with sub-bin jitter 0 (bin centres)
    std::vector<float3> c;
    auto U = [&](int k) { return (1.5f + (float)k / 256.0f) / (float)n; };   // exactly representable
    for (int ky = 0; ky <= 256; ++ky) for (int kx = 0; kx <= 256; ++kx) c.push_back(make_float3(U(kx), U(ky), U(0)));
    const size_t nA = c.size();
    const int fx[6] = {0, 37, 128, 200, 255, 77}, fy[6] = {0, 91, 128, 13, 255, 77};
    for (int p = 0; p < 6; ++p) for (int kz = 0; kz <= 256; ++kz) c.push_back(make_float3(U(fx[p]), U(fy[p]), U(kz)));
    const size_t nB = c.size() - nA;
    uint32_t s = 777u;
    const int nC = 1 << 16;
    std::vector<int> kk;
    for (int i = 0; i < nC; ++i) { int a = lcg(s) >> 24, b = lcg(s) >> 24, d = lcg(s) >> 24; c.push_back(make_float3(U(a), U(b), U(d))); }
    dump("p2_coords.bin", c.data(), c.size() * sizeof(float3));
    printf("nA %zu nB %zu nC %d\n", nA, nB, nC);

    for (int layer = 0; layer < 2; ++layer) {
        // half4 one-hot
        {
            std::vector<__half> d((size_t)n * n * n * 4, __float2half(0.f));
            auto at = [&](int x, int y, int z, int ch) -> __half& { return d[(((size_t)z * n + y) * n + x) * 4 + ch]; };
            const int z = 1 + layer;
            at(1, 1, z, 0) = __float2half(1.f); at(2, 1, z, 1) = __float2half(1.f);
            at(1, 2, z, 2) = __float2half(1.f); at(2, 2, z, 3) = __float2half(1.f);
            cudaArray_t arr; cudaTextureObject_t tex = make_tex<__half>(d, n, cudaCreateChannelDescHalf4(), &arr);
            char nm[64]; snprintf(nm, sizeof nm, "p2_h4_layer%d.bin", layer); sample_dump(tex, c, nm);
            cudaDestroyTextureObject(tex); cudaFreeArray(arr);
        }
        // float4 one-hot
        {
            std::vector<float> d((size_t)n * n * n * 4, 0.f);
            auto at = [&](int x, int y, int z, int ch) -> float& { return d[(((size_t)z * n + y) * n + x) * 4 + ch]; };
            const int z = 1 + layer;
            at(1, 1, z, 0) = 1.f; at(2, 1, z, 1) = 1.f; at(1, 2, z, 2) = 1.f; at(2, 2, z, 3) = 1.f;
            cudaArray_t arr; cudaTextureObject_t tex = make_tex<float>(d, n, cudaCreateChannelDesc<float4>(), &arr);
            char nm[64]; snprintf(nm, sizeof nm, "p2_f4_layer%d.bin", layer); sample_dump(tex, c, nm);
            cudaDestroyTextureObject(tex); cudaFreeArray(arr);
        }
    }
    // value-dependence probe: same coords, texel values = assorted magnitudes in each corner (half4),
    // to see where rounding happens. channel c of corner j holds v[j][c].
    {
        std::vector<__half> d((size_t)n * n * n * 4, __float2half(0.f));
        uint32_t s2 = 4242u; std::vector<float> vals;
        for (int z = 1; z <= 2; ++z) for (int y = 1; y <= 2; ++y) for (int x = 1; x <= 2; ++x) for (int ch = 0; ch < 4; ++ch) {
            float v = (float)(lcg(s2) >> 8) / 16777216.0f; v = v * v * 4.f;
            __half h = __float2half(v); d[(((size_t)z * n + y) * n + x) * 4 + ch] = h; vals.push_back(__half2float(h));
        }
        dump("p2_vals.bin", vals.data(), vals.size() * sizeof(float));
        cudaArray_t arr; cudaTextureObject_t tex = make_tex<__half>(d, n, cudaCreateChannelDescHalf4(), &arr);
        sample_dump(tex, c, "p2_h4_vals.bin");
        cudaDestroyTextureObject(tex); cudaFreeArray(arr);
    }
    // staircases for non power-of-two sizes (light maps are 96^3)
    for (int m : {96, 48, 100}) {
        std::vector<__half> d((size_t)m * m * m * 4);
        for (int z = 0; z < m; ++z) for (int y = 0; y < m; ++y) for (int x = 0; x < m; ++x) {
            size_t o = (((size_t)z * m + y) * m + x) * 4;
            d[o] = __float2half((float)(x & 1)); d[o + 1] = __float2half((float)(y & 1)); d[o + 2] = __float2half((float)(z & 1)); d[o + 3] = __float2half((float)x);
        }
        cudaArray_t arr; cudaTextureObject_t tex = make_tex<__half>(d, m, cudaCreateChannelDescHalf4(), &arr);
        std::vector<float3> cs; uint32_t s3 = 31u + m;
        for (int i = 0; i < (1 << 16); ++i) { float a = (float)(lcg(s3) >> 8) / 16777216.0f; cs.push_back(make_float3(a, a * 0.5f + 0.25f, 0.5f)); }
        char nm[64]; snprintf(nm, sizeof nm, "p2_np2_%d_coords.bin", m); dump(nm, cs.data(), cs.size() * sizeof(float3));
        snprintf(nm, sizeof nm, "p2_np2_%d_out.bin", m); sample_dump(tex, cs, nm);
        cudaDestroyTextureObject(tex); cudaFreeArray(arr);
    }
    printf("done\n");
    return 0;
}
