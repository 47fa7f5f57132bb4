// tex_latency.cu — dependent-chain latency of trilinear RGBA16F 3-D texture fetches on B200, and of the
// ALU part of one light-march step (get_step with IEEE division, --fmad=false as in libmv_b200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -o tools/bin/tex_latency tools/tex_latency.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <vector>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// one thread walks a chain: the next coordinate depends on the fetched value (always 0 -> no change in
// the path, but the hardware cannot know). stride = distance between consecutive samples in texels.
// laneSpread = distance in texels between the rays of neighbouring lanes (8 x 4 lane grid in y, z)
__global__ void chain(cudaTextureObject_t tex, int steps, float stride, float n, long long* cycles, float* sink, int withAlu, float laneSpread)
{
    float x = 0.5f / n, y = 0.37f + (threadIdx.x & 7) * laneSpread / n, z = 0.41f + (threadIdx.x >> 3) * laneSpread / n;
    float acc = 0.0f, transm = 1.0f, prev = 0.0f, step = 0.036f;
    // warm-up of the instruction cache
    for (int i = 0; i < 4; ++i) acc += tex3D<float4>(tex, x, y, z).w;
    x += acc;
    const long long t0 = clock64();
    for (int i = 0; i < steps; ++i) {
        const float d = tex3D<float4>(tex, x, y, z).w;   // 0 everywhere
        if (withAlu) {   // the arithmetic of cast_light_ray between two fetches
            const float dD = d - prev;
            const float op = fminf(fmaxf(d * step, 0.0f), 1.0f);
            const float fEv = fminf(1.0f / 256.0f / fabsf(dD + 0.004f), 2.0f);
            const float fUi = fminf(1.0f - op, 1.0f);
            const float fTh = 1.0f - transm;
            const float ns = 0.036f * fmaxf(1.5f * fEv * fUi * fTh, 1.0f);
            prev = d;
            transm *= 1.0f - d * 0.8f;
            if (transm < 0.01f) break;
            step = ns;
            x += (step - 0.036f) + stride / n;           // depends on the fetched value
        } else x += d + stride / n;
        y += d; z += d;
        if (x > 1.0f) { x -= 1.0f; y += 7.3f / n; if (y > 1.0f) { y -= 1.0f; z += 5.1f / n; if (z > 1.0f) z -= 1.0f; } }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { *cycles = t1 - t0; }
    sink[threadIdx.x] = acc + x + y + z + transm;
}

__global__ void touch(cudaTextureObject_t tex, int n, float* sink)   // pull the whole volume through L2
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), z = blockIdx.z;
    const float v = tex3D<float4>(tex, (x + 0.5f) / n, (y + 0.5f) / n, (z + 0.5f) / n).w;
    if (v > 1.0f) sink[0] = v;
}

int main()
{
    int clockKHz = 0; cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0);
    for (int n : {128, 256}) {
        cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDescHalf4();
        CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n)));
        std::vector<unsigned short> zeros((size_t)n * n * n * 4, 0);
        cudaMemcpy3DParms p{}; p.srcPtr = make_cudaPitchedPtr(zeros.data(), (size_t)n * 8, n, n); p.dstArray = arr; p.extent = make_cudaExtent(n, n, n); p.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&p));
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear; td.normalizedCoords = 1;
        cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        long long* dc; float* sink; CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&sink, 4096));
        void* flush; const size_t fb = 512u << 20; CK(cudaMalloc(&flush, fb));
        for (int withAlu = 1; withAlu < 2; ++withAlu)
            for (float spread : {0.0f, 1.0f, 1.33f, 2.67f, 8.0f})
            for (float stride : {0.87f, 2.3f, 4.6f}) {
                for (int warm = 0; warm < 2; ++warm) {
                    const int threads = 32;
                    if (warm) { dim3 g((n + 31) / 32, (n + 7) / 8, n); touch<<<g, 256>>>(tex, n, sink); }
                    else CK(cudaMemset(flush, 1, fb));   // evict L2
                    CK(cudaDeviceSynchronize());
                    const int steps = 2000;
                    chain<<<1, threads>>>(tex, steps, stride, (float)n, dc, sink, withAlu, spread);
                    CK(cudaDeviceSynchronize());
                    long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
                    printf("{\"n\": %d, \"mb\": %.0f, \"alu\": %d, \"lane_spread\": %.2f, \"stride_texels\": %.2f, \"l2_warm\": %d, \"cycles_per_step\": %.1f, \"ns_per_step\": %.1f}\n",
                           n, (double)n * n * n * 8 / 1e6, withAlu, spread, stride, warm, (double)c / steps, (double)c / steps / (clockKHz * 1e-6));
                }
            }
        cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(dc); cudaFree(sink); cudaFree(flush);
    }
    return 0;
}
