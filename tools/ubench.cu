// Micro-benchmarks that size the FIR/FFT kernel design on B200: FFMA vs FFMA2 issue rate, shared-memory
// load bandwidth, and HBM copy bandwidth with 128-bit accesses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/ubench.bin tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) fma_kernel(float2* out, int iters, float a) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 m = make_float2(a, a * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {            // scalar FFMA x2 (same flops as one FFMA2)
                    acc[i].x = fmaf(acc[i].x, m.x, a);
                    acc[i].y = fmaf(acc[i].y, m.y, a);
                } else if (MODE == 1) {     // FFMA2, three register operands
                    acc[i] = __ffma2_rn(acc[i], m, make_float2(a, a));
                } else {                    // FFMA2 with scalar-broadcast multiplier
                    acc[i] = __ffma2_rn(acc[i], make_float2(a, a), m);
                }
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W>
__global__ void __launch_bounds__(256) lds_kernel(float* out, int iters) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (W == 16) { float4 v = sm[(idx + u * 256) & 2047]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            else if (W == 8) { float2 v = reinterpret_cast<float2*>(sm)[(idx + u * 256) & 4095]; acc.x += v.x; acc.y += v.y; }
            else { float v = reinterpret_cast<float*>(sm)[(idx + u * 256) & 8191]; acc.x += v; }
        }
        idx = (idx + 17) & 2047;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(out + i, __ldcs(in + i));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs %d clock %d kHz smem/SM %zu\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate, p.sharedMemPerMultiprocessor);
    float2* out;
    CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2) * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            int blocks = p.multiProcessorCount * 4;
            if (mode == 0) fma_kernel<0><<<blocks, 256>>>(out, iters, 1.0001f);
            if (mode == 1) fma_kernel<1><<<blocks, 256>>>(out, iters, 1.0001f);
            if (mode == 2) fma_kernel<2><<<blocks, 256>>>(out, iters, 1.0001f);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = (double)blocks * 256 * iters * 64 * 2;   // FMA lanes
            if (rep) printf("fma mode %d (%s): %.3f ms  %.2f TFMA/s = %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", mode,
                            mode == 0 ? "2x FFMA" : mode == 1 ? "FFMA2 rrr" : "FFMA2 bcast", ms, fmas / ms * 1e-9, 2 * fmas / ms * 1e-9,
                            fmas / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
        }
    }
    for (int w = 4; w <= 16; w *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            int blocks = p.multiProcessorCount * 4;
            if (w == 4) lds_kernel<4><<<blocks, 256>>>((float*)out, 5000);
            if (w == 8) lds_kernel<8><<<blocks, 256>>>((float*)out, 5000);
            if (w == 16) lds_kernel<16><<<blocks, 256>>>((float*)out, 5000);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double bytes = (double)blocks * 256 * 5000 * 16 * w;
            if (rep) printf("lds %2d B: %.3f ms  %.1f TB/s  (%.1f B/clk/SM at 1.9 GHz)\n", w, ms, bytes / ms * 1e-9, bytes / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
        }
    }
    size_t n4 = (size_t)1 << 28;   // 4 GiB
    float4 *a, *b;
    CK(cudaMalloc(&a, n4 * 16)); CK(cudaMalloc(&b, n4 * 16));
    CK(cudaMemset(a, 1, n4 * 16));
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        copy_kernel<<<p.multiProcessorCount * 16, 256>>>(a, b, n4);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("copy 4 GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n4 * 16 / ms * 1e-6);
    }
    return 0;
}
