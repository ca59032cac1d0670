// Direct-form FIR kernels.
//
// fir_generic_kernel: any tap count, any decimation, all four type combinations.  One output per
// thread, taps staged in shared memory in chunks, input read through L1 (every input sample is
// re-read by up to M/D neighbouring threads of the same CTA, so it is L1-resident after the first
// touch).  This is the catch-all; the hot configurations run fir_fft.cu / tuner.cu instead.
//
// Reference semantics: radio/blocks/signal/firfilter.lua:111-163 (history of M-1 samples, zero
// initial state), hilberttransform.lua:91-128, downsampler.lua:40-56 for the fused decimation.
#include "common.cuh"

namespace lrb {

namespace {

constexpr int GEN_THREADS = 256;
constexpr int GEN_TAP_CHUNK = 1024;

template <typename T>
__device__ __forceinline__ T load_logical(const T* __restrict__ x, const T* __restrict__ hist, int Hm1, long long i) {
    return (i >= 0) ? __ldg(x + i) : __ldg(hist + (Hm1 + i));
}

// acc += h * v for the four type combinations
__device__ __forceinline__ void mac(float2& acc, float2 v, float h) { acc = __ffma2_rn(v, make_float2(h, h), acc); }
__device__ __forceinline__ void mac(float2& acc, float2 v, float2 h) {
    acc = __ffma2_rn(v, make_float2(h.x, h.x), acc);
    acc = __ffma2_rn(make_float2(-v.y, v.x), make_float2(h.y, h.y), acc);
}
__device__ __forceinline__ void mac(float& acc, float v, float h) { acc = fmaf(v, h, acc); }

template <typename IN, typename TAP, bool HILBERT>
__global__ void __launch_bounds__(GEN_THREADS)
fir_generic_kernel(const IN* __restrict__ x, const IN* __restrict__ hist, const TAP* __restrict__ taps, int M, int D,
                   long long first, long long n_out, void* __restrict__ yv) {
    __shared__ TAP s_taps[GEN_TAP_CHUNK];
    const long long j = (long long)blockIdx.x * GEN_THREADS + threadIdx.x;
    const bool active = j < n_out;
    const long long c = first + j * (long long)D;   // input index this output is aligned with
    // four interleaved partial sums: shortens the float32 rounding chain (128 taps -> 32 per chain)
    IN acc0 = IN(), acc1 = IN(), acc2 = IN(), acc3 = IN();
    for (int k0 = 0; k0 < M; k0 += GEN_TAP_CHUNK) {
        const int kn = min(GEN_TAP_CHUNK, M - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < kn; k += GEN_THREADS) s_taps[k] = taps[k0 + k];
        __syncthreads();
        if (active) {
            int k = 0;
            for (; k + 4 <= kn; k += 4) {
                mac(acc0, load_logical(x, hist, M - 1, c - (k0 + k)), s_taps[k]);
                mac(acc1, load_logical(x, hist, M - 1, c - (k0 + k + 1)), s_taps[k + 1]);
                mac(acc2, load_logical(x, hist, M - 1, c - (k0 + k + 2)), s_taps[k + 2]);
                mac(acc3, load_logical(x, hist, M - 1, c - (k0 + k + 3)), s_taps[k + 3]);
            }
            for (; k < kn; ++k) mac(acc0, load_logical(x, hist, M - 1, c - (k0 + k)), s_taps[k]);
        }
    }
    if (!active) return;
    if constexpr (HILBERT) {
        // hilberttransform.lua:120-124: real = x delayed by (M-1)/2, imag = filter output
        float re = load_logical(x, hist, M - 1, c - (M - 1) / 2);
        float im = (acc0 + acc1) + (acc2 + acc3);
        reinterpret_cast<float2*>(yv)[j] = make_float2(re, im);
    } else if constexpr (sizeof(IN) == 8) {
        float2 r = __fadd2_rn(__fadd2_rn(acc0, acc1), __fadd2_rn(acc2, acc3));
        reinterpret_cast<float2*>(yv)[j] = r;
    } else {
        reinterpret_cast<float*>(yv)[j] = (acc0 + acc1) + (acc2 + acc3);
    }
}

// new_hist[j] = logical(n - H + j) over [old_hist | x]; one CTA-strided pass, distinct buffers.
template <typename T>
__global__ void hist_update_kernel(const T* __restrict__ x, long long n, const T* __restrict__ hist_old,
                                   T* __restrict__ hist_new, int H) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < H; j += gridDim.x * blockDim.x) {
        long long i = n - H + j;
        hist_new[j] = (i >= 0) ? x[i] : hist_old[H + i];
    }
}

__global__ void copy_last_kernel(const char* __restrict__ x, long long n, char* __restrict__ dst, int elem_size) {
    int t = threadIdx.x;
    if (t < elem_size) dst[t] = x[(n - 1) * elem_size + t];
}

struct ZeroSegs { void* p[32]; unsigned long long bytes[32]; int count; };
__global__ void zero_segments_kernel(ZeroSegs z) {
    for (int k = blockIdx.x; k < z.count; k += gridDim.x) {
        char* p = (char*)z.p[k];
        for (unsigned long long i = threadIdx.x; i < z.bytes[k]; i += blockDim.x) p[i] = 0;
    }
}

}  // namespace

int launch_zero_segments(void* const* ptrs, const size_t* bytes, int count, cudaStream_t s) {
    for (int done = 0; done < count; done += 32) {
        ZeroSegs z;
        z.count = count - done < 32 ? count - done : 32;
        for (int k = 0; k < z.count; ++k) { z.p[k] = ptrs[done + k]; z.bytes[k] = bytes[done + k]; }
        zero_segments_kernel<<<z.count, 256, 0, s>>>(z);
        count_launch();
    }
    LRB_CHECK(cudaGetLastError());
    return 0;
}

int launch_fir_generic(FirKind kind, const void* x, const void* hist, const void* taps, int M, int D,
                       long long first, long long n_out, void* y, cudaStream_t s) {
    if (n_out <= 0) return 0;
    const long long max_per = (long long)GEN_THREADS * 2147483647LL;
    if (n_out > max_per) { set_error("fir: n_out too large for one launch"); return -1; }
    dim3 grid((unsigned)((n_out + GEN_THREADS - 1) / GEN_THREADS)), block(GEN_THREADS);
    switch (kind) {
        case FIR_CRCF:
            fir_generic_kernel<float2, float, false><<<grid, block, 0, s>>>((const float2*)x, (const float2*)hist, (const float*)taps, M, D, first, n_out, y);
            break;
        case FIR_CCCF:
            fir_generic_kernel<float2, float2, false><<<grid, block, 0, s>>>((const float2*)x, (const float2*)hist, (const float2*)taps, M, D, first, n_out, y);
            break;
        case FIR_RRRF:
            fir_generic_kernel<float, float, false><<<grid, block, 0, s>>>((const float*)x, (const float*)hist, (const float*)taps, M, D, first, n_out, y);
            break;
        case FIR_HILBERT:
            fir_generic_kernel<float, float, true><<<grid, block, 0, s>>>((const float*)x, (const float*)hist, (const float*)taps, M, D, first, n_out, y);
            break;
    }
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

int launch_hist_update(const void* x, long long n, const void* hist_old, void* hist_new, int H, int elem_size, cudaStream_t s) {
    if (H <= 0) return 0;
    int threads = 256, blocks = (H + threads - 1) / threads;
    if (blocks > 64) blocks = 64;
    if (elem_size == 8)
        hist_update_kernel<float2><<<blocks, threads, 0, s>>>((const float2*)x, n, (const float2*)hist_old, (float2*)hist_new, H);
    else
        hist_update_kernel<float><<<blocks, threads, 0, s>>>((const float*)x, n, (const float*)hist_old, (float*)hist_new, H);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

int launch_copy_last(const void* x, long long n, void* dst, int elem_size, cudaStream_t s) {
    if (n <= 0) return 0;
    copy_last_kernel<<<1, 32, 0, s>>>((const char*)x, n, (char*)dst, elem_size);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace lrb
