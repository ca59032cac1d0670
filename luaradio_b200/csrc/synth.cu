// Synthetic IQ sources generated on the device.  Counter-based: sample n depends only
// on (n, seed), so any window of the stream can be regenerated on any GPU.  The integer hash is the
// one oracle/lr_oracle.py:splitmix64 implements; white IQ is bit-identical to the oracle's.
// Reference analogues: radio/blocks/sources/uniformrandom.lua, radio/blocks/sources/signal.lua.
#include "common.cuh"

namespace lrb {

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float2 white_pair(uint64_t idx, uint32_t seed) {
    uint64_t h = splitmix64(idx * 2ull + ((uint64_t)seed << 40));
    uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    // 24-bit uniform in [-1, 1)
    return make_float2((float)(lo >> 8) * 1.1920928955078125e-07f - 1.0f,
                       (float)(hi >> 8) * 1.1920928955078125e-07f - 1.0f);
}

__global__ void synth_white_kernel(float2* __restrict__ dst, uint64_t n0, long long n, uint32_t seed) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = white_pair(n0 + (uint64_t)i, seed);
}

// x[n] = amp*exp(j(2*pi*fc/fs*n + 2*pi*dev*Int m)) + noise*w[n];  m = 0.5 sin(2 pi 1k t) + 0.5 sin(2 pi 7k t)
__global__ void synth_fm_kernel(float2* __restrict__ dst, uint64_t n0, long long n, uint32_t seed, double rate,
                                uint64_t carrier_fix, double deviation, float amp, float noise) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double f1 = 1e3, f2 = 7e3;
    const double two_pi = 6.283185307179586476925286766559;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t g = n0 + (uint64_t)i;
        double t = (double)g / rate;
        double integ = 0.5 * (1.0 - cos(two_pi * f1 * t)) / (two_pi * f1) + 0.5 * (1.0 - cos(two_pi * f2 * t)) / (two_pi * f2);
        double carrier_turns = (double)(carrier_fix * g) * 5.421010862427522e-20;   // 2^-64
        double ph = two_pi * carrier_turns + two_pi * deviation * integ;
        double s, c;
        sincos(ph, &s, &c);
        float2 w = white_pair(g, seed);
        dst[i] = make_float2((float)((double)amp * c + (double)noise * (double)w.x),
                             (float)((double)amp * s + (double)noise * (double)w.y));
    }
}

}  // namespace

int launch_synth_white(float2* dst, uint64_t n0, long long n, uint32_t seed, cudaStream_t s) {
    if (n <= 0) return 0;
    int blocks = ctx().sm_count * 8;
    synth_white_kernel<<<blocks, 256, 0, s>>>(dst, n0, n, seed);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

int launch_synth_fm(float2* dst, uint64_t n0, long long n, uint32_t seed, double rate, double carrier,
                    double deviation, float amp, float noise, cudaStream_t s) {
    if (n <= 0) return 0;
    uint64_t fix = turns_to_fix(carrier / rate);
    int blocks = ctx().sm_count * 8;
    synth_fm_kernel<<<blocks, 256, 0, s>>>(dst, n0, n, seed, rate, fix, deviation, amp, noise);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace lrb
