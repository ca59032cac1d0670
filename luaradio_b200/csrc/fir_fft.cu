// Fused overlap-save FIR: load -> 1024-point FFT -> x tap spectrum -> inverse FFT -> store, one kernel,
// the spectrum never leaves the SM.
//
// Reference algorithm: FIRFilterBlock:process_fft (radio/blocks/signal/firfilter.lua:320-398) with
// spectrum_utils.DFT/IDFT (radio/utilities/spectrum_utils.lua:86-113,317-349): N = 2^floor(log2(8M)),
// L = N-M+1, per block two FFTW executes + a VOLK multiply + a 1/N scaling loop + three ffi.copy, i.e. five
// passes over memory.  For M = 128 the reference picks N = 1024, L = 897; this kernel uses N = 1024 for
// every M it accepts (L = 1025 - M).  Unlike process_fft the block stays length-preserving:
// the M-1 sample overlap is re-read from the input (or the carried history), never delayed.
//
// Why FFT at all on a GPU: direct form costs M packed FFMA2 per complex sample, 128 at M = 128,
// which caps the kernel at ~35 % of the HBM roofline on the FP32 pipe; overlap-save costs ~34 per sample.
//
// Mapping.  One WARP owns one 1024-point block: N = 32 x 32, every lane holds 32 complex points in 32 packed
// f32x2 registers and runs a fully unrolled radix-2 DIT network on them (tools/gen_fft32.py); the only data
// exchange is one 32x32 transpose through a warp-private padded shared-memory tile per transform, so there is
// no CTA-wide barrier in the block loop at all (only __syncwarp).  With n = 32*n1 + n2 and k = k1 + 32*k2:
//   forward:  lane = n2: DFT32 over n1 -> * W1024^(n2*k1) -> transpose -> lane = k1: DFT32 over n2  => X[k1+32k2]
//   multiply: X[k] * H[k]/N      (H in shared memory in [k2][k1] order: conflict-free)
//   inverse:  lane = k1: IDFT32 over k2 -> * conj W1024^(k1*n2) -> transpose -> lane = n2: IDFT32 over k1 => y[n2+32n1]
// so global loads and stores are both lane-contiguous (256 B per warp instruction).
#include "common.cuh"
#include "blocks.h"
#include "fft32_gen.cuh"
#include "../../include/lrb200.h"

#include <cmath>
#include <complex>
#include <new>
#include <vector>

namespace lrb {

namespace {

constexpr int FF_N = 1024;
constexpr int FF_WARPS = 8;                       // warps (= concurrent FFT blocks) per CTA
constexpr int FF_THREADS = FF_WARPS * 32;
constexpr int FF_XSTRIDE = 33;                    // padded row stride of the transpose tile (float2 units)
constexpr int FF_XCH = 32 * FF_XSTRIDE;           // float2 per warp-private tile
constexpr size_t FF_SMEM = (size_t)(FF_N + FF_N + FF_WARPS * FF_XCH) * sizeof(float2);

__device__ __forceinline__ float2 cmul_conj_if(float2 a, float2 w, bool conj) {
    // a * w  or  a * conj(w)
    const float wi = conj ? -w.y : w.y;
    float2 t = __fmul2_rn(make_float2(-a.y, a.x), make_float2(wi, wi));
    return __ffma2_rn(a, make_float2(w.x, w.x), t);
}

// MODE 0: complex input / complex output (crcf, cccf).  MODE 1: real input packed two blocks per FFT (rrrf).
// EDGE false: interior blocks [b_lo, b_hi) whose N inputs and L outputs all lie inside x / y: unconditional,
//             fully coalesced loads and stores, one code path (the compiler otherwise clones the butterfly
//             networks behind each branch of a guarded load).
// EDGE true : the few blocks that touch the carried history (b = 0) or the end of the input (b >= b_hi):
//             edge index e = 0 -> block 0, e >= 1 -> block b_hi + e - 1; every access is bounds-checked.
template <int MODE, bool EDGE>
__global__ void __launch_bounds__(FF_THREADS, 2)
fir_fft1024_kernel(const void* __restrict__ xv, const void* __restrict__ histv, long long n, void* __restrict__ yv,
                   const float2* __restrict__ Hg, const float2* __restrict__ twg, int M, long long b_lo, long long b_hi,
                   long long nwork) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw = sm;                            // [k1][n2]  W1024^(k1*n2)
    float2* s_H = sm + FF_N;                      // [k2][k1]  H[k1 + 32 k2] / N
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* xch = sm + 2 * FF_N + warp * FF_XCH;
    for (int i = tid; i < FF_N; i += FF_THREADS) { s_tw[i] = twg[i]; s_H[i] = Hg[i]; }
    __syncthreads();

    const int L = FF_N - (M - 1);
    const int Hm1 = M - 1;
    const long long wstride = (long long)gridDim.x * FF_WARPS;
    for (long long wi = (long long)blockIdx.x * FF_WARPS + warp; wi < nwork; wi += wstride) {
        const long long b = EDGE ? (wi == 0 ? 0 : b_hi + wi - 1) : (b_lo + wi);
        float2 v[32];
        // ---- load: v[r] = X[base + 32 r + lane]
        if constexpr (MODE == 0) {
            const float2* x = reinterpret_cast<const float2*>(xv);
            const float2* hist = reinterpret_cast<const float2*>(histv);
            const long long base = b * L - Hm1;
            if constexpr (!EDGE) {
                const float2* xb = x + base + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = __ldcs(xb + 32 * r);
            } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const long long i = base + 32 * r + lane;
                    v[r] = (i >= 0) ? (i < n ? __ldg(x + i) : make_float2(0.f, 0.f)) : __ldg(hist + (Hm1 + i));
                }
            }
        } else {
            // two consecutive real blocks 2b, 2b+1 as real / imaginary part
            const float* x = reinterpret_cast<const float*>(xv);
            const float* hist = reinterpret_cast<const float*>(histv);
            const long long base0 = (2 * b) * L - Hm1, base1 = base0 + L;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const long long i0 = base0 + 32 * r + lane, i1 = base1 + 32 * r + lane;
                if constexpr (!EDGE) {
                    v[r] = make_float2(__ldcs(x + i0), __ldcs(x + i1));
                } else {
                    const float a = (i0 >= 0) ? (i0 < n ? __ldg(x + i0) : 0.f) : __ldg(hist + (Hm1 + i0));
                    const float c = (i1 >= 0) ? (i1 < n ? __ldg(x + i1) : 0.f) : __ldg(hist + (Hm1 + i1));
                    v[r] = make_float2(a, c);
                }
            }
        }

        // ---- forward pass 1: lane = n2, registers n1 -> k1 (result for k1 in v[bitrev5(k1)])
        fft32_nat2br<false>(v);
        __syncwarp();                               // previous iteration's tile reads are done
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            float2 t = v[bitrev5(k1)];
            if (k1 > 0) t = cmul_conj_if(t, s_tw[k1 * 32 + lane], false);
            xch[k1 * FF_XSTRIDE + lane] = t;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = xch[lane * FF_XSTRIDE + r];      // lane = k1, r = n2
        // ---- forward pass 2: registers n2 -> k2 (X[k1 + 32 k2] in v[bitrev5(k2)])
        fft32_nat2br<false>(v);
        // ---- multiply by the tap spectrum (1/N folded in)
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) v[bitrev5(k2)] = cmul_conj_if(v[bitrev5(k2)], s_H[k2 * 32 + lane], false);
        // ---- inverse pass 1: lane = k1, registers k2 (bit-reversed placement) -> n2 (natural)
        fft32_br2nat<true>(v);
        __syncwarp();                               // tile reads of the forward transpose are done
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) {
            float2 t = v[n2];
            if (n2 > 0) t = cmul_conj_if(t, s_tw[n2 * 32 + lane], true);
            xch[n2 * FF_XSTRIDE + lane] = t;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = xch[lane * FF_XSTRIDE + r];      // lane = n2, r = k1
        // ---- inverse pass 2: registers k1 -> n1 (y[n2 + 32 n1] in v[bitrev5(n1)])
        fft32_nat2br<true>(v);

        // ---- store the L valid outputs: block index n = 32 n1 + lane >= M-1  ->  y[b L + n - (M-1)]
        if constexpr (MODE == 0) {
            float2* y = reinterpret_cast<float2*>(yv);
            const long long obase = b * L - Hm1;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int nn = 32 * n1 + lane;
                const long long o = obase + nn;
                if (nn >= Hm1 && (!EDGE || o < n)) __stcs(y + o, v[bitrev5(n1)]);
            }
        } else {
            float* y = reinterpret_cast<float*>(yv);
            const long long obase0 = (2 * b) * L - Hm1, obase1 = obase0 + L;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int nn = 32 * n1 + lane;
                if (nn >= Hm1) {
                    const float2 t = v[bitrev5(n1)];
                    if (!EDGE || obase0 + nn < n) y[obase0 + nn] = t.x;
                    if (!EDGE || obase1 + nn < n) y[obase1 + nn] = t.y;
                }
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host-side plan: tap spectrum (float64 DFT of the zero-extended taps, scaled by 1/N, as
// firfilter.lua:337-343 does with spectrum_utils.DFT) and the 32x32 inter-pass twiddle table.
// ---------------------------------------------------------------------------------------------
struct FirFast {
    float2* d_H = nullptr;
    float2* d_tw = nullptr;
    int mode = 0;
    bool configured = false;
};

static constexpr int FFT_MAX_TAPS = 513;       // L >= 512: at most half of every block is overlap

int FirBlock::fast_init() {
    if (kind == FIR_CRCF && D > 1) poly = polyphase_prepare((const float*)h_taps.data(), M, D, 0.0);
    if (D != 1 || kind == FIR_HILBERT || M > FFT_MAX_TAPS) return 0;
    fast = new (std::nothrow) FirFast();
    if (!fast) { set_error("out of memory"); return -1; }
    fast->mode = (kind == FIR_RRRF) ? 1 : 0;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<std::complex<double>> h(M);
    for (int k = 0; k < M; ++k) {
        if (kind == FIR_CCCF) h[k] = std::complex<double>(((const float*)h_taps.data())[2 * k], ((const float*)h_taps.data())[2 * k + 1]);
        else h[k] = std::complex<double>(((const float*)h_taps.data())[k], 0.0);
    }
    std::vector<float2> H(FF_N), tw(FF_N);
    for (int k = 0; k < FF_N; ++k) {
        std::complex<double> acc(0.0, 0.0);
        for (int m = 0; m < M; ++m) {
            const int e = (int)(((long long)k * m) % FF_N);
            acc += h[m] * std::complex<double>(std::cos(two_pi * e / FF_N), -std::sin(two_pi * e / FF_N));
        }
        acc /= (double)FF_N;
        const int k1 = k % 32, k2 = k / 32;
        H[k2 * 32 + k1] = make_float2((float)acc.real(), (float)acc.imag());
    }
    for (int a = 0; a < 32; ++a)
        for (int c = 0; c < 32; ++c) {
            const int e = (a * c) % FF_N;
            tw[a * 32 + c] = make_float2((float)std::cos(two_pi * e / FF_N), (float)(-std::sin(two_pi * e / FF_N)));
        }
    LRB_CHECK(cudaMalloc(&fast->d_H, sizeof(float2) * FF_N));
    LRB_CHECK(cudaMalloc(&fast->d_tw, sizeof(float2) * FF_N));
    LRB_CHECK(cudaMemcpy(fast->d_H, H.data(), sizeof(float2) * FF_N, cudaMemcpyHostToDevice));
    LRB_CHECK(cudaMemcpy(fast->d_tw, tw.data(), sizeof(float2) * FF_N, cudaMemcpyHostToDevice));
    return 0;
}

void FirBlock::fast_free() {
    polyphase_release(poly);
    poly = nullptr;
    if (fast) {
        cudaFree(fast->d_H);
        cudaFree(fast->d_tw);
        delete fast;
        fast = nullptr;
    }
}

// The algorithm that would run for a long input (what lrb200_fir_get_algorithm reports).
int FirBlock::effective_algorithm() const {
    if (!fast || algo == LRB200_FIR_DIRECT) return LRB200_FIR_DIRECT;
    if (algo == LRB200_FIR_FFT) return LRB200_FIR_FFT;
    // automatic: overlap-save once the direct form would be FP32-bound.  Packed FFMA2 per sample:
    // direct = M (crcf), 2M (cccf), M/2 (rrrf);  overlap-save ~ 31 / (L/N) (half for packed real blocks)
    const int direct_cost = kind == FIR_CCCF ? 2 * M : kind == FIR_RRRF ? (M + 1) / 2 : M;
    const double fft_cost = 31.0 * FF_N / (double)(FF_N - M + 1) * (kind == FIR_RRRF ? 0.5 : 1.0);
    return (double)direct_cost > 1.15 * fft_cost ? LRB200_FIR_FFT : LRB200_FIR_DIRECT;
}

int FirBlock::set_algorithm(int a) {
    if (a < LRB200_FIR_AUTO || a > LRB200_FIR_FFT) { set_error("fir: unknown algorithm %d", a); return -1; }
    algo = a;
    return 0;
}

int FirBlock::fast_run(const void* dx, size_t n, void* dy, long long first, long long n_out, cudaStream_t s) {
    if (poly && D > 1)
        return launch_polyphase_crcf(poly, (const float2*)dx, (const float2*)d_hist[cur], (long long)n, (float2*)dy,
                                     first, n_out, false, 0, consumed, s);
    if (!fast || effective_algorithm() != LRB200_FIR_FFT) return 0;
    const int L = FF_N - (M - 1);
    // a forced FFT always runs; the automatic choice leaves short calls (a few blocks) to the direct kernel
    if (algo != LRB200_FIR_FFT && (long long)n < 8LL * L) return 0;
    // blocks of L outputs; in packed-real mode one FFT covers two of them
    const long long per = (fast->mode == 1) ? 2LL * L : (long long)L;
    const long long nblocks = ((long long)n + per - 1) / per;
    long long b_hi = (long long)n / per;                 // blocks [1, b_hi) are interior
    if (b_hi < 1) b_hi = 1;
    if (b_hi > nblocks) b_hi = nblocks;
    const long long n_int = b_hi - 1, n_edge = 1 + (nblocks - b_hi);
    if (!fast->configured) {
        LRB_CHECK(cudaFuncSetAttribute(fir_fft1024_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        LRB_CHECK(cudaFuncSetAttribute(fir_fft1024_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        LRB_CHECK(cudaFuncSetAttribute(fir_fft1024_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        LRB_CHECK(cudaFuncSetAttribute(fir_fft1024_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        fast->configured = true;
    }
    const long long max_ctas = (long long)ctx().sm_count * 2;
    if (n_int > 0) {
        long long ctas = (n_int + FF_WARPS - 1) / FF_WARPS;
        if (ctas > max_ctas) ctas = max_ctas;
        if (fast->mode == 0)
            fir_fft1024_kernel<0, false><<<(unsigned)ctas, FF_THREADS, FF_SMEM, s>>>(dx, d_hist[cur], (long long)n, dy, fast->d_H, fast->d_tw, M, 1, b_hi, n_int);
        else
            fir_fft1024_kernel<1, false><<<(unsigned)ctas, FF_THREADS, FF_SMEM, s>>>(dx, d_hist[cur], (long long)n, dy, fast->d_H, fast->d_tw, M, 1, b_hi, n_int);
        count_launch();
    }
    {
        long long ctas = (n_edge + FF_WARPS - 1) / FF_WARPS;
        if (fast->mode == 0)
            fir_fft1024_kernel<0, true><<<(unsigned)ctas, FF_THREADS, FF_SMEM, s>>>(dx, d_hist[cur], (long long)n, dy, fast->d_H, fast->d_tw, M, 1, b_hi, n_edge);
        else
            fir_fft1024_kernel<1, true><<<(unsigned)ctas, FF_THREADS, FF_SMEM, s>>>(dx, d_hist[cur], (long long)n, dy, fast->d_H, fast->d_tw, M, 1, b_hi, n_edge);
        count_launch();
    }
    LRB_CHECK(cudaGetLastError());
    return 1;
}

}  // namespace lrb
