// Resampling family: MultiplyConstantBlock, UpsamplerBlock and the fused polyphase
// interpolating / rational-resampling FIR behind InterpolatorBlock and RationalResamplerBlock.
//
// Reference:
//   radio/blocks/signal/multiplyconstant.lua        y[i] = x[i] * c          (complex x complex, complex x real, real x real)
//   radio/blocks/signal/upsampler.lua:44-52         y[i*L] = x[i], zeros in between
//   radio/composites/interpolator.lua:31-41         MultiplyConstant(L) -> Upsampler(L) -> Lowpass(ntaps, 1/L, nyquist 1.0)
//   radio/composites/rationalresampler.lua:33-46    ... -> Downsampler(D), cutoff min(1/L, 1/D)
// The reference filters the zero-stuffed stream with the ordinary dot-product FIR (L-1 of every L products are
// 0 * tap, and D-1 of every D outputs are dropped).  The fused kernel computes only kept outputs from only the
// non-zero products: with u the zero-stuffed stream (u[iL] = c*x[i]) and J = m*D the upsampled index of output m,
//     y[m] = sum_k h[k] u[J-k] = sum_{t >= 0, k = (J mod L) + tL < M} h[k] * (c * x[floor(J/L) - t]),
// i.e. ceil(M/L) multiply-adds per output instead of M*D; the stuffed stream never exists in memory.
// Algorithmic bytes: 8 (in) + 8 L/D (out) per complex input sample.
#include "common.cuh"
#include "blocks.h"

#include <algorithm>
#include <new>
#include <vector>

namespace lrb {

namespace {

template <bool CDATA, bool CCONST>
__global__ void __launch_bounds__(256)
scale_kernel(const void* __restrict__ xv, void* __restrict__ yv, long long n, float cre, float cim) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if constexpr (CDATA) {
            const float2 v = reinterpret_cast<const float2*>(xv)[i];
            float2 o;
            if constexpr (CCONST) o = make_float2(v.x * cre - v.y * cim, v.x * cim + v.y * cre);
            else o = make_float2(v.x * cre, v.y * cre);
            reinterpret_cast<float2*>(yv)[i] = o;
        } else {
            reinterpret_cast<float*>(yv)[i] = reinterpret_cast<const float*>(xv)[i] * cre;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample_kernel(const T* __restrict__ x, T* __restrict__ y, long long n_out, int L) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += stride) {
        const long long q = j / L;
        T v{};
        if (j - q * L == 0) v = x[q];
        y[j] = v;
    }
}

__device__ __forceinline__ float2 fma_tap(float2 acc, float2 v, float h) { return __ffma2_rn(v, make_float2(h, h), acc); }
__device__ __forceinline__ float fma_tap(float acc, float v, float h) { return fmaf(v, h, acc); }
__device__ __forceinline__ float2 scaled(float2 v, float c) { return make_float2(v.x * c, v.y * c); }
__device__ __forceinline__ float scaled(float v, float c) { return v * c; }

// one thread per kept output m (global decimated index m0 + idx): J = m*D, q = floor(J/L), k0 = J mod L
template <typename T, bool SCALE>
__global__ void __launch_bounds__(256)
interp_fir_kernel(const T* __restrict__ x, const T* __restrict__ hist, T* __restrict__ y, const float* __restrict__ taps,
                  long long n_out, long long m0, long long c0, int Hn, int L, int D, int M, float c) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_out; idx += stride) {
        const long long J = (m0 + idx) * D;
        const long long q = J / L;
        int k = (int)(J - q * L);
        int i = (int)(q - c0);                     // local input index of the newest contributing sample (< n)
        T acc{};
        for (; k < M; k += L, --i) {
            T v{};
            if (i >= 0) v = x[i];
            else if (Hn + i >= 0) v = hist[Hn + i];
            else break;                            // before the start of the stream: zeros
            if constexpr (SCALE) v = scaled(v, c);
            acc = fma_tap(acc, v, __ldg(taps + k));
        }
        y[idx] = acc;
    }
}

// ---- Interpolator (D == 1) and small-D rational resampler (D <= 4), 2 <= L <= 8: register-tiled polyphase kernel
// (for D > 1 all phases are computed and D-1 of D dropped at the store: still ~3x faster than one thread per kept output).  A thread owns IT_R consecutive inputs and
// all L phases (IT_R * L accumulators); per tap row t it needs x[i - t] for its IT_R inputs, a window that slides by one
// sample per row, so one new shared-memory value per row feeds IT_R * L multiply-adds.  The window is held as two aligned
// register chunks A = x[s0 - tb*R .. +R) and B = x[s0 - (tb+1)*R .. +R): every index below is a compile-time constant and
// the chunk loads are 128-bit.  Taps sit in shared memory as [t][p] and are read as warp-wide broadcasts.
constexpr int IT_R = 4;
constexpr int IT_THREADS = 128;

// 4 consecutive samples from a 16-byte aligned shared-memory position (index a multiple of 4)
__device__ __forceinline__ void load_chunk4(const float2* p, float2 (&o)[4]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void load_chunk4(const float* p, float (&o)[4]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
constexpr int IT_TILE = IT_R * IT_THREADS;        // inputs per CTA tile

template <typename T, int L, bool SCALE>
__global__ void __launch_bounds__(IT_THREADS)
interp_tiled_kernel(const T* __restrict__ x, const T* __restrict__ hist, T* __restrict__ y, const float* __restrict__ taps_tp,
                    long long n, int Hn, int Tt, float c, int D, long long J0, long long m_lo) {
    extern __shared__ __align__(16) unsigned char it_smem[];
    float* hs = reinterpret_cast<float*>(it_smem);                       // [Tt][L]
    T* xs = reinterpret_cast<T*>(it_smem + (((size_t)Tt * L * sizeof(float) + 15) & ~(size_t)15));   // [Tt + IT_TILE], xs[Tt + j] = x[tile0 + j]
    const int tid = threadIdx.x;
    for (int i = tid; i < Tt * L; i += IT_THREADS) hs[i] = taps_tp[i];
    const long long ntiles = (n + IT_TILE - 1) / IT_TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long tile0 = tile * IT_TILE;
        __syncthreads();                                                 // previous tile's reads are done (and hs is staged)
        for (int j = tid; j < Tt + IT_TILE; j += IT_THREADS) {
            const long long i = tile0 - Tt + j;
            T v{};
            if (i >= 0) { if (i < n) v = x[i]; }
            else if (Hn + i >= 0) v = hist[Hn + i];
            if constexpr (SCALE) v = scaled(v, c);
            xs[j] = v;
        }
        __syncthreads();
        const int s0 = Tt + tid * IT_R;                                  // smem position of this thread's first input
        T acc[IT_R][L];
#pragma unroll
        for (int r = 0; r < IT_R; ++r)
#pragma unroll
            for (int p = 0; p < L; ++p) acc[r][p] = T{};
        T A[IT_R], B[IT_R];
        load_chunk4(xs + s0, A);
#pragma unroll 1
        for (int tb = 0; tb < Tt / IT_R; ++tb) {
            load_chunk4(xs + s0 - (tb + 1) * IT_R, B);
#pragma unroll
            for (int sft = 0; sft < IT_R; ++sft) {
                float h[L];
#pragma unroll
                for (int p = 0; p < L; ++p) h[p] = hs[(tb * IT_R + sft) * L + p];
#pragma unroll
                for (int r = 0; r < IT_R; ++r) {
                    const T xv = (r - sft >= 0) ? A[(r - sft >= 0) ? r - sft : 0] : B[(r - sft < 0) ? IT_R + r - sft : 0];
#pragma unroll
                    for (int p = 0; p < L; ++p) acc[r][p] = fma_tap(acc[r][p], xv, h[p]);
                }
            }
#pragma unroll
            for (int r = 0; r < IT_R; ++r) A[r] = B[r];
        }
        const long long i0 = tile0 + (long long)tid * IT_R;
        if (D == 1) {
#pragma unroll
            for (int r = 0; r < IT_R; ++r) {
                if (i0 + r < n) {
                    T* yo = y + (i0 + r) * L;
#pragma unroll
                    for (int p = 0; p < L; ++p) yo[p] = acc[r][p];
                }
            }
        } else {
            // rational resampling with a small D: every phase was computed, keep the upsampled indices J = 0 (mod D).
            // J0 = global upsampled index of this call's first input, m_lo = global index of this call's first output.
            long long J = J0 + i0 * L;
            long long m = (J + D - 1) / D;                         // first kept output at or after J
            int gap = (int)(m * D - J);                            // its distance in upsampled samples
#pragma unroll
            for (int r = 0; r < IT_R; ++r) {
#pragma unroll
                for (int p = 0; p < L; ++p) {
                    if (gap == 0) {
                        if (i0 + r < n) y[m - m_lo] = acc[r][p];
                        ++m;
                        gap = D;
                    }
                    --gap;
                }
            }
        }
    }
}

int grid_for(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)ctx().sm_count * 16;
    return (int)std::max<long long>(1, std::min(b, cap));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
ScaleBlock::ScaleBlock(float re, float im, bool cdata, bool cconst, bool dev) : cre(re), cim(im), complex_data(cdata), complex_const(cconst) {
    name = "mulconst";
    in_size = out_size = cdata ? 8 : 4;
    dev_ptrs = dev;
}
int ScaleBlock::run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) {
    *n_out = n;
    consumed += n;
    if (n == 0) return 0;
    const int g = grid_for((long long)n);
    if (complex_data && complex_const) scale_kernel<true, true><<<g, 256, 0, s>>>(dx, dy, (long long)n, cre, cim);
    else if (complex_data) scale_kernel<true, false><<<g, 256, 0, s>>>(dx, dy, (long long)n, cre, cim);
    else scale_kernel<false, false><<<g, 256, 0, s>>>(dx, dy, (long long)n, cre, cim);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

UpsampleBlock::UpsampleBlock(unsigned factor, unsigned elem, bool dev) {
    name = "upsample";
    in_size = out_size = elem;
    dev_ptrs = dev;
    L = (int)factor;
}
int UpsampleBlock::run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) {
    const long long no = (long long)n * L;
    *n_out = (size_t)no;
    consumed += n;
    if (n == 0) return 0;
    if (in_size == 8) upsample_kernel<float2><<<grid_for(no), 256, 0, s>>>((const float2*)dx, (float2*)dy, no, L);
    else upsample_kernel<float><<<grid_for(no), 256, 0, s>>>((const float*)dx, (float*)dy, no, L);
    count_launch();
    LRB_CHECK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
InterpFirBlock::InterpFirBlock(bool cdata, const float* taps_host, int ntaps, int interp, int decim, bool has_scale_, float scale_, bool dev)
    : complex_data(cdata), L(interp), D(decim), M(ntaps), has_scale(has_scale_), scale(scale_) {
    label = std::string(has_scale ? "mulconst+" : "") + "upsample+fir" + (D > 1 ? "+down" : "") + "(" + std::to_string(M) + ",x" +
            std::to_string(L) + (D > 1 ? "/" + std::to_string(D) : "") + ")";
    name = label.c_str();
    in_size = out_size = cdata ? 8 : 4;
    dev_ptrs = dev;
    h_taps.assign(taps_host, taps_host + ntaps);
    Hn = (M + L - 1) / L;
}
InterpFirBlock::~InterpFirBlock() {
    cudaFree(d_taps);
    cudaFree(d_taps_tp);
    cudaFree(d_hist[0]);
    cudaFree(d_hist[1]);
}
int InterpFirBlock::init() {
    LRB_CHECK(cudaMalloc(&d_taps, sizeof(float) * (size_t)M));
    LRB_CHECK(cudaMemcpy(d_taps, h_taps.data(), sizeof(float) * (size_t)M, cudaMemcpyHostToDevice));
    Tt = ((M + L - 1) / L + IT_R - 1) / IT_R * IT_R;
    if (D <= 4 && L >= 2 && L <= 8 && (size_t)Tt * L * sizeof(float) + (size_t)(Tt + IT_TILE) * in_size + 16 <= 48 * 1024) {
        std::vector<float> tp((size_t)Tt * L, 0.0f);
        for (int k = 0; k < M; ++k) tp[(size_t)(k / L) * L + (k % L)] = h_taps[k];
        LRB_CHECK(cudaMalloc(&d_taps_tp, sizeof(float) * tp.size()));
        LRB_CHECK(cudaMemcpy(d_taps_tp, tp.data(), sizeof(float) * tp.size(), cudaMemcpyHostToDevice));
        if (Hn < Tt) Hn = Tt;                     // the tile staging reads Tt samples of history
    }
    for (int i = 0; i < 2; ++i) {
        LRB_CHECK(cudaMalloc(&d_hist[i], in_size * (size_t)Hn));
        LRB_CHECK(cudaMemset(d_hist[i], 0, in_size * (size_t)Hn));
    }
    return 0;
}
void InterpFirBlock::state_buffers(std::vector<std::pair<void*, size_t>>& segs) {
    segs.emplace_back(d_hist[0], in_size * (size_t)Hn);
    segs.emplace_back(d_hist[1], in_size * (size_t)Hn);
}
uint64_t InterpFirBlock::outputs_before(uint64_t idx) const { return (idx * (uint64_t)L + (uint64_t)D - 1) / (uint64_t)D; }
size_t InterpFirBlock::max_output(size_t n) const { return (size_t)(((unsigned long long)n * L) / D + 2); }
int InterpFirBlock::run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) {
    const long long m_lo = (long long)outputs_before(consumed), m_hi = (long long)outputs_before(consumed + n);
    const long long no = m_hi - m_lo;
    *n_out = (size_t)no;
    if (n == 0) return 0;
    if (no > 0 && d_taps_tp) {
        const long long ntiles = ((long long)n + IT_TILE - 1) / IT_TILE;
        const int g = (int)std::min<long long>(ntiles, (long long)ctx().sm_count * 8);
        const size_t smem = (((size_t)Tt * L * sizeof(float) + 15) & ~(size_t)15) + (size_t)(Tt + IT_TILE) * in_size;
#define LRB_IT2(T, LL, S) interp_tiled_kernel<T, LL, S><<<g, IT_THREADS, smem, s>>>((const T*)dx, (const T*)d_hist[cur], (T*)dy, d_taps_tp, (long long)n, Hn, Tt, scale, D, (long long)consumed * L, m_lo)
#define LRB_IT(LL) \
        if (complex_data) { if (has_scale) LRB_IT2(float2, LL, true); else LRB_IT2(float2, LL, false); } \
        else { if (has_scale) LRB_IT2(float, LL, true); else LRB_IT2(float, LL, false); }
        switch (L) {
            case 2: LRB_IT(2); break;
            case 3: LRB_IT(3); break;
            case 4: LRB_IT(4); break;
            case 5: LRB_IT(5); break;
            case 6: LRB_IT(6); break;
            case 7: LRB_IT(7); break;
            default: LRB_IT(8); break;
        }
#undef LRB_IT
#undef LRB_IT2
        count_launch();
        LRB_CHECK(cudaGetLastError());
    } else if (no > 0) {
        const int g = grid_for(no);
#define LRB_IF(T, S) interp_fir_kernel<T, S><<<g, 256, 0, s>>>((const T*)dx, (const T*)d_hist[cur], (T*)dy, d_taps, no, m_lo, (long long)consumed, Hn, L, D, M, scale)
        if (complex_data) { if (has_scale) LRB_IF(float2, true); else LRB_IF(float2, false); }
        else { if (has_scale) LRB_IF(float, true); else LRB_IF(float, false); }
#undef LRB_IF
        count_launch();
        LRB_CHECK(cudaGetLastError());
    }
    if (launch_hist_update(dx, (long long)n, d_hist[cur], d_hist[cur ^ 1], Hn, (int)in_size, s) != 0) return -1;
    cur ^= 1;
    consumed += n;
    return 0;
}

}  // namespace lrb
