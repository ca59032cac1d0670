// Register-tiled polyphase decimating FIR (complex in, real taps) with an optional fused frequency
// translator: the TunerBlock / DecimatorBlock kernel.
//
// Reference dataflow being fused (composites/tuner.lua:40-47):
//     FrequencyTranslatorBlock(offset) -> LowpassFilterBlock(128, bw/2) -> DownsamplerBlock(D)
// i.e. three processes and two socket hops, with the FIR computing D times more outputs than are
// kept (firfilter.lua:121-124 runs one dot product per INPUT sample).  Here only the kept outputs are
// computed (M/D complex-by-real MACs per input sample), the rotation is applied once per input sample
// while the tile is staged into shared memory, and the input is read from HBM exactly once:
// algorithmic traffic 8 + 8/D bytes per input sample.
//
// Tiling.  A CTA produces TO = 128 threads x R = 8 consecutive decimated outputs.  With reversed,
// zero-padded taps hr[0 .. Q*D), Q = ceil(M/D), and B the first input index the tile needs,
//     y[m0 + r] = sum_p sum_q hr[q*D + p] * S_p[r + q],      S_p[j] = X[B + j*D + p]
// so the tile is staged de-interleaved by polyphase branch p.  Per branch a thread slides an R-wide
// register window over S_p: one new 8-byte LDS per R FFMA2 (the complex sample is one packed f32x2
// register, the real tap is the scalar-broadcast operand of FFMA2).  Branch arrays are padded by
// 2 samples every 8 so that the per-thread stride is 10 samples = 20 banks and 128-bit shared loads of
// a quarter warp fall in distinct banks.  Taps live in the kernel-parameter constant bank.
#include "common.cuh"
#include "blocks.h"

#include <cmath>
#include <cstring>
#include <new>

namespace lrb {

namespace {

constexpr int PT_THREADS = 128;
constexpr int PT_R = 8;
constexpr int PT_TO = PT_THREADS * PT_R;     // outputs per tile
constexpr int PT_MAXTAPS = 144;              // Q*D upper bound for the instantiated shapes
constexpr int PT_MAXIT = 48;                 // staging iterations (pairs per thread) upper bound

struct PolyParams {
    float hr[PT_MAXTAPS];        // reversed taps, zero padded to Q*D
    float2 step[PT_MAXIT];       // exp(j*2*pi*turns * 2*PT_THREADS*it): staging-iteration phasor
    float2 rot1;                 // exp(j*2*pi*turns): one-sample rotation
    uint64_t turns_fix;          // turns per sample, 2^-64 units
    uint64_t g0;                 // global index of x[0]
    int M;
};

__host__ __device__ constexpr int pad_idx(int j) { return j + 2 * (j / 8); }

template <int D, int Q>
struct PolyShape {
    static constexpr int J = PT_TO + Q;                    // samples per branch (one spare)
    static constexpr int JP = (pad_idx(J) + 3) & ~1;       // padded branch length (even: keeps 16 B alignment)
    static constexpr int SPAN = (PT_TO + Q - 1) * D + 2;   // input samples staged per tile (incl. alignment slack)
    static constexpr int PAIRS = (SPAN + 1) / 2;
    static constexpr int ITERS = (PAIRS + PT_THREADS - 1) / PT_THREADS;
    static constexpr size_t SMEM = (size_t)D * JP * sizeof(float2);
};

// EDGE false: interior tiles -- every staged input lies inside x and x is 16-byte aligned, so the stage is
//             batches of unconditional 128-bit streaming loads issued back to back (7 in flight per thread)
//             before any of them is consumed.  (The first version guarded every load; ncu showed 85 % of the
//             stall samples on the first use of each load, i.e. one load in flight per thread.)
// EDGE true : tiles that touch the carried history or the end of the input; tile index = tile_ids-free mapping
//             e -> (e < n_head ? e : t_hi + (e - n_head)); every load is bounds-checked.
template <int D, int Q, bool ROT, bool EDGE>
__global__ void __launch_bounds__(PT_THREADS)
polyphase_crcf_kernel(const float2* __restrict__ x, const float2* __restrict__ hist, long long n,
                      float2* __restrict__ y, long long first, long long n_out,
                      const __grid_constant__ PolyParams P, long long t_lo, long long t_hi) {
    using S = PolyShape<D, Q>;
    extern __shared__ __align__(16) float2 smem[];
    const int tid = threadIdx.x;
    const long long tile = EDGE ? ((long long)blockIdx.x < t_lo ? (long long)blockIdx.x : t_hi + ((long long)blockIdx.x - t_lo))
                                : (t_lo + (long long)blockIdx.x);
    const long long m0 = tile * PT_TO;
    const int M = P.M;
    // first input index needed by output m0 with the taps padded to Q*D at the OLD end:
    // y[m] = sum_{i'} hr[i'] X[c_m - (Q*D - 1) + i'],  c_m = first + m*D
    const long long B = first + m0 * (long long)D - (long long)(Q * D - 1);
    const long long Beven = B & ~1LL;                 // floor to even (two's complement: correct for negatives)
    const int shift = (int)(B - Beven);               // 0 or 1

    // ---- stage: global -> (rotate) -> de-interleaved shared memory
    float2 c0 = make_float2(1.f, 0.f), c1 = make_float2(1.f, 0.f);
    if constexpr (ROT) {
        // phasor of this thread's first sample pair; later pairs advance by the per-iteration step table
        uint64_t g = P.g0 + (uint64_t)(Beven + 2LL * tid);
        c0 = phasor_from_fix(P.turns_fix * g);
        c1 = cmul(c0, P.rot1);
    }
    const int Hm1 = M - 1;
    // tile-relative index of the pair's first sample, e0 = 2u - shift, tracked as (j, p) = (e0 / D, e0 % D);
    // one staging iteration advances e0 by 2*PT_THREADS
    constexpr int ADV_J = (2 * PT_THREADS) / D, ADV_P = (2 * PT_THREADS) % D;
    int ej, ep;
    {
        const int e0 = 2 * tid - shift + D;           // + D keeps the division non-negative (e0 >= -1)
        ej = e0 / D - 1;
        ep = e0 - (ej + 1) * D;
    }
    auto scatter = [&](float2 a, float2 b) {
        if (ej >= 0 && ej < S::J) smem[ep * S::JP + pad_idx(ej)] = a;
        int j1 = ej, p1 = ep + 1;
        if (p1 == D) { p1 = 0; ++j1; }
        if (j1 >= 0 && j1 < S::J) smem[p1 * S::JP + pad_idx(j1)] = b;
        ej += ADV_J;
        ep += ADV_P;
        if (ep >= D) { ep -= D; ++ej; }
    };
    if constexpr (!EDGE) {
        constexpr int BATCH = 7;
        const float4* x4 = reinterpret_cast<const float4*>(x + Beven) + tid;
#pragma unroll 1
        for (int it0 = 0; it0 < S::ITERS; it0 += BATCH) {
            float4 buf[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k)
                if (it0 + k < S::ITERS) buf[k] = __ldcs(x4 + (it0 + k) * PT_THREADS);
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                if (it0 + k < S::ITERS) {
                    float2 a = make_float2(buf[k].x, buf[k].y), b = make_float2(buf[k].z, buf[k].w);
                    if constexpr (ROT) {
                        const float2 st = P.step[it0 + k];
                        a = cmul(a, cmul(c0, st));
                        b = cmul(b, cmul(c1, st));
                    }
                    scatter(a, b);
                }
            }
        }
    } else {
#pragma unroll 2
        for (int it = 0; it < S::ITERS; ++it) {
            const int u = tid + it * PT_THREADS;      // pair index within the tile
            const long long i0 = Beven + 2LL * u;     // logical input index of the pair's first sample
            float2 a = (i0 >= 0) ? (i0 < n ? __ldg(x + i0) : make_float2(0.f, 0.f))
                                 : ((Hm1 + i0 >= 0) ? __ldg(hist + (Hm1 + i0)) : make_float2(0.f, 0.f));
            const long long i1 = i0 + 1;
            float2 b = (i1 >= 0) ? (i1 < n ? __ldg(x + i1) : make_float2(0.f, 0.f))
                                 : ((Hm1 + i1 >= 0) ? __ldg(hist + (Hm1 + i1)) : make_float2(0.f, 0.f));
            if constexpr (ROT) {
                const float2 st = P.step[it];
                a = cmul(a, cmul(c0, st));
                b = cmul(b, cmul(c1, st));
            }
            scatter(a, b);
        }
    }
    __syncthreads();

    // ---- compute: R outputs per thread, sliding register window per polyphase branch
    float2 acc[PT_R];
#pragma unroll
    for (int r = 0; r < PT_R; ++r) acc[r] = make_float2(0.f, 0.f);
    const float2* sbase = smem + tid * (PT_R + 2);    // pad_idx(tid*8) = tid*10
#pragma unroll
    for (int p = 0; p < D; ++p) {
        const float2* sp = sbase + p * S::JP;
        float2 w[PT_R];
#pragma unroll
        for (int r = 0; r < PT_R; r += 2) {
            float4 v = *reinterpret_cast<const float4*>(sp + r);
            w[r] = make_float2(v.x, v.y);
            w[r + 1] = make_float2(v.z, v.w);
        }
        float2 nx0 = make_float2(0.f, 0.f), nx1 = nx0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if ((q & 1) == 0 && q + 1 < Q) {
                // elements R+q and R+q+1 of this thread's branch window (contiguous in the padded layout)
                float4 v = *reinterpret_cast<const float4*>(sp + pad_idx(PT_R + q));
                nx0 = make_float2(v.x, v.y);
                nx1 = make_float2(v.z, v.w);
            }
            const float h = P.hr[q * D + p];
#pragma unroll
            for (int r = 0; r < PT_R; ++r)
                acc[r] = __ffma2_rn(w[(r + q) % PT_R], make_float2(h, h), acc[r]);
            w[q % PT_R] = (q & 1) ? nx1 : nx0;
        }
    }

    // ---- store: 8 consecutive complex outputs per thread
    const long long mbase = m0 + (long long)tid * PT_R;
    if (mbase + PT_R <= n_out && ((reinterpret_cast<uintptr_t>(y + mbase) & 15) == 0)) {
        float4* y4 = reinterpret_cast<float4*>(y + mbase);
#pragma unroll
        for (int r = 0; r < PT_R; r += 2)
            __stcs(y4 + r / 2, make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y));
    } else {
#pragma unroll
        for (int r = 0; r < PT_R; ++r)
            if (mbase + r < n_out) y[mbase + r] = acc[r];
    }
}

template <int D, int Q, bool ROT>
int launch_shape(const PolyParams& P, const float2* x, const float2* hist, long long n, float2* y, long long first,
                 long long n_out, cudaStream_t s) {
    using S = PolyShape<D, Q>;
    static_assert(Q * D <= PT_MAXTAPS, "taps table too small");
    static_assert(S::ITERS <= PT_MAXIT, "step table too small");
    static bool configured = false;
    auto kern_i = polyphase_crcf_kernel<D, Q, ROT, false>;
    auto kern_e = polyphase_crcf_kernel<D, Q, ROT, true>;
    if (!configured) {
        LRB_CHECK(cudaFuncSetAttribute(kern_i, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        LRB_CHECK(cudaFuncSetAttribute(kern_e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        configured = true;
    }
    const long long tiles = (n_out + PT_TO - 1) / PT_TO;
    // interior tiles: staged span [Beven, Beven + 2*ITERS*THREADS) inside [0, n), x 16-byte aligned
    long long t_lo = 0, t_hi = 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const long long span = 2LL * S::ITERS * PT_THREADS;
        // Beven(t) >= B(t) - 1,  B(t) = first + t*TO*D - (Q*D - 1)
        const long long step = (long long)PT_TO * D;
        const long long need_lo = (long long)(Q * D - 1) + 1 - first;            // B(t) - 1 >= 0
        t_lo = need_lo <= 0 ? 0 : (need_lo + step - 1) / step;
        const long long lim = n - span - first + (long long)(Q * D - 1);         // B(t) + span <= n
        t_hi = lim < 0 ? 0 : lim / step + 1;
        if (t_hi > tiles) t_hi = tiles;
        if (t_lo > t_hi) t_lo = t_hi;
    }
    const long long n_int = t_hi - t_lo, n_edge = tiles - n_int;
    if (n_int > 0) {
        kern_i<<<(unsigned)n_int, PT_THREADS, S::SMEM, s>>>(x, hist, n, y, first, n_out, P, t_lo, t_hi);
        count_launch();
    }
    if (n_edge > 0) {
        kern_e<<<(unsigned)n_edge, PT_THREADS, S::SMEM, s>>>(x, hist, n, y, first, n_out, P, t_lo, t_hi);
        count_launch();
    }
    LRB_CHECK(cudaGetLastError());
    return 1;
}

}  // namespace

struct PolyTaps {
    int M, D, Q;
    PolyParams P;
};

static int shape_q(int M, int D) {
    // instantiated (D, Q) shapes: Q = ceil(Mpad / D) for Mpad in {64, 128}
    if (D != 2 && D != 3 && D != 4 && D != 5 && D != 8 && D != 10) return 0;
    if (M <= 64) return (64 + D - 1) / D;
    if (M <= 128) return (128 + D - 1) / D;
    return 0;
}

PolyTaps* polyphase_prepare(const float* taps, int M, int D, double turns_per_sample) {
    int Q = shape_q(M, D);
    if (!Q) return nullptr;
    PolyTaps* p = new (std::nothrow) PolyTaps();
    if (!p) return nullptr;
    p->M = M; p->D = D; p->Q = Q;
    std::memset(&p->P, 0, sizeof(p->P));
    // hr[i'] multiplies X[c - (Q*D-1) + i']  =>  hr[i'] = h[Q*D-1-i'] (zero for tap index >= M)
    for (int i = 0; i < Q * D; ++i) {
        int k = Q * D - 1 - i;
        p->P.hr[i] = (k < M) ? taps[k] : 0.0f;
    }
    p->P.M = M;
    const double two_pi = 6.283185307179586476925286766559;
    p->P.turns_fix = turns_to_fix(turns_per_sample);
    // staging-step phasors from the SAME fixed-point turns the kernel uses for the absolute phase
    const double tq = std::ldexp((double)p->P.turns_fix, -64);
    for (int it = 0; it < PT_MAXIT; ++it) {
        double a = tq * (double)(2 * PT_THREADS) * (double)it;
        a -= std::floor(a);
        p->P.step[it] = make_float2((float)std::cos(two_pi * a), (float)std::sin(two_pi * a));
    }
    p->P.rot1 = make_float2((float)std::cos(two_pi * tq), (float)std::sin(two_pi * tq));
    return p;
}

void polyphase_release(PolyTaps* p) { delete p; }

#define LRB_SHAPE(DD, QQ)                                                                                   \
    if (p->D == DD && p->Q == QQ)                                                                           \
        return rotate ? launch_shape<DD, QQ, true>(P, x, hist, n, y, first, n_out, s)                       \
                      : launch_shape<DD, QQ, false>(P, x, hist, n, y, first, n_out, s);

int launch_polyphase_crcf(const PolyTaps* p, const float2* x, const float2* hist, long long n, float2* y,
                          long long first, long long n_out, bool rotate, uint64_t turns_fix, uint64_t g0,
                          cudaStream_t s) {
    if (!p) return 0;
    if (n_out <= 0) return 1;
    PolyParams P = p->P;
    P.g0 = g0;
    (void)turns_fix;
    LRB_SHAPE(2, 32) LRB_SHAPE(2, 64)
    LRB_SHAPE(3, 22) LRB_SHAPE(3, 43)
    LRB_SHAPE(4, 16) LRB_SHAPE(4, 32)
    LRB_SHAPE(5, 13) LRB_SHAPE(5, 26)
    LRB_SHAPE(8, 8) LRB_SHAPE(8, 16)
    LRB_SHAPE(10, 7) LRB_SHAPE(10, 13)
    return 0;
}

// ---------------------------------------------------------------------------------------------
// TunerBlock: Rotator -> FIR(crcf) -> Downsampler as one stage of the graph
// ---------------------------------------------------------------------------------------------
struct TunerBlock : Block {
    int M, D;
    PolyTaps* pt = nullptr;
    void* d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    std::string label;

    TunerBlock(PolyTaps* p) : M(p->M), D(p->D), pt(p) {
        in_size = out_size = 8;
        dev_ptrs = true;
        label = "tuner(" + std::to_string(M) + ",/" + std::to_string(D) + ")";
        name = label.c_str();
    }
    ~TunerBlock() override {
        polyphase_release(pt);
        cudaFree(d_hist[0]);
        cudaFree(d_hist[1]);
    }
    int init() override {
        size_t hb = (size_t)(M > 1 ? M - 1 : 1) * 8;
        for (int i = 0; i < 2; ++i) {
            LRB_CHECK(cudaMalloc(&d_hist[i], hb));
            LRB_CHECK(cudaMemset(d_hist[i], 0, hb));
        }
        return 0;
    }
    size_t max_output(size_t n) const override { return n / D + 1; }
    uint64_t outputs_before(uint64_t idx) const override { return (idx + D - 1) / D; }
    int reset() override {
        consumed = 0;
        cur = 0;
        size_t hb = (size_t)(M > 1 ? M - 1 : 1) * 8;
        LRB_CHECK(cudaMemsetAsync(d_hist[0], 0, hb, ctx().stream));
        LRB_CHECK(cudaMemsetAsync(d_hist[1], 0, hb, ctx().stream));
        return 0;
    }
    int run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) override {
        uint64_t r = consumed % (uint64_t)D;
        long long first = (long long)(((uint64_t)D - r) % (uint64_t)D);
        long long no = ((long long)n > first) ? (((long long)n - first + D - 1) / D) : 0;
        *n_out = (size_t)no;
        if (n == 0) return 0;
        int rc = launch_polyphase_crcf(pt, (const float2*)dx, (const float2*)d_hist[cur], (long long)n, (float2*)dy,
                                       first, no, true, pt->P.turns_fix, consumed, s);
        if (rc <= 0) { if (rc == 0) set_error("tuner: unsupported shape"); return -1; }
        if (M > 1) {
            if (launch_hist_update(dx, (long long)n, d_hist[cur], d_hist[cur ^ 1], M - 1, 8, s) != 0) return -1;
            cur ^= 1;
        }
        consumed += n;
        return 0;
    }
};

Block* make_tuner(double turns_per_sample, const float* taps, int ntaps, int decim) {
    PolyTaps* p = polyphase_prepare(taps, ntaps, decim, turns_per_sample);
    if (!p) return nullptr;      // unsupported shape: the graph keeps the blocks separate
    TunerBlock* t = new (std::nothrow) TunerBlock(p);
    if (!t) { polyphase_release(p); set_error("out of memory"); return nullptr; }
    if (t->init() != 0) { delete t; return nullptr; }
    return t;
}

}  // namespace lrb
