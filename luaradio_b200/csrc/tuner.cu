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

// One tile = PT_TO complex filter outputs starting at output index m0.  With the fused discriminator
// (DISC) consecutive tiles overlap by ONE output (tile stride PT_TO - 1): slot 0 of a tile is the output just
// before its first discriminator output, so y[m-1] is always in a neighbouring register (or one shuffle /
// one shared-memory word away) and never has to be fetched from another CTA.
//
// Interior kernel (EDGE = false): persistent CTAs, grid-stride over the interior tiles.  Staging is batches
// of unconditional 128-bit streaming loads; the first two batches of the NEXT tile are issued before the
// compute phase of the current one and stay in registers across it, so HBM latency overlaps the FFMA2 loop
// (ncu on the first version: 85 % of stall samples on the first use of a guarded load; on the second, still
// ~32 % -- hence the cross-tile prefetch).  Edge kernel (EDGE = true): the few tiles that touch the carried
// history or the end of the input; every load bounds-checked, one tile per CTA.
constexpr int PT_BATCH = 7;

template <int D, int Q, bool ROT, bool DISC>
struct TileCtx {
    long long m0, B, Beven;
    int shift;
};

template <int D, int Q, bool ROT, bool DISC, bool EDGE>
__global__ void __launch_bounds__(PT_THREADS, 4)
polyphase_crcf_kernel(const float2* __restrict__ x, const float2* __restrict__ hist, long long n,
                      void* __restrict__ yv, long long first, long long n_out,
                      const __grid_constant__ PolyParams P, long long t_lo, long long t_hi,
                      const float2* __restrict__ prev_in, float2* __restrict__ prev_out, float inv_gain) {
    using S = PolyShape<D, Q>;
    constexpr int TS = DISC ? PT_TO - 1 : PT_TO;       // tile stride in outputs
    constexpr int NPRE = S::ITERS < PT_BATCH ? S::ITERS : PT_BATCH;
    extern __shared__ __align__(16) float2 smem[];
    __shared__ float2 s_edge[PT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M;
    const int Hm1 = M - 1;
    constexpr int ADV_J = (2 * PT_THREADS) / D, ADV_P = (2 * PT_THREADS) % D;

    auto tile_of = [&](long long idx) -> long long {
        return EDGE ? (idx < t_lo ? idx : t_hi + (idx - t_lo)) : (t_lo + idx);
    };
    const long long n_work = EDGE ? 0 : (t_hi - t_lo);
    long long widx = blockIdx.x;
    float4 pre[NPRE];
    auto first_input = [&](long long tile) -> long long {
        // first input index needed by slot 0 of the tile, taps padded to Q*D at the OLD end:
        // y[m] = sum_{i'} hr[i'] X[c_m - (Q*D - 1) + i'],  c_m = first + m*D,  m = tile*TS - (DISC ? 1 : 0)
        const long long m0 = tile * TS - (DISC ? 1 : 0);
        return first + m0 * (long long)D - (long long)(Q * D - 1);
    };
    if constexpr (!EDGE) {
        if (widx < n_work) {
            const long long Bev = first_input(tile_of(widx)) & ~1LL;
            const float4* x4 = reinterpret_cast<const float4*>(x + Bev) + tid;
#pragma unroll
            for (int k = 0; k < NPRE; ++k) pre[k] = __ldcs(x4 + k * PT_THREADS);
        }
    }

    for (;;) {
        if constexpr (!EDGE) { if (widx >= n_work) break; }
        const long long tile = tile_of(widx);
        const long long m0 = tile * TS - (DISC ? 1 : 0);          // output index of slot 0
        const long long B = first_input(tile);
        const long long Beven = B & ~1LL;              // floor to even (two's complement: correct for negatives)
        const int shift = (int)(B - Beven);            // 0 or 1

        // ---- stage: global -> (rotate) -> de-interleaved shared memory
        float2 c0 = make_float2(1.f, 0.f), c1 = make_float2(1.f, 0.f);
        if constexpr (ROT) {
            // phasor of this thread's first sample pair; later pairs advance by the per-iteration step table
            uint64_t g = P.g0 + (uint64_t)(Beven + 2LL * tid);
            c0 = phasor_from_fix(P.turns_fix * g);
            c1 = cmul(c0, P.rot1);
        }
        // tile-relative index of the pair's first sample, e0 = 2u - shift, tracked as (j, p) = (e0 / D, e0 % D)
        int ej, ep;
        {
            const int e0 = 2 * tid - shift + D;        // + D keeps the division non-negative (e0 >= -1)
            ej = e0 / D - 1;
            ep = e0 - (ej + 1) * D;
        }
        auto stage_pair = [&](float4 v, int it) {
            float2 a = make_float2(v.x, v.y), b = make_float2(v.z, v.w);
            if constexpr (ROT) {
                const float2 st = P.step[it];
                a = cmul(a, cmul(c0, st));
                b = cmul(b, cmul(c1, st));
            }
            if (ej >= 0 && ej < S::J) smem[ep * S::JP + pad_idx(ej)] = a;
            int j1 = ej, p1 = ep + 1;
            if (p1 == D) { p1 = 0; ++j1; }
            if (j1 >= 0 && j1 < S::J) smem[p1 * S::JP + pad_idx(j1)] = b;
            ej += ADV_J;
            ep += ADV_P;
            if (ep >= D) { ep -= D; ++ej; }
        };
        if constexpr (!EDGE) {
#pragma unroll
            for (int k = 0; k < NPRE; ++k) stage_pair(pre[k], k);
            const float4* x4 = reinterpret_cast<const float4*>(x + Beven) + tid;
#pragma unroll 1
            for (int it0 = NPRE; it0 < S::ITERS; it0 += PT_BATCH) {
                float4 buf[PT_BATCH];
#pragma unroll
                for (int k = 0; k < PT_BATCH; ++k)
                    if (it0 + k < S::ITERS) buf[k] = __ldcs(x4 + (it0 + k) * PT_THREADS);
#pragma unroll
                for (int k = 0; k < PT_BATCH; ++k)
                    if (it0 + k < S::ITERS) stage_pair(buf[k], it0 + k);
            }
        } else {
#pragma unroll 2
            for (int it = 0; it < S::ITERS; ++it) {
                const long long i0 = Beven + 2LL * (tid + it * PT_THREADS);   // logical index of the pair's first sample
                const float2 a = (i0 >= 0) ? (i0 < n ? __ldg(x + i0) : make_float2(0.f, 0.f))
                                           : ((Hm1 + i0 >= 0) ? __ldg(hist + (Hm1 + i0)) : make_float2(0.f, 0.f));
                const long long i1 = i0 + 1;
                const float2 b = (i1 >= 0) ? (i1 < n ? __ldg(x + i1) : make_float2(0.f, 0.f))
                                           : ((Hm1 + i1 >= 0) ? __ldg(hist + (Hm1 + i1)) : make_float2(0.f, 0.f));
                stage_pair(make_float4(a.x, a.y, b.x, b.y), it);
            }
        }
        __syncthreads();

        // ---- prefetch the first batches of this CTA's next tile; they stay in registers across the compute phase
        if constexpr (!EDGE) {
            const long long nidx = widx + gridDim.x;
            if (nidx < n_work) {
                const long long Bev = first_input(tile_of(nidx)) & ~1LL;
                const float4* x4n = reinterpret_cast<const float4*>(x + Bev) + tid;
#pragma unroll
                for (int k = 0; k < NPRE; ++k) pre[k] = __ldcs(x4n + k * PT_THREADS);
            }
        }

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

        const long long mbase = m0 + (long long)tid * PT_R;     // output index of acc[0]
        if constexpr (!DISC) {
            // ---- store: 8 consecutive complex outputs per thread
            float2* y = reinterpret_cast<float2*>(yv);
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
            __syncthreads();                               // shared tile is reused by the next iteration
        } else {
            // ---- fused FrequencyDiscriminator (frequencydiscriminator.lua:68-88):
            //      d[m] = atan2(im, re of y[m] * conj(y[m-1])) * (1/gain);  slot 0 of the tile is y[m0] with m0 = first disc output - 1
            float* yd = reinterpret_cast<float*>(yv);
            float2 left;                                   // y just before acc[0]
            left.x = __shfl_up_sync(0xffffffffu, acc[PT_R - 1].x, 1);
            left.y = __shfl_up_sync(0xffffffffu, acc[PT_R - 1].y, 1);
            if (lane == 31) s_edge[warp] = acc[PT_R - 1];
            __syncthreads();                               // also fences the shared tile for the next iteration
            if (lane == 0 && warp > 0) left = s_edge[warp - 1];
            if (tid == 0 && tile == 0) acc[0] = __ldg(prev_in);   // stream state: the previous call's last output
#pragma unroll
            for (int r = 0; r < PT_R; ++r) {
                const long long m = mbase + r;             // acc[r] = y[m]; it yields d[m] unless it is the tile's slot 0
                const float2 cur = acc[r];
                const float2 pv = (r == 0) ? left : acc[r - 1];
                if (!(tid == 0 && r == 0) && m < n_out) {
                    const float re = fmaf(cur.x, pv.x, cur.y * pv.y);
                    const float im = fmaf(cur.y, pv.x, -cur.x * pv.y);
                    yd[m] = atan2f(im, re) * inv_gain;
                }
                if (m == n_out - 1) *prev_out = cur;       // carried to the next call
            }
            // (s_edge is rewritten only after the next iteration's staging barrier)
        }
        if constexpr (EDGE) break;
        widx += gridDim.x;
    }
}

template <int D, int Q, bool ROT, bool DISC>
int launch_shape(const PolyParams& P, const float2* x, const float2* hist, long long n, void* y, long long first,
                 long long n_out, const float2* prev_in, float2* prev_out, float inv_gain, cudaStream_t s) {
    using S = PolyShape<D, Q>;
    static_assert(Q * D <= PT_MAXTAPS, "taps table too small");
    static_assert(S::ITERS <= PT_MAXIT, "step table too small");
    static bool configured = false;
    static int ctas_per_sm = 1;
    auto kern_i = polyphase_crcf_kernel<D, Q, ROT, DISC, false>;
    auto kern_e = polyphase_crcf_kernel<D, Q, ROT, DISC, true>;
    if (!configured) {
        LRB_CHECK(cudaFuncSetAttribute(kern_i, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        LRB_CHECK(cudaFuncSetAttribute(kern_e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        LRB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern_i, PT_THREADS, S::SMEM));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured = true;
    }
    constexpr int TS = DISC ? PT_TO - 1 : PT_TO;
    // DISC: tile t yields discriminator outputs [t*TS, t*TS + TS); its slot 0 is output t*TS - 1
    const long long tiles = (n_out + TS - 1) / TS;
    // interior tiles: staged span [Beven, Beven + 2*ITERS*THREADS) inside [0, n), x 16-byte aligned
    long long t_lo = 0, t_hi = 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const long long span = 2LL * S::ITERS * PT_THREADS;
        const long long step = (long long)TS * D;
        const long long off = first - (DISC ? D : 0) - (long long)(Q * D - 1);   // B(t) = off + t*step
        // Beven >= B - 1 >= 0   and   B + span <= n
        const long long need_lo = 1 - off;
        t_lo = need_lo <= 0 ? 0 : (need_lo + step - 1) / step;
        const long long lim = n - span - off;
        t_hi = lim < 0 ? 0 : lim / step + 1;
        if (t_hi > tiles) t_hi = tiles;
        if (t_lo > t_hi) t_lo = t_hi;
    }
    const long long n_int = t_hi - t_lo, n_edge = tiles - n_int;
    if (n_int > 0) {
        long long grid = (long long)ctx().sm_count * ctas_per_sm;
        if (grid > n_int) grid = n_int;
        kern_i<<<(unsigned)grid, PT_THREADS, S::SMEM, s>>>(x, hist, n, y, first, n_out, P, t_lo, t_hi, prev_in, prev_out, inv_gain);
        count_launch();
    }
    if (n_edge > 0) {
        kern_e<<<(unsigned)n_edge, PT_THREADS, S::SMEM, s>>>(x, hist, n, y, first, n_out, P, t_lo, t_hi, prev_in, prev_out, inv_gain);
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
    // instantiated (D, Q) shapes: Q = ceil(Mpad / D) for Mpad in {64, 128}; D = 1 (plain short FIR): Q in {16, 32}
    if (D == 1) return M <= 16 ? 16 : (M <= 32 ? 32 : 0);
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

#define LRB_SHAPE(DD, QQ)                                                                                        \
    if (p->D == DD && p->Q == QQ) {                                                                              \
        if (disc) return launch_shape<DD, QQ, true, true>(P, x, hist, n, y, first, n_out, prev_in, prev_out, inv_gain, s); \
        return rotate ? launch_shape<DD, QQ, true, false>(P, x, hist, n, y, first, n_out, nullptr, nullptr, 0.f, s)    \
                      : launch_shape<DD, QQ, false, false>(P, x, hist, n, y, first, n_out, nullptr, nullptr, 0.f, s);  \
    }

static int launch_polyphase_any(const PolyTaps* p, const float2* x, const float2* hist, long long n, void* y,
                                long long first, long long n_out, bool rotate, bool disc, uint64_t g0,
                                const float2* prev_in, float2* prev_out, float inv_gain, cudaStream_t s) {
    if (!p) return 0;
    if (n_out <= 0) return 1;
    PolyParams P = p->P;
    P.g0 = g0;
    LRB_SHAPE(1, 16) LRB_SHAPE(1, 32)
    LRB_SHAPE(2, 32) LRB_SHAPE(2, 64)
    LRB_SHAPE(3, 22) LRB_SHAPE(3, 43)
    LRB_SHAPE(4, 16) LRB_SHAPE(4, 32)
    LRB_SHAPE(5, 13) LRB_SHAPE(5, 26)
    LRB_SHAPE(8, 8) LRB_SHAPE(8, 16)
    LRB_SHAPE(10, 7) LRB_SHAPE(10, 13)
    return 0;
}

int launch_polyphase_crcf(const PolyTaps* p, const float2* x, const float2* hist, long long n, float2* y,
                          long long first, long long n_out, bool rotate, uint64_t turns_fix, uint64_t g0,
                          cudaStream_t s) {
    (void)turns_fix;
    return launch_polyphase_any(p, x, hist, n, y, first, n_out, rotate, false, g0, nullptr, nullptr, 0.f, s);
}

// ---------------------------------------------------------------------------------------------
// TunerBlock: Rotator -> FIR(crcf) -> Downsampler as one stage of the graph
// ---------------------------------------------------------------------------------------------
struct TunerBlock : Block {
    int M, D;
    PolyTaps* pt = nullptr;
    void* d_hist[2] = {nullptr, nullptr};
    void* d_prev[2] = {nullptr, nullptr};      // fused discriminator: previous tuner output (ping-pong)
    int cur = 0, pcur = 0;
    bool disc = false;
    float gain = 1.f;
    std::string label;

    TunerBlock(PolyTaps* p, float disc_gain) : M(p->M), D(p->D), pt(p) {
        disc = disc_gain != 0.0f;
        gain = disc_gain;
        in_size = 8;
        out_size = disc ? 4 : 8;
        dev_ptrs = true;
        label = std::string(disc ? "tuner+discrim(" : "tuner(") + std::to_string(M) + ",/" + std::to_string(D) + ")";
        name = label.c_str();
    }
    ~TunerBlock() override {
        polyphase_release(pt);
        for (int i = 0; i < 2; ++i) { cudaFree(d_hist[i]); cudaFree(d_prev[i]); }
    }
    int init() override {
        size_t hb = (size_t)(M > 1 ? M - 1 : 1) * 8;
        for (int i = 0; i < 2; ++i) {
            LRB_CHECK(cudaMalloc(&d_hist[i], hb));
            LRB_CHECK(cudaMemset(d_hist[i], 0, hb));
            LRB_CHECK(cudaMalloc(&d_prev[i], 8));
            LRB_CHECK(cudaMemset(d_prev[i], 0, 8));
        }
        return 0;
    }
    size_t max_output(size_t n) const override { return n / D + 1; }
    uint64_t outputs_before(uint64_t idx) const override { return (idx + D - 1) / D; }
    int reset() override {
        consumed = 0;
        cur = pcur = 0;
        size_t hb = (size_t)(M > 1 ? M - 1 : 1) * 8;
        for (int i = 0; i < 2; ++i) {
            LRB_CHECK(cudaMemsetAsync(d_hist[i], 0, hb, ctx().stream));
            LRB_CHECK(cudaMemsetAsync(d_prev[i], 0, 8, ctx().stream));
        }
        return 0;
    }
    int run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) override {
        uint64_t r = consumed % (uint64_t)D;
        long long first = (long long)(((uint64_t)D - r) % (uint64_t)D);
        long long no = ((long long)n > first) ? (((long long)n - first + D - 1) / D) : 0;
        *n_out = (size_t)no;
        if (n == 0) return 0;
        int rc = launch_polyphase_any(pt, (const float2*)dx, (const float2*)d_hist[cur], (long long)n, dy, first, no, true,
                                      disc, consumed, (const float2*)d_prev[pcur], (float2*)d_prev[pcur ^ 1],
                                      disc ? 1.0f / gain : 0.f, s);
        if (rc <= 0) { if (rc == 0) set_error("tuner: unsupported shape"); return -1; }
        if (disc && no > 0) pcur ^= 1;
        if (M > 1) {
            if (launch_hist_update(dx, (long long)n, d_hist[cur], d_hist[cur ^ 1], M - 1, 8, s) != 0) return -1;
            cur ^= 1;
        }
        consumed += n;
        return 0;
    }
};

Block* make_tuner(double turns_per_sample, const float* taps, int ntaps, int decim, float disc_gain) {
    PolyTaps* p = polyphase_prepare(taps, ntaps, decim, turns_per_sample);
    if (!p) return nullptr;      // unsupported shape: the graph keeps the blocks separate
    TunerBlock* t = new (std::nothrow) TunerBlock(p, disc_gain);
    if (!t) { polyphase_release(p); set_error("out of memory"); return nullptr; }
    if (t->init() != 0) { delete t; return nullptr; }
    return t;
}

}  // namespace lrb
