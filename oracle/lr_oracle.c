/*
 * C restatement of the reference's CPU path for the WBFM-mono chain -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and --impl reference) may load this;
 * nothing under luaradio_b200/ does.  It exists because the reference stack (LuaJIT + VOLK 2.1.0 +
 * liquid-dsp 1.3.2 + FFTW3f) is not installed here and cannot be built without network access.
 *
 * It follows the reference's VOLK code path block by block (file:line relative to the reference root):
 *   rotator       radio/blocks/signal/frequencytranslator.lua:26-53    (VOLK rotator: recurrent phasor)
 *   fir_crcf      radio/blocks/signal/firfilter.lua:129-145            (history + one dot product per INPUT sample,
 *                                                                      volk_32fc_32f_dot_prod_32fc)
 *   fir_rrrf      radio/blocks/signal/firfilter.lua:147-163
 *   downsample    radio/blocks/signal/downsampler.lua:40-56
 *   discriminator radio/blocks/signal/frequencydiscriminator.lua:47-64 (multiply-conjugate, atan2, 1/gain)
 *   iir1          radio/blocks/signal/iirfilter.lua:113-179            (direct-form recurrence)
 * As in the reference, the filters compute EVERY output and the downsampler then discards D-1 of D
 * (composites/tuner.lua:40-47).  Dot products are plain loops that gcc -O3 -march=native vectorises with
 * AVX-512 FMA, standing in for VOLK's hand-written SIMD kernels.
 *
 * Parallelism: the reference runs one process per block (7 blocks, <= 7 cores busy).  This port instead
 * splits the stream into time chunks with a lead-in halo and runs the whole chain per chunk on every core
 * (pthreads), which is at least as fast on an 8-core host; `cores` in bench.py's JSON says how many were used.
 *
 * Pinned against the numpy oracle (itself pinned on the reference's golden vectors) by tests/test_oracle_c.py.
 */
#define _GNU_SOURCE
#include <sched.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* minimal parallel-for over chunk indices with pthreads (libgomp is not in this image) */
typedef void (*lro_job_fn)(long c, void* arg);
typedef struct { lro_job_fn fn; void* arg; long begin, end; } lro_job;
static void* lro_job_main(void* p) {
    lro_job* j = (lro_job*)p;
    for (long c = j->begin; c < j->end; ++c) j->fn(c, j->arg);
    return NULL;
}
static void lro_parallel_for(long count, int threads, lro_job_fn fn, void* arg) {
    if (threads > count) threads = (int)count;
    if (threads <= 1) { for (long c = 0; c < count; ++c) fn(c, arg); return; }
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)threads);
    lro_job* jobs = malloc(sizeof(lro_job) * (size_t)threads);
    long per = (count + threads - 1) / threads;
    /* pin worker t to the t-th allowed CPU: on these hosts freshly created threads otherwise all start on the
     * creator's CPU and are not migrated within a 0.1 s job (measured: 8 threads, 1.0x speed-up) */
    cpu_set_t allowed;
    int cpus[1024], ncpu = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; ++c) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
    for (int t = 0; t < threads; ++t) {
        jobs[t].fn = fn; jobs[t].arg = arg;
        jobs[t].begin = t * per; jobs[t].end = (t + 1) * per < count ? (t + 1) * per : count;
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        if (ncpu > 0) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[t % ncpu], &one);
            pthread_attr_setaffinity_np(&attr, sizeof(one), &one);
        }
        if (pthread_create(&th[t], &attr, lro_job_main, &jobs[t]) != 0) pthread_create(&th[t], NULL, lro_job_main, &jobs[t]);
        pthread_attr_destroy(&attr);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct { float re, im; } cf32;

/* ---- frequencytranslator.lua:26-53 (VOLK path): volk_32fc_s32fc_x2_rotator_32fc -- a recurrent complex phasor,
 * phase *= phase_inc per sample, re-normalised every 512 samples as VOLK 2.x does.  (The pure-Lua path,
 * :93-110, calls cos/sin per sample and is ~10x slower; the reference's test tolerance for this block is 1e-5.)
 * The start phase of a chunk comes from the exact turn fraction of its global index. */
void lro_rotator(const cf32* x, cf32* y, long n, double omega, double* phase) {
    float pr = (float)cos(*phase), pi = (float)sin(*phase);
    const float ir = (float)cos(omega), ii = (float)sin(omega);
    for (long i0 = 0; i0 < n; i0 += 512) {
        const long i1 = i0 + 512 < n ? i0 + 512 : n;
        for (long i = i0; i < i1; ++i) {
            const float xr = x[i].re, xi = x[i].im;
            y[i].re = xr * pr - xi * pi;
            y[i].im = xr * pi + xi * pr;
            const float nr = pr * ir - pi * ii;
            pi = pr * ii + pi * ir;
            pr = nr;
        }
        const float inv = 1.0f / sqrtf(pr * pr + pi * pi);
        pr *= inv;
        pi *= inv;
    }
    /* carry the phase exactly: the recurrence above is only used within a vector */
    double ph = *phase + omega * (double)n;
    ph -= 2 * M_PI * floor(ph / (2 * M_PI));
    *phase = ph;
}

/* ---- firfilter.lua:129-145: state = [last M-1 inputs | x]; out[i] = dot(state[i..i+M), reversed taps).
 * The complex-by-real dot product runs over the interleaved float view with each tap duplicated (t0,t0,t1,t1,..)
 * into 16 independent partial sums, the shape VOLK's volk_32fc_32f_dot_prod_32fc SIMD kernels use, so that
 * gcc emits packed FMAs (the naive re/im loop does not vectorise). */
#define LRO_LANES 16
void lro_fir_crcf(const cf32* x, long n, const float* taps_rev, int M, cf32* state /* M-1+n */, cf32* y) {
    memcpy(state + (M - 1), x, (size_t)n * sizeof(cf32));
    const int M2 = 2 * M;
    const int M2v = M2 / LRO_LANES * LRO_LANES;
    float* td = malloc(sizeof(float) * (size_t)(M2 + LRO_LANES));
    for (int k = 0; k < M; ++k) { td[2 * k] = taps_rev[k]; td[2 * k + 1] = taps_rev[k]; }
    for (long i = 0; i < n; ++i) {
        const float* s = (const float*)(state + i);
        float acc[LRO_LANES];
        for (int l = 0; l < LRO_LANES; ++l) acc[l] = 0.f;
        for (int k = 0; k < M2v; k += LRO_LANES)
            for (int l = 0; l < LRO_LANES; ++l) acc[l] += s[k + l] * td[k + l];
        float ar = 0.f, ai = 0.f;
        for (int l = 0; l < LRO_LANES; l += 2) { ar += acc[l]; ai += acc[l + 1]; }
        for (int k = M2v; k < M2; k += 2) { ar += s[k] * td[k]; ai += s[k + 1] * td[k + 1]; }
        y[i].re = ar;
        y[i].im = ai;
    }
    free(td);
    memmove(state, state + n, (size_t)(M - 1) * sizeof(cf32));
}

/* ---- firfilter.lua:147-163 */
void lro_fir_rrrf(const float* x, long n, const float* taps_rev, int M, float* state /* M-1+n */, float* y) {
    memcpy(state + (M - 1), x, (size_t)n * sizeof(float));
    const int Mv = M / LRO_LANES * LRO_LANES;
    for (long i = 0; i < n; ++i) {
        const float* s = state + i;
        float acc[LRO_LANES];
        for (int l = 0; l < LRO_LANES; ++l) acc[l] = 0.f;
        for (int k = 0; k < Mv; k += LRO_LANES)
            for (int l = 0; l < LRO_LANES; ++l) acc[l] += s[k + l] * taps_rev[k + l];
        float a = 0.f;
        for (int l = 0; l < LRO_LANES; ++l) a += acc[l];
        for (int k = Mv; k < M; ++k) a += s[k] * taps_rev[k];
        y[i] = a;
    }
    memmove(state, state + n, (size_t)(M - 1) * sizeof(float));
}

/* ---- downsampler.lua:40-56 */
long lro_downsample_c(const cf32* x, long n, int D, long* index, cf32* y) {
    long m = 0, idx = *index;
    for (; idx < n; idx += D) y[m++] = x[idx];
    *index = idx - n;
    return m;
}
long lro_downsample_f(const float* x, long n, int D, long* index, float* y) {
    long m = 0, idx = *index;
    for (; idx < n; idx += D) y[m++] = x[idx];
    *index = idx - n;
    return m;
}

/* ---- frequencydiscriminator.lua:47-64 */
void lro_discriminator(const cf32* x, long n, float gain, cf32* prev, float* y) {
    cf32 p = *prev;
    float inv = 1.0f / gain;
    for (long i = 0; i < n; ++i) {
        float re = x[i].re * p.re + x[i].im * p.im;
        float im = x[i].im * p.re - x[i].re * p.im;
        y[i] = atan2f(im, re) * inv;
        p = x[i];
    }
    *prev = p;
}

/* ---- iirfilter.lua:147-179 with nb = na = 2 */
void lro_iir1(const float* x, long n, const float* b, const float* a, float* xprev, float* yprev, float* y) {
    float xp = *xprev, yp = *yprev;
    for (long i = 0; i < n; ++i) {
        float v = (b[0] * x[i] + b[1] * xp - a[1] * yp) / a[0];
        xp = x[i];
        yp = v;
        y[i] = v;
    }
    *xprev = xp;
    *yprev = yp;
}

/* ---- examples/rtlsdr_wbfm_mono.lua:12-28 on one contiguous piece, block by block, vectors of `vec` samples
 * (the reference moves 8192-sample vectors between blocks, radio/blocks/sources/zero.lua:30). */
typedef struct {
    double omega;
    const float *t1_rev, *t2_rev; /* tuner / audio lowpass taps, reversed */
    int M1, M2, D1, D2;
    float gain;
    float b[2], a[2];
} lro_wbfm_params;

static long wbfm_piece(const lro_wbfm_params* p, const cf32* x, long n, uint64_t n0, float* out, long vec) {
    cf32* rot = malloc((size_t)vec * sizeof(cf32));
    cf32* st1 = calloc((size_t)(p->M1 - 1 + vec), sizeof(cf32));
    cf32* f1 = malloc((size_t)vec * sizeof(cf32));
    cf32* d1 = malloc((size_t)(vec / p->D1 + 2) * sizeof(cf32));
    float* dm = malloc((size_t)(vec / p->D1 + 2) * sizeof(float));
    float* st2 = calloc((size_t)(p->M2 - 1 + vec / p->D1 + 2), sizeof(float));
    float* f2 = malloc((size_t)(vec / p->D1 + 2) * sizeof(float));
    float* de = malloc((size_t)(vec / p->D1 + 2) * sizeof(float));
    double phase = fmod(p->omega * (double)(n0 % 1102500000ull), 2 * M_PI);   /* chunk start phase */
    {   /* exact start phase from the turn fraction of the global index */
        long double turns = (long double)p->omega / (2 * M_PIl) * (long double)n0;
        turns -= floorl(turns);
        phase = (double)(turns * 2 * M_PIl);
    }
    long idx1 = (long)((p->D1 - (long)(n0 % (uint64_t)p->D1)) % p->D1), idx2 = 0;
    {
        uint64_t m0 = (n0 + (uint64_t)p->D1 - 1) / (uint64_t)p->D1;         /* tuner outputs before n0 */
        idx2 = (long)((p->D2 - (long)(m0 % (uint64_t)p->D2)) % p->D2);
    }
    cf32 prev = {0.f, 0.f};
    float xp = 0.f, yp = 0.f;
    long produced = 0;
    for (long off = 0; off < n; off += vec) {
        long nv = n - off < vec ? n - off : vec;
        lro_rotator(x + off, rot, nv, p->omega, &phase);
        lro_fir_crcf(rot, nv, p->t1_rev, p->M1, st1, f1);
        long n1 = lro_downsample_c(f1, nv, p->D1, &idx1, d1);
        lro_discriminator(d1, n1, p->gain, &prev, dm);
        lro_fir_rrrf(dm, n1, p->t2_rev, p->M2, st2, f2);
        lro_iir1(f2, n1, p->b, p->a, &xp, &yp, de);
        produced += lro_downsample_f(de, n1, p->D2, &idx2, out + produced);
    }
    free(rot); free(st1); free(f1); free(d1); free(dm); free(st2); free(f2); free(de);
    return produced;
}

typedef struct { const lro_wbfm_params* p; const cf32* x; long n; uint64_t n0; float* out; long per, halo, DD; } wbfm_job;
static void wbfm_chunk(long c, void* arg) {
    wbfm_job* j = (wbfm_job*)arg;
    long start = c * j->per, end = start + j->per < j->n ? start + j->per : j->n;
    long lead = c == 0 ? 0 : (j->halo < start ? j->halo : start);
    float* tmp = malloc(sizeof(float) * (size_t)((end - start + lead) / j->DD + 4));
    long got = wbfm_piece(j->p, j->x + start - lead, end - start + lead, j->n0 + (uint64_t)(start - lead), tmp, 8192);
    long skip = lead / j->DD;
    memcpy(j->out + start / j->DD, tmp + skip, sizeof(float) * (size_t)(got - skip));
    free(tmp);
}

/* Whole chain over x[0..n) whose first sample has global index n0 (multiple of D1*D2).
 * threads <= 1: one sequential pass (bit-for-bit the streaming semantics).
 * threads > 1 : time-chunk parallel; every chunk but the first starts `halo` samples early from a cold
 *               state and discards the corresponding outputs (halo must be a multiple of D1*D2).
 * Returns the number of float outputs written. */
long lro_wbfm_chain(const cf32* x, long n, uint64_t n0, const float* taps1, int M1, int D1, double omega, float gain,
                    const float* taps2, int M2, const float* b, const float* a, int D2, float* out, int threads, long halo) {
    float* r1 = malloc(sizeof(float) * (size_t)M1);
    float* r2 = malloc(sizeof(float) * (size_t)M2);
    for (int k = 0; k < M1; ++k) r1[k] = taps1[M1 - 1 - k];
    for (int k = 0; k < M2; ++k) r2[k] = taps2[M2 - 1 - k];
    lro_wbfm_params p = {omega, r1, r2, M1, M2, D1, D2, gain, {b[0], b[1]}, {a[0], a[1]}};
    const long DD = (long)D1 * D2;
    long total = 0;
    if (threads <= 1) {
        total = wbfm_piece(&p, x, n, n0, out, 8192);
    } else {
        long per = ((n / threads + DD - 1) / DD) * DD;
        long nchunks = (n + per - 1) / per;
        total = (n + DD - 1) / DD;
        wbfm_job job = {&p, x, n, n0, out, per, halo, DD};
        lro_parallel_for(nchunks, threads, wbfm_chunk, &job);
    }
    free(r1); free(r2);
    return total;
}

typedef struct { const cf32* x; long n; const float* r; int M; cf32* y; long per; } fir_job;
static void fir_chunk(long c, void* arg) {
    fir_job* j = (fir_job*)arg;
    long start = c * j->per, end = start + j->per < j->n ? start + j->per : j->n;
    if (start >= end) return;
    const long vec = 8192;
    int M = j->M;
    cf32* st = calloc((size_t)(M - 1 + vec), sizeof(cf32));
    long lead = start < M - 1 ? start : M - 1;              /* real history instead of zeros */
    memcpy(st + (M - 1 - lead), j->x + start - lead, (size_t)lead * sizeof(cf32));
    for (long off = start; off < end; off += vec) {
        long nv = end - off < vec ? end - off : vec;
        lro_fir_crcf(j->x + off, nv, j->r, M, st, j->y + off);
    }
    free(st);
}

/* 128-tap complex-input real-tap FIR alone (BASELINE config 0), chunk-parallel the same way. */
void lro_fir_crcf_parallel(const cf32* x, long n, const float* taps, int M, cf32* y, int threads) {
    float* r = malloc(sizeof(float) * (size_t)M);
    for (int k = 0; k < M; ++k) r[k] = taps[M - 1 - k];
    long per = (n + threads - 1) / threads;
    fir_job job = {x, n, r, M, y, per};
    lro_parallel_for(threads, threads, fir_chunk, &job);
    free(r);
}

int lro_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
