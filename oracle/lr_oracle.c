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
 * BOTH FIR algorithms of the reference are here: the dot-product form above and the overlap-save form
 * (firfilter.lua:320-398, the reference's DEFAULT for > 16 taps when FFTW is present, firfilter.lua:55-62):
 * N = 2^floor(log2(8M)), L = N - M + 1, forward DFT, spectrum multiply (volk_32fc_x2_multiply_32fc), inverse DFT with
 * the 1/N scale (folded into the tap spectrum here).  FFTW is replaced by an in-house float32 FFT (32 x 32 four-step,
 * structure-of-arrays so that gcc vectorises the butterflies); real-input filters pack two consecutive blocks
 * into one complex transform, the same work FFTW's r2c/c2r plans save.  lro_wbfm_chain(..., use_fft) selects.
 *
 * Parallelism: the reference runs one process per block (7 blocks, <= 7 cores busy).  This port instead
 * splits the stream into time chunks with a lead-in halo and runs the whole chain per chunk on every USABLE core
 * (persistent pthread pool, dynamic chunk queue, per-thread scratch allocated once; usable = affinity mask capped by
 * the cgroup CPU quota), which is at least as fast; `cores` in bench.py's JSON says how many were used.
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

/* ---- persistent thread pool with a dynamic chunk queue (libgomp is not in this image) -------------------- */
#include <stdio.h>
typedef void (*lro_job_fn)(long c, void* arg);
static struct {
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    pthread_t th[1024];
    int nthreads;              /* workers created so far */
    int active;                /* workers allowed to take part in the current job */
    unsigned long gen;         /* job generation */
    lro_job_fn fn; void* arg;
    long count; long next;     /* next chunk index (under mu: chunks are ~ms of work) */
    int running;               /* workers still inside the current job */
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, 0, NULL, NULL, 0, 0, 0};

static void lro_pool_drain(void) {
    for (;;) {
        pthread_mutex_lock(&g_pool.mu);
        long c = g_pool.next < g_pool.count ? g_pool.next++ : -1;
        pthread_mutex_unlock(&g_pool.mu);
        if (c < 0) return;
        g_pool.fn(c, g_pool.arg);
    }
}
static void* lro_pool_main(void* p) {
    const int id = (int)(long)p;
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.gen == seen || id >= g_pool.active) {
            if (g_pool.gen != seen) seen = g_pool.gen;       /* not invited to this job */
            pthread_cond_wait(&g_pool.cv_start, &g_pool.mu);
        }
        seen = g_pool.gen;
        pthread_mutex_unlock(&g_pool.mu);
        lro_pool_drain();
        pthread_mutex_lock(&g_pool.mu);
        if (--g_pool.running == 0) pthread_cond_signal(&g_pool.cv_done);
        pthread_mutex_unlock(&g_pool.mu);
    }
    return NULL;
}
static void lro_parallel_for(long count, int threads, lro_job_fn fn, void* arg) {
    if (threads > count) threads = (int)count;
    if (threads > 1024) threads = 1024;
    if (threads <= 1) { for (long c = 0; c < count; ++c) fn(c, arg); return; }
    const int workers = threads - 1;                         /* the caller is the last worker */
    pthread_mutex_lock(&g_pool.mu);
    if (g_pool.nthreads < workers) {
        /* pin worker t to the t-th allowed CPU: on these hosts freshly created threads otherwise all start on the
         * creator's CPU and are not migrated within a 0.1 s job (measured: 8 threads, 1.0x speed-up) */
        cpu_set_t allowed;
        int cpus[1024], ncpu = 0;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
            for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; ++c) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
        for (int t = g_pool.nthreads; t < workers; ++t) {
            pthread_attr_t attr;
            pthread_attr_init(&attr);
            if (ncpu > 0 && workers < ncpu) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[(t + 1) % ncpu], &one);
                pthread_attr_setaffinity_np(&attr, sizeof(one), &one);
            }
            if (pthread_create(&g_pool.th[t], &attr, lro_pool_main, (void*)(long)t) != 0)
                pthread_create(&g_pool.th[t], NULL, lro_pool_main, (void*)(long)t);
            pthread_attr_destroy(&attr);
        }
        g_pool.nthreads = workers;
    }
    g_pool.fn = fn; g_pool.arg = arg; g_pool.count = count; g_pool.next = 0;
    g_pool.active = workers; g_pool.running = workers;
    g_pool.gen++;
    pthread_cond_broadcast(&g_pool.cv_start);
    pthread_mutex_unlock(&g_pool.mu);
    lro_pool_drain();
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.running > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
    pthread_mutex_unlock(&g_pool.mu);
}

/* CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota (v2 cpu.max, v1 cfs_quota) */
int lro_usable_cpus(void) {
    int n = 0;
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) n = CPU_COUNT(&allowed);
    if (n <= 0) { long k = sysconf(_SC_NPROCESSORS_ONLN); n = k > 0 ? (int)k : 1; }
    double quota = -1, period = -1;
    FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64];
        if (fscanf(f, "%63s %lf", q, &period) == 2 && strcmp(q, "max") != 0) quota = atof(q);
        fclose(f);
    } else {
        FILE* fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r");
        FILE* fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
        if (fq && fp && fscanf(fq, "%lf", &quota) == 1 && fscanf(fp, "%lf", &period) == 1) { /* ok */ } else quota = -1;
        if (fq) fclose(fq);
        if (fp) fclose(fp);
    }
    if (quota > 0 && period > 0) {
        int q = (int)ceil(quota / period);
        if (q >= 1 && q < n) n = q;
    }
    return n;
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

/* =================================================================================================================
 * Overlap-save FIR (firfilter.lua:320-398) with an in-house FFT standing in for FFTW3f.
 *
 * FFT of N = 1024 = 32 x 32 (four-step): x[32 n1 + n2] as a 32 x 32 matrix, 32-point FFTs down the columns
 * (vectorised across the 32 columns of a row), twiddle W_N^(n2 k1), transpose, 32-point FFTs down the columns again:
 * the result is X[k] in natural order.  Structure-of-arrays (re[], im[]) so that gcc -O3 emits packed FMAs.
 * Other power-of-two sizes run a plain radix-2 (only used by tests with other tap counts).
 * Conventions: forward e^{-j 2 pi nk/N}, inverse e^{+j...} scaled by 1/N (spectrum_utils.lua:206-244, 456-504);
 * here the 1/N is folded into the tap spectrum.
 * ================================================================================================================= */
typedef struct {
    int N;
    float *w32r, *w32i;          /* W_32^j, j < 16 (forward sign) */
    float *twr, *twi;            /* W_N^(n2 k1), [k1][n2], forward sign (N == 1024) */
    float *wNr, *wNi;            /* W_N^j, j < N/2 (generic radix-2) */
} lro_fft_plan;

static lro_fft_plan* lro_fft_plan_new(int N) {
    lro_fft_plan* p = calloc(1, sizeof(*p));
    p->N = N;
    p->w32r = malloc(16 * sizeof(float)); p->w32i = malloc(16 * sizeof(float));
    for (int j = 0; j < 16; ++j) { p->w32r[j] = (float)cos(2 * M_PI * j / 32); p->w32i[j] = (float)(-sin(2 * M_PI * j / 32)); }
    if (N == 1024) {
        p->twr = malloc(1024 * sizeof(float)); p->twi = malloc(1024 * sizeof(float));
        for (int k1 = 0; k1 < 32; ++k1)
            for (int n2 = 0; n2 < 32; ++n2) {
                p->twr[k1 * 32 + n2] = (float)cos(2 * M_PI * (k1 * n2) / 1024.0);
                p->twi[k1 * 32 + n2] = (float)(-sin(2 * M_PI * (k1 * n2) / 1024.0));
            }
    }
    p->wNr = malloc((size_t)(N / 2 + 1) * sizeof(float)); p->wNi = malloc((size_t)(N / 2 + 1) * sizeof(float));
    for (int j = 0; j < N / 2; ++j) { p->wNr[j] = (float)cos(2 * M_PI * j / N); p->wNi[j] = (float)(-sin(2 * M_PI * j / N)); }
    return p;
}
static void lro_fft_plan_free(lro_fft_plan* p) {
    if (!p) return;
    free(p->w32r); free(p->w32i); free(p->twr); free(p->twi); free(p->wNr); free(p->wNi); free(p);
}

static const unsigned char lro_rev32[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                             1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/* 32-point FFTs down the columns of a 32 x 32 SoA matrix, in place; sgn = +1 forward, -1 inverse */
static void lro_fft32_cols(float* restrict re, float* restrict im, const float* wr, const float* wi, float sgn) {
    for (int i = 0; i < 32; ++i) {
        const int j = lro_rev32[i];
        if (j > i) {
            float* restrict a = re + i * 32; float* restrict b = re + j * 32;
            float* restrict c = im + i * 32; float* restrict d = im + j * 32;
            for (int k = 0; k < 32; ++k) { float t = a[k]; a[k] = b[k]; b[k] = t; float u = c[k]; c[k] = d[k]; d[k] = u; }
        }
    }
    for (int len = 2; len <= 32; len <<= 1) {
        const int half = len >> 1, step = 32 / len;
        for (int i = 0; i < 32; i += len)
            for (int j = 0; j < half; ++j) {
                const float c = wr[j * step], s = sgn * wi[j * step];
                float* restrict ar = re + (i + j) * 32; float* restrict ai = im + (i + j) * 32;
                float* restrict br = re + (i + j + half) * 32; float* restrict bi = im + (i + j + half) * 32;
                for (int k = 0; k < 32; ++k) {
                    const float tr = br[k] * c - bi[k] * s, ti = br[k] * s + bi[k] * c;
                    br[k] = ar[k] - tr; bi[k] = ai[k] - ti;
                    ar[k] += tr; ai[k] += ti;
                }
            }
    }
}

static void lro_transpose32(const float* restrict a, float* restrict b) {
    for (int i0 = 0; i0 < 32; i0 += 8)
        for (int j0 = 0; j0 < 32; j0 += 8)
            for (int i = i0; i < i0 + 8; ++i)
                for (int j = j0; j < j0 + 8; ++j) b[j * 32 + i] = a[i * 32 + j];
}

/* in place on (re, im), natural order in and out; tr/ti: scratch of N floats each; inverse is unscaled */
static void lro_fft(const lro_fft_plan* p, float* restrict re, float* restrict im, float* restrict tr, float* restrict ti, int inverse) {
    const int N = p->N;
    const float sgn = inverse ? -1.f : 1.f;
    if (N == 1024) {
        lro_fft32_cols(re, im, p->w32r, p->w32i, sgn);
        for (int k = 0; k < 1024; ++k) {
            const float c = p->twr[k], s = sgn * p->twi[k];
            const float a = re[k], b = im[k];
            re[k] = a * c - b * s;
            im[k] = a * s + b * c;
        }
        lro_transpose32(re, tr);
        lro_transpose32(im, ti);
        lro_fft32_cols(tr, ti, p->w32r, p->w32i, sgn);
        memcpy(re, tr, 1024 * sizeof(float));
        memcpy(im, ti, 1024 * sizeof(float));
        return;
    }
    /* generic radix-2, decimation in time */
    int bits = 0;
    while ((1 << bits) < N) ++bits;
    for (int i = 0; i < N; ++i) {
        int j = 0;
        for (int b = 0; b < bits; ++b) if (i & (1 << b)) j |= 1 << (bits - 1 - b);
        if (j > i) { float t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= N; len <<= 1) {
        const int half = len >> 1, step = N / len;
        for (int i = 0; i < N; i += len)
            for (int j = 0; j < half; ++j) {
                const float c = p->wNr[j * step], s = sgn * p->wNi[j * step];
                const float xr = re[i + j + half] * c - im[i + j + half] * s, xi = re[i + j + half] * s + im[i + j + half] * c;
                re[i + j + half] = re[i + j] - xr; im[i + j + half] = im[i + j] - xi;
                re[i + j] += xr; im[i + j] += xi;
            }
    }
    (void)tr; (void)ti;
}

/* firfilter.lua:320-359 + :361-398.  Streaming state exactly as the reference keeps it: input_block holds the M-1
 * overlap samples followed by `fill` pending input samples; a block is transformed when fill reaches L. */
typedef struct {
    int M, N, L, fill, real_in;
    lro_fft_plan* plan;
    float *Hr, *Hi;              /* tap spectrum / N */
    float *inr, *ini;            /* input block, SoA (ini unused for real input: see pending block) */
    float *wr_, *wi_, *tr, *ti;  /* work */
    float *pend;                 /* real input: a completed block waiting for its partner (N floats incl. overlap) */
    int have_pend;
} lro_firfft;

lro_firfft* lro_firfft_new(const float* taps /* natural order */, int M, int complex_taps, int real_in) {
    lro_firfft* f = calloc(1, sizeof(*f));
    int N = 1;
    while (N * 2 <= 8 * M) N *= 2;                       /* 2^floor(log2(8M)) */
    f->M = M; f->N = N; f->L = N - M + 1; f->real_in = real_in;
    f->plan = lro_fft_plan_new(N);
    size_t nb = (size_t)N * sizeof(float);
    f->Hr = calloc(1, nb); f->Hi = calloc(1, nb); f->inr = calloc(1, nb); f->ini = calloc(1, nb);
    f->wr_ = calloc(1, nb); f->wi_ = calloc(1, nb); f->tr = calloc(1, nb); f->ti = calloc(1, nb); f->pend = calloc(1, nb);
    /* taps DFT in double (the reference transforms float32 taps with FFTW; double keeps this port's own error down) */
    for (int k = 0; k < N; ++k) {
        double ar = 0, ai = 0;
        for (int m = 0; m < M; ++m) {
            const double ang = -2 * M_PI * (double)((long)k * m % N) / N;
            const double tr_ = complex_taps ? taps[2 * m] : taps[m], ti_ = complex_taps ? taps[2 * m + 1] : 0.0;
            ar += tr_ * cos(ang) - ti_ * sin(ang);
            ai += tr_ * sin(ang) + ti_ * cos(ang);
        }
        f->Hr[k] = (float)(ar / N);
        f->Hi[k] = (float)(ai / N);
    }
    return f;
}
void lro_firfft_free(lro_firfft* f) {
    if (!f) return;
    lro_fft_plan_free(f->plan);
    free(f->Hr); free(f->Hi); free(f->inr); free(f->ini); free(f->wr_); free(f->wi_); free(f->tr); free(f->ti); free(f->pend); free(f);
}
void lro_firfft_reset(lro_firfft* f) {
    memset(f->inr, 0, (size_t)f->N * sizeof(float)); memset(f->ini, 0, (size_t)f->N * sizeof(float));
    f->fill = 0; f->have_pend = 0;
}

/* one block through FFT -> multiply (volk_32fc_x2_multiply_32fc) -> IFFT; (wr_, wi_) in/out */
static void firfft_block(lro_firfft* f) {
    const int N = f->N;
    lro_fft(f->plan, f->wr_, f->wi_, f->tr, f->ti, 0);
    for (int k = 0; k < N; ++k) {
        const float a = f->wr_[k], b = f->wi_[k];
        f->wr_[k] = a * f->Hr[k] - b * f->Hi[k];
        f->wi_[k] = a * f->Hi[k] + b * f->Hr[k];
    }
    lro_fft(f->plan, f->wr_, f->wi_, f->tr, f->ti, 1);
}

/* complex input: returns the number of outputs written (a multiple of L), firfilter.lua:361-398 */
long lro_firfft_process_c(lro_firfft* f, const cf32* x, long n, cf32* y) {
    const int M1 = f->M - 1, L = f->L, N = f->N;
    long i = 0, out = 0;
    while (i < n) {
        long len = n - i < L - f->fill ? n - i : L - f->fill;
        float* restrict dr = f->inr + M1 + f->fill; float* restrict di = f->ini + M1 + f->fill;
        for (long k = 0; k < len; ++k) { dr[k] = x[i + k].re; di[k] = x[i + k].im; }
        f->fill += (int)len;
        i += len;
        if (f->fill < L) break;
        memcpy(f->wr_, f->inr, (size_t)N * sizeof(float));
        memcpy(f->wi_, f->ini, (size_t)N * sizeof(float));
        firfft_block(f);
        for (int k = 0; k < L; ++k) { y[out + k].re = f->wr_[M1 + k]; y[out + k].im = f->wi_[M1 + k]; }
        out += L;
        memmove(f->inr, f->inr + (N - M1), (size_t)M1 * sizeof(float));
        memmove(f->ini, f->ini + (N - M1), (size_t)M1 * sizeof(float));
        f->fill = 0;
    }
    return out;
}

/* real input, real taps: two consecutive blocks ride one complex transform (block A in the real lane, block B in the
 * imaginary lane; real taps keep the lanes separate).  A completed block waits in `pend` for its partner, so outputs
 * appear 2L at a time; lro_firfft_flush_r pushes out a waiting block. */
static long firfft_emit_pair(lro_firfft* f, const float* blkB /* NULL: zeros */, float* y, int emitB) {
    const int M1 = f->M - 1, L = f->L, N = f->N;
    memcpy(f->wr_, f->pend, (size_t)N * sizeof(float));
    if (blkB) memcpy(f->wi_, blkB, (size_t)N * sizeof(float)); else memset(f->wi_, 0, (size_t)N * sizeof(float));
    firfft_block(f);
    memcpy(y, f->wr_ + M1, (size_t)L * sizeof(float));
    if (emitB) memcpy(y + L, f->wi_ + M1, (size_t)L * sizeof(float));
    f->have_pend = 0;
    return emitB ? 2L * L : L;
}
long lro_firfft_process_r(lro_firfft* f, const float* x, long n, float* y) {
    const int M1 = f->M - 1, L = f->L, N = f->N;
    long i = 0, out = 0;
    while (i < n) {
        long len = n - i < L - f->fill ? n - i : L - f->fill;
        memcpy(f->inr + M1 + f->fill, x + i, (size_t)len * sizeof(float));
        f->fill += (int)len;
        i += len;
        if (f->fill < L) break;
        if (!f->have_pend) {
            memcpy(f->pend, f->inr, (size_t)N * sizeof(float));
            f->have_pend = 1;
        } else {
            out += firfft_emit_pair(f, f->inr, y + out, 1);
        }
        memmove(f->inr, f->inr + (N - M1), (size_t)M1 * sizeof(float));
        f->fill = 0;
    }
    return out;
}
/* end of stream: zero-pad the partial block and push everything out; writes fill (+ L if a block was waiting) samples */
long lro_firfft_flush_c(lro_firfft* f, cf32* y) {
    const int M1 = f->M - 1, N = f->N, fill = f->fill;
    if (fill == 0) return 0;
    memset(f->inr + M1 + fill, 0, (size_t)(N - M1 - fill) * sizeof(float));
    memset(f->ini + M1 + fill, 0, (size_t)(N - M1 - fill) * sizeof(float));
    memcpy(f->wr_, f->inr, (size_t)N * sizeof(float));
    memcpy(f->wi_, f->ini, (size_t)N * sizeof(float));
    firfft_block(f);
    for (int k = 0; k < fill; ++k) { y[k].re = f->wr_[M1 + k]; y[k].im = f->wi_[M1 + k]; }
    f->fill = 0;
    return fill;
}
long lro_firfft_flush_r(lro_firfft* f, float* y) {
    const int M1 = f->M - 1, L = f->L, N = f->N, fill = f->fill;
    long out = 0;
    if (fill > 0) memset(f->inr + M1 + fill, 0, (size_t)(N - M1 - fill) * sizeof(float));
    if (f->have_pend) {
        float* tmp = malloc((size_t)2 * L * sizeof(float));
        firfft_emit_pair(f, fill > 0 ? f->inr : NULL, tmp, 1);
        memcpy(y, tmp, (size_t)(L + fill) * sizeof(float));
        free(tmp);
        out = L + fill;
    } else if (fill > 0) {
        memcpy(f->pend, f->inr, (size_t)N * sizeof(float));
        float* tmp = malloc((size_t)L * sizeof(float));
        firfft_emit_pair(f, NULL, tmp, 0);
        memcpy(y, tmp, (size_t)fill * sizeof(float));
        free(tmp);
        out = fill;
    }
    f->fill = 0;
    return out;
}

/* ---- examples/rtlsdr_wbfm_mono.lua:12-28 on one contiguous piece, block by block, vectors of `vec` samples
 * (the reference moves 8192-sample vectors between blocks, radio/blocks/sources/zero.lua:30).  use_fft selects the
 * FIR algorithm of both filters (firfilter.lua:55-62).  All scratch lives in a per-thread context allocated once. */
typedef struct {
    double omega;
    const float *t1_rev, *t2_rev; /* tuner / audio lowpass taps, reversed */
    const float *t1, *t2;         /* natural order */
    int M1, M2, D1, D2;
    float gain;
    float b[2], a[2];
    int use_fft;
} lro_wbfm_params;

typedef struct {
    long vec; int M1, M2, use_fft;
    cf32 *rot, *st1, *f1, *d1;
    float *dm, *st2, *f2, *de, *tmp;
    long tmp_cap;
    lro_firfft *ff1, *ff2;
    const float *t1_id, *t2_id;   /* taps the FFT plans were built for */
} wbfm_scratch;
static __thread wbfm_scratch* tl_scratch = NULL;

static wbfm_scratch* wbfm_get_scratch(const lro_wbfm_params* p, long vec) {
    wbfm_scratch* s = tl_scratch;
    if (s && (s->vec != vec || s->M1 != p->M1 || s->M2 != p->M2)) {
        free(s->rot); free(s->st1); free(s->f1); free(s->d1); free(s->dm); free(s->st2); free(s->f2); free(s->de); free(s->tmp);
        lro_firfft_free(s->ff1); lro_firfft_free(s->ff2);
        free(s); s = NULL;
    }
    if (!s) {
        s = calloc(1, sizeof(*s));
        s->vec = vec; s->M1 = p->M1; s->M2 = p->M2;
        const long big = vec + 8 * (long)(p->M1 > p->M2 ? p->M1 : p->M2) + 16;      /* overlap-save emits up to L more per call */
        s->rot = malloc((size_t)vec * sizeof(cf32));
        s->st1 = calloc((size_t)(p->M1 - 1 + vec), sizeof(cf32));
        s->f1 = malloc((size_t)big * sizeof(cf32));
        s->d1 = malloc((size_t)big * sizeof(cf32));
        s->dm = malloc((size_t)big * sizeof(float));
        s->st2 = calloc((size_t)(p->M2 - 1 + big), sizeof(float));
        s->f2 = malloc((size_t)(big + 16 * (long)p->M2) * sizeof(float));
        s->de = malloc((size_t)(big + 16 * (long)p->M2) * sizeof(float));
        tl_scratch = s;
    }
    if (p->use_fft && (!s->ff1 || s->t1_id != p->t1 || s->t2_id != p->t2)) {
        lro_firfft_free(s->ff1); lro_firfft_free(s->ff2);
        s->ff1 = lro_firfft_new(p->t1, p->M1, 0, 0);
        s->ff2 = lro_firfft_new(p->t2, p->M2, 0, 1);
        s->t1_id = p->t1; s->t2_id = p->t2;
    }
    return s;
}

static long wbfm_piece(const lro_wbfm_params* p, const cf32* x, long n, uint64_t n0, float* out, long vec) {
    wbfm_scratch* s = wbfm_get_scratch(p, vec);
    cf32 *rot = s->rot, *st1 = s->st1, *f1 = s->f1, *d1 = s->d1;
    float *dm = s->dm, *st2 = s->st2, *f2 = s->f2, *de = s->de;
    memset(st1, 0, (size_t)(p->M1 - 1) * sizeof(cf32));
    memset(st2, 0, (size_t)(p->M2 - 1) * sizeof(float));
    if (p->use_fft) { lro_firfft_reset(s->ff1); lro_firfft_reset(s->ff2); }
    double phase;
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
    for (long off = 0; off <= n; off += vec) {
        const int last = off + vec >= n;
        long nv = n - off < vec ? n - off : vec;
        long nf1 = nv, n2;
        if (nv > 0) lro_rotator(x + off, rot, nv, p->omega, &phase);
        if (p->use_fft) {
            nf1 = nv > 0 ? lro_firfft_process_c(s->ff1, rot, nv, f1) : 0;
            if (last) nf1 += lro_firfft_flush_c(s->ff1, f1 + nf1);            /* end of the piece: push the tail out */
        } else if (nv > 0) {
            lro_fir_crcf(rot, nv, p->t1_rev, p->M1, st1, f1);
        }
        long n1 = lro_downsample_c(f1, nf1, p->D1, &idx1, d1);
        lro_discriminator(d1, n1, p->gain, &prev, dm);
        if (p->use_fft) {
            n2 = lro_firfft_process_r(s->ff2, dm, n1, f2);
            if (last) n2 += lro_firfft_flush_r(s->ff2, f2 + n2);
        } else {
            lro_fir_rrrf(dm, n1, p->t2_rev, p->M2, st2, f2);
            n2 = n1;
        }
        lro_iir1(f2, n2, p->b, p->a, &xp, &yp, de);
        produced += lro_downsample_f(de, n2, p->D2, &idx2, out + produced);
        if (last) break;
    }
    return produced;
}

typedef struct { const lro_wbfm_params* p; const cf32* x; long n; uint64_t n0; float* out; long per, halo, DD; } wbfm_job;
static void wbfm_chunk(long c, void* arg) {
    wbfm_job* j = (wbfm_job*)arg;
    long start = c * j->per, end = start + j->per < j->n ? start + j->per : j->n;
    long lead = c == 0 ? 0 : (j->halo < start ? j->halo : start);
    wbfm_scratch* s = wbfm_get_scratch(j->p, 8192);
    const long need = (end - start + lead) / j->DD + 4;
    if (s->tmp_cap < need) { free(s->tmp); s->tmp = malloc(sizeof(float) * (size_t)need); s->tmp_cap = need; }
    float* tmp = s->tmp;
    long got = wbfm_piece(j->p, j->x + start - lead, end - start + lead, j->n0 + (uint64_t)(start - lead), tmp, 8192);
    long skip = lead / j->DD;
    memcpy(j->out + start / j->DD, tmp + skip, sizeof(float) * (size_t)(got - skip));
}

/* Whole chain over x[0..n) whose first sample has global index n0 (multiple of D1*D2).
 * threads <= 1: one sequential pass (bit-for-bit the streaming semantics).
 * threads > 1 : time-chunk parallel (2 chunks per thread, taken dynamically); every chunk but the first starts `halo`
 *               samples early from a cold state and discards the corresponding outputs (halo: multiple of D1*D2).
 * use_fft     : 0 = dot-product FIRs (VOLK path, firfilter.lua:111-163), 1 = overlap-save FIRs (:320-398, the
 *               reference's default with FFTW present).
 * Returns the number of float outputs written. */
long lro_wbfm_chain2(const cf32* x, long n, uint64_t n0, const float* taps1, int M1, int D1, double omega, float gain,
                     const float* taps2, int M2, const float* b, const float* a, int D2, float* out, int threads, long halo,
                     int use_fft) {
    /* reversed copies live as long as the process: the per-thread FFT plans are keyed on the tap pointers */
    static const float *k1 = NULL, *k2 = NULL;
    static float *r1 = NULL, *r2 = NULL, *c1 = NULL, *c2 = NULL;
    static int kM1 = 0, kM2 = 0;
    if (k1 != taps1 || k2 != taps2 || kM1 != M1 || kM2 != M2 || !c1 || memcmp(c1, taps1, sizeof(float) * (size_t)M1) || memcmp(c2, taps2, sizeof(float) * (size_t)M2)) {
        /* (leaks the previous set on a tap change: baseline infrastructure, a handful of calls per process) */
        r1 = malloc(sizeof(float) * (size_t)M1); r2 = malloc(sizeof(float) * (size_t)M2);
        c1 = malloc(sizeof(float) * (size_t)M1); c2 = malloc(sizeof(float) * (size_t)M2);
        memcpy(c1, taps1, sizeof(float) * (size_t)M1); memcpy(c2, taps2, sizeof(float) * (size_t)M2);
        for (int k = 0; k < M1; ++k) r1[k] = taps1[M1 - 1 - k];
        for (int k = 0; k < M2; ++k) r2[k] = taps2[M2 - 1 - k];
        k1 = taps1; k2 = taps2; kM1 = M1; kM2 = M2;
    }
    lro_wbfm_params p = {omega, r1, r2, c1, c2, M1, M2, D1, D2, gain, {b[0], b[1]}, {a[0], a[1]}, use_fft};
    const long DD = (long)D1 * D2;
    long total = 0;
    if (threads <= 1) {
        total = wbfm_piece(&p, x, n, n0, out, 8192);
    } else {
        long per = ((n / (2L * threads) + DD - 1) / DD) * DD;
        if (per < 64 * DD) per = 64 * DD;
        long nchunks = (n + per - 1) / per;
        total = (n + DD - 1) / DD;
        wbfm_job job = {&p, x, n, n0, out, per, halo, DD};
        lro_parallel_for(nchunks, threads, wbfm_chunk, &job);
    }
    return total;
}
long lro_wbfm_chain(const cf32* x, long n, uint64_t n0, const float* taps1, int M1, int D1, double omega, float gain,
                    const float* taps2, int M2, const float* b, const float* a, int D2, float* out, int threads, long halo) {
    return lro_wbfm_chain2(x, n, n0, taps1, M1, D1, omega, gain, taps2, M2, b, a, D2, out, threads, halo, 0);
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

/* the same filter by overlap-save (firfilter.lua:361-398), chunk-parallel: each chunk re-reads M-1 samples of history */
typedef struct { const cf32* x; long n; const float* taps; int M; cf32* y; long per; } firfft_job;
static __thread lro_firfft* tl_ff = NULL;
static __thread const float* tl_ff_taps = NULL;
static void firfft_chunk(long c, void* arg) {
    firfft_job* j = (firfft_job*)arg;
    long start = c * j->per, end = start + j->per < j->n ? start + j->per : j->n;
    if (start >= end) return;
    if (!tl_ff || tl_ff_taps != j->taps || tl_ff->M != j->M) {
        lro_firfft_free(tl_ff);
        tl_ff = lro_firfft_new(j->taps, j->M, 0, 0);
        tl_ff_taps = j->taps;
    }
    lro_firfft* f = tl_ff;
    lro_firfft_reset(f);
    const int M1 = j->M - 1;
    long lead = start < M1 ? start : M1;                    /* real history instead of zeros */
    for (long k = 0; k < lead; ++k) { f->inr[M1 - lead + k] = j->x[start - lead + k].re; f->ini[M1 - lead + k] = j->x[start - lead + k].im; }
    long pos = start;
    for (long off = start; off < end; off += 8192) {
        long nv = end - off < 8192 ? end - off : 8192;
        pos += lro_firfft_process_c(f, j->x + off, nv, j->y + pos);
    }
    lro_firfft_flush_c(f, j->y + pos);
}
void lro_fir_crcf_fft_parallel(const cf32* x, long n, const float* taps, int M, cf32* y, int threads) {
    long per = (n + 2L * threads - 1) / (2L * threads);
    if (per < 8192) per = 8192;
    firfft_job job = {x, n, taps, M, y, per};
    lro_parallel_for((n + per - 1) / per, threads, firfft_chunk, &job);
}

/* ---- the synthetic FM IQ of SURVEY.md 8d (oracle/lr_oracle.py: synth_fm_iq, splitmix64), so that the reference arm of
 * bench.py can make its 256 Mi-sample input without the GPU library */
static inline uint64_t lro_splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct { cf32* dst; uint64_t n0; long n; uint32_t seed; double rate, carrier, deviation; float amp, noise; long per; } synth_job;
static void synth_chunk(long c, void* arg) {
    synth_job* j = (synth_job*)arg;
    long start = c * j->per, end = start + j->per < j->n ? start + j->per : j->n;
    const double f1 = 1e3, f2 = 7e3, two_pi = 2 * M_PI;
    const long double cps = (long double)(j->carrier / j->rate);
    for (long i = start; i < end; ++i) {
        const uint64_t g = j->n0 + (uint64_t)i;
        const double t = (double)g / j->rate;
        const double integ = 0.5 * (1.0 - cos(two_pi * f1 * t)) / (two_pi * f1) + 0.5 * (1.0 - cos(two_pi * f2 * t)) / (two_pi * f2);
        long double ct = cps * (long double)g;
        ct -= floorl(ct);
        const double ph = two_pi * (double)ct + two_pi * j->deviation * integ;
        const uint64_t h = lro_splitmix64(g * 2ull + ((uint64_t)j->seed << 40));
        const float wa = (float)((uint32_t)h >> 8) * 1.1920928955078125e-07f - 1.0f;
        const float wb = (float)((uint32_t)(h >> 32) >> 8) * 1.1920928955078125e-07f - 1.0f;
        j->dst[i].re = (float)((double)j->amp * cos(ph) + (double)j->noise * (double)wa);
        j->dst[i].im = (float)((double)j->amp * sin(ph) + (double)j->noise * (double)wb);
    }
}
void lro_synth_fm_iq(cf32* dst, uint64_t n0, long n, uint32_t seed, double rate, double carrier, double deviation,
                     float amp, float noise, int threads) {
    long per = 1 << 16;
    synth_job job = {dst, n0, n, seed, rate, carrier, deviation, amp, noise, per};
    lro_parallel_for((n + per - 1) / per, threads, synth_chunk, &job);
}

int lro_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
