"""Line-by-line numpy model of `rs_poly_kernel<T,L,D,RB>` (luaradio_b200/csrc/resample.cu): the same tile / thread
decomposition, transposed shared-memory tile with the position table, one-tile-ahead register prefetch, circular register
window with the unconditional (clamped) refill, output staging with the odd pitch, and streaming history -- with 8 model
threads per CTA instead of 128.  The CUDA kernel was written from this model; tests/test_resampler_model.py runs it against
the oracle's MultiplyConstant -> Upsampler -> Lowpass [-> Downsampler] on every instantiated (L, D) pair."""
import numpy as np


def ceil_div(a, b):
    return -((-a) // b)


NT, C = 8, 4            # NT: model thread count (kernel: RS_THREADS = 128); C = RS_C


def geometry(L, D, RB, M, banks=16):
    RI = RB * D
    Tt = ceil_div(ceil_div(M, L), C) * C
    HB = max(1, ceil_div(Tt - 1, RI)); H = HB * RI
    k = ceil_div(banks, RI); ntp = HB + NT + 1
    while ntp % banks != k % banks: ntp += 1
    NPOS = RI + Tt + 2 * C
    pos = []
    for kk in range(NPOS):
        ep = H + RI - 1 - kk
        pos.append((ep % RI) * ntp + ep // RI if ep >= 0 else 0)
    RP = (RB * L) | 1
    smem = max(RI * ntp, NT * RP)
    return Tt, H, HB, ntp, pos, smem

def run_call(x, hist, Hn, c0, taps, L, D, RB, scale, grid=3):
    n = len(x); M = len(taps)
    Tt, H, HB, NTP, pos, smem = geometry(L, D, RB, M)
    assert H <= NT * 16   # model only
    hp = np.zeros(Tt * L + 64, np.float32); hp[:M] = taps
    R, RI = RB * L, RB * D
    WN = RI + C - 1; U = (WN + C + C - 1) // C; WP = U * C
    m_lo, m_hi = ceil_div(c0 * L, D), ceil_div((c0 + n) * L, D)
    y = np.zeros(m_hi - m_lo, x.dtype)
    if m_hi <= m_lo: return y
    TO = NT * R; Mbase = (m_lo // L) * L
    ntiles = ceil_div(m_hi - Mbase, TO)
    E = H + NT * RI
    KE = ceil_div(E, NT)          # kernel: RI + 1 with H <= NT
    nb = Tt // C
    cd = np.complex128 if np.iscomplexobj(x) else np.float64
    for cta in range(grid):
        S = np.zeros(smem, x.dtype)
        def fetch(tile):
            lbase = ((Mbase + tile * TO) // L) * D - H - c0
            pre = np.zeros((NT, KE), x.dtype)
            for tid in range(NT):
                for k in range(KE):
                    e = tid + k * NT
                    if e < E:
                        i = lbase + e
                        if i >= 0:
                            if i < n: pre[tid, k] = x[i]
                        elif Hn + i >= 0: pre[tid, k] = hist[Hn + i]
            return pre
        tile = cta
        if tile < ntiles: pre = fetch(tile)
        while tile < ntiles:
            mt = Mbase + tile * TO
            for tid in range(NT):
                for k in range(KE):
                    e = tid + k * NT
                    if e < E: S[(e % RI) * NTP + e // RI] = pre[tid, k] * np.float32(scale)
            if tile + grid < ntiles: pre = fetch(tile + grid)
            accs = np.zeros((NT, R), cd)
            for tid in range(NT):
                W = [None] * WP
                for j in range(WN): W[j] = S[tid + pos[RI - 1 + C - 1 - j]]
                for tb0 in range(0, nb, U):
                    for u in range(U):
                        tb = tb0 + u
                        if tb < nb:
                            for j in range(C):
                                W[(j - (u + 1) * C) % WP] = S[tid + pos[RI - 1 + (tb + 1) * C + C - 1 - j]]
                            for s in range(C):
                                for r in range(R):
                                    accs[tid, r] += hp[(tb * C + s) * L + (r * D) % L] * W[((r * D) // L - s + C - 1 - u * C) % WP]
            RP = R | 1
            for tid in range(NT):
                for r in range(R): S[tid * RP + r] = accs[tid, r]
            for tid in range(NT):
                for o in range(tid, NT * R, NT):
                    m = mt + o
                    if m_lo <= m < m_hi: y[m - m_lo] = S[(o // R) * RP + o % R]
            tile += grid
    return y

def stream(x, taps, L, D, RB, scale, chunks):
    Tt = geometry(L, D, RB, len(taps))[0]
    Hn = Tt
    hist = np.zeros(Hn, x.dtype); c0 = 0; outs = []
    for ch in chunks:
        xc = x[c0:c0 + ch]
        outs.append(run_call(xc, hist, Hn, c0, taps, L, D, RB, scale))
        hist = np.concatenate([hist, xc])[-Hn:]
        c0 += ch
    return np.concatenate(outs)

def rs_rb(L, D):
    if D == 1: return 8 if L == 2 else (4 if L <= 4 else (3 if L == 5 else 2))
    return {(2,3):4,(2,5):4,(3,2):4,(3,4):3,(3,5):3,(4,3):3,(4,5):3,(5,2):3,(5,3):3,(5,4):3,(7,5):2}[(L,D)]
