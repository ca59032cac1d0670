// Fast FIR paths (placeholder dispatch until the fused overlap-save kernel lands).
#include "common.cuh"
#include "blocks.h"
#include "../../include/lrb200.h"

namespace lrb {

int FirBlock::fast_init() {
    if (kind == FIR_CRCF && D > 1) poly = polyphase_prepare((const float*)h_taps.data(), M, D, 0.0);
    return 0;
}
void FirBlock::fast_free() { polyphase_release(poly); poly = nullptr; }
int FirBlock::fast_run(const void* dx, size_t n, void* dy, long long first, long long n_out, cudaStream_t s) {
    if (poly && algo != LRB200_FIR_DIRECT + 100)
        return launch_polyphase_crcf(poly, (const float2*)dx, (const float2*)d_hist[cur], (long long)n, (float2*)dy,
                                     first, n_out, false, 0, consumed, s);
    return 0;
}
int FirBlock::set_algorithm(int a) {
    if (a < LRB200_FIR_AUTO || a > LRB200_FIR_FFT) { set_error("fir: unknown algorithm %d", a); return -1; }
    algo = a;
    return 0;
}
int FirBlock::effective_algorithm() const { return LRB200_FIR_DIRECT; }

}  // namespace lrb
