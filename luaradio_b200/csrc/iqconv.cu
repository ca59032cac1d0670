// Source boundary: IQFileSource sample formats -> ComplexFloat32 on the device, so that a
// host-fed graph moves the file's own bytes over PCIe (2 B/sample for RTL-SDR style u8 IQ instead of 8).
//
// Reference: radio/blocks/sources/iqfile.lua:96-108 (byte swap loop, then (value - offset) / scale per component) with
// the table radio/utilities/format_utils.lua:82-97.  The reference computes in double and stores float32; here
// 8/16-bit formats use correctly rounded float32 (operands exact, one division), 32-bit and f64 formats double.
#include "common.cuh"
#include "blocks.h"

#include <cstring>
#include <new>

namespace lrb {

namespace {

enum IqFmt { F_U8, F_S8, F_U16, F_S16, F_U32, F_S32, F_F32, F_F64 };

__device__ __forceinline__ uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
    return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32));
}

template <int FMT, bool SWAP>
__device__ __forceinline__ float conv_one(const unsigned char* p) {
    if constexpr (FMT == F_U8) return __fdiv_rn((float)p[0] - 127.5f, 127.5f);
    if constexpr (FMT == F_S8) return __fdiv_rn((float)(signed char)p[0], 127.5f);
    if constexpr (FMT == F_U16 || FMT == F_S16) {
        uint16_t v = *reinterpret_cast<const uint16_t*>(p);
        if (SWAP) v = bswap16(v);
        return FMT == F_U16 ? __fdiv_rn((float)v - 32767.5f, 32767.5f) : __fdiv_rn((float)(int16_t)v, 32767.5f);
    }
    if constexpr (FMT == F_U32 || FMT == F_S32) {
        uint32_t v = *reinterpret_cast<const uint32_t*>(p);
        if (SWAP) v = bswap32(v);
        const double d = FMT == F_U32 ? ((double)v - 2147483647.5) : (double)(int32_t)v;
        return (float)(d / 2147483647.5);
    }
    if constexpr (FMT == F_F32) {
        uint32_t v = *reinterpret_cast<const uint32_t*>(p);
        if (SWAP) v = bswap32(v);
        return __uint_as_float(v);
    }
    if constexpr (FMT == F_F64) {
        uint64_t v = *reinterpret_cast<const uint64_t*>(p);
        if (SWAP) v = bswap64(v);
        return (float)__longlong_as_double((long long)v);
    }
    return 0.f;
}

template <int FMT, bool SWAP, int BYTES>
__global__ void __launch_bounds__(256)
iqconv_kernel(const unsigned char* __restrict__ x, float2* __restrict__ y, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned char* p = x + i * (2 * BYTES);
        y[i] = make_float2(conv_one<FMT, SWAP>(p), conv_one<FMT, SWAP>(p + BYTES));
    }
}

// u8 IQ (the RTL-SDR wire format): 8 samples = 16 bytes per 128-bit load, two 128-bit stores per 4 samples
__global__ void __launch_bounds__(256)
iqconv_u8_vec_kernel(const uint4* __restrict__ x, float4* __restrict__ y, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldcs(x + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = __fdiv_rn((float)(w[k] & 0xff) - 127.5f, 127.5f);
            o.y = __fdiv_rn((float)((w[k] >> 8) & 0xff) - 127.5f, 127.5f);
            o.z = __fdiv_rn((float)((w[k] >> 16) & 0xff) - 127.5f, 127.5f);
            o.w = __fdiv_rn((float)(w[k] >> 24) - 127.5f, 127.5f);
            __stcs(y + 4 * i + k, o);
        }
    }
}

struct FmtInfo { const char* name; int fmt; int bytes; bool big_endian; };
const FmtInfo FORMATS[] = {
    {"u8", F_U8, 1, false}, {"s8", F_S8, 1, false},
    {"u16le", F_U16, 2, false}, {"u16be", F_U16, 2, true}, {"s16le", F_S16, 2, false}, {"s16be", F_S16, 2, true},
    {"u32le", F_U32, 4, false}, {"u32be", F_U32, 4, true}, {"s32le", F_S32, 4, false}, {"s32be", F_S32, 4, true},
    {"f32le", F_F32, 4, false}, {"f32be", F_F32, 4, true}, {"f64le", F_F64, 8, false}, {"f64be", F_F64, 8, true},
};

struct IqConvBlock : Block {
    FmtInfo info;
    std::string label;
    explicit IqConvBlock(const FmtInfo& f, bool dev) : info(f) {
        label = std::string("iqconv(") + f.name + ")";
        name = label.c_str();
        in_size = (size_t)2 * f.bytes;
        out_size = 8;
        dev_ptrs = dev;
    }
    int run(const void* dx, size_t n, void* dy, size_t* n_out, cudaStream_t s) override {
        *n_out = n;
        consumed += n;
        if (n == 0) return 0;
        const unsigned char* x = (const unsigned char*)dx;
        float2* y = (float2*)dy;
        const long long nn = (long long)n;
        int blocks = (int)((nn + 255) / 256);
        const int cap = ctx().sm_count * 16;
        if (blocks > cap) blocks = cap;
        const bool sw = info.big_endian;          // device is little endian
#define LRB_CONV(F, B) \
        if (sw) iqconv_kernel<F, true, B><<<blocks, 256, 0, s>>>(x, y, nn); else iqconv_kernel<F, false, B><<<blocks, 256, 0, s>>>(x, y, nn)
        switch (info.fmt) {
            case F_U8:
                if ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && nn >= 8) {
                    const long long n16 = nn / 8;
                    int vb = (int)((n16 + 255) / 256);
                    if (vb > cap) vb = cap;
                    iqconv_u8_vec_kernel<<<vb, 256, 0, s>>>((const uint4*)x, (float4*)y, n16);
                    if (nn % 8) { iqconv_kernel<F_U8, false, 1><<<1, 32, 0, s>>>(x + n16 * 16, y + n16 * 8, nn % 8); count_launch(); }
                } else {
                    iqconv_kernel<F_U8, false, 1><<<blocks, 256, 0, s>>>(x, y, nn);
                }
                break;
            case F_S8: iqconv_kernel<F_S8, false, 1><<<blocks, 256, 0, s>>>(x, y, nn); break;
            case F_U16: LRB_CONV(F_U16, 2); break;
            case F_S16: LRB_CONV(F_S16, 2); break;
            case F_U32: LRB_CONV(F_U32, 4); break;
            case F_S32: LRB_CONV(F_S32, 4); break;
            case F_F32: LRB_CONV(F_F32, 4); break;
            default: LRB_CONV(F_F64, 8); break;
        }
#undef LRB_CONV
        count_launch();
        LRB_CHECK(cudaGetLastError());
        return 0;
    }
};

}  // namespace

Block* make_iqconv(const char* format, bool dev) {
    if (!format) { set_error("iqconv: format is NULL"); return nullptr; }
    for (const FmtInfo& f : FORMATS)
        if (std::strcmp(f.name, format) == 0) {
            Block* b = new (std::nothrow) IqConvBlock(f, dev);
            if (!b) set_error("out of memory");
            return b;
        }
    set_error("Unsupported format (\"%s\")", format);     // iqfile.lua:46
    return nullptr;
}

}  // namespace lrb
