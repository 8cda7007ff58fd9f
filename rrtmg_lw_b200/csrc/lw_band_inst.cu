// lw_band_inst.cu -- instantiations of the band kernel; compiled once per -DLW_TU=n so the 16
// bands x cloud modes x idrv variants build in parallel.
#include <cstdlib>

#include "lw_band.cuh"
#include "_generated/lw_band32.cuh"
#include "lw_launch.h"

#ifndef LW_TU
#error "compile with -DLW_TU=0..5"
#endif

#if LW_TU == 0
#define LW_TU_BANDS(X) X(1) X(2) X(6) X(10)
#elif LW_TU == 1
#define LW_TU_BANDS(X) X(3) X(14)
#elif LW_TU == 2
#define LW_TU_BANDS(X) X(4) X(8)
#elif LW_TU == 3
#define LW_TU_BANDS(X) X(5) X(11)
#elif LW_TU == 4
#define LW_TU_BANDS(X) X(7) X(12) X(16)
#else
#define LW_TU_BANDS(X) X(9) X(13) X(15)
#endif

#define LW_CAT2(a, b) a##b
#define LW_CAT(a, b) LW_CAT2(a, b)

namespace lw {

template <int B>
static void launch_b(int cld, int idrv, const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                     cudaStream_t st) {
    const int threads = BAND_THREADS;
    const int blocks = (r.ncol + threads - 1) / threads;
    // RRTMG_LW_SMEM_PAD (bytes, dev experiments): extra dynamic shared memory per CTA, i.e. fewer resident CTAs
    static const size_t pad = std::getenv("RRTMG_LW_SMEM_PAD") ? (size_t)std::atol(std::getenv("RRTMG_LW_SMEM_PAD")) : 0;
    const size_t smem = Ring<B>::BYTES + pad;
#define LW_GO(CLD, IDRV)                                                                                  \
    do {                                                                                                  \
        static unsigned attr_set = 0u; /* bit d = done on device d (the attribute is per device) */      \
        int dev_ = 0;                                                                                     \
        cudaGetDevice(&dev_);                                                                             \
        if (!((attr_set >> dev_) & 1u)) {                                                                 \
            cudaFuncSetAttribute(band_kernel<B, CLD, IDRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            attr_set |= 1u << dev_;                                                                       \
        }                                                                                                 \
        band_kernel<B, CLD, IDRV><<<blocks, threads, smem, st>>>(in, out, s, r);                          \
    } while (0)
    if (cld == CLD_NONE) { if (idrv) LW_GO(CLD_NONE, 1); else LW_GO(CLD_NONE, 0); }
    else if (cld == CLD_MASK) { if (idrv) LW_GO(CLD_MASK, 1); else LW_GO(CLD_MASK, 0); }
    else { if (idrv) LW_GO(CLD_ARRAYS, 1); else LW_GO(CLD_ARRAYS, 0); }
#undef LW_GO
}

// FP32 mode (generated float kernels, lw::p32): clear sky and the fused McICA entry
template <int B>
static void launch_b32(int cld, int idrv, const ColInF& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                       cudaStream_t st) {
    const int threads = BAND_THREADS;
    const int blocks = (r.ncol + threads - 1) / threads;
    const size_t smem = p32::Ring<B>::BYTES;
#define LW_GO(CLD, IDRV)                                                                                  \
    do {                                                                                                  \
        static unsigned attr_set = 0u; /* bit d = done on device d (the attribute is per device) */      \
        int dev_ = 0;                                                                                     \
        cudaGetDevice(&dev_);                                                                             \
        if (!((attr_set >> dev_) & 1u)) {                                                                 \
            cudaFuncSetAttribute(p32::band_kernel<B, CLD, IDRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            attr_set |= 1u << dev_;                                                                       \
        }                                                                                                 \
        p32::band_kernel<B, CLD, IDRV><<<blocks, threads, smem, st>>>(in, out, s, r);                     \
    } while (0)
    if (cld == CLD_NONE) { if (idrv) LW_GO(CLD_NONE, 1); else LW_GO(CLD_NONE, 0); }
    else { if (idrv) LW_GO(CLD_MASK, 1); else LW_GO(CLD_MASK, 0); }
#undef LW_GO
}

bool LW_CAT(launch_band32_tu, LW_TU)(int band, int cld, int idrv, const ColInF& in, const ColOut& out, const Scratch& s,
                                     const RunCfg& r, cudaStream_t st) {
    switch (band) {
#define LW_CASE(b) case b: launch_b32<b>(cld, idrv, in, out, s, r, st); return true;
        LW_TU_BANDS(LW_CASE)
#undef LW_CASE
        default: return false;
    }
}

bool LW_CAT(launch_band_tu, LW_TU)(int band, int cld, int idrv, const ColIn& in, const ColOut& out, const Scratch& s,
                                   const RunCfg& r, cudaStream_t st) {
    switch (band) {
#define LW_CASE(b) case b: launch_b<b>(cld, idrv, in, out, s, r, st); return true;
        LW_TU_BANDS(LW_CASE)
#undef LW_CASE
        default: return false;
    }
}

cudaError_t LW_CAT(upload_constants_tu, LW_TU)(const ConstTab& c, const DevPtrs& d, const double (*refrat)[5],
                                               const ConstTabF* cf) {
    return upload_constants_here(c, d, refrat, cf);
}

}  // namespace lw
