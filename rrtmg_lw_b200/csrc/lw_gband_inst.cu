// lw_gband_inst.cu -- instantiations of the g-lane band kernel (lw_gband.cuh); compiled once per
// -DLW_GTU=n (two bands each) so the 16 bands x 3 cloud modes build in parallel.
#include "lw_gband.cuh"
#include "lw_launch.h"

#ifndef LW_GTU
#error "compile with -DLW_GTU=0..7"
#endif

#if LW_GTU == 0
#define LW_TU_BANDS(X) X(3) X(14)
#elif LW_GTU == 1
#define LW_TU_BANDS(X) X(4) X(15)
#elif LW_GTU == 2
#define LW_TU_BANDS(X) X(5) X(16)
#elif LW_GTU == 3
#define LW_TU_BANDS(X) X(7) X(10)
#elif LW_GTU == 4
#define LW_TU_BANDS(X) X(9) X(6)
#elif LW_GTU == 5
#define LW_TU_BANDS(X) X(1) X(8)
#elif LW_GTU == 6
#define LW_TU_BANDS(X) X(2) X(11)
#else
#define LW_TU_BANDS(X) X(12) X(13)
#endif

#define LW_CAT2(a, b) a##b
#define LW_CAT(a, b) LW_CAT2(a, b)

namespace lw {

template <int B>
static void launch_g(int cld, int grid, const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
#define LW_GO(CLD)                                                                                                  \
    do {                                                                                                            \
        static bool attr_set = false;                                                                               \
        if (!attr_set) {                                                                                            \
            cudaFuncSetAttribute(gband_kernel<B, CLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GB_SMEM_BYTES); \
            attr_set = true;                                                                                        \
        }                                                                                                           \
        gband_kernel<B, CLD><<<grid, GB_THREADS, GB_SMEM_BYTES, st>>>(in, s, r);                                   \
    } while (0)
    if (cld == CLD_NONE) LW_GO(CLD_NONE);
    else if (cld == CLD_MASK) LW_GO(CLD_MASK);
    else LW_GO(CLD_ARRAYS);
#undef LW_GO
}

bool LW_CAT(launch_gband_tu, LW_GTU)(int band, int cld, int grid, const ColIn& in, const Scratch& s, const RunCfg& r,
                                     cudaStream_t st) {
    switch (band) {
#define LW_CASE(b) case b: launch_g<b>(cld, grid, in, s, r, st); return true;
        LW_TU_BANDS(LW_CASE)
#undef LW_CASE
        default: return false;
    }
}

cudaError_t LW_CAT(upload_constants_gtu, LW_GTU)(const ConstTab& c, const DevPtrs& d, const double (*refrat)[5]) {
    return upload_constants_here(c, d, refrat);
}

}  // namespace lw
