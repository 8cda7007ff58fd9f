// lw_twophase_inst.cu -- instantiations of the two-phase kernels (lw_taumol.cuh, lw_rtrnmc.cuh); compiled once per
// -DLW_STU=n so that the band groups x cloud modes x idrv variants build in parallel.
//   STU 0: column scalars + the (column, layer) taumol kernel
//   STU 1-3: the default band groups of the sweep kernel
//   STU 4-5: single-band sweep kernels (band-range runs, e.g. the per-band output blocks of the column-mode harness)
//   STU 6: the non-McICA path (lw_gridcloud.cuh): cldprop + overlap factors, rtrn / rtrnmr sweeps
#include <cstdlib>

#include "lw_rtrnmc.cuh"
#if LW_STU == 6
#include "lw_gridcloud.cuh"
#endif
#include "lw_launch.h"

#ifndef LW_STU
#error "compile with -DLW_STU=0..6"
#endif

#define LW_CAT2(a, b) a##b
#define LW_CAT(a, b) LW_CAT2(a, b)

namespace lw {

cudaError_t LW_CAT(upload_constants_stu, LW_STU)(const ConstTab& c, const DevPtrs& d, const double (*refrat)[5],
                                                  const ConstTabF* cf) {
    return upload_constants_here(c, d, refrat, cf);
}

#if LW_STU == 0
// The 16 bands are evaluated by a few launches of the (column, layer) kernel, each instantiated for a band group
// (mask: bit b-1 = band b) and compiled for its own number of resident CTAs; every launch redoes inatm / setcoef / the
// sort for its bands (about 2 ms per 1 M columns).  LW_TAUMOL_GROUPS lists them as X(mask, min_ctas_per_sm).
// Default (variant 0): the five bands with two key species in BOTH atmosphere halves (3 4 5 7 9) at 128 registers /
// 4 CTAs, the other eleven at 80 registers / 6 CTAs -- the kernel is latency-bound, its rate follows the resident warps,
// and those eleven spill < 60 bytes each at 80 registers while the five lose at anything below 128.  81.1 -> 70.7 ms per
// 1 M columns against the one-launch form (variant 9); variants 1-8 are the other groupings that were measured
// (profiles/r02_taumol_band_groups_experiment.log).
#ifndef LW_TAUMOL_VARIANT
#define LW_TAUMOL_VARIANT 0
#endif
#if LW_TAUMOL_VARIANT == 0
#define LW_TAUMOL_GROUPS X(0x015cu, 4) X(0xfea3u, 6)
#elif LW_TAUMOL_VARIANT == 1
#define LW_TAUMOL_GROUPS X(0x001cu, 4) X(0x0140u, 4) X(0xfea3u, 6)
#elif LW_TAUMOL_VARIANT == 2
#define LW_TAUMOL_GROUPS X(0x015cu, 4) X(0x00a3u, 6) X(0xfe00u, 6)
#elif LW_TAUMOL_VARIANT == 3
#define LW_TAUMOL_GROUPS X(0x001cu, 4) X(0x0140u, 4) X(0x00a3u, 6) X(0xfe00u, 6)
#elif LW_TAUMOL_VARIANT == 4
#define LW_TAUMOL_GROUPS X(0x4u, 4) X(0x8u, 4) X(0x10u, 4) X(0x40u, 4) X(0x100u, 4) X(0xfea3u, 6)
#elif LW_TAUMOL_VARIANT == 5
#define LW_TAUMOL_GROUPS X(0x015cu, 4) X(0xfea3u, 7)
#elif LW_TAUMOL_VARIANT == 6
#define LW_TAUMOL_GROUPS X(0x4u, 4) X(0x8u, 4) X(0x10u, 4) X(0x40u, 4) X(0x100u, 4) X(0x00a3u, 6) X(0xfe00u, 6)
#elif LW_TAUMOL_VARIANT == 9
#define LW_TAUMOL_GROUPS X(0xffffu, LW_TAUMOL_MINB)
#elif LW_TAUMOL_VARIANT == 7
#define LW_TAUMOL_GROUPS X(0x4u, 5) X(0x8u, 5) X(0x10u, 5) X(0x40u, 5) X(0x100u, 5) X(0xfea3u, 6)
#elif LW_TAUMOL_VARIANT == 8
#define LW_TAUMOL_GROUPS X(0x015cu, 4) X(0x00a3u, 6) X(0x0e00u, 6) X(0xf000u, 6)
#endif
template <unsigned BMASK, int MINB>
static void launch_taumol_part(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    const dim3 grid((unsigned)((r.ncol + TAUMOL_THREADS - 1) / TAUMOL_THREADS), (unsigned)r.nlay);
    // RRTMG_LW_TAUMOL_SMEM_PAD (bytes, dev knob): extra dynamic shared memory per CTA, to lower the residency
    static const size_t pad = std::getenv("RRTMG_LW_TAUMOL_SMEM_PAD") ? (size_t)std::atol(std::getenv("RRTMG_LW_TAUMOL_SMEM_PAD")) : 0;
    const size_t smem = (size_t)NFIELD * TAUMOL_THREADS * sizeof(double) + pad;
    auto kern = taumol_kernel<BMASK, MINB>;
    if (smem > 40 * 1024) {  // static (sort arrays) + dynamic shared memory beyond 48 KB needs the opt-in
        static unsigned attr_set = 0u;  // bit d = done on device d
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((attr_set >> dev) & 1u)) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set |= 1u << dev;
        }
    }
    kern<<<grid, TAUMOL_THREADS, smem, st>>>(in, s, r);
}
int launch_taumol(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    int n = 0;
    constexpr unsigned all = 0u
#define X(mask, minb) | (mask)
        LW_TAUMOL_GROUPS
#undef X
        ;
    constexpr unsigned sum = 0u
#define X(mask, minb) + (mask)
        LW_TAUMOL_GROUPS
#undef X
        ;
    static_assert(all == 0xffffu && sum == 0xffffu, "LW_TAUMOL_GROUPS must hold every band exactly once");
#define X(mask, minb) launch_taumol_part<(mask), (minb)>(in, s, r, st); ++n;
    LW_TAUMOL_GROUPS
#undef X
    return n;
}
#elif LW_STU == 6
void launch_gridcloud_prep(const ColIn& in, const Scratch& s, const RunCfg& r, int maxran, cudaStream_t st) {
    gridcloud_prep_kernel<<<(r.ncol + 127) / 128, 128, 0, st>>>(in, s, r, maxran);
}
template <int B>
static void launch_gc(int maxran, int idrv, const ColIn& in, const Scratch& s, const RunCfg& r, const SweepTab& tb,
                      cudaStream_t st) {
    const int blocks = (r.ncol + 127) / 128;
    if (maxran) {
        if (idrv) gridcloud_sweep_kernel<B, 1, 1><<<blocks, 128, 0, st>>>(in, s, r, tb);
        else gridcloud_sweep_kernel<B, 1, 0><<<blocks, 128, 0, st>>>(in, s, r, tb);
    } else {
        if (idrv) gridcloud_sweep_kernel<B, 0, 1><<<blocks, 128, 0, st>>>(in, s, r, tb);
        else gridcloud_sweep_kernel<B, 0, 0><<<blocks, 128, 0, st>>>(in, s, r, tb);
    }
}
bool launch_gridcloud_sweep(int band, int maxran, int idrv, const ColIn& in, const Scratch& s, const RunCfg& r,
                            const SweepTab& tb, cudaStream_t st) {
    switch (band) {
#define LW_CASE(B) case B: launch_gc<B>(maxran, idrv, in, s, r, tb, st); return true;
        LW_CASE(1) LW_CASE(2) LW_CASE(3) LW_CASE(4) LW_CASE(5) LW_CASE(6) LW_CASE(7) LW_CASE(8) LW_CASE(9) LW_CASE(10)
        LW_CASE(11) LW_CASE(12) LW_CASE(13) LW_CASE(14) LW_CASE(15) LW_CASE(16)
#undef LW_CASE
    }
    return false;
}
#else

// registers a sweep kernel wants: the group's recurrences (total + clear stream, + their dT derivatives) and ~70 more
#ifndef LW_SWEEP_BASE_REGS
#define LW_SWEEP_BASE_REGS 72
#endif
#ifndef LW_SWEEP_MINB_6G
#define LW_SWEEP_MINB_6G 6  // band 10: 80 registers, 8 bytes of spill; 3.38 -> 3.13 ms per 1 M columns
#endif
#ifndef LW_SWEEP_MINB_8G
#define LW_SWEEP_MINB_8G 5  // bands 6 8 11 12: 96 registers, no spill; -2 % (80 registers: spills, slower)
#endif
#ifndef LW_SWEEP_MINB_10G
#define LW_SWEEP_MINB_10G 4  // band 1 and the 13-16 group spill at 96 registers: +17 %
#endif
#ifndef LW_SWEEP_MINB_12G
#define LW_SWEEP_MINB_12G 4  // bands 2 7 9 spill at 96 registers: +15 %
#endif
constexpr int sweep_minb(int ngg, int cld, int idrv) {
    // part of the clear-sky stream of a wide cloudy group lives in shared memory (SweepRing::clr_smem): 2 registers less each
    const int clr_sm = (cld != CLD_NONE && ngg >= LW_SWEEP_CLR_SMEM_MIN) ? LW_SWEEP_CLR_SMEM_N : 0;
    const int est = ((cld != CLD_NONE ? 4 : 2) * ngg * (idrv ? 2 : 1) - 2 * clr_sm) + LW_SWEEP_BASE_REGS;
    int m = est <= 128 ? 4 : (est <= 168 ? 3 : 2);
    // narrow groups without the dT derivatives: more resident CTAs (the sweeps are latency-bound too)
    if (m == 4 && idrv == 0) {
        if (ngg <= 6) m = LW_SWEEP_MINB_6G;
        else if (ngg <= 8) m = LW_SWEEP_MINB_8G;
        else if (ngg <= 10) m = LW_SWEEP_MINB_10G;
        else if (ngg <= 12) m = LW_SWEEP_MINB_12G;
    }
    return m;
}

template <int B0, int NB>
static void launch_g(int cld, int idrv, const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                     const SweepTab& tb, cudaStream_t st) {
    const int threads = SWEEP_THREADS;
    const int blocks = (r.ncol + threads - 1) / threads;
    constexpr int NGG = Grp<B0, NB>::NGG;
#define LW_GO(CLD, IDRV)                                                                                         \
    do {                                                                                                         \
        const size_t smem = SweepRing<B0, NB>::bytes(CLD);                                                       \
        auto kern = rtrnmc_kernel<B0, NB, CLD, IDRV, sweep_minb(NGG, CLD, IDRV)>;                                \
        static unsigned attr_set = 0u; /* bit d = done on device d (the attribute is per device) */             \
        int dev_ = 0;                                                                                            \
        cudaGetDevice(&dev_);                                                                                    \
        if (!((attr_set >> dev_) & 1u)) {                                                                        \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
            attr_set |= 1u << dev_;                                                                              \
        }                                                                                                        \
        kern<<<blocks, threads, smem, st>>>(in, out, s, r, tb);                                                    \
    } while (0)
    if (cld == CLD_NONE) { if (idrv) LW_GO(CLD_NONE, 1); else LW_GO(CLD_NONE, 0); }
    else if (cld == CLD_MASK) { if (idrv) LW_GO(CLD_MASK, 1); else LW_GO(CLD_MASK, 0); }
    else { if (idrv) LW_GO(CLD_ARRAYS, 1); else LW_GO(CLD_ARRAYS, 0); }
#undef LW_GO
}

#if LW_STU == 1
#define LW_STU_GROUPS(X) X(1, 2) X(6, 2)
#elif LW_STU == 2
#define LW_STU_GROUPS(X) X(3, 1) X(4, 1) X(5, 1)
#elif LW_STU == 3
#define LW_STU_GROUPS(X) X(8, 2) X(10, 3) X(13, 4)
#elif LW_STU == 4
#define LW_STU_GROUPS(X) X(1, 1) X(2, 1) X(6, 1) X(7, 1) X(8, 1)
#else
#define LW_STU_GROUPS(X) X(9, 1) X(10, 1) X(11, 1) X(12, 1) X(13, 1) X(14, 1) X(15, 1) X(16, 1)
#endif

bool LW_CAT(launch_sweep_stu, LW_STU)(int b0, int nb, int cld, int idrv, const ColIn& in, const ColOut& out,
                                      const Scratch& s, const RunCfg& r, const SweepTab& tb, cudaStream_t st) {
#define LW_CASE(B0, NB) if (b0 == B0 && nb == NB) { launch_g<B0, NB>(cld, idrv, in, out, s, r, tb, st); return true; }
    LW_STU_GROUPS(LW_CASE)
#undef LW_CASE
    return false;
}
#endif

}  // namespace lw
