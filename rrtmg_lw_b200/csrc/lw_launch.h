// lw_launch.h -- host-callable launchers of the kernels (one set per translation unit).
#pragma once
#include <cuda_runtime.h>

#include "lw_device.cuh"

#ifndef LW_MAXLAY
#define LW_MAXLAY 640  // >= mxlay = 603 of modules/parrrtm.f90:31
#endif
#define LW_NUM_BAND_TUS 6

namespace lw {
struct DevPtrs;

void launch_prep(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);
void launch_mcica_kiss(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);
void kiss_work_stats(unsigned long long* out8);
bool launch_mcica_kiss_work(const ColIn&, const Scratch&, const RunCfg&, int resident, cudaStream_t);
// Mersenne-Twister generator (irng = 1): masks of a chunk whose first column is global column mt_c0 of a call over
// mt_ncol columns; returns the kernels launched or -1 (allocation failure)
int launch_mcica_mt(const ColIn&, const Scratch&, const RunCfg&, long mt_ncol, long mt_c0, cudaStream_t);
void free_mt_buffers();
#define MTJ_NPOW 44  // jump distances below 2^44 words (140 sub-columns x 2^24 columns x 2 x 640 layers < 2^43)
const unsigned* mt_jump_pow2_table();  // lw_mtjump.cpp: x^(2^b) mod phi(MT19937), MTJ_NPOW x 624 words; NULL on failure
int mt_jump_selftest();                // host-only check of the jump algebra against plain stepping
// FP32 mode: float copies of the caller arrays the band kernels read, into Scratch::inf; fills *out with pointers
// biased by the chunk's first column (the kernels index caller arrays with c0 + i); returns kernels launched
int launch_cvt_inputs(const ColIn&, const Scratch&, const RunCfg&, int cloudy, int inflag, ColInF* out, cudaStream_t);
void launch_cldflags(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);
void launch_finalize(const ColIn&, const ColOut&, const Scratch&, const RunCfg&, int cloudy, cudaStream_t);
void launch_finalize_slots(const ColIn&, const ColOut&, const Scratch&, const RunCfg&, int cloudy, unsigned slots, cudaStream_t);
void launch_get_alpha(int ncol, long ld, int nlay, int icld, int idcor, double decorr_con, const double* dz,
                      const double* lat, int juldat, const double* cldfrac, double* alpha, double am1, double am2,
                      double am4, cudaStream_t);
void launch_materialize(const ColIn&, const Scratch&, const RunCfg&, double* cldfmcl, double* ciwpmcl, double* clwpmcl,
                        double* reicmcl, double* relqmcl, double* taucmcl, cudaStream_t);
int measure_fp64_peak(double* tflops, int reps);
int kiss_jump_selftest();
cudaError_t upload_constants_kernels(const ConstTab&, const DevPtrs&, const double (*refrat)[5]);

// band kernels: each TU handles a subset of bands; returns false when the band is not its own
#define LW_DECL_TU(n)                                                                                         \
    bool launch_band_tu##n(int band, int cld, int idrv, const ColIn&, const ColOut&, const Scratch&, const RunCfg&, \
                           cudaStream_t);                                                                     \
    bool launch_band32_tu##n(int band, int cld, int idrv, const ColInF&, const ColOut&, const Scratch&, const RunCfg&, \
                             cudaStream_t);                                                                   \
    cudaError_t upload_constants_tu##n(const ConstTab&, const DevPtrs&, const double (*refrat)[5], const ConstTabF*);
LW_DECL_TU(0) LW_DECL_TU(1) LW_DECL_TU(2) LW_DECL_TU(3) LW_DECL_TU(4) LW_DECL_TU(5)
#undef LW_DECL_TU

// two-phase path (lw_taumol.cuh, lw_rtrnmc.cuh; instantiated in lw_twophase_inst.cu, one TU per -DLW_STU=n)
#ifndef LW_TAUMOL_MINB
#define LW_TAUMOL_MINB 4  // resident CTAs per SM the (column, layer) taumol kernel is compiled for
#endif
#define LW_NUM_SWEEP_TUS 6
void set_kiss_sparse(int on);  // KISS generator: step over the draws of cloud-free layers (default on)
void launch_colprep(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);
void launch_cldarrays(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);  // strict entry, two-phase path; needs secdiff
void launch_cldoptics(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);  // needs the masks and secdiff
int launch_taumol(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);  // returns the number of launches
#define LW_DECL_STU(n)                                                                                             \
    bool launch_sweep_stu##n(int b0, int nb, int cld, int idrv, const ColIn&, const ColOut&, const Scratch&, const RunCfg&, \
                             const SweepTab&, cudaStream_t);                                                                        \
    cudaError_t upload_constants_stu##n(const ConstTab&, const DevPtrs&, const double (*refrat)[5], const ConstTabF*);
LW_DECL_STU(1) LW_DECL_STU(2) LW_DECL_STU(3) LW_DECL_STU(4) LW_DECL_STU(5)
#undef LW_DECL_STU
cudaError_t upload_constants_stu0(const ConstTab&, const DevPtrs&, const double (*refrat)[5], const ConstTabF*);
// non-McICA path (lw_gridcloud.cuh, STU 6): cldprop + overlap factors, then one rtrn / rtrnmr sweep kernel per band
cudaError_t upload_constants_stu6(const ConstTab&, const DevPtrs&, const double (*refrat)[5], const ConstTabF*);
void launch_gridcloud_prep(const ColIn&, const Scratch&, const RunCfg&, int maxran, cudaStream_t);
bool launch_gridcloud_sweep(int band, int maxran, int idrv, const ColIn&, const Scratch&, const RunCfg&, const SweepTab&,
                            cudaStream_t);
}  // namespace lw
