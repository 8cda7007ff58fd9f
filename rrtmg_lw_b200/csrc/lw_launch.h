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
// Mersenne-Twister generator (irng = 1): masks of a chunk whose first column is global column mt_c0 of a call over
// mt_ncol columns; returns the kernels launched or -1 (allocation failure)
int launch_mcica_mt(const ColIn&, const Scratch&, const RunCfg&, long mt_ncol, long mt_c0, cudaStream_t);
void free_mt_buffers();
// FP32 mode: float copies of the caller arrays the band kernels read, into Scratch::inf; fills *out with pointers
// biased by the chunk's first column (the kernels index caller arrays with c0 + i); returns kernels launched
int launch_cvt_inputs(const ColIn&, const Scratch&, const RunCfg&, int cloudy, int inflag, ColInF* out, cudaStream_t);
void launch_cldflags(const ColIn&, const Scratch&, const RunCfg&, cudaStream_t);
void launch_finalize(const ColIn&, const ColOut&, const Scratch&, const RunCfg&, int cloudy, cudaStream_t);
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

// g-lane band kernels (lw_gband.cuh): grid = persistent CTAs
#define LW_NUM_GBAND_TUS 8
#define LW_DECL_GTU(n)                                                                                      \
    bool launch_gband_tu##n(int band, int cld, int grid, const ColIn&, const Scratch&, const RunCfg&, cudaStream_t); \
    cudaError_t upload_constants_gtu##n(const ConstTab&, const DevPtrs&, const double (*refrat)[5]);
LW_DECL_GTU(0) LW_DECL_GTU(1) LW_DECL_GTU(2) LW_DECL_GTU(3) LW_DECL_GTU(4) LW_DECL_GTU(5) LW_DECL_GTU(6) LW_DECL_GTU(7)
#undef LW_DECL_GTU
void launch_transposes(const Scratch&, const RunCfg&, int with_mask, cudaStream_t);
void launch_finalizeT(const ColIn&, const ColOut&, const Scratch&, const RunCfg&, int cloudy, cudaStream_t);
}  // namespace lw
