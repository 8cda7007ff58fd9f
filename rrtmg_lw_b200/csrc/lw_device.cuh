// lw_device.cuh -- device-side data structures of the B200 RRTMG_LW path.
#pragma once
#include <cuda_runtime.h>

#include "lw_tables.h"

namespace lw {

// Caller arrays (API layout, see include/rrtmg_lw_b200.h).  Column i of the current chunk is
// caller column c0 + i.
struct ColIn {
    const double *play, *plev, *tlay, *tlev, *tsfc;
    const double *h2o, *o3, *co2, *ch4, *n2o, *o2, *cfc11, *cfc12, *cfc22, *ccl4;
    const double *emis, *tauaer;
    // strict rrtmg_lw() McICA arrays (140, ld, nlay) / (ld, nlay)
    const double *cldfmcl, *taucmcl, *ciwpmcl, *clwpmcl, *reicmcl, *relqmcl;
    // grid-mean cloud inputs of the fused entry (ld, nlay); tauc (16, ld, nlay)
    const double *cldfrac, *ciwp, *clwp, *rei, *rel, *tauc, *alpha;
    long ld;
    int nlay;
};

// FP32 mode: float copies (made by the conversion pass of the prep stage, pitch-strided like the scratch) of the
// caller arrays the band kernels read; same member names as ColIn.
struct ColInF {
    const float *play, *tlay, *tlev, *tsfc, *emis, *tauaer;
    const float *cldfmcl, *taucmcl, *ciwpmcl, *clwpmcl, *reicmcl, *relqmcl;  // unused (strict entry stays FP64)
    const float *ciwp, *clwp, *rei, *rel, *tauc;
    long ld;
    int nlay;
};

struct ColOut {
    double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc, *duflx_dt, *duflxc_dt;
    long ld;
};

// per-(layer, column) fields written by the prep kernel (inatm + setcoef)
enum Field {
    F_FAC00, F_FAC01, F_FAC10, F_FAC11,
    F_COLH2O, F_COLCO2, F_COLO3, F_COLN2O, F_COLCH4, F_COLO2, F_COLBRD, F_COLDRY,
    F_SELFFAC, F_SELFFRAC, F_FORFAC, F_FORFRAC, F_MINORFRAC, F_SCALEMINOR, F_SCALEMINORN2,
    F_WX1, F_WX2, F_WX3, F_WX4,
    NFIELD
};

// cloud treatment of the band kernel
enum CloudMode { CLD_NONE = 0, CLD_MASK = 1, CLD_ARRAYS = 2, CLD_GRID = 3 };  // CLD_GRID: lw_gridcloud.cuh kernels

// per-(column, layer) record written by gridcloud_prep_kernel: [layer][GC_NREC][pitch] doubles
enum GridCloudRec {
    GC_CLDFRAC = 0,
    GC_FLAGS,      // bit 0 icldlyr, bit 1 istcld(lev), bit 2 istcldd(lev)  (as a double holding a small integer)
    GC_FACCLR1, GC_FACCLR2, GC_FACCLD1, GC_FACCLD2, GC_FACCMB1, GC_FACCMB2,        // index lev+1 of the reference (upward)
    GC_FACCLR1D, GC_FACCLR2D, GC_FACCLD1D, GC_FACCLD2D, GC_FACCMB1D, GC_FACCMB2D,  // index lev-1 of the reference (downward)
    GC_ODCLD,                  // 16 bands: secdiff(ib) * taucloud(lay, ib), ib = ipat(band, ncbands)
    GC_EFCL = GC_ODCLD + 16,   // 16 bands: (1 - exp(-odcld)) * cldfrac  (rtrn only)
    GC_NREC = GC_EFCL + 16
};

struct Scratch {
    double* fld;        // NFIELD x nlay x pitch
    int* ipack;         // nlay x pitch : jp | jt<<6 | jt1<<9 | indself<<12 | indfor<<16 | indminor<<18
    double* pwvcm;      // pitch
    int* laytrop;       // pitch
    unsigned* mask;     // nlay x 5 x pitch : 140 sub-column cloud bits; bit 31 of word 4 = icldlyr
    int* ctop;          // pitch : highest layer (1-based) with icldlyr = 1, 0 = cloud-free column
    double* state;      // 5 x 16 x nlay x pitch : atrans, bbugas, atot, bbutot, efclfrac per (g, layer)
    double* abscld;     // nlay x pitch : 1 - exp(-odcld) of the current band (CLD_MASK)
    double* fluxB;      // [pitch / 128][nlay + 1][6][128] running band sums of dflx, dflxc, uflx, uflxc, duflx_dt,
                        // duflxc_dt (column-lane kernel; the finalize kernel writes the caller's arrays from them)
    // FP32 mode (set_option("precision", 32)): float views / copies used by the generated float band kernels
    float* fldf;        // NFIELD x nlay x pitch (written by the prep kernel instead of fld)
    float* statef;      // aliases state
    float* abscldf;     // aliases abscld
    float* inf;         // float copies of caller arrays, see lw_api.cu: carve_fp32_inputs
    unsigned long long* status;  // sticky error word: code<<56 | (col+1)<<24 | lay
    long err_base;      // added to the column an input error is reported with (first column of a staged host pass)
    long pitch;         // columns per row of every scratch array (multiple of 32)
    // ---- two-phase path (lw_taumol.cuh -> lw_rtrnmc.cuh)
    double* od;         // [pitch / 128][nlay][70][128][2] : odepth = max(0, secdiff * (taug + tauaer)) of g-point pairs
    double* pfs;        // [pitch / 128][nlay][9][128] : eta (specmult) of the Planck-fraction interpolation, binary bands
    double* secdiff;    // 16 x pitch : diffusivity angle per band (rtrnmc.f90:258-281)
    double* cldb;       // [pitch / 128][nlay][16][2][128] : {odcld, 1 - exp(-odcld)} of a band's cloudy sub-columns (CLD_MASK)
    double* fluxP;      // [pitch / 128][nlay + 1][16][6][128] : level-flux band sums, one slot per sweep launch (first band)
    double* gcld;       // 2 x 140 x nlay x pitch : efclfrac, odcld per g-point (strict rrtmg_lw() entry only)
    double* gridcld;    // nlay x GC_NREC x pitch : cloud fraction, overlap factors, band cloud optical depths of the
                        // non-McICA entry (lw_gridcloud.cuh)
    // debug capture of one column
    int dbg_col;        // chunk-relative column or -1
    double *dbg_taug, *dbg_fracs;  // (140, nlay)
    int* dbg_itr;       // (140, nlay, 2)
    int* dbg_ints;      // (nlay, 6)
    double* dbg_facs;   // (nlay, 4)
    double* dbg_scal;   // 2
};

// Tables and constants of the sweep kernels (lw_rtrnmc.cuh), passed BY VALUE as a kernel parameter: parameters sit in
// constant bank 0 and are read as immediate operands with compile-time offsets, whereas the big __constant__ table
// block needs an LDC per use whose misses showed up as the top stall of the vertical loops.
struct SweepTab {
    const double2* exptfn;        // [itr] = {exp_tbl, tfn_tbl}
    const double* tau_tbl;
    const double* totplnk;        // (181,17): column 16 (0-based) = totplk16
    const double* totplnkderiv;   // (181,17)
    const double* fracrefa[NBND]; // padded rows (lw_tables.h NGPAD)
    const double* fracrefb[NBND];
    double delwave[NBND];
    double bpade, fluxfac;
};

struct RunCfg {
    int ncol;        // columns in this chunk
    long c0;         // first caller column of the chunk
    int nlay;
    int idrv;
    int inflag, iceflag, liqflag;
    int icld, irng, permuteseed;
    int fp32;        // 1 = the prep kernel writes float fields (Scratch::fldf) for the FP32-mode band kernels
    int istart;      // first band of the run (two-phase path; 16 = band 16 alone, which then takes its own Planck table)
    int fld_block;   // layout of Scratch::fld: 0 = [field][layer][column], 1 = blocks of 128 columns,
                     // [column / 128][layer][field][column % 128] (column-lane kernel: a CTA's fields of one layer are
                     // one contiguous run, addressed as one base + compile-time offsets); pitch is a multiple of 128
};

}  // namespace lw
