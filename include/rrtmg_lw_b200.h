/* rrtmg_lw_b200.h -- C ABI of the B200-native RRTMG_LW longwave hot path.
 *
 * These are the entry points a thin ISO_C_BINDING shim binds behind the reference's
 * UNCHANGED Fortran GCM interface (fortran/rrtmg_lw_b200_shim.f90):
 *
 *   reference interface (file:line under the reference tree)          C entry point
 *   ---------------------------------------------------------------  ------------------------------
 *   rrtmg_lw_ini(cpdair)            src/rrtmg_lw_init.f90:47          rrtmg_lw_b200_ini
 *   rrtmg_lw(ncol,nlay,icld,...)    src/rrtmg_lw_rad.f90:99-108       rrtmg_lw_b200_rrtmg_lw
 *   mcica_subcol_lw(iplon,ncol,...) src/mcica_subcol_gen_lw.f90:183   rrtmg_lw_b200_mcica_subcol_lw
 *   get_alpha(iplon,ncol,...)       src/mcica_subcol_gen_lw.f90:68    rrtmg_lw_b200_get_alpha
 *   [mcica_subcol_lw -> rrtmg_lw fused, sub-columns never materialised] rrtmg_lw_b200_rrtmg_lw_mcica
 *   rrtmg_lw(ncol,nlay,icld,...)    src/rrtmg_lw_rad.nomcica.f90:99   rrtmg_lw_b200_rrtmg_lw_nomcica
 *
 * Conventions
 *   - Every array keeps the reference's Fortran layout, passed as a plain pointer:
 *       (ncol,nlay) arrays:     element (i,k) at  i + ld*k            (column index fastest)
 *       emis(ncol,16):          element (i,b) at  i + ld*b
 *       tauaer(ncol,nlay,16):   element (i,k,b) at i + ld*(k + nlay*b)
 *       McICA (140,ncol,nlay):  element (g,i,k) at g + 140*(i + ld*k)
 *       tauc(16,ncol,nlay):     element (b,i,k) at b + 16*(i + ld*k)
 *     `ld` is the extent of the column dimension of the caller's arrays (size(x,1) in the
 *     shim; >= ncol).  Layer 1 is the surface layer (src/rrtmg_lw_rad.f90:785-787).
 *   - memspace RRTMG_LW_HOST: pointers are host memory, the call copies in/out and is
 *     synchronous.  RRTMG_LW_DEVICE: pointers are device memory on the current CUDA device,
 *     work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = default stream);
 *     the status word is checked on the next HOST-mode call or rrtmg_lw_b200_sync().
 *   - Return value: 0 on success, otherwise an RRTMG_LW_ERR_* code.  The reference STOPs at
 *     these sites (src/rrtmg_lw_cldprmc.f90:191,204,208,217,231,253;
 *     src/mcica_subcol_gen_lw.f90:259,465); the shim turns a non-zero code into
 *     `error stop` with the reference's message (rrtmg_lw_b200_strerror).
 *   - No CPU fallback exists: without a CUDA device every compute entry returns
 *     RRTMG_LW_ERR_CUDA.
 */
#ifndef RRTMG_LW_B200_H
#define RRTMG_LW_B200_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RRTMG_LW_OK = 0,
    RRTMG_LW_ERR_FLAG = 1,          /* invalid icld / inflag=1 with McICA */
    RRTMG_LW_ERR_ICE_RADIUS = 2,    /* 'ICE RADIUS OUT OF BOUNDS' etc. */
    RRTMG_LW_ERR_LIQ_RADIUS = 3,    /* 'LIQUID EFFECTIVE RADIUS OUT OF BOUNDS' */
    RRTMG_LW_ERR_KISS_PRESSURE = 4, /* 'KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS' */
    RRTMG_LW_ERR_CUDA = 5,          /* CUDA runtime failure / no device */
    RRTMG_LW_ERR_NOT_INIT = 6,      /* rrtmg_lw_b200_ini not called */
    RRTMG_LW_ERR_ARG = 7            /* bad argument (NULL pointer, ld < ncol, nlay out of range, ...) */
};

enum { RRTMG_LW_HOST = 0, RRTMG_LW_DEVICE = 1 };

/* Replaces rrtmg_lw_ini (src/rrtmg_lw_init.f90:47-194): constants, g-point reduction maps,
 * tau/exp/tfn lookup tables, reduction of the 16-g k-data to 140 g-points, upload to the
 * current CUDA device.
 *   kdata           : unreduced k-distribution blob, include/rrtmg_lw_kdata_layout.h
 *                     (the reference compiles these data in via rrtmg_lw_k_g.f90)
 *   nkdata          : number of doubles in kdata (must equal RRTMG_LW_KDATA_DOUBLES)
 *   default_real_kind: 4 = semantics of a plain `gfortran -O3` build of the reference,
 *                     8 = of the shipped `-r8` makefiles (affects default-kind literals only) */
int rrtmg_lw_b200_ini(double cpdair, const double* kdata, long nkdata, int default_real_kind);

/* The two halves of rrtmg_lw_ini, for deployments that do not want the k-data or the start-up reduction on the
 * compute nodes.  rrtmg_lw_b200_tables_build is HOST-ONLY (needs no CUDA device): it runs the rrtmg_lw_ini restatement
 * (src/rrtmg_lw_init.f90:125-173, 215-298, 385-2034: constants, lookup tables, the 16-g -> 140-g reduction of every
 * k-table) and writes a versioned binary image of the result; returns the image size in bytes (call with out = NULL to
 * size the buffer), or -RRTMG_LW_ERR_ARG.  rrtmg_lw_b200_ini_tables uploads such an image (checksum, version and
 * shape checked: RRTMG_LW_ERR_ARG otherwise) and leaves the library in the same state as rrtmg_lw_b200_ini with the
 * k-data the image was built from -- bit-identical tables and results.  tools/make_table_blob.py is the command-line
 * form. */
long rrtmg_lw_b200_tables_build(double cpdair, const double* kdata, long nkdata, int default_real_kind,
                                void* out, long cap);
int rrtmg_lw_b200_ini_tables(const void* blob, long nbytes);

/* Array-by-array alternative to building the blob: hand over one unreduced k-array of
 * modules/rrlw_kgNN.f90 (e.g. band 3, "kao", 585*16 doubles, Fortran element order) as the
 * reference's lw_kgbNN left it; after all RRTMG_LW_NKARRAYS arrays were put,
 * rrtmg_lw_b200_ini(cpdair, NULL, 0, kind) consumes them.  n must equal 16*lead of the array. */
int rrtmg_lw_b200_kdata_put(int band, const char* name, const double* data, long n);

/* Replaces rrtmg_lw (src/rrtmg_lw_rad.f90:99-595), argument for argument.
 * icld is inout (clamped to 2 when outside [0,3], :470).  duflx_dt/duflxc_dt may be NULL
 * when idrv == 0.  The McICA arrays are read only when *icld >= 1. */
int rrtmg_lw_b200_rrtmg_lw(int ncol, int ld, int nlay, int* icld, int idrv,
                           const double* play, const double* plev, const double* tlay, const double* tlev,
                           const double* tsfc, const double* h2ovmr, const double* o3vmr, const double* co2vmr,
                           const double* ch4vmr, const double* n2ovmr, const double* o2vmr,
                           const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr,
                           const double* ccl4vmr, const double* emis, int inflglw, int iceflglw, int liqflglw,
                           const double* cldfmcl, const double* taucmcl, const double* ciwpmcl,
                           const double* clwpmcl, const double* reicmcl, const double* relqmcl,
                           const double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc,
                           double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                           int memspace, void* stream);

/* Replaces rrtmg_lw of the NON-McICA build (src/rrtmg_lw_rad.nomcica.f90:99-588), argument for argument:
 * grid-mean cloud arrays cldfr, cicewp, cliqwp, reice, reliq (ncol,nlay) and taucld (16,ncol,nlay);
 * cldprop (src/rrtmg_lw_cldprop.f90) then rtrn (random overlap, *icld == 1, src/rrtmg_lw_rtrn.f90) or
 * rtrnmr (maximum-random overlap, *icld == 0, 2, 3, src/rrtmg_lw_rtrnmr.f90).  icld is inout (clamped to 2 when
 * outside [0,3], :458).  The cloud arrays are read only when *icld >= 1.  Radius errors of cldprop
 * (:212,217,228,244,272) come back as RRTMG_LW_ERR_ICE_RADIUS / RRTMG_LW_ERR_LIQ_RADIUS. */
int rrtmg_lw_b200_rrtmg_lw_nomcica(int ncol, int ld, int nlay, int* icld, int idrv,
                                   const double* play, const double* plev, const double* tlay, const double* tlev,
                                   const double* tsfc, const double* h2ovmr, const double* o3vmr,
                                   const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                                   const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                                   const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                                   int iceflglw, int liqflglw, const double* cldfr, const double* taucld,
                                   const double* cicewp, const double* cliqwp, const double* reice,
                                   const double* reliq, const double* tauaer, double* uflx, double* dflx, double* hr,
                                   double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                                   int memspace, void* stream);

/* Replaces mcica_subcol_lw (src/mcica_subcol_gen_lw.f90:183-291); irng is inout (:442).
 * iplon is accepted and ignored, as in the reference.  icld in [0,5]; alpha read for icld 4/5.
 * irng = 0: KISS, one stream per column seeded from the column's four lowest pressures (needs nlay >= 4).
 * irng = 1: Mersenne Twister (src/mcica_random_numbers.f90), ONE stream per call in the order
 *           (sub-column, column, layer) seeded with permuteseed -- reproduced exactly, including for calls
 *           that the library processes in several column passes.  140 generator CTAs (one per sub-column) each
 *           reach their place in the stream with a GF(2) polynomial jump-ahead and emit their run in parallel. */
int rrtmg_lw_b200_mcica_subcol_lw(int iplon, int ncol, int ld, int nlay, int icld, int permuteseed, int* irng,
                                  const double* play, const double* cldfrac, const double* ciwp,
                                  const double* clwp, const double* rei, const double* rel, const double* tauc,
                                  const double* alpha, double* cldfmcl, double* ciwpmcl, double* clwpmcl,
                                  double* reicmcl, double* relqmcl, double* taucmcl, int memspace, void* stream);

/* Replaces get_alpha (src/mcica_subcol_gen_lw.f90:68-180). */
int rrtmg_lw_b200_get_alpha(int iplon, int ncol, int ld, int nlayers, int icld, int idcor, double decorr_con,
                            const double* dz, const double* lat, int juldat, const double* cldfrac,
                            double* alpha, int memspace, void* stream);

/* mcica_subcol_lw followed by rrtmg_lw in ONE device pass: the (140,ncol,nlay) sub-column
 * arrays are never materialised (140 mask bits per (column,layer) instead of 4x140 doubles).
 * Results are identical to calling the two entry points above in sequence.
 * icld_overlap in [0,5] selects the generator's overlap (the value rrtmg_lw itself would see
 * is min(icld,3)->2 clamp per :470, which does not change its arithmetic). */
int rrtmg_lw_b200_rrtmg_lw_mcica(int ncol, int ld, int nlay, int icld_overlap, int idrv, int permuteseed, int irng,
                                 const double* play, const double* plev, const double* tlay, const double* tlev,
                                 const double* tsfc, const double* h2ovmr, const double* o3vmr,
                                 const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                                 const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                                 const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                                 int iceflglw, int liqflglw, const double* cldfrac, const double* ciwp,
                                 const double* clwp, const double* rei, const double* rel, const double* tauc,
                                 const double* alpha, const double* tauaer, double* uflx, double* dflx, double* hr,
                                 double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                                 int memspace, void* stream);

/* Wait for all enqueued work and return the sticky device status (first error wins);
 * col/lay (1-based, may be NULL) receive the first offending column and layer. */
int rrtmg_lw_b200_sync(int* col, int* lay);

/* Stage dump of ONE column: rrtmg_lw_b200_debug_capture(icol + 1) arms the capture of caller column
 * icol (0-based) for the following rrtmg_lw/rrtmg_lw_mcica calls (0 disarms);
 * rrtmg_lw_b200_debug_stages(icol, ...) reads it back.  Host pointers, any may be NULL:
 *   ints6 (6,nlay) jp,jt,jt1,indself,indfor,indminor ; facs (4,nlay) fac00,fac01,fac10,fac11
 *   taug,fracs (nlay,140) ; itr (2,nlay,140) ; scal = {laytrop, pwvcm} ; mask (nlay,140) 0/1 bytes */
int rrtmg_lw_b200_debug_capture(int icol_plus_1);
int rrtmg_lw_b200_debug_stages(int icol, int* ints6, double* facs, double* taug, double* fracs, int* itr,
                               double* scal, unsigned char* mask);

/* Readback of an init-time table for parity tests: "tau_tbl","exp_tbl","tfn_tbl","rwgt", the rrlw_con scalars
 * "heatfac" / "fluxfac" (what a column-mode driver needs for its idrv adjustment, src/rrtmg_lw.1col.f90:587-610) or
 * "<band>:<reduced array name>" (e.g. "3:absa", reference element order).  Returns the element
 * count (0 = unknown name); copies min(count, cap) doubles when out != NULL. */
long rrtmg_lw_b200_get_table(const char* name, double* out, long cap);

/* Multi-GPU from one host process: the library keeps one context (tables, scratch, streams, options) per
 * CUDA device.  rrtmg_lw_b200_set_device(d) makes the following calls of this process run on device d (and makes d
 * the thread's current CUDA device); d = -1 (default) follows whatever device is current when an entry is called, so
 * one-process-per-GPU hosts need not call it.  Each device is initialised with its own rrtmg_lw_b200_ini call. */
int rrtmg_lw_b200_set_device(int device);

/* Upper bound of columns processed per internal pass of a device-pointer call (default 262144; scratch is
 * sized to min(ncol, max_columns): about 1.3 KB per (column, layer) on the default two-phase path). */
int rrtmg_lw_b200_set_chunk(int max_columns);

/* Tuning switches (no reference counterpart).
 * "band_kernel": 4 (default) = two-phase path: inatm + setcoef + taumol with one thread per (column, layer), then
 *   cldprmc + rtrnmc sweep kernels with one thread per (column, band group) that stream the 8-byte layer optical
 *   depths; 2 = fused per-band kernels (one thread per (column, band), 16-byte sweep state; the optional FP32 mode
 *   runs on these).
 * "sweep_groups": 0 (default) = one sweep kernel per band (see "tail_group"); 1 = sweep kernels over the band groups
 *   {1,2} {3} {4} {5} {6,7} {8,9} {10-12} {13-16} (fewer launches and flux slots, more registers per thread: slower).
 * "tail_group": 1 (default) = bands 13-16 (10 g-points together) run as one sweep kernel when the band range holds
 *   all four; 0 = one kernel each.
 * "band_lo" / "band_hi": band range istart..iend of the following runs (default 1..16): only those bands enter the
 *   flux sums, as in the column-mode program's per-band output (src/rrtmg_lw.1col.f90:463-467); two-phase path.
 * "concurrency": 1 (default) = independent kernels of a pass run on side streams of the library, forked from and joined
 *   to the caller's stream: the McICA generator and the cloud optics beside taumol, and -- for passes of up to 393 216
 *   columns, where the overlap of the kernels' tails pays -- the sweep kernels beside each other; 0 = everything on the
 *   caller's stream; 2 / 3 = only the first / only the second half of that schedule, 4 / 5 = the sweeps over two /
 *   three streams (A/B runs; profiles/r02_stream_schedule_experiment.log).
 * "mcica_beside": n = 1..4: the KISS generator and the cloud optics run as ONE work-queue kernel with n CTAs per SM resident
 *   beside taumol (launched ahead of it on a high-priority stream), plus a full-grid launch after taumol for the items left;
 *   0 (default) = generator and cloud optics as their own kernels.  A measured and rejected scheduling variant (bit-identical
 *   results; the step time does not change, profiles/r02_mcica_beside_taumol_experiment.log), kept for A/B runs.
 * "kiss_sparse": 1 (default) = the KISS generator (irng = 0, icld 1 / 2) stops at the highest layer with a cloud
 *   fraction and steps over the draws of cloud-free layers with closed-form jumps (those draws cannot reach the
 *   mask); 0 = every number of the reference's stream is drawn.  The masks are bit-identical either way.
 * "out_ld": column extent of the OUTPUT arrays of the following device-pointer calls when it differs from ld (0 =
 *   same as ld): a call on a column sub-range of a resident batch (inputs offset, ld = the batch's extent) can write
 *   its fluxes straight into a packed, chunk-sized send buffer.  out_ld >= ncol.
 * "release_scratch": frees the library-owned scratch now (any value); the next call re-allocates what it needs.
 * "host_chunk": columns per staged pass of a host-pointer call (default 32768); passes alternate between two
 *   staging arenas so the H2D copy of pass k+1 and the D2H copy of pass k-1 overlap the kernels of pass k.
 * "precision": 64 (default) or 32 = the optional FP32 mode of BASELINE.json's north star: float band kernels
 *   (taumol, cldprmc, rtrnmc) with double flux accumulators for clear sky and the fused McICA entry, within
 *   0.05 W m-2 / 0.01 K day-1 of the FP64 results; setcoef indices and McICA masks are unchanged.
 * "mt_global_ncol" / "mt_col_offset" (irng = 1 only): the columns of the following calls are columns
 *   [offset, offset + ncol) of a batch of mt_global_ncol columns; the single Mersenne-Twister stream of that batch
 *   (src/mcica_subcol_gen_lw.f90:523-530) is reproduced for the block, so a column-sharded run equals the unsharded
 *   one (0 = off: every call is its own stream). */
int rrtmg_lw_b200_set_option(const char* name, int value);

/* Multi-GPU: the output gather behind the C ABI.  One rank per GPU, started by the host's launcher; columns
 * are block-partitioned over the ranks and nothing is exchanged on the data path.  A host that wants the ranks' output
 * blocks on one rank:
 *   rank 0:     rrtmg_lw_b200_comm_unique_id(id)   128 bytes, which the host broadcasts to every rank (MPI_Bcast, ...)
 *   all ranks:  rrtmg_lw_b200_comm_init(nranks, rank, id)
 *   per step:   rrtmg_lw_b200_gather(send, count, recv, root, stream)
 *               every rank contributes `count` doubles of DEVICE memory (e.g. a packed (rows, columns) output buffer
 *               written with the "out_ld" option); the root receives rank r's block at recv + r * count (recv: DEVICE
 *               memory, nranks * count doubles, ignored elsewhere).  NCCL send / receive pairs, enqueued on `stream`
 *               after the rank's kernels; several gathers of column chunks may be in flight while later chunks compute.
 *               rrtmg_lw_b200_comm_finalize()
 * NCCL (libnccl.so.2) is loaded at the first of these calls; the library has no link-time dependency on it.  Errors:
 * RRTMG_LW_ERR_CUDA (NCCL missing or failing), RRTMG_LW_ERR_NOT_INIT (gather before comm_init), RRTMG_LW_ERR_ARG. */
int rrtmg_lw_b200_comm_unique_id(void* id128);
int rrtmg_lw_b200_comm_init(int nranks, int rank, const void* id128);
int rrtmg_lw_b200_gather(const double* send, long count, double* recv, int root, void* stream);
int rrtmg_lw_b200_comm_finalize(void);

/* Host-only self test of the n-step jumps of the KISS generator used by the sub-column-parallel
 * McICA kernel (src/mcica_subcol_gen_lw.f90:731-741 stepped n times == one jump); returns the number
 * of mismatches (0 = ok).  Needs no CUDA device. */
int rrtmg_lw_b200_kiss_jump_selftest(void);
/* Host-only self test of the Mersenne-Twister jump-ahead of the irng = 1 path (src/mcica_random_numbers.f90:134-152
 * stepped n times == the GF(2) polynomial jump by n, for several n up to 10^6): builds the jump table (characteristic
 * polynomial by Berlekamp-Massey, x^(2^b) mod phi by squaring) and returns the number of mismatching state words
 * (0 = ok, -1 = the table could not be built).  Needs no CUDA device. */
int rrtmg_lw_b200_mt_jump_selftest(void);
/* Diagnostic of the "mcica_beside" variant: globaltimer stamps (ns) of the work-queue kernel's last pass -- [0] / [1] first
 * CTA start / last CTA end of the resident launch, [2] / [3] of the mop-up launch, [4] items taken when the mop-up began.
 * Synchronises the device. */
void rrtmg_lw_b200_kiss_work_stats(unsigned long long* out8);
/* Number of kernels launched by this library since the last call with reset != 0. */
long rrtmg_lw_b200_launch_count(int reset);
/* CUDA-event time (ms) spent in the band kernels / all kernels of the last HOST- or DEVICE-mode
 * call when timing capture is enabled with rrtmg_lw_b200_set_timing(1). */
int rrtmg_lw_b200_set_timing(int enable);
int rrtmg_lw_b200_last_timing(float* ms_band_kernels, float* ms_all_kernels);
/* Same capture, per band: ms16[b-1] = CUDA-event time of band b's kernel summed over the passes (fused band
 * kernels; on the two-phase path slot 0 is the taumol kernel and slots 1.. the sweep kernels). */
int rrtmg_lw_b200_last_band_timing(float* ms16);
/* Same capture, per timed stage of the pipeline between the prep kernels and the finalize kernel: returns the number
 * of stages n and copies min(n, cap) times (two-phase path: taumol, then each sweep kernel in launch order). */
int rrtmg_lw_b200_last_stage_timing(float* ms, int cap);
/* (column, layer), both 1-based, of the input error most recently reported by a compute entry or by
 * rrtmg_lw_b200_sync; the column counts from the first column of the CALL, also for host-pointer calls that the
 * library cuts into staged passes. */
int rrtmg_lw_b200_last_error(int* col, int* lay);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (an FMA = 2 flops); the FP64
 * roofline denominator used by bench.py. */
int rrtmg_lw_b200_fp64_peak(double* tflops);

const char* rrtmg_lw_b200_strerror(int code);
const char* rrtmg_lw_b200_version(void);
void rrtmg_lw_b200_finalize(void);

#ifdef __cplusplus
}
#endif
#endif
