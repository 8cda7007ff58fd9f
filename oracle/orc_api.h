/* orc_api.h -- C entry points of the CPU ORACLE (test infrastructure only; see orc_common.h).
 * Argument lists mirror the reference's GCM entry points:
 *   rrtmg_lw_ini     src/rrtmg_lw_init.f90:47
 *   rrtmg_lw         src/rrtmg_lw_rad.f90:99-108
 *   mcica_subcol_lw  src/mcica_subcol_gen_lw.f90:183-185
 *   get_alpha        src/mcica_subcol_gen_lw.f90:68-69
 * with every array in the reference's Fortran layout and an explicit leading dimension
 * `ld` (>= ncol) for the column axis.
 */
#ifndef ORC_API_H
#define ORC_API_H
#ifdef __cplusplus
extern "C" {
#endif

int orc_ini(double cpdair, const double* static_tables, const double* kdata, int default_real_kind);

/* band range istart..iend of the following orc_rrtmg_lw / orc_column_stages runs (1..16 by default): only those bands
 * enter the flux sums, and band 16 alone takes the totplk16 Planck variant (src/rrtmg_lw_setcoef.f90:233-251) */
int orc_set_band_range(int istart, int iend);

int orc_rrtmg_lw(int ncol, int ld, int nlay, int* icld, int idrv, const double* play, const double* plev,
                 const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                 const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                 const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr,
                 const double* ccl4vmr, const double* emis, int inflglw, int iceflglw, int liqflglw,
                 const double* cldfmcl, const double* taucmcl, const double* ciwpmcl, const double* clwpmcl,
                 const double* reicmcl, const double* relqmcl, const double* tauaer, double* uflx, double* dflx,
                 double* hr, double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                 int nthreads);

/* rrtmg_lw of src/rrtmg_lw_rad.nomcica.f90:99-108 (no McICA): cldfr, cicewp, cliqwp, reice, reliq (ncol, nlay);
 * taucld (nbndlw, ncol, nlay); icld = 1 -> rtrn (random overlap), else rtrnmr (maximum-random) */
int orc_rrtmg_lw_nomcica(int ncol, int ld, int nlay, int* icld, int idrv, const double* play, const double* plev,
                         const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                         const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                         const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr,
                         const double* ccl4vmr, const double* emis, int inflglw, int iceflglw, int liqflglw,
                         const double* cldfr, const double* taucld, const double* cicewp, const double* cliqwp,
                         const double* reice, const double* reliq, const double* tauaer, double* uflx, double* dflx,
                         double* hr, double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                         int nthreads);

int orc_mcica_subcol_lw(int ncol, int ld, int nlay, int icld, int permuteseed, int* irng, const double* play,
                        const double* cldfrac, const double* ciwp, const double* clwp, const double* rei,
                        const double* rel, const double* tauc, const double* alpha, double* cldfmcl,
                        double* ciwpmcl, double* clwpmcl, double* reicmcl, double* relqmcl, double* taucmcl);

int orc_get_alpha(int ncol, int ld, int nlayers, int icld, int idcor, double decorr_con, const double* dz,
                  const double* lat, int juldat, const double* cldfrac, double* alpha);

/* Stage dump for ONE column (index icol, 0-based) of an rrtmg_lw call: integer outputs of
 * setcoef, taug/fracs, table indices of rtrnmc.  Any output pointer may be NULL.
 *   ints6   : (6, nlay)  jp, jt, jt1, indself, indfor, indminor  (indself = 0 above laytrop)
 *   facs    : (4, nlay)  fac00, fac01, fac10, fac11
 *   taug, fracs : (nlay, 140)
 *   itr     : (2, nlay, 140)  itr|itgas , ittot  (-1 where the branch takes no lookup)
 *   scal    : [laytrop, pwvcm]
 */
int orc_column_stages(int icol, int ncol, int ld, int nlay, int icld, int idrv, const double* play,
                      const double* plev, const double* tlay, const double* tlev, const double* tsfc,
                      const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                      const double* n2ovmr, const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                      const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                      int iceflglw, int liqflglw, const double* cldfmcl, const double* taucmcl,
                      const double* ciwpmcl, const double* clwpmcl, const double* reicmcl, const double* relqmcl,
                      const double* tauaer, int* ints6, double* facs, double* taug, double* fracs, int* itr,
                      double* scal);

/* raw generator access for known-answer tests */
void orc_mt_stream(int seed, int n, unsigned int* ints_out, double* reals_out);
void orc_kiss_stream(int s1, int s2, int s3, int s4, int n, double* reals_out, int* state_out);
/* reduced-table / lookup-table readback: name in {"tau_tbl","exp_tbl","tfn_tbl","rwgt"} or
 * "<band>:<reduced array>" e.g. "3:absa"; returns element count (0 if unknown) */
long orc_get_table(const char* name, double* out, long cap);

/* branch-coverage counters of taumol / rtrnmc (143 of them, orc_common.h): counted only by the coverage build
 * liborc_cov.so (same arithmetic + counters); the library timed as the CPU baseline carries no instrumentation */
int orc_branch_counts(long long* out, int cap, int reset);

/* op-counting build (liborc_count.so) only: {add/sub, mul, div, exp, log, pow} executed by the calling
 * thread since the last reset; -1 in the normal build */
void orc_opcount(long long* out6, int reset);

#ifdef __cplusplus
}
#endif
#endif
