/* lw_c_caller.c -- a non-Python caller of the C ABI (include/rrtmg_lw_b200.h), written the way the Fortran shim
 * (fortran/rrtmg_lw_b200_shim.f90) drives it:
 *   1. the unreduced k-arrays are handed over ONE BY ONE with rrtmg_lw_b200_kdata_put, in the order of
 *      include/rrtmg_lw_kdata_layout.h (what the shim does after the reference's lw_kgb01..16), then
 *      rrtmg_lw_b200_ini(cpdair, NULL, 0, kind) consumes the staged arrays;
 *   2. rrtmg_lw_b200_rrtmg_lw is called with HOST arrays whose column extent ld is LARGER than ncol (a GCM's pcols
 *      chunk): only columns 0..ncol-1 may be read or written;
 *   3. the result is checked against the CPU oracle (oracle/orc_api.h, test infrastructure) fed with the same
 *      numbers: fluxes within 1e-6 W m-2, heating rates within 1e-6 K day-1 (BASELINE.json north_star).
 * Exit code 0 = pass, 1 = mismatch, 2 = the library reported an error (e.g. no CUDA device: there is no CPU path).
 * Built by __graft_entry__.build() with gcc; run by tests/test_c_caller.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/rrtmg_lw_b200.h"
#include "../../include/rrtmg_lw_kdata_layout.h"
#include "../../oracle/orc_api.h"

#define NLAY 40
#define NCOL 37
#define LD 48 /* > NCOL on purpose */
#define NG 140

static double frand(unsigned long long* s) { /* splitmix64 -> (0,1) */
    unsigned long long z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}

static double* arr(size_t n, double fill) {
    double* p = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; ++i) p[i] = fill;
    return p;
}

int main(int argc, char** argv) {
    const char* static_path = argc > 1 ? argv[1] : "rrtmg_lw_b200/data/lw_static_tables.bin";
    unsigned long long seed = 20261017ULL;
    /* ---- synthetic k-data, array by array (positive, smooth in the leading index, decreasing in g) */
    double* blob = arr(RRTMG_LW_KDATA_DOUBLES, 0.0);
    size_t off = 0;
    for (int k = 0; k < RRTMG_LW_NKARRAYS; ++k) {
        const rrtmg_lw_karray_desc* d = &RRTMG_LW_KARRAYS[k];
        const size_t cnt = (size_t)16 * d->lead;
        double* a = blob + off;
        const int frac = strncmp(d->name, "fracref", 7) == 0;
        for (int g = 0; g < 16; ++g)
            for (int i = 0; i < d->lead; ++i) {
                const size_t e = d->gfirst ? (size_t)g + 16 * (size_t)i : (size_t)i + (size_t)d->lead * g;
                if (frac) a[e] = (1.0 + 0.02 * (i % 7) + 0.1 * frand(&seed)) / 16.0;
                else a[e] = 1.0e-3 * exp(0.35 * g) * (1.0 + 0.3 * sin(0.37 * i + d->band)) * (0.8 + 0.4 * frand(&seed));
            }
        int rc = rrtmg_lw_b200_kdata_put(d->band, d->name, a, (long)cnt);
        if (rc) { fprintf(stderr, "kdata_put(%d,%s) -> %d\n", d->band, d->name, rc); return 2; }
        off += cnt;
    }
    /* an unknown (band, name) and a wrong size are refused */
    if (rrtmg_lw_b200_kdata_put(3, "nonsense", blob, 16) == RRTMG_LW_OK) return 1;
    if (rrtmg_lw_b200_kdata_put(1, "kao", blob, 17) == RRTMG_LW_OK) return 1;
    int rc = rrtmg_lw_b200_ini(1004.0, NULL, 0, 4);
    if (rc) { fprintf(stderr, "rrtmg_lw_b200_ini -> %d (%s)\n", rc, rrtmg_lw_b200_strerror(rc)); return 2; }

    /* ---- the checker, initialised with the same numbers (blob) and the reference's static tables */
    FILE* f = fopen(static_path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", static_path); return 2; }
    double* st = arr(9110, 0.0);
    if (fread(st, sizeof(double), 9110, f) != 9110) { fprintf(stderr, "short static table file\n"); return 2; }
    fclose(f);
    if (orc_ini(1004.0, st, blob, 4)) { fprintf(stderr, "orc_ini failed\n"); return 2; }

    /* ---- a clear-sky + aerosol batch with dF/dTs, Fortran layout (column fastest), ld > ncol; the columns beyond
     * ncol hold NaN on input and a sentinel on output */
    const size_t n2 = (size_t)LD * NLAY, n2p = (size_t)LD * (NLAY + 1);
    double *play = arr(n2, NAN), *plev = arr(n2p, NAN), *tlay = arr(n2, NAN), *tlev = arr(n2p, NAN), *tsfc = arr(LD, NAN);
    double *h2o = arr(n2, NAN), *o3 = arr(n2, NAN), *co2 = arr(n2, NAN), *ch4 = arr(n2, NAN), *n2o = arr(n2, NAN),
           *o2 = arr(n2, NAN), *c11 = arr(n2, NAN), *c12 = arr(n2, NAN), *c22 = arr(n2, NAN), *ccl4 = arr(n2, NAN);
    double *emis = arr((size_t)LD * 16, NAN), *tauaer = arr((size_t)LD * NLAY * 16, NAN);
    for (int i = 0; i < NCOL; ++i) {
        const double ps = 1013.25 * (0.85 + 0.2 * frand(&seed)), dt = 10.0 * (frand(&seed) - 0.5);
        for (int l = 0; l <= NLAY; ++l) {
            const double sig = exp(-11.5 * pow((double)l / NLAY, 1.6));
            plev[i + LD * l] = ps * sig;
            tlev[i + LD * l] = 216.0 + 78.0 * pow(sig, 0.19) + dt;
        }
        tsfc[i] = tlev[i] + 2.0 * (frand(&seed) - 0.5);
        for (int l = 0; l < NLAY; ++l) {
            const size_t a = i + (size_t)LD * l;
            play[a] = 0.5 * (plev[a] + plev[a + LD]);
            tlay[a] = 0.5 * (tlev[a] + tlev[a + LD]);
            const double pr = play[a] / ps;
            h2o[a] = fmax(3.0e-6, 0.02 * pr * pr * pr * (0.5 + frand(&seed)));
            o3[a] = 3.0e-8 + 8.0e-6 * exp(-0.5 * pow(log(play[a] / 10.0), 2));
            co2[a] = 4.0e-4; ch4[a] = 1.8e-6; n2o[a] = 3.2e-7; o2[a] = 0.209;
            c11[a] = 2.3e-10; c12[a] = 5.2e-10; c22[a] = 2.3e-10; ccl4[a] = 8.0e-11;
            for (int b = 0; b < 16; ++b) tauaer[i + (size_t)LD * (l + (size_t)NLAY * b)] = play[a] > 500.0 ? 0.01 * frand(&seed) : 0.0;
        }
        for (int b = 0; b < 16; ++b) emis[i + LD * b] = 0.96 + 0.04 * frand(&seed);
    }
    const double sentinel = -777.0;
    double* out[2][8];
    for (int w = 0; w < 2; ++w)
        for (int k = 0; k < 8; ++k) out[w][k] = arr(n2p, sentinel);
    int icld = 0;
    rc = rrtmg_lw_b200_rrtmg_lw(NCOL, LD, NLAY, &icld, 1, play, plev, tlay, tlev, tsfc, h2o, o3, co2, ch4, n2o, o2, c11, c12,
                                c22, ccl4, emis, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, tauaer, out[0][0], out[0][1],
                                out[0][2], out[0][3], out[0][4], out[0][5], out[0][6], out[0][7], RRTMG_LW_HOST, NULL);
    if (rc) { fprintf(stderr, "rrtmg_lw_b200_rrtmg_lw -> %d (%s)\n", rc, rrtmg_lw_b200_strerror(rc)); return 2; }
    int icld2 = 0;
    double dummy = 0.0;
    rc = orc_rrtmg_lw(NCOL, LD, NLAY, &icld2, 1, play, plev, tlay, tlev, tsfc, h2o, o3, co2, ch4, n2o, o2, c11, c12, c22, ccl4,
                      emis, 0, 0, 0, &dummy, &dummy, &dummy, &dummy, &dummy, &dummy, tauaer, out[1][0], out[1][1], out[1][2],
                      out[1][3], out[1][4], out[1][5], out[1][6], out[1][7], 1);
    if (rc) { fprintf(stderr, "oracle -> %d\n", rc); return 2; }
    const char* names[8] = {"uflx", "dflx", "hr", "uflxc", "dflxc", "hrc", "duflx_dt", "duflxc_dt"};
    int bad = 0;
    double worst = 0.0;
    for (int k = 0; k < 8; ++k) {
        const int nrow = (k == 2 || k == 5) ? NLAY : NLAY + 1;
        for (int l = 0; l < nrow; ++l)
            for (int i = 0; i < LD; ++i) {
                const double g = out[0][k][i + (size_t)LD * l], r = out[1][k][i + (size_t)LD * l];
                if (i >= NCOL) { /* padding columns untouched by the library */
                    if (g != sentinel) { if (!bad) fprintf(stderr, "%s: column %d >= ncol was written\n", names[k], i); bad = 1; }
                    continue;
                }
                const double e = fabs(g - r);
                if (!(e <= 1.0e-6)) { if (!bad) fprintf(stderr, "%s(%d,%d): %.12g vs %.12g\n", names[k], i, l, g, r); bad = 1; }
                if (e > worst) worst = e;
            }
    }
    if (!(out[0][0][0] > 100.0 && out[0][0][0] < 600.0)) bad = 1; /* a surface upward flux, not zeros */
    int col = 0, lay = 0;
    if (rrtmg_lw_b200_sync(&col, &lay) != RRTMG_LW_OK) bad = 1;
    printf("lw_c_caller: %d columns (ld %d) x %d layers, idrv=1, worst |gpu - oracle| = %.3g -> %s\n", NCOL, LD, NLAY, worst,
           bad ? "FAIL" : "PASS");
    rrtmg_lw_b200_finalize();
    return bad ? 1 : 0;
}
