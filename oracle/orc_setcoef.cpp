// orc_setcoef.cpp -- ORACLE (test infrastructure): restatement of setcoef.
// Follows src/rrtmg_lw_setcoef.f90:169-432 statement by statement.
#include "orc_common.h"

namespace orc {

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void setcoef(Column& c, int istart, int idrv) {
    const int nlayers = c.nlayers;
    const double stpfac = 296. / 1013.;  // :171

    int indbound = (int)(c.tbound - 159.);  // :173 real->int truncation
    indbound = clampi(indbound, 1, 180);
    double tbndfrac = c.tbound - 159. - (double)indbound;
    int indlev0 = (int)(c.tz[0] - 159.);
    indlev0 = clampi(indlev0, 1, 180);
    double t0frac = c.tz[0] - 159. - (double)indlev0;
    c.laytrop = 0;

    for (int lay = 1; lay <= nlayers; ++lay) {
        const double tav = c.tavel[lay - 1], pav = c.pavel[lay - 1];
        int indlay = clampi((int)(tav - 159.), 1, 180);
        double tlayfrac = tav - 159. - (double)indlay;
        int indlev = clampi((int)(c.tz[lay] - 159.), 1, 180);
        double tlevfrac = c.tz[lay] - 159. - (double)indlev;
        double dbdtlev, dbdtlay;

        // bands 1-15 :209-227
        for (int iband = 1; iband <= 15; ++iband) {
            if (lay == 1) {
                dbdtlev = G.tp(indbound + 1, iband) - G.tp(indbound, iband);
                c.plankbnd[iband - 1] = c.semiss[iband - 1] * (G.tp(indbound, iband) + tbndfrac * dbdtlev);
                dbdtlev = G.tp(indlev0 + 1, iband) - G.tp(indlev0, iband);
                c.PLEV(0, iband) = G.tp(indlev0, iband) + t0frac * dbdtlev;
                if (idrv == 1) {
                    dbdtlev = G.tpd(indbound + 1, iband) - G.tpd(indbound, iband);
                    c.dplankbnd_dt[iband - 1] = c.semiss[iband - 1] * (G.tpd(indbound, iband) + tbndfrac * dbdtlev);
                }
            }
            dbdtlev = G.tp(indlev + 1, iband) - G.tp(indlev, iband);
            dbdtlay = G.tp(indlay + 1, iband) - G.tp(indlay, iband);
            c.PLAY(lay, iband) = G.tp(indlay, iband) + tlayfrac * dbdtlay;
            c.PLEV(lay, iband) = G.tp(indlev, iband) + tlevfrac * dbdtlev;
        }
        // band 16 :233-269  (totplk16 variant only when istart == 16; T8 incl. the :244 mix)
        const int iband = 16;
        if (istart == 16) {
            if (lay == 1) {
                dbdtlev = G.totplk16[indbound] - G.totplk16[indbound - 1];
                c.plankbnd[15] = c.semiss[15] * (G.totplk16[indbound - 1] + tbndfrac * dbdtlev);
                if (idrv == 1) {
                    dbdtlev = G.totplk16deriv[indbound] - G.totplk16deriv[indbound - 1];
                    c.dplankbnd_dt[15] = c.semiss[15] * (G.totplk16deriv[indbound - 1] + tbndfrac * dbdtlev);
                }
                dbdtlev = G.tp(indlev0 + 1, iband) - G.tp(indlev0, iband);
                c.PLEV(0, iband) = G.totplk16[indlev0 - 1] + t0frac * dbdtlev;
            }
            dbdtlev = G.totplk16[indlev] - G.totplk16[indlev - 1];
            dbdtlay = G.totplk16[indlay] - G.totplk16[indlay - 1];
            c.PLAY(lay, iband) = G.totplk16[indlay - 1] + tlayfrac * dbdtlay;
            c.PLEV(lay, iband) = G.totplk16[indlev - 1] + tlevfrac * dbdtlev;
        } else {
            if (lay == 1) {
                dbdtlev = G.tp(indbound + 1, iband) - G.tp(indbound, iband);
                c.plankbnd[15] = c.semiss[15] * (G.tp(indbound, iband) + tbndfrac * dbdtlev);
                if (idrv == 1) {
                    dbdtlev = G.tpd(indbound + 1, iband) - G.tpd(indbound, iband);
                    c.dplankbnd_dt[15] = c.semiss[15] * (G.tpd(indbound, iband) + tbndfrac * dbdtlev);
                }
                dbdtlev = G.tp(indlev0 + 1, iband) - G.tp(indlev0, iband);
                c.PLEV(0, iband) = G.tp(indlev0, iband) + t0frac * dbdtlev;
            }
            dbdtlev = G.tp(indlev + 1, iband) - G.tp(indlev, iband);
            dbdtlay = G.tp(indlay + 1, iband) - G.tp(indlay, iband);
            c.PLAY(lay, iband) = G.tp(indlay, iband) + tlayfrac * dbdtlay;
            c.PLEV(lay, iband) = G.tp(indlev, iband) + tlevfrac * dbdtlev;
        }

        // pressure interpolation :276-284
        double plog = std::log(pav);
        int jp = (int)(36. - 5 * (plog + 0.04));
        jp = clampi(jp, 1, 58);
        c.jp[lay - 1] = jp;
        int jp1 = jp + 1;
        double fp = 5. * (G.preflog[jp - 1] - plog);
        // temperature interpolation :293-306
        int jt = (int)(3. + (tav - G.tref[jp - 1]) / 15.);
        jt = clampi(jt, 1, 4);
        c.jt[lay - 1] = jt;
        double ft = ((tav - G.tref[jp - 1]) / 15.) - (double)(jt - 3);
        int jt1 = (int)(3. + (tav - G.tref[jp1 - 1]) / 15.);
        jt1 = clampi(jt1, 1, 4);
        c.jt1[lay - 1] = jt1;
        double ft1 = ((tav - G.tref[jp1 - 1]) / 15.) - (double)(jt1 - 3);
        double water = c.WKL(1, lay) / c.coldry[lay - 1];
        double scalefac = pav * stpfac / tav;

        const double coldry = c.coldry[lay - 1];
        if (!(plog <= 4.56)) {
            // lower atmosphere :313-366
            c.laytrop = c.laytrop + 1;
            c.forfac[lay - 1] = scalefac / (1. + water);
            double factor = (332.0 - tav) / 36.0;
            c.indfor[lay - 1] = std::min(2, std::max(1, (int)factor));
            c.forfrac[lay - 1] = factor - (double)c.indfor[lay - 1];
            c.selffac[lay - 1] = water * c.forfac[lay - 1];
            factor = (tav - 188.0) / 7.2;
            c.indself[lay - 1] = std::min(9, std::max(1, (int)factor - 7));
            c.selffrac[lay - 1] = factor - (double)(c.indself[lay - 1] + 7);
            c.scaleminor[lay - 1] = pav / tav;
            c.scaleminorn2[lay - 1] = (pav / tav) * (c.wbrodl[lay - 1] / (coldry + c.WKL(1, lay)));
            factor = (tav - 180.8) / 7.2;
            c.indminor[lay - 1] = std::min(18, std::max(1, (int)factor));
            c.minorfrac[lay - 1] = factor - (double)c.indminor[lay - 1];
            c.rat_h2oco2[lay - 1] = G.chi(1, jp) / G.chi(2, jp);
            c.rat_h2oco2_1[lay - 1] = G.chi(1, jp + 1) / G.chi(2, jp + 1);
            c.rat_h2oo3[lay - 1] = G.chi(1, jp) / G.chi(3, jp);
            c.rat_h2oo3_1[lay - 1] = G.chi(1, jp + 1) / G.chi(3, jp + 1);
            c.rat_h2on2o[lay - 1] = G.chi(1, jp) / G.chi(4, jp);
            c.rat_h2on2o_1[lay - 1] = G.chi(1, jp + 1) / G.chi(4, jp + 1);
            c.rat_h2och4[lay - 1] = G.chi(1, jp) / G.chi(6, jp);
            c.rat_h2och4_1[lay - 1] = G.chi(1, jp + 1) / G.chi(6, jp + 1);
            c.rat_n2oco2[lay - 1] = G.chi(4, jp) / G.chi(2, jp);
            c.rat_n2oco2_1[lay - 1] = G.chi(4, jp + 1) / G.chi(2, jp + 1);
        } else {
            // upper atmosphere :369-411
            c.forfac[lay - 1] = scalefac / (1. + water);
            double factor = (tav - 188.0) / 36.0;
            c.indfor[lay - 1] = 3;
            c.forfrac[lay - 1] = factor - 1.0;
            c.selffac[lay - 1] = water * c.forfac[lay - 1];
            c.scaleminor[lay - 1] = pav / tav;
            c.scaleminorn2[lay - 1] = (pav / tav) * (c.wbrodl[lay - 1] / (coldry + c.WKL(1, lay)));
            factor = (tav - 180.8) / 7.2;
            c.indminor[lay - 1] = std::min(18, std::max(1, (int)factor));
            c.minorfrac[lay - 1] = factor - (double)c.indminor[lay - 1];
            c.rat_h2oco2[lay - 1] = G.chi(1, jp) / G.chi(2, jp);
            c.rat_h2oco2_1[lay - 1] = G.chi(1, jp + 1) / G.chi(2, jp + 1);
            c.rat_o3co2[lay - 1] = G.chi(3, jp) / G.chi(2, jp);
            c.rat_o3co2_1[lay - 1] = G.chi(3, jp + 1) / G.chi(2, jp + 1);
        }
        // column amounts (identical in both branches) :354-366 / :399-411
        c.colh2o[lay - 1] = 1.e-20 * c.WKL(1, lay);
        c.colco2[lay - 1] = 1.e-20 * c.WKL(2, lay);
        c.colo3[lay - 1] = 1.e-20 * c.WKL(3, lay);
        c.coln2o[lay - 1] = 1.e-20 * c.WKL(4, lay);
        c.colco[lay - 1] = 1.e-20 * c.WKL(5, lay);
        c.colch4[lay - 1] = 1.e-20 * c.WKL(6, lay);
        c.colo2[lay - 1] = 1.e-20 * c.WKL(7, lay);
        if (c.colco2[lay - 1] == 0.) c.colco2[lay - 1] = 1.e-32 * coldry;
        if (c.colo3[lay - 1] == 0.) c.colo3[lay - 1] = 1.e-32 * coldry;
        if (c.coln2o[lay - 1] == 0.) c.coln2o[lay - 1] = 1.e-32 * coldry;
        if (c.colco[lay - 1] == 0.) c.colco[lay - 1] = 1.e-32 * coldry;
        if (c.colch4[lay - 1] == 0.) c.colch4[lay - 1] = 1.e-32 * coldry;
        c.colbrd[lay - 1] = 1.e-20 * c.wbrodl[lay - 1];

        // :421-429
        double compfp = 1. - fp;
        c.fac10[lay - 1] = compfp * ft;
        c.fac00[lay - 1] = compfp * (1. - ft);
        c.fac11[lay - 1] = fp * ft1;
        c.fac01[lay - 1] = fp * (1. - ft1);
        c.selffac[lay - 1] = c.colh2o[lay - 1] * c.selffac[lay - 1];
        c.forfac[lay - 1] = c.colh2o[lay - 1] * c.forfac[lay - 1];
    }
}

}  // namespace orc
