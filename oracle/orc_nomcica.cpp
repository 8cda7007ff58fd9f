// orc_nomcica.cpp -- ORACLE (test infrastructure): restatement of the reference's NON-McICA cloudy path:
//   cldprop  src/rrtmg_lw_cldprop.f90:166-295   band cloud optical depths from grid-mean cloud properties
//   rtrn     src/rrtmg_lw_rtrn.f90:296-560       radiative transfer, random cloud overlap (icld = 1)
//   rtrnmr   src/rrtmg_lw_rtrnmr.f90:296-806     radiative transfer, maximum-random overlap (icld = 2, 3)
// behind rrtmg_lw of src/rrtmg_lw_rad.nomcica.f90:99-588.  Statement order, branch structure and association follow
// the Fortran.  Two places of rtrnmr read variables the Fortran never assigned on that path (faccmb1/2(lev+1) and
// faccmb1d/2d(lev-1) when the cloudy layer starts a cloud block, istcld = 1; rat1 / rat2 before their first
// assignment cannot be reached): they are taken as 0 here, which is what zero-initialised stack memory gives and
// the only value that keeps radmod meaningful (documented in DESIGN.md as trap T17).
#include "orc_common.h"

namespace orc {

static inline double MAXV(double a, double b) { return a > b ? a : b; }
static inline double MINV(double a, double b) { return a < b ? a : b; }

// cldprop, rrtmg_lw_cldprop.f90:166-295.  tauc is (nbndlw, nlayers) as in the reference; taucloud (nlayers, nbndlw).
int cldprop(int nlayers, int inflag, int iceflag, int liqflag, const double* cldfrac, const double* tauc,
            const double* ciwp, const double* clwp, const double* rei, const double* rel, int& ncbands,
            std::vector<double>& taucloud) {
    static const int icb[3][16] = {{1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                                   {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                                   {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
    const double cldmin = 1.e-20;  // rrlw_cld / cldprop data statement
    ncbands = 1;
    std::vector<double> tauctot(nlayers, 0.);
    taucloud.assign((size_t)nlayers * NBND, 0.0);
    auto TC = [&](int lay, int ib) -> double& { return taucloud[(lay - 1) + (size_t)nlayers * (ib - 1)]; };
    for (int lay = 1; lay <= nlayers; ++lay)
        for (int ib = 1; ib <= NBND; ++ib) tauctot[lay - 1] = tauctot[lay - 1] + tauc[(ib - 1) + NBND * (lay - 1)];
    double abscoice[NBND], abscoliq[NBND];
    for (int lay = 1; lay <= nlayers; ++lay) {
        double cwp = ciwp[lay - 1] + clwp[lay - 1];
        if (cldfrac[lay - 1] >= cldmin && (cwp >= cldmin || tauctot[lay - 1] >= cldmin)) {
            if (inflag == 0) {
                ncbands = 16;
                for (int ib = 1; ib <= ncbands; ++ib) TC(lay, ib) = tauc[(ib - 1) + NBND * (lay - 1)];
            } else if (inflag == 1) {
                ncbands = 16;
                for (int ib = 1; ib <= ncbands; ++ib) TC(lay, ib) = G.abscld1 * cwp;
            } else if (inflag == 2) {
                double radice = rei[lay - 1];
                int iceind = 0, liqind = 0;
                if (ciwp[lay - 1] == 0.0) {
                    abscoice[0] = 0.0;
                    iceind = 0;
                } else if (iceflag == 0) {
                    if (radice < 10.0) return ORC_ERR_ICE_RADIUS;  // 'ICE RADIUS TOO SMALL'
                    abscoice[0] = G.absice0[0] + G.absice0[1] / radice;
                    iceind = 0;
                } else if (iceflag == 1) {
                    if (radice < 13.0 || radice > 130.) return ORC_ERR_ICE_RADIUS;
                    ncbands = 5;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoice[ib - 1] = G.absice1[0 + 2 * (ib - 1)] + G.absice1[1 + 2 * (ib - 1)] / radice;
                    iceind = 1;
                } else if (iceflag == 2) {
                    if (radice < 5.0 || radice > 131.0) return ORC_ERR_ICE_RADIUS;
                    ncbands = 16;
                    double factor = (radice - 2.) / 3.;
                    int index = (int)factor;
                    if (index == 43) index = 42;
                    double fint = factor - (double)index;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoice[ib - 1] = G.absice2[(index - 1) + 43 * (ib - 1)] +
                                           fint * (G.absice2[index + 43 * (ib - 1)] - (G.absice2[(index - 1) + 43 * (ib - 1)]));
                    iceind = 2;
                } else if (iceflag == 3) {
                    if (radice < 5.0 || radice > 140.0) return ORC_ERR_ICE_RADIUS;
                    ncbands = 16;
                    double factor = (radice - 2.) / 3.;
                    int index = (int)factor;
                    if (index == 46) index = 45;
                    double fint = factor - (double)index;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoice[ib - 1] = G.absice3[(index - 1) + 46 * (ib - 1)] +
                                           fint * (G.absice3[index + 46 * (ib - 1)] - (G.absice3[(index - 1) + 46 * (ib - 1)]));
                    iceind = 2;
                }
                if (clwp[lay - 1] == 0.0) {
                    abscoliq[0] = 0.0;
                    liqind = 0;
                    if (iceind == 1) iceind = 2;
                } else if (liqflag == 0) {
                    abscoliq[0] = G.absliq0;
                    liqind = 0;
                    if (iceind == 1) iceind = 2;
                } else if (liqflag == 1) {
                    double radliq = rel[lay - 1];
                    if (radliq < 2.5 || radliq > 60.) return ORC_ERR_LIQ_RADIUS;
                    int index = (int)(radliq - 1.5);
                    if (index == 0) index = 1;
                    if (index == 58) index = 57;
                    double fint = radliq - 1.5 - (double)index;
                    ncbands = 16;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoliq[ib - 1] = G.absliq1[(index - 1) + 58 * (ib - 1)] +
                                           fint * (G.absliq1[index + 58 * (ib - 1)] - (G.absliq1[(index - 1) + 58 * (ib - 1)]));
                    liqind = 2;
                }
                for (int ib = 1; ib <= ncbands; ++ib)
                    TC(lay, ib) = ciwp[lay - 1] * abscoice[icb[iceind][ib - 1] - 1] + clwp[lay - 1] * abscoliq[icb[liqind][ib - 1] - 1];
            }
        }
    }
    return ORC_OK;
}

// rtrn (maxrand = false) / rtrnmr (maxrand = true)
void rtrn_nomcica(Column& c, bool maxrand, const double* cldfrac, int ncbands, const std::vector<double>& taucloud,
                  int istart, int iend, int iout, int idrv) {
    const int nlayers = c.nlayers;
    const double wtdiff = 0.5, rec_6 = 0.166667;
    static const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33,
                                  1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    static const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40,
                                  -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062,
                                  0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const int ipat[3][16] = {{1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                                    {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                                    {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
    const double bpade = G.bpade, tblint = 10000.0;
    double secdiff[NBND];
    for (int ibnd = 1; ibnd <= NBND; ++ibnd) {
        if (ibnd == 1 || ibnd == 4 || ibnd >= 10) {
            secdiff[ibnd - 1] = 1.66;
        } else {
            secdiff[ibnd - 1] = a0[ibnd - 1] + a1[ibnd - 1] * std::exp(a2[ibnd - 1] * c.pwvcm);
            if (secdiff[ibnd - 1] > 1.80) secdiff[ibnd - 1] = 1.80;
            if (secdiff[ibnd - 1] < 1.50) secdiff[ibnd - 1] = 1.50;
        }
    }
    const int n1 = nlayers + 1;
    std::vector<double> urad(n1, 0.), drad(n1, 0.), clrurad(n1, 0.), clrdrad(n1, 0.), d_urad_dt(n1, 0.), d_clrurad_dt(n1, 0.);
    std::vector<double> atrans(n1, 0.), bbugas(n1, 0.), atot(n1, 0.), bbutot(n1, 0.);
    std::vector<int> icldlyr(n1 + 1, 0);
    std::vector<double> odcld((size_t)nlayers * NBND, 0.), abscld((size_t)nlayers * NBND, 0.), efclfrac((size_t)nlayers * NBND, 0.);
    auto LB = [&](std::vector<double>& v, int lay, int ib) -> double& { return v[(lay - 1) + (size_t)nlayers * (ib - 1)]; };
    auto TCL = [&](int lay, int ib) { return taucloud[(lay - 1) + (size_t)nlayers * (ib - 1)]; };
    auto CF = [&](int lev) { return cldfrac[lev - 1]; };
    for (int l = 0; l <= nlayers; ++l) {
        c.totuflux[l] = c.totdflux[l] = c.totuclfl[l] = c.totdclfl[l] = 0.;
        c.dtotuflux_dt[l] = c.dtotuclfl_dt[l] = 0.;
    }
    for (int lay = 1; lay <= nlayers; ++lay) {
        for (int ib = 1; ib <= ncbands; ++ib) {
            if (CF(lay) >= 1.e-6) {
                LB(odcld, lay, ib) = secdiff[ib - 1] * TCL(lay, ib);
                if (!maxrand) {
                    double transcld = std::exp(-LB(odcld, lay, ib));
                    LB(abscld, lay, ib) = 1. - transcld;
                    LB(efclfrac, lay, ib) = LB(abscld, lay, ib) * CF(lay);
                }
                icldlyr[lay] = 1;
            } else {
                LB(odcld, lay, ib) = 0.0;
                icldlyr[lay] = 0;
            }
        }
    }

    // maximum/random overlap factors, rtrnmr.f90:345-505 (index lev+1 up to nlayers+1; lev-1 down to 0)
    std::vector<double> faccld1(n1 + 1, 0.), faccld2(n1 + 1, 0.), facclr1(n1 + 1, 0.), facclr2(n1 + 1, 0.), faccmb1(n1 + 1, 0.),
        faccmb2(n1 + 1, 0.);
    std::vector<double> faccld1d(n1, 0.), faccld2d(n1, 0.), facclr1d(n1, 0.), facclr2d(n1, 0.), faccmb1d(n1, 0.), faccmb2d(n1, 0.);
    std::vector<int> istcld(n1 + 1, 0), istcldd(n1, 0);
    if (maxrand) {
        double rat1 = 0., rat2 = 0.;
        istcld[1] = 1;
        istcldd[nlayers] = 1;
        for (int lev = 1; lev <= nlayers; ++lev) {
            if (icldlyr[lev] == 1) {
                istcld[lev + 1] = 0;
                if (lev == nlayers) {
                    faccld1[lev + 1] = 0.; faccld2[lev + 1] = 0.; facclr1[lev + 1] = 0.; facclr2[lev + 1] = 0.;
                    faccmb1[lev + 1] = 0.; faccmb2[lev + 1] = 0.;
                } else if (CF(lev + 1) >= CF(lev)) {
                    faccld1[lev + 1] = 0.;
                    faccld2[lev + 1] = 0.;
                    if (istcld[lev] == 1) {
                        facclr1[lev + 1] = 0.;
                        facclr2[lev + 1] = 0.;
                        if (CF(lev) < 1.) facclr2[lev + 1] = (CF(lev + 1) - CF(lev)) / (1. - CF(lev));
                        facclr2[lev] = 0.;
                        faccld2[lev] = 0.;
                    } else {
                        double fmax = MAXV(CF(lev), CF(lev - 1));
                        if (CF(lev + 1) > fmax) {
                            facclr1[lev + 1] = rat2;
                            facclr2[lev + 1] = (CF(lev + 1) - fmax) / (1. - fmax);
                        } else if (CF(lev + 1) < fmax) {
                            facclr1[lev + 1] = (CF(lev + 1) - CF(lev)) / (CF(lev - 1) - CF(lev));
                            facclr2[lev + 1] = 0.;
                        } else {
                            facclr1[lev + 1] = rat2;
                            facclr2[lev + 1] = 0.;
                        }
                    }
                    if (facclr1[lev + 1] > 0. || facclr2[lev + 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                    else { rat1 = 0.; rat2 = 0.; }
                } else {
                    facclr1[lev + 1] = 0.;
                    facclr2[lev + 1] = 0.;
                    if (istcld[lev] == 1) {
                        faccld1[lev + 1] = 0.;
                        faccld2[lev + 1] = (CF(lev) - CF(lev + 1)) / CF(lev);
                        facclr2[lev] = 0.;
                        faccld2[lev] = 0.;
                    } else {
                        double fmin = MINV(CF(lev), CF(lev - 1));
                        if (CF(lev + 1) <= fmin) {
                            faccld1[lev + 1] = rat1;
                            faccld2[lev + 1] = (fmin - CF(lev + 1)) / fmin;
                        } else {
                            faccld1[lev + 1] = (CF(lev) - CF(lev + 1)) / (CF(lev) - fmin);
                            faccld2[lev + 1] = 0.;
                        }
                    }
                    if (faccld1[lev + 1] > 0. || faccld2[lev + 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                    else { rat1 = 0.; rat2 = 0.; }
                }
                if (istcld[lev] != 1) {
                    // max(0., ...) with a default-kind 0.: same value in either kind
                    faccmb1[lev + 1] = MAXV(0., MINV(CF(lev + 1) - CF(lev), CF(lev - 1) - CF(lev)));
                    faccmb2[lev + 1] = MAXV(0., MINV(CF(lev) - CF(lev + 1), CF(lev) - CF(lev - 1)));
                }
            } else {
                istcld[lev + 1] = 1;
            }
        }
        for (int lev = nlayers; lev >= 1; --lev) {
            if (icldlyr[lev] == 1) {
                istcldd[lev - 1] = 0;
                if (lev == 1) {
                    faccld1d[lev - 1] = 0.; faccld2d[lev - 1] = 0.; facclr1d[lev - 1] = 0.; facclr2d[lev - 1] = 0.;
                    faccmb1d[lev - 1] = 0.; faccmb2d[lev - 1] = 0.;
                } else if (CF(lev - 1) >= CF(lev)) {
                    faccld1d[lev - 1] = 0.;
                    faccld2d[lev - 1] = 0.;
                    if (istcldd[lev] == 1) {
                        facclr1d[lev - 1] = 0.;
                        facclr2d[lev - 1] = 0.;
                        if (CF(lev) < 1.) facclr2d[lev - 1] = (CF(lev - 1) - CF(lev)) / (1. - CF(lev));
                        facclr2d[lev] = 0.;
                        faccld2d[lev] = 0.;
                    } else {
                        double fmax = MAXV(CF(lev), CF(lev + 1));
                        if (CF(lev - 1) > fmax) {
                            facclr1d[lev - 1] = rat2;
                            facclr2d[lev - 1] = (CF(lev - 1) - fmax) / (1. - fmax);
                        } else if (CF(lev - 1) < fmax) {
                            facclr1d[lev - 1] = (CF(lev - 1) - CF(lev)) / (CF(lev + 1) - CF(lev));
                            facclr2d[lev - 1] = 0.;
                        } else {
                            facclr1d[lev - 1] = rat2;
                            facclr2d[lev - 1] = 0.;
                        }
                    }
                    if (facclr1d[lev - 1] > 0. || facclr2d[lev - 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                    else { rat1 = 0.; rat2 = 0.; }
                } else {
                    facclr1d[lev - 1] = 0.;
                    facclr2d[lev - 1] = 0.;
                    if (istcldd[lev] == 1) {
                        faccld1d[lev - 1] = 0.;
                        faccld2d[lev - 1] = (CF(lev) - CF(lev - 1)) / CF(lev);
                        facclr2d[lev] = 0.;
                        faccld2d[lev] = 0.;
                    } else {
                        double fmin = MINV(CF(lev), CF(lev + 1));
                        if (CF(lev - 1) <= fmin) {
                            faccld1d[lev - 1] = rat1;
                            faccld2d[lev - 1] = (fmin - CF(lev - 1)) / fmin;
                        } else {
                            faccld1d[lev - 1] = (CF(lev) - CF(lev - 1)) / (CF(lev) - fmin);
                            faccld2d[lev - 1] = 0.;
                        }
                    }
                    if (faccld1d[lev - 1] > 0. || faccld2d[lev - 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                    else { rat1 = 0.; rat2 = 0.; }
                }
                if (istcldd[lev] != 1) {
                    faccmb1d[lev - 1] = MAXV(0., MINV(CF(lev + 1) - CF(lev), CF(lev - 1) - CF(lev)));
                    faccmb2d[lev - 1] = MAXV(0., MINV(CF(lev) - CF(lev + 1), CF(lev) - CF(lev - 1)));
                }
            } else {
                istcldd[lev - 1] = 1;
            }
        }
    }

    int igc = 1;
    for (int iband = istart; iband <= iend; ++iband) {
        if (iout > 0 && iband >= 2) igc = G.ngs[iband - 2] + 1;
        int ib = ncbands == 1 ? ipat[0][iband - 1] : (ncbands == 5 ? ipat[1][iband - 1] : ipat[2][iband - 1]);
        do {
            double radld = 0., radclrd = 0.;
            int iclddn = 0;
            double bbd = 0.;
            double cldradd = 0., clrradd = 0., oldcld = 0., oldclr = 0., rad = 0.;
            for (int lev = nlayers; lev >= 1; --lev) {
                double plfrac = c.LG(c.fracs, lev, igc);
                double blay = c.PLAY(lev, iband);
                double dplankup = c.PLEV(lev, iband) - blay;
                double dplankdn = c.PLEV(lev - 1, iband) - blay;
                double odepth = secdiff[iband - 1] * c.LG(c.taut, lev, igc);
                if (odepth < 0.0) odepth = 0.0;
                if (icldlyr[lev] == 1) {
                    iclddn = 1;
                    double odtot = odepth + LB(odcld, lev, ib);
                    double gassrc, bbdtot;
                    if (odtot < 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        atot[lev] = odtot - 0.5 * odtot * odtot;
                        double odtot_rec = rec_6 * odtot;
                        bbdtot = plfrac * (blay + dplankdn * odtot_rec);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + dplankup * odtot_rec);
                    } else if (odepth <= 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        odtot = odepth + LB(odcld, lev, ib);
                        double tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        double tfactot = G.tfn_tbl[ittot];
                        bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        atot[lev] = 1. - G.exp_tbl[ittot];
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    } else {
                        double tblind = odepth / (bpade + odepth);
                        int itgas = (int)(tblint * tblind + 0.5);
                        odepth = G.tau_tbl[itgas];
                        atrans[lev] = 1. - G.exp_tbl[itgas];
                        double tfacgas = G.tfn_tbl[itgas];
                        gassrc = atrans[lev] * plfrac * (blay + tfacgas * dplankdn);
                        odtot = odepth + LB(odcld, lev, ib);
                        tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        double tfactot = G.tfn_tbl[ittot];
                        bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + tfacgas * dplankdn);
                        atot[lev] = 1. - G.exp_tbl[ittot];
                        bbugas[lev] = plfrac * (blay + tfacgas * dplankup);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    }
                    if (!maxrand) {  // rtrn.f90:380-384 and twins
                        radld = radld - radld * (atrans[lev] + LB(efclfrac, lev, ib) * (1. - atrans[lev])) + gassrc +
                                CF(lev) * (bbdtot * atot[lev] - gassrc);
                        drad[lev - 1] = drad[lev - 1] + radld;
                    } else {  // rtrnmr.f90:596-621
                        if (istcldd[lev] == 1) {
                            cldradd = CF(lev) * radld;
                            clrradd = radld - cldradd;
                            oldcld = cldradd;
                            oldclr = clrradd;
                            rad = 0.;
                        }
                        double ttot = 1. - atot[lev];
                        double cldsrc = bbdtot * atot[lev];
                        cldradd = cldradd * ttot + CF(lev) * cldsrc;
                        clrradd = clrradd * (1. - atrans[lev]) + (1. - CF(lev)) * gassrc;
                        radld = cldradd + clrradd;
                        drad[lev - 1] = drad[lev - 1] + radld;
                        double radmod = rad * (facclr1d[lev - 1] * (1. - atrans[lev]) + faccld1d[lev - 1] * ttot) -
                                        faccmb1d[lev - 1] * gassrc + faccmb2d[lev - 1] * cldsrc;
                        oldcld = cldradd - radmod;
                        oldclr = clrradd + radmod;
                        rad = -radmod + facclr2d[lev - 1] * oldclr - faccld2d[lev - 1] * oldcld;
                        cldradd = cldradd + rad;
                        clrradd = clrradd - rad;
                    }
                } else {
                    if (odepth <= 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        odepth = rec_6 * odepth;
                        bbd = plfrac * (blay + dplankdn * odepth);
                        bbugas[lev] = plfrac * (blay + dplankup * odepth);
                    } else {
                        double tblind = odepth / (bpade + odepth);
                        int itr = (int)(tblint * tblind + 0.5);
                        double transc = G.exp_tbl[itr];
                        atrans[lev] = 1. - transc;
                        double tausfac = G.tfn_tbl[itr];
                        bbd = plfrac * (blay + tausfac * dplankdn);
                        bbugas[lev] = plfrac * (blay + tausfac * dplankup);
                    }
                    radld = radld + (bbd - radld) * atrans[lev];
                    drad[lev - 1] = drad[lev - 1] + radld;
                }
                if (iclddn == 1) {
                    radclrd = radclrd + (bbd - radclrd) * atrans[lev];
                    clrdrad[lev - 1] = clrdrad[lev - 1] + radclrd;
                } else {
                    radclrd = radld;
                    clrdrad[lev - 1] = drad[lev - 1];
                }
            }
            double rad0 = c.LG(c.fracs, 1, igc) * c.plankbnd[iband - 1];
            double d_rad0_dt = 0.;
            if (idrv == 1) d_rad0_dt = c.LG(c.fracs, 1, igc) * c.dplankbnd_dt[iband - 1];
            double reflect = 1. - c.semiss[iband - 1];
            double radlu = rad0 + reflect * radld;
            double radclru = rad0 + reflect * radclrd;
            urad[0] = urad[0] + radlu;
            clrurad[0] = clrurad[0] + radclru;
            double d_radlu_dt = 0., d_radclru_dt = 0.;
            if (idrv == 1) {
                d_radlu_dt = d_rad0_dt;
                d_urad_dt[0] = d_urad_dt[0] + d_radlu_dt;
                d_radclru_dt = d_rad0_dt;
                d_clrurad_dt[0] = d_clrurad_dt[0] + d_radclru_dt;
            }
            double cldradu = 0., clrradu = 0.;
            for (int lev = 1; lev <= nlayers; ++lev) {
                if (icldlyr[lev] == 1) {
                    double gassrc = bbugas[lev] * atrans[lev];
                    if (!maxrand) {  // rtrn.f90:499-510
                        radlu = radlu - radlu * (atrans[lev] + LB(efclfrac, lev, ib) * (1. - atrans[lev])) + gassrc +
                                CF(lev) * (bbutot[lev] * atot[lev] - gassrc);
                        urad[lev] = urad[lev] + radlu;
                    } else {  // rtrnmr.f90:694-718
                        if (istcld[lev] == 1) {
                            cldradu = CF(lev) * radlu;
                            clrradu = radlu - cldradu;
                            oldcld = cldradu;
                            oldclr = clrradu;
                            rad = 0.;
                        }
                        double ttot = 1. - atot[lev];
                        double cldsrc = bbutot[lev] * atot[lev];
                        cldradu = cldradu * ttot + CF(lev) * cldsrc;
                        clrradu = clrradu * (1.0 - atrans[lev]) + (1. - CF(lev)) * gassrc;
                        radlu = cldradu + clrradu;
                        urad[lev] = urad[lev] + radlu;
                        double radmod = rad * (facclr1[lev + 1] * (1.0 - atrans[lev]) + faccld1[lev + 1] * ttot) -
                                        faccmb1[lev + 1] * gassrc + faccmb2[lev + 1] * cldsrc;
                        oldcld = cldradu - radmod;
                        oldclr = clrradu + radmod;
                        rad = -radmod + facclr2[lev + 1] * oldclr - faccld2[lev + 1] * oldcld;
                        cldradu = cldradu + rad;
                        clrradu = clrradu - rad;
                    }
                    if (idrv == 1) {
                        d_radlu_dt = d_radlu_dt * CF(lev) * (1.0 - atot[lev]) + d_radlu_dt * (1.0 - CF(lev)) * (1.0 - atrans[lev]);
                        d_urad_dt[lev] = d_urad_dt[lev] + d_radlu_dt;
                    }
                } else {
                    radlu = radlu + (bbugas[lev] - radlu) * atrans[lev];
                    urad[lev] = urad[lev] + radlu;
                    if (idrv == 1) {
                        d_radlu_dt = d_radlu_dt * (1.0 - atrans[lev]);
                        d_urad_dt[lev] = d_urad_dt[lev] + d_radlu_dt;
                    }
                }
                if (iclddn == 1) {
                    radclru = radclru + (bbugas[lev] - radclru) * atrans[lev];
                    clrurad[lev] = clrurad[lev] + radclru;
                } else {
                    radclru = radlu;
                    clrurad[lev] = urad[lev];
                }
                if (idrv == 1) {
                    if (iclddn == 1) {
                        d_radclru_dt = d_radclru_dt * (1.0 - atrans[lev]);
                        d_clrurad_dt[lev] = d_clrurad_dt[lev] + d_radclru_dt;
                    } else {
                        d_radclru_dt = d_radlu_dt;
                        d_clrurad_dt[lev] = d_urad_dt[lev];
                    }
                }
            }
            igc = igc + 1;
        } while (igc <= G.ngs[iband - 1]);

        for (int lev = nlayers; lev >= 0; --lev) {
            double uflux = urad[lev] * wtdiff;
            double dflux = drad[lev] * wtdiff;
            urad[lev] = 0.0; drad[lev] = 0.0;
            c.totuflux[lev] = c.totuflux[lev] + uflux * G.delwave[iband - 1];
            c.totdflux[lev] = c.totdflux[lev] + dflux * G.delwave[iband - 1];
            double uclfl = clrurad[lev] * wtdiff;
            double dclfl = clrdrad[lev] * wtdiff;
            clrurad[lev] = 0.0; clrdrad[lev] = 0.0;
            c.totuclfl[lev] = c.totuclfl[lev] + uclfl * G.delwave[iband - 1];
            c.totdclfl[lev] = c.totdclfl[lev] + dclfl * G.delwave[iband - 1];
        }
        if (idrv == 1) {
            for (int lev = nlayers; lev >= 0; --lev) {
                double duflux_dt = d_urad_dt[lev] * wtdiff;
                d_urad_dt[lev] = 0.0;
                c.dtotuflux_dt[lev] = c.dtotuflux_dt[lev] + duflux_dt * G.delwave[iband - 1] * G.fluxfac;
                double duclfl_dt = d_clrurad_dt[lev] * wtdiff;
                d_clrurad_dt[lev] = 0.0;
                c.dtotuclfl_dt[lev] = c.dtotuclfl_dt[lev] + duclfl_dt * G.delwave[iband - 1] * G.fluxfac;
            }
        }
    }

    c.totuflux[0] = c.totuflux[0] * G.fluxfac;
    c.totdflux[0] = c.totdflux[0] * G.fluxfac;
    c.fnet[0] = c.totuflux[0] - c.totdflux[0];
    c.totuclfl[0] = c.totuclfl[0] * G.fluxfac;
    c.totdclfl[0] = c.totdclfl[0] * G.fluxfac;
    c.fnetc[0] = c.totuclfl[0] - c.totdclfl[0];
    for (int lev = 1; lev <= nlayers; ++lev) {
        c.totuflux[lev] = c.totuflux[lev] * G.fluxfac;
        c.totdflux[lev] = c.totdflux[lev] * G.fluxfac;
        c.fnet[lev] = c.totuflux[lev] - c.totdflux[lev];
        c.totuclfl[lev] = c.totuclfl[lev] * G.fluxfac;
        c.totdclfl[lev] = c.totdclfl[lev] * G.fluxfac;
        c.fnetc[lev] = c.totuclfl[lev] - c.totdclfl[lev];
        int l = lev - 1;
        c.htr[l] = G.heatfac * (c.fnet[l] - c.fnet[lev]) / (c.pz[l] - c.pz[lev]);
        c.htrc[l] = G.heatfac * (c.fnetc[l] - c.fnetc[lev]) / (c.pz[l] - c.pz[lev]);
    }
    c.htr[nlayers] = 0.0;
    c.htrc[nlayers] = 0.0;
}

}  // namespace orc
