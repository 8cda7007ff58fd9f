// orc_rtrnmc.cpp -- ORACLE (test infrastructure): restatement of rtrnmc.
// Follows src/rrtmg_lw_rtrnmc.f90:251-593 statement by statement (same branch
// structure, same association order in every expression).
#include "orc_common.h"

namespace orc {

void rtrnmc(Column& c, int istart, int iend, int iout, int idrv, bool dump_itr) {
    const int nlayers = c.nlayers;
    const double wtdiff = 0.5, rec_6 = 0.166667;  // :251-252 (T3)
    static const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33,
                                  1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    static const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40,
                                  -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062,
                                  0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
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
    std::vector<double> urad(n1, 0.), drad(n1, 0.), clrurad(n1, 0.), clrdrad(n1, 0.), d_urad_dt(n1, 0.),
        d_clrurad_dt(n1, 0.);
    std::vector<double> atrans(n1, 0.), bbugas(n1, 0.), atot(n1, 0.), bbutot(n1, 0.);
    std::vector<int> icldlyr(n1, 0);
    std::vector<double> odcld((size_t)nlayers * NGPT, 0.), abscld((size_t)nlayers * NGPT, 0.),
        efclfrac((size_t)nlayers * NGPT, 0.);
    for (int l = 0; l <= nlayers; ++l) {
        c.totuflux[l] = c.totdflux[l] = c.totuclfl[l] = c.totdclfl[l] = 0.;
        c.dtotuflux_dt[l] = c.dtotuclfl_dt[l] = 0.;
    }
    if (dump_itr) c.itr_dump.assign((size_t)nlayers * NGPT * 2, -1);

    for (int lay = 1; lay <= nlayers; ++lay) {
        for (int ig = 1; ig <= NGPT; ++ig) {
            if (c.CLD(c.cldfmc, ig, lay) == 1.) {
                int ib = G.ngb[ig - 1];
                c.LG(odcld, lay, ig) = secdiff[ib - 1] * c.CLD(c.taucmc, ig, lay);
                double transcld = std::exp(-c.LG(odcld, lay, ig));
                c.LG(abscld, lay, ig) = 1. - transcld;
                c.LG(efclfrac, lay, ig) = c.LG(abscld, lay, ig) * c.CLD(c.cldfmc, ig, lay);
                icldlyr[lay] = 1;
            } else {
                c.LG(odcld, lay, ig) = 0.0;
                c.LG(abscld, lay, ig) = 0.0;
                c.LG(efclfrac, lay, ig) = 0.0;
            }
        }
    }

    int igc = 1;
    for (int iband = istart; iband <= iend; ++iband) {
        if (iout > 0 && iband >= 2) igc = G.ngs[iband - 2] + 1;
        do {  // 1000 continue ... go to 1000
            double radld = 0., radclrd = 0.;
            int iclddn = 0;
            double bbd = 0.;
            // downward loop :350-455
            for (int lev = nlayers; lev >= 1; --lev) {
                double plfrac = c.LG(c.fracs, lev, igc);
                double blay = c.PLAY(lev, iband);
                double dplankup = c.PLEV(lev, iband) - blay;
                double dplankdn = c.PLEV(lev - 1, iband) - blay;
                double odepth = secdiff[iband - 1] * c.LG(c.taut, lev, igc);
                if (odepth < 0.0) odepth = 0.0;
                int d0 = -1, d1 = -1;
                if (icldlyr[lev] == 1) {
                    iclddn = 1;
                    double odtot = odepth + c.LG(odcld, lev, igc);
                    double cf = c.CLD(c.cldfmc, igc, lev), ef = c.LG(efclfrac, lev, igc);
                    if (odtot < 0.06) {
                        branch_hit(138);
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        double gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        atot[lev] = odtot - 0.5 * odtot * odtot;
                        double odtot_rec = rec_6 * odtot;
                        double bbdtot = plfrac * (blay + dplankdn * odtot_rec);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        radld = radld - radld * (atrans[lev] + ef * (1. - atrans[lev])) + gassrc +
                                cf * (bbdtot * atot[lev] - gassrc);
                        drad[lev - 1] = drad[lev - 1] + radld;
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + dplankup * odtot_rec);
                    } else if (odepth <= 0.06) {
                        branch_hit(139);
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        double gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        odtot = odepth + c.LG(odcld, lev, igc);
                        double tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        d1 = ittot;
                        double tfactot = G.tfn_tbl[ittot];
                        double bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        atot[lev] = 1. - G.exp_tbl[ittot];
                        radld = radld - radld * (atrans[lev] + ef * (1. - atrans[lev])) + gassrc +
                                cf * (bbdtot * atot[lev] - gassrc);
                        drad[lev - 1] = drad[lev - 1] + radld;
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    } else {
                        branch_hit(140);
                        double tblind = odepth / (bpade + odepth);
                        int itgas = (int)(tblint * tblind + 0.5);
                        d0 = itgas;
                        odepth = G.tau_tbl[itgas];
                        atrans[lev] = 1. - G.exp_tbl[itgas];
                        double tfacgas = G.tfn_tbl[itgas];
                        double gassrc = atrans[lev] * plfrac * (blay + tfacgas * dplankdn);
                        odtot = odepth + c.LG(odcld, lev, igc);
                        tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        d1 = ittot;
                        double tfactot = G.tfn_tbl[ittot];
                        double bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + tfacgas * dplankdn);
                        atot[lev] = 1. - G.exp_tbl[ittot];
                        radld = radld - radld * (atrans[lev] + ef * (1. - atrans[lev])) + gassrc +
                                cf * (bbdtot * atot[lev] - gassrc);
                        drad[lev - 1] = drad[lev - 1] + radld;
                        bbugas[lev] = plfrac * (blay + tfacgas * dplankup);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    }
                } else {
                    if (odepth <= 0.06) {
                        branch_hit(141);
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        odepth = rec_6 * odepth;
                        bbd = plfrac * (blay + dplankdn * odepth);
                        bbugas[lev] = plfrac * (blay + dplankup * odepth);
                    } else {
                        branch_hit(142);
                        double tblind = odepth / (bpade + odepth);
                        int itr = (int)(tblint * tblind + 0.5);
                        d0 = itr;
                        double transc = G.exp_tbl[itr];
                        atrans[lev] = 1. - transc;
                        double tausfac = G.tfn_tbl[itr];
                        bbd = plfrac * (blay + tausfac * dplankdn);
                        bbugas[lev] = plfrac * (blay + tausfac * dplankup);
                    }
                    radld = radld + (bbd - radld) * atrans[lev];
                    drad[lev - 1] = drad[lev - 1] + radld;
                }
                if (dump_itr) {
                    c.itr_dump[2 * ((size_t)(lev - 1) + (size_t)nlayers * (igc - 1))] = d0;
                    c.itr_dump[2 * ((size_t)(lev - 1) + (size_t)nlayers * (igc - 1)) + 1] = d1;
                }
                if (iclddn == 1) {
                    radclrd = radclrd + (bbd - radclrd) * atrans[lev];
                    clrdrad[lev - 1] = clrdrad[lev - 1] + radclrd;
                } else {
                    radclrd = radld;
                    clrdrad[lev - 1] = drad[lev - 1];
                }
            }
            // surface :465-473
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
            // upward loop :486-529
            for (int lev = 1; lev <= nlayers; ++lev) {
                if (icldlyr[lev] == 1) {
                    double cf = c.CLD(c.cldfmc, igc, lev), ef = c.LG(efclfrac, lev, igc);
                    double gassrc = bbugas[lev] * atrans[lev];
                    radlu = radlu - radlu * (atrans[lev] + ef * (1. - atrans[lev])) + gassrc +
                            cf * (bbutot[lev] * atot[lev] - gassrc);
                    urad[lev] = urad[lev] + radlu;
                    if (idrv == 1) {
                        d_radlu_dt = d_radlu_dt * cf * (1.0 - atot[lev]) + d_radlu_dt * (1.0 - cf) * (1.0 - atrans[lev]);
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

        // band close-out :538-563
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

    // :569-593
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
