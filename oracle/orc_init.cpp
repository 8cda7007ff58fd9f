// orc_init.cpp -- ORACLE (test infrastructure): restatement of rrtmg_lw_ini and callees.
// Follows src/rrtmg_lw_init.f90:47-194 (rrtmg_lw_ini), :197-300 (lwdatinit),
// :303-382 (lwcmbdat), :385-2034 (cmbgb1..16).  The constant tables of lwatmref /
// lwavplank / lwavplankderiv (src/rrtmg_lw_setcoef.f90:437-2009) and lwcldpr
// (src/rrtmg_lw_init.f90:2037-2675) arrive as a numeric blob written by
// tools/extract_static_tables.py; the 16-g k-data arrive in the layout of
// include/rrtmg_lw_kdata_layout.h (the reference's lw_kgbNN sources are missing).
#include "orc_common.h"
#include "../include/rrtmg_lw_kdata_layout.h"

namespace orc {

State G;
#ifdef ORC_COUNT_OPS
thread_local OpCount g_ops = {0, 0, 0, 0, 0, 0};
#endif

void Column::resize(int nl) {
    nlayers = nl;
    auto L = [&](std::vector<double>& v) { v.assign(nl, 0.0); };
    L(pavel); L(tavel); L(coldry); L(wbrodl); L(reicmc); L(relqmc);
    pz.assign(nl + 1, 0.0); tz.assign(nl + 1, 0.0);
    wkl.assign(7 * nl, 0.0); wx.assign(4 * nl, 0.0); taua.assign(nl * 16, 0.0);
    cldfmc.assign(NGPT * nl, 0.0); taucmc.assign(NGPT * nl, 0.0);
    ciwpmc.assign(NGPT * nl, 0.0); clwpmc.assign(NGPT * nl, 0.0);
    jp.assign(nl, 0); jt.assign(nl, 0); jt1.assign(nl, 0);
    indself.assign(nl, 0); indfor.assign(nl, 0); indminor.assign(nl, 0);
    planklay.assign(nl * 16, 0.0); planklev.assign((nl + 1) * 16, 0.0);
    L(colh2o); L(colco2); L(colo3); L(coln2o); L(colco); L(colch4); L(colo2); L(colbrd);
    L(fac00); L(fac01); L(fac10); L(fac11);
    L(rat_h2oco2); L(rat_h2oco2_1); L(rat_h2oo3); L(rat_h2oo3_1); L(rat_h2on2o); L(rat_h2on2o_1);
    L(rat_h2och4); L(rat_h2och4_1); L(rat_n2oco2); L(rat_n2oco2_1); L(rat_o3co2); L(rat_o3co2_1);
    L(selffac); L(selffrac); L(forfac); L(forfrac); L(minorfrac); L(scaleminor); L(scaleminorn2);
    fracs.assign(nl * NGPT, 0.0); taug.assign(nl * NGPT, 0.0); taut.assign(nl * NGPT, 0.0);
    auto V = [&](std::vector<double>& v) { v.assign(nl + 2, 0.0); };
    V(totuflux); V(totdflux); V(fnet); V(htr); V(totuclfl); V(totdclfl); V(fnetc); V(htrc);
    V(dtotuflux_dt); V(dtotuclfl_dt);
}

// src/rrtmg_lw_init.f90:197-300
static void lwdatinit(double cpdair) {
    static const double dw[16] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
    static const int nspa[16] = {1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9};
    static const int nspb[16] = {1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0};
    for (int i = 0; i < 16; ++i) { G.delwave[i] = dw[i]; G.ng[i] = 16; G.nspa[i] = nspa[i]; G.nspb[i] = nspb[i]; }
    G.grav = 9.8066;
    G.planck = 6.62606876e-27;
    G.boltz = 1.3806503e-16;
    G.clight = 2.99792458e+10;
    G.avogad = 6.02214199e+23;
    G.alosmt = 2.6867775e+19;
    G.gascon = 8.31447200e+07;
    G.radcn1 = 1.191042722e-12;
    G.radcn2 = 1.4387752;
    G.sbcnst = 5.670400e-04;
    G.secdy = 8.6400e4;
    G.heatfac = G.grav * G.secdy / (cpdair * 1.e2);  // :298
}

// src/rrtmg_lw_init.f90:303-382
static void lwcmbdat() {
    static const int ngc[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
    static const int ngs[16] = {10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140};
    static const int ngm[256] = {
        1,2,3,3,4,4,5,5,6,6,7,7,8,8,9,10,       1,2,3,4,5,6,7,8,9,9,10,10,11,11,12,12,
        1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16, 1,2,3,4,5,6,7,8,9,10,11,12,13,14,14,14,
        1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16, 1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,
        1,1,2,2,3,4,5,6,7,8,9,10,11,11,12,12,   1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,
        1,2,3,4,5,6,7,8,9,9,10,10,11,11,12,12,  1,1,2,2,3,3,4,4,5,5,5,5,6,6,6,6,
        1,2,3,3,4,4,5,5,6,6,7,7,7,8,8,8,        1,2,3,4,5,5,6,6,7,7,7,7,8,8,8,8,
        1,1,1,2,2,2,3,3,3,3,4,4,4,4,4,4,        1,1,1,1,1,1,1,1,2,2,2,2,2,2,2,2,
        1,1,1,1,1,1,1,1,2,2,2,2,2,2,2,2,        1,1,1,1,2,2,2,2,2,2,2,2,2,2,2,2};
    static const int ngn[140] = {
        1,1,2,2,2,2,2,2,1,1,  1,1,1,1,1,1,1,1,2,2,2,2,  1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
        1,1,1,1,1,1,1,1,1,1,1,1,1,3,  1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,  2,2,2,2,2,2,2,2,
        2,2,1,1,1,1,1,1,1,1,2,2,  2,2,2,2,2,2,2,2,  1,1,1,1,1,1,1,1,2,2,2,2,  2,2,2,2,4,4,
        1,1,2,2,2,2,3,3,  1,1,1,1,2,2,4,4,  3,3,4,6,  8,8,  8,8,  4,12};
    static const double wt[16] = {0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
                                  0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                                  0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};
    for (int i = 0; i < 16; ++i) { G.ngc[i] = ngc[i]; G.ngs[i] = ngs[i]; G.wt[i] = wt[i]; }
    for (int i = 0; i < 256; ++i) G.ngm[i] = ngm[i];
    for (int i = 0; i < 140; ++i) G.ngn[i] = ngn[i];
    int k = 0;
    for (int b = 0; b < 16; ++b)
        for (int i = 0; i < ngc[b]; ++i) G.ngb[k++] = b + 1;  // ngb(:) :362-377
}

static std::string reduced_name(const std::string& n) {
    if (n.rfind("kao", 0) == 0) return n.size() == 3 ? "absa" : "ka" + n.substr(3);
    if (n.rfind("kbo", 0) == 0) return n.size() == 3 ? "absb" : "kb" + n.substr(3);
    return n.substr(0, n.size() - 1);  // selfrefo->selfref, fracrefao->fracrefa, ccl4o->ccl4, ...
}

// One array of one band through cmbgbN: reduced(:,igc) = sum over the ngn(ngs(b-1)+igc)
// original g-points of orig(:,iprsm) [* rwgt(iprsm + 16*(b-1)) unless a Planck fraction].
// src/rrtmg_lw_init.f90:385-2034 -- every cmbgbN has this one loop shape (sums start
// at 0. and run in iprsm order).
static void cmbgb_array(int b, const rrtmg_lw_karray_desc& d, const double* orig) {
    Band& B = G.band[b - 1];
    const int ngc = G.ngc[b - 1];
    const int goff = (b == 1) ? 0 : G.ngs[b - 2];
    std::vector<double> red((size_t)d.lead * ngc, 0.0);
    for (int i = 0; i < d.lead; ++i) {
        int iprsm = 0;
        for (int igc = 1; igc <= ngc; ++igc) {
            double sum = 0.;
            for (int ipr = 1; ipr <= G.ngn[goff + igc - 1]; ++ipr) {
                iprsm++;
                if (d.gfirst) {
                    sum = sum + orig[(iprsm - 1) + 16 * i];
                } else if (d.weighted) {
                    sum = sum + orig[i + (size_t)d.lead * (iprsm - 1)] * G.rwgt[iprsm - 1 + 16 * (b - 1)];
                } else {
                    sum = sum + orig[i + (size_t)d.lead * (iprsm - 1)];
                }
            }
            if (d.gfirst) red[(igc - 1) + (size_t)ngc * i] = sum;
            else red[i + (size_t)d.lead * (igc - 1)] = sum;
        }
    }
    std::string rn = reduced_name(d.name);
    B.arr[rn] = red;
    B.lead[rn] = d.lead;
}

// src/rrtmg_lw_init.f90:47-194
int init(double cpdair, const double* st, const double* kdata, int drk) {
    G = State();
    G.drk = drk;
    lwdatinit(cpdair);
    lwcmbdat();
    // static tables blob, order of tools/extract_static_tables.py STATIC_LAYOUT
    const double* p = st;
    auto take = [&](double* dst, int n) { std::memcpy(dst, p, sizeof(double) * n); p += n; };
    take(G.pref, 59); take(G.preflog, 59); take(G.tref, 59); take(G.chi_mls, 7 * 59);
    take(G.totplnk, 181 * 16); take(G.totplk16, 181); take(G.totplnkderiv, 181 * 16); take(G.totplk16deriv, 181);
    take(&G.abscld1, 1); take(G.absice0, 2); take(G.absice1, 10); take(G.absice2, 43 * 16); take(G.absice3, 46 * 16);
    take(&G.absliq0, 1); take(G.absliq1, 58 * 16);

    // lookup tables :125-142
    const double expeps = 1.e-20;
    const double pade = 0.278;
    G.tau_tbl[0] = 0.0; G.tau_tbl[NTBL] = 1.e10;
    G.exp_tbl[0] = 1.0; G.exp_tbl[NTBL] = expeps;
    G.tfn_tbl[0] = 0.0; G.tfn_tbl[NTBL] = 1.0;
    G.bpade = 1.0 / pade;
    for (int itr = 1; itr <= NTBL - 1; ++itr) {
        // tfn = real(itr) / real(ntbl): default-kind division
        double tfn = (drk == 4) ? (double)((float)itr / (float)NTBL) : (double)itr / (double)NTBL;
        G.tau_tbl[itr] = G.bpade * tfn / (1. - tfn);
        G.exp_tbl[itr] = std::exp(-G.tau_tbl[itr]);
        if (G.exp_tbl[itr] <= expeps) G.exp_tbl[itr] = expeps;
        if (G.tau_tbl[itr] < 0.06) {
            G.tfn_tbl[itr] = G.tau_tbl[itr] / 6.;
        } else {
            G.tfn_tbl[itr] = 1. - 2. * ((1. / G.tau_tbl[itr]) - (G.exp_tbl[itr] / (1. - G.exp_tbl[itr])));
        }
    }

    // rwgt :149-173
    int igcsm = 0;
    for (int ibnd = 1; ibnd <= NBND; ++ibnd) {
        int iprsm = 0;
        double wtsm[MG];
        if (G.ngc[ibnd - 1] < MG) {
            for (int igc = 1; igc <= G.ngc[ibnd - 1]; ++igc) {
                igcsm++;
                double wtsum = 0.;
                for (int ipr = 1; ipr <= G.ngn[igcsm - 1]; ++ipr) {
                    iprsm++;
                    wtsum = wtsum + G.wt[iprsm - 1];
                }
                wtsm[igc - 1] = wtsum;
            }
            for (int ig = 1; ig <= G.ng[ibnd - 1]; ++ig) {
                int ind = (ibnd - 1) * MG + ig;
                G.rwgt[ind - 1] = G.wt[ig - 1] / wtsm[G.ngm[ind - 1] - 1];
            }
        } else {
            for (int ig = 1; ig <= G.ng[ibnd - 1]; ++ig) {
                igcsm++;
                int ind = (ibnd - 1) * MG + ig;
                G.rwgt[ind - 1] = 1.0;
            }
        }
    }

    // cmbgb1..16
    const double* k = kdata;
    for (int i = 0; i < RRTMG_LW_NKARRAYS; ++i) {
        const rrtmg_lw_karray_desc& d = RRTMG_LW_KARRAYS[i];
        G.band[d.band - 1].ng = G.ngc[d.band - 1];
        cmbgb_array(d.band, d, k);
        k += (size_t)d.lead * 16;
    }
    // rrtmg_lw sets these on every call (src/rrtmg_lw_rad.f90:451-453)
    G.oneminus = 1. - 1.e-6;
    G.pi = 2. * std::asin(1.);
    G.fluxfac = G.pi * 2.e4;
    G.inited = true;
    return ORC_OK;
}

}  // namespace orc
