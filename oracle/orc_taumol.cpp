// orc_taumol.cpp -- ORACLE (test infrastructure): restatement of taumol / taugb1..16.
// Follows src/rrtmg_lw_taumol.f90:299-3164.  One function per reference band routine,
// same statement order and the same left-to-right association inside every sum, so
// that a build with -ffp-contract=off reproduces the reference's rounding.
// The eta-interpolation weight block and the 6/4-term tau_major sums are token-identical
// in bands 3,4,5,7,9,12,13,15,16 of the reference (checked mechanically); they are
// written once here as eta_fac()/tau_major().
#include "orc_common.h"

namespace orc {
namespace {

struct Ctx {
    Column& c;
    const Band& B;
    int b;      // band 1..16
    int goff;   // ngs(b-1)
    int ng;
    const double *absa, *absb, *selfref, *forref, *fracrefa, *fracrefb;
    int nra, nrb;  // rows of absa / absb
    Ctx(Column& c_, int band) : c(c_), B(G.band[band - 1]), b(band) {
        g_cur_band = band;
        branch_hit(96 + (band - 1) * 2, c_.laytrop);
        branch_hit(96 + (band - 1) * 2 + 1, c_.nlayers - c_.laytrop);
        goff = band == 1 ? 0 : G.ngs[band - 2];
        ng = G.ngc[band - 1];
        absa = B.a("absa"); absb = B.a("absb"); selfref = B.a("selfref"); forref = B.a("forref");
        fracrefa = B.a("fracrefa"); fracrefb = B.a("fracrefb");
        nra = 65 * G.nspa[band - 1];
        nrb = absb ? B.lead.at("absb") : 0;
    }
    // 1-based Fortran element access
    double A(int row, int ig) const { return absa[(row - 1) + (size_t)nra * (ig - 1)]; }
    double Bk(int row, int ig) const { return absb[(row - 1) + (size_t)nrb * (ig - 1)]; }
    double S(int i, int ig) const { return selfref[(i - 1) + 10 * (ig - 1)]; }
    double F(int i, int ig) const { return forref[(i - 1) + 4 * (ig - 1)]; }
    double FA1(int ig) const { return fracrefa[ig - 1]; }
    double FB1(int ig) const { return fracrefb[ig - 1]; }
    double FA(int ig, int j) const { return fracrefa[(ig - 1) + ng * (j - 1)]; }
    double FB(int ig, int j) const { return fracrefb[(ig - 1) + ng * (j - 1)]; }
    double& taug(int lay, int ig) { return c.LG(c.taug, lay, goff + ig); }
    double& fracs(int lay, int ig) { return c.LG(c.fracs, lay, goff + ig); }
};

// minor-gas tables: M2 = (19, ng), M3 = (neta, 19, ng)
inline double M2(const double* t, int i, int ig) { return t[(i - 1) + 19 * (ig - 1)]; }
inline double M3(const double* t, int neta, int j, int i, int ig) {
    return t[(j - 1) + neta * ((i - 1) + 19 * (ig - 1))];
}

// speccomb/specparm/specmult/js/fs block, e.g. taumol.f90:523-528
struct Spec { double comb, parm, fs; int js; };
inline Spec spec(double colA, double rat, double colB, double mult) {
    Spec s;
    s.comb = colA + rat * colB;
    s.parm = colA / s.comb;
    if (s.parm >= G.oneminus) s.parm = G.oneminus;
    double specmult = mult * (s.parm);
    s.js = 1 + (int)specmult;
    s.fs = std::fmod(specmult, 1.0);
    return s;
}

// eta weights for one (p,T) corner pair, taumol.f90:569-598 (and its twins)
struct Fac { int br; double f0, f1, f2, g0, g1, g2; };  // f* multiply row block 0, g* row block +nsp
inline Fac eta_fac(double specparm, double fs, double facA, double facB, int corner) {
    Fac w;
    const int site = ((g_cur_band - 1) * 2 + corner) * 3;
    if (specparm < 0.125) {
        w.br = 0;
        branch_hit(site);
        double p = fs - 1;
        double p2 = p * p;
        double p4 = p2 * p2;  // p**4
        double fk0 = p4;
        double fk1 = 1 - p - 2.0 * p4;
        double fk2 = p + p4;
        w.f0 = fk0 * facA; w.f1 = fk1 * facA; w.f2 = fk2 * facA;
        w.g0 = fk0 * facB; w.g1 = fk1 * facB; w.g2 = fk2 * facB;
    } else if (specparm > 0.875) {
        w.br = 1;
        branch_hit(site + 1);
        double p = -fs;
        double p2 = p * p;
        double p4 = p2 * p2;
        double fk0 = p4;
        double fk1 = 1 - p - 2.0 * p4;
        double fk2 = p + p4;
        w.f0 = fk0 * facA; w.f1 = fk1 * facA; w.f2 = fk2 * facA;
        w.g0 = fk0 * facB; w.g1 = fk1 * facB; w.g2 = fk2 * facB;
    } else {
        w.br = 2;
        branch_hit(site + 2);
        w.f0 = (1. - fs) * facA; w.g0 = (1. - fs) * facB;
        w.f1 = fs * facA; w.g1 = fs * facB;
        w.f2 = w.g2 = 0.;
    }
    return w;
}
// tau_major / tau_major1, taumol.f90:641-687 (nspa = 9 rows per T step)
inline double tau_major(const Ctx& x, const Fac& w, double speccomb, int ind, int ig) {
    if (w.br == 0)
        return speccomb * (w.f0 * x.A(ind, ig) + w.f1 * x.A(ind + 1, ig) + w.f2 * x.A(ind + 2, ig) +
                           w.g0 * x.A(ind + 9, ig) + w.g1 * x.A(ind + 10, ig) + w.g2 * x.A(ind + 11, ig));
    if (w.br == 1)
        return speccomb * (w.f2 * x.A(ind - 1, ig) + w.f1 * x.A(ind, ig) + w.f0 * x.A(ind + 1, ig) +
                           w.g2 * x.A(ind + 8, ig) + w.g1 * x.A(ind + 9, ig) + w.g0 * x.A(ind + 10, ig));
    return speccomb * (w.f0 * x.A(ind, ig) + w.f1 * x.A(ind + 1, ig) + w.g0 * x.A(ind + 9, ig) + w.g1 * x.A(ind + 10, ig));
}

#define LAYVARS                                                                              \
    const double fac00 = c.fac00[lay - 1], fac01 = c.fac01[lay - 1], fac10 = c.fac10[lay - 1], \
                 fac11 = c.fac11[lay - 1];                                                   \
    const int jp = c.jp[lay - 1], jt = c.jt[lay - 1], jt1 = c.jt1[lay - 1];                   \
    (void)fac00; (void)fac01; (void)fac10; (void)fac11; (void)jp; (void)jt; (void)jt1;

inline double tauself_(const Ctx& x, const Column& c, int lay, int ig) {
    int inds = c.indself[lay - 1];
    return c.selffac[lay - 1] * (x.S(inds, ig) + c.selffrac[lay - 1] * (x.S(inds + 1, ig) - x.S(inds, ig)));
}
inline double taufor_(const Ctx& x, const Column& c, int lay, int ig) {
    int indf = c.indfor[lay - 1];
    return c.forfac[lay - 1] * (x.F(indf, ig) + c.forfrac[lay - 1] * (x.F(indf + 1, ig) - x.F(indf, ig)));
}
inline double minor2(const double* t, const Column& c, int lay, int ig) {
    int indm = c.indminor[lay - 1];
    return M2(t, indm, ig) + c.minorfrac[lay - 1] * (M2(t, indm + 1, ig) - M2(t, indm, ig));
}
inline double minor3(const double* t, int neta, int jm, double fm, const Column& c, int lay, int ig) {
    int indm = c.indminor[lay - 1];
    double m1 = M3(t, neta, jm, indm, ig) + fm * (M3(t, neta, jm + 1, indm, ig) - M3(t, neta, jm, indm, ig));
    double m2 = M3(t, neta, jm, indm + 1, ig) + fm * (M3(t, neta, jm + 1, indm + 1, ig) - M3(t, neta, jm, indm + 1, ig));
    return m1 + c.minorfrac[lay - 1] * (m2 - m1);
}
// single-key-species index pairs, e.g. taumol.f90:337-338 / :369-370
inline void ind_lo(int jp, int jt, int jt1, int nspa, int js, int js1, int& ind0, int& ind1) {
    ind0 = ((jp - 1) * 5 + (jt - 1)) * nspa + js;
    ind1 = (jp * 5 + (jt1 - 1)) * nspa + js1;
}
inline void ind_up(int jp, int jt, int jt1, int nspb, int js, int js1, int& ind0, int& ind1) {
    ind0 = ((jp - 13) * 5 + (jt - 1)) * nspb + js;
    ind1 = ((jp - 12) * 5 + (jt1 - 1)) * nspb + js1;
}
inline double key4a(const Ctx& x, int ind0, int ind1, double f00, double f10, double f01, double f11, int ig) {
    return (f00 * x.A(ind0, ig) + f10 * x.A(ind0 + 1, ig) + f01 * x.A(ind1, ig) + f11 * x.A(ind1 + 1, ig));
}
inline double key4b(const Ctx& x, int ind0, int ind1, double f00, double f10, double f01, double f11, int ig) {
    return (f00 * x.Bk(ind0, ig) + f10 * x.Bk(ind0 + 1, ig) + f01 * x.Bk(ind1, ig) + f11 * x.Bk(ind1 + 1, ig));
}
// upper-atmosphere binary-species sum (nspb = 5), e.g. taumol.f90:762-771
inline double key8b(const Ctx& x, const Spec& s, const Spec& s1, int ind0, int ind1, double fac00, double fac10,
                    double fac01, double fac11, int ig) {
    double fac000 = (1. - s.fs) * fac00, fac010 = (1. - s.fs) * fac10, fac100 = s.fs * fac00, fac110 = s.fs * fac10;
    double fac001 = (1. - s1.fs) * fac01, fac011 = (1. - s1.fs) * fac11, fac101 = s1.fs * fac01, fac111 = s1.fs * fac11;
    return s.comb * (fac000 * x.Bk(ind0, ig) + fac100 * x.Bk(ind0 + 1, ig) + fac010 * x.Bk(ind0 + 5, ig) +
                     fac110 * x.Bk(ind0 + 6, ig)) +
           s1.comb * (fac001 * x.Bk(ind1, ig) + fac101 * x.Bk(ind1 + 1, ig) + fac011 * x.Bk(ind1 + 5, ig) +
                      fac111 * x.Bk(ind1 + 6, ig));
}
// gas-amount adjustment with pow, e.g. taumol.f90:547-554
inline double adjcol(int site, double col, double coldry, double e20, double chiref, double thresh, double base,
                     double expo, double chiref2) {
    double chi = col / coldry;
    double rat = e20 * chi / chiref;
    if (rat > thresh) {
        branch_hit(128 + site);
        double adjfac = base + std::pow(rat - base, expo);
        return adjfac * chiref2 * coldry * 1.e-20;
    }
    return col;
}

// ---- band 1: 10-350 cm-1 (low key h2o; low minor n2) (high key h2o; high minor n2)  :304-389
void taugb1(Column& c) {
    Ctx x(c, 1);
    const double* ka_mn2 = x.B.a("ka_mn2");
    const double* kb_mn2 = x.B.a("kb_mn2");
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[0], 1, 1, ind0, ind1);
        double pp = c.pavel[lay - 1];
        double corradj = 1.;
        if (pp < 250.) corradj = 1. - 0.15 * (250. - pp) / 154.4;
        double scalen2 = c.colbrd[lay - 1] * c.scaleminorn2[lay - 1];
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double taun2 = scalen2 * (minor2(ka_mn2, c, lay, ig));
            x.taug(lay, ig) = corradj * (c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) +
                                         tauself + taufor + taun2);
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[0], 1, 1, ind0, ind1);
        double pp = c.pavel[lay - 1];
        double corradj = 1. - 0.15 * (pp / 95.6);
        double scalen2 = c.colbrd[lay - 1] * c.scaleminorn2[lay - 1];
        for (int ig = 1; ig <= x.ng; ++ig) {
            double taufor = taufor_(x, c, lay, ig);
            double taun2 = scalen2 * (minor2(kb_mn2, c, lay, ig));
            x.taug(lay, ig) =
                corradj * (c.colh2o[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + taufor + taun2);
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 2: 350-500 cm-1 (h2o)  :392-470
void taugb2(Column& c) {
    Ctx x(c, 2);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[1], 1, 1, ind0, ind1);
        double pp = c.pavel[lay - 1];
        double corradj = 1. - .05 * (pp - 100.) / 900.;
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            x.taug(lay, ig) =
                corradj * (c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself + taufor);
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[1], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double taufor = taufor_(x, c, lay, ig);
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + taufor;
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 3: 500-630 cm-1 (low key h2o,co2; low minor n2o) (high key h2o,co2; high minor n2o)  :473-780
void taugb3(Column& c) {
    Ctx x(c, 3);
    const double* ka_mn2o = x.B.a("ka_mn2o");
    const double* kb_mn2o = x.B.a("kb_mn2o");
    double refrat_planck_a = G.chi(1, 9) / G.chi(2, 9);
    double refrat_planck_b = G.chi(1, 13) / G.chi(2, 13);
    double refrat_m_a = G.chi(1, 3) / G.chi(2, 3);
    double refrat_m_b = G.chi(1, 13) / G.chi(2, 13);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oco2[lay - 1], colco2, 8.);
        Spec s1 = spec(colh2o, c.rat_h2oco2_1[lay - 1], colco2, 8.);
        Spec sm = spec(colh2o, refrat_m_a, colco2, 8.);
        double adjcoln2o = adjcol(0, c.coln2o[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(4, jp + 1), 1.5, 0.5, 0.65,
                                  G.chi(4, jp + 1));
        Spec sp = spec(colh2o, refrat_planck_a, colco2, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[2], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absn2o = minor3(ka_mn2o, 9, sm.js, sm.fs, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + adjcoln2o * absn2o;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oco2[lay - 1], colco2, 4.);
        Spec s1 = spec(colh2o, c.rat_h2oco2_1[lay - 1], colco2, 4.);
        Spec sm = spec(colh2o, refrat_m_b, colco2, 4.);
        // :734 "1.e20" is a default-kind literal here (T1)
        double adjcoln2o = adjcol(1, c.coln2o[lay - 1], c.coldry[lay - 1], lit(1.e20), G.chi(4, jp + 1), 1.5, 0.5, 0.65,
                                  G.chi(4, jp + 1));
        Spec sp = spec(colh2o, refrat_planck_b, colco2, 4.);
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[2], s.js, s1.js, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double taufor = taufor_(x, c, lay, ig);
            double absn2o = minor3(kb_mn2o, 5, sm.js, sm.fs, c, lay, ig);
            x.taug(lay, ig) = key8b(x, s, s1, ind0, ind1, fac00, fac10, fac01, fac11, ig) + taufor + adjcoln2o * absn2o;
            x.fracs(lay, ig) = x.FB(ig, sp.js) + sp.fs * (x.FB(ig, sp.js + 1) - x.FB(ig, sp.js));
        }
    }
}

// ---- band 4: 630-700 cm-1 (low key h2o,co2) (high key o3,co2)  :783-1037
void taugb4(Column& c) {
    Ctx x(c, 4);
    double refrat_planck_a = G.chi(1, 11) / G.chi(2, 11);
    double refrat_planck_b = G.chi(3, 13) / G.chi(2, 13);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oco2[lay - 1], colco2, 8.);
        Spec s1 = spec(colh2o, c.rat_h2oco2_1[lay - 1], colco2, 8.);
        Spec sp = spec(colh2o, refrat_planck_a, colco2, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[3], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    // :1028-1034 default-kind literals (T1)
    const double sc[7] = {lit(0.92), lit(0.88), lit(1.07), lit(1.1), lit(0.99), lit(0.88), lit(0.943)};
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        const double colo3 = c.colo3[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colo3, c.rat_o3co2[lay - 1], colco2, 4.);
        Spec s1 = spec(colo3, c.rat_o3co2_1[lay - 1], colco2, 4.);
        Spec sp = spec(colo3, refrat_planck_b, colco2, 4.);
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[3], s.js, s1.js, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = key8b(x, s, s1, ind0, ind1, fac00, fac10, fac01, fac11, ig);
            x.fracs(lay, ig) = x.FB(ig, sp.js) + sp.fs * (x.FB(ig, sp.js + 1) - x.FB(ig, sp.js));
        }
        for (int k = 0; k < 7; ++k) x.taug(lay, 8 + k) = x.taug(lay, 8 + k) * sc[k];
    }
}

// ---- band 5: 700-820 cm-1 (low key h2o,co2; low minor o3, ccl4) (high key o3,co2)  :1040-1316
void taugb5(Column& c) {
    Ctx x(c, 5);
    const double* ka_mo3 = x.B.a("ka_mo3");
    const double* ccl4 = x.B.a("ccl4");
    double refrat_planck_a = G.chi(1, 5) / G.chi(2, 5);
    double refrat_planck_b = G.chi(3, 43) / G.chi(2, 43);
    double refrat_m_a = G.chi(1, 7) / G.chi(2, 7);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oco2[lay - 1], colco2, 8.);
        Spec s1 = spec(colh2o, c.rat_h2oco2_1[lay - 1], colco2, 8.);
        Spec sm = spec(colh2o, refrat_m_a, colco2, 8.);
        Spec sp = spec(colh2o, refrat_planck_a, colco2, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[4], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double abso3 = minor3(ka_mo3, 9, sm.js, sm.fs, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + abso3 * c.colo3[lay - 1] + c.WX(1, lay) * ccl4[ig - 1];
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        const double colo3 = c.colo3[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colo3, c.rat_o3co2[lay - 1], colco2, 4.);
        Spec s1 = spec(colo3, c.rat_o3co2_1[lay - 1], colco2, 4.);
        Spec sp = spec(colo3, refrat_planck_b, colco2, 4.);
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[4], s.js, s1.js, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = key8b(x, s, s1, ind0, ind1, fac00, fac10, fac01, fac11, ig) + c.WX(1, lay) * ccl4[ig - 1];
            x.fracs(lay, ig) = x.FB(ig, sp.js) + sp.fs * (x.FB(ig, sp.js + 1) - x.FB(ig, sp.js));
        }
    }
}

// ---- band 6: 820-980 cm-1 (low key h2o; low minor co2) (high: cfc11, cfc12 only)  :1319-1400
void taugb6(Column& c) {
    Ctx x(c, 6);
    const double* ka_mco2 = x.B.a("ka_mco2");
    const double* cfc11adj = x.B.a("cfc11adj");
    const double* cfc12 = x.B.a("cfc12");
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        double adjcolco2 =
            adjcol(2, c.colco2[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(2, jp + 1), 3.0, 2.0, 0.77, G.chi(2, jp + 1));
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[5], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absco2 = (minor2(ka_mco2, c, lay, ig));
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself +
                              taufor + adjcolco2 * absco2 + c.WX(2, lay) * cfc11adj[ig - 1] +
                              c.WX(3, lay) * cfc12[ig - 1];
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = 0.0 + c.WX(2, lay) * cfc11adj[ig - 1] + c.WX(3, lay) * cfc12[ig - 1];
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
}

// ---- band 7: 980-1080 cm-1 (low key h2o,o3; low minor co2) (high key o3; high minor co2)  :1403-1672
void taugb7(Column& c) {
    Ctx x(c, 7);
    const double* ka_mco2 = x.B.a("ka_mco2");
    const double* kb_mco2 = x.B.a("kb_mco2");
    double refrat_planck_a = G.chi(1, 3) / G.chi(3, 3);
    double refrat_m_a = G.chi(1, 3) / G.chi(3, 3);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colo3 = c.colo3[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oo3[lay - 1], colo3, 8.);
        Spec s1 = spec(colh2o, c.rat_h2oo3_1[lay - 1], colo3, 8.);
        Spec sm = spec(colh2o, refrat_m_a, colo3, 8.);
        // :1481 default-kind "1.e20" (T1)
        double adjcolco2 = adjcol(3, c.colco2[lay - 1], c.coldry[lay - 1], lit(1.e20), G.chi(2, jp + 1), 3.0, 3.0, 0.79,
                                  G.chi(2, jp + 1));
        Spec sp = spec(colh2o, refrat_planck_a, colo3, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[6], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absco2 = minor3(ka_mco2, 9, sm.js, sm.fs, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + adjcolco2 * absco2;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    const double sc[6] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.855};  // :1664-1669 (_rb literals)
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        // :1637 default-kind "1.e20" (T1); base 2.0 in the upper atmosphere
        double adjcolco2 = adjcol(4, c.colco2[lay - 1], c.coldry[lay - 1], lit(1.e20), G.chi(2, jp + 1), 3.0, 2.0, 0.79,
                                  G.chi(2, jp + 1));
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[6], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double absco2 = minor2(kb_mco2, c, lay, ig);
            x.taug(lay, ig) =
                c.colo3[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + adjcolco2 * absco2;
            x.fracs(lay, ig) = x.FB1(ig);
        }
        for (int k = 0; k < 6; ++k) x.taug(lay, 6 + k) = x.taug(lay, 6 + k) * sc[k];
    }
}

// ---- band 8: 1080-1180 cm-1 (low key h2o; low minor co2,o3,n2o) (high key o3; high minor co2,n2o)  :1675-1806
void taugb8(Column& c) {
    Ctx x(c, 8);
    const double *ka_mco2 = x.B.a("ka_mco2"), *ka_mn2o = x.B.a("ka_mn2o"), *ka_mo3 = x.B.a("ka_mo3");
    const double *kb_mco2 = x.B.a("kb_mco2"), *kb_mn2o = x.B.a("kb_mn2o");
    const double *cfc12 = x.B.a("cfc12"), *cfc22adj = x.B.a("cfc22adj");
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        double adjcolco2 =
            adjcol(5, c.colco2[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(2, jp + 1), 3.0, 2.0, 0.65, G.chi(2, jp + 1));
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[7], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absco2 = (minor2(ka_mco2, c, lay, ig));
            double abso3 = (minor2(ka_mo3, c, lay, ig));
            double absn2o = (minor2(ka_mn2o, c, lay, ig));
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself +
                              taufor + adjcolco2 * absco2 + c.colo3[lay - 1] * abso3 + c.coln2o[lay - 1] * absn2o +
                              c.WX(3, lay) * cfc12[ig - 1] + c.WX(4, lay) * cfc22adj[ig - 1];
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        double adjcolco2 =
            adjcol(6, c.colco2[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(2, jp + 1), 3.0, 2.0, 0.65, G.chi(2, jp + 1));
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[7], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double absco2 = (minor2(kb_mco2, c, lay, ig));
            double absn2o = (minor2(kb_mn2o, c, lay, ig));
            x.taug(lay, ig) = c.colo3[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) +
                              adjcolco2 * absco2 + c.coln2o[lay - 1] * absn2o + c.WX(3, lay) * cfc12[ig - 1] +
                              c.WX(4, lay) * cfc22adj[ig - 1];
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 9: 1180-1390 cm-1 (low key h2o,ch4; low minor n2o) (high key ch4; high minor n2o)  :1809-2066
void taugb9(Column& c) {
    Ctx x(c, 9);
    const double* ka_mn2o = x.B.a("ka_mn2o");
    const double* kb_mn2o = x.B.a("kb_mn2o");
    double refrat_planck_a = G.chi(1, 9) / G.chi(6, 9);
    double refrat_m_a = G.chi(1, 3) / G.chi(6, 3);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colch4 = c.colch4[lay - 1];
        Spec s = spec(colh2o, c.rat_h2och4[lay - 1], colch4, 8.);
        Spec s1 = spec(colh2o, c.rat_h2och4_1[lay - 1], colch4, 8.);
        Spec sm = spec(colh2o, refrat_m_a, colch4, 8.);
        double adjcoln2o = adjcol(7, c.coln2o[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(4, jp + 1), 1.5, 0.5, 0.65,
                                  G.chi(4, jp + 1));
        Spec sp = spec(colh2o, refrat_planck_a, colch4, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[8], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absn2o = minor3(ka_mn2o, 9, sm.js, sm.fs, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + adjcoln2o * absn2o;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        double adjcoln2o = adjcol(8, c.coln2o[lay - 1], c.coldry[lay - 1], 1.e20, G.chi(4, jp + 1), 1.5, 0.5, 0.65,
                                  G.chi(4, jp + 1));
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[8], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double absn2o = minor2(kb_mn2o, c, lay, ig);
            x.taug(lay, ig) =
                c.colch4[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + adjcoln2o * absn2o;
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 10: 1390-1480 cm-1 (h2o)  :2069-2140
void taugb10(Column& c) {
    Ctx x(c, 10);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[9], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself + taufor;
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[9], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double taufor = taufor_(x, c, lay, ig);
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + taufor;
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 11: 1480-1800 cm-1 (h2o; minor o2)  :2143-2232
void taugb11(Column& c) {
    Ctx x(c, 11);
    const double* ka_mo2 = x.B.a("ka_mo2");
    const double* kb_mo2 = x.B.a("kb_mo2");
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[10], 1, 1, ind0, ind1);
        double scaleo2 = c.colo2[lay - 1] * c.scaleminor[lay - 1];
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double tauo2 = scaleo2 * (minor2(ka_mo2, c, lay, ig));
            x.taug(lay, ig) =
                c.colh2o[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself + taufor + tauo2;
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[10], 1, 1, ind0, ind1);
        double scaleo2 = c.colo2[lay - 1] * c.scaleminor[lay - 1];
        for (int ig = 1; ig <= x.ng; ++ig) {
            double taufor = taufor_(x, c, lay, ig);
            double tauo2 = scaleo2 * (minor2(kb_mo2, c, lay, ig));
            x.taug(lay, ig) = c.colh2o[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + taufor + tauo2;
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 12: 1800-2080 cm-1 (low key h2o,co2) (high: nothing)  :2235-2410
void taugb12(Column& c) {
    Ctx x(c, 12);
    double refrat_planck_a = G.chi(1, 10) / G.chi(2, 10);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(colh2o, c.rat_h2oco2[lay - 1], colco2, 8.);
        Spec s1 = spec(colh2o, c.rat_h2oco2_1[lay - 1], colco2, 8.);
        Spec sp = spec(colh2o, refrat_planck_a, colco2, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[11], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay)
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = 0.0;
            x.fracs(lay, ig) = 0.0;
        }
}

// ---- band 13: 2080-2250 cm-1 (low key h2o,n2o; low minor co2, co) (high minor o3)  :2413-2670
void taugb13(Column& c) {
    Ctx x(c, 13);
    const double *ka_mco2 = x.B.a("ka_mco2"), *ka_mco = x.B.a("ka_mco"), *kb_mo3 = x.B.a("kb_mo3");
    double refrat_planck_a = G.chi(1, 5) / G.chi(4, 5);
    double refrat_m_a = G.chi(1, 1) / G.chi(4, 1);
    double refrat_m_a3 = G.chi(1, 3) / G.chi(4, 3);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], coln2o = c.coln2o[lay - 1];
        Spec s = spec(colh2o, c.rat_h2on2o[lay - 1], coln2o, 8.);
        Spec s1 = spec(colh2o, c.rat_h2on2o_1[lay - 1], coln2o, 8.);
        Spec sm = spec(colh2o, refrat_m_a, coln2o, 8.);
        // :2494-2501  ratco2 uses 3.55e-4_rb, adjcolco2 the default-kind 3.55e-4 (T1)
        double adjcolco2 =
            adjcol(9, c.colco2[lay - 1], c.coldry[lay - 1], 1.e20, 3.55e-4, 3.0, 2.0, 0.68, lit(3.55e-4));
        Spec smco = spec(colh2o, refrat_m_a3, coln2o, 8.);
        Spec sp = spec(colh2o, refrat_planck_a, coln2o, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[12], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double absco2 = minor3(ka_mco2, 9, sm.js, sm.fs, c, lay, ig);
            double absco = minor3(ka_mco, 9, smco.js, smco.fs, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + adjcolco2 * absco2 + c.colco[lay - 1] * absco;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        for (int ig = 1; ig <= x.ng; ++ig) {
            double abso3 = minor2(kb_mo3, c, lay, ig);
            x.taug(lay, ig) = c.colo3[lay - 1] * abso3;
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 14: 2250-2380 cm-1 (co2)  :2673-2740
void taugb14(Column& c) {
    Ctx x(c, 14);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[13], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            x.taug(lay, ig) = c.colco2[lay - 1] * key4a(x, ind0, ind1, fac00, fac10, fac01, fac11, ig) + tauself + taufor;
            x.fracs(lay, ig) = x.FA1(ig);
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        ind_up(jp, jt, jt1, G.nspb[13], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = c.colco2[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig);
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

// ---- band 15: 2380-2600 cm-1 (low key n2o,co2; low minor n2) (high: nothing)  :2743-2956
void taugb15(Column& c) {
    Ctx x(c, 15);
    const double* ka_mn2 = x.B.a("ka_mn2");
    double refrat_planck_a = G.chi(4, 1) / G.chi(2, 1);
    double refrat_m_a = G.chi(4, 1) / G.chi(2, 1);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double coln2o = c.coln2o[lay - 1], colco2 = c.colco2[lay - 1];
        Spec s = spec(coln2o, c.rat_n2oco2[lay - 1], colco2, 8.);
        Spec s1 = spec(coln2o, c.rat_n2oco2_1[lay - 1], colco2, 8.);
        Spec sm = spec(coln2o, refrat_m_a, colco2, 8.);
        Spec sp = spec(coln2o, refrat_planck_a, colco2, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[14], s.js, s1.js, ind0, ind1);
        double scalen2 = c.colbrd[lay - 1] * c.scaleminor[lay - 1];
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double taun2 = scalen2 * (minor3(ka_mn2, 9, sm.js, sm.fs, c, lay, ig));
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor + taun2;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay)
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = 0.0;
            x.fracs(lay, ig) = 0.0;
        }
}

// ---- band 16: 2600-3250 cm-1 (low key h2o,ch4) (high key ch4)  :2959-3164
void taugb16(Column& c) {
    Ctx x(c, 16);
    double refrat_planck_a = G.chi(1, 6) / G.chi(6, 6);
    for (int lay = 1; lay <= c.laytrop; ++lay) {
        LAYVARS
        const double colh2o = c.colh2o[lay - 1], colch4 = c.colch4[lay - 1];
        Spec s = spec(colh2o, c.rat_h2och4[lay - 1], colch4, 8.);
        Spec s1 = spec(colh2o, c.rat_h2och4_1[lay - 1], colch4, 8.);
        Spec sp = spec(colh2o, refrat_planck_a, colch4, 8.);
        int ind0, ind1;
        ind_lo(jp, jt, jt1, G.nspa[15], s.js, s1.js, ind0, ind1);
        Fac w0 = eta_fac(s.parm, s.fs, fac00, fac10, 0);
        Fac w1 = eta_fac(s1.parm, s1.fs, fac01, fac11, 1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            double tauself = tauself_(x, c, lay, ig);
            double taufor = taufor_(x, c, lay, ig);
            double tm = tau_major(x, w0, s.comb, ind0, ig);
            double tm1 = tau_major(x, w1, s1.comb, ind1, ig);
            x.taug(lay, ig) = tm + tm1 + tauself + taufor;
            x.fracs(lay, ig) = x.FA(ig, sp.js) + sp.fs * (x.FA(ig, sp.js + 1) - x.FA(ig, sp.js));
        }
    }
    for (int lay = c.laytrop + 1; lay <= c.nlayers; ++lay) {
        LAYVARS
        int ind0, ind1;
        // nspb(16) = 0 -> ind0 = ind1 = 1 for every layer (T2; reproduced, not fixed)
        ind_up(jp, jt, jt1, G.nspb[15], 1, 1, ind0, ind1);
        for (int ig = 1; ig <= x.ng; ++ig) {
            x.taug(lay, ig) = c.colch4[lay - 1] * key4b(x, ind0, ind1, fac00, fac10, fac01, fac11, ig);
            x.fracs(lay, ig) = x.FB1(ig);
        }
    }
}

}  // namespace

void taumol(Column& c) {
    taugb1(c); taugb2(c); taugb3(c); taugb4(c); taugb5(c); taugb6(c); taugb7(c); taugb8(c);
    taugb9(c); taugb10(c); taugb11(c); taugb12(c); taugb13(c); taugb14(c); taugb15(c); taugb16(c);
}

}  // namespace orc
