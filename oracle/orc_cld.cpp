// orc_cld.cpp -- ORACLE (test infrastructure): cldprmc, the McICA sub-column generator,
// get_alpha and the two random-number generators.
// Follows src/rrtmg_lw_cldprmc.f90:164-271, src/mcica_subcol_gen_lw.f90:129-178, 257-289,
// 442-679, 731-741 and src/mcica_random_numbers.f90:116-185, 243-295.
// The GCM generator file does not compile as shipped; for icld = 4/5 the
// per-column semantics of src/mcica_subcol_gen_lw.1col.f90:543-611 are applied to every
// column (KISS: each column owns its stream; MT: one global stream in
// (isubcol, column, layer) order, two draws per point).
#include "orc_common.h"
#include "orc_api.h"

namespace orc {

// ---------------------------------------------------------------- cldprmc
int cldprmc(Column& c) {
    static const int icb[16] = {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
    const double cldmin = 1.e-20;
    c.ncbands = 1;
    double abscoice[NGPT], abscoliq[NGPT];
    for (int lay = 1; lay <= c.nlayers; ++lay) {
        for (int ig = 1; ig <= NGPT; ++ig) {
            double ciwp = c.CLD(c.ciwpmc, ig, lay), clwp = c.CLD(c.clwpmc, ig, lay);
            double cwp = ciwp + clwp;
            if (c.CLD(c.cldfmc, ig, lay) >= cldmin && (cwp >= cldmin || c.CLD(c.taucmc, ig, lay) >= cldmin)) {
                if (c.inflag == 0) {
                    return ORC_OK;  // :186-188 returns from inside the loops (T10)
                } else if (c.inflag == 1) {
                    return ORC_ERR_FLAG;  // STOP 'INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA'
                } else if (c.inflag == 2) {
                    double radice = c.reicmc[lay - 1];
                    if (ciwp == 0.0) {
                        abscoice[ig - 1] = 0.0;
                    } else if (c.iceflag == 0) {
                        if (radice < 10.0) return ORC_ERR_ICE_RADIUS;
                        abscoice[ig - 1] = G.absice0[0] + G.absice0[1] / radice;
                    } else if (c.iceflag == 1) {
                        if (radice < 13.0 || radice > 130.) return ORC_ERR_ICE_RADIUS;
                        c.ncbands = 5;
                        int ib = icb[G.ngb[ig - 1] - 1];
                        abscoice[ig - 1] = G.absice1[0 + 2 * (ib - 1)] + G.absice1[1 + 2 * (ib - 1)] / radice;
                    } else if (c.iceflag == 2) {
                        if (radice < 5.0 || radice > 131.0) return ORC_ERR_ICE_RADIUS;
                        c.ncbands = 16;
                        double factor = (radice - 2.) / 3.;
                        int index = (int)factor;
                        if (index == 43) index = 42;
                        double fint = factor - (double)index;
                        int ib = G.ngb[ig - 1];
                        const double* t = G.absice2 + 43 * (ib - 1);
                        abscoice[ig - 1] = t[index - 1] + fint * (t[index] - (t[index - 1]));
                    } else if (c.iceflag == 3) {
                        if (radice < 5.0 || radice > 140.0) return ORC_ERR_ICE_RADIUS;
                        c.ncbands = 16;
                        double factor = (radice - 2.) / 3.;
                        int index = (int)factor;
                        if (index == 46) index = 45;
                        double fint = factor - (double)index;
                        int ib = G.ngb[ig - 1];
                        const double* t = G.absice3 + 46 * (ib - 1);
                        abscoice[ig - 1] = t[index - 1] + fint * (t[index] - (t[index - 1]));
                    }
                    if (clwp == 0.0) {
                        abscoliq[ig - 1] = 0.0;
                    } else if (c.liqflag == 0) {
                        abscoliq[ig - 1] = G.absliq0;
                    } else if (c.liqflag == 1) {
                        double radliq = c.relqmc[lay - 1];
                        if (radliq < 2.5 || radliq > 60.) return ORC_ERR_LIQ_RADIUS;
                        int index = (int)(radliq - 1.5);
                        if (index == 0) index = 1;
                        if (index == 58) index = 57;
                        double fint = radliq - 1.5 - (double)index;
                        int ib = G.ngb[ig - 1];
                        const double* t = G.absliq1 + 58 * (ib - 1);
                        abscoliq[ig - 1] = t[index - 1] + fint * (t[index] - (t[index - 1]));
                    }
                    c.CLD(c.taucmc, ig, lay) = ciwp * abscoice[ig - 1] + clwp * abscoliq[ig - 1];
                }
            }
        }
    }
    return ORC_OK;
}

// ---------------------------------------------------------------- KISS  (mcica:711-743)
void kiss(int32_t& s1, int32_t& s2, int32_t& s3, int32_t& s4, double& r) {
    // two's-complement wrap-around INT32 arithmetic, logical shifts (ishft)
    auto m = [](uint32_t k, int n) -> uint32_t { return k ^ (n >= 0 ? (k << n) : (k >> (-n))); };
    uint32_t u1 = (uint32_t)s1, u2 = (uint32_t)s2, u3 = (uint32_t)s3, u4 = (uint32_t)s4;
    u1 = 69069u * u1 + 1327217885u;
    u2 = m(m(m(u2, 13), -17), 5);
    u3 = 18000u * (u3 & 65535u) + (u3 >> 16);
    u4 = 30903u * (u4 & 65535u) + (u4 >> 16);
    uint32_t k = u1 + u2 + (u3 << 16) + u4;
    s1 = (int32_t)u1; s2 = (int32_t)u2; s3 = (int32_t)u3; s4 = (int32_t)u4;
    int32_t kiss = (int32_t)k;
    r = kiss * 2.328306e-10 + 0.5;
}

// ---------------------------------------------------------------- MT19937
void MT::seed(int32_t s) {  // initialize_scalar :169-185
    uint32_t* st = reinterpret_cast<uint32_t*>(state);
    st[0] = (uint32_t)s;
    for (int i = 1; i < 624; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    cur = 624;
}
int32_t MT::next_int() {
    uint32_t* st = reinterpret_cast<uint32_t*>(state);
    if (cur >= 624) {  // nextState :134-152
        auto twist = [](uint32_t u, uint32_t v) -> uint32_t {
            uint32_t mix = (u & 0x80000000u) | (v & 0x7fffffffu);
            return (mix >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        };
        for (int k = 0; k <= 624 - 397 - 1; ++k) st[k] = st[k + 397] ^ twist(st[k], st[k + 1]);
        for (int k = 624 - 397; k <= 624 - 2; ++k) st[k] = st[k + 397 - 624] ^ twist(st[k], st[k + 1]);
        st[623] = st[396] ^ twist(st[623], st[0]);
        cur = 0;
    }
    uint32_t y = st[cur++];  // temper :154-165
    uint32_t x = y ^ (y >> 11);
    x = x ^ ((x << 7) & 0x9d2c5680u);
    x = x ^ ((x << 15) & 0xefc60000u);
    return (int32_t)(x ^ (x >> 18));
}
double MT::next_real() {  // getRandomReal :276-295 incl. the default-kind 2.0**32 (T12)
    int32_t li = next_int();
    const double denom = 4294967296.0 - 1.0;
    if (li < 0) {
        if (G.drk == 4) {
            float num = (float)li + 4294967296.0f;  // INT32 + REAL(4) -> REAL(4)
            return (double)num / denom;
        }
        return ((double)li + 4294967296.0) / denom;
    }
    return ((double)li) / denom;
}

}  // namespace orc

using namespace orc;

// ---------------------------------------------------------------- get_alpha (mcica:68-180)
extern "C" int orc_get_alpha(int ncol, int ld, int nlayers, int icld, int idcor, double decorr_con, const double* dz,
                             const double* lat, int juldat, const double* cldfrac, double* alpha) {
    const double am1 = lit(1.4315), am2 = lit(2.1219), am4 = lit(-25.584), amr = 7.0, f_one = 1.0;
    std::vector<double> decorr_len(ncol, 0.0), decorr_inv(ncol, f_one);
    auto D = [&](const double* a, int i, int k) { return a[(i - 1) + (size_t)ld * (k - 1)]; };
    if (icld == 4 || icld == 5) {
        if (idcor == 1) {
            double am3;
            if (juldat > 181) am3 = -4. * amr / 365. * (juldat - 272);
            else am3 = 4. * amr / 365. * (juldat - 91);
            for (int i = 1; i <= ncol; ++i) {
                double d = lat[i - 1] - am3;
                double decorr_lat = am1 + am2 * std::exp(-(d * d) / (am4 * am4));
                decorr_len[i - 1] = decorr_lat * 1.e3;
            }
        } else {
            for (int i = 0; i < ncol; ++i) decorr_len[i] = decorr_con;
        }
        for (int i = 0; i < ncol; ++i)
            if (decorr_len[i] >= 0.0) decorr_inv[i] = f_one / decorr_len[i];
    }
    if (icld == 4 || icld == 5) {
        for (int i = 1; i <= ncol; ++i) {
            alpha[(i - 1)] = 0.0;
            for (int k = 2; k <= nlayers; ++k) {
                double a = std::exp(-(0.5 * (D(dz, i, k) + D(dz, i, k - 1))) * decorr_inv[i - 1]);
                if (icld == 5 && D(cldfrac, i, k) == 0.0 && D(cldfrac, i, k - 1) > 0.0) a = 0.0;
                alpha[(i - 1) + (size_t)ld * (k - 1)] = a;
            }
        }
    }
    return ORC_OK;
}

// ---------------------------------------------------------------- mcica_subcol_lw (mcica:183-703)
// Arrays in the reference's Fortran layout: (ncol,nlay) column-fastest with leading
// dimension ld; tauc (16,ld,nlay); outputs (140,ld,nlay).
extern "C" int orc_mcica_subcol_lw(int ncol, int ld, int nlay, int icld, int permuteseed, int* irng,
                                   const double* play, const double* cldfrac, const double* ciwp, const double* clwp,
                                   const double* rei, const double* rel, const double* tauc, const double* alpha,
                                   double* cldfmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl, double* relqmcl,
                                   double* taucmcl) {
    if (!G.inited) return ORC_ERR_NOT_INIT;
    if (icld == 0) return ORC_OK;
    if (icld < 0 || icld > 5) return ORC_ERR_FLAG;
    const int nsubcol = NGPT;
    const double cldmin = 1.0e-20;
    auto at = [&](int i, int k) { return (size_t)(i - 1) + (size_t)ld * (k - 1); };
    for (int k = 1; k <= nlay; ++k)
        for (int i = 1; i <= ncol; ++i) {
            reicmcl[at(i, k)] = rei[at(i, k)];
            relqmcl[at(i, k)] = rel[at(i, k)];
        }
    if (*irng != 0) *irng = 1;
    std::vector<double> cldf((size_t)ncol * nlay), pmid((size_t)ncol * nlay);
    auto cf = [&](int i, int k) -> double& { return cldf[(size_t)(i - 1) + (size_t)ncol * (k - 1)]; };
    auto pm = [&](int i, int k) -> double& { return pmid[(size_t)(i - 1) + (size_t)ncol * (k - 1)]; };
    for (int k = 1; k <= nlay; ++k)
        for (int i = 1; i <= ncol; ++i) {
            pm(i, k) = play[at(i, k)] * 1.e2;
            double v = cldfrac[at(i, k)];
            cf(i, k) = (v < cldmin) ? 0. : v;
        }
    // CDF(isubcol, i, ilev)
    std::vector<double> CDF((size_t)nsubcol * ncol * nlay), CDF2;
    auto cdf = [&](std::vector<double>& a, int s, int i, int k) -> double& {
        return a[(size_t)(s - 1) + (size_t)nsubcol * ((size_t)(i - 1) + (size_t)ncol * (k - 1))];
    };
    if (icld >= 4) CDF2.assign(CDF.size(), 0.0);

    std::vector<int32_t> s1(ncol), s2(ncol), s3(ncol), s4(ncol);
    MT mt;
    if (*irng == 0) {
        for (int i = 1; i <= ncol; ++i) {
            if (pm(i, 1) < pm(i, 2)) return ORC_ERR_KISS_PRESSURE;
            // seedN = (pmid - int(pmid)) * 1000000000_im : REAL*INT -> REAL, truncated on assignment
            auto sd = [&](double p) { return (int32_t)((p - (double)(int)p) * 1000000000.0); };
            s1[i - 1] = sd(pm(i, 1)); s2[i - 1] = sd(pm(i, 2)); s3[i - 1] = sd(pm(i, 3)); s4[i - 1] = sd(pm(i, 4));
        }
        double r;
        for (int n = 1; n <= permuteseed; ++n)
            for (int i = 0; i < ncol; ++i) kiss(s1[i], s2[i], s3[i], s4[i], r);
    } else {
        mt.seed(permuteseed);
    }
    const bool two = icld >= 4;
    if (*irng == 0) {
        // every kissvec call advances ALL columns by one draw; per column the stream is
        // (isubcol outer, ilev inner) [icld 1,2,4,5] or one draw per isubcol [icld 3]
        for (int i = 1; i <= ncol; ++i) {
            double r;
            for (int s = 1; s <= nsubcol; ++s) {
                if (icld == 3) {
                    kiss(s1[i - 1], s2[i - 1], s3[i - 1], s4[i - 1], r);
                    for (int k = 1; k <= nlay; ++k) cdf(CDF, s, i, k) = r;
                } else {
                    for (int k = 1; k <= nlay; ++k) {
                        kiss(s1[i - 1], s2[i - 1], s3[i - 1], s4[i - 1], r);
                        cdf(CDF, s, i, k) = r;
                        if (two) {
                            kiss(s1[i - 1], s2[i - 1], s3[i - 1], s4[i - 1], r);
                            cdf(CDF2, s, i, k) = r;
                        }
                    }
                }
            }
        }
    } else {
        for (int s = 1; s <= nsubcol; ++s)
            for (int i = 1; i <= ncol; ++i) {
                if (icld == 3) {
                    double r = mt.next_real();
                    for (int k = 1; k <= nlay; ++k) cdf(CDF, s, i, k) = r;
                } else {
                    for (int k = 1; k <= nlay; ++k) {
                        cdf(CDF, s, i, k) = mt.next_real();
                        if (two) cdf(CDF2, s, i, k) = mt.next_real();
                    }
                }
            }
    }
    if (icld == 2) {  // :533-543
        for (int k = 2; k <= nlay; ++k)
            for (int i = 1; i <= ncol; ++i)
                for (int s = 1; s <= nsubcol; ++s) {
                    if (cdf(CDF, s, i, k - 1) > 1. - cf(i, k - 1)) cdf(CDF, s, i, k) = cdf(CDF, s, i, k - 1);
                    else cdf(CDF, s, i, k) = cdf(CDF, s, i, k) * (1. - cf(i, k - 1));
                }
    } else if (icld >= 4) {  // :599-603 / 1col:572-576
        for (int k = 2; k <= nlay; ++k)
            for (int i = 1; i <= ncol; ++i) {
                double a = alpha[at(i, k)];
                for (int s = 1; s <= nsubcol; ++s)
                    if (cdf(CDF2, s, i, k) < a) cdf(CDF, s, i, k) = cdf(CDF, s, i, k - 1);
            }
    }
    // :649-679
    for (int k = 1; k <= nlay; ++k)
        for (int i = 1; i <= ncol; ++i)
            for (int s = 1; s <= nsubcol; ++s) {
                size_t o = (size_t)(s - 1) + (size_t)NGPT * at(i, k);
                bool cloudy = cdf(CDF, s, i, k) >= 1. - cf(i, k);
                if (cloudy) {
                    cldfmcl[o] = 1.;
                    clwpmcl[o] = clwp[at(i, k)];
                    ciwpmcl[o] = ciwp[at(i, k)];
                    int n = G.ngb[s - 1];
                    taucmcl[o] = tauc[(size_t)(n - 1) + (size_t)NBND * at(i, k)];
                } else {
                    cldfmcl[o] = 0.; clwpmcl[o] = 0.; ciwpmcl[o] = 0.; taucmcl[o] = 0.;
                }
            }
    return ORC_OK;
}
