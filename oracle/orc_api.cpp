// orc_api.cpp -- ORACLE (test infrastructure): rrtmg_lw driver + inatm and the C entry points.
// Follows src/rrtmg_lw_rad.f90:449-593 (rrtmg_lw) and :755-922 (inatm).
#include <cstdio>
#include <cstdlib>

#include "orc_common.h"
#include <atomic>

#include "orc_api.h"

namespace orc {
int init(double cpdair, const double* st, const double* kdata, int drk);

struct In {
    int ncol, ld, nlay, icld, idrv;
    const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2, *cfc11, *cfc12, *cfc22, *ccl4,
        *emis;
    int inflg, iceflg, liqflg;
    const double *cldfmcl, *taucmcl, *ciwpmcl, *clwpmcl, *reicmcl, *relqmcl, *tauaer;
    double at2(const double* a, int iplon, int k) const { return a[(size_t)(iplon - 1) + (size_t)ld * (k - 1)]; }
};

// src/rrtmg_lw_rad.f90:755-922
static void inatm(const In& in, int iplon, int iaer, Column& c) {
    const double amd = 28.9660, amw = 18.0160;
    const int nlayers = in.nlay;
    c.resize(nlayers);
    double amttl = 0.0, wvttl = 0.0;
    c.tbound = in.tsfc[iplon - 1];
    c.pz[0] = in.at2(in.plev, iplon, 1);
    c.tz[0] = in.at2(in.tlev, iplon, 1);
    for (int l = 1; l <= nlayers; ++l) {
        c.pavel[l - 1] = in.at2(in.play, iplon, l);
        c.tavel[l - 1] = in.at2(in.tlay, iplon, l);
        c.pz[l] = in.at2(in.plev, iplon, l + 1);
        c.tz[l] = in.at2(in.tlev, iplon, l + 1);
        c.WKL(1, l) = in.at2(in.h2o, iplon, l);
        c.WKL(2, l) = in.at2(in.co2, iplon, l);
        c.WKL(3, l) = in.at2(in.o3, iplon, l);
        c.WKL(4, l) = in.at2(in.n2o, iplon, l);
        c.WKL(6, l) = in.at2(in.ch4, iplon, l);
        c.WKL(7, l) = in.at2(in.o2, iplon, l);
        double amm = (1. - c.WKL(1, l)) * amd + c.WKL(1, l) * amw;
        c.coldry[l - 1] = (c.pz[l - 1] - c.pz[l]) * 1.e3 * G.avogad / (1.e2 * G.grav * amm * (1. + c.WKL(1, l)));
    }
    for (int l = 1; l <= nlayers; ++l) {
        c.WX(1, l) = in.at2(in.ccl4, iplon, l);
        c.WX(2, l) = in.at2(in.cfc11, iplon, l);
        c.WX(3, l) = in.at2(in.cfc12, iplon, l);
        c.WX(4, l) = in.at2(in.cfc22, iplon, l);
    }
    for (int l = 1; l <= nlayers; ++l) {
        double summol = 0.0;
        for (int imol = 2; imol <= 7; ++imol) summol = summol + c.WKL(imol, l);
        c.wbrodl[l - 1] = c.coldry[l - 1] * (1. - summol);
        for (int imol = 1; imol <= 7; ++imol) c.WKL(imol, l) = c.coldry[l - 1] * c.WKL(imol, l);
        amttl = amttl + c.coldry[l - 1] + c.WKL(1, l);
        wvttl = wvttl + c.WKL(1, l);
        for (int ix = 1; ix <= 4; ++ix) c.WX(ix, l) = c.coldry[l - 1] * c.WX(ix, l) * 1.e-20;  // ixindx(1:4)=1..4
    }
    double wvsh = (amw * wvttl) / (amd * amttl);
    c.pwvcm = wvsh * (1.e3 * c.pz[0]) / (1.e2 * G.grav);
    for (int n = 1; n <= NBND; ++n) c.semiss[n - 1] = in.emis[(size_t)(iplon - 1) + (size_t)in.ld * (n - 1)];
    if (iaer >= 1) {
        for (int l = 1; l <= nlayers; ++l)
            for (int ib = 1; ib <= NBND; ++ib)
                c.TAUA(l, ib) = in.tauaer[(size_t)(iplon - 1) + (size_t)in.ld * ((l - 1) + (size_t)nlayers * (ib - 1))];
    }
    if (in.icld >= 1) {
        c.inflag = in.inflg; c.iceflag = in.iceflg; c.liqflag = in.liqflg;
        for (int l = 1; l <= nlayers; ++l) {
            size_t base = (size_t)NGPT * ((size_t)(iplon - 1) + (size_t)in.ld * (l - 1));
            for (int ig = 1; ig <= NGPT; ++ig) {
                c.CLD(c.cldfmc, ig, l) = in.cldfmcl[base + ig - 1];
                c.CLD(c.taucmc, ig, l) = in.taucmcl[base + ig - 1];
                c.CLD(c.ciwpmc, ig, l) = in.ciwpmcl[base + ig - 1];
                c.CLD(c.clwpmc, ig, l) = in.clwpmcl[base + ig - 1];
            }
            c.reicmc[l - 1] = in.at2(in.reicmcl, iplon, l);
            c.relqmc[l - 1] = in.at2(in.relqmcl, iplon, l);
        }
    }
}

// body of "do iplon" in rrtmg_lw, src/rrtmg_lw_rad.f90:486-571
// band range of the following runs (the column-mode program's istart / iend / iout of a per-band output block,
// src/rrtmg_lw.1col.f90:463-467): set by orc_set_band_range, 1..16 by default
static int g_istart = 1, g_iend = 16;
static int column(const In& in, int iplon, Column& c, bool dump) {
    const int istart = g_istart, iend = g_iend, iout = (g_istart == 1 && g_iend == 16) ? 0 : g_istart, iaer = 10;
    inatm(in, iplon, iaer, c);
    int rc = cldprmc(c);
    if (rc != ORC_OK) return rc;
    setcoef(c, istart, in.idrv);
    taumol(c);
    for (int k = 1; k <= c.nlayers; ++k)
        for (int ig = 1; ig <= NGPT; ++ig) c.LG(c.taut, k, ig) = c.LG(c.taug, k, ig) + c.TAUA(k, G.ngb[ig - 1]);
    rtrnmc(c, istart, iend, iout, in.idrv, dump);
    return ORC_OK;
}
}  // namespace orc

using namespace orc;

extern "C" int orc_set_band_range(int istart, int iend) {
    if (istart < 1 || iend > 16 || istart > iend) return ORC_ERR_FLAG;
    orc::g_istart = istart; orc::g_iend = iend;
    return ORC_OK;
}

extern "C" int orc_ini(double cpdair, const double* st, const double* kdata, int drk) {
    if (drk != 4 && drk != 8) return ORC_ERR_FLAG;
    return init(cpdair, st, kdata, drk);
}

#define FILL_IN(in)                                                                                        \
    in.ncol = ncol; in.ld = ld; in.nlay = nlay; in.icld = icld_v; in.idrv = idrv;                          \
    in.play = play; in.plev = plev; in.tlay = tlay; in.tlev = tlev; in.tsfc = tsfc; in.h2o = h2ovmr;       \
    in.o3 = o3vmr; in.co2 = co2vmr; in.ch4 = ch4vmr; in.n2o = n2ovmr; in.o2 = o2vmr; in.cfc11 = cfc11vmr;  \
    in.cfc12 = cfc12vmr; in.cfc22 = cfc22vmr; in.ccl4 = ccl4vmr; in.emis = emis; in.inflg = inflglw;       \
    in.iceflg = iceflglw; in.liqflg = liqflglw; in.cldfmcl = cldfmcl; in.taucmcl = taucmcl;                \
    in.ciwpmcl = ciwpmcl; in.clwpmcl = clwpmcl; in.reicmcl = reicmcl; in.relqmcl = relqmcl; in.tauaer = tauaer;

extern "C" int orc_rrtmg_lw(int ncol, int ld, int nlay, int* icld, int idrv, const double* play, const double* plev,
                            const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                            const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                            const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                            const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                            int iceflglw, int liqflglw, const double* cldfmcl, const double* taucmcl,
                            const double* ciwpmcl, const double* clwpmcl, const double* reicmcl,
                            const double* relqmcl, const double* tauaer, double* uflx, double* dflx, double* hr,
                            double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt,
                            int nthreads) {
    if (!G.inited) return ORC_ERR_NOT_INIT;
    if (*icld < 0 || *icld > 3) *icld = 2;  // src/rrtmg_lw_rad.f90:470 (T6)
    int icld_v = *icld;
    In in;
    FILL_IN(in)
    int status = ORC_OK;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        Column c;
#pragma omp for schedule(dynamic, 16)
        for (int iplon = 1; iplon <= ncol; ++iplon) {
            int rc = column(in, iplon, c, false);
            if (rc != ORC_OK) {
#pragma omp critical
                if (status == ORC_OK) status = rc;
                continue;
            }
            auto O = [&](double* a, int k) -> double& { return a[(size_t)(iplon - 1) + (size_t)ld * k]; };
            for (int k = 0; k <= nlay; ++k) {
                O(uflx, k) = c.totuflux[k]; O(dflx, k) = c.totdflux[k];
                O(uflxc, k) = c.totuclfl[k]; O(dflxc, k) = c.totdclfl[k];
            }
            for (int k = 0; k <= nlay - 1; ++k) { O(hr, k) = c.htr[k]; O(hrc, k) = c.htrc[k]; }
            if (idrv == 1 && duflx_dt && duflxc_dt)
                for (int k = 0; k <= nlay; ++k) { O(duflx_dt, k) = c.dtotuflux_dt[k]; O(duflxc_dt, k) = c.dtotuclfl_dt[k]; }
        }
    }
    return status;
}

// rrtmg_lw of src/rrtmg_lw_rad.nomcica.f90:99-588: grid-mean cloud properties, cldprop + rtrn (icld = 1, random
// overlap) or rtrnmr (icld = 0, 2, 3: maximum-random overlap).  cldfr, cicewp, cliqwp, reice, reliq (ncol, nlay);
// taucld (nbndlw, ncol, nlay).
namespace orc {
int cldprop(int nlayers, int inflag, int iceflag, int liqflag, const double* cldfrac, const double* tauc,
            const double* ciwp, const double* clwp, const double* rei, const double* rel, int& ncbands,
            std::vector<double>& taucloud);
void rtrn_nomcica(Column& c, bool maxrand, const double* cldfrac, int ncbands, const std::vector<double>& taucloud,
                  int istart, int iend, int iout, int idrv);
}
extern "C" int orc_rrtmg_lw_nomcica(int ncol, int ld, int nlay, int* icld, int idrv, const double* play, const double* plev,
                                    const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                                    const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                                    const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                                    const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                                    int iceflglw, int liqflglw, const double* cldfr, const double* taucld,
                                    const double* cicewp, const double* cliqwp, const double* reice, const double* reliq,
                                    const double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc,
                                    double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt, int nthreads) {
    if (!G.inited) return ORC_ERR_NOT_INIT;
    if (*icld < 0 || *icld > 3) *icld = 2;  // rad.nomcica.f90:458
    const int icld_run = *icld;
    int icld_v = 0;  // inatm below copies no McICA arrays
    const double *cldfmcl = nullptr, *taucmcl = nullptr, *ciwpmcl = nullptr, *clwpmcl = nullptr, *reicmcl = nullptr,
                 *relqmcl = nullptr;
    In in;
    FILL_IN(in)
    int status = ORC_OK;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        Column c;
        // cfpad: one element of slack either side -- rtrnmr reads cldfrac(lev+1) at the top layer and cldfrac(lev-1) at
        // layer 1 when it fills faccmb* (rtrnmr.f90:420-423,498-501); those values never reach an output
        std::vector<double> cfpad(nlay + 2, 0.);
        double* cf = cfpad.data() + 1;
        std::vector<double> ci(nlay), cl(nlay), re(nlay), rl(nlay), tc((size_t)NBND * nlay), taucloud;
#pragma omp for schedule(dynamic, 16)
        for (int iplon = 1; iplon <= ncol; ++iplon) {
            const int istart = g_istart, iend = g_iend, iout = (g_istart == 1 && g_iend == 16) ? 0 : g_istart, iaer = 10;
            inatm(in, iplon, iaer, c);
            for (int l = 1; l <= nlay; ++l) {
                if (icld_run >= 1) {
                    cf[l - 1] = in.at2(cldfr, iplon, l); ci[l - 1] = in.at2(cicewp, iplon, l);
                    cl[l - 1] = in.at2(cliqwp, iplon, l); re[l - 1] = in.at2(reice, iplon, l);
                    rl[l - 1] = in.at2(reliq, iplon, l);
                    for (int n = 1; n <= NBND; ++n)
                        tc[(n - 1) + NBND * (l - 1)] = taucld[(n - 1) + NBND * ((size_t)(iplon - 1) + (size_t)ld * (l - 1))];
                } else {
                    cf[l - 1] = ci[l - 1] = cl[l - 1] = re[l - 1] = rl[l - 1] = 0.;
                    for (int n = 1; n <= NBND; ++n) tc[(n - 1) + NBND * (l - 1)] = 0.;
                }
            }
            int ncbands = 1;
            int rc = cldprop(nlay, icld_run >= 1 ? inflglw : 0, iceflglw, liqflglw, cf, tc.data(), ci.data(), cl.data(),
                             re.data(), rl.data(), ncbands, taucloud);
            if (rc != ORC_OK) {
#pragma omp critical
                if (status == ORC_OK) status = rc;
                continue;
            }
            setcoef(c, istart, idrv);
            taumol(c);
            for (int k = 1; k <= c.nlayers; ++k)
                for (int ig = 1; ig <= NGPT; ++ig) c.LG(c.taut, k, ig) = c.LG(c.taug, k, ig) + c.TAUA(k, G.ngb[ig - 1]);
            rtrn_nomcica(c, icld_run != 1, cf, ncbands, taucloud, istart, iend, iout, idrv);
            auto O = [&](double* a, int k) -> double& { return a[(size_t)(iplon - 1) + (size_t)ld * k]; };
            for (int k = 0; k <= nlay; ++k) {
                O(uflx, k) = c.totuflux[k]; O(dflx, k) = c.totdflux[k];
                O(uflxc, k) = c.totuclfl[k]; O(dflxc, k) = c.totdclfl[k];
            }
            for (int k = 0; k <= nlay - 1; ++k) { O(hr, k) = c.htr[k]; O(hrc, k) = c.htrc[k]; }
            if (idrv == 1 && duflx_dt && duflxc_dt)
                for (int k = 0; k <= nlay; ++k) { O(duflx_dt, k) = c.dtotuflux_dt[k]; O(duflxc_dt, k) = c.dtotuclfl_dt[k]; }
        }
    }
    return status;
}

extern "C" int orc_column_stages(int icol, int ncol, int ld, int nlay, int icld_v, int idrv, const double* play,
                                 const double* plev, const double* tlay, const double* tlev, const double* tsfc,
                                 const double* h2ovmr, const double* o3vmr, const double* co2vmr,
                                 const double* ch4vmr, const double* n2ovmr, const double* o2vmr,
                                 const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr,
                                 const double* ccl4vmr, const double* emis, int inflglw, int iceflglw, int liqflglw,
                                 const double* cldfmcl, const double* taucmcl, const double* ciwpmcl,
                                 const double* clwpmcl, const double* reicmcl, const double* relqmcl,
                                 const double* tauaer, int* ints6, double* facs, double* taug, double* fracs, int* itr,
                                 double* scal) {
    if (!G.inited) return ORC_ERR_NOT_INIT;
    In in;
    FILL_IN(in)
    Column c;
    int rc = column(in, icol + 1, c, true);
    if (rc != ORC_OK) return rc;
    for (int l = 0; l < nlay; ++l) {
        if (ints6) {
            bool lower = l < c.laytrop;
            ints6[6 * l + 0] = c.jp[l]; ints6[6 * l + 1] = c.jt[l]; ints6[6 * l + 2] = c.jt1[l];
            ints6[6 * l + 3] = lower ? c.indself[l] : 0; ints6[6 * l + 4] = c.indfor[l]; ints6[6 * l + 5] = c.indminor[l];
        }
        if (facs) { facs[4 * l] = c.fac00[l]; facs[4 * l + 1] = c.fac01[l]; facs[4 * l + 2] = c.fac10[l]; facs[4 * l + 3] = c.fac11[l]; }
    }
    if (taug) std::memcpy(taug, c.taug.data(), sizeof(double) * nlay * NGPT);
    if (fracs) std::memcpy(fracs, c.fracs.data(), sizeof(double) * nlay * NGPT);
    if (itr) std::memcpy(itr, c.itr_dump.data(), sizeof(int) * nlay * NGPT * 2);
    if (scal) { scal[0] = c.laytrop; scal[1] = c.pwvcm; }
    return ORC_OK;
}

extern "C" void orc_mt_stream(int seed, int n, unsigned int* ints_out, double* reals_out) {
    MT a, b;
    a.seed(seed); b.seed(seed);
    for (int i = 0; i < n; ++i) {
        if (ints_out) ints_out[i] = (unsigned int)a.next_int();
        if (reals_out) reals_out[i] = b.next_real();
    }
}

extern "C" void orc_kiss_stream(int s1, int s2, int s3, int s4, int n, double* reals_out, int* state_out) {
    int32_t a = s1, b = s2, c = s3, d = s4;
    for (int i = 0; i < n; ++i) {
        double r;
        kiss(a, b, c, d, r);
        if (reals_out) reals_out[i] = r;
    }
    if (state_out) { state_out[0] = a; state_out[1] = b; state_out[2] = c; state_out[3] = d; }
}

extern "C" long orc_get_table(const char* name, double* out, long cap) {
    std::string n(name);
    const double* src = nullptr;
    long cnt = 0;
    if (n == "tau_tbl") { src = G.tau_tbl; cnt = NTBL + 1; }
    else if (n == "exp_tbl") { src = G.exp_tbl; cnt = NTBL + 1; }
    else if (n == "tfn_tbl") { src = G.tfn_tbl; cnt = NTBL + 1; }
    else if (n == "rwgt") { src = G.rwgt; cnt = 256; }
    else if (n == "heatfac") { src = &G.heatfac; cnt = 1; }
    else if (n == "fluxfac") { src = &G.fluxfac; cnt = 1; }
    else {
        size_t p = n.find(':');
        if (p == std::string::npos) return 0;
        int b = std::atoi(n.substr(0, p).c_str());
        if (b < 1 || b > 16) return 0;
        auto it = G.band[b - 1].arr.find(n.substr(p + 1));
        if (it == G.band[b - 1].arr.end()) return 0;
        src = it->second.data(); cnt = (long)it->second.size();
    }
    if (out) for (long i = 0; i < cnt && i < cap; ++i) out[i] = src[i];
    return cnt;
}

// branch-coverage counters (orc_common.h NBRANCH): copies them out and optionally resets them
// (counted only in the coverage build liborc_cov.so; all zero otherwise)
static std::atomic<long long> g_branch[orc::NBRANCH];
namespace orc {
thread_local int g_cur_band = 0;
void branch_hit_slow(int idx, long long n) {
    if (idx >= 0 && idx < NBRANCH) g_branch[idx].fetch_add(n, std::memory_order_relaxed);
}
}  // namespace orc

extern "C" int orc_branch_counts(long long* out, int cap, int reset) {
    for (int i = 0; i < orc::NBRANCH; ++i) {
        if (out && i < cap) out[i] = g_branch[i].load();
        if (reset) g_branch[i].store(0);
    }
    return orc::NBRANCH;
}

// op-counting build only (liborc_count.so): counters of the calling thread since the last reset
extern "C" void orc_opcount(long long* out6, int reset) {
#ifdef ORC_COUNT_OPS
    if (out6) { out6[0] = g_ops.add; out6[1] = g_ops.mul; out6[2] = g_ops.div; out6[3] = g_ops.exp; out6[4] = g_ops.log; out6[5] = g_ops.pow; }
    if (reset) g_ops = OpCount{0, 0, 0, 0, 0, 0};
#else
    if (out6) for (int i = 0; i < 6; ++i) out6[i] = -1;
    (void)reset;
#endif
}
