// orc_common.h -- shared state of the CPU ORACLE (test infrastructure only).
//
// The oracle is a plain C++ restatement of the reference's Fortran algorithm for the
// rrtmg_lw hot path.  It exists to CHECK the CUDA product; nothing under
// rrtmg_lw_b200/ may include, link or call it.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY STATUS: the reference's own Fortran cannot be built here (no Fortran
// compiler in the image) and its k-distribution data (src/rrtmg_lw_k_g.f90,
// data/rrtmg_lw.nc) are missing from the mount, so this oracle is pinned only by
// the data-independent known answers of the reference (Planck surface fluxes of
// the 4 standard atmospheres in run_examples_std_atm/output_rrtm_*, the MT19937
// test vector, structural sums of the g-point reduction maps, closed-form lookup
// tables).  For every k-dependent quantity: "parity unpinned" -- see DESIGN.md.
//
// Module-level state mirrors modules/rrlw_*.f90 of the reference.
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#ifdef ORC_COUNT_OPS
// Op-counting build (liborc_count.so): every `double` of the oracle sources becomes a wrapper
// that counts executed +,-,*,/ and libm calls, giving the ALGORITHMIC FP64 work per column that
// the roofline figure in bench.py / DESIGN.md uses.
namespace orc {
struct OpCount { long long add, mul, div, exp, log, pow; };
extern thread_local OpCount g_ops;
struct Real {
    double v;
    Real() = default;
    Real(double x) : v(x) {}
    Real(float x) : v(x) {}
    Real(int x) : v(x) {}
    Real(long x) : v((double)x) {}
    explicit operator int() const { return (int)v; }
    explicit operator float() const { return (float)v; }
    explicit operator bool() const { return v != 0; }
    Real operator-() const { return Real(-v); }
};
#define ORC_BIN(op, ctr)                                                                   \
    inline Real operator op(Real a, Real b) { g_ops.ctr++; return Real(a.v op b.v); }      \
    inline Real operator op(Real a, double b) { g_ops.ctr++; return Real(a.v op b); }      \
    inline Real operator op(double a, Real b) { g_ops.ctr++; return Real(a op b.v); }      \
    inline Real operator op(Real a, int b) { g_ops.ctr++; return Real(a.v op b); }         \
    inline Real operator op(int a, Real b) { g_ops.ctr++; return Real(a op b.v); }
ORC_BIN(+, add) ORC_BIN(-, add) ORC_BIN(*, mul) ORC_BIN(/, div)
#undef ORC_BIN
#define ORC_CMP(op)                                                    \
    inline bool operator op(Real a, Real b) { return a.v op b.v; }     \
    inline bool operator op(Real a, double b) { return a.v op b; }     \
    inline bool operator op(double a, Real b) { return a op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP
}  // namespace orc
namespace std {
inline orc::Real exp(orc::Real x) { orc::g_ops.exp++; return orc::Real(std::exp(x.v)); }
inline orc::Real log(orc::Real x) { orc::g_ops.log++; return orc::Real(std::log(x.v)); }
inline orc::Real asin(orc::Real x) { return orc::Real(std::asin(x.v)); }
inline orc::Real pow(orc::Real x, orc::Real y) { orc::g_ops.pow++; return orc::Real(std::pow(x.v, y.v)); }
inline orc::Real pow(orc::Real x, double y) { orc::g_ops.pow++; return orc::Real(std::pow(x.v, y)); }
inline orc::Real fmod(orc::Real x, double y) { return orc::Real(std::fmod(x.v, y)); }
}  // namespace std
#define double orc::Real
#endif

namespace orc {

constexpr int NBND = 16;    // nbndlw  modules/parrrtm.f90:35
constexpr int NGPT = 140;   // ngptlw  modules/parrrtm.f90:36
constexpr int MG = 16;      // mg      modules/parrrtm.f90:34
constexpr int NTBL = 10000; // ntbl    modules/rrlw_tbl.f90:34

// one spectral band's REDUCED tables (modules/rrlw_kgNN.f90, the non-"o" arrays)
struct Band {
    int ng = 0;  // ngc(band)
    std::map<std::string, std::vector<double>> arr;  // reduced arrays by reference name
    std::map<std::string, int> lead;                 // leading extent of each
    const double* a(const char* n) const {
        auto it = arr.find(n);
        return it == arr.end() ? nullptr : it->second.data();
    }
};

struct State {
    bool inited = false;
    int drk = 4;  // default real kind of the Fortran build being restated (4: plain gfortran, 8: -r8)
    // rrlw_con
    double grav, planck, boltz, clight, avogad, alosmt, gascon, radcn1, radcn2, sbcnst, secdy;
    double heatfac, fluxfac, oneminus, pi;
    // rrlw_wvn
    double delwave[NBND];
    int ng[NBND], nspa[NBND], nspb[NBND];
    int ngc[NBND], ngs[NBND], ngn[NGPT], ngb[NGPT], ngm[NBND * MG];
    double wt[MG], rwgt[NBND * MG];
    // rrlw_ref  (1-based access through helpers below)
    double pref[59], preflog[59], tref[59], chi_mls[7 * 59];
    // Planck tables
    double totplnk[181 * 16], totplk16[181], totplnkderiv[181 * 16], totplk16deriv[181];
    // rrlw_cld
    double abscld1, absice0[2], absice1[2 * 5], absice2[43 * 16], absice3[46 * 16], absliq0, absliq1[58 * 16];
    // rrlw_tbl
    double tau_tbl[NTBL + 1], exp_tbl[NTBL + 1], tfn_tbl[NTBL + 1], bpade;
    Band band[NBND];

    double chi(int sp, int lev) const { return chi_mls[(sp - 1) + 7 * (lev - 1)]; }
    double tp(int it, int ib) const { return totplnk[(it - 1) + 181 * (ib - 1)]; }
    double tpd(int it, int ib) const { return totplnkderiv[(it - 1) + 181 * (ib - 1)]; }
};

extern State G;

// A default-kind REAL literal of the Fortran source: float32-rounded when the
// build being restated is plain gfortran (drk == 4), exact under -r8.  SURVEY T1.
inline double lit(double x) { return G.drk == 4 ? (double)(float)x : x; }

// per-column work arrays (the locals of rrtmg_lw, src/rrtmg_lw_rad.f90:349-446); 1-based layers
struct Column {
    int nlayers = 0;
    std::vector<double> pavel, tavel, pz, tz, coldry, wbrodl;
    std::vector<double> wkl;  // (7, nlayers)  only the 7 used species
    std::vector<double> wx;   // (4, nlayers)
    double tbound = 0, pwvcm = 0;
    double semiss[NBND];
    std::vector<double> taua;  // (nlayers, 16)
    int inflag = 0, iceflag = 0, liqflag = 0;
    std::vector<double> cldfmc, taucmc, ciwpmc, clwpmc;  // (140, nlayers)
    std::vector<double> reicmc, relqmc;
    // setcoef outputs
    int laytrop = 0;
    std::vector<int> jp, jt, jt1, indself, indfor, indminor;
    std::vector<double> planklay, planklev;  // (nlayers,16), (0:nlayers,16)
    double plankbnd[NBND], dplankbnd_dt[NBND];
    std::vector<double> colh2o, colco2, colo3, coln2o, colco, colch4, colo2, colbrd;
    std::vector<double> fac00, fac01, fac10, fac11;
    std::vector<double> rat_h2oco2, rat_h2oco2_1, rat_h2oo3, rat_h2oo3_1, rat_h2on2o, rat_h2on2o_1,
        rat_h2och4, rat_h2och4_1, rat_n2oco2, rat_n2oco2_1, rat_o3co2, rat_o3co2_1;
    std::vector<double> selffac, selffrac, forfac, forfrac, minorfrac, scaleminor, scaleminorn2;
    // taumol outputs
    std::vector<double> fracs, taug, taut;  // (nlayers, 140)
    int ncbands = 0;
    // rtrnmc outputs (0:nlayers)
    std::vector<double> totuflux, totdflux, fnet, htr, totuclfl, totdclfl, fnetc, htrc, dtotuflux_dt, dtotuclfl_dt;
    std::vector<int> itr_dump;  // optional: table indices (nlayers,140,2) for bit-exact checks

    void resize(int nl);
    // accessors, 1-based
    double& WKL(int m, int l) { return wkl[(m - 1) + 7 * (l - 1)]; }
    double& WX(int m, int l) { return wx[(m - 1) + 4 * (l - 1)]; }
    double& TAUA(int l, int b) { return taua[(l - 1) + nlayers * (b - 1)]; }
    double& CLD(std::vector<double>& v, int ig, int l) { return v[(ig - 1) + NGPT * (l - 1)]; }
    double& PLAY(int l, int b) { return planklay[(l - 1) + nlayers * (b - 1)]; }
    double& PLEV(int l, int b) { return planklev[l + (nlayers + 1) * (b - 1)]; }
    double& LG(std::vector<double>& v, int l, int ig) { return v[(l - 1) + nlayers * (ig - 1)]; }
};

// Branch-coverage counters of the oracle (test infrastructure for the parity suite): how often each data-dependent branch
// of the reference algorithm was taken since the last reset, summed over all threads.  tests/ assert that the batches
// the CUDA path is compared on really exercise every branch.  Layout (orc_branch_counts):
//   [0, 96)   eta interpolation, taumol.f90:569-598 and twins: band (1..16) x corner (ind0 / ind1) x branch
//             (specparm < 0.125, specparm > 0.875, linear): index ((band-1)*2 + corner)*3 + branch
//   [96, 128) atmosphere halves: (band-1)*2 + (0 lower: lay <= laytrop, 1 upper)   -- layers evaluated
//   [128,138) adjfac sites taken (rat > threshold), in source order taumol.f90:550,736,1355,1483,1639,1721,1768,1879,2036,2497
//   [138,141) cloudy-layer sub-branches of rtrnmc.f90:358-425 (odtot < 0.06, odepth <= 0.06, else)
//   [141,143) clear-layer branches rtrnmc.f90:429-440 (odepth <= 0.06, table)
constexpr int NBRANCH = 143;
// compiled in only in the coverage build (liborc_cov.so, -DORC_BRANCH_COUNTERS): the library that is timed as the CPU
// baseline carries no instrumentation
void branch_hit_slow(int idx, long long n);
inline void branch_hit(int idx, long long n = 1) {
#ifdef ORC_BRANCH_COUNTERS
    branch_hit_slow(idx, n);
#else
    (void)idx; (void)n;
#endif
}
extern thread_local int g_cur_band;  // set by the taugbN being evaluated (for the shared eta_fac helper)

// status codes shared by oracle entry points (mirror the reference's STOP sites)
enum {
    ORC_OK = 0,
    ORC_ERR_FLAG = 1,
    ORC_ERR_ICE_RADIUS = 2,
    ORC_ERR_LIQ_RADIUS = 3,
    ORC_ERR_KISS_PRESSURE = 4,
    ORC_ERR_NOT_INIT = 6
};

void setcoef(Column& c, int istart, int idrv);
void taumol(Column& c);
int cldprmc(Column& c);
void rtrnmc(Column& c, int istart, int iend, int iout, int idrv, bool dump_itr);

// Mersenne Twister (src/mcica_random_numbers.f90)
struct MT {
    int32_t state[624];
    int cur;
    void seed(int32_t s);
    int32_t next_int();
    double next_real();
};
void kiss(int32_t& s1, int32_t& s2, int32_t& s3, int32_t& s4, double& r);

}  // namespace orc
#endif
