// lw_init.cpp -- host side of rrtmg_lw_b200_ini: the init-time work of the reference's
// rrtmg_lw_ini (src/rrtmg_lw_init.f90:47-194) done once on the host, producing tables already
// in the device layout of lw_tables.h.
//
// What the reference does here, and where:  physical constants and band limits (lwdatinit,
// :197-300); the 256 -> 140 g-point grouping ngn/ngm and Gauss weights wt (lwcmbdat, :303-382);
// 10001-entry tau/exp/tfn lookup tables (:125-142); relative weights rwgt (:149-173); the
// weighted reduction of every 16-g table to the band's ngc g-points (cmbgb1..16, :385-2034).
// This file computes the same numbers with the same summation order (so they are bit-identical
// to the reference's), but table by table from the descriptor list of
// include/rrtmg_lw_kdata_layout.h, writing each result g-fastest.
#include "lw_tables.h"

#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/rrtmg_lw_kdata_layout.h"

namespace lw {
namespace {

// number of original g-points merged into each reduced g-point (ngn), per band
const int kGroup[NBND][16] = {
    {1, 1, 2, 2, 2, 2, 2, 2, 1, 1},
    {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {2, 2, 2, 2, 2, 2, 2, 2},
    {2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2},
    {2, 2, 2, 2, 2, 2, 2, 2},
    {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2},
    {2, 2, 2, 2, 4, 4},
    {1, 1, 2, 2, 2, 2, 3, 3},
    {1, 1, 1, 1, 2, 2, 4, 4},
    {3, 3, 4, 6},
    {8, 8},
    {8, 8},
    {4, 12}};

const double kGaussWt[16] = {0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
                             0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                             0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};

struct StaticView {  // views into the blob of tools/extract_static_tables.py (STATIC_LAYOUT order)
    const double *pref, *preflog, *tref, *chi, *totplnk, *totplk16, *totplnkderiv, *totplk16deriv, *abscld1,
        *absice0, *absice1, *absice2, *absice3, *absliq0, *absliq1;
    explicit StaticView(const double* p) {
        pref = p; p += 59; preflog = p; p += 59; tref = p; p += 59; chi = p; p += 7 * 59;
        totplnk = p; p += 181 * 16; totplk16 = p; p += 181; totplnkderiv = p; p += 181 * 16; totplk16deriv = p; p += 181;
        abscld1 = p; p += 1; absice0 = p; p += 2; absice1 = p; p += 10; absice2 = p; p += 43 * 16;
        absice3 = p; p += 46 * 16; absliq0 = p; p += 1; absliq1 = p;
    }
};

inline double dk(int drk, double x) { return drk == 4 ? (double)(float)x : x; }

std::string reduced_name(const std::string& n) {
    if (n.compare(0, 3, "kao") == 0) return n.size() == 3 ? "absa" : "ka" + n.substr(3);
    if (n.compare(0, 3, "kbo") == 0) return n.size() == 3 ? "absb" : "kb" + n.substr(3);
    return n.substr(0, n.size() - 1);
}

}  // namespace

void build_host_tables(double cpdair, const double* st, const double* kdata, int drk, HostTables& H) {
    StaticView S(st);
    ConstTab& c = H.ct;
    std::memset(&c, 0, sizeof(c));
    c.drk = drk;
    const double dw[NBND] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
    const double a0[NBND] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    const double a1[NBND] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    const double a2[NBND] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    for (int b = 0; b < NBND; ++b) { c.delwave[b] = dw[b]; c.a0[b] = a0[b]; c.a1[b] = a1[b]; c.a2[b] = a2[b]; }
    std::memcpy(c.pref, S.pref, sizeof(c.pref));
    std::memcpy(c.preflog, S.preflog, sizeof(c.preflog));
    std::memcpy(c.tref, S.tref, sizeof(c.tref));
    std::memcpy(c.chi_mls, S.chi, sizeof(c.chi_mls));
    c.grav = 9.8066;
    c.avogad = 6.02214199e+23;
    const double secdy = 8.6400e4;
    c.heatfac = c.grav * secdy / (cpdair * 1.e2);
    c.oneminus = 1. - 1.e-6;
    c.fluxfac = (2. * std::asin(1.)) * 2.e4;
    c.bpade = 1.0 / 0.278;
    c.e20_b3up = dk(drk, 1.e20);
    c.e20_b7 = dk(drk, 1.e20);
    const double b4[7] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.88, 0.943};
    for (int i = 0; i < 7; ++i) c.b4scale[i] = dk(drk, b4[i]);
    c.co2ref_b13 = dk(drk, 3.55e-4);
    c.absice0[0] = S.absice0[0]; c.absice0[1] = S.absice0[1];
    std::memcpy(c.absice1, S.absice1, sizeof(c.absice1));
    c.abscld1 = *S.abscld1;
    c.absliq0 = *S.absliq0;

    // (181, 16) in the reference's order, followed by a 17th column: totplk16 / totplk16deriv, band 16 integrated over
    // 2600-3250 cm-1 only, which setcoef uses when the run is band 16 alone (istart = 16, setcoef.f90:233-251)
    H.totplnk.assign(S.totplnk, S.totplnk + 181 * 16);
    H.totplnk.insert(H.totplnk.end(), S.totplk16, S.totplk16 + 181);
    H.totplnkderiv.assign(S.totplnkderiv, S.totplnkderiv + 181 * 16);
    H.totplnkderiv.insert(H.totplnkderiv.end(), S.totplk16deriv, S.totplk16deriv + 181);
    H.absice2.assign(S.absice2, S.absice2 + 43 * 16);
    H.absice3.assign(S.absice3, S.absice3 + 46 * 16);
    H.absliq1.assign(S.absliq1, S.absliq1 + 58 * 16);

    // ---- transmittance lookup tables: x = i/ntbl is the Pade variable tau/(bpade+tau)
    H.tau_tbl.assign(NTBL + 1, 0.0);
    H.exp_tbl.assign(NTBL + 1, 0.0);
    H.tfn_tbl.assign(NTBL + 1, 0.0);
    const double tiny = 1.e-20;
    H.tau_tbl[NTBL] = 1.e10; H.exp_tbl[0] = 1.0; H.exp_tbl[NTBL] = tiny; H.tfn_tbl[NTBL] = 1.0;
    for (int i = 1; i < NTBL; ++i) {
        const double x = drk == 4 ? (double)((float)i / (float)NTBL) : (double)i / (double)NTBL;
        const double tau = c.bpade * x / (1. - x);
        double tr = std::exp(-tau);
        if (tr <= tiny) tr = tiny;
        H.tau_tbl[i] = tau;
        H.exp_tbl[i] = tr;
        H.tfn_tbl[i] = tau < 0.06 ? tau / 6. : 1. - 2. * ((1. / tau) - (tr / (1. - tr)));
    }

    // ---- relative weights of the original g-points inside their reduced g-point
    for (int b = 0; b < NBND; ++b) {
        int g16 = 0;
        for (int ig = 0; ig < NGC[b]; ++ig) {
            const int n = kGroup[b][ig];
            double wsum = 0.;
            for (int j = 0; j < n; ++j) wsum = wsum + kGaussWt[g16 + j];
            for (int j = 0; j < n; ++j) H.rwgt[b * 16 + g16 + j] = NGC[b] < 16 ? kGaussWt[g16 + j] / wsum : 1.0;
            g16 += n;
        }
    }

    // ---- reduce every table 16 -> NG and lay it out g-fastest in one arena
    for (int b = 0; b < NBND; ++b)
        for (int s = 0; s < 13; ++s) H.off[b][s] = H.offp[b][s] = std::numeric_limits<size_t>::max();
    H.arena.clear();
    H.arena_p.clear();
    H.named.clear();
    int nminor[NBND] = {0}, nvec[NBND] = {0};
    const double* src = kdata;
    for (int a = 0; a < RRTMG_LW_NKARRAYS; ++a) {
        const rrtmg_lw_karray_desc& d = RRTMG_LW_KARRAYS[a];
        const int b = d.band - 1, ng = NGC[b], lead = d.lead;
        // pad each table start to 16 bytes (double2 loads)
        if (H.arena.size() & 1) H.arena.push_back(0.0);
        const size_t base = H.arena.size();
        H.arena.resize(base + (size_t)lead * ng, 0.0);
        const int ngp = NGPAD[b];
        while (H.arena_p.size() & 15) H.arena_p.push_back(0.0);  // every table starts a 128-byte line
        const size_t basep = H.arena_p.size();
        H.arena_p.resize(basep + (size_t)lead * ngp, 0.0);
        std::vector<double> ref((size_t)lead * ng, 0.0);  // reference element order, for get_table
        for (int i = 0; i < lead; ++i) {
            int g16 = 0;
            for (int ig = 0; ig < ng; ++ig) {
                double acc = 0.;
                for (int j = 0; j < kGroup[b][ig]; ++j, ++g16) {
                    const double x = d.gfirst ? src[g16 + 16 * (size_t)i] : src[i + (size_t)lead * g16];
                    acc = d.weighted ? acc + x * H.rwgt[b * 16 + g16] : acc + x;
                }
                H.arena[base + (size_t)i * ng + ig] = acc;
                H.arena_p[basep + (size_t)i * ngp + ig] = acc;
                if (d.gfirst) ref[ig + (size_t)ng * i] = acc; else ref[i + (size_t)lead * ig] = acc;
            }
        }
        src += (size_t)lead * 16;
        const std::string nm = d.name;
        int slot;
        if (nm == "kao") slot = 0;
        else if (nm == "kbo") slot = 1;
        else if (nm == "selfrefo") slot = 2;
        else if (nm == "forrefo") slot = 3;
        else if (nm == "fracrefao") slot = 4;
        else if (nm == "fracrefbo") slot = 5;
        else if (nm.find("_m") != std::string::npos) slot = 6 + nminor[b]++;
        else slot = 11 + nvec[b]++;
        H.off[b][slot] = base;
        H.offp[b][slot] = basep;
        H.named.emplace_back(std::to_string(d.band) + ":" + reduced_name(nm), std::move(ref));
    }
    if (H.arena.size() & 1) H.arena.push_back(0.0);
    while (H.arena_p.size() & 3) H.arena_p.push_back(0.0);
}


// ---------------------------------------------------------------------------------------------------------------------
// Table image: 8-byte words throughout.  Header: magic, version, sizeof(ConstTab), default real kind, section count;
// then length-prefixed sections in a fixed order; last word: a checksum (sum of all preceding words as uint64).
namespace {
constexpr unsigned long long kTblMagic = 0x314c42544c574c52ull;  // "RLWLTBL1" little-endian
constexpr unsigned long long kTblVersion = 1;

struct Writer {
    unsigned long long* out;
    long cap_words;
    long n = 0;
    unsigned long long sum = 0;
    void word(unsigned long long v) {
        if (out && n < cap_words) out[n] = v;
        sum += v;
        ++n;
    }
    void raw(const void* p, size_t bytes) {  // bytes is a multiple of 8 or padded with zeros
        const size_t w = (bytes + 7) / 8;
        word((unsigned long long)bytes);
        const unsigned char* q = static_cast<const unsigned char*>(p);
        for (size_t k = 0; k < w; ++k) {
            unsigned long long v = 0;
            const size_t m = bytes - 8 * k < 8 ? bytes - 8 * k : 8;
            std::memcpy(&v, q + 8 * k, m);
            word(v);
        }
    }
    void vec(const std::vector<double>& v) { raw(v.data(), v.size() * sizeof(double)); }
};

struct Reader {
    const unsigned long long* in;
    long nwords;
    long n = 0;
    bool ok = true;
    unsigned long long word() {
        if (n >= nwords) { ok = false; return 0; }
        return in[n++];
    }
    bool raw(void* p, size_t expect_bytes) {
        const unsigned long long bytes = word();
        if (!ok || bytes != expect_bytes) { ok = false; return false; }
        const size_t w = (bytes + 7) / 8;
        if (n + (long)w > nwords) { ok = false; return false; }
        std::memcpy(p, in + n, bytes);
        n += (long)w;
        return true;
    }
    bool vec(std::vector<double>& v, size_t max_doubles) {
        if (n >= nwords) { ok = false; return false; }
        const unsigned long long bytes = in[n];
        if (bytes % 8 != 0 || bytes / 8 > max_doubles) { ok = false; return false; }
        v.resize(bytes / 8);
        return raw(v.data(), bytes);
    }
};
}  // namespace

long host_tables_write(const HostTables& H, void* out, long cap) {
    Writer w{static_cast<unsigned long long*>(out), out ? cap / 8 : 0};
    w.word(kTblMagic);
    w.word(kTblVersion);
    w.word(sizeof(ConstTab));
    w.word((unsigned long long)H.ct.drk);
    ConstTab ct = H.ct;  // device pointers are not part of the image
    std::memset(ct.band, 0, sizeof(ct.band));
    std::memset(ct.bandp, 0, sizeof(ct.bandp));
    w.raw(&ct, sizeof(ConstTab));
    w.vec(H.arena);
    w.raw(H.off, sizeof(H.off));
    w.vec(H.arena_p);
    w.raw(H.offp, sizeof(H.offp));
    w.vec(H.totplnk); w.vec(H.totplnkderiv);
    w.vec(H.exp_tbl); w.vec(H.tfn_tbl); w.vec(H.tau_tbl);
    w.vec(H.absice2); w.vec(H.absice3); w.vec(H.absliq1);
    w.raw(H.rwgt, sizeof(H.rwgt));
    w.word((unsigned long long)H.named.size());
    for (const auto& kv : H.named) {
        w.raw(kv.first.data(), kv.first.size());
        w.vec(kv.second);
    }
    const unsigned long long sum = w.sum;
    w.word(sum);
    const long bytes = w.n * 8;
    return bytes;
}

bool host_tables_read(const void* blob, long nbytes, HostTables& H) {
    if (!blob || nbytes < 64 || nbytes % 8 != 0) return false;
    Reader r{static_cast<const unsigned long long*>(blob), nbytes / 8};
    unsigned long long sum = 0;
    for (long k = 0; k + 1 < r.nwords; ++k) sum += r.in[k];
    if (sum != r.in[r.nwords - 1]) return false;
    if (r.word() != kTblMagic || r.word() != kTblVersion || r.word() != sizeof(ConstTab)) return false;
    const unsigned long long drk = r.word();
    if (drk != 4 && drk != 8) return false;
    const size_t big = (size_t)1 << 26;
    if (!r.raw(&H.ct, sizeof(ConstTab))) return false;
    if ((unsigned long long)H.ct.drk != drk) return false;
    if (!r.vec(H.arena, big) || !r.raw(H.off, sizeof(H.off)) || !r.vec(H.arena_p, big) || !r.raw(H.offp, sizeof(H.offp)))
        return false;
    if (!r.vec(H.totplnk, big) || !r.vec(H.totplnkderiv, big) || !r.vec(H.exp_tbl, big) || !r.vec(H.tfn_tbl, big) ||
        !r.vec(H.tau_tbl, big) || !r.vec(H.absice2, big) || !r.vec(H.absice3, big) || !r.vec(H.absliq1, big))
        return false;
    if (!r.raw(H.rwgt, sizeof(H.rwgt))) return false;
    const unsigned long long nnamed = r.word();
    if (!r.ok || nnamed > 4096) return false;
    H.named.clear();
    for (unsigned long long k = 0; k < nnamed; ++k) {
        if (r.n >= r.nwords) return false;
        const unsigned long long len = r.in[r.n];
        if (len > 256) return false;
        std::string name(len, ' ');
        if (!r.raw(&name[0], len)) return false;
        std::vector<double> v;
        if (!r.vec(v, big)) return false;
        H.named.emplace_back(std::move(name), std::move(v));
    }
    // consistency the kernels rely on
    if (H.totplnk.size() != 181 * 17 || H.totplnkderiv.size() != 181 * 17 || H.exp_tbl.size() != NTBL + 1 ||
        H.tfn_tbl.size() != NTBL + 1 || H.tau_tbl.size() != NTBL + 1 || H.absice2.size() != 43 * 16 ||
        H.absice3.size() != 46 * 16 || H.absliq1.size() != 58 * 16)
        return false;
    for (int b = 0; b < NBND; ++b)
        for (int k = 0; k < 13; ++k) {
            if (H.off[b][k] != (size_t)-1 && H.off[b][k] >= H.arena.size()) return false;
            if (H.offp[b][k] != (size_t)-1 && H.offp[b][k] >= H.arena_p.size()) return false;
        }
    return r.ok && r.n == r.nwords - 1;
}

}  // namespace lw
