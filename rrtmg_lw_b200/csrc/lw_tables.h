// lw_tables.h -- init-time tables of the B200 RRTMG_LW path: host build + device layout.
//
// Device layout (DESIGN.md "Data layout in HBM"): every reduced k-table of a band is stored
// as (rows, NG) row-major with the g-point FASTEST, i.e. the transpose of the reference's
// absa(rows, ng) (modules/rrlw_kgNN.f90).  One thread walks the g-points of its band for a
// fixed set of rows, so its loads of one row are contiguous (and 16-byte vectorisable, every
// NG is even) and the rows shared by neighbouring columns are L1/shared-memory hits.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace lw {

constexpr int NBND = 16;
constexpr int NGPT = 140;
constexpr int NTBL = 10000;

// g-points per band / first g-point (0-based) of each band: ngc, ngs of src/rrtmg_lw_init.f90:325-326
constexpr int NGC[NBND] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
constexpr int GOFF[NBND] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138};
// row stride of the band kernel's copy of the k-tables: NG rounded up to 4 g-points, so that a row
// is a whole number of 32-byte sectors and four g-points are one aligned 256-bit load (the pad is 0)
constexpr int NGPAD[NBND] = {12, 12, 16, 16, 16, 8, 12, 8, 12, 8, 8, 8, 4, 4, 4, 4};
constexpr int NSPA[NBND] = {1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9};
constexpr int NSPB[NBND] = {1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0};

// slot of a binary-species band in Scratch::pfs (band order 3, 4, 5, 7, 9, 12, 13, 15, 16), -1 for single-key bands
#ifdef __CUDACC__
#define LW_HD __host__ __device__
#else
#define LW_HD
#endif
LW_HD constexpr int pf_slot(int b) {
    switch (b) {
        case 3: return 0;
        case 4: return 1;
        case 5: return 2;
        case 7: return 3;
        case 9: return 4;
        case 12: return 5;
        case 13: return 6;
        case 15: return 7;
        case 16: return 8;
        default: return -1;
    }
}
constexpr int NPF = 9;

// Pointers into the device table arena for one band.  Minor-gas tables m[] and per-g vectors
// v[] are numbered in the order the band lists them in include/rrtmg_lw_kdata_layout.h:
//   band 1: m0 ka_mn2 m1 kb_mn2        band 3: m0 ka_mn2o(9,19) m1 kb_mn2o(5,19)
//   band 5: m0 ka_mo3(9,19) v0 ccl4    band 6: m0 ka_mco2 v0 cfc11adj v1 cfc12
//   band 7: m0 ka_mco2(9,19) m1 kb_mco2
//   band 8: m0 ka_mco2 m1 ka_mn2o m2 ka_mo3 m3 kb_mco2 m4 kb_mn2o v0 cfc12 v1 cfc22adj
//   band 9: m0 ka_mn2o(9,19) m1 kb_mn2o   band 11: m0 ka_mo2 m1 kb_mo2
//   band 13: m0 ka_mco2(9,19) m1 ka_mco(9,19) m2 kb_mo3   band 15: m0 ka_mn2(9,19)
template <class T>
struct BandTabT {
    const T* ka;    // (65*nspa, NG)
    const T* kb;    // (235*max(nspb,1), NG) or null
    const T* selfref;  // (10, NG)
    const T* forref;   // (4, NG)
    const T* fracrefa; // (9 or 1, NG)
    const T* fracrefb; // (5 or 1, NG) or null
    const T* m[5];
    const T* v[2];
};
using BandTab = BandTabT<double>;
using BandTabF = BandTabT<float>;   // optional FP32 mode: float copy of the padded arena

// Scalars and small tables that live in __constant__ memory.
struct ConstTab {
    BandTab band[NBND];    // rows of NG doubles (g-lane kernel)
    BandTab bandp[NBND];   // rows of NGPAD doubles, 32-byte aligned (column-lane kernel)
    double delwave[NBND];
    double a0[NBND], a1[NBND], a2[NBND];  // diffusivity-angle fit, src/rrtmg_lw_rtrnmc.f90:258-269
    double pref[59], preflog[59], tref[59];
    double chi_mls[7 * 59];               // (species, level) species fastest
    double heatfac, fluxfac, oneminus, bpade;
    double avogad, grav;
    // default-kind literals whose value depends on the Fortran build
    double e20_b3up, e20_b7;       // 1.e20 at taumol.f90:734 and :1481,1637
    double b4scale[7];             // taumol.f90:1028-1034
    double co2ref_b13;             // 3.55e-4 at taumol.f90:2498
    double absice0[2], absice1[10], abscld1, absliq0;
    int drk;
};

// Float copies of what the FP32-mode band kernels read from constant memory (same member names as ConstTab, so
// the generated FP32 kernel source only swaps the object).
struct ConstTabF {
    BandTabF band[NBND];   // = bandp (the FP32 kernels only use padded rows)
    BandTabF bandp[NBND];
    float delwave[NBND];
    float a0[NBND], a1[NBND], a2[NBND];
    float fluxfac, oneminus, bpade;
    float e20_b3up, e20_b7;
    float b4scale[7];
    float co2ref_b13;
    float absice0[2], absice1[10], absliq0;
};

struct DeviceTables {
    double* arena = nullptr;        // all k-tables
    size_t arena_doubles = 0;
    double* totplnk = nullptr;      // (16,181)+... see lw_init.cpp
    double* totplnkderiv = nullptr;
    double* exptfn = nullptr;       // (NTBL+1) x {exp_tbl, tfn_tbl} interleaved -> one 16-byte gather
    double* tau_tbl = nullptr;      // (NTBL+1)
    double* absice2 = nullptr;      // (43,16) reference order
    double* absice3 = nullptr;      // (46,16)
    double* absliq1 = nullptr;      // (58,16)
};

// Host-side result of the rrtmg_lw_ini restatement.
struct HostTables {
    ConstTab ct;                       // band pointers filled after upload
    std::vector<double> arena;         // device-layout k tables
    size_t off[NBND][13];              // arena offsets: ka,kb,self,for,fra,frb,m0..m4,v0,v1 (SIZE_MAX = absent)
    std::vector<double> arena_p;       // same tables with rows padded to NGPAD, every table start 32-byte aligned
    size_t offp[NBND][13];
    std::vector<double> totplnk, totplnkderiv;  // (181,16) reference order
    std::vector<double> exp_tbl, tfn_tbl, tau_tbl;
    std::vector<double> absice2, absice3, absliq1;
    double rwgt[256];
    // reference-order copies for rrtmg_lw_b200_get_table
    std::vector<std::pair<std::string, std::vector<double>>> named;
};

// Builds everything on the host.  static_tables = blob of tools/extract_static_tables.py.
void build_host_tables(double cpdair, const double* static_tables, const double* kdata, int drk, HostTables& out);

// Versioned binary image of a HostTables (the result of the rrtmg_lw_ini restatement: reduced k-tables in device
// layout, lookup tables, constants): what a deployment ships instead of the k-data + the init-time reduction.
// write: returns the image size in bytes; fills `out` when it is non-NULL and `cap` is large enough.
// read: false when the image is truncated, of another version or inconsistent.
long host_tables_write(const HostTables& H, void* out, long cap);
bool host_tables_read(const void* blob, long nbytes, HostTables& H);

}  // namespace lw
