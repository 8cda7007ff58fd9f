// lw_taumol.cuh -- phase 1 of the two-phase longwave path: column scalars, then inatm + setcoef + taumol with one
// thread per (column, LAYER).
//
// What it computes is the reference's inatm (src/rrtmg_lw_rad.f90:755-870), setcoef (src/rrtmg_lw_setcoef.f90:271-432)
// and taugb1..16 (src/rrtmg_lw_taumol.f90:304-3164) followed by the aerosol add and the diffusivity-angle scaling that
// open rtrnmc's layer loop (rad.f90:548-553, rtrnmc.f90:258-281, 352): the layer optical depth along the diffusivity
// angle  odepth(g) = max(0, secdiff(band) * (taug(g) + tauaer(band)))  for all 140 g-points of one layer of one column.
// How it computes it is specific to this implementation:
//   * nothing on this part of the path is a vertical recurrence, so the layer axis is parallel too: one thread owns one
//     (column, layer), a CTA owns 128 consecutive columns of one layer.  Lanes are consecutive columns (every input is a
//     coalesced load, neighbouring lanes gather neighbouring k-table rows), there are ncol x nlay independent threads,
//     no loop-carried registers and no per-layer scratch fields: the setcoef results of the layer stay in the CTA's
//     shared memory and are consumed by the 16 band evaluations that follow in the same thread;
//   * the two whole-column quantities the layer needs (laytrop, and pwvcm -> secdiff per band) come from a light
//     column-scalar kernel that runs the reference's sequential sums over the layers;
//   * the 140 optical depths go to the sweep state as 8-byte values, column-innermost in blocks of 128 columns
//     ([column / 128][layer][g / 2][column % 128][2]): one coalesced 16-byte store per lane and g-point pair here, one coalesced read
//     per sweep direction in lw_rtrnmc.cuh.  Binary-species bands also pass the eta of their Planck-fraction
//     interpolation (one value per band and layer) so that the sweeps can rebuild fracs(g) bit for bit;
//   * k-tables: g-fastest rows padded to four g-points, 256-bit loads (lw_band.cuh BandLayer, unchanged arithmetic).
// Arithmetic keeps the reference's association order (translation unit compiled with -fmad=false).
#pragma once
#include "lw_band.cuh"

namespace lw {

// ===== LW_COMMON_BEGIN =====
#ifndef LW_TAUMOL_THREADS
#define LW_TAUMOL_THREADS 128
#endif
constexpr int TAUMOL_THREADS = LW_TAUMOL_THREADS;  // a multiple of the 128-column block of the scratch layout
constexpr int OD_BLK = 128;
static_assert(TAUMOL_THREADS % OD_BLK == 0, "a CTA covers whole 128-column blocks");
// ===== LW_COMMON_END =====

// ===== LW_PRECISION_BEGIN =====
// The 8-byte sweep-state word of one (g, layer, column).  rtrnmc uses the layer optical depth odepth in one of two ways
// (rtrnmc.f90:429-440): odepth <= 0.06 -> the small-optical-depth series in odepth itself; otherwise ONLY through the
// table index itr = int(1e4 * odepth / (bpade + odepth) + 0.5) (transmittance exp_tbl(itr), Planck weight
// tfn_tbl(itr), and in cloudy layers the table optical depth tau_tbl(itr), :386-387).  Phase 1 therefore stores the
// optical depth itself in the first case and the index in the second -- as the low word of a double whose sign bit is
// set (an optical depth is never negative) -- so that the sweeps, which run the vertical recurrences and are latency
// bound, neither divide nor convert: the division happens once here, where the FP64 pipe has idle slots.
__device__ __forceinline__ double od_state(double odepth, double bpade) {
    const double tblind = div_fast(odepth, bpade + odepth);
    const int itr = (int)(10000.0 * tblind + 0.5);
    return odepth <= 0.06 ? odepth : __hiloint2double((int)0x80000000u, itr);
}
__device__ __forceinline__ bool od_state_is_index(double st) { return __double2hiint(st) < 0; }
__device__ __forceinline__ int od_state_index(double st) { return __double2loint(st); }

// setcoef fields of this thread's (column, layer): its own column of a [field][thread] block of shared memory
struct LayerFs {
    const double* base;
    __device__ __forceinline__ double f(int field) const { return base[field * TAUMOL_THREADS]; }
};

// gas optical depth of band B for one (column, layer) -> odepth of its NG g-points; eta of the Planck fractions
template <int B>
__device__ __forceinline__ void band_odepth(const LayerFs& L, int ipk, bool lower, double pavel, double secdiff, double taua,
                                            double* __restrict__ odp, double* __restrict__ pfp, bool live, bool dbg,
                                            const Scratch& s, int l, int nlay) {
    constexpr int NG = NGC[B - 1];
    constexpr int G0 = GOFF[B - 1];
#ifdef LW_F32
    constexpr int VEC = NGPAD[B - 1] % 8 == 0 ? 8 : 4;
#else
#ifndef LW_TAUMOL_VEC
#define LW_TAUMOL_VEC 4  // g-points evaluated per table-row burst: 4 = one 256-bit load per row in flight (8 = two: more
                         // registers in flight, spills at 128 registers; measured 5 % slower, profiles/r02_taumol_vec_experiment.log)
#endif
    constexpr int VEC = NGPAD[B - 1] < LW_TAUMOL_VEC ? NGPAD[B - 1] : LW_TAUMOL_VEC;
#endif
    using GV = typename std::conditional<VEC == 4, D4, VA<VEC / 4>>::type;
    using BL = BandLayer<B, true>;
    BL W;
    W.smpl = 0.;
    W.setup(L, ipk, lower, pavel);
    // pairs of adjacent g-points are one 16-byte store: [pair][column % 128][2] (every NG and every G0 is even)
    auto emit = [&](int g, const GV& tv) {
        double a[VEC];
        unpack(tv, a);
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            if (g + j < NG) {
                double od[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double taut = a[j + k] + taua;  // rad.f90:548-553
                    const double odepth = secdiff * taut; // rtrnmc.f90:352
                    od[k] = od_state(odepth < 0.0 ? 0.0 : odepth, C.bpade);
                    if (dbg) s.dbg_taug[(long)(G0 + g + j + k) * nlay + l] = a[j + k];
                }
                if (live) __stcs(reinterpret_cast<double2*>(odp + (G0 + g + j) * OD_BLK), make_double2(od[0], od[1]));
            }
        }
    };
    if (lower) {
#pragma unroll
        for (int g = 0; g < NG; g += VEC) {
            GV tv, pv;
            W.template evalv_t<GV, true>(g, tv, pv);
            emit(g, tv);
        }
    } else {
#pragma unroll
        for (int g = 0; g < NG; g += VEC) {
            GV tv, pv;
            W.template evalv_t<GV, false>(g, tv, pv);
            emit(g, tv);
        }
    }
    if constexpr (pf_slot(B) >= 0) {
        if (live) __stcs(pfp + pf_slot(B) * OD_BLK, W.smpl);
    }
}

// BMASK: bit b-1 set = band b is evaluated by this instantiation (all 16 by default; subsets recompute setcoef)
template <unsigned BMASK, int MINB>
__global__ void __launch_bounds__(TAUMOL_THREADS, MINB) taumol_kernel(ColIn in, Scratch s, RunCfg r) {
    extern __shared__ double fs[];  // NFIELD x TAUMOL_THREADS
    const int tid = threadIdx.x;
    const int l = blockIdx.y;
    const int i_raw = blockIdx.x * TAUMOL_THREADS + tid;
    const bool live = i_raw < r.ncol;
    const int i = live ? i_raw : r.ncol - 1;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long a = ci + in.ld * l;
    const bool dbg = live && (s.dbg_col == i);
    const GTab& g = *D.g;

    // ---- inatm (rad.f90:785-867) for this layer
    const double amd = 28.9660, amw = 18.0160;
    const double stpfac = 296. / 1013.;
    const double pavel = in.play[a], tavel = in.tlay[a];
    const double pz_lo = in.plev[a], pz_hi = in.plev[a + in.ld];
    const double h2o = in.h2o[a], co2 = in.co2[a], o3 = in.o3[a], n2o = in.n2o[a], ch4 = in.ch4[a], o2 = in.o2[a];
    const double amm = (1. - h2o) * amd + h2o * amw;
    const double coldry = (pz_lo - pz_hi) * 1.e3 * C.avogad / (1.e2 * C.grav * amm * (1. + h2o));
    double summol = 0.0;
    summol = summol + co2; summol = summol + o3; summol = summol + n2o; summol = summol + 0.0;
    summol = summol + ch4; summol = summol + o2;
    const double wbrodl = coldry * (1. - summol);
    const double w1 = coldry * h2o, w2 = coldry * co2, w3 = coldry * o3, w4 = coldry * n2o, w6 = coldry * ch4,
                 w7 = coldry * o2;
    double* fq = fs + tid;
    auto put = [&](int f, double v) { fq[f * TAUMOL_THREADS] = v; };
    put(F_WX1, coldry * in.ccl4[a] * 1.e-20);
    put(F_WX2, coldry * in.cfc11[a] * 1.e-20);
    put(F_WX3, coldry * in.cfc12[a] * 1.e-20);
    put(F_WX4, coldry * in.cfc22[a] * 1.e-20);

    // ---- setcoef (setcoef.f90:271-432) for this layer
    const double plog = log(pavel);
    int jp = (int)(36. - 5 * (plog + 0.04));
    jp = jp < 1 ? 1 : (jp > 58 ? 58 : jp);
    const double fp = 5. * (ldg(&g.preflog[jp - 1]) - plog);
    const double tr0 = ldg(&g.tref[jp - 1]), tr1 = ldg(&g.tref[jp]);
    int jt = (int)(3. + (tavel - tr0) / 15.);
    jt = jt < 1 ? 1 : (jt > 4 ? 4 : jt);
    const double ft = ((tavel - tr0) / 15.) - (double)(jt - 3);
    int jt1 = (int)(3. + (tavel - tr1) / 15.);
    jt1 = jt1 < 1 ? 1 : (jt1 > 4 ? 4 : jt1);
    const double ft1 = ((tavel - tr1) / 15.) - (double)(jt1 - 3);
    const double water = w1 / coldry;
    const double scalefac = pavel * stpfac / tavel;
    const double forfac = scalefac / (1. + water);
    const double selffac = water * forfac;
    int indfor, indself = 0;
    double forfrac, selffrac = 0.;
    if (!(plog <= 4.56)) {
        double factor = (332.0 - tavel) / 36.0;
        indfor = min(2, max(1, (int)factor));
        forfrac = factor - (double)indfor;
        factor = (tavel - 188.0) / 7.2;
        indself = min(9, max(1, (int)factor - 7));
        selffrac = factor - (double)(indself + 7);
    } else {
        const double factor = (tavel - 188.0) / 36.0;
        indfor = 3;
        forfrac = factor - 1.0;
    }
    const double scaleminor = pavel / tavel;
    const double scaleminorn2 = (pavel / tavel) * (wbrodl / (coldry + w1));
    const double factor = (tavel - 180.8) / 7.2;
    const int indminor = min(18, max(1, (int)factor));
    const double minorfrac = factor - (double)indminor;
    const double colh2o = 1.e-20 * w1;
    double colco2 = 1.e-20 * w2, colo3 = 1.e-20 * w3, coln2o = 1.e-20 * w4, colch4 = 1.e-20 * w6;
    const double colo2 = 1.e-20 * w7;
    if (colco2 == 0.) colco2 = 1.e-32 * coldry;
    if (colo3 == 0.) colo3 = 1.e-32 * coldry;
    if (coln2o == 0.) coln2o = 1.e-32 * coldry;
    if (colch4 == 0.) colch4 = 1.e-32 * coldry;
    const double compfp = 1. - fp;
    put(F_FAC10, compfp * ft);
    put(F_FAC00, compfp * (1. - ft));
    put(F_FAC11, fp * ft1);
    put(F_FAC01, fp * (1. - ft1));
    put(F_COLH2O, colh2o);
    put(F_COLCO2, colco2);
    put(F_COLO3, colo3);
    put(F_COLN2O, coln2o);
    put(F_COLCH4, colch4);
    put(F_COLO2, colo2);
    put(F_COLBRD, 1.e-20 * wbrodl);
    put(F_COLDRY, coldry);
    put(F_SELFFAC, colh2o * selffac);
    put(F_SELFFRAC, selffrac);
    put(F_FORFAC, colh2o * forfac);
    put(F_FORFRAC, forfrac);
    put(F_MINORFRAC, minorfrac);
    put(F_SCALEMINOR, scaleminor);
    put(F_SCALEMINORN2, scaleminorn2);
    const int ipk = jp | (jt << 6) | (jt1 << 9) | (indself << 12) | (indfor << 16) | (indminor << 18);
    if (dbg) {
        int* q = s.dbg_ints + 6 * l;
        q[0] = jp; q[1] = jt; q[2] = jt1; q[3] = indself; q[4] = indfor; q[5] = indminor;
        double* dq = s.dbg_facs + 4 * l;
        dq[0] = compfp * (1. - ft); dq[1] = fp * (1. - ft1); dq[2] = compfp * ft; dq[3] = fp * ft1;
    }

    // ---- taumol, band by band (loops 1..laytrop / laytrop+1..nlayers, taumol.f90:335, 367)
    const bool lower = l < s.laytrop[i];
#ifndef LW_TAUMOL_NO_SORT
    // The band evaluations are bound by the L1 data pipe: every lane pulls its own k-table rows into registers, and a
    // 256-bit load costs the fewer L1 wavefronts the fewer distinct rows the lanes of an 8-lane group touch.  Columns of
    // a CTA are independent, so the (column -> lane) assignment is free: a counting sort on the (pressure level jp,
    // temperature index jt) pair puts columns that read the same (p, T) corner rows in neighbouring lanes.  Each thread
    // then evaluates the bands of the column it was handed (fields, indices and flags come from shared memory).
    // Third key (LW_TAUMOL_SORT_ETA): the eta bin of the binary-species bands -- H2O / CO2 below the tropopause (bands 3 4 5
    // 12), O3 / CO2 above it (bands 4 5) -- picks the js rows inside a (jp, jt) group of the k-tables; lanes that agree in
    // it touch the same 128-byte lines, and the wavefronts of a warp-wide row load are the distinct lines it touches.
#ifndef LW_TAUMOL_SORT_ETA
#define LW_TAUMOL_SORT_ETA 1
#endif
    constexpr int NBIN = LW_TAUMOL_SORT_ETA ? 128 : 64;
    __shared__ int sh_hist[NBIN];
    __shared__ int sh_perm[TAUMOL_THREADS];
    __shared__ int sh_ipk[TAUMOL_THREADS];
    __shared__ double sh_pavel[TAUMOL_THREADS];
    __shared__ int sh_jpmin;
    for (int k = tid; k < NBIN; k += TAUMOL_THREADS) sh_hist[k] = 0;
    if (tid == 0) sh_jpmin = 64;
    sh_ipk[tid] = ipk | (lower ? (1 << 30) : 0);
    sh_pavel[tid] = pavel;
    __syncthreads();
    atomicMin(&sh_jpmin, jp);
    __syncthreads();
    int bin = (jp - sh_jpmin) * 5 + (jt - 1);
#if LW_TAUMOL_SORT_ETA
    {
        const double a = lower ? colh2o : colo3;
        const double sp = a + ldg(&g.rat[lower ? 0 : 5][jp - 1]) * colco2;
        int eta = (int)((lower ? 8. : 4.) * (a / sp));
        eta = eta < 0 ? 0 : (eta > 7 ? 7 : eta);
        bin = bin * 8 + eta;
    }
#endif
    bin = bin > NBIN - 1 ? NBIN - 1 : bin;
    const int pos = atomicAdd(&sh_hist[bin], 1);
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the bin counts
        constexpr int PER = NBIN / 32;
        int c[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { c[k] = sh_hist[PER * tid + k]; sum += c[k]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= d) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) { sh_hist[PER * tid + k] = run; run += c[k]; }
    }
    __syncthreads();
    sh_perm[sh_hist[bin] + pos] = tid;
    __syncthreads();
    const int src = sh_perm[tid];                       // the column (of this CTA) this thread evaluates
    const int i_src_raw = blockIdx.x * TAUMOL_THREADS + src;
    const bool live_e = i_src_raw < r.ncol;
    const int i_e = live_e ? i_src_raw : r.ncol - 1;
    const int ipk_e = sh_ipk[src] & ~(1 << 30);
    const bool lower_e = (sh_ipk[src] >> 30) & 1;
    const double pavel_e = sh_pavel[src];
    const bool dbg_e = live_e && (s.dbg_col == i_e);
#else
    const int src = tid;
    const bool live_e = live, lower_e = lower, dbg_e = dbg;
    const int i_e = i, ipk_e = ipk;
    const double pavel_e = pavel;
#endif
    const LayerFs L{fs + src};
    const long rec = ((long)blockIdx.x * (TAUMOL_THREADS / OD_BLK) + (src / OD_BLK)) * nlay + l;  // (128-column block, layer)
    double* odp = s.od + rec * NGPT * OD_BLK + 2 * (src % OD_BLK);
    double* pfp = s.pfs + rec * NPF * OD_BLK + (src % OD_BLK);
    const double* sdp = s.secdiff + i_e;
    const double* tap = in.tauaer + (r.c0 + i_e) + in.ld * l;
    const long tas = in.ld * (long)nlay;
#ifdef LW_TAUMOL_BAND_SYNC
#define LW_BAND_SYNC() __syncthreads()  /* keeps the warps of a CTA in the same band (same table rows, same code) */
#else
#define LW_BAND_SYNC()
#endif
#define LW_BAND(B)                                                                                                   \
    if constexpr ((BMASK >> (B - 1)) & 1u) {                                                                         \
        band_odepth<B>(L, ipk_e, lower_e, pavel_e, sdp[(long)(B - 1) * s.pitch], tap[tas * (B - 1)], odp, pfp, live_e, dbg_e, s, l, nlay); \
        LW_BAND_SYNC();                                                                                              \
    }
    LW_BAND(1) LW_BAND(2) LW_BAND(3) LW_BAND(4) LW_BAND(5) LW_BAND(6) LW_BAND(7) LW_BAND(8)
    LW_BAND(9) LW_BAND(10) LW_BAND(11) LW_BAND(12) LW_BAND(13) LW_BAND(14) LW_BAND(15) LW_BAND(16)
#undef LW_BAND
}
// ===== LW_PRECISION_END =====

}  // namespace lw
