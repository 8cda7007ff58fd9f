// lw_rtrnmc.cuh -- phase 2 of the two-phase longwave path: cloud optics + the radiative-transfer sweeps.
//
// What it computes is the reference's cldprmc (src/rrtmg_lw_cldprmc.f90:164-271) and rtrnmc
// (src/rrtmg_lw_rtrnmc.f90:251-563) from the layer optical depths of lw_taumol.cuh; how it computes it is specific to
// this implementation:
//   * one thread owns one column and a GROUP of consecutive spectral bands (<= 32 g-points): lanes are consecutive
//     columns, the vertical recurrences of the group's g-points (radld / radlu, total and clear stream) live in
//     registers, the g-point -> band reduction is a thread-local sum in g order and the band -> broadband reduction a
//     running sum in band order, exactly the reference's summation order (rtrnmc.f90:538-563);
//   * the only per-(g, layer) state is the 8-byte optical depth written by phase 1.  It is streamed once per sweep
//     direction (16-byte L1-bypassing cp.async of a g-point pair into the thread's private slots of a two-stage
//     shared-memory ring, one layer ahead); the upward sweep recomputes the layer transmittance and the Planck source
//     from it with the very same expressions the downward sweep used (same table indices, same rounding) instead of
//     reading back a stored {transmittance, source} pair -- 16 bytes of DRAM traffic per (g, layer) instead of 32;
//   * the Planck fraction of a g-point is rebuilt from the band's reduced fracref table (a g-constant for single-key
//     bands, a two-row interpolation with the eta passed by phase 1 for binary-species bands);
//   * a group's bands share the per-layer temperatures, cloud-mask words, cloud-optics inputs and the
//     read-modify-write of the running level fluxes, so those are touched once per group instead of once per band.
// Everything upstream of the table indices itr / itgas / ittot keeps the reference's association order (-fmad=false);
// downstream of them the clear-layer expressions use explicit fused multiply-adds (lw_band.cuh fmadd).
#pragma once
#include "lw_taumol.cuh"

namespace lw {

// ===== LW_COMMON_BEGIN =====
constexpr int SWEEP_THREADS = 128;
template <int K> using IC = std::integral_constant<int, K>;

// a group of NB consecutive bands starting at band B0 (1-based)
template <int B0, int NB>
struct Grp {
    static constexpr int ngg() { int n = 0; for (int b = B0; b < B0 + NB; ++b) n += NGC[b - 1]; return n; }
    static constexpr int NGG = ngg();
    static constexpr int G0 = GOFF[B0 - 1];
    __host__ __device__ static constexpr int goff(int k) { return GOFF[B0 - 1 + k] - GOFF[B0 - 1]; }
    static constexpr int npf() { int n = 0; for (int b = B0; b < B0 + NB; ++b) n += pf_slot(b) >= 0; return n; }
    static constexpr int NPFG = npf();
    static constexpr int pf0() { for (int b = B0; b < B0 + NB; ++b) if (pf_slot(b) >= 0) return pf_slot(b); return 0; }
    static constexpr int PF0 = pf0();
    static_assert(NGG <= 32, "a group's cloud bits must fit two mask words");
    static_assert(NGG % 2 == 0, "optical depths are stored as g-point pairs");
};
// ===== LW_COMMON_END =====

// ===== LW_PRECISION_BEGIN =====
// shared-memory ring of one sweep kernel: rows of SWEEP_THREADS doubles per stage.  Rows [0, NGG): the optical depths as
// NGG / 2 pair rows (thread t owns doubles [2 t, 2 t + 1] of a pair row); then one row each for tlay, tlev, the group's
// Planck-fraction etas, two rows of 4-byte words (icldlyr word, mask word lo, mask word hi) and four cloud-optics inputs.
template <int B0, int NB>
struct SweepRing {
    using G = Grp<B0, NB>;
    static constexpr int IPS = (int)(sizeof(double) / sizeof(int));
    static constexpr int S_TLAY = G::NGG, S_TLEV = G::NGG + 1, S_PF = G::NGG + 2;
    static constexpr int S_INT = S_PF + G::NPFG;
    static constexpr int S_CLD = S_INT + 4 / IPS;
    static constexpr int SLOTS = S_CLD + 4;
    static constexpr int NST = 2;
    static constexpr size_t BYTES = (size_t)NST * SLOTS * SWEEP_THREADS * sizeof(double);
};

// Planck function of band B at the temperature whose table position is (it, frac), setcoef.f90:173-226
template <int B>
__device__ __forceinline__ double planck_at(int it, double frac) {
    const double* tp = D.totplnk + 181 * (B - 1);
    const double a = ldg(tp + it - 1), b = ldg(tp + it);
    const double dbdt = b - a;
    return fmadd(frac, dbdt, a);
}
__device__ __forceinline__ void planck_pos(double t, int& it, double& frac) {
    it = (int)(t - 159.);
    it = it < 1 ? 1 : (it > 180 ? 180 : it);
    frac = t - 159. - (double)it;
}

// Planck fractions fracs(lay, g) of band B, rebuilt from the reduced fracref tables exactly as taugbB forms them
// (e.g. taumol.f90:556-561, 692-693 for the eta-interpolated bands; a g-constant otherwise); sm = specmult of the
// Planck-fraction interpolation passed by phase 1 (jpl = 1 + int(sm), fpl = mod(sm, 1))
template <int B>
__device__ __forceinline__ void planck_fracs(bool lower, double sm, double* pf) {
    constexpr int NG = NGC[B - 1], RS = NGPAD[B - 1];
    constexpr bool BIN_LO = NSPA[B - 1] == 9, BIN_UP = NSPB[B - 1] == 5;
    const BandTab& T = C.bandp[B - 1];
    double buf[RS];
    auto rows2 = [&](const double* tab) {
        const int jpl = 1 + (int)sm;
        const double fpl = sm - (double)(int)sm;
        const double* p = tab + (jpl - 1) * RS;
#pragma unroll
        for (int g = 0; g < RS; g += 4) {
            const D4 p0 = ldv<D4>(p + g), p1 = ldv<D4>(p + RS + g);
            unpack(p0 + fpl * (p1 - p0), buf + g);
        }
    };
    auto row1 = [&](const double* tab) {
#pragma unroll
        for (int g = 0; g < RS; g += 4) unpack(ldv<D4>(tab + g), buf + g);
    };
    if (lower) {
        if constexpr (BIN_LO) rows2(T.fracrefa); else row1(T.fracrefa);
    } else {
        if constexpr (BIN_UP) rows2(T.fracrefb);
        else if constexpr (B == 6) row1(T.fracrefa);
        else if constexpr (B == 12 || B == 15) {
#pragma unroll
            for (int g = 0; g < RS; ++g) buf[g] = 0.0;
        } else row1(T.fracrefb);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) pf[g] = buf[g];
}

// cldprmc for one (sub-column, layer) of band B: cloud optical depth from the water paths and effective radii
// (cldprmc.f90:177-266); tauc_in is the caller's value, returned unchanged unless inflag = 2 finds cloud water
template <int B>
__device__ __forceinline__ double cld_tauc(double ciwp, double clwp, double radice, double radliq, double tauc_in,
                                           const RunCfg& r) {
    double tauc = tauc_in;
    const double cwp = ciwp + clwp;
    if (r.inflag == 2 && (cwp >= 1.e-20 || tauc >= 1.e-20)) {
        double abscoice = 0.0, abscoliq = 0.0;
        if (ciwp != 0.0) {
            if (r.iceflag == 0) {
                abscoice = C.absice0[0] + C.absice0[1] / radice;
            } else if (r.iceflag == 1) {
                constexpr int icb[16] = {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
                abscoice = C.absice1[0 + 2 * (icb[B - 1] - 1)] + C.absice1[1 + 2 * (icb[B - 1] - 1)] / radice;
            } else {
                const double factor = (radice - 2.) / 3.;
                int index = (int)factor;
                const int nrad = r.iceflag == 2 ? 43 : 46;
                if (index == nrad) index = nrad - 1;
                if (index < 1) index = 1;        // unreachable for valid radii; keeps loads in bounds
                if (index > nrad - 1) index = nrad - 1;
                const double fint = factor - (double)index;
                const double* t = (r.iceflag == 2 ? D.absice2 : D.absice3) + nrad * (B - 1);
                const double x0 = ldg(t + index - 1), x1 = ldg(t + index);
                abscoice = x0 + fint * (x1 - (x0));
            }
        }
        if (clwp != 0.0) {
            if (r.liqflag == 0) {
                abscoliq = C.absliq0;
            } else {
                int index = (int)(radliq - 1.5);
                if (index == 0) index = 1;
                if (index == 58) index = 57;
                if (index < 1) index = 1;
                if (index > 57) index = 57;
                const double fint = radliq - 1.5 - (double)index;
                const double* t = D.absliq1 + 58 * (B - 1);
                const double x0 = ldg(t + index - 1), x1 = ldg(t + index);
                abscoliq = x0 + fint * (x1 - (x0));
            }
        }
        tauc = ciwp * abscoice + clwp * abscoliq;
    }
    return tauc;
}

// One g-point of a layer that holds cloud in some sub-column (icldlyr = 1), rtrnmc.f90:358-425.  dpl = dplankdn on the
// way down, dplankup on the way up; bb_gas / bb_tot are the Planck sources of the gas-only and the gas + cloud path
// (bbd / bbdtot down, bbugas / bbutot up); gassrc is formed only on the way down, with the reference's association.
// itr = int(tblint * tblind + 0.5) (rtrnmc.f90:386, 404, 412), kept inside the table for out-of-range cloud inputs
// (a negative cloud optical depth would index the reference's tables out of bounds; here it must not fault the device)
__device__ __forceinline__ int tbl_index(double tblind) {
    const int it = (int)(10000.0 * tblind + 0.5);
    return it < 0 ? 0 : (it > NTBL ? NTBL : it);
}

template <bool DOWN>
__device__ __forceinline__ void cloudy_point(double odepth, double odc, double plfrac, double blay, double dpl,
                                             double bpade, double& atrans, double& atot, double& bb_gas, double& bb_tot,
                                             double& gassrc, int& d0, int& d1) {
    const double rec_6 = 0.166667;
    double odtot = odepth + odc;
    d0 = -1; d1 = -1;
    gassrc = 0.;
    if (odtot < 0.06) {
        atrans = odepth - 0.5 * odepth * odepth;
        const double odepth_rec = rec_6 * odepth;
        if (DOWN) gassrc = plfrac * fmadd(dpl, odepth_rec, blay) * atrans;
        atot = odtot - 0.5 * odtot * odtot;
        const double odtot_rec = rec_6 * odtot;
        bb_tot = plfrac * fmadd(dpl, odtot_rec, blay);
        bb_gas = plfrac * fmadd(dpl, odepth_rec, blay);
    } else if (odepth <= 0.06) {
        atrans = odepth - 0.5 * odepth * odepth;
        const double odepth_rec = rec_6 * odepth;
        if (DOWN) gassrc = plfrac * fmadd(dpl, odepth_rec, blay) * atrans;
        odtot = odepth + odc;
        const double tblind = odtot / (bpade + odtot);
        const int ittot = tbl_index(tblind);
        d1 = ittot;
        const double2 et = __ldg(D.exptfn + ittot);
        const double tfactot = et.y;
        bb_tot = plfrac * fmadd(tfactot, dpl, blay);
        bb_gas = plfrac * fmadd(dpl, odepth_rec, blay);
        atot = 1. - et.x;
    } else {
        double tblind = odepth / (bpade + odepth);
        const int itgas = tbl_index(tblind);
        d0 = itgas;
        odepth = ldg(D.tau_tbl + itgas);
        const double2 eg = __ldg(D.exptfn + itgas);
        atrans = 1. - eg.x;
        const double tfacgas = eg.y;
        if (DOWN) gassrc = atrans * plfrac * fmadd(tfacgas, dpl, blay);
        odtot = odepth + odc;
        tblind = odtot / (bpade + odtot);
        const int ittot = tbl_index(tblind);
        d1 = ittot;
        const double2 et = __ldg(D.exptfn + ittot);
        const double tfactot = et.y;
        bb_tot = plfrac * fmadd(tfactot, dpl, blay);
        bb_gas = plfrac * fmadd(tfacgas, dpl, blay);
        atot = 1. - et.x;
    }
}

// One g-point of a cloud-free layer, rtrnmc.f90:426-445, as selects instead of branches: both the small-optical-depth
// series and the table lookup are evaluated and one is chosen, so that the g-points of a thread form independent
// instruction streams.  bb = plfrac * (blay + tausfac * dpl) is bbd on the way down and bbugas on the way up.
__device__ __forceinline__ void clear_point(double odepth, double plfrac, double blay, double dpl, double bpade,
                                            double& atrans, double& bb, int& itr_out) {
    const double rec_6 = 0.166667;
    const bool small = odepth <= 0.06;
    const double tblind = div_fast(odepth, bpade + odepth);
    const int itr = (int)(10000.0 * tblind + 0.5);
    const double2 e = __ldg(D.exptfn + itr);
#ifdef LW_TAIL_FMA
    atrans = small ? fmadd(-0.5 * odepth, odepth, odepth) : 1. - e.x;
#else
    atrans = small ? odepth - 0.5 * odepth * odepth : 1. - e.x;
#endif
    const double tausfac = small ? rec_6 * odepth : e.y;
    bb = plfrac * fmadd(tausfac, dpl, blay);
    itr_out = small ? -1 : itr;
}

// ------------------------------------------------------------------------------------------
// the sweep kernel of bands B0 .. B0 + NB - 1
template <int B0, int NB, int CLD, int IDRV, int MINB>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB) rtrnmc_kernel(ColIn in, ColOut out, Scratch s, RunCfg r) {
    using GR = Grp<B0, NB>;
    using RG = SweepRing<B0, NB>;
    constexpr int NGG = GR::NGG;
    constexpr int G0 = GR::G0;
    extern __shared__ double ring[];
    const int tid = threadIdx.x;
    const int i_raw = blockIdx.x * blockDim.x + tid;
    // out-of-range threads of the last CTA shadow the last valid column (no early return: keeps all lanes of a warp alive)
    const bool live = i_raw < r.ncol;
    const int i = live ? i_raw : r.ncol - 1;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const long plane = (long)nlay * P;
    const double bpade = C.bpade;
    const bool dbg = live && (s.dbg_col == i);
    const bool first = r.first_band != 0;  // this launch opens the band sum (band istart of the run)
    const int laytrop = s.laytrop[i];
    const int ctop = (CLD != CLD_NONE) ? s.ctop[i] : 0;
    (void)out;

    // blocks of 128 columns, [column / 128][layer][...][column % 128]
    const long blk_lay0 = (long)(i >> 7) * nlay;
    const int blk_t = i & 127;
    auto od_src = [&](int l0) { return reinterpret_cast<const double2*>(s.od + ((blk_lay0 + l0) * NGPT + G0) * 128) + blk_t; };
    auto pf_src = [&](int l0) { return s.pfs + ((blk_lay0 + l0) * NPF + GR::PF0) * 128 + blk_t; };
    // {odcld, 1 - exp(-odcld)} of a (layer, band) whose sub-columns hold cloud (CLD_MASK): [block][layer][band][2][128]
    auto cldb = [&](int l0, int b) { return s.cldb + (((blk_lay0 + l0) * NBND + (b - 1)) * 2) * 128 + blk_t; };
    const long blk_lev0 = (long)(i >> 7) * (nlay + 1);
    // running band sums of the level fluxes: [column / 128][level][6][column % 128]; 0 dflx, 1 dflxc, 2 uflx, 3 uflxc,
    // 4 duflx_dt, 5 duflxc_dt (the finalize kernel turns them into the caller's arrays)
    auto flx = [&](int lev) { return s.fluxB + ((blk_lev0 + lev) * 6) * 128 + blk_t; };
    // strict rrtmg_lw(): per-g {efclfrac, odcld} of cloudy layers
    double* S_EF = s.gcld;
    double* S_OC = s.gcld + (long)NGPT * plane;

    // ---- asynchronous prefetch of one layer's inputs into this thread's slots of ring stage `stage`.
    // DOWN: tlev = level below the layer (planklev(lev-1)), cloud-optics inputs ride along; UP: tlev = level above.
    auto issue = [&](int stage, int lev, bool down) {
        if (lev >= 1 && lev <= nlay) {
            const int l0 = lev - 1;
            const long a = ci + in.ld * l0;
            double* dst = ring + (size_t)stage * RG::SLOTS * SWEEP_THREADS;
            const double2* od = od_src(l0);
#pragma unroll
            for (int p = 0; p < NGG / 2; ++p) cp_async_pair(dst + (2 * p) * SWEEP_THREADS + 2 * tid, od + p * 128);
            cp_async_val(dst + RG::S_TLAY * SWEEP_THREADS + tid, in.tlay + a);
            cp_async_val(dst + RG::S_TLEV * SWEEP_THREADS + tid, in.tlev + (down ? a : a + in.ld));
            if constexpr (GR::NPFG > 0) {
                const double* pf = pf_src(l0);
#pragma unroll
                for (int k = 0; k < GR::NPFG; ++k) cp_async_val(dst + (RG::S_PF + k) * SWEEP_THREADS + tid, pf + k * 128);
            }
            if constexpr (CLD != CLD_NONE) {
                auto di = [&](int k) { return reinterpret_cast<int*>(dst + (RG::S_INT + k / RG::IPS) * SWEEP_THREADS + tid) + (k % RG::IPS); };
                const unsigned* mw = s.mask + ((long)l0 * 5) * P + i;
                cp_async4(di(0), mw + 4 * P);
                if constexpr (CLD == CLD_MASK) {
                    constexpr int wlo = G0 >> 5, sh = G0 & 31;
                    cp_async4(di(1), mw + (long)wlo * P);
                    if constexpr (sh + NGG > 32) cp_async4(di(2), mw + (long)(wlo + 1) * P);
                    if (down && lev <= ctop) {  // cloud-optics inputs: only layers at or below the column's cloud top can need them
                        cp_async_val(dst + (RG::S_CLD + 0) * SWEEP_THREADS + tid, in.ciwp + a);
                        cp_async_val(dst + (RG::S_CLD + 1) * SWEEP_THREADS + tid, in.clwp + a);
                        cp_async_val(dst + (RG::S_CLD + 2) * SWEEP_THREADS + tid, in.rei + a);
                        cp_async_val(dst + (RG::S_CLD + 3) * SWEEP_THREADS + tid, in.rel + a);
                    }
                }
            }
        }
        cp_commit();
    };
    // views of a stage
    auto st_d = [&](int stage, int slot) { return ring[((size_t)stage * RG::SLOTS + slot) * SWEEP_THREADS + tid]; };
    auto st_i = [&](int stage, int k) {
        return reinterpret_cast<const int*>(ring + ((size_t)stage * RG::SLOTS + RG::S_INT + k / RG::IPS) * SWEEP_THREADS + tid)[k % RG::IPS];
    };
    auto st_od = [&](int stage, int p) {
        return *reinterpret_cast<const double2*>(ring + ((size_t)stage * RG::SLOTS + 2 * p) * SWEEP_THREADS + 2 * tid);
    };
    // the group's cloud bits of the layer in `stage` (bit k = g-point G0 + k)
    auto group_bits = [&](int stage) -> unsigned {
        constexpr int sh = G0 & 31;
        unsigned lo = ((unsigned)st_i(stage, 1)) >> sh;
        if constexpr (sh + NGG > 32) lo |= ((unsigned)st_i(stage, 2)) << (32 - sh);
        return NGG == 32 ? lo : (lo & ((1u << (NGG & 31)) - 1u));
    };
    auto for_bands = [&](auto&& f) {
        f(IC<0>{});
        if constexpr (NB > 1) f(IC<1>{});
        if constexpr (NB > 2) f(IC<2>{});
        if constexpr (NB > 3) f(IC<3>{});
        if constexpr (NB > 4) f(IC<4>{});
        if constexpr (NB > 5) f(IC<5>{});
        if constexpr (NB > 6) f(IC<6>{});
        if constexpr (NB > 7) f(IC<7>{});
        static_assert(NB <= 8, "at most 8 bands per group");
    };

    acc_t radld[NGG], radclrd[NGG];
#pragma unroll
    for (int g = 0; g < NGG; ++g) { radld[g] = 0.; radclrd[g] = 0.; }

    // ---------------------------------------------------------------- downward sweep, rtrnmc.f90:350-455
    issue(0, nlay, true);
    issue(1, nlay - 1, true);
    int stage = 0;
    for (int lev = nlay; lev >= 1; --lev) {
        const int l0 = lev - 1;
        const long idx = (long)l0 * P + i;
        const long a = ci + in.ld * l0;
        cp_wait<RG::NST - 1>();
        // running band sums of this level, fetched now so their latency hides behind the g loops
        acc_t acc_d = 0., acc_dc = 0.;
        if (!first && live) {
            acc_d = __ldcg(flx(l0));
            if (CLD != CLD_NONE && ctop > 0) acc_dc = __ldcg(flx(l0) + 128);
        }
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(st_d(stage, RG::S_TLAY), it_lay, fr_lay);
        planck_pos(st_d(stage, RG::S_TLEV), it_lev, fr_lev);
        const bool lower = lev <= laytrop;
        bool cldlyr = false;
        unsigned gbits = 0;
        if constexpr (CLD != CLD_NONE) cldlyr = (((unsigned)st_i(stage, 0)) >> 31) != 0;
        if constexpr (CLD == CLD_MASK) { if (cldlyr) gbits = group_bits(stage); }

        for_bands([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int B = B0 + K;
            constexpr int NG = NGC[B - 1];
            constexpr int GO = GR::goff(K);
            const double blay = planck_at<B>(it_lay, fr_lay);
            const double dplankdn = planck_at<B>(it_lev, fr_lev) - blay;  // planklev(lev-1) - planklay(lev)
            double pf[NG];
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = st_d(stage, RG::S_PF + pf_slot(B) - GR::PF0);
            planck_fracs<B>(lower, sm, pf);
            // cloud state of this (column, layer) for this band
            unsigned bits = 0;
            double odcld_b = 0., abscld_b = 0.;
            if constexpr (CLD == CLD_MASK) {
                if (cldlyr) {
                    bits = (gbits >> GO) & ((1u << NG) - 1u);
                    if (bits) {
                        const double tauc_in = in.tauc ? in.tauc[(B - 1) + 16 * a] : 0.0;
                        const double tauc = cld_tauc<B>(st_d(stage, RG::S_CLD), st_d(stage, RG::S_CLD + 1), st_d(stage, RG::S_CLD + 2),
                                                        st_d(stage, RG::S_CLD + 3), tauc_in, r);
                        odcld_b = ldg(s.secdiff + (long)(B - 1) * P + i) * tauc;
                        abscld_b = 1. - exp(-odcld_b);
                        if (live) { double* q = cldb(l0, B); q[0] = odcld_b; q[128] = abscld_b; }
                    }
                }
            }
            acc_t dsum = 0., dcsum = 0.;
            if (CLD != CLD_NONE && cldlyr) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double odepth = ((GO + g) & 1) ? odp.y : odp.x;
                    double cf, odc, ef;
                    if constexpr (CLD == CLD_MASK) {
                        const bool on = (bits >> g) & 1u;
                        cf = on ? 1. : 0.; odc = on ? odcld_b : 0.; ef = on ? abscld_b : 0.;
                    } else {
                        // strict rrtmg_lw(): per-g cloud arrays, cldprmc evaluated on the fly
                        const long ag = (G0 + GO + g) + 140 * a;
                        cf = in.cldfmcl[ag];
                        double tauc = in.taucmcl[ag];
                        if (cf >= 1.e-20) tauc = cld_tauc<B>(in.ciwpmcl[ag], in.clwpmcl[ag], in.reicmcl[a], in.relqmcl[a], tauc, r);
                        if (cf == 1.) {
                            odc = ldg(s.secdiff + (long)(B - 1) * P + i) * tauc;
                            ef = (1. - exp(-odc)) * cf;
                        } else {
                            odc = 0.; ef = 0.;
                        }
                        if (live) { S_EF[(long)(G0 + GO + g) * plane + idx] = ef; S_OC[(long)(G0 + GO + g) * plane + idx] = odc; }
                    }
                    double atrans, atot, bbd, bbdtot, gassrc;
                    int d0, d1;
                    cloudy_point<true>(odepth, odc, pf[g], blay, dplankdn, bpade, atrans, atot, bbd, bbdtot, gassrc, d0, d1);
                    radld[GO + g] = relaxed_cld(radld[GO + g], atrans + ef * (1. - atrans), gassrc, cf * (bbdtot * atot - gassrc));
                    dsum = dsum + radld[GO + g];
                    // clear-sky stream: identical to the total stream until the first cloudy layer from the top
                    const acc_t sep = relaxed(radclrd[GO + g], bbd, atrans);
                    radclrd[GO + g] = (lev <= ctop) ? sep : radld[GO + g];
                    dcsum = dcsum + radclrd[GO + g];
                    if (dbg) {
                        const long o = (long)(G0 + GO + g) * nlay + l0;
                        s.dbg_fracs[o] = pf[g];
                        s.dbg_itr[2 * o] = d0; s.dbg_itr[2 * o + 1] = d1;
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double odepth = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbd;
                    int itr;
                    clear_point(odepth, pf[g], blay, dplankdn, bpade, atrans, bbd, itr);
                    radld[GO + g] = relaxed(radld[GO + g], bbd, atrans);
                    dsum = dsum + radld[GO + g];
                    if (CLD != CLD_NONE) {
                        const acc_t sep = relaxed(radclrd[GO + g], bbd, atrans);
                        radclrd[GO + g] = (lev <= ctop) ? sep : radld[GO + g];
                        dcsum = dcsum + radclrd[GO + g];
                    }
                    if (dbg) {
                        const long o = (long)(G0 + GO + g) * nlay + l0;
                        s.dbg_fracs[o] = pf[g];
                        s.dbg_itr[2 * o] = itr; s.dbg_itr[2 * o + 1] = -1;
                    }
                }
            }
            // band close-out for level lev-1: flux += (rad * wtdiff) * delwave, rtrnmc.f90:538-550
            acc_d = acc_d + (dsum * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
            if (CLD != CLD_NONE) acc_dc = acc_dc + ((lev <= ctop ? dcsum : dsum) * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
        });
        if (live) {
            acc_t* o = flx(l0);
            o[0] = acc_d;
            if (CLD != CLD_NONE && ctop > 0) o[128] = acc_dc;
        }
        issue(stage, lev - RG::NST, true);
        stage ^= 1;
    }
    cp_wait<0>();

    // ---------------------------------------------------------------- surface, rtrnmc.f90:465-484
    // radld / radclrd are dead after the sweep; they become the upward radiances
    acc_t (&radlu)[NGG] = radld;
    acc_t (&radclru)[NGG] = radclrd;
    acc_t d_radlu[IDRV ? NGG : 1], d_radclru[IDRV ? NGG : 1];
    issue(0, 1, false);
    issue(1, 2, false);
    {
        acc_t acc_u = 0., acc_uc = 0., acc_du = 0., acc_duc = 0.;
        if (first) {
            // downward flux at the top of the atmosphere is never touched by the sweep: drad(nlayers) = 0
            if (live) {
                flx(nlay)[0] = 0.0;
                if (CLD != CLD_NONE && ctop > 0) flx(nlay)[128] = 0.0;
            }
        } else if (live) {
            const acc_t* o = flx(0);
            acc_u = __ldcg(o + 256);
            if (CLD != CLD_NONE && ctop > 0) acc_uc = __ldcg(o + 384);
            if constexpr (IDRV == 1) {
                acc_du = __ldcg(o + 512);
                if (CLD != CLD_NONE && ctop > 0) acc_duc = __ldcg(o + 640);
            }
        }
        const double tbound = in.tsfc[ci];
        int it_b;
        double fr_b;
        planck_pos(tbound, it_b, fr_b);
        const bool lower1 = 1 <= laytrop;
        for_bands([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int B = B0 + K;
            constexpr int NG = NGC[B - 1];
            constexpr int GO = GR::goff(K);
            // surface emission, setcoef.f90:211-219 / rtrnmc.f90:465-469
            const double semiss = in.emis[ci + in.ld * (B - 1)];
            const double* tp = D.totplnk + 181 * (B - 1);
            const double pa = ldg(tp + it_b - 1), pb = ldg(tp + it_b);
            const double plankbnd = semiss * (pa + fr_b * (pb - pa));
            double dplankbnd_dt = 0.;
            if constexpr (IDRV == 1) {
                const double* td = D.totplnkderiv + 181 * (B - 1);
                const double da = ldg(td + it_b - 1), db = ldg(td + it_b);
                dplankbnd_dt = semiss * (da + fr_b * (db - da));
            }
            const double reflect = 1. - semiss;
            double pf[NG];
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = pf_src(0)[(pf_slot(B) - GR::PF0) * 128];
            planck_fracs<B>(lower1, sm, pf);  // fracs(1, g)
            acc_t usum = 0., ucsum = 0., dusum = 0.;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double rad0 = pf[g] * plankbnd;
                radlu[GO + g] = rad0 + reflect * radld[GO + g];
                radclru[GO + g] = rad0 + reflect * radclrd[GO + g];
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                usum = usum + radlu[GO + g];
                ucsum = ucsum + radclru[GO + g];
            }
            if constexpr (IDRV == 1) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    d_radlu[GO + g] = pf[g] * dplankbnd_dt;
                    d_radclru[GO + g] = d_radlu[GO + g];
                    dusum = dusum + d_radlu[GO + g];
                }
            }
            acc_u = acc_u + (usum * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
            acc_uc = acc_uc + (ucsum * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
            if constexpr (IDRV == 1) {
                const acc_t addd = (dusum * (acc_t)0.5) * (acc_t)C.delwave[B - 1] * (acc_t)C.fluxfac;
                acc_du = acc_du + addd;
                acc_duc = acc_duc + addd;
            }
        });
        if (live) {
            acc_t* o = flx(0);
            o[256] = acc_u;
            if (CLD != CLD_NONE && ctop > 0) o[384] = acc_uc;
            if constexpr (IDRV == 1) {
                o[512] = acc_du;
                if (CLD != CLD_NONE && ctop > 0) o[640] = acc_duc;
            }
        }
    }

    // ---------------------------------------------------------------- upward sweep, rtrnmc.f90:486-529
    stage = 0;
    for (int lev = 1; lev <= nlay; ++lev) {
        const int l0 = lev - 1;
        const long idx = (long)l0 * P + i;
        const long a = ci + in.ld * l0;
        cp_wait<RG::NST - 1>();
        acc_t acc_u = 0., acc_uc = 0., acc_du = 0., acc_duc = 0.;
        if (!first && live) {
            const acc_t* o = flx(lev);
            acc_u = __ldcg(o + 256);
            if (CLD != CLD_NONE && ctop > 0) acc_uc = __ldcg(o + 384);
            if constexpr (IDRV == 1) {
                acc_du = __ldcg(o + 512);
                if (CLD != CLD_NONE && ctop > 0) acc_duc = __ldcg(o + 640);
            }
        }
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(st_d(stage, RG::S_TLAY), it_lay, fr_lay);
        planck_pos(st_d(stage, RG::S_TLEV), it_lev, fr_lev);
        const bool lower = lev <= laytrop;
        bool cldlyr = false;
        unsigned gbits = 0;
        if constexpr (CLD != CLD_NONE) cldlyr = (((unsigned)st_i(stage, 0)) >> 31) != 0;
        if constexpr (CLD == CLD_MASK) { if (cldlyr) gbits = group_bits(stage); }

        for_bands([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int B = B0 + K;
            constexpr int NG = NGC[B - 1];
            constexpr int GO = GR::goff(K);
            const double blay = planck_at<B>(it_lay, fr_lay);
            const double dplankup = planck_at<B>(it_lev, fr_lev) - blay;  // planklev(lev) - planklay(lev)
            double pf[NG];
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = st_d(stage, RG::S_PF + pf_slot(B) - GR::PF0);
            planck_fracs<B>(lower, sm, pf);
            unsigned bits = 0;
            double odcld_b = 0., abscld_b = 0.;
            if constexpr (CLD == CLD_MASK) {
                if (cldlyr) {
                    bits = (gbits >> GO) & ((1u << NG) - 1u);
                    if (bits) { const double* q = cldb(l0, B); odcld_b = __ldcg(q); abscld_b = __ldcg(q + 128); }
                }
            }
            acc_t usum = 0., ucsum = 0., dusum = 0., ducsum = 0.;
            if (CLD != CLD_NONE && cldlyr) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double odepth = ((GO + g) & 1) ? odp.y : odp.x;
                    double cf, odc, ef;
                    if constexpr (CLD == CLD_MASK) {
                        const bool on = (bits >> g) & 1u;
                        cf = on ? 1. : 0.; odc = on ? odcld_b : 0.; ef = on ? abscld_b : 0.;
                    } else {
                        cf = in.cldfmcl[(G0 + GO + g) + 140 * a];
                        ef = S_EF[(long)(G0 + GO + g) * plane + idx];
                        odc = S_OC[(long)(G0 + GO + g) * plane + idx];
                    }
                    double atrans, atot, bbugas, bbutot, unused;
                    int d0, d1;
                    cloudy_point<false>(odepth, odc, pf[g], blay, dplankup, bpade, atrans, atot, bbugas, bbutot, unused, d0, d1);
                    const double gassrc = bbugas * atrans;
                    radlu[GO + g] = relaxed_cld(radlu[GO + g], atrans + ef * (1. - atrans), gassrc, cf * (bbutot * atot - gassrc));
                    if constexpr (IDRV == 1)
                        d_radlu[GO + g] = d_radlu[GO + g] * cf * (1.0 - atot) + d_radlu[GO + g] * (1.0 - cf) * (1.0 - atrans);
                    usum = usum + radlu[GO + g];
                    if constexpr (IDRV == 1) dusum = dusum + d_radlu[GO + g];
                    radclru[GO + g] = relaxed(radclru[GO + g], bbugas, atrans);
                    ucsum = ucsum + radclru[GO + g];
                    if constexpr (IDRV == 1) {
                        d_radclru[GO + g] = d_radclru[GO + g] * (1.0 - atrans);
                        ducsum = ducsum + d_radclru[GO + g];
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double odepth = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbugas;
                    int itr;
                    clear_point(odepth, pf[g], blay, dplankup, bpade, atrans, bbugas, itr);
                    radlu[GO + g] = relaxed(radlu[GO + g], bbugas, atrans);
                    usum = usum + radlu[GO + g];
                    if constexpr (IDRV == 1) {
                        d_radlu[GO + g] = d_radlu[GO + g] * (1.0 - atrans);
                        dusum = dusum + d_radlu[GO + g];
                    }
                    if constexpr (CLD != CLD_NONE) {
                        radclru[GO + g] = relaxed(radclru[GO + g], bbugas, atrans);
                        ucsum = ucsum + radclru[GO + g];
                        if constexpr (IDRV == 1) {
                            d_radclru[GO + g] = d_radclru[GO + g] * (1.0 - atrans);
                            ducsum = ducsum + d_radclru[GO + g];
                        }
                    }
                }
            }
            acc_u = acc_u + (usum * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
            if (CLD != CLD_NONE) acc_uc = acc_uc + (ucsum * (acc_t)0.5) * (acc_t)C.delwave[B - 1];
            if constexpr (IDRV == 1) {
                acc_du = acc_du + (dusum * (acc_t)0.5) * (acc_t)C.delwave[B - 1] * (acc_t)C.fluxfac;
                if (CLD != CLD_NONE) acc_duc = acc_duc + (ducsum * (acc_t)0.5) * (acc_t)C.delwave[B - 1] * (acc_t)C.fluxfac;
            }
        });
        if (live) {
            acc_t* o = flx(lev);
            o[256] = acc_u;
            if (CLD != CLD_NONE && ctop > 0) o[384] = acc_uc;
            if constexpr (IDRV == 1) {
                o[512] = acc_du;
                if (CLD != CLD_NONE && ctop > 0) o[640] = acc_duc;
            }
        }
        issue(stage, lev + RG::NST, false);
        stage ^= 1;
    }
    cp_wait<0>();
}
// ===== LW_PRECISION_END =====

}  // namespace lw
