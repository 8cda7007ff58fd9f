// lw_gridcloud.cuh -- the NON-McICA cloudy path of the reference's second GCM entry (src/rrtmg_lw_rad.nomcica.f90):
// grid-mean cloud properties, band cloud optical depths from cldprop (src/rrtmg_lw_cldprop.f90:166-295), radiative
// transfer with random cloud overlap (rtrn, src/rrtmg_lw_rtrn.f90:296-560, icld = 1) or maximum-random overlap
// (rtrnmr, src/rrtmg_lw_rtrnmr.f90:296-806, icld = 2 / 3 / 0).
//
// It rides on the two-phase path: inatm + setcoef + taumol are the same (column, layer) kernel (lw_taumol.cuh), whose
// 8-byte state words carry the gas optical depth / table index of every (g, layer).  What is specific to this path:
//   * gridcloud_prep_kernel, one thread per column: cldprop for every layer (the band count ncbands it leaves behind
//     is the value after the LAST cloudy layer, a column scalar that selects the band -> cloud-band map ipat), the
//     per-band cloud optical depth along the diffusivity angle, and the layer-to-layer overlap factors of rtrnmr
//     (two sequential passes with the rat1 / rat2 memory of the reference, rtrnmr.f90:345-505);
//   * gridcloud_sweep_kernel<B, MAXRAN, IDRV>, one thread per (column, band): the sweeps, with the cloudy / clear
//     stream split and the radmod / rad exchange terms of the maximum-random scheme carried per g-point in registers.
// This is a compatibility path (it exists so that every entry point of the reference has a drop-in, and for the one
// shipped example that uses it, run_examples_std_atm/input_rrtm_MLS-cld5-imca0-icld2): plain loads, no prefetch ring.
#pragma once
#include "lw_rtrnmc.cuh"

namespace lw {


__device__ __forceinline__ void gc_raise(const Scratch& s, int code, long col, int lay) {
    const unsigned long long v = ((unsigned long long)code << 56) | ((unsigned long long)(col + s.err_base + 1) << 24) |
                                 (unsigned long long)(lay & 0xffffff);
    atomicCAS(s.status, 0ull, v);
}

static __global__ void __launch_bounds__(128) gridcloud_prep_kernel(ColIn in, Scratch s, RunCfg r, int maxran) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    auto REC = [&](int l0, int k) -> double& { return s.gridcld[((long)l0 * GC_NREC + k) * P + i]; };
    const int icb[3][16] = {{1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                            {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                            {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
    const double cldmin = 1.e-20;
    int ncbands = 1;
    const bool cloudy = r.icld >= 1;
    // ---- cldprop: taucloud(lay, 1..16) parked in the GC_ODCLD slots until ncbands is final
    for (int l = 0; l < nlay; ++l) {
        const long a = ci + in.ld * l;
        const double cf = cloudy ? in.cldfrac[a] : 0., ciwp = cloudy ? in.ciwp[a] : 0., clwp = cloudy ? in.clwp[a] : 0.;
        double tauctot = 0.;
        if (cloudy)
            for (int ib = 0; ib < 16; ++ib) tauctot = tauctot + in.tauc[ib + 16 * a];
        double tc[16];
#pragma unroll
        for (int ib = 0; ib < 16; ++ib) tc[ib] = 0.;
        const double cwp = ciwp + clwp;
        if (cf >= cldmin && (cwp >= cldmin || tauctot >= cldmin)) {
            if (r.inflag == 0) {
                ncbands = 16;
                for (int ib = 0; ib < 16; ++ib) tc[ib] = in.tauc[ib + 16 * a];
            } else if (r.inflag == 1) {
                ncbands = 16;
                for (int ib = 0; ib < 16; ++ib) tc[ib] = C.abscld1 * cwp;
            } else if (r.inflag == 2) {
                const double radice = in.rei[a];
                double abscoice[16], abscoliq[16];
                int iceind = 0, liqind = 0;
                if (ciwp == 0.0) {
                    abscoice[0] = 0.0;
                } else if (r.iceflag == 0) {
                    if (radice < 10.0) gc_raise(s, 2, ci, l + 1);
                    abscoice[0] = C.absice0[0] + C.absice0[1] / radice;
                } else if (r.iceflag == 1) {
                    if (radice < 13.0 || radice > 130.) gc_raise(s, 2, ci, l + 1);
                    ncbands = 5;
                    for (int ib = 0; ib < 5; ++ib) abscoice[ib] = C.absice1[0 + 2 * ib] + C.absice1[1 + 2 * ib] / radice;
                    iceind = 1;
                } else {
                    const int nrad = r.iceflag == 2 ? 43 : 46;
                    if (radice < 5.0 || radice > (r.iceflag == 2 ? 131.0 : 140.0)) gc_raise(s, 2, ci, l + 1);
                    ncbands = 16;
                    const double factor = (radice - 2.) / 3.;
                    int index = (int)factor;
                    if (index == nrad) index = nrad - 1;
                    index = index < 1 ? 1 : (index > nrad - 1 ? nrad - 1 : index);  // keeps loads in bounds after an error
                    const double fint = factor - (double)index;
                    const double* t = r.iceflag == 2 ? D.absice2 : D.absice3;
                    for (int ib = 0; ib < 16; ++ib) {
                        const double x0 = ldg(t + (index - 1) + nrad * ib), x1 = ldg(t + index + nrad * ib);
                        abscoice[ib] = x0 + fint * (x1 - (x0));
                    }
                    iceind = 2;
                }
                if (clwp == 0.0) {
                    abscoliq[0] = 0.0;
                    if (iceind == 1) iceind = 2;
                } else if (r.liqflag == 0) {
                    abscoliq[0] = C.absliq0;
                    if (iceind == 1) iceind = 2;
                } else {
                    const double radliq = in.rel[a];
                    if (radliq < 2.5 || radliq > 60.) gc_raise(s, 3, ci, l + 1);
                    int index = (int)(radliq - 1.5);
                    if (index == 0) index = 1;
                    if (index == 58) index = 57;
                    index = index < 1 ? 1 : (index > 57 ? 57 : index);
                    const double fint = radliq - 1.5 - (double)index;
                    ncbands = 16;
                    for (int ib = 0; ib < 16; ++ib) {
                        const double x0 = ldg(D.absliq1 + (index - 1) + 58 * ib), x1 = ldg(D.absliq1 + index + 58 * ib);
                        abscoliq[ib] = x0 + fint * (x1 - (x0));
                    }
                    liqind = 2;
                }
                for (int ib = 0; ib < ncbands; ++ib)
                    tc[ib] = ciwp * abscoice[icb[iceind][ib] - 1] + clwp * abscoliq[icb[liqind][ib] - 1];
            }
        }
        REC(l, GC_CLDFRAC) = cf;
        for (int ib = 0; ib < 16; ++ib) REC(l, GC_ODCLD + ib) = tc[ib];
    }
    // ---- band cloud optical depth along the diffusivity angle (rtrn.f90:321-335 / rtrnmr.f90:333-343): cloud band
    // ib = ipat(band, ncbands), and secdiff is taken at the CLOUD band index, as the reference does
    const int sel = ncbands == 1 ? 0 : (ncbands == 5 ? 1 : 2);
    for (int l = 0; l < nlay; ++l) {
        const double cf = REC(l, GC_CLDFRAC);
        double tc[16];
        for (int ib = 0; ib < 16; ++ib) tc[ib] = REC(l, GC_ODCLD + ib);
        const bool cl = cf >= 1.e-6;
        for (int b = 0; b < 16; ++b) {
            const int ib = icb[sel][b] - 1;
            const double od = cl ? s.secdiff[(long)ib * P + i] * tc[ib] : 0.0;
            REC(l, GC_ODCLD + b) = od;
            REC(l, GC_EFCL + b) = (cl && !maxran) ? (1. - exp(-od)) * cf : 0.0;
        }
        REC(l, GC_FLAGS) = cl ? 1.0 : 0.0;
        for (int k = GC_FACCLR1; k <= GC_FACCMB2D; ++k) REC(l, k) = 0.0;
    }
    int top = 0;
    for (int l = 0; l < nlay; ++l)
        if (REC(l, GC_FLAGS) != 0.0) top = l + 1;
    s.ctop[i] = top;  // > 0: the finalize kernel keeps the clear-sky sums apart
    if (!maxran) return;
    // ---- maximum-random overlap factors, rtrnmr.f90:345-505.  lev = l + 1 (1-based); the upward factors of index lev+1
    // and the downward ones of index lev-1 are stored with layer lev; the "facclr2(lev) = 0 / faccld2(lev) = 0" resets
    // of the reference touch what was stored with layer lev-1 (upward) / lev+1 (downward).
    auto CF = [&](int lev) { return REC(lev - 1, GC_CLDFRAC); };
    auto CLD = [&](int lev) { return REC(lev - 1, GC_FLAGS) != 0.0; };
    double rat1 = 0., rat2 = 0.;
    int ist = 1;  // istcld(lev)
    for (int lev = 1; lev <= nlay; ++lev) {
        const int ist_lev = ist;
        if (CLD(lev)) {
            ist = 0;
            double fcl1 = 0., fcl2 = 0., fcd1 = 0., fcd2 = 0., fmb1 = 0., fmb2 = 0.;
            if (lev == nlay) {
                // all zero
            } else if (CF(lev + 1) >= CF(lev)) {
                if (ist_lev == 1) {
                    if (CF(lev) < 1.) fcl2 = (CF(lev + 1) - CF(lev)) / (1. - CF(lev));
                    if (lev >= 2) { REC(lev - 2, GC_FACCLR2) = 0.; REC(lev - 2, GC_FACCLD2) = 0.; }  // facclr2(lev), faccld2(lev)
                } else {
                    const double fmx = fmax(CF(lev), CF(lev - 1));
                    if (CF(lev + 1) > fmx) { fcl1 = rat2; fcl2 = (CF(lev + 1) - fmx) / (1. - fmx); }
                    else if (CF(lev + 1) < fmx) { fcl1 = (CF(lev + 1) - CF(lev)) / (CF(lev - 1) - CF(lev)); fcl2 = 0.; }
                    else { fcl1 = rat2; fcl2 = 0.; }
                }
                if (fcl1 > 0. || fcl2 > 0.) { rat1 = 1.; rat2 = 0.; } else { rat1 = 0.; rat2 = 0.; }
            } else {
                if (ist_lev == 1) {
                    fcd2 = (CF(lev) - CF(lev + 1)) / CF(lev);
                    if (lev >= 2) { REC(lev - 2, GC_FACCLR2) = 0.; REC(lev - 2, GC_FACCLD2) = 0.; }
                } else {
                    const double fmn = fmin(CF(lev), CF(lev - 1));
                    if (CF(lev + 1) <= fmn) { fcd1 = rat1; fcd2 = (fmn - CF(lev + 1)) / fmn; }
                    else { fcd1 = (CF(lev) - CF(lev + 1)) / (CF(lev) - fmn); fcd2 = 0.; }
                }
                if (fcd1 > 0. || fcd2 > 0.) { rat1 = 0.; rat2 = 1.; } else { rat1 = 0.; rat2 = 0.; }
            }
            if (ist_lev != 1 && lev < nlay) {
                fmb1 = fmax(0., fmin(CF(lev + 1) - CF(lev), CF(lev - 1) - CF(lev)));
                fmb2 = fmax(0., fmin(CF(lev) - CF(lev + 1), CF(lev) - CF(lev - 1)));
            }
            REC(lev - 1, GC_FACCLR1) = fcl1; REC(lev - 1, GC_FACCLR2) = fcl2; REC(lev - 1, GC_FACCLD1) = fcd1;
            REC(lev - 1, GC_FACCLD2) = fcd2; REC(lev - 1, GC_FACCMB1) = fmb1; REC(lev - 1, GC_FACCMB2) = fmb2;
        } else {
            ist = 1;
        }
        if (ist_lev) REC(lev - 1, GC_FLAGS) = REC(lev - 1, GC_FLAGS) + 2.0;
    }
    ist = 1;  // istcldd(lev)
    for (int lev = nlay; lev >= 1; --lev) {
        const int ist_lev = ist;
        if (CLD(lev)) {
            ist = 0;
            double fcl1 = 0., fcl2 = 0., fcd1 = 0., fcd2 = 0., fmb1 = 0., fmb2 = 0.;
            if (lev == 1) {
                // all zero
            } else if (CF(lev - 1) >= CF(lev)) {
                if (ist_lev == 1) {
                    if (CF(lev) < 1.) fcl2 = (CF(lev - 1) - CF(lev)) / (1. - CF(lev));
                    if (lev <= nlay - 1) { REC(lev, GC_FACCLR2D) = 0.; REC(lev, GC_FACCLD2D) = 0.; }  // facclr2d(lev), faccld2d(lev)
                } else {
                    const double fmx = fmax(CF(lev), CF(lev + 1));
                    if (CF(lev - 1) > fmx) { fcl1 = rat2; fcl2 = (CF(lev - 1) - fmx) / (1. - fmx); }
                    else if (CF(lev - 1) < fmx) { fcl1 = (CF(lev - 1) - CF(lev)) / (CF(lev + 1) - CF(lev)); fcl2 = 0.; }
                    else { fcl1 = rat2; fcl2 = 0.; }
                }
                if (fcl1 > 0. || fcl2 > 0.) { rat1 = 1.; rat2 = 0.; } else { rat1 = 0.; rat2 = 0.; }
            } else {
                if (ist_lev == 1) {
                    fcd2 = (CF(lev) - CF(lev - 1)) / CF(lev);
                    if (lev <= nlay - 1) { REC(lev, GC_FACCLR2D) = 0.; REC(lev, GC_FACCLD2D) = 0.; }
                } else {
                    const double fmn = fmin(CF(lev), CF(lev + 1));
                    if (CF(lev - 1) <= fmn) { fcd1 = rat1; fcd2 = (fmn - CF(lev - 1)) / fmn; }
                    else { fcd1 = (CF(lev) - CF(lev - 1)) / (CF(lev) - fmn); fcd2 = 0.; }
                }
                if (fcd1 > 0. || fcd2 > 0.) { rat1 = 0.; rat2 = 1.; } else { rat1 = 0.; rat2 = 0.; }
            }
            if (ist_lev != 1 && lev > 1) {
                fmb1 = fmax(0., fmin(CF(lev + 1) - CF(lev), CF(lev - 1) - CF(lev)));
                fmb2 = fmax(0., fmin(CF(lev) - CF(lev + 1), CF(lev) - CF(lev - 1)));
            }
            REC(lev - 1, GC_FACCLR1D) = fcl1; REC(lev - 1, GC_FACCLR2D) = fcl2; REC(lev - 1, GC_FACCLD1D) = fcd1;
            REC(lev - 1, GC_FACCLD2D) = fcd2; REC(lev - 1, GC_FACCMB1D) = fmb1; REC(lev - 1, GC_FACCMB2D) = fmb2;
        } else {
            ist = 1;
        }
        if (ist_lev) REC(lev - 1, GC_FLAGS) = REC(lev - 1, GC_FLAGS) + 4.0;
    }
}

// The sweeps of band B: thread = column.  Gas part of a layer from the state word (gas_point), total path of a cloudy
// layer from cloud_point (every g-point of the band sees the band's cloud optical depth); the layer update is rtrn's
// (random overlap, MAXRAN = 0) or rtrnmr's (MAXRAN = 1).
template <int B, int MAXRAN, int IDRV>
__global__ void __launch_bounds__(128) gridcloud_sweep_kernel(ColIn in, Scratch s, RunCfg r, const SweepTab tb) {
    constexpr int NG = NGC[B - 1];
    constexpr int G0 = GOFF[B - 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const int laytrop = s.laytrop[i];
    const bool own16 = (B == 16) && r.istart == 16;
    const long blk_lay0 = (long)(i >> 7) * nlay;
    const int blk_t = i & 127;
    auto state = [&](int l0, int g) {
        const double2 w = reinterpret_cast<const double2*>(s.od + ((blk_lay0 + l0) * NGPT + G0) * 128)[((g >> 1) * 128) + blk_t];
        return (g & 1) ? w.y : w.x;
    };
    auto REC = [&](int l0, int k) { return s.gridcld[((long)l0 * GC_NREC + k) * P + i]; };
    const long blk_lev0 = (long)(i >> 7) * (nlay + 1);
    auto flx = [&](int lev) { return s.fluxP + (((blk_lev0 + lev) * NBND + (B - 1)) * 6) * 128 + blk_t; };
    auto pf_of = [&](int l0, PlFrac<B>& pf) {
        double sm = 0.;
        if constexpr (pf_slot(B) >= 0) sm = s.pfs[((blk_lay0 + l0) * NPF + pf_slot(B)) * 128 + blk_t];
        pf.setup(tb, (l0 + 1) <= laytrop, sm);
    };
    double radld[NG], radclrd[NG], cldrad[NG], clrrad[NG], radx[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) { radld[g] = 0.; radclrd[g] = 0.; cldrad[g] = 0.; clrrad[g] = 0.; radx[g] = 0.; }
    int iclddn = 0;
    // ---- downward, rtrn.f90:355-458 / rtrnmr.f90:537-650
    for (int lev = nlay; lev >= 1; --lev) {
        const int l0 = lev - 1;
        const long a = ci + in.ld * l0;
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(in.tlay[a], it_lay, fr_lay);
        planck_pos(in.tlev[a], it_lev, fr_lev);
        const double blay = planck_at<B>(tb.totplnk, it_lay, fr_lay, own16);
        const double dplankdn = planck_at<B>(tb.totplnk, it_lev, fr_lev, own16, lev == 1) - blay;
        PlFrac<B> pf;
        pf_of(l0, pf);
        const int flags = (int)REC(l0, GC_FLAGS);
        const bool cld = flags & 1;
        acc_t dsum = 0., dcsum = 0.;
        if (cld) {
            iclddn = 1;
            const double cf = REC(l0, GC_CLDFRAC), odc = REC(l0, GC_ODCLD + B - 1);
            const double ef = MAXRAN ? 0. : REC(l0, GC_EFCL + B - 1);
            const double fclr1 = REC(l0, GC_FACCLR1D), fclr2 = REC(l0, GC_FACCLR2D), fcld1 = REC(l0, GC_FACCLD1D),
                         fcld2 = REC(l0, GC_FACCLD2D), fcmb1 = REC(l0, GC_FACCMB1D), fcmb2 = REC(l0, GC_FACCMB2D);
            const bool start = flags & 4;  // istcldd(lev) = 1
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double st = state(l0, g);
                double atrans, bbd, atot, bbdtot;
                bool big;
                int itr, ittot;
                gas_point(tb.exptfn, st, pf.get(g), blay, dplankdn, atrans, bbd, big, itr);
                cloud_point(tb, st, big, itr, odc, true, pf.get(g), blay, dplankdn, atot, bbdtot, ittot);
                const double gassrc = bbd * atrans;
                if (MAXRAN) {
                    if (start) {
                        cldrad[g] = cf * radld[g];
                        clrrad[g] = radld[g] - cldrad[g];
                        radx[g] = 0.;
                    }
                    const double ttot = 1. - atot;
                    const double cldsrc = bbdtot * atot;
                    cldrad[g] = cldrad[g] * ttot + cf * cldsrc;
                    clrrad[g] = clrrad[g] * (1. - atrans) + (1. - cf) * gassrc;
                    radld[g] = cldrad[g] + clrrad[g];
                    const double radmod = radx[g] * (fclr1 * (1. - atrans) + fcld1 * ttot) - fcmb1 * gassrc + fcmb2 * cldsrc;
                    const double oldcld = cldrad[g] - radmod;
                    const double oldclr = clrrad[g] + radmod;
                    radx[g] = -radmod + fclr2 * oldclr - fcld2 * oldcld;
                    cldrad[g] = cldrad[g] + radx[g];
                    clrrad[g] = clrrad[g] - radx[g];
                } else {
                    radld[g] = radld[g] - radld[g] * (atrans + ef * (1. - atrans)) + gassrc + cf * (bbdtot * atot - gassrc);
                }
                dsum = dsum + radld[g];
                radclrd[g] = radclrd[g] + (bbd - radclrd[g]) * atrans;
                dcsum = dcsum + radclrd[g];
            }
        } else {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double st = state(l0, g);
                double atrans, bbd;
                bool big;
                int itr;
                gas_point(tb.exptfn, st, pf.get(g), blay, dplankdn, atrans, bbd, big, itr);
                radld[g] = relaxed(radld[g], bbd, atrans);
                dsum = dsum + radld[g];
                radclrd[g] = iclddn ? relaxed(radclrd[g], bbd, atrans) : radld[g];
                dcsum = dcsum + radclrd[g];
            }
        }
        acc_t* o = flx(l0);
        o[0] = (dsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        o[128] = ((iclddn ? dcsum : dsum) * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
    }
    flx(nlay)[0] = 0.0;
    flx(nlay)[128] = 0.0;
    // ---- surface, rtrnmr.f90:663-688
    double d_radlu[IDRV ? NG : 1], d_radclru[IDRV ? NG : 1];
    {
        int it_b;
        double fr_b;
        planck_pos(in.tsfc[ci], it_b, fr_b);
        const double semiss = in.emis[ci + in.ld * (B - 1)];
        const double* tp = tb.totplnk + 181 * ((B == 16 && own16) ? 16 : B - 1);
        const double pa = ldg(tp + it_b - 1), pb = ldg(tp + it_b);
        const double plankbnd = semiss * (pa + fr_b * (pb - pa));
        double dplankbnd_dt = 0.;
        if constexpr (IDRV == 1) {
            const double* td = tb.totplnkderiv + 181 * ((B == 16 && own16) ? 16 : B - 1);
            const double da = ldg(td + it_b - 1), db = ldg(td + it_b);
            dplankbnd_dt = semiss * (da + fr_b * (db - da));
        }
        const double reflect = 1. - semiss;
        PlFrac<B> pf;
        pf_of(0, pf);
        acc_t usum = 0., ucsum = 0., dusum = 0.;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const double rad0 = pf.get(g) * plankbnd;
            radld[g] = rad0 + reflect * radld[g];       // radlu
            radclrd[g] = rad0 + reflect * radclrd[g];   // radclru
            usum = usum + radld[g];
            ucsum = ucsum + radclrd[g];
            if constexpr (IDRV == 1) {
                d_radlu[g] = pf.get(g) * dplankbnd_dt;
                d_radclru[g] = d_radlu[g];
                dusum = dusum + d_radlu[g];
            }
        }
        acc_t* o = flx(0);
        o[256] = (usum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        o[384] = (ucsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        if constexpr (IDRV == 1) {
            const acc_t addd = (dusum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
            o[512] = addd;
            o[640] = addd;
        }
    }
    // ---- upward, rtrn.f90:486-545 / rtrnmr.f90:690-757
    for (int lev = 1; lev <= nlay; ++lev) {
        const int l0 = lev - 1;
        const long a = ci + in.ld * l0;
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(in.tlay[a], it_lay, fr_lay);
        planck_pos(in.tlev[a + in.ld], it_lev, fr_lev);
        const double blay = planck_at<B>(tb.totplnk, it_lay, fr_lay, own16);
        const double dplankup = planck_at<B>(tb.totplnk, it_lev, fr_lev, own16) - blay;
        PlFrac<B> pf;
        pf_of(l0, pf);
        const int flags = (int)REC(l0, GC_FLAGS);
        const bool cld = flags & 1;
        acc_t usum = 0., ucsum = 0., dusum = 0., ducsum = 0.;
        if (cld) {
            const double cf = REC(l0, GC_CLDFRAC), odc = REC(l0, GC_ODCLD + B - 1);
            const double ef = MAXRAN ? 0. : REC(l0, GC_EFCL + B - 1);
            const double fclr1 = REC(l0, GC_FACCLR1), fclr2 = REC(l0, GC_FACCLR2), fcld1 = REC(l0, GC_FACCLD1),
                         fcld2 = REC(l0, GC_FACCLD2), fcmb1 = REC(l0, GC_FACCMB1), fcmb2 = REC(l0, GC_FACCMB2);
            const bool start = flags & 2;  // istcld(lev) = 1
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double st = state(l0, g);
                double atrans, bbugas, atot, bbutot;
                bool big;
                int itr, ittot;
                gas_point(tb.exptfn, st, pf.get(g), blay, dplankup, atrans, bbugas, big, itr);
                cloud_point(tb, st, big, itr, odc, true, pf.get(g), blay, dplankup, atot, bbutot, ittot);
                const double gassrc = bbugas * atrans;
                if (MAXRAN) {
                    if (start) {
                        cldrad[g] = cf * radld[g];
                        clrrad[g] = radld[g] - cldrad[g];
                        radx[g] = 0.;
                    }
                    const double ttot = 1. - atot;
                    const double cldsrc = bbutot * atot;
                    cldrad[g] = cldrad[g] * ttot + cf * cldsrc;
                    clrrad[g] = clrrad[g] * (1.0 - atrans) + (1. - cf) * gassrc;
                    radld[g] = cldrad[g] + clrrad[g];
                    const double radmod = radx[g] * (fclr1 * (1.0 - atrans) + fcld1 * ttot) - fcmb1 * gassrc + fcmb2 * cldsrc;
                    const double oldcld = cldrad[g] - radmod;
                    const double oldclr = clrrad[g] + radmod;
                    radx[g] = -radmod + fclr2 * oldclr - fcld2 * oldcld;
                    cldrad[g] = cldrad[g] + radx[g];
                    clrrad[g] = clrrad[g] - radx[g];
                } else {
                    radld[g] = radld[g] - radld[g] * (atrans + ef * (1. - atrans)) + gassrc + cf * (bbutot * atot - gassrc);
                }
                usum = usum + radld[g];
                if constexpr (IDRV == 1) {
                    d_radlu[g] = d_radlu[g] * cf * (1.0 - atot) + d_radlu[g] * (1.0 - cf) * (1.0 - atrans);
                    dusum = dusum + d_radlu[g];
                }
                radclrd[g] = iclddn ? relaxed(radclrd[g], bbugas, atrans) : radld[g];
                ucsum = ucsum + radclrd[g];
                if constexpr (IDRV == 1) {
                    d_radclru[g] = iclddn ? d_radclru[g] * (1.0 - atrans) : d_radlu[g];
                    ducsum = ducsum + d_radclru[g];
                }
            }
        } else {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double st = state(l0, g);
                double atrans, bbugas;
                bool big;
                int itr;
                gas_point(tb.exptfn, st, pf.get(g), blay, dplankup, atrans, bbugas, big, itr);
                radld[g] = relaxed(radld[g], bbugas, atrans);
                usum = usum + radld[g];
                if constexpr (IDRV == 1) {
                    d_radlu[g] = d_radlu[g] * (1.0 - atrans);
                    dusum = dusum + d_radlu[g];
                }
                radclrd[g] = iclddn ? relaxed(radclrd[g], bbugas, atrans) : radld[g];
                ucsum = ucsum + radclrd[g];
                if constexpr (IDRV == 1) {
                    d_radclru[g] = iclddn ? d_radclru[g] * (1.0 - atrans) : d_radlu[g];
                    ducsum = ducsum + d_radclru[g];
                }
            }
        }
        acc_t* o = flx(lev);
        o[256] = (usum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        o[384] = (ucsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        if constexpr (IDRV == 1) {
            o[512] = (dusum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
            o[640] = (ducsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
        }
    }
}

}  // namespace lw
