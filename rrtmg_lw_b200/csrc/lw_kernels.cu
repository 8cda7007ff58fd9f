// lw_kernels.cu -- column-parallel kernels around the band kernel: inatm + setcoef (prep), the
// McICA sub-column generator (KISS / MT19937), cloud-flag scan for the strict rrtmg_lw() entry,
// flux finalisation (fnet, heating rate), get_alpha, sub-column materialisation.
#include <cstdio>

#include "lw_band.cuh"
#include "lw_launch.h"

namespace lw {

cudaError_t upload_constants_kernels(const ConstTab& c, const DevPtrs& d, const double (*refrat)[5]) {
    return upload_constants_here(c, d, refrat);
}

__device__ __forceinline__ void raise(const Scratch& s, int code, long col, int lay) {
    const unsigned long long v = ((unsigned long long)code << 56) | ((unsigned long long)(col + s.err_base + 1) << 24) |
                                 (unsigned long long)(lay & 0xffffff);
    atomicCAS(s.status, 0ull, v);
}

static inline int nblk(long n, int b) { return (int)((n + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// inatm (src/rrtmg_lw_rad.f90:755-870) + setcoef (src/rrtmg_lw_setcoef.f90:271-432) for one column
// per thread.  The two whole-column sums (amttl, wvttl -> pwvcm) run in layer order, as in the
// reference; everything else is per layer and goes to the column-innermost scratch.
__global__ void __launch_bounds__(128) prep_kernel(ColIn in, Scratch s, RunCfg r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch, plane = (long)nlay * P;
    const double amd = 28.9660, amw = 18.0160;
    const double stpfac = 296. / 1013.;
    const GTab& g = *D.g;
    const bool dbg = s.dbg_col == i;
    double amttl = 0.0, wvttl = 0.0;
    int laytrop = 0;
    double pz_lo = in.plev[ci];
    const double pz0 = pz_lo;
    for (int l = 0; l < nlay; ++l) {
        const long a = ci + in.ld * l;
        const long idx = (long)l * P + i;
        const double pavel = in.play[a], tavel = in.tlay[a];
        const double pz_hi = in.plev[ci + in.ld * (l + 1)];
        const double h2o = in.h2o[a], co2 = in.co2[a], o3 = in.o3[a], n2o = in.n2o[a], ch4 = in.ch4[a], o2 = in.o2[a];
        const double amm = (1. - h2o) * amd + h2o * amw;
        const double coldry = (pz_lo - pz_hi) * 1.e3 * C.avogad / (1.e2 * C.grav * amm * (1. + h2o));
        pz_lo = pz_hi;
        double summol = 0.0;
        summol = summol + co2; summol = summol + o3; summol = summol + n2o; summol = summol + 0.0;
        summol = summol + ch4; summol = summol + o2;
        const double wbrodl = coldry * (1. - summol);
        const double w1 = coldry * h2o, w2 = coldry * co2, w3 = coldry * o3, w4 = coldry * n2o, w6 = coldry * ch4,
                     w7 = coldry * o2;
        amttl = amttl + coldry + w1;
        wvttl = wvttl + w1;
        // FP32 mode: the float band kernels read float fields from the same memory
        auto put = [&](int f, double v) {
            const long o = r.fld_block ? ((((long)(i >> 7) * nlay + l) * NFIELD + f) * 128 + (i & 127)) : ((long)f * plane + idx);
            if (r.fp32) s.fldf[o] = (float)v;
            else s.fld[o] = v;
        };
        put(F_WX1, coldry * in.ccl4[a] * 1.e-20);
        put(F_WX2, coldry * in.cfc11[a] * 1.e-20);
        put(F_WX3, coldry * in.cfc12[a] * 1.e-20);
        put(F_WX4, coldry * in.cfc22[a] * 1.e-20);

        // setcoef
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
        double forfac = scalefac / (1. + water);
        double selffac = water * forfac;
        int indfor, indself = 0;
        double forfrac, selffrac = 0.;
        if (!(plog <= 4.56)) {
            laytrop = laytrop + 1;
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
        s.ipack[idx] = jp | (jt << 6) | (jt1 << 9) | (indself << 12) | (indfor << 16) | (indminor << 18);
        if (dbg) {
            int* q = s.dbg_ints + 6 * l;
            q[0] = jp; q[1] = jt; q[2] = jt1; q[3] = indself; q[4] = indfor; q[5] = indminor;
            double* fq = s.dbg_facs + 4 * l;
            fq[0] = compfp * (1. - ft); fq[1] = fp * (1. - ft1); fq[2] = compfp * ft; fq[3] = fp * ft1;
        }
    }
    const double wvsh = (amw * wvttl) / (amd * amttl);
    const double pwvcm = wvsh * (1.e3 * pz0) / (1.e2 * C.grav);
    s.pwvcm[i] = pwvcm;
    s.laytrop[i] = laytrop;
    if (dbg) { s.dbg_scal[0] = laytrop; s.dbg_scal[1] = pwvcm; }
}

// ---------------------------------------------------------------------------------------------
// Column scalars: laytrop (setcoef.f90:312-313), pwvcm (rad.f90:860-870, sums in layer order) and the diffusivity
// angle of every band (rtrnmc.f90:258-281).  One thread per column.
__global__ void __launch_bounds__(128) colprep_kernel(ColIn in, Scratch s, RunCfg r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const double amd = 28.9660, amw = 18.0160;
    double amttl = 0.0, wvttl = 0.0;
    int laytrop = 0;
    double pz_lo = in.plev[ci];
    const double pz0 = pz_lo;
    for (int l = 0; l < nlay; ++l) {
        const long a = ci + in.ld * l;
        const double pavel = in.play[a];
        const double pz_hi = in.plev[ci + in.ld * (l + 1)];
        const double h2o = in.h2o[a];
        const double amm = (1. - h2o) * amd + h2o * amw;
        const double coldry = (pz_lo - pz_hi) * 1.e3 * C.avogad / (1.e2 * C.grav * amm * (1. + h2o));
        pz_lo = pz_hi;
        const double w1 = coldry * h2o;
        amttl = amttl + coldry + w1;
        wvttl = wvttl + w1;
        const double plog = log(pavel);
        if (!(plog <= 4.56)) laytrop = laytrop + 1;
    }
    const double wvsh = (amw * wvttl) / (amd * amttl);
    const double pwvcm = wvsh * (1.e3 * pz0) / (1.e2 * C.grav);
    s.pwvcm[i] = pwvcm;
    s.laytrop[i] = laytrop;
#pragma unroll
    for (int b = 1; b <= NBND; ++b) {
        double secdiff;
        if (b == 1 || b == 4 || b >= 10) {
            secdiff = 1.66;
        } else {
            secdiff = C.a0[b - 1] + C.a1[b - 1] * exp(C.a2[b - 1] * pwvcm);
            if (secdiff > 1.80) secdiff = 1.80;
            if (secdiff < 1.50) secdiff = 1.50;
        }
        s.secdiff[(long)(b - 1) * s.pitch + i] = secdiff;
    }
    if (s.dbg_col == i) { s.dbg_scal[0] = laytrop; s.dbg_scal[1] = pwvcm; }
}

// cldprmc for one sub-column of band b (0-based), cldprmc.f90:177-266: cloud optical depth from the water paths and
// effective radii; tauc_in is the caller's value, returned unchanged unless inflag = 2 finds cloud water
__device__ __forceinline__ double cld_tauc_band(int b, double ciwp, double clwp, double radice, double radliq, double tauc_in,
                                                const RunCfg& r) {
    double tauc = tauc_in;
    const double cwp = ciwp + clwp;
    if (r.inflag == 2 && (cwp >= 1.e-20 || tauc >= 1.e-20)) {
        double abscoice = 0.0, abscoliq = 0.0;
        if (ciwp != 0.0) {
            if (r.iceflag == 0) {
                abscoice = C.absice0[0] + C.absice0[1] / radice;
            } else if (r.iceflag == 1) {
                const int icb[16] = {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
                abscoice = C.absice1[0 + 2 * (icb[b] - 1)] + C.absice1[1 + 2 * (icb[b] - 1)] / radice;
            } else {
                const double factor = (radice - 2.) / 3.;
                int index = (int)factor;
                const int nrad = r.iceflag == 2 ? 43 : 46;
                if (index == nrad) index = nrad - 1;
                if (index < 1) index = 1;        // unreachable for valid radii; keeps loads in bounds
                if (index > nrad - 1) index = nrad - 1;
                const double fint = factor - (double)index;
                const double* t = (r.iceflag == 2 ? D.absice2 : D.absice3) + nrad * b;
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
                const double* t = D.absliq1 + 58 * b;
                const double x0 = ldg(t + index - 1), x1 = ldg(t + index);
                abscoliq = x0 + fint * (x1 - (x0));
            }
        }
        tauc = ciwp * abscoice + clwp * abscoliq;
    }
    return tauc;
}

// ---------------------------------------------------------------------------------------------
// cldprmc for the McICA mask entry of the two-phase path (cldprmc.f90:177-266 + rtrnmc.f90:316-326): one thread per
// (column, layer).  Every sub-column of a band that holds cloud in this layer has the same water paths and radii
// (they are the grid-mean values, mcica:658-679), hence the same band optical depth: the kernel writes, per band whose
// sub-column bits are not all clear, odcld = secdiff * taucmc and abscld = 1 - exp(-odcld) for the sweep kernels.
// The layer axis is parallel here, so the exp() and the table lookups stay out of the sweeps' vertical recurrences.
// (column i of the pass, layer l) whose five mask words are w (flag bit of word 4 cleared)
__device__ __forceinline__ void cldoptics_point(const ColIn& in, const Scratch& s, const RunCfg& r, const int i, const int l,
                                                const unsigned (&w)[5]) {
    const long P = s.pitch;
    const long ci = r.c0 + i;
    const long a = ci + in.ld * l;
    const double ciwp = in.ciwp[a], clwp = in.clwp[a];
    const double radice = in.rei[a], radliq = in.rel[a];
    double* q = s.cldb + ((((long)(i >> 7) * r.nlay + l) * NBND) * 2) * 128 + (i & 127);
#pragma unroll
    for (int b = 0; b < NBND; ++b) {
        // the band's sub-column bits: g-points GOFF[b] .. GOFF[b] + NGC[b] - 1
        const int goff[NBND] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138};
        const int ngc[NBND] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
        const int g0 = goff[b], ng = ngc[b];
        const int wlo = g0 >> 5, sh = g0 & 31;
        unsigned bits = w[wlo] >> sh;
        if (sh + ng > 32) bits |= w[wlo + 1] << (32 - sh);
        bits &= (1u << ng) - 1u;
        if (!bits) continue;
        const double tauc = cld_tauc_band(b, ciwp, clwp, radice, radliq, in.tauc ? in.tauc[b + 16 * a] : 0.0, r);
        const double odcld = s.secdiff[(long)b * P + i] * tauc;
        q[(b * 2) * 128] = odcld;
        q[(b * 2 + 1) * 128] = 1. - exp(-odcld);
    }
}
__global__ void __launch_bounds__(128) cldoptics_kernel(ColIn in, Scratch s, RunCfg r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (i >= r.ncol) return;
    const long P = s.pitch;
    const unsigned* mw = s.mask + ((long)l * 5) * P + i;
    const unsigned w4 = mw[4 * P];
    if ((w4 >> 31) == 0) return;  // icldlyr = 0
    const unsigned w[5] = {mw[0], mw[P], mw[2 * P], mw[3 * P], w4 & 0x7fffffffu};
    cldoptics_point(in, s, r, i, l, w);
}

// ---------------------------------------------------------------------------------------------
// KISS generator step, src/mcica_subcol_gen_lw.f90:731-741 (INT32 wrap-around, logical shifts)
struct Kiss {
    unsigned s1, s2, s3, s4;
    __device__ __forceinline__ double next() {
        s1 = 69069u * s1 + 1327217885u;
        s2 ^= s2 << 13; s2 ^= s2 >> 17; s2 ^= s2 << 5;
        s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
        s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
        const int kiss = (int)(s1 + s2 + (s3 << 16) + s4);
        return (double)kiss * 2.328306e-10 + 0.5;
    }
    __device__ __forceinline__ void step() {  // advance without forming the number
        s1 = 69069u * s1 + 1327217885u;
        s2 ^= s2 << 13; s2 ^= s2 >> 17; s2 ^= s2 << 5;
        s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
        s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
    }
};

// per-layer cloud flag (icldlyr, bit 31 of mask word 4), column cloud top, and the validation of the cloud-optics
// inputs the way cldprmc would STOP (cldprmc.f90:191-253), for column i whose 140 mask bits per layer are in place
__device__ void finish_column(const ColIn& in, const Scratch& s, const RunCfg& r, int i) {
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const double cldmin = 1.0e-20;
    int ctop = 0;
    for (int l = 0; l < nlay; ++l) {
        unsigned any = 0;
        for (int w = 0; w < 5; ++w) any |= s.mask[((long)l * 5 + w) * P + i];
        if (any) {
            s.mask[((long)l * 5 + 4) * P + i] |= 0x80000000u;
            ctop = l + 1;
            const long a = ci + in.ld * l;
            const double ciwp = in.ciwp[a], clwp = in.clwp[a];
            bool taubig = false;
            if (in.tauc) for (int b = 0; b < 16; ++b) taubig |= in.tauc[b + 16 * a] >= cldmin;
            if ((ciwp + clwp) >= cldmin || taubig) {
                if (r.inflag == 1) raise(s, 1, ci, l + 1);
                if (r.inflag == 2) {
                    if (ciwp != 0.0) {
                        const double rad = in.rei[a];
                        if ((r.iceflag == 0 && rad < 10.0) || (r.iceflag == 1 && (rad < 13.0 || rad > 130.)) ||
                            (r.iceflag == 2 && (rad < 5.0 || rad > 131.0)) || (r.iceflag == 3 && (rad < 5.0 || rad > 140.0)))
                            raise(s, 2, ci, l + 1);
                    }
                    if (clwp != 0.0 && r.liqflag == 1) {
                        const double rad = in.rel[a];
                        if (rad < 2.5 || rad > 60.) raise(s, 3, ci, l + 1);
                    }
                }
            }
        }
    }
    s.ctop[i] = ctop;}

// One thread = one column: draws its 140 x nlay (or 2 x that) numbers in the reference's order
// (sub-column outer, layer inner), applies the overlap recurrence on the fly and packs the cloud
// mask as 140 bits per layer.  generate_stochastic_clouds, mcica:442-679; icld 4/5 per the column
// version mcica_subcol_gen_lw.1col.f90:543-611 (the GCM file does not compile as shipped).
// Also validates the cloud-optics inputs the way cldprmc would STOP (cldprmc.f90:191-253).
__global__ void __launch_bounds__(128) mcica_kiss_kernel(ColIn in, Scratch s, RunCfg r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const int icld = r.icld;
    unsigned mloc[LW_MAXLAY];  // mask word of the current 32 sub-columns, one per layer (L1-backed)
    Kiss k;
    {
        const double p1 = in.play[ci] * 1.e2, p2 = in.play[ci + in.ld] * 1.e2;
        if (p1 < p2) {
            raise(s, 4, ci, 1);
            s.ctop[i] = 0;
            for (int k = 0; k < nlay * 5; ++k) s.mask[(long)k * P + i] = 0u;
            return;
        }
        const double p3 = in.play[ci + 2 * in.ld] * 1.e2, p4 = in.play[ci + 3 * in.ld] * 1.e2;
        k.s1 = (unsigned)(int)((p1 - (double)(int)p1) * 1000000000.0);
        k.s2 = (unsigned)(int)((p2 - (double)(int)p2) * 1000000000.0);
        k.s3 = (unsigned)(int)((p3 - (double)(int)p3) * 1000000000.0);
        k.s4 = (unsigned)(int)((p4 - (double)(int)p4) * 1000000000.0);
        for (int n = 0; n < r.permuteseed; ++n) k.next();
    }
    const double cldmin = 1.0e-20;
    for (int sc = 0; sc < NGPT; ++sc) {
        const unsigned bit = 1u << (sc & 31);
        if ((sc & 31) == 0)
            for (int l = 0; l < nlay; ++l) mloc[l] = 0u;
        double prev = 0.;
        double rmax = 0.;
        if (icld == 3) rmax = k.next();
        for (int l = 0; l < nlay; ++l) {
            double cf = in.cldfrac[ci + in.ld * l];
            if (cf < cldmin) cf = 0.;
            double x;
            if (icld == 3) {
                x = rmax;
            } else {
                x = k.next();
                if (icld >= 4) {
                    const double x2 = k.next();
                    if (l > 0 && x2 < in.alpha[ci + in.ld * l]) x = prev;
                } else if (icld == 2 && l > 0) {
                    double cfm = in.cldfrac[ci + in.ld * (l - 1)];
                    if (cfm < cldmin) cfm = 0.;
                    if (prev > 1. - cfm) x = prev; else x = x * (1. - cfm);
                }
            }
            prev = x;
            if (x >= 1. - cf) mloc[l] |= bit;
        }
        if ((sc & 31) == 31 || sc == NGPT - 1)
            for (int l = 0; l < nlay; ++l) s.mask[((long)l * 5 + (sc >> 5)) * P + i] = mloc[l];
    }
    finish_column(in, s, r, i);
}

// ---------------------------------------------------------------------------------------------
// Mersenne-Twister path (irng = 1).  The reference draws ONE global stream in the order
// (sub-column, column, layer) [two draws per point for icld 4/5, one per (sub-column, column) for icld 3]
// from new_RandomNumberSequence(seed = permuteseed) (mcica:478, 523-530, 586-595; MT19937 itself
// src/mcica_random_numbers.f90:116-185).  The 140 sub-column runs of a column block are generated side by side:
// generator CTA sc starts (sc * ncol_of_the_call + first column) * draws_per_column words into the stream, a position
// it reaches with the GF(2) jump-ahead of lw_mtjump.cpp (mt_jump_kernel), regenerates the 624-word state in three
// data-parallel phases per period and emits the tempered words of its run tile by tile; a (column, sub-column)-parallel
// kernel applies the overlap recurrence and sets the sub-column's bit in the packed cloud mask.  The words are the
// reference's, bit for bit; what changes is that 140 CTAs advance 140 copies of the stream instead of one.
struct MtState {
    unsigned mt[624];
    int cur;
};
__global__ void mt_seed_kernel(MtState* st, unsigned seed) {  // initialize_scalar, random_numbers:169-185
    unsigned x = seed;
    st->mt[0] = x;
    for (int i = 1; i < 624; ++i) { x = 1812433253u * (x ^ (x >> 30)) + (unsigned)i; st->mt[i] = x; }
    st->cur = 624;
}
__device__ __forceinline__ unsigned mt_twist(unsigned u, unsigned v) {
    const unsigned mix = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (mix >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
// CTA q advances generator state st[q] and writes its next `emit` tempered words to out + q * out_stride (after
// stepping over `skip` words)
__global__ void __launch_bounds__(256) mt_fill_kernel(MtState* st, unsigned* __restrict__ out, long out_stride, long skip,
                                                      long emit) {
    __shared__ unsigned mt[624];
    const int t = threadIdx.x;
    st += blockIdx.x;
    out += (long)blockIdx.x * out_stride;
    for (int k = t; k < 624; k += 256) mt[k] = st->mt[k];
    int cur = st->cur;
    __syncthreads();
    long pos = -skip;  // index of the next word relative to out[0]
    while (pos < emit) {
        if (cur >= 624) {  // nextState, random_numbers:134-152
            unsigned v = 0;
            if (t < 227) v = mt[t + 397] ^ mt_twist(mt[t], mt[t + 1]);
            __syncthreads();
            if (t < 227) mt[t] = v;
            __syncthreads();
            if (t < 227) v = mt[t] ^ mt_twist(mt[t + 227], mt[t + 228]);
            __syncthreads();
            if (t < 227) mt[t + 227] = v;
            __syncthreads();
            if (t < 169) v = mt[t + 227] ^ mt_twist(mt[t + 454], mt[t + 455]);
            __syncthreads();
            if (t < 169) mt[t + 454] = v;
            __syncthreads();
            if (t == 0) mt[623] = mt[396] ^ mt_twist(mt[623], mt[0]);
            __syncthreads();
            cur = 0;
        }
        const long left = emit - pos;
        const int n = (624 - cur) < left ? (624 - cur) : (int)left;
        for (int k = t; k < n; k += 256) {
            const long o = pos + k;
            if (o >= 0) {  // temper, random_numbers:154-165
                unsigned y = mt[cur + k];
                y ^= y >> 11;
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                out[o] = y ^ (y >> 18);
            }
        }
        pos += n;
        cur += n;
        __syncthreads();
    }
    for (int k = t; k < 624; k += 256) st->mt[k] = mt[k];
    if (t == 0) st->cur = cur;
}
// Jump-ahead: out[q] = the state n0 + q * stride words after the freshly seeded state *seed (lw_mtjump.cpp).  For every
// set bit b of the distance: expand the current window to the next 19937 + 623 raw words, then word j of the new window
// is the XOR of x[i + j] over the set coefficients i of x^(2^b) mod phi.  One CTA per target state.
constexpr int MTJ_DEG = 19937;
constexpr int MTJ_XS = MTJ_DEG + 624;  // raw words a pass needs
__global__ void __launch_bounds__(640) mt_jump_kernel(const MtState* __restrict__ seed, MtState* __restrict__ out,
                                                      const unsigned* __restrict__ pow2, unsigned long long n0,
                                                      unsigned long long stride) {
    extern __shared__ unsigned xs[];  // MTJ_XS words
    const int t = threadIdx.x;
    const unsigned long long n = n0 + (unsigned long long)blockIdx.x * stride;
    for (int k = t; k < 624; k += blockDim.x) xs[k] = seed->mt[k];
    __syncthreads();
    for (int b = 0; b < MTJ_NPOW && (n >> b) != 0ull; ++b) {
        if (!((n >> b) & 1ull)) continue;  // uniform over the CTA
        // x[k + 624] = x[k + 397] ^ twist(x[k], x[k + 1]): 227 words per step are independent of each other
        for (int base = 0; base < MTJ_DEG; base += 227) {
            const int k = base + t;
            if (t < 227 && k < MTJ_DEG) xs[k + 624] = xs[k + 397] ^ mt_twist(xs[k], xs[k + 1]);
            __syncthreads();
        }
        const unsigned* g = pow2 + (size_t)b * 624;
        unsigned y = 0u;
        if (t < 624) {
            for (int w = 0; w < 624; ++w) {
                const unsigned gw = __ldg(g + w);  // uniform
                if (gw == 0u) continue;
                const unsigned* x = xs + w * 32 + t;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if ((gw >> i) & 1u) y ^= x[i];
            }
        }
        __syncthreads();
        if (t < 624) xs[t] = y;
        __syncthreads();
    }
    MtState* o = out + blockIdx.x;
    for (int k = t; k < 624; k += blockDim.x) o->mt[k] = xs[k];
    if (t == 0) o->cur = 624;
}

// getRandomReal, random_numbers:276-295: for a negative INT32 the sum localInt + 2.0**32 is evaluated in the
// default real kind (float under plain gfortran), then divided by the kind-rb 2**32 - 1
__device__ __forceinline__ double mt_real(unsigned w) {
    const int li = (int)w;
    const double denom = 4294967296.0 - 1.0;
    if (li < 0) {
        if (C.drk == 4) return (double)((float)li + 4294967296.0f) / denom;
        return ((double)li + 4294967296.0) / denom;
    }
    return (double)li / denom;
}
// one thread = one column of the tile [t0, t0 + tn) of the chunk and sub-column sc = blockIdx.y: its draws are
// buf[sc * stride + ((i - t0) * nlay + l) * d (+1)]; the chunk's mask words were zeroed beforehand
__global__ void __launch_bounds__(128) mcica_mt_apply_kernel(ColIn in, Scratch s, RunCfg r, const unsigned* __restrict__ buf,
                                                             long stride, int t0, int tn) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= tn) return;
    const int i = t0 + it;
    const int sc = blockIdx.y;
    buf += (long)sc * stride;
    const long ci = r.c0 + i;
    const int nlay = r.nlay, icld = r.icld;
    const long P = s.pitch;
    const double cldmin = 1.0e-20;
    const unsigned bit = 1u << (sc & 31);
    const int d = icld >= 4 ? 2 : 1;
    const unsigned* q = icld == 3 ? buf + it : buf + (long)it * nlay * d;
    double prev = 0., cfm = 0.;
    const double rmax = icld == 3 ? mt_real(q[0]) : 0.;
    for (int l = 0; l < nlay; ++l) {
        double cf = in.cldfrac[ci + in.ld * l];
        if (cf < cldmin) cf = 0.;
        double x;
        if (icld == 3) {
            x = rmax;
        } else {
            x = mt_real(q[l * d]);
            if (icld >= 4) {
                const double x2 = mt_real(q[l * d + 1]);
                if (l > 0 && x2 < in.alpha[ci + in.ld * l]) x = prev;
            } else if (icld == 2 && l > 0) {
                if (prev > 1. - cfm) x = prev; else x = x * (1. - cfm);
            }
        }
        prev = x;
        cfm = cf;
        if (x >= 1. - cf) atomicOr(s.mask + ((long)l * 5 + (sc >> 5)) * P + i, bit);
    }
}
__global__ void __launch_bounds__(128) mcica_finish_kernel(ColIn in, Scratch s, RunCfg r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r.ncol) finish_column(in, s, r, i);
}

// generator state / word buffer and the KISS jump table live on the device that uses them (one set per device)
constexpr int LWK_MAX_DEVICES = 16;
static inline int cur_dev() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= LWK_MAX_DEVICES) d = 0;
    return d;
}
struct MtBuffers {
    MtState* seed = nullptr;    // the freshly seeded state of the call
    MtState* run = nullptr;     // NGPT states, one per sub-column generator
    unsigned* pow2 = nullptr;   // MTJ_NPOW x 624 words: x^(2^b) mod phi
    unsigned* buf = nullptr;    // NGPT x buf_stride tempered words of the current tile
    size_t buf_words = 0;
    bool jump_attr = false;
};
static MtBuffers g_mt_d[LWK_MAX_DEVICES];
#define g_mt g_mt_d[cur_dev()]

void free_mt_buffers() {
    MtBuffers& m = g_mt;
    if (m.seed) cudaFree(m.seed);
    if (m.run) cudaFree(m.run);
    if (m.pow2) cudaFree(m.pow2);
    if (m.buf) cudaFree(m.buf);
    m = MtBuffers{};
}

#ifndef LW_MT_TILE_WORDS
#define LW_MT_TILE_WORDS (48L << 20)  // words of the tile buffer (192 MB): 140 sub-column runs of a column tile
#endif

// Masks of chunk columns [mt_c0, mt_c0 + r.ncol) of a call over mt_ncol columns; returns the number of kernel
// launches, or -1 when a buffer cannot be allocated or the jump table cannot be built.
int launch_mcica_mt(const ColIn& in, const Scratch& s, const RunCfg& r, long mt_ncol, long mt_c0, cudaStream_t st) {
    const long per_col = r.icld == 3 ? 1 : (r.icld >= 4 ? 2L * r.nlay : (long)r.nlay);  // draws per (sub-column, column)
    MtBuffers& m = g_mt;
    if (!m.seed && cudaMalloc((void**)&m.seed, sizeof(MtState)) != cudaSuccess) return -1;
    if (!m.run && cudaMalloc((void**)&m.run, NGPT * sizeof(MtState)) != cudaSuccess) return -1;
    if (!m.pow2) {
        const uint32_t* h = mt_jump_pow2_table();
        if (!h) return -1;
        if (cudaMalloc((void**)&m.pow2, (size_t)MTJ_NPOW * 624 * sizeof(unsigned)) != cudaSuccess) return -1;
        if (cudaMemcpy(m.pow2, h, (size_t)MTJ_NPOW * 624 * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    }
    // column tile: all 140 runs of it fit the buffer
    long tile = LW_MT_TILE_WORDS / (NGPT * per_col);
    tile = tile < 128 ? 128 : (tile / 128) * 128;
    if (tile > r.ncol) tile = r.ncol;
    const size_t stride = (size_t)tile * per_col;
    const size_t need = stride * NGPT;
    if (need > m.buf_words) {
        cudaStreamSynchronize(st);
        if (m.buf) cudaFree(m.buf);
        m.buf = nullptr; m.buf_words = 0;
        if (cudaMalloc((void**)&m.buf, need * sizeof(unsigned)) != cudaSuccess) return -1;
        m.buf_words = need;
    }
    const size_t jsmem = (size_t)MTJ_XS * sizeof(unsigned);
    if (!m.jump_attr) {
        if (cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsmem) != cudaSuccess) return -1;
        m.jump_attr = true;
    }
    cudaMemsetAsync(s.mask, 0, (size_t)r.nlay * 5 * s.pitch * sizeof(unsigned), st);
    mt_seed_kernel<<<1, 1, 0, st>>>(m.seed, (unsigned)r.permuteseed);
    // generator sc starts at the chunk's first column of sub-column sc
    mt_jump_kernel<<<NGPT, 640, jsmem, st>>>(m.seed, m.run, m.pow2, (unsigned long long)(mt_c0 * per_col),
                                             (unsigned long long)(mt_ncol * per_col));
    int launches = 3;
    for (long t0 = 0; t0 < r.ncol; t0 += tile) {
        const int tn = (int)((r.ncol - t0) < tile ? (r.ncol - t0) : tile);
        mt_fill_kernel<<<NGPT, 256, 0, st>>>(m.run, m.buf, (long)stride, 0L, (long)tn * per_col);
        mcica_mt_apply_kernel<<<dim3((unsigned)nblk(tn, 128), NGPT), 128, 0, st>>>(in, s, r, m.buf, (long)stride, (int)t0, tn);
        launches += 2;
    }
    mcica_finish_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, s, r);
    return launches + 1;
}

// ---------------------------------------------------------------------------------------------
// Sub-column-parallel KISS generator.  The reference draws a column's numbers as ONE stream
// (sub-column outer, layer inner, mcica:523-543); the four component generators of KISS all have
// closed-form n-step jumps, so thread (column, sub-column) starts its own copy of the stream at
// draw  permuteseed + isubcol * draws_per_subcolumn :
//   s1  x -> 69069 x + 1327217885 (mod 2^32)            affine map, n-fold composition
//   s2  32-bit xorshift (13, 17, 5)                      linear over GF(2): a 32x32 bit matrix power
//   s3  multiply-with-carry, a = 18000 (low/high half)   x -> a x mod (a 2^16 - 1)
//   s4  multiply-with-carry, a = 30903                   x -> a x mod (a 2^16 - 1)
// The host tabulates the 140 jumps once per (permuteseed, draws_per_subcolumn) (KissJump); the
// numbers drawn are bit-identical to the sequential stream.  One CTA = 8 consecutive columns taken one after the
// other (their mask words then leave as whole 32-byte sectors), one warp = 32 sub-columns, so a layer's 32 cloud bits
// are one __ballot_sync.
struct KissJump {
    unsigned a1[NGPT], c1[NGPT];   // s1_n = a1 * s1_0 + c1
    unsigned p3[NGPT], p4[NGPT];   // a^(n-1) mod m (applied after one explicit step; n = 0 -> unused)
    unsigned n[NGPT];              // the jump distance itself
    unsigned w2[NGPT][8][16];      // xorshift matrix power as 4-bit window tables: y = XOR_k w2[.][k][(x >> 4k) & 15]
                                   // (8 lookups instead of 32 conditional XORs of matrix columns)
};
constexpr unsigned KISS_M3 = 18000u * 65536u - 1u, KISS_M4 = 30903u * 65536u - 1u;
// n-step jumps for n = 1 .. LW_MAXLAY (seed-independent, built once per device): a run of n cloud-free layers draws n
// numbers that cannot reach the mask (icld 1: every layer's draw is compared with 1 - cldfrac = 1 only; icld 2: the
// value a clear layer hands upward is never kept, as prev > 1 - 0 cannot hold), so the stream is stepped over them
struct KissRun {
    unsigned a1[LW_MAXLAY + 1], c1[LW_MAXLAY + 1];
    unsigned p3[LW_MAXLAY + 1], p4[LW_MAXLAY + 1];  // a^(n-1) mod m
    unsigned w2[LW_MAXLAY + 1][8][16];
};
#ifndef LW_KISS_JUMP_MIN
#define LW_KISS_JUMP_MIN 8  // shorter clear runs are stepped draw by draw (a jump costs about as much as 6 steps)
#endif
__device__ __forceinline__ void kiss_jump(Kiss& k, const KissRun* __restrict__ R, int n) {
    k.s1 = R->a1[n] * k.s1 + R->c1[n];
    unsigned y = 0;
    const unsigned z2 = k.s2;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) y ^= R->w2[n][kk][(z2 >> (4 * kk)) & 15u];
    k.s2 = y;
    unsigned t3 = 18000u * (k.s3 & 65535u) + (k.s3 >> 16);
    unsigned t4 = 30903u * (k.s4 & 65535u) + (k.s4 >> 16);
    if (t3 != KISS_M3) t3 = (unsigned)(((unsigned long long)R->p3[n] * t3) % KISS_M3);
    if (t4 != KISS_M4) t4 = (unsigned)(((unsigned long long)R->p4[n] * t4) % KISS_M4);
    k.s3 = t3; k.s4 = t4;
}

constexpr int KISS_CPB = 8;  // columns per CTA: a CTA's mask words of one (layer, word) are one 32-byte sector
// One work item = KISS_CPB consecutive columns starting at i0.  FUSE_CLD: the cloud optics of those columns' cloudy layers
// (cldoptics_point) follow from the mask words still in shared memory (needs colprep's secdiff in place).
template <bool FUSE_CLD>
__device__ __forceinline__ void kiss_item(const ColIn& in, const Scratch& s, const RunCfg& r, const KissJump* __restrict__ J,
                                          const KissRun* __restrict__ R, unsigned* mk_sm, const int i0) {
    const int nlay = r.nlay;
    const int W5 = nlay * 5;
    unsigned* words = mk_sm;                                     // KISS_CPB x nlay x 5 mask words
    int* lst = reinterpret_cast<int*>(mk_sm + KISS_CPB * W5);    // KISS_CPB x nlay: the layers a column draws at, ascending
    int* nl = lst + KISS_CPB * nlay;                             // KISS_CPB: their number
    unsigned* top = reinterpret_cast<unsigned*>(nl + KISS_CPB);  // KISS_CPB: cloud top
    const long P = s.pitch;
    const int icld = r.icld;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int sc = w * 32 + lane;
    const bool act = sc < NGPT;
    const double cldmin = 1.0e-20;
    // icld 1 / 2 (R given): only layers with a cloud fraction draw numbers that can reach the mask; the stream is
    // stepped (short runs) or jumped (long runs) over the draws of the cloud-free layers between them
    const bool sparse = (icld == 1 || icld == 2) && R != nullptr;
    for (int k = tid; k < KISS_CPB * W5; k += blockDim.x) words[k] = 0u;
    if (tid < KISS_CPB) { nl[tid] = 0; top[tid] = 0u; }
    __syncthreads();
    for (int c = w; c < KISS_CPB; c += 5) {  // one warp lists a column's drawing layers
        const int i = i0 + c;
        if (i >= r.ncol) continue;
        const long ci = r.c0 + i;
        int cnt = 0;
        for (int l0 = 0; l0 < nlay; l0 += 32) {
            const int l = l0 + lane;
            const bool on = l < nlay && (!sparse || in.cldfrac[ci + in.ld * l] >= cldmin);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) lst[c * nlay + cnt + __popc(m & ((1u << lane) - 1u))] = l;
            cnt += __popc(m);
        }
        if (lane == 0) nl[c] = cnt;
    }
    __syncthreads();
    for (int c = 0; c < KISS_CPB; ++c) {
        const int i = i0 + c;
        if (i >= r.ncol) break;
        const long ci = r.c0 + i;
        const double p1 = in.play[ci] * 1.e2, p2 = in.play[ci + in.ld] * 1.e2;
        if (p1 < p2) {  // the reference STOPs (mcica:465); the column is left cloud-free
            if (tid == 0) raise(s, 4, ci, 1);
            continue;
        }
        Kiss k;
        if (act) {
            const double p3 = in.play[ci + 2 * in.ld] * 1.e2, p4 = in.play[ci + 3 * in.ld] * 1.e2;
            const unsigned z1 = (unsigned)(int)((p1 - (double)(int)p1) * 1000000000.0);
            const unsigned z2 = (unsigned)(int)((p2 - (double)(int)p2) * 1000000000.0);
            const unsigned z3 = (unsigned)(int)((p3 - (double)(int)p3) * 1000000000.0);
            const unsigned z4 = (unsigned)(int)((p4 - (double)(int)p4) * 1000000000.0);
            k.s1 = J->a1[sc] * z1 + J->c1[sc];
            unsigned y = 0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) y ^= J->w2[sc][kk][(z2 >> (4 * kk)) & 15u];
            k.s2 = y;
            if (J->n[sc] == 0u) {
                k.s3 = z3; k.s4 = z4;
            } else {
                unsigned t3 = 18000u * (z3 & 65535u) + (z3 >> 16);
                unsigned t4 = 30903u * (z4 & 65535u) + (z4 >> 16);
                if (t3 != KISS_M3) t3 = (unsigned)(((unsigned long long)J->p3[sc] * t3) % KISS_M3);
                if (t4 != KISS_M4) t4 = (unsigned)(((unsigned long long)J->p4[sc] * t4) % KISS_M4);
                k.s3 = t3; k.s4 = t4;
            }
        } else {
            k.s1 = k.s2 = k.s3 = k.s4 = 0u;
        }
        double prev = 0., rmax = 0., cfm = 0.;
        if (icld == 3 && act) rmax = k.next();
        const int n_draw = nl[c];
        int lprev = -1;
        for (int q = 0; q < n_draw; ++q) {
            const int l = lst[c * nlay + q];
            const int gap = l - lprev - 1;  // cloud-free layers stepped over (sparse only; uniform over the CTA)
            if (gap > 0) {
                if (gap >= LW_KISS_JUMP_MIN) kiss_jump(k, R, gap);
                else for (int j = 0; j < gap; ++j) k.step();
                prev = 0.;  // any value < 1: the layer above a clear one never keeps it (prev > 1 - 0 cannot hold)
                cfm = 0.;
            }
            lprev = l;
            double cf = in.cldfrac[ci + in.ld * l];
            if (cf < cldmin) cf = 0.;
            double x;
            if (icld == 3) {
                x = rmax;
            } else {
                x = k.next();
                if (icld >= 4) {
                    const double x2 = k.next();
                    if (l > 0 && x2 < in.alpha[ci + in.ld * l]) x = prev;
                } else if (icld == 2 && l > 0) {
                    if (prev > 1. - cfm) x = prev; else x = x * (1. - cfm);
                }
            }
            prev = x;
            cfm = cf;
            const unsigned word = __ballot_sync(0xffffffffu, act && (x >= 1. - cf));
            if (lane == 0) words[c * W5 + l * 5 + w] = word;
        }
    }
    __syncthreads();
    // per-layer cloud flag (icldlyr, bit 31 of word 4), column cloud top, cldprmc input validation, mask write-out:
    // the KISS_CPB columns' words of a (layer, word) leave as one 32-byte store
    for (int l = tid; l < nlay; l += blockDim.x) {
        unsigned any[KISS_CPB];
#pragma unroll
        for (int c = 0; c < KISS_CPB; ++c) {
            unsigned a = 0;
#pragma unroll
            for (int q = 0; q < 5; ++q) a |= words[c * W5 + l * 5 + q];
            any[c] = a;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            unsigned v[KISS_CPB];
#pragma unroll
            for (int c = 0; c < KISS_CPB; ++c) v[c] = words[c * W5 + l * 5 + q] | ((q == 4 && any[c]) ? 0x80000000u : 0u);
            uint4* dst = reinterpret_cast<uint4*>(s.mask + ((long)l * 5 + q) * P + i0);
            dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
            dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
        }
#pragma unroll
        for (int c = 0; c < KISS_CPB; ++c) {
            if (!any[c]) continue;
            const long ci = r.c0 + i0 + c;
            atomicMax(&top[c], (unsigned)(l + 1));
            const long a = ci + in.ld * l;
            const double ciwp = in.ciwp[a], clwp = in.clwp[a];
            bool taubig = false;
            if (in.tauc) for (int b = 0; b < 16; ++b) taubig |= in.tauc[b + 16 * a] >= cldmin;
            if ((ciwp + clwp) >= cldmin || taubig) {
                if (r.inflag == 1) raise(s, 1, ci, l + 1);
                if (r.inflag == 2) {
                    if (ciwp != 0.0) {
                        const double rad = in.rei[a];
                        if ((r.iceflag == 0 && rad < 10.0) || (r.iceflag == 1 && (rad < 13.0 || rad > 130.)) ||
                            (r.iceflag == 2 && (rad < 5.0 || rad > 131.0)) || (r.iceflag == 3 && (rad < 5.0 || rad > 140.0)))
                            raise(s, 2, ci, l + 1);
                    }
                    if (clwp != 0.0 && r.liqflag == 1) {
                        const double rad = in.rel[a];
                        if (rad < 2.5 || rad > 60.) raise(s, 3, ci, l + 1);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid < KISS_CPB && i0 + tid < r.ncol) s.ctop[i0 + tid] = (int)top[tid];
    if constexpr (FUSE_CLD) {
        // (column, layer) pairs of the item, column fastest: a warp's stores of a (layer, band) are 64-byte runs
        for (int p = tid; p < KISS_CPB * nlay; p += blockDim.x) {
            const int c = p % KISS_CPB, l = p / KISS_CPB;
            if (i0 + c >= r.ncol) continue;
            const unsigned* wp = words + c * W5 + l * 5;
            const unsigned w5[5] = {wp[0], wp[1], wp[2], wp[3], wp[4]};
            if ((w5[0] | w5[1] | w5[2] | w5[3] | w5[4]) == 0u) continue;
            cldoptics_point(in, s, r, i0 + c, l, w5);
        }
    }
}
__global__ void __launch_bounds__(160) mcica_kiss_par_kernel(ColIn in, Scratch s, RunCfg r, const KissJump* __restrict__ J,
                                                             const KissRun* __restrict__ R) {
    extern __shared__ unsigned mk_sm[];
    kiss_item<false>(in, s, r, J, R, mk_sm, (int)blockIdx.x * KISS_CPB);
}
// Work-queue form (set_option("mcica_beside", 1)): a grid of one CTA per SM, launched BEFORE taumol, stays resident in
// the registers taumol's CTAs cannot use (65 536 - 3 x 16 384 per SM) and generates masks + cloud optics in the issue
// slots the L1-bound band evaluations leave idle; a second launch of the same kernel with a full grid, ordered after
// taumol, takes whatever items are left on the shared counter.
#define LW_KISS_WORK_STATS 1
__device__ unsigned long long g_kiss_stats[8];
__global__ void __launch_bounds__(160, 8) mcica_kiss_work_kernel(ColIn in, Scratch s, RunCfg r, const KissJump* __restrict__ J,
                                                              const KissRun* __restrict__ R, unsigned* __restrict__ ctr,
                                                              const unsigned nitems, const int mop) {
    extern __shared__ unsigned mk_sm[];
    __shared__ unsigned sh_item;
#ifdef LW_KISS_WORK_STATS
    if (threadIdx.x == 0) {  // [0/2] first CTA start, [1/3] last CTA end (ns), [4] items taken when the mop-up launch began
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (atomicCAS(&g_kiss_stats[mop ? 2 : 0], 0ull, t) == 0ull && mop) g_kiss_stats[4] = *(volatile unsigned*)ctr;
    }
#endif
    for (;;) {
        __syncthreads();  // the previous item's shared words are free
        if (threadIdx.x == 0) sh_item = atomicAdd(ctr, 1u);
        __syncthreads();
        const unsigned it = sh_item;
        if (it >= nitems) break;
        kiss_item<true>(in, s, r, J, R, mk_sm, (int)(it * KISS_CPB));
    }
#ifdef LW_KISS_WORK_STATS
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(&g_kiss_stats[mop ? 3 : 1], t);
    }
#endif
}

// host side of the jumps ---------------------------------------------------------------------
namespace {
struct KissMap {               // n-step map of the four component generators
    unsigned a1, c1;           // affine map of s1
    unsigned m2[32];           // matrix columns of s2
    unsigned long long p3, p4; // multipliers of s3, s4 (mod m)
};
unsigned xs_step(unsigned v) { v ^= v << 13; v ^= v >> 17; v ^= v << 5; return v; }
unsigned m2_apply(const unsigned* m, unsigned x) {
    unsigned y = 0;
    for (int b = 0; b < 32; ++b) if ((x >> b) & 1u) y ^= m[b];
    return y;
}
// 4-bit window tables of a 32x32 GF(2) matrix given by its columns: w[k][v] = XOR of the columns 4k + b over the bits b of v
void m2_windows(const unsigned* m, unsigned (*w)[16]) {
    for (int k = 0; k < 8; ++k)
        for (unsigned v = 0; v < 16; ++v) {
            unsigned y = 0;
            for (int b = 0; b < 4; ++b) if ((v >> b) & 1u) y ^= m[4 * k + b];
            w[k][v] = y;
        }
}
KissMap kiss_identity() {
    KissMap k; k.a1 = 1u; k.c1 = 0u; k.p3 = 1; k.p4 = 1;
    for (int b = 0; b < 32; ++b) k.m2[b] = 1u << b;
    return k;
}
KissMap kiss_one() {
    KissMap k; k.a1 = 69069u; k.c1 = 1327217885u; k.p3 = 18000; k.p4 = 30903;
    for (int b = 0; b < 32; ++b) k.m2[b] = xs_step(1u << b);
    return k;
}
// first y, then x
KissMap kiss_compose(const KissMap& x, const KissMap& y) {
    KissMap k;
    k.a1 = x.a1 * y.a1; k.c1 = x.a1 * y.c1 + x.c1;
    for (int b = 0; b < 32; ++b) k.m2[b] = m2_apply(x.m2, y.m2[b]);
    k.p3 = (x.p3 * y.p3) % KISS_M3; k.p4 = (x.p4 * y.p4) % KISS_M4;
    return k;
}
KissMap kiss_power(unsigned long long n) {
    KissMap res = kiss_identity(), base = kiss_one();
    while (n) {
        if (n & 1ull) res = kiss_compose(base, res);
        base = kiss_compose(base, base);
        n >>= 1;
    }
    return res;
}
static inline int cur_dev_kj() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 16) d = 0;
    return d;
}
struct KissJumpCache {
    KissJump* dev = nullptr;
    long key_seed = -1, key_draws = -1;
} g_kj_d[16];
#define g_kj g_kj_d[cur_dev_kj()]
}  // namespace

// Host-only check of the jump algebra against plain stepping: returns the number of mismatching
// generator states over a few (distance, seed) pairs (0 = all n-step jumps reproduce n single steps).
int kiss_jump_selftest() {
    const unsigned long long dist[] = {0ull, 1ull, 2ull, 71ull, 72ull, 144ull, 1000ull, 10079ull, 20161ull, 123457ull};
    const unsigned seeds[][4] = {{123456789u, 362436069u, 521288629u, 916191069u}, {0u, 0u, 0u, 0u},
                                 {999999999u, 1u, 65535u, 65536u}, {250000000u, 777777777u, 1179647998u, 2025259006u}};
    int bad = 0;
    for (unsigned long long n : dist) {
        const KissMap m = kiss_power(n);
        const KissMap m1 = kiss_power(n > 0 ? n - 1 : 0);
        for (const auto& z : seeds) {
            unsigned s1 = z[0], s2 = z[1], s3 = z[2], s4 = z[3];
            for (unsigned long long k = 0; k < n; ++k) {
                s1 = 69069u * s1 + 1327217885u;
                s2 = xs_step(s2);
                s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
                s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
            }
            const unsigned j1 = m.a1 * z[0] + m.c1;
            unsigned j2 = m2_apply(m.m2, z[1]);
            {   // the device applies the matrix through its 4-bit window tables
                unsigned w[8][16], y = 0;
                m2_windows(m.m2, w);
                for (int k = 0; k < 8; ++k) y ^= w[k][(z[1] >> (4 * k)) & 15u];
                if (y != j2) j2 = ~s2;  // counted as a mismatch below
            }
            unsigned j3 = z[2], j4 = z[3];
            if (n > 0) {
                j3 = 18000u * (z[2] & 65535u) + (z[2] >> 16);
                j4 = 30903u * (z[3] & 65535u) + (z[3] >> 16);
                if (j3 != KISS_M3) j3 = (unsigned)((m1.p3 * j3) % KISS_M3);
                if (j4 != KISS_M4) j4 = (unsigned)((m1.p4 * j4) % KISS_M4);
            }
            bad += (j1 != s1) + (j2 != s2) + (j3 != s3) + (j4 != s4);
        }
    }
    return bad;
}

// fills (and caches) the device jump table for jumps permuteseed + isubcol * draws, isubcol = 0..139
static const KissJump* kiss_jump_table(int permuteseed, long draws, cudaStream_t st) {
    if (!g_kj.dev) {
        if (cudaMalloc((void**)&g_kj.dev, sizeof(KissJump)) != cudaSuccess) return nullptr;
    }
    if (g_kj.key_seed == permuteseed && g_kj.key_draws == draws) return g_kj.dev;
    static KissJump h;  // host staging (pageable; the copy below is synchronous with respect to the host buffer)
    const unsigned long long n0 = permuteseed > 0 ? (unsigned long long)permuteseed : 0ull;
    // multipliers are tabulated for n - 1 steps (the kernel takes the first MWC step explicitly, which
    // maps any 32-bit seed into the range where the modular form is exact)
    const KissMap stepD = kiss_power((unsigned long long)draws);
    KissMap cur = kiss_power(n0);
    const KissMap inv_guard = kiss_identity();
    (void)inv_guard;
    for (int sc = 0; sc < NGPT; ++sc) {
        const unsigned long long n = n0 + (unsigned long long)sc * (unsigned long long)draws;
        h.a1[sc] = cur.a1; h.c1[sc] = cur.c1;
        m2_windows(cur.m2, h.w2[sc]);
        h.n[sc] = n > 0xffffffffull ? 0xffffffffu : (unsigned)n;
        if (n == 0) { h.p3[sc] = 1u; h.p4[sc] = 1u; }
        else {
            const KissMap m1 = kiss_power(n - 1);
            h.p3[sc] = (unsigned)m1.p3; h.p4[sc] = (unsigned)m1.p4;
        }
        cur = kiss_compose(stepD, cur);
    }
    if (cudaMemcpyAsync(g_kj.dev, &h, sizeof(KissJump), cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    cudaStreamSynchronize(st);
    g_kj.key_seed = permuteseed; g_kj.key_draws = draws;
    return g_kj.dev;
}

// seed-independent n-step jump table of the sparse path, one per device
static const KissRun* kiss_run_table(cudaStream_t st) {
    static KissRun* dev[16] = {};
    KissRun*& d = dev[cur_dev_kj()];
    if (d) return d;
    static KissRun h;
    KissMap cur = kiss_identity();          // n - 1 steps
    const KissMap one = kiss_one();
    for (int n = 1; n <= LW_MAXLAY; ++n) {
        h.p3[n] = (unsigned)cur.p3; h.p4[n] = (unsigned)cur.p4;
        cur = kiss_compose(one, cur);       // n steps
        h.a1[n] = cur.a1; h.c1[n] = cur.c1;
        m2_windows(cur.m2, h.w2[n]);
    }
    h.a1[0] = 1u; h.c1[0] = 0u; h.p3[0] = 1u; h.p4[0] = 1u;
    {
        unsigned id[32];
        for (int b = 0; b < 32; ++b) id[b] = 1u << b;
        m2_windows(id, h.w2[0]);
    }
    if (cudaMalloc((void**)&d, sizeof(KissRun)) != cudaSuccess) { d = nullptr; return nullptr; }
    if (cudaMemcpyAsync(d, &h, sizeof(KissRun), cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    cudaStreamSynchronize(st);
    return d;
}

// Cloud flags for the strict rrtmg_lw() entry: one warp scans the 140 sub-column fractions of a
// (column, layer); icldlyr = any(cldfmc == 1) (rtrnmc.f90:316-329); validates like cldprmc.
__global__ void __launch_bounds__(128) cldflags_kernel(ColIn in, Scratch s, RunCfg r) {
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long total = (long)r.ncol * r.nlay;
    if (wid >= total) return;
    const int l = (int)(wid / r.ncol);
    const int i = (int)(wid % r.ncol);
    const long ci = r.c0 + i;
    const long a = ci + in.ld * l;
    const double* cf = in.cldfmcl + 140 * a;
    bool one = false;
    int bad = 0;
    for (int g = lane; g < 140; g += 32) {
        const double c = cf[g];
        one |= (c == 1.);
        if (c >= 1.e-20) {
            const double ciwp = in.ciwpmcl[140 * a + g], clwp = in.clwpmcl[140 * a + g];
            if ((ciwp + clwp) >= 1.e-20 || in.taucmcl[140 * a + g] >= 1.e-20) {
                if (r.inflag == 1) bad = 1;
                if (r.inflag == 2) {
                    if (ciwp != 0.0) {
                        const double rad = in.reicmcl[a];
                        if ((r.iceflag == 0 && rad < 10.0) || (r.iceflag == 1 && (rad < 13.0 || rad > 130.)) ||
                            (r.iceflag == 2 && (rad < 5.0 || rad > 131.0)) || (r.iceflag == 3 && (rad < 5.0 || rad > 140.0)))
                            bad = 2;
                    }
                    if (clwp != 0.0 && r.liqflag == 1) {
                        const double rad = in.relqmcl[a];
                        if ((rad < 2.5 || rad > 60.) && bad == 0) bad = 3;
                    }
                }
            }
        }
    }
    const unsigned any = __ballot_sync(0xffffffffu, one);
    const unsigned anybad = __ballot_sync(0xffffffffu, bad != 0);
    if (anybad) {
        const int src = __ffs(anybad) - 1;
        const int code = __shfl_sync(0xffffffffu, bad, src);
        if (lane == 0) raise(s, code, ci, l + 1);
    }
    if (lane == 0) {
        s.mask[((long)l * 5 + 4) * s.pitch + i] = any ? 0x80000000u : 0u;
        if (any) atomicMax(&s.ctop[i], l + 1);
    }
}

// Strict rrtmg_lw() on the two-phase path: everything the sweeps need from the caller's per-g cloud arrays, computed with
// the arrays read the way they lie in memory ((140, ncol, nlay): a column's 140 sub-columns are one contiguous run).
// CTA = 16 consecutive columns of one layer, thread = sub-column g.  Per (column, layer): the cldfmc == 1 flags as 140
// mask bits + the icldlyr flag (rtrnmc.f90:316-329), the column's cloud top, cldprmc's validation (STOP sites), and, for
// layers in which any of the 16 columns holds cloud, odcld = secdiff * taucmc and efclfrac = 1 - exp(-odcld) of every
// sub-column (cldprmc.f90:177-266, rtrnmc.f90:331-345), transposed through shared memory so that the [g][layer][column]
// scratch is written in 128-byte runs.
constexpr int CA_COLS = 16;
__global__ void __launch_bounds__(160) cldarrays_kernel(ColIn in, Scratch s, RunCfg r) {
    __shared__ double sh_ef[NGPT][CA_COLS + 1], sh_oc[NGPT][CA_COLS + 1];
    __shared__ unsigned sh_w[CA_COLS][5];
    __shared__ int sh_any;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int g = tid;
    const int l = blockIdx.y;
    const int i0 = blockIdx.x * CA_COLS;
    const long P = s.pitch;
    const long plane = (long)r.nlay * P;
    double* S_EF = s.gcld;
    double* S_OC = s.gcld + (long)NGPT * plane;
    const int goff[NBND] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138};
    int b = 0;
#pragma unroll
    for (int k = 1; k < NBND; ++k) b += (g >= goff[k]);
    if (tid == 0) sh_any = 0;
    __syncthreads();
    for (int c = 0; c < CA_COLS; ++c) {
        const int i = i0 + c;
        bool on = false;
        int bad = 0;
        double ef = 0., odc = 0.;
        if (i < r.ncol && g < NGPT) {
            const long ci = r.c0 + i;
            const long a = ci + in.ld * l;
            const double cf = in.cldfmcl[140 * a + g];
            if (cf >= 1.e-20) {
                const double ciwp = in.ciwpmcl[140 * a + g], clwp = in.clwpmcl[140 * a + g];
                const double tauc_in = in.taucmcl[140 * a + g];
                if ((ciwp + clwp) >= 1.e-20 || tauc_in >= 1.e-20) {
                    if (r.inflag == 1) bad = 1;
                    if (r.inflag == 2) {
                        if (ciwp != 0.0) {
                            const double rad = in.reicmcl[a];
                            if ((r.iceflag == 0 && rad < 10.0) || (r.iceflag == 1 && (rad < 13.0 || rad > 130.)) ||
                                (r.iceflag == 2 && (rad < 5.0 || rad > 131.0)) || (r.iceflag == 3 && (rad < 5.0 || rad > 140.0)))
                                bad = 2;
                        }
                        if (clwp != 0.0 && r.liqflag == 1) {
                            const double rad = in.relqmcl[a];
                            if ((rad < 2.5 || rad > 60.) && bad == 0) bad = 3;
                        }
                    }
                }
                if (cf == 1.) {
                    on = true;
                    const double tauc = cld_tauc_band(b, ciwp, clwp, in.reicmcl[a], in.relqmcl[a], tauc_in, r);
                    odc = s.secdiff[(long)b * P + i] * tauc;
                    ef = (1. - exp(-odc)) * cf;
                }
            }
        }
        const unsigned word = __ballot_sync(0xffffffffu, on);
        const unsigned anybad = __ballot_sync(0xffffffffu, bad != 0);
        if (anybad) {
            const int src = __ffs(anybad) - 1;
            const int code = __shfl_sync(0xffffffffu, bad, src);
            if (lane == 0) raise(s, code, r.c0 + i, l + 1);
        }
        if (lane == 0) {
            sh_w[c][w] = word;
            if (word) sh_any = 1;
        }
        if (g < NGPT) { sh_ef[g][c] = ef; sh_oc[g][c] = odc; }
    }
    __syncthreads();
    // mask words (bit 31 of word 4 = icldlyr), cloud top
    if (tid < CA_COLS * 5) {
        const int c = tid / 5, q = tid % 5;
        const int i = i0 + c;
        if (i < r.ncol) {
            const unsigned any = sh_w[c][0] | sh_w[c][1] | sh_w[c][2] | sh_w[c][3] | sh_w[c][4];
            s.mask[((long)l * 5 + q) * P + i] = sh_w[c][q] | ((q == 4 && any) ? 0x80000000u : 0u);
            if (q == 0 && any) atomicMax(&s.ctop[i], l + 1);
        }
    }
    if (!sh_any) return;
    // per-g rows of this layer: 16 columns = 128 contiguous bytes
    for (int k = tid; k < NGPT * CA_COLS; k += blockDim.x) {
        const int gg = k / CA_COLS, c = k % CA_COLS;
        const int i = i0 + c;
        if (i < r.ncol) {
            const long o = (long)gg * plane + (long)l * P + i;
            S_EF[o] = sh_ef[gg][c];
            S_OC[o] = sh_oc[gg][c];
        }
    }
}

__global__ void zero_int_kernel(int* p, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// flux close-out: x fluxfac, net flux, heating rate; clear = total where no cloud was seen.
// rtrnmc.f90:569-593 and the output index shift of rad.f90:572-590.
__global__ void __launch_bounds__(128) finalize_kernel(ColIn in, ColOut out, Scratch s, RunCfg r, int cloudy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const bool sep = cloudy && s.ctop[i] > 0;
    // band sums of the column-lane kernels: [column / 128][level][6][column % 128] (lw_band.cuh)
    const double* fb = s.fluxB + ((long)(i >> 7) * (nlay + 1) * 6) * 128 + (i & 127);
    double fnet_lo = 0., fnetc_lo = 0., pz_lo = 0.;
    for (int lev = 0; lev <= nlay; ++lev) {
        const long o = ci + out.ld * lev;
        const double* f = fb + (long)lev * 6 * 128;
        const double u = f[256] * C.fluxfac;
        const double d = f[0] * C.fluxfac;
        double uc, dc;
        if (sep) { uc = f[384] * C.fluxfac; dc = f[128] * C.fluxfac; }
        else { uc = u; dc = d; }
        out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = uc; out.dflxc[o] = dc;
        if (r.idrv == 1) {
            const double du = f[512];
            out.duflx_dt[o] = du;
            out.duflxc_dt[o] = sep ? f[640] : du;
        }
        const double fnet = u - d, fnetc = uc - dc;
        const double pz = in.plev[ci + in.ld * lev];
        if (lev > 0) {
            const long oh = ci + out.ld * (lev - 1);
            out.hr[oh] = C.heatfac * (fnet_lo - fnet) / (pz_lo - pz);
            out.hrc[oh] = C.heatfac * (fnetc_lo - fnetc) / (pz_lo - pz);
        }
        fnet_lo = fnet; fnetc_lo = fnetc; pz_lo = pz;
    }
}

// flux close-out of the two-phase path: the sweep launches left their band sums in separate slots of s.fluxP
// ([column / 128][level][slot][6][128], slot = first band of a launch); `slots` = bit mask of the slots written.
// Slots are added in band order starting from 0.0: the reference's summation order (rtrnmc.f90:538-563).
__global__ void __launch_bounds__(128) finalize_slots_kernel(ColIn in, ColOut out, Scratch s, RunCfg r, int cloudy, unsigned slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.ncol) return;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const bool sep = cloudy && s.ctop[i] > 0;
    const double* fb = s.fluxP + ((long)(i >> 7) * (nlay + 1) * NBND * 6) * 128 + (i & 127);
    double fnet_lo = 0., fnetc_lo = 0., pz_lo = 0.;
    for (int lev = 0; lev <= nlay; ++lev) {
        const long o = ci + out.ld * lev;
        const double* f = fb + (long)lev * NBND * 6 * 128;
        double su = 0., sd = 0., suc = 0., sdc = 0., sdu = 0., sduc = 0.;
#pragma unroll
        for (int b = 0; b < NBND; ++b) {
            if ((slots >> b) & 1u) {
                const double* q = f + b * 6 * 128;
                sd = sd + q[0];
                su = su + q[256];
                if (sep) { sdc = sdc + q[128]; suc = suc + q[384]; }
                if (r.idrv == 1) {
                    sdu = sdu + q[512];
                    if (sep) sduc = sduc + q[640];
                }
            }
        }
        const double u = su * C.fluxfac;
        const double d = sd * C.fluxfac;
        double uc, dc;
        if (sep) { uc = suc * C.fluxfac; dc = sdc * C.fluxfac; }
        else { uc = u; dc = d; }
        out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = uc; out.dflxc[o] = dc;
        if (r.idrv == 1) {
            out.duflx_dt[o] = sdu;
            out.duflxc_dt[o] = sep ? sduc : sdu;
        }
        const double fnet = u - d, fnetc = uc - dc;
        const double pz = in.plev[ci + in.ld * lev];
        if (lev > 0) {
            const long oh = ci + out.ld * (lev - 1);
            out.hr[oh] = C.heatfac * (fnet_lo - fnet) / (pz_lo - pz);
            out.hrc[oh] = C.heatfac * (fnetc_lo - fnetc) / (pz_lo - pz);
        }
        fnet_lo = fnet; fnetc_lo = fnetc; pz_lo = pz;
    }
}

// ---------------------------------------------------------------------------------------------
// get_alpha, src/mcica_subcol_gen_lw.f90:129-178
__global__ void get_alpha_kernel(int ncol, long ld, int nlay, int icld, int idcor, double decorr_con, const double* dz,
                                 const double* lat, int juldat, const double* cldfrac, double* alpha, double am1,
                                 double am2, double am4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncol) return;
    if (icld != 4 && icld != 5) return;
    double decorr_len;
    if (idcor == 1) {
        const double amr = 7.0;
        double am3;
        if (juldat > 181) am3 = -4. * amr / 365. * (juldat - 272);
        else am3 = 4. * amr / 365. * (juldat - 91);
        const double d = lat[i] - am3;
        const double decorr_lat = am1 + am2 * exp(-(d * d) / (am4 * am4));
        decorr_len = decorr_lat * 1.e3;
    } else {
        decorr_len = decorr_con;
    }
    double decorr_inv = 1.0;
    if (decorr_len >= 0.0) decorr_inv = 1.0 / decorr_len;
    alpha[i] = 0.0;
    for (int k = 1; k < nlay; ++k) {
        double a = exp(-(0.5 * (dz[i + ld * k] + dz[i + ld * (k - 1)])) * decorr_inv);
        if (icld == 5 && cldfrac[i + ld * k] == 0.0 && cldfrac[i + ld * (k - 1)] > 0.0) a = 0.0;
        alpha[i + ld * k] = a;
    }
}

// Sub-column arrays of mcica_subcol_lw from the packed mask (mcica:658-679): one thread per
// (g, column, layer) element, g fastest -> coalesced 8-byte stores.
__global__ void materialize_kernel(ColIn in, Scratch s, RunCfg r, double* cldfmcl, double* ciwpmcl, double* clwpmcl,
                                   double* reicmcl, double* relqmcl, double* taucmcl) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)r.ncol * r.nlay * NGPT;
    if (t >= total) return;
    const int g = (int)(t % NGPT);
    const long cl = t / NGPT;
    const int i = (int)(cl % r.ncol);
    const int l = (int)(cl / r.ncol);
    const long ci = r.c0 + i;
    const long a = ci + in.ld * l;
    const unsigned w = s.mask[((long)l * 5 + (g >> 5)) * s.pitch + i];
    const bool on = (w >> (g & 31)) & 1u;
    const long o = g + 140 * a;
    // 0-based band of g-point g (ngb, src/rrtmg_lw_init.f90:362-377)
    constexpr int goff[NBND] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138};
    int band = 0;
#pragma unroll
    for (int b = 1; b < NBND; ++b) band += (g >= goff[b]);
    cldfmcl[o] = on ? 1. : 0.;
    clwpmcl[o] = on ? in.clwp[a] : 0.;
    ciwpmcl[o] = on ? in.ciwp[a] : 0.;
    taucmcl[o] = on ? in.tauc[band + 16 * a] : 0.;
    if (g == 0) { reicmcl[a] = in.rei[a]; relqmcl[a] = in.rel[a]; }
}

// ---------------------------------------------------------------------------------------------

void launch_colprep(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    colprep_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, s, r);
}
void launch_cldoptics(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    const dim3 grid((unsigned)nblk(r.ncol, 128), (unsigned)r.nlay);
    cldoptics_kernel<<<grid, 128, 0, st>>>(in, s, r);
}
void launch_prep(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    prep_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, s, r);
}
int g_kiss_sparse = 1;  // set_option("kiss_sparse", 0) draws every number (the round-1 kernel), for A/B timing
void set_kiss_sparse(int on) { g_kiss_sparse = on; }
void launch_mcica_kiss(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    const long draws = r.icld == 3 ? 1 : (r.icld >= 4 ? 2L * r.nlay : (long)r.nlay);
    const KissJump* J = kiss_jump_table(r.permuteseed, draws, st);
    if (!J) { mcica_kiss_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, s, r); return; }
    const size_t smem = ((size_t)KISS_CPB * r.nlay * 6 + 2 * KISS_CPB) * sizeof(unsigned);
    if (smem > 48 * 1024) {
        static size_t attr_d[16] = {};
        size_t& cur = attr_d[cur_dev_kj()];
        if (smem > cur) {
            cudaFuncSetAttribute(mcica_kiss_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cur = smem;
        }
    }
    const KissRun* R = g_kiss_sparse ? kiss_run_table(st) : nullptr;
    mcica_kiss_par_kernel<<<(unsigned)nblk(r.ncol, KISS_CPB), 160, smem, st>>>(in, s, r, J, R);
}
// Work-queue form of the KISS generator with the cloud optics fused (mcica_kiss_work_kernel).  resident > 0: resets the
// counter and launches `resident` CTAs per SM (the launch that runs beside taumol: 2 x 160 threads x 48 registers fit the
// 16 384 registers three taumol CTAs leave); resident = 0: a full grid that takes the items still on the counter.  Returns false when the jump tables are not available (the caller uses the grid form).
__device__ unsigned g_kiss_ctr;
bool launch_mcica_kiss_work(const ColIn& in, const Scratch& s, const RunCfg& r, int resident, cudaStream_t st) {
    const long draws = r.icld == 3 ? 1 : (r.icld >= 4 ? 2L * r.nlay : (long)r.nlay);
    const KissJump* J = kiss_jump_table(r.permuteseed, draws, st);
    if (!J) return false;
    const size_t smem = ((size_t)KISS_CPB * r.nlay * 6 + 2 * KISS_CPB) * sizeof(unsigned);
    if (smem > 48 * 1024) {
        static size_t attr_d[16] = {};
        size_t& cur = attr_d[cur_dev_kj()];
        if (smem > cur) {
            cudaFuncSetAttribute(mcica_kiss_work_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cur = smem;
        }
    }
    static int nsm_d[16] = {};
    int& nsm = nsm_d[cur_dev_kj()];
    if (!nsm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    const KissRun* R = g_kiss_sparse ? kiss_run_table(st) : nullptr;
    unsigned* ctr = nullptr;
    cudaGetSymbolAddress((void**)&ctr, g_kiss_ctr);
    const unsigned nitems = (unsigned)nblk(r.ncol, KISS_CPB);
    if (resident > 0) {
        cudaMemsetAsync(ctr, 0, sizeof(unsigned), st);
#ifdef LW_KISS_WORK_STATS
        void* stats = nullptr;
        cudaGetSymbolAddress(&stats, g_kiss_stats);
        cudaMemsetAsync(stats, 0, sizeof(g_kiss_stats), st);
#endif
    }
    unsigned grid = (unsigned)nsm * (resident > 0 ? (unsigned)resident : 8u);
    if (grid > nitems) grid = nitems;
    mcica_kiss_work_kernel<<<grid, 160, smem, st>>>(in, s, r, J, R, ctr, nitems, resident > 0 ? 0 : 1);
    return true;
}
void kiss_work_stats(unsigned long long* out8) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_kiss_stats, sizeof(g_kiss_stats));
}
// dst[e + w*(i + ldd*k)] = (float) src[e + w*(c0 + i + lds*k)]   for e < w, i < ncol, k < nrows
__global__ void cvt_kernel(const double* __restrict__ src, float* __restrict__ dst, int w, long lds, long ldd, long c0,
                           int ncol, long nrows) {
    const long n = (long)w * ncol;
    for (long k = blockIdx.y; k < nrows; k += gridDim.y)
        for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
            dst[t + (long)w * ldd * k] = (float)src[t + (long)w * (c0 + lds * k)];
}

int launch_cvt_inputs(const ColIn& in, const Scratch& s, const RunCfg& r, int cloudy, int inflag, ColInF* out,
                      cudaStream_t st) {
    const long P = s.pitch;
    const int nlay = r.nlay;
    float* p = s.inf;
    int launches = 0;
    auto cvt = [&](const double* src, int w, long nrows) -> const float* {
        float* d = p;
        p += (size_t)w * P * nrows;
        dim3 g((unsigned)nblk((long)w * r.ncol, 256), (unsigned)(nrows < 65535 ? nrows : 65535));
        cvt_kernel<<<g, 256, 0, st>>>(src, d, w, in.ld, P, r.c0, r.ncol, nrows);
        ++launches;
        return d - (long)w * r.c0;  // the band kernels index with c0 + i
    };
    ColInF f{};
    f.ld = P; f.nlay = nlay;
    f.play = cvt(in.play, 1, nlay);
    f.tlay = cvt(in.tlay, 1, nlay);
    f.tlev = cvt(in.tlev, 1, nlay + 1);
    f.tsfc = cvt(in.tsfc, 1, 1);
    f.emis = cvt(in.emis, 1, 16);
    f.tauaer = cvt(in.tauaer, 1, (long)nlay * 16);
    if (cloudy) {
        f.ciwp = cvt(in.ciwp, 1, nlay); f.clwp = cvt(in.clwp, 1, nlay);
        f.rei = cvt(in.rei, 1, nlay); f.rel = cvt(in.rel, 1, nlay);
        // tauc only enters the arithmetic when the cloud optical depth is prescribed (inflag 0); with inflag 2 it only
        // takes part in the "is there any cloud" test, which the float copy answers identically
        f.tauc = in.tauc ? cvt(in.tauc, 16, nlay) : nullptr;
        (void)inflag;
    }
    *out = f;
    return launches;
}

void launch_cldflags(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    zero_int_kernel<<<nblk(r.ncol, 256), 256, 0, st>>>(s.ctop, r.ncol);
    cldflags_kernel<<<nblk((long)r.ncol * r.nlay * 32, 128), 128, 0, st>>>(in, s, r);
}
void launch_cldarrays(const ColIn& in, const Scratch& s, const RunCfg& r, cudaStream_t st) {
    zero_int_kernel<<<nblk(r.ncol, 256), 256, 0, st>>>(s.ctop, r.ncol);
    const dim3 grid((unsigned)nblk(r.ncol, CA_COLS), (unsigned)r.nlay);
    cldarrays_kernel<<<grid, 160, 0, st>>>(in, s, r);
}
void launch_finalize(const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r, int cloudy, cudaStream_t st) {
    finalize_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, out, s, r, cloudy);
}
void launch_finalize_slots(const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r, int cloudy, unsigned slots,
                           cudaStream_t st) {
    finalize_slots_kernel<<<nblk(r.ncol, 128), 128, 0, st>>>(in, out, s, r, cloudy, slots);
}
void launch_get_alpha(int ncol, long ld, int nlay, int icld, int idcor, double decorr_con, const double* dz,
                      const double* lat, int juldat, const double* cldfrac, double* alpha, double am1, double am2,
                      double am4, cudaStream_t st) {
    get_alpha_kernel<<<nblk(ncol, 128), 128, 0, st>>>(ncol, ld, nlay, icld, idcor, decorr_con, dz, lat, juldat, cldfrac,
                                                      alpha, am1, am2, am4);
}
void launch_materialize(const ColIn& in, const Scratch& s, const RunCfg& r, double* cldfmcl, double* ciwpmcl,
                        double* clwpmcl, double* reicmcl, double* relqmcl, double* taucmcl, cudaStream_t st) {
    const long total = (long)r.ncol * r.nlay * NGPT;
    materialize_kernel<<<nblk(total, 256), 256, 0, st>>>(in, s, r, cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl);
}

}  // namespace lw

// ---------------------------------------------------------------------------------------------
// FP64 pipe peak: 8 independent DFMA chains per thread.  Used by bench.py as the measured FP64
// roofline denominator (MEASURED_PEAKS.json carries HBM and bf16 only).
namespace lw {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}
// returns TFLOP/s counting an FMA as 2 flops, best of `reps`
int measure_fp64_peak(double* tflops, int reps) {
    double* d = nullptr;
    if (cudaMalloc((void**)&d, 8) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
}  // namespace lw
