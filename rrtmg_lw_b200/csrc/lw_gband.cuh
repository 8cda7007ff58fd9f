// lw_gband.cuh -- the g-lane band kernel: the g-points of one (column, band) are spread over the
// lanes of a "unit", two adjacent g-points per thread.
//
// Same physics as lw_band.cuh (taugbN -> cldprmc -> rtrnmc of one band; src/rrtmg_lw_taumol.f90,
// src/rrtmg_lw_cldprmc.f90:164-271, src/rrtmg_lw_rtrnmc.f90:251-563), mapped for the memory system:
//   * a warp holds UPW = 32/LPU columns ("units"), LPU = lanes per unit (NG/2 rounded up to
//     2/4/8); lane j of a unit owns g-points 2j, 2j+1, so one k-table row (stored g-fastest,
//     lw_tables.h) is ONE contiguous 32..128-byte access per unit made of 16-byte loads, instead
//     of 32 scattered lines per load when lanes are columns;
//   * the g-independent work of a layer (binary-species interpolation weights, continuum and
//     minor-gas factors, Planck terms, cloud optics) is NOT replicated across the unit's lanes:
//     in the "S-step" lane k of a unit does it for layer (top - k) of the current block of LPU
//     layers and writes a record into the warp's shared-memory ring; in the "G-step" all lanes
//     of the unit read the record of the current layer (a broadcast) and evaluate their g-points;
//   * the exp/tfn transmittance table (10 001 x 16 B) lives in shared memory: a gathered lookup
//     costs bank conflicts, not one L1 tag pass per 128-byte line;
//   * the vertical recurrences (radld/radlu, total and clear stream, d/dTs) run in registers;
//     the per-(g,layer) transmittance/source pair needed again on the way up goes to a
//     per-resident-warp slot (CTAs are persistent, so the slots total a few tens of MB and
//     stay in L2), written and read as full 256-byte rows, read four layers ahead;
//   * the g-point -> band reduction is a butterfly over the unit's lanes; the lane that owns a
//     (level, stream) keeps the sum and adds it to the band-flux scratch, whose running value
//     it fetched at the start of the block;
//   * wide bands (LPU = 8) read layer-fastest copies of the per-layer inputs (a unit's S-step
//     reads 8 consecutive layers of one column), narrow bands (LPU <= 4) the column-fastest
//     originals (8..16 consecutive columns of one layer): every S-step load is >= 64 contiguous
//     bytes per unit group; the band-flux scratch follows the same choice.
// Bit-exactness policy as in lw_band.cuh: everything feeding a table index keeps the reference's
// association order (-fmad=false); only the g-sum order differs (tree instead of sequential,
// SURVEY T16).
#pragma once
#include <type_traits>

#include "lw_band.cuh"

namespace lw {

#ifndef GB_NWARPS
#define GB_NWARPS 8
#endif
constexpr int GB_WARPS = GB_NWARPS;
constexpr int GB_THREADS = GB_WARPS * 32;
#ifndef GB_ETAB_SMEM
#define GB_ETAB_SMEM 1   // 1: exp/tfn table in shared memory; 0: gathered through L1 (leaves ~160 KB of L1)
#endif
constexpr int GB_ETAB_BYTES = GB_ETAB_SMEM ? ((NTBL + 1) * 16 + 255) / 256 * 256 : 0;

template <int B>
struct GLanes {
    static constexpr int NG = NGC[B - 1];
    static constexpr int GPT = NG >= 4 ? 2 : 1;        // g-points per thread
    static constexpr int NH = NG / GPT;                // threads a unit needs
    static constexpr int LPU = NH > 4 ? 8 : (NH > 2 ? 4 : 2);  // lanes per unit
    static constexpr int UPW = 32 / LPU;               // units (columns) per warp
    static constexpr bool LAYER_FASTEST = LPU >= 8;    // which copy of the per-layer inputs / flux scratch is used
};

// per-band feature switches of the S-step record
template <int B>
struct GRec {
    static constexpr bool BIN_LO = NSPA[B - 1] == 9;
    static constexpr bool BIN_UP = NSPB[B - 1] == 5;
    static constexpr bool BIN = BIN_LO || BIN_UP;
    static constexpr int NW = BIN_LO ? 12 : (BIN_UP ? 8 : 5);  // weights: c0.w,c1.w | u[8] | f00,f10,f01,f11,keycol
    static constexpr bool MINOR = B == 1 || B == 3 || B == 5 || B == 6 || B == 7 || B == 8 || B == 9 || B == 11 ||
                                  B == 13 || B == 15;
    static constexpr bool MINOR_ETA = B == 3 || B == 5 || B == 7 || B == 9 || B == 13 || B == 15;
    static constexpr bool MUL1 = B == 8 || B == 13;
    static constexpr bool MUL2 = B == 8;
    static constexpr bool WXA = B == 5 || B == 6 || B == 8;
    static constexpr bool WXB = B == 6 || B == 8;
    static constexpr bool CORR = B == 1 || B == 2;
    // slot numbering
    static constexpr int S_W = 0;
    static constexpr int S_COMB = S_W + NW;                 // 2 if BIN
    static constexpr int S_SELF = S_COMB + (BIN ? 2 : 0);   // selffac, selffrac, forfac, forfrac
    static constexpr int S_FPL = S_SELF + 4;                // 1 if BIN
    static constexpr int S_MIN = S_FPL + (BIN ? 1 : 0);     // mul0, minorfrac
    static constexpr int S_FM0 = S_MIN + (MINOR ? 2 : 0);
    static constexpr int S_MUL1 = S_FM0 + (MINOR_ETA ? 1 : 0);
    static constexpr int S_FM1 = S_MUL1 + (MUL1 ? 1 : 0);   // band 13 only
    static constexpr int S_MUL2 = S_FM1 + (B == 13 ? 1 : 0);
    static constexpr int S_WXA = S_MUL2 + (MUL2 ? 1 : 0);
    static constexpr int S_WXB = S_WXA + (WXA ? 1 : 0);
    static constexpr int S_CORR = S_WXB + (WXB ? 1 : 0);
    static constexpr int S_INTS = S_CORR + (CORR ? 1 : 0);  // two packed 32-bit words
    static constexpr int S_PLK = S_INTS + 1;                // blay, dplankup, dplankdn, taua
    static constexpr int S_CLD = S_PLK + 4;                 // odcld, abscld, {cloud word, spare}
    static constexpr int NSLOT = S_CLD + 3;
};
constexpr int GB_MAXSLOT = 32;
constexpr int GB_RING_POS = 32;                   // records a warp's ring holds
constexpr int GB_RING_DOUBLES = GB_MAXSLOT * GB_RING_POS;  // per warp
constexpr size_t GB_SMEM_BYTES = (size_t)GB_ETAB_BYTES + (size_t)GB_WARPS * GB_RING_DOUBLES * sizeof(double);

// layer-fastest accessor of the prep fields for BandLayer::setup
struct LayerT {
    const double* base;  // fldT + i*LP + l0
    long fstride;        // pitch * LP
    __device__ __forceinline__ double f(int field) const { return __ldg(base + (long)field * fstride); }
};

// cldprmc for one (column, layer, band): cloud optical depth of a cloudy sub-column, cldprmc.f90:177-266
template <int B>
__device__ __forceinline__ double cld_tauc(const RunCfg& r, double ciwp, double clwp, double tauc, double radice,
                                           double radliq) {
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

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}

// transposing butterfly: every lane of an aligned group of N lanes holds N quantities q[0..N-1];
// afterwards q[0] of lane k (k = lane % N) is the sum over the group's lanes of quantity k.
template <int N>
__device__ __forceinline__ void xreduce(double (&q)[N], int g) {
#pragma unroll
    for (int s = N / 2; s >= 1; s >>= 1) {
        const bool hi = (g & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const double send = hi ? q[k] : q[k + s];
            const double keep = hi ? q[k + s] : q[k];
            q[k] = keep + shfl_xor_d(send, s);
        }
    }
}

// pack / unpack of the S-step record (ring column p, stride 32 doubles between slots)
template <int B>
__device__ __forceinline__ void rec_store(double* p, const BandLayer<B>& W) {
    using R = GRec<B>;
    auto put = [&](int slot, double v) { p[slot * GB_RING_POS] = v; };
    int rowa = 0, rowb = 0, na = 0, nb = 0;
    if (W.lower) {
        if constexpr (R::BIN_LO) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { put(R::S_W + k, W.c0.w[k]); put(R::S_W + 6 + k, W.c1.w[k]); }
            rowa = W.c0.row; rowb = W.c1.row; na = W.c0.n == 6; nb = W.c1.n == 6;
        } else {
            put(R::S_W + 0, W.f00); put(R::S_W + 1, W.f10); put(R::S_W + 2, W.f01); put(R::S_W + 3, W.f11);
            put(R::S_W + 4, W.keycol);
            rowa = W.r0; rowb = W.r1;
        }
    } else {
        if constexpr (R::BIN_UP) {
#pragma unroll
            for (int k = 0; k < 8; ++k) put(R::S_W + k, W.u[k]);
        } else {
            put(R::S_W + 0, W.f00); put(R::S_W + 1, W.f10); put(R::S_W + 2, W.f01); put(R::S_W + 3, W.f11);
            put(R::S_W + 4, W.keycol);
        }
        rowa = W.r0; rowb = W.r1;
    }
    if constexpr (R::BIN) { put(R::S_COMB, W.s.comb); put(R::S_COMB + 1, W.s1.comb); put(R::S_FPL, W.fpl); }
    put(R::S_SELF, W.selffac); put(R::S_SELF + 1, W.selffrac); put(R::S_SELF + 2, W.forfac); put(R::S_SELF + 3, W.forfrac);
    if constexpr (R::MINOR) { put(R::S_MIN, W.mul0); put(R::S_MIN + 1, W.minorfrac); }
    if constexpr (R::MINOR_ETA) put(R::S_FM0, W.fm0);
    if constexpr (R::MUL1) put(R::S_MUL1, W.mul1);
    if constexpr (B == 13) put(R::S_FM1, W.fm1);
    if constexpr (R::MUL2) put(R::S_MUL2, W.mul2);
    if constexpr (R::WXA) put(R::S_WXA, W.wxa);
    if constexpr (R::WXB) put(R::S_WXB, W.wxb);
    if constexpr (R::CORR) put(R::S_CORR, W.corradj);
    const int wa = (rowa & 4095) | ((rowb & 4095) << 12) | (na << 24) | (nb << 25) | ((W.lower ? 1 : 0) << 26);
    const int wb = (W.inds & 15) | ((W.indf & 3) << 4) | ((W.indm & 31) << 6) | ((W.jm0 & 15) << 11) |
                   ((W.jm1 & 15) << 15) | ((W.jpl & 15) << 19);
    put(R::S_INTS, __hiloint2double(wb, wa));
}

template <int B>
__device__ __forceinline__ void rec_load(const double* p, BandLayer<B>& W) {
    using R = GRec<B>;
    auto get = [&](int slot) { return p[slot * GB_RING_POS]; };
    const double iw = get(R::S_INTS);
    const int wa = __double2loint(iw), wb = __double2hiint(iw);
    const int rowa = wa & 4095, rowb = (wa >> 12) & 4095;
    W.lower = ((wa >> 26) & 1) != 0;
    W.inds = wb & 15; W.indf = (wb >> 4) & 3; W.indm = (wb >> 6) & 31; W.jm0 = (wb >> 11) & 15;
    W.jm1 = (wb >> 15) & 15; W.jpl = (wb >> 19) & 15;
    W.r0 = rowa; W.r1 = rowb;
    if constexpr (R::BIN_LO) {
        // lower: corner weights; upper (BIN_UP or single species): u[] / fac overlay the same slots
#pragma unroll
        for (int k = 0; k < 6; ++k) { W.c0.w[k] = get(R::S_W + k); W.c1.w[k] = get(R::S_W + 6 + k); }
        W.c0.row = rowa; W.c1.row = rowb;
        W.c0.n = ((wa >> 24) & 1) ? 6 : 4;
        W.c1.n = ((wa >> 25) & 1) ? 6 : 4;
        if constexpr (R::BIN_UP) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { W.u[k] = W.c0.w[k]; }
            W.u[6] = W.c1.w[0]; W.u[7] = W.c1.w[1];
        } else {
            W.f00 = W.c0.w[0]; W.f10 = W.c0.w[1]; W.f01 = W.c0.w[2]; W.f11 = W.c0.w[3]; W.keycol = W.c0.w[4];
        }
    } else if constexpr (R::BIN_UP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) W.u[k] = get(R::S_W + k);
        W.f00 = W.u[0]; W.f10 = W.u[1]; W.f01 = W.u[2]; W.f11 = W.u[3]; W.keycol = W.u[4];
    } else {
        W.f00 = get(R::S_W + 0); W.f10 = get(R::S_W + 1); W.f01 = get(R::S_W + 2); W.f11 = get(R::S_W + 3);
        W.keycol = get(R::S_W + 4);
    }
    if constexpr (R::BIN) { W.s.comb = get(R::S_COMB); W.s1.comb = get(R::S_COMB + 1); W.fpl = get(R::S_FPL); }
    W.selffac = get(R::S_SELF); W.selffrac = get(R::S_SELF + 1); W.forfac = get(R::S_SELF + 2); W.forfrac = get(R::S_SELF + 3);
    if constexpr (R::MINOR) { W.mul0 = get(R::S_MIN); W.minorfrac = get(R::S_MIN + 1); }
    if constexpr (R::MINOR_ETA) W.fm0 = get(R::S_FM0);
    if constexpr (R::MUL1) W.mul1 = get(R::S_MUL1);
    if constexpr (B == 13) W.fm1 = get(R::S_FM1);
    if constexpr (R::MUL2) W.mul2 = get(R::S_MUL2);
    if constexpr (R::WXA) W.wxa = get(R::S_WXA);
    if constexpr (R::WXB) W.wxb = get(R::S_WXB);
    if constexpr (R::CORR) W.corradj = get(R::S_CORR);
}


// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ double comp(const D2& v, int k) { return k == 0 ? v.x : v.y; }
__device__ __forceinline__ double comp(double v, int) { return v; }

// sum over the LPU lanes of a unit of NS interleaved quantities: on return every lane holds the
// unit total of quantity (j % NS) (NS = 1 or 2), j = lane index inside the unit
template <int LPU, int NS>
__device__ __forceinline__ double unit_sum(double a, double b, int j) {
    double v;
    if constexpr (NS == 2) {
        const bool hi = (j & 1) != 0;
        const double send = hi ? a : b;
        const double keep = hi ? b : a;
        v = keep + shfl_xor_d(send, 1);
#pragma unroll
        for (int m = 2; m < LPU; m <<= 1) v = v + shfl_xor_d(v, m);
    } else {
        v = a;
#pragma unroll
        for (int m = 1; m < LPU; m <<= 1) v = v + shfl_xor_d(v, m);
    }
    return v;
}

template <int B, int CLD>
__global__ void __launch_bounds__(GB_THREADS, 1) gband_kernel(ColIn in, Scratch s, RunCfg r) {
    using GL = GLanes<B>;
    using R = GRec<B>;
    constexpr int NG = GL::NG, GPT = GL::GPT, LPU = GL::LPU, UPW = GL::UPW;
    constexpr bool LF = GL::LAYER_FASTEST;
    constexpr int G0 = GOFF[B - 1];
    constexpr int NSD = CLD != CLD_NONE ? 2 : 1;  // streams per level: total (+ clear)
    constexpr int SB = GB_RING_POS / UPW;         // layers per S-step ( = LPU with 8 warps)
    constexpr int LPR = LPU / NSD;                // levels per flush: lane j owns (level j / NSD, stream j % NSD)
    constexpr int UB = LPR > 4 ? LPR : 4;         // levels per batch of state loads on the way up
    constexpr int RS = GB_RING_POS;
    constexpr bool FIRST = B == 1 || B == 6;      // first band writing its flux scratch (layer-fastest: 1, column-fastest: 6)
    using V = typename std::conditional<GPT == 2, D2, double>::type;
    static_assert(R::NSLOT <= GB_MAXSLOT, "ring record too large");
    static_assert(SB == LPU && SB % LPR == 0 && UB % LPR == 0, "block geometry");
    extern __shared__ __align__(16) unsigned char gb_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ring = reinterpret_cast<double*>(gb_smem + GB_ETAB_BYTES) + warp * GB_RING_DOUBLES;
#if GB_ETAB_SMEM
    double2* etab = reinterpret_cast<double2*>(gb_smem);
    for (int k = threadIdx.x; k <= NTBL; k += GB_THREADS) etab[k] = __ldg(D.exptfn + k);
    __syncthreads();
#else
    const double2* etab = D.exptfn;
#endif

    const int u = lane / LPU, j = lane % LPU;
    const int g0 = j * GPT;          // first g-point of this thread
    const bool gact = g0 < NG;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const int LP = s.LP, LP1 = s.LP1;
    const long plane = (long)nlay * P;
    const double rec_6 = 0.166667;
    const double bpade = C.bpade;
    const double* tp = D.totplnk + 181 * (B - 1);
    const double delw = C.delwave[B - 1];
    const int gw = blockIdx.x * GB_WARPS + warp;
    const int total_warps = gridDim.x * GB_WARPS;
    // state slots of this resident warp: [l0][2*GPT][32] and (cloudy layers) [l0][4*GPT][32]
    double* st0 = s.gstate + (size_t)gw * nlay * 128 + lane;
    double* stB = s.gstateB + (size_t)gw * nlay * 256 + lane;
    const long ngroups = (r.ncol + UPW - 1) / UPW;
    // ring position of (unit u, record k) = k*UPW + u: the UPW units read consecutive 8-byte words
    double* ring_w = ring + (j * UPW + u);
    const double* ring_r = ring + u;
    const int og = j / NSD, ost = j % NSD;  // flush ownership of this lane

    auto planck = [&](double t) {
        int it = (int)(t - 159.);
        it = it < 1 ? 1 : (it > 180 ? 180 : it);
        const double frac = t - 159. - (double)it;
        const double a = ldg(tp + it - 1), b = ldg(tp + it);
        const double dbdt = b - a;
        return a + frac * dbdt;
    };

    for (long grp = gw; grp < ngroups; grp += total_warps) {
        const long i_raw = grp * UPW + u;
        const bool live = i_raw < r.ncol;
        const int i = live ? (int)i_raw : r.ncol - 1;
        const long ci = r.c0 + i;
        const bool dbg = live && gact && (s.dbg_col == i);

        // diffusivity angle for this band, rtrnmc.f90:273-281
        double secdiff;
        if constexpr (B == 1 || B == 4 || B >= 10) {
            secdiff = 1.66;
        } else {
            secdiff = C.a0[B - 1] + C.a1[B - 1] * exp(C.a2[B - 1] * s.pwvcm[i]);
            if (secdiff > 1.80) secdiff = 1.80;
            if (secdiff < 1.50) secdiff = 1.50;
        }
        const int laytrop = s.laytrop[i];
        const int ctop = (CLD != CLD_NONE) ? s.ctop[i] : 0;
        // surface emission, setcoef.f90:211-219 / rtrnmc.f90:465-469
        const double semiss = in.emis[ci + in.ld * (B - 1)];
        double plankbnd, dplankbnd_dt = 0.;
        {
            const double tbound = in.tsfc[ci];
            int it = (int)(tbound - 159.);
            it = it < 1 ? 1 : (it > 180 ? 180 : it);
            const double frac = tbound - 159. - (double)it;
            const double a = ldg(tp + it - 1), b = ldg(tp + it);
            const double dbdt = b - a;
            plankbnd = semiss * (a + frac * dbdt);
            if (r.idrv == 1) {
                const double* td = D.totplnkderiv + 181 * (B - 1);
                const double da = ldg(td + it - 1), db = ldg(td + it);
                const double ddbdt = db - da;
                dplankbnd_dt = semiss * (da + frac * ddbdt);
            }
        }
        const double reflect = 1. - semiss;
        // band-flux scratch: element (stream, level) of this column
        auto fxp = [&](int stream, int level) -> double* {
            if constexpr (LF) return s.fluxT + ((long)stream * P + i) * LP1 + level;
            else return s.fluxC + ((long)stream * (nlay + 1) + level) * P + i;
        };
        // L2 prefetch of the per-layer inputs of the S-block that starts at layer `top`
        auto prefetch_block = [&](int top) {
            const int lev_s = top - j;
            if (lev_s < 1) return;
            const int l0 = lev_s - 1;
            const long a = ci + in.ld * l0;
            prefetch_l2(in.tlay + a);
            prefetch_l2(in.tlev + a);
            prefetch_l2(in.tauaer + ci + in.ld * (l0 + (long)nlay * (B - 1)));
            constexpr unsigned fmask = band_fields(B);
            if constexpr (LF) {
                // one line of a field row covers 16 layers: lane j fetches fields j, j + LPU, ...
                const int lb = top - 1 - (SB - 1) < 0 ? 0 : top - SB;
                for (int f = j; f < NFIELD; f += LPU)
                    if (fmask & (1u << f)) prefetch_l2(s.fldT + ((long)f * P + i) * LP + lb);
            } else {
#pragma unroll
                for (int f = 0; f < NFIELD; ++f)
                    if (fmask & (1u << f)) prefetch_l2(s.fld + (long)f * plane + (long)l0 * P + i);
            }
        };

        double radld[GPT], radclrd[GPT], pf_sfc[GPT];
#pragma unroll
        for (int k = 0; k < GPT; ++k) { radld[k] = 0.; radclrd[k] = 0.; pf_sfc[k] = 0.; }

        // ------------------------------------------------------------ downward sweep, rtrnmc.f90:350-455
        double acc = 0., oldf = 0.;
        int flev = 0;  // level this lane flushes next
        for (int top = nlay; top >= 1; top -= SB) {
            // ---- S-step: lane j prepares layer top - j of its unit
            {
                const int lev_s = top - j;
                if (lev_s >= 1) {
                    const int l0 = lev_s - 1;
                    const long a = ci + in.ld * l0;
                    double pavel = 0.;
                    if constexpr (B == 1 || B == 2) pavel = in.play[a];
                    BandLayer<B> W;
                    W.jm0 = 0; W.jm1 = 0; W.jpl = 0; W.inds = 0; W.indf = 0; W.indm = 0; W.r0 = 0; W.r1 = 0;
                    unsigned w4 = 0, mlo = 0, mhi = 0;
                    constexpr int wlo = G0 >> 5, sh = G0 & 31;
                    if constexpr (LF) {
                        const LayerT L{s.fldT + (long)i * LP + l0, P * LP};
                        W.setup(L, s.ipackT[(long)i * LP + l0], lev_s <= laytrop, pavel);
                        if constexpr (CLD != CLD_NONE) {
                            w4 = s.maskT[((long)4 * P + i) * LP + l0];
                            if constexpr (CLD == CLD_MASK) {
                                mlo = s.maskT[((long)wlo * P + i) * LP + l0];
                                if constexpr (sh + NG > 32) mhi = s.maskT[((long)(wlo + 1) * P + i) * LP + l0];
                            }
                        }
                    } else {
                        const LayerRef L{s, (long)l0 * P + i, plane};
                        W.setup(L, s.ipack[(long)l0 * P + i], lev_s <= laytrop, pavel);
                        if constexpr (CLD != CLD_NONE) {
                            const unsigned* mw = s.mask + ((long)l0 * 5) * P + i;
                            w4 = mw[4 * P];
                            if constexpr (CLD == CLD_MASK) {
                                mlo = mw[(long)wlo * P];
                                if constexpr (sh + NG > 32) mhi = mw[(long)(wlo + 1) * P];
                            }
                        }
                    }
                    rec_store<B>(ring_w, W);
                    const double blay = planck(in.tlay[a]);
                    const double plev_dn = planck(in.tlev[a]);            // planklev(lev-1)
                    const double plev_up = planck(in.tlev[a + in.ld]);    // planklev(lev)
                    ring_w[(R::S_PLK + 0) * RS] = blay;
                    ring_w[(R::S_PLK + 1) * RS] = plev_up - blay;
                    ring_w[(R::S_PLK + 2) * RS] = plev_dn - blay;
                    ring_w[(R::S_PLK + 3) * RS] = in.tauaer[ci + in.ld * (l0 + (long)nlay * (B - 1))];
                    if constexpr (CLD != CLD_NONE) {
                        int cw = (w4 >> 31) ? 1 : 0;
                        if constexpr (CLD == CLD_MASK) {
                            double odcld_b = 0., abscld_b = 0.;
                            if (cw) {
                                unsigned lo = mlo >> sh;
                                if constexpr (sh + NG > 32) lo |= mhi << (32 - sh);
                                const unsigned bits = lo & ((1u << NG) - 1u);
                                if (bits) {
                                    const double ciwp = in.ciwp[a], clwp = in.clwp[a];
                                    const double tauc0 = in.tauc ? in.tauc[(B - 1) + 16 * a] : 0.0;
                                    const double tauc = cld_tauc<B>(r, ciwp, clwp, tauc0, in.rei[a], in.rel[a]);
                                    odcld_b = secdiff * tauc;
                                    abscld_b = 1. - exp(-odcld_b);
                                }
                                cw |= (int)(bits << 1);
                            }
                            ring_w[(R::S_CLD + 0) * RS] = odcld_b;
                            ring_w[(R::S_CLD + 1) * RS] = abscld_b;
                        }
                        ring_w[(R::S_CLD + 2) * RS] = __hiloint2double(0, cw);
                    }
                }
                if (top - SB >= 1) prefetch_block(top - SB);
            }
            __syncwarp();
            const int nin = top < SB ? top : SB;
            for (int jj = 0; jj < nin; ++jj) {
                const int lev = top - jj;
                const int l0 = lev - 1;
                if ((jj % LPR) == 0) {
                    // this lane will own (level lev - 1 - og, stream ost) of the coming flush: fetch the running band sum early
                    flev = lev - 1 - og;
                    if constexpr (!FIRST) {
                        if (live && flev >= 0) oldf = __ldcg(fxp(ost, flev));
                    }
                }
                // ---- G-step for layer lev
                double v_tot = 0., v_clr = 0.;
                {
                    const double* rp = ring_r + jj * UPW;
                    BandLayer<B> W;
                    rec_load<B>(rp, W);
                    const double blay = rp[(R::S_PLK + 0) * RS];
                    const double dplankup = rp[(R::S_PLK + 1) * RS];
                    const double dplankdn = rp[(R::S_PLK + 2) * RS];
                    const double taua = rp[(R::S_PLK + 3) * RS];
                    bool cldlyr = false;
                    unsigned bits = 0;
                    double odcld_b = 0., abscld_b = 0.;
                    if constexpr (CLD != CLD_NONE) {
                        const int cw = __double2loint(rp[(R::S_CLD + 2) * RS]);
                        cldlyr = (cw & 1) != 0;
                        if constexpr (CLD == CLD_MASK) {
                            bits = ((unsigned)cw) >> 1;
                            odcld_b = rp[(R::S_CLD + 0) * RS];
                            abscld_b = rp[(R::S_CLD + 1) * RS];
                        }
                    }
                    if (gact) {
                        V tgv, pfv;
                        W.template evalv<V>(g0, tgv, pfv);
#pragma unroll
                        for (int k = 0; k < GPT; ++k) {
                            const int g = g0 + k;
                            const double tg = comp(tgv, k), plfrac = comp(pfv, k);
                            const double taut = tg + taua;  // rad.f90:548-553
                            double odepth = secdiff * taut;
                            if (odepth < 0.0) odepth = 0.0;
                            int d0 = -1, d1 = -1;
                            double atrans, bbugas, bbd;
                            if (CLD != CLD_NONE && cldlyr) {
                                double cf, odc, ef;
                                if constexpr (CLD == CLD_MASK) {
                                    const bool on = (bits >> g) & 1u;
                                    cf = on ? 1. : 0.; odc = on ? odcld_b : 0.; ef = on ? abscld_b : 0.;
                                } else {
                                    // strict rrtmg_lw(): per-g cloud arrays (g fastest -> coalesced over the unit's lanes)
                                    const long a = ci + in.ld * l0;
                                    const long ag = (G0 + g) + 140 * a;
                                    cf = in.cldfmcl[ag];
                                    double tauc = in.taucmcl[ag];
                                    const double ciwp = in.ciwpmcl[ag], clwp = in.clwpmcl[ag];
                                    if (cf >= 1.e-20) tauc = cld_tauc<B>(r, ciwp, clwp, tauc, in.reicmcl[a], in.relqmcl[a]);
                                    if (cf == 1.) {
                                        odc = secdiff * tauc;
                                        ef = (1. - exp(-odc)) * cf;
                                    } else {
                                        odc = 0.; ef = 0.;
                                    }
                                }
                                // cloudy layer, rtrnmc.f90:358-425
                                double odtot = odepth + odc;
                                double atot, bbutot, gassrc, bbdtot;
                                if (odtot < 0.06) {
                                    atrans = odepth - 0.5 * odepth * odepth;
                                    const double odepth_rec = rec_6 * odepth;
                                    gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
                                    atot = odtot - 0.5 * odtot * odtot;
                                    const double odtot_rec = rec_6 * odtot;
                                    bbdtot = plfrac * (blay + dplankdn * odtot_rec);
                                    bbd = plfrac * (blay + dplankdn * odepth_rec);
                                    bbugas = plfrac * (blay + dplankup * odepth_rec);
                                    bbutot = plfrac * (blay + dplankup * odtot_rec);
                                } else if (odepth <= 0.06) {
                                    atrans = odepth - 0.5 * odepth * odepth;
                                    const double odepth_rec = rec_6 * odepth;
                                    gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
                                    odtot = odepth + odc;
                                    const double tblind = odtot / (bpade + odtot);
                                    const int ittot = (int)(10000.0 * tblind + 0.5);
                                    d1 = ittot;
                                    const double2 et = etab[ittot];
                                    const double tfactot = et.y;
                                    bbdtot = plfrac * (blay + tfactot * dplankdn);
                                    bbd = plfrac * (blay + dplankdn * odepth_rec);
                                    atot = 1. - et.x;
                                    bbugas = plfrac * (blay + dplankup * odepth_rec);
                                    bbutot = plfrac * (blay + tfactot * dplankup);
                                } else {
                                    double tblind = odepth / (bpade + odepth);
                                    const int itgas = (int)(10000.0 * tblind + 0.5);
                                    d0 = itgas;
                                    odepth = ldg(D.tau_tbl + itgas);
                                    const double2 eg = etab[itgas];
                                    atrans = 1. - eg.x;
                                    const double tfacgas = eg.y;
                                    gassrc = atrans * plfrac * (blay + tfacgas * dplankdn);
                                    odtot = odepth + odc;
                                    tblind = odtot / (bpade + odtot);
                                    const int ittot = (int)(10000.0 * tblind + 0.5);
                                    d1 = ittot;
                                    const double2 et = etab[ittot];
                                    const double tfactot = et.y;
                                    bbdtot = plfrac * (blay + tfactot * dplankdn);
                                    bbd = plfrac * (blay + tfacgas * dplankdn);
                                    atot = 1. - et.x;
                                    bbugas = plfrac * (blay + tfacgas * dplankup);
                                    bbutot = plfrac * (blay + tfactot * dplankup);
                                }
                                radld[k] = radld[k] - radld[k] * (atrans + ef * (1. - atrans)) + gassrc + cf * (bbdtot * atot - gassrc);
                                double* sB = stB + ((size_t)l0 * 4 * GPT + 4 * k) * 32;
                                __stcg(sB, atot);
                                __stcg(sB + 32, bbutot);
                                __stcg(sB + 64, ef);
                                __stcg(sB + 96, cf);
                            } else {
                                // clear layer, rtrnmc.f90:426-445
                                if (odepth <= 0.06) {
                                    atrans = odepth - 0.5 * odepth * odepth;
                                    odepth = rec_6 * odepth;
                                    bbd = plfrac * (blay + dplankdn * odepth);
                                    bbugas = plfrac * (blay + dplankup * odepth);
                                } else {
                                    const double tblind = odepth / (bpade + odepth);
                                    const int itr = (int)(10000.0 * tblind + 0.5);
                                    d0 = itr;
                                    const double2 e = etab[itr];
                                    atrans = 1. - e.x;
                                    const double tausfac = e.y;
                                    bbd = plfrac * (blay + tausfac * dplankdn);
                                    bbugas = plfrac * (blay + tausfac * dplankup);
                                }
                                radld[k] = radld[k] + (bbd - radld[k]) * atrans;
                            }
                            __stcg(st0 + ((size_t)l0 * 2 * GPT + 2 * k) * 32, atrans);
                            __stcg(st0 + ((size_t)l0 * 2 * GPT + 2 * k + 1) * 32, bbugas);
                            if (CLD != CLD_NONE) {
                                // clear-sky stream: identical to the total stream above the highest cloudy layer
                                if (lev <= ctop) radclrd[k] = radclrd[k] + (bbd - radclrd[k]) * atrans;
                                else radclrd[k] = radld[k];
                            }
                            if (lev == 1) pf_sfc[k] = plfrac;
                            if (dbg) {
                                const long o = (long)(G0 + g) * nlay + l0;
                                s.dbg_taug[o] = tg; s.dbg_fracs[o] = plfrac;
                                s.dbg_itr[2 * o] = d0; s.dbg_itr[2 * o + 1] = d1;
                            }
                        }
                        v_tot = radld[0];
                        v_clr = radclrd[0];
                        if constexpr (GPT == 2) { v_tot = v_tot + radld[1]; v_clr = v_clr + radclrd[1]; }
                    }
                }
                // g-sum of this level's radiances; the owning lane keeps it until the flush
                {
                    const double sum = unit_sum<LPU, NSD>(v_tot, v_clr, j);
                    if (og == (jj % LPR)) acc = sum;
                    if ((jj % LPR) == LPR - 1 || lev == 1) {
                        if (live && flev >= 0 && og <= (jj % LPR)) {
                            const double add = (acc * 0.5) * delw;
                            if constexpr (FIRST) __stcg(fxp(ost, flev), 0.0 + add); else __stcg(fxp(ost, flev), oldf + add);
                        }
                    }
                }
            }
            __syncwarp();
        }
        if constexpr (FIRST) {
            // downward flux at the top of the atmosphere is never touched by the sweep: drad(nlayers) = 0
            if (live && j < NSD) __stcg(fxp(j, nlay), 0.0);
        }

        // ------------------------------------------------------------ surface, rtrnmc.f90:465-484
        double radlu[GPT], radclru[GPT], d_radlu[GPT], d_radclru[GPT];
#pragma unroll
        for (int k = 0; k < GPT; ++k) {
            radlu[k] = 0.; radclru[k] = 0.; d_radlu[k] = 0.; d_radclru[k] = 0.;
            if (gact) {
                const double rad0 = pf_sfc[k] * plankbnd;
                radlu[k] = rad0 + reflect * radld[k];
                radclru[k] = rad0 + reflect * radclrd[k];
                d_radlu[k] = pf_sfc[k] * dplankbnd_dt;
                d_radclru[k] = d_radlu[k];
            }
        }
        // ------------------------------------------------------------ upward sweep, rtrnmc.f90:486-529
        // levels 0..nlay in batches of UB; level 0 is the surface (no layer update)
        double accd = 0., oldd = 0.;
        for (int base = 0; base <= nlay; base += UB) {
            double sa[UB][GPT], sb[UB][GPT];
            unsigned cfl[UB];
#pragma unroll
            for (int q = 0; q < UB; ++q) {
                const int lev = base + q;
                cfl[q] = 0;
                if (lev >= 1 && lev <= nlay) {
                    const int l0 = lev - 1;
                    if (gact) {
#pragma unroll
                        for (int k = 0; k < GPT; ++k) {
                            sa[q][k] = __ldcg(st0 + ((size_t)l0 * 2 * GPT + 2 * k) * 32);
                            sb[q][k] = __ldcg(st0 + ((size_t)l0 * 2 * GPT + 2 * k + 1) * 32);
                        }
                    }
                    if constexpr (CLD != CLD_NONE) {
                        if constexpr (LF) cfl[q] = s.maskT[((long)4 * P + i) * LP + l0];
                        else cfl[q] = s.mask[((long)l0 * 5 + 4) * P + i];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < UB; ++q) {
                const int lev = base + q;
                if (lev <= nlay) {  // warp-uniform
                    if ((q % LPR) == 0) {
                        flev = lev + og;
                        if constexpr (!FIRST) {
                            if (live && flev <= nlay) {
                                oldf = __ldcg(fxp(2 + ost, flev));
                                if (r.idrv == 1) oldd = __ldcg(fxp(4 + ost, flev));
                            }
                        }
                    }
                    if (lev >= 1 && gact) {
                        const int l0 = lev - 1;
                        const bool cldlyr = (CLD != CLD_NONE) && (cfl[q] >> 31) != 0;
#pragma unroll
                        for (int k = 0; k < GPT; ++k) {
                            const double atrans = sa[q][k], bbugas = sb[q][k];
                            if (CLD != CLD_NONE && cldlyr) {
                                const double* sB = stB + ((size_t)l0 * 4 * GPT + 4 * k) * 32;
                                const double atot = __ldcg(sB);
                                const double bbutot = __ldcg(sB + 32);
                                const double ef = __ldcg(sB + 64);
                                const double cf = __ldcg(sB + 96);
                                const double gassrc = bbugas * atrans;
                                radlu[k] = radlu[k] - radlu[k] * (atrans + ef * (1. - atrans)) + gassrc + cf * (bbutot * atot - gassrc);
                                if (r.idrv == 1)
                                    d_radlu[k] = d_radlu[k] * cf * (1.0 - atot) + d_radlu[k] * (1.0 - cf) * (1.0 - atrans);
                            } else {
                                radlu[k] = radlu[k] + (bbugas - radlu[k]) * atrans;
                                if (r.idrv == 1) d_radlu[k] = d_radlu[k] * (1.0 - atrans);
                            }
                            if (CLD != CLD_NONE) {
                                radclru[k] = radclru[k] + (bbugas - radclru[k]) * atrans;
                                if (r.idrv == 1) d_radclru[k] = d_radclru[k] * (1.0 - atrans);
                            }
                        }
                    }
                    const bool mine = og == (q % LPR);
                    const bool flush = (q % LPR) == LPR - 1 || lev == nlay;
                    const bool st = live && flev <= nlay && og <= (q % LPR);
                    double vt = radlu[0], vc = radclru[0];
                    if constexpr (GPT == 2) { vt = vt + radlu[1]; vc = vc + radclru[1]; }
                    const double sum = unit_sum<LPU, NSD>(vt, vc, j);
                    if (mine) acc = sum;
                    if (flush && st) {
                        const double add = (acc * 0.5) * delw;
                        if constexpr (FIRST) __stcg(fxp(2 + ost, flev), 0.0 + add); else __stcg(fxp(2 + ost, flev), oldf + add);
                    }
                    if (r.idrv == 1) {
                        double dt = d_radlu[0], dc = d_radclru[0];
                        if constexpr (GPT == 2) { dt = dt + d_radlu[1]; dc = dc + d_radclru[1]; }
                        const double sumd = unit_sum<LPU, NSD>(dt, dc, j);
                        if (mine) accd = sumd;
                        if (flush && st) {
                            const double add = (accd * 0.5) * delw * C.fluxfac;
                            if constexpr (FIRST) __stcg(fxp(4 + ost, flev), 0.0 + add); else __stcg(fxp(4 + ost, flev), oldd + add);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace lw
