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
// Planck-fraction etas and two rows of 4-byte words (icldlyr word, mask word lo, mask word hi).
template <int B0, int NB>
struct SweepRing {
    using G = Grp<B0, NB>;
    static constexpr int IPS = (int)(sizeof(double) / sizeof(int));
    static constexpr int S_TLAY = G::NGG, S_TLEV = G::NGG + 1, S_PF = G::NGG + 2;
    static constexpr int S_INT = S_PF + G::NPFG;
    static constexpr int SLOTS = S_INT + 4 / IPS;
#ifndef LW_SWEEP_L2PF
#define LW_SWEEP_L2PF 0  // layers beyond the ring's own prefetch distance whose state words are pulled into L2
#endif
#ifndef LW_SWEEP_NST
#define LW_SWEEP_NST 2
#endif
    static constexpr int NST = LW_SWEEP_NST;  // ring stages = layers of prefetch distance
    static constexpr size_t RING_BYTES = (size_t)NST * SLOTS * SWEEP_THREADS * sizeof(double);
    // Rejected variant, off by default (LW_SWEEP_CLR_SMEM_MIN=16 switches it on): groups of >= MIN g-points keep the
    // clear-sky recurrences of their first LW_SWEEP_CLR_SMEM_N g-points in shared memory ([g][thread], behind the ring)
    // instead of registers, so that a 16-g band fits 128 registers and four CTAs per SM.  Measured 35 % SLOWER on bands
    // 3 and 5 (3.21 vs 2.37 ms per 262 144 columns, profiles/r02_sweep_clear_stream_smem_experiment.log): the extra
    // shared-memory round trip sits on the recurrence's dependency chain.
#ifndef LW_SWEEP_CLR_SMEM_MIN
#define LW_SWEEP_CLR_SMEM_MIN 99
#endif
#ifndef LW_SWEEP_CLR_SMEM_N
#define LW_SWEEP_CLR_SMEM_N 8
#endif
    LW_HD static constexpr int clr_smem(int cld) { return (cld != CLD_NONE && G::NGG >= LW_SWEEP_CLR_SMEM_MIN) ? LW_SWEEP_CLR_SMEM_N : 0; }
    LW_HD static constexpr size_t bytes(int cld) {
        return RING_BYTES + (size_t)clr_smem(cld) * SWEEP_THREADS * sizeof(double);
    }
    static constexpr size_t BYTES = RING_BYTES;
};

// Planck function of band B at the temperature whose table position is (it, frac), setcoef.f90:173-226
// own16: the run is band 16 alone (istart = 16): setcoef then integrates its Planck function over 2600-3250 cm-1 only
// (totplk16, column 17 of the table) instead of 2600 cm-1 to infinity (setcoef.f90:233-251).  level0: planklev(0, 16) of
// that variant takes the VALUE from totplk16 and the SLOPE from totplnk(:,16) (setcoef.f90:244-246).
template <int B>
__device__ __forceinline__ double planck_at(const double* totplnk, int it, double frac, bool own16 = false, bool level0 = false) {
    const double* tp = totplnk + 181 * (B - 1);
    if (B == 16 && own16) {
        const double* t16 = totplnk + 181 * 16;
        const double a = ldg(t16 + it - 1);
        const double dbdt = level0 ? ldg(tp + it) - ldg(tp + it - 1) : ldg(t16 + it) - a;
        return fmadd(frac, dbdt, a);
    }
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
// Planck-fraction interpolation passed by phase 1 (jpl = 1 + int(sm), fpl = mod(sm, 1)).  The fraction of a g-point is
// fetched where it is used (one or two L1-resident loads) instead of being held in registers across the g loop: the
// registers buy more g-points in flight.  One straight-line form for every case: a + fpl * (b - a) with fpl = 0 and
// b = a where the band takes a g-constant (bit-identical: a + 0 * 0 == a), times 0 where the reference sets fracs = 0.
template <int B>
struct PlFrac {
    static constexpr int RS = NGPAD[B - 1];
    static constexpr bool BIN_LO = NSPA[B - 1] == 9, BIN_UP = NSPB[B - 1] == 5;
    static constexpr bool ANY_BIN = BIN_LO || BIN_UP;
    static constexpr bool HAS_ZERO = (B == 12 || B == 15);
    const double* p0;
    int dnext;     // element offset of the second row (0 = no interpolation)
    double fpl;
    bool zero;
    __device__ __forceinline__ void setup(const SweepTab& tb, bool lower, double sm) {
        struct { const double *fracrefa, *fracrefb; } T{tb.fracrefa[B - 1], tb.fracrefb[B - 1]};
        const int jpl = 1 + (int)sm;
        const double f = sm - (double)(int)sm;
        zero = false;
        if (lower) {
            if constexpr (BIN_LO) { p0 = T.fracrefa + (jpl - 1) * RS; dnext = RS; fpl = f; }
            else { p0 = T.fracrefa; dnext = 0; fpl = 0.; }
        } else {
            if constexpr (BIN_UP) { p0 = T.fracrefb + (jpl - 1) * RS; dnext = RS; fpl = f; }
            else if constexpr (B == 6) { p0 = T.fracrefa; dnext = 0; fpl = 0.; }
            else if constexpr (HAS_ZERO) { p0 = T.fracrefa; dnext = 0; fpl = 0.; zero = true; }
            else { p0 = T.fracrefb; dnext = 0; fpl = 0.; }
        }
    }
    __device__ __forceinline__ double get(int g) const {
        const double a = ldg(p0 + g);
        double v = a;
        if constexpr (ANY_BIN) {
            const double b = ldg(p0 + dnext + g);
            v = a + fpl * (b - a);
        }
        if constexpr (HAS_ZERO) v = zero ? 0.0 : v;
        return v;
    }
};

// itr = int(tblint * tblind + 0.5) (rtrnmc.f90:386, 404, 412), kept inside the table for out-of-range cloud inputs
// (a negative cloud optical depth would index the reference's tables out of bounds; here it must not fault the device)
__device__ __forceinline__ int tbl_index(double tblind) {
    const int it = (int)(10000.0 * tblind + 0.5);
    return it < 0 ? 0 : (it > NTBL ? NTBL : it);
}

// predicated gathers: lanes that do not need the value issue no memory access
__device__ __forceinline__ double2 ldg2_if(const double2* p, bool on) {
    double2 v;
#ifdef LW_F32
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; mov.f32 %0, 0f00000000; mov.f32 %1, 0f00000000;"
        " @q ld.global.nc.v2.f32 {%0,%1}, [%2]; }" : "=f"(v.x), "=f"(v.y) : "l"(p), "r"((int)on));
#else
    v.x = 0.; v.y = 0.;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.global.nc.v2.f64 {%0,%1}, [%2]; }" : "+d"(v.x), "+d"(v.y) : "l"(p), "r"((int)on));
#endif
    return v;
}
__device__ __forceinline__ double ldg1_if(const double* p, bool on) {
    double v;
#ifdef LW_F32
    asm("{ .reg .pred q; setp.ne.s32 q, %2, 0; mov.f32 %0, 0f00000000; @q ld.global.nc.f32 %0, [%1]; }"
        : "=f"(v) : "l"(p), "r"((int)on));
#else
    asm("{ .reg .pred q; setp.ne.s32 q, %2, 0; mov.f64 %0, 0d0000000000000000; @q ld.global.nc.f64 %0, [%1]; }"
        : "=d"(v) : "l"(p), "r"((int)on));
#endif
    return v;
}

// Gas-only transmittance and Planck source of one (g, layer) from the state word of phase 1 (lw_taumol.cuh
// od_state): the optical depth itself where the reference takes its small-optical-depth series (odepth <= 0.06,
// rtrnmc.f90:429-431), the table index itr where it takes the lookup (:433-440).  bb = plfrac * (blay + tausfac * dpl)
// is bbd on the way down (dpl = dplankdn) and bbugas on the way up (dpl = dplankup).  Selects, no branches: the
// g-points of a thread form independent instruction streams; the table gather is issued only by the lanes that use it.
__device__ __forceinline__ void gas_point(const double2* exptfn, double st, double plfrac, double blay, double dpl,
                                          double& atrans, double& bb, bool& big, int& itr) {
    const double rec_6 = 0.166667;
    big = od_state_is_index(st);
    itr = od_state_index(st);
    const double2 e = ldg2_if(exptfn + itr, big);
#ifdef LW_TAIL_FMA
    atrans = big ? 1. - e.x : fmadd(-0.5 * st, st, st);
#else
    atrans = big ? 1. - e.x : st - 0.5 * st * st;
#endif
    const double tausfac = big ? e.y : rec_6 * st;
    bb = plfrac * fmadd(tausfac, dpl, blay);
}

// Gas + cloud path of a sub-column that holds cloud (cldfmc = 1) in this layer, rtrnmc.f90:358-425: total optical depth
// odtot = odepth + odcld, where the gas part is the table value tau_tbl(itgas) once it has gone through the lookup
// (:387); atot / bb_tot from the series (odtot < 0.06) or from the lookup at ittot.  Lanes with on = false get
// finite don't-care values and touch no memory.
__device__ __forceinline__ void cloud_point(const SweepTab& tb, double st, bool big, int itr, double odc, bool on,
                                            double plfrac, double blay, double dpl, double& atot, double& bb_tot,
                                            int& ittot_out) {
    const double bpade = tb.bpade;
    const double rec_6 = 0.166667;
    const double odg = big ? ldg1_if(tb.tau_tbl + itr, on && big) : st;
    const double odtot = odg + odc;
    // the series only where the reference takes it: odtot < 0.06 with the gas optical depth itself (branch 1,
    // rtrnmc.f90:361-376); once the gas part has gone through the lookup (odepth > 0.06, :397-425) the total is looked
    // up too, whatever tau_tbl(itgas) + odcld comes to
    const bool tsmall = !big && odtot < 0.06;
    const double tblind = div_fast(odtot, bpade + odtot);
    const int ittot = tbl_index(tblind);
    const double2 et = ldg2_if(tb.exptfn + ittot, on && !tsmall);
    atot = tsmall ? odtot - 0.5 * odtot * odtot : 1. - et.x;
    const double tfactot = tsmall ? rec_6 * odtot : et.y;
    bb_tot = plfrac * fmadd(tfactot, dpl, blay);
    ittot_out = tsmall ? -1 : ittot;
}

// per-g values of a thread: the first NSM in the thread's column of a [g][thread] shared-memory block, the rest in
// registers (g is a compile-time constant wherever it is indexed: the loops are unrolled)
template <int N, int NSM>
struct HybArr {
    acc_t v[N - NSM > 0 ? N - NSM : 1];
    acc_t* p;
    __device__ __forceinline__ acc_t& operator[](int g) { return g < NSM ? p[g * SWEEP_THREADS] : v[g - NSM]; }
    __device__ __forceinline__ void bind(void* q) { p = reinterpret_cast<acc_t*>(q); }
};

// ------------------------------------------------------------------------------------------
// Experiment LW_SWEEP_TMA=1 (rejected variants): a layer's optical depths of the CTA's 128 columns are ONE
// contiguous run of NGG x 1 KB in phase 1's layout, so instead of 128 threads x NGG / 2 16-byte cp.async one elected
// thread issues one bulk copy of the TMA unit (cp.async.bulk, completion on an mbarrier with a transaction count) per
// ring stage; the consumers wait on the barrier's phase.  The price: the copy overwrites every thread's slots, so the
// stage has to be released by the whole CTA (__syncthreads per layer) where the per-thread cp.async needs no barrier.
#ifndef LW_SWEEP_TMA
#define LW_SWEEP_TMA 0
#endif
#if LW_SWEEP_TMA
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LW_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LW_DONE_%=;\n"
        "bra LW_WAIT_%=;\n"
        "LW_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
#endif

// the sweep kernel of bands B0 .. B0 + NB - 1
template <int B0, int NB, int CLD, int IDRV, int MINB>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB) rtrnmc_kernel(ColIn in, ColOut out, Scratch s, RunCfg r, const SweepTab tb) {
    using GR = Grp<B0, NB>;
    using RG = SweepRing<B0, NB>;
    constexpr int NGG = GR::NGG;
    constexpr int G0 = GR::G0;
    extern __shared__ __align__(128) double ring[];
    const int tid = threadIdx.x;
#if LW_SWEEP_TMA
    __shared__ __align__(8) unsigned long long mbar[RG::NST];
    if (tid == 0) {
        for (int k = 0; k < RG::NST; ++k) mbar_init(&mbar[k], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned mpar = 0u;  // phase parity per stage
#endif
    const int i_raw = blockIdx.x * blockDim.x + tid;
    // out-of-range threads of the last CTA shadow the last valid column (no early return: keeps all lanes of a warp alive)
    const bool live = i_raw < r.ncol;
    const int i = live ? i_raw : r.ncol - 1;
    const long ci = r.c0 + i;
    const int nlay = r.nlay;
    const long P = s.pitch;
    const long plane = (long)nlay * P;
    const bool dbg = live && (s.dbg_col == i);
    const int laytrop = s.laytrop[i];
    const int ctop = (CLD != CLD_NONE) ? s.ctop[i] : 0;
    const bool own16 = (B0 + NB - 1 == 16) && r.istart == 16;
    (void)own16;
    (void)out;

    // blocks of 128 columns, [column / 128][layer][...][column % 128]
    const long blk_lay0 = (long)(i >> 7) * nlay;
    const int blk_t = i & 127;
    auto od_src = [&](int l0) { return reinterpret_cast<const double2*>(s.od + ((blk_lay0 + l0) * NGPT + G0) * 128) + blk_t; };
    auto pf_src = [&](int l0) { return s.pfs + ((blk_lay0 + l0) * NPF + GR::PF0) * 128 + blk_t; };
    // {odcld, 1 - exp(-odcld)} of a (layer, band) whose sub-columns hold cloud (CLD_MASK): [block][layer][band][2][128]
    auto cldb = [&](int l0, int b) { return s.cldb + (((blk_lay0 + l0) * NBND + (b - 1)) * 2) * 128 + blk_t; };
    const long blk_lev0 = (long)(i >> 7) * (nlay + 1);
    // band sums of the level fluxes of THIS launch's bands: [column / 128][level][slot][6][column % 128], slot = first band
    // of the launch; 0 dflx, 1 dflxc, 2 uflx, 3 uflxc, 4 duflx_dt, 5 duflxc_dt.  Every launch owns its slot (no
    // read-modify-write, no ordering between the sweep kernels, which therefore run concurrently on several streams);
    // the finalize kernel adds the slots in band order, which is the reference's summation order (rtrnmc.f90:538-563).
    auto flx = [&](int lev) { return s.fluxP + (((blk_lev0 + lev) * NBND + (B0 - 1)) * 6) * 128 + blk_t; };
    // strict rrtmg_lw(): per-g {efclfrac, odcld} of cloudy layers
    double* S_EF = s.gcld;
    double* S_OC = s.gcld + (long)NGPT * plane;

    // ---- asynchronous prefetch of one layer's inputs into this thread's slots of ring stage `stage`.
    // DOWN: tlev = level below the layer (planklev(lev-1)); UP: tlev = level above.
    auto issue = [&](int stage, int lev, bool down) {
        if (lev >= 1 && lev <= nlay) {
            const int l0 = lev - 1;
            const long a = ci + in.ld * l0;
            double* dst = ring + (size_t)stage * RG::SLOTS * SWEEP_THREADS;
#if LW_SWEEP_TMA
            if (tid == 0) {  // one bulk copy for the CTA's optical depths of the layer, one for its Planck-fraction etas
                constexpr unsigned OD_BYTES = NGG * SWEEP_THREADS * sizeof(double);
                constexpr unsigned PF_BYTES = GR::NPFG * SWEEP_THREADS * sizeof(double);
                mbar_expect_tx(&mbar[stage], OD_BYTES + PF_BYTES);
                bulk_g2s(dst, s.od + ((blk_lay0 + l0) * NGPT + G0) * 128, OD_BYTES, &mbar[stage]);
                if constexpr (GR::NPFG > 0)
                    bulk_g2s(dst + RG::S_PF * SWEEP_THREADS, s.pfs + ((blk_lay0 + l0) * NPF + GR::PF0) * 128, PF_BYTES, &mbar[stage]);
            }
#else
            const double2* od = od_src(l0);
#pragma unroll
            for (int p = 0; p < NGG / 2; ++p) cp_async_pair(dst + (2 * p) * SWEEP_THREADS + 2 * tid, od + p * 128);
            if constexpr (GR::NPFG > 0) {
                const double* pf = pf_src(l0);
#pragma unroll
                for (int k = 0; k < GR::NPFG; ++k) cp_async_val(dst + (RG::S_PF + k) * SWEEP_THREADS + tid, pf + k * 128);
            }
#endif
            cp_async_val(dst + RG::S_TLAY * SWEEP_THREADS + tid, in.tlay + a);
            cp_async_val(dst + RG::S_TLEV * SWEEP_THREADS + tid, in.tlev + (down ? a : a + in.ld));
            if constexpr (CLD != CLD_NONE) {
                auto di = [&](int k) { return reinterpret_cast<int*>(dst + (RG::S_INT + k / RG::IPS) * SWEEP_THREADS + tid) + (k % RG::IPS); };
                const unsigned* mw = s.mask + ((long)l0 * 5) * P + i;
                cp_async4(di(0), mw + 4 * P);
                {   // sub-column cloud bits: McICA mask (CLD_MASK) or cldfmc == 1 flags of the caller's arrays (CLD_ARRAYS)
                    constexpr int wlo = G0 >> 5, sh = G0 & 31;
                    cp_async4(di(1), mw + (long)wlo * P);
                    if constexpr (sh + NGG > 32) cp_async4(di(2), mw + (long)(wlo + 1) * P);
                }
            }
        }
        cp_commit();
#if LW_SWEEP_L2PF > 0
        {   // the state words of the layer LW_SWEEP_L2PF further on: one L2 prefetch per 128-byte line of the CTA's block
            const int lev2 = down ? lev - LW_SWEEP_L2PF : lev + LW_SWEEP_L2PF;
            if (lev2 >= 1 && lev2 <= nlay) {
                const char* base = reinterpret_cast<const char*>(s.od + ((blk_lay0 + (lev2 - 1)) * NGPT + G0) * 128);
                for (int ln = tid; ln < NGG * 8; ln += SWEEP_THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ln * 128));
            }
        }
#endif
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

    acc_t radld[NGG];
    HybArr<NGG, RG::clr_smem(CLD)> radclrd;
    radclrd.bind(reinterpret_cast<char*>(ring) + RG::RING_BYTES + tid * sizeof(acc_t));
#pragma unroll
    for (int g = 0; g < NGG; ++g) { radld[g] = 0.; radclrd[g] = 0.; }

    // ---------------------------------------------------------------- downward sweep, rtrnmc.f90:350-455
#pragma unroll
    for (int k = 0; k < RG::NST; ++k) issue(k, nlay - k, true);
    int stage = 0;
    for (int lev = nlay; lev >= 1; --lev) {
        const int l0 = lev - 1;
        const long idx = (long)l0 * P + i;
        const long a = ci + in.ld * l0;
        cp_wait<RG::NST - 1>();
#if LW_SWEEP_TMA
        mbar_wait(&mbar[stage], (mpar >> stage) & 1u);
        mpar ^= 1u << stage;
#endif
        acc_t acc_d = 0., acc_dc = 0.;
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(st_d(stage, RG::S_TLAY), it_lay, fr_lay);
        planck_pos(st_d(stage, RG::S_TLEV), it_lev, fr_lev);
        const bool lower = lev <= laytrop;
        bool cldlyr = false;
        unsigned gbits = 0;
        if constexpr (CLD != CLD_NONE) cldlyr = (((unsigned)st_i(stage, 0)) >> 31) != 0;
        if constexpr (CLD != CLD_NONE) { if (cldlyr) gbits = group_bits(stage); }

        for_bands([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int B = B0 + K;
            constexpr int NG = NGC[B - 1];
            constexpr int GO = GR::goff(K);
            const double blay = planck_at<B>(tb.totplnk, it_lay, fr_lay, own16);
            const double dplankdn = planck_at<B>(tb.totplnk, it_lev, fr_lev, own16, lev == 1) - blay;  // planklev(lev-1) - planklay(lev)
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = st_d(stage, RG::S_PF + pf_slot(B) - GR::PF0);
            PlFrac<B> pf;
            pf.setup(tb, lower, sm);
            // cloud state of this (column, layer) for this band: sub-column bits of the McICA mask; the band's cloud
            // optical depth along the diffusivity angle and its absorptivity come from the cloud-optics kernel
            unsigned bits = 0;
            double odcld_b = 0., abscld_b = 0.;
            if constexpr (CLD != CLD_NONE) {
                if (cldlyr) {
                    bits = (gbits >> GO) & ((1u << NG) - 1u);
                    if constexpr (CLD == CLD_MASK) {
                        if (bits) { const double* q = cldb(l0, B); odcld_b = __ldg(q); abscld_b = __ldg(q + 128); }
                    }
                }
            }
            acc_t dsum = 0., dcsum = 0.;
            // does any column of the warp hold cloud in this band's sub-columns of this layer?  (warp-uniform: the
            // loop with the cloud path is straight-line code for all lanes, the one without it never sees a cloud)
            bool wcld = false;
            if constexpr (CLD != CLD_NONE) wcld = __any_sync(0xffffffffu, bits != 0);
            if (CLD != CLD_NONE && wcld) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double st = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbd;
                    bool big;
                    int itr;
                    gas_point(tb.exptfn, st, pf.get(g), blay, dplankdn, atrans, bbd, big, itr);
                    bool on;
                    double odc, ef;
                    if constexpr (CLD == CLD_MASK) {
                        on = (bits >> g) & 1u;
                        odc = on ? odcld_b : 0.; ef = on ? abscld_b : 0.;
                    } else {
                        // strict rrtmg_lw(): per-g cloud optical depth and absorptivity of the sub-columns with cldfmc == 1
                        // (the only ones that are cloud to rtrnmc, rtrnmc.f90:316-329), from the cloud-arrays kernel
                        on = (bits >> g) & 1u;
                        ef = ldg1_if(S_EF + (long)(G0 + GO + g) * plane + idx, on);
                        odc = ldg1_if(S_OC + (long)(G0 + GO + g) * plane + idx, on);
                    }
                    double atot, bbdtot;
                    int ittot;
                    cloud_point(tb, st, big, itr, odc, on, pf.get(g), blay, dplankdn, atot, bbdtot, ittot);
                    const double gassrc = bbd * atrans;
                    const acc_t xc = relaxed_cld(radld[GO + g], atrans + ef * (1. - atrans), gassrc, bbdtot * atot - gassrc);
                    const acc_t xr = relaxed(radld[GO + g], bbd, atrans);
                    radld[GO + g] = on ? xc : xr;
                    dsum = dsum + radld[GO + g];
                    // clear-sky stream: identical to the total stream until the first cloudy layer from the top
                    const acc_t sep = relaxed(radclrd[GO + g], bbd, atrans);
                    radclrd[GO + g] = (lev <= ctop) ? sep : radld[GO + g];
                    dcsum = dcsum + radclrd[GO + g];
                }
            } else {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double st = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbd;
                    bool big;
                    int itr;
                    gas_point(tb.exptfn, st, pf.get(g), blay, dplankdn, atrans, bbd, big, itr);
                    radld[GO + g] = relaxed(radld[GO + g], bbd, atrans);
                    dsum = dsum + radld[GO + g];
                    if (CLD != CLD_NONE) {
                        const acc_t sep = relaxed(radclrd[GO + g], bbd, atrans);
                        radclrd[GO + g] = (lev <= ctop) ? sep : radld[GO + g];
                        dcsum = dcsum + radclrd[GO + g];
                    }
                }
            }
            // stage dump of the captured column (kept out of the g loops: a branch per g-point would cut the straight-line
            // code the scheduler interleaves): Planck fractions and the table indices itgas / ittot the reference forms
            // (ittot for every sub-column of a layer with icldlyr = 1, rtrnmc.f90:386, 404, 412)
            if (dbg) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double st = ((GO + g) & 1) ? odp.y : odp.x;
                    const bool big = od_state_is_index(st);
                    const int itr = od_state_index(st);
                    const long o = (long)(G0 + GO + g) * nlay + l0;
                    s.dbg_fracs[o] = pf.get(g);
                    s.dbg_itr[2 * o] = big ? itr : -1;
                    int d1 = -1;
                    if (CLD != CLD_NONE && cldlyr) {
                        double odc = 0.;
                        if constexpr (CLD == CLD_MASK) odc = ((bits >> g) & 1u) ? odcld_b : 0.;
                        if constexpr (CLD == CLD_ARRAYS) odc = S_OC[(long)(G0 + GO + g) * plane + idx];
                        double t0, t1;
                        cloud_point(tb, st, big, itr, odc, true, pf.get(g), blay, dplankdn, t0, t1, d1);
                    }
                    s.dbg_itr[2 * o + 1] = d1;
                }
            }
            // band close-out for level lev-1: flux += (rad * wtdiff) * delwave, rtrnmc.f90:538-550
            acc_d = acc_d + (dsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
            if (CLD != CLD_NONE) acc_dc = acc_dc + ((lev <= ctop ? dcsum : dsum) * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
        });
        if (live) {
            acc_t* o = flx(l0);
            o[0] = acc_d;
            if (CLD != CLD_NONE && ctop > 0) o[128] = acc_dc;
        }
#if LW_SWEEP_TMA
        __syncthreads();  // every thread is done with the stage the bulk copy is about to overwrite
#endif
        issue(stage, lev - RG::NST, true);
        stage = stage + 1 == RG::NST ? 0 : stage + 1;
    }
    cp_wait<0>();

    // ---------------------------------------------------------------- surface, rtrnmc.f90:465-484
    // radld / radclrd are dead after the sweep; they become the upward radiances
    acc_t (&radlu)[NGG] = radld;
    auto& radclru = radclrd;
    acc_t d_radlu[IDRV ? NGG : 1], d_radclru[IDRV ? NGG : 1];
#if LW_SWEEP_TMA
    __syncthreads();
#endif
#pragma unroll
    for (int k = 0; k < RG::NST; ++k) issue(k, 1 + k, false);
    {
        acc_t acc_u = 0., acc_uc = 0., acc_du = 0., acc_duc = 0.;
        // downward flux at the top of the atmosphere is never touched by the sweep: drad(nlayers) = 0
        if (live) {
            flx(nlay)[0] = 0.0;
            if (CLD != CLD_NONE && ctop > 0) flx(nlay)[128] = 0.0;
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
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = pf_src(0)[(pf_slot(B) - GR::PF0) * 128];
            PlFrac<B> pf;
            pf.setup(tb, lower1, sm);  // fracs(1, g)
            acc_t usum = 0., ucsum = 0., dusum = 0.;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double rad0 = pf.get(g) * plankbnd;
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
                    d_radlu[GO + g] = pf.get(g) * dplankbnd_dt;
                    d_radclru[GO + g] = d_radlu[GO + g];
                    dusum = dusum + d_radlu[GO + g];
                }
            }
            acc_u = acc_u + (usum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
            acc_uc = acc_uc + (ucsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
            if constexpr (IDRV == 1) {
                const acc_t addd = (dusum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
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
#if LW_SWEEP_TMA
        mbar_wait(&mbar[stage], (mpar >> stage) & 1u);
        mpar ^= 1u << stage;
#endif
        acc_t acc_u = 0., acc_uc = 0., acc_du = 0., acc_duc = 0.;
        int it_lay, it_lev;
        double fr_lay, fr_lev;
        planck_pos(st_d(stage, RG::S_TLAY), it_lay, fr_lay);
        planck_pos(st_d(stage, RG::S_TLEV), it_lev, fr_lev);
        const bool lower = lev <= laytrop;
        bool cldlyr = false;
        unsigned gbits = 0;
        if constexpr (CLD != CLD_NONE) cldlyr = (((unsigned)st_i(stage, 0)) >> 31) != 0;
        if constexpr (CLD != CLD_NONE) { if (cldlyr) gbits = group_bits(stage); }

        for_bands([&](auto kc) {
            constexpr int K = decltype(kc)::value;
            constexpr int B = B0 + K;
            constexpr int NG = NGC[B - 1];
            constexpr int GO = GR::goff(K);
            const double blay = planck_at<B>(tb.totplnk, it_lay, fr_lay, own16);
            const double dplankup = planck_at<B>(tb.totplnk, it_lev, fr_lev, own16) - blay;  // planklev(lev) - planklay(lev)
            double sm = 0.;
            if constexpr (pf_slot(B) >= 0) sm = st_d(stage, RG::S_PF + pf_slot(B) - GR::PF0);
            PlFrac<B> pf;
            pf.setup(tb, lower, sm);
            unsigned bits = 0;
            double odcld_b = 0., abscld_b = 0.;
            if constexpr (CLD != CLD_NONE) {
                if (cldlyr) {
                    bits = (gbits >> GO) & ((1u << NG) - 1u);
                    if constexpr (CLD == CLD_MASK) {
                        if (bits) { const double* q = cldb(l0, B); odcld_b = __ldg(q); abscld_b = __ldg(q + 128); }
                    }
                }
            }
            acc_t usum = 0., ucsum = 0., dusum = 0., ducsum = 0.;
            bool wcld = false;
            if constexpr (CLD != CLD_NONE) wcld = __any_sync(0xffffffffu, bits != 0);
            if (CLD != CLD_NONE && wcld) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2 odp = st_od(stage, (GO + g) >> 1);
                    const double st = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbugas;
                    bool big;
                    int itr;
                    gas_point(tb.exptfn, st, pf.get(g), blay, dplankup, atrans, bbugas, big, itr);
                    bool on;
                    double odc, ef;
                    if constexpr (CLD == CLD_MASK) {
                        on = (bits >> g) & 1u;
                        odc = on ? odcld_b : 0.; ef = on ? abscld_b : 0.;
                    } else {
                        on = (bits >> g) & 1u;
                        ef = ldg1_if(S_EF + (long)(G0 + GO + g) * plane + idx, on);
                        odc = ldg1_if(S_OC + (long)(G0 + GO + g) * plane + idx, on);
                    }
                    double atot, bbutot;
                    int ittot;
                    cloud_point(tb, st, big, itr, odc, on, pf.get(g), blay, dplankup, atot, bbutot, ittot);
                    const double gassrc = bbugas * atrans;
                    const acc_t xc = relaxed_cld(radlu[GO + g], atrans + ef * (1. - atrans), gassrc, bbutot * atot - gassrc);
                    const acc_t xr = relaxed(radlu[GO + g], bbugas, atrans);
                    radlu[GO + g] = on ? xc : xr;
                    usum = usum + radlu[GO + g];
                    if constexpr (IDRV == 1) {
                        // d_radlu * cldfmc * (1 - atot) + d_radlu * (1 - cldfmc) * (1 - atrans) with cldfmc = 0 or 1
                        d_radlu[GO + g] = d_radlu[GO + g] * (1.0 - (on ? atot : atrans));
                        dusum = dusum + d_radlu[GO + g];
                    }
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
                    const double st = ((GO + g) & 1) ? odp.y : odp.x;
                    double atrans, bbugas;
                    bool big;
                    int itr;
                    gas_point(tb.exptfn, st, pf.get(g), blay, dplankup, atrans, bbugas, big, itr);
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
            acc_u = acc_u + (usum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
            if (CLD != CLD_NONE) acc_uc = acc_uc + (ucsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1];
            if constexpr (IDRV == 1) {
                acc_du = acc_du + (dusum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
                if (CLD != CLD_NONE) acc_duc = acc_duc + (ducsum * (acc_t)0.5) * (acc_t)tb.delwave[B - 1] * (acc_t)tb.fluxfac;
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
#if LW_SWEEP_TMA
        __syncthreads();
#endif
        issue(stage, lev + RG::NST, false);
        stage = stage + 1 == RG::NST ? 0 : stage + 1;
    }
    cp_wait<0>();
}
// ===== LW_PRECISION_END =====

}  // namespace lw
