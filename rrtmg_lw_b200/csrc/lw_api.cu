// lw_api.cu -- the C ABI of include/rrtmg_lw_b200.h: table upload, scratch management, chunked
// execution of the kernel pipeline, host<->device staging for host-pointer calls.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rrtmg_lw_b200.h"
#include "../../include/rrtmg_lw_kdata_layout.h"
#include "lw_band.cuh"
#include "lw_launch.h"

// numeric blob of tools/extract_static_tables.py, embedded at build time
static const double kStaticTables[] = {
#include "_generated/lw_static_tables.inc"
};

namespace lw {
namespace {

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
struct Ctx {
    bool inited = false;
    HostTables H;
    double* d_arena = nullptr;
    double* d_arena_p = nullptr;
    // FP32 mode: float copies of every table the band kernels read
    std::vector<float*> f32_allocs;
    int precision = 64;
    GTab* d_gtab = nullptr;
    double *d_totplnk = nullptr, *d_totplnkderiv = nullptr, *d_exptfn = nullptr, *d_tau = nullptr;
    double *d_absice2 = nullptr, *d_absice3 = nullptr, *d_absliq1 = nullptr;
    // scratch, sized for (cap_cols, cap_lay)
    Scratch S{};
    SweepTab sweep_tab{};     // tables of the sweep kernels (kernel parameter)
    // two-phase path: the McICA generator / cloud optics run beside the taumol kernel, and the sweep kernels (which own
    // their flux slots) beside each other, on side streams forked from and joined to the caller's stream
    static constexpr int NSIDE = 3;
    cudaStream_t side[NSIDE] = {};
    int mcica_beside = env_int("RRTMG_LW_MCICA_BESIDE", 0);  // set_option("mcica_beside", n): KISS masks + cloud optics as a work-queue kernel resident
                              // beside taumol (n CTAs per SM, launched first) + a full-grid launch after taumol for the
                              // items left; needs concurrency, KISS, the fused McICA entry
    cudaEvent_t e_kw = nullptr, e_mop = nullptr;
    cudaStream_t hi = nullptr;  // highest-priority stream for the resident work-queue kernel: its few CTAs are placed
                                // ahead of taumol's whenever registers come free
    cudaEvent_t e_fork = nullptr, e_colprep = nullptr, e_t = nullptr, e_cld = nullptr, e_side[NSIDE] = {};
    int tail_group = 1;       // set_option("tail_group", 0): bands 13-16 as four sweep kernels instead of one
    int concurrency = 1;      // set_option("concurrency", 0) keeps every kernel on the caller's stream; 1 = auto (see run_device)
    long out_ld = 0;          // set_option("out_ld", n): column extent of the OUTPUT arrays of device-pointer calls (0 = ld)
    long cap_cols = 0;
    int cap_lay = 0;
    unsigned have = 0;        // ScratchPiece bits allocated at the current capacity
    std::vector<void*> scratch_allocs;
    // host-mode staging arena
    double* stage = nullptr;
    size_t stage_doubles = 0;
    int chunk = 262144;       // columns per internal pass of a device-pointer call (scratch is sized to min(ncol, chunk))
    int host_chunk = 32768;   // columns per staged pass of a host-pointer call (H2D / compute / D2H overlap)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_in[2] = {nullptr, nullptr}, e_done[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};
    int band_kernel = 4;  // 4 = two-phase path (lw_taumol.cuh + lw_rtrnmc.cuh), 2 = fused per-band kernels (lw_band.cuh)
    int sweep_groups = 0; // two-phase path: 0 = one sweep kernel per band (band sums added in the reference's order),
                          // 1 = sweep kernels over band groups {1,2} {3} {4} {5} {6,7} {8,9} {10-12} {13-16}
    int band_lo = 1, band_hi = NBND;  // band range of a run (two-phase path; the column-mode harness' istart / iend)
    int nstage = 0;       // timed stages of the last timed pass (ms_bands[0 .. nstage))
    int sms = 0;
    long launches = 0;
    bool dbg = false;
    long dbg_col = -1;  // requested by debug_capture: always column given at debug_stages time? (we capture col 0..)
    int dbg_nlay = 0;
    bool timing = false;
    // Mersenne-Twister stream bookkeeping of a staged host call: columns of the whole call / first column of the pass
    long mt_total = 0, mt_base = 0;
    // set_option("mt_global_ncol" / "mt_col_offset"): the call's columns are columns [offset, offset + ncol) of a
    // larger batch of mt_global_ncol columns whose single Mersenne-Twister stream is to be reproduced (sharded runs)
    long mt_opt_total = 0, mt_opt_base = 0;
    // (column, layer) of the input error a host-pointer call returned (1-based column of the CALL, not of the pass)
    long last_err_col = 0, err_col0 = 0;
    int last_err_lay = 0;
    float ms_band = 0.f, ms_all = 0.f;
    float ms_bands[NBND + 2] = {};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evb[NBND + 3] = {};
};
// One context per CUDA device: tables, scratch, streams and options belong to the device they were created on.
// rrtmg_lw_b200_set_device(d) binds the calling process's following calls to device d (one host process can drive
// several GPUs by switching between them); without it the library follows the caller's current device.
constexpr int LW_MAX_DEVICES = 16;
Ctx g_ctx[LW_MAX_DEVICES];
int g_cur = 0;        // context of the call in progress
int g_explicit = -1;  // device chosen with set_device, -1 = follow cudaGetDevice
std::mutex g_mu;
#define G (g_ctx[g_cur])
void select_ctx() {
    int d = 0;
    if (g_explicit >= 0) {
        d = g_explicit;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != d) cudaSetDevice(d);
    } else if (cudaGetDevice(&d) != cudaSuccess) {
        d = 0;
    }
    g_cur = (d >= 0 && d < LW_MAX_DEVICES) ? d : 0;
}
#define LW_ENTER() std::lock_guard<std::mutex> lk(g_mu); select_ctx()

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e__ = (x);                                                           \
        if (e__ != cudaSuccess) {                                                        \
            std::fprintf(stderr, "rrtmg_lw_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return RRTMG_LW_ERR_CUDA;                                                    \
        }                                                                                \
    } while (0)

template <typename T>
int dalloc(T** p, size_t n) {
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}

void free_scratch() {
    for (void* p : G.scratch_allocs) cudaFree(p);
    G.scratch_allocs.clear();
    G.cap_cols = 0;
    G.cap_lay = 0;
    G.have = 0;
}

// what a pass needs from the scratch (allocated lazily, piece by piece: a path that is not used costs no memory)
enum ScratchPiece {
    SP_FUSED = 1,    // fused band kernels (band_kernel = 2): setcoef fields, packed indices, 16-byte sweep state
    SP_TWOPHASE = 2, // two-phase path (band_kernel = 4): 8-byte optical depths, Planck-fraction etas, secdiff
    SP_CLDB = 4,     // two-phase path, McICA mask entry: per-band cloud optical depth of cloudy layers
    SP_GCLD = 8,     // two-phase path, strict rrtmg_lw() entry with clouds: per-g efclfrac / odcld
    SP_FP32IN = 16,  // FP32 mode: float copies of the caller arrays
    SP_GRID = 32,    // non-McICA entry: cloud fraction, overlap factors, band cloud optical depths per layer
};

int ensure_scratch(long cols, int nlay, unsigned need) {
    if (cols > G.cap_cols || nlay > G.cap_lay) {
        free_scratch();
        const long P = ((cols + 127) / 128) * 128;  // whole 128-column blocks
        Scratch& S = G.S;
        S = Scratch{};
        S.pitch = P;
        G.cap_cols = P;
        G.cap_lay = nlay;
    }
    Scratch& S = G.S;
    const long P = S.pitch;
    nlay = G.cap_lay;
    auto A = [&](auto** p, size_t n) -> int {
        int rc = dalloc(p, n);
        if (rc == 0) G.scratch_allocs.push_back((void*)*p);
        return rc;
    };
    const size_t plane = (size_t)nlay * P;
    if (!S.status) {  // pieces every path uses
        if (A(&S.pwvcm, (size_t)P)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.laytrop, (size_t)P)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.mask, 5 * plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.ctop, (size_t)P)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.status, (size_t)1)) return RRTMG_LW_ERR_CUDA;
        CK(cudaMemset(S.status, 0, sizeof(unsigned long long)));
        if (A(&S.dbg_taug, (size_t)NGPT * nlay)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.dbg_fracs, (size_t)NGPT * nlay)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.dbg_itr, (size_t)2 * NGPT * nlay)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.dbg_ints, (size_t)6 * nlay)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.dbg_facs, (size_t)4 * nlay)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.dbg_scal, (size_t)2)) return RRTMG_LW_ERR_CUDA;
        S.dbg_col = -1;
    }
    const unsigned miss = need & ~G.have;
    if (miss & SP_FUSED) {
        if (A(&S.fld, (size_t)NFIELD * plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.ipack, plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.state, (size_t)5 * 16 * plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.abscld, plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.fluxB, (size_t)6 * (nlay + 1) * P)) return RRTMG_LW_ERR_CUDA;
        // FP32 mode: the float kernels reuse the same memory through float views
        S.fldf = reinterpret_cast<float*>(S.fld);
        S.statef = reinterpret_cast<float*>(S.state);
        S.abscldf = reinterpret_cast<float*>(S.abscld);
    }
    if (miss & SP_TWOPHASE) {
        if (A(&S.od, (size_t)NGPT * plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.pfs, (size_t)NPF * plane)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.secdiff, (size_t)NBND * P)) return RRTMG_LW_ERR_CUDA;
        if (A(&S.fluxP, (size_t)NBND * 6 * (nlay + 1) * P)) return RRTMG_LW_ERR_CUDA;
    }
    if (miss & SP_CLDB) {
        if (A(&S.cldb, (size_t)2 * NBND * plane)) return RRTMG_LW_ERR_CUDA;
    }
    if (miss & SP_GCLD) {
        if (A(&S.gcld, (size_t)2 * NGPT * plane)) return RRTMG_LW_ERR_CUDA;
    }
    if (miss & SP_GRID) {
        if (A(&S.gridcld, (size_t)GC_NREC * plane)) return RRTMG_LW_ERR_CUDA;
    }
    if (miss & SP_FP32IN) {
        if (A(&S.inf, (size_t)(40 * (size_t)(nlay + 1) + 32) * P)) return RRTMG_LW_ERR_CUDA;
    }
    G.have |= need;
    return 0;
}

int ensure_stage(size_t doubles) {
    if (doubles <= G.stage_doubles) return 0;
    if (G.stage) cudaFree(G.stage);
    G.stage = nullptr;
    G.stage_doubles = 0;
    CK(cudaMalloc((void**)&G.stage, doubles * sizeof(double)));
    G.stage_doubles = doubles;
    return 0;
}

int decode_status(int* col, int* lay) {
    unsigned long long v = 0;
    if (!G.S.status) return RRTMG_LW_OK;
    CK(cudaMemcpy(&v, G.S.status, sizeof(v), cudaMemcpyDeviceToHost));
    if (v == 0) return RRTMG_LW_OK;
    CK(cudaMemset(G.S.status, 0, sizeof(v)));
    G.last_err_col = (long)((v >> 24) & 0xffffffffull);
    G.last_err_lay = (int)(v & 0xffffff);
    if (col) *col = (int)G.last_err_col;
    if (lay) *lay = G.last_err_lay;
    return (int)(v >> 56);
}

bool launch_band(int band, int cld, int idrv, const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                 cudaStream_t st) {
    return launch_band_tu0(band, cld, idrv, in, out, s, r, st) || launch_band_tu1(band, cld, idrv, in, out, s, r, st) ||
           launch_band_tu2(band, cld, idrv, in, out, s, r, st) || launch_band_tu3(band, cld, idrv, in, out, s, r, st) ||
           launch_band_tu4(band, cld, idrv, in, out, s, r, st) || launch_band_tu5(band, cld, idrv, in, out, s, r, st);
}

bool launch_band32(int band, int cld, int idrv, const ColInF& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                   cudaStream_t st) {
    return launch_band32_tu0(band, cld, idrv, in, out, s, r, st) || launch_band32_tu1(band, cld, idrv, in, out, s, r, st) ||
           launch_band32_tu2(band, cld, idrv, in, out, s, r, st) || launch_band32_tu3(band, cld, idrv, in, out, s, r, st) ||
           launch_band32_tu4(band, cld, idrv, in, out, s, r, st) || launch_band32_tu5(band, cld, idrv, in, out, s, r, st);
}

// sweep kernels of bands [b0, b0 + nb) of the two-phase path
bool launch_sweep(int b0, int nb, int cld, int idrv, const ColIn& in, const ColOut& out, const Scratch& s, const RunCfg& r,
                  cudaStream_t st) {
    const SweepTab& tb = G.sweep_tab;
    return launch_sweep_stu1(b0, nb, cld, idrv, in, out, s, r, tb, st) || launch_sweep_stu2(b0, nb, cld, idrv, in, out, s, r, tb, st) ||
           launch_sweep_stu3(b0, nb, cld, idrv, in, out, s, r, tb, st) || launch_sweep_stu4(b0, nb, cld, idrv, in, out, s, r, tb, st) ||
           launch_sweep_stu5(b0, nb, cld, idrv, in, out, s, r, tb, st);
}
// default band groups of the sweep kernels: {first band, bands}
const int kSweepGroups[][2] = {{1, 2}, {3, 1}, {4, 1}, {5, 1}, {6, 2}, {8, 2}, {10, 3}, {13, 4}};

enum Mode { MODE_STRICT = 0, MODE_FUSED = 1, MODE_GRID = 2 };  // MODE_GRID: the non-McICA entry (lw_gridcloud.cuh)

int ensure_side_streams() {
    if (G.side[0]) return 0;
    for (auto& q : G.side) CK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&G.e_fork, &G.e_colprep, &G.e_t, &G.e_cld}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (auto& e : G.e_side) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&G.e_kw, &G.e_mop}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    int prio_least = 0, prio_greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CK(cudaStreamCreateWithPriority(&G.hi, cudaStreamNonBlocking, prio_greatest));
    return 0;
}

// Device-resident pipeline over column chunks.  `in`/`out` hold device pointers with their ld.
int run_device(Mode mode, int ncol, int nlay, int icld, int idrv, int inflag, int iceflag, int liqflag, int permuteseed,
               int irng, ColIn in, ColOut out, cudaStream_t st) {
    if (nlay < 4 && mode == MODE_FUSED && icld > 0 && irng == 0) return RRTMG_LW_ERR_ARG;  // KISS seeds need 4 layers
    if (nlay < 1 || nlay > LW_MAXLAY) return RRTMG_LW_ERR_ARG;
    const int chunk = G.chunk;
    const int cld = icld == 0 ? CLD_NONE : (mode == MODE_FUSED ? CLD_MASK : (mode == MODE_GRID ? CLD_GRID : CLD_ARRAYS));
    // optional FP32 mode: float band kernels for clear sky and the fused McICA entry (fused band kernels only; the
    // strict entry with per-g cloud arrays stays FP64)
    const bool fp32 = G.precision == 32 && cld != CLD_ARRAYS && mode != MODE_GRID;
    // the FP32 mode lives in the fused band kernels; the non-McICA entry exists on the two-phase path only
    const bool two_phase = (G.band_kernel == 4 && !fp32) || mode == MODE_GRID;
    unsigned need = two_phase ? SP_TWOPHASE : SP_FUSED;
    if (two_phase && cld == CLD_MASK) need |= SP_CLDB;
    if (two_phase && cld == CLD_ARRAYS) need |= SP_GCLD;
    if (cld == CLD_GRID) need |= SP_GRID;
    if (fp32) need |= SP_FP32IN;
    int rc = ensure_scratch(ncol < chunk ? ncol : chunk, nlay, need);
    if (rc) return rc;
    const int b_lo = G.band_lo, b_hi = G.band_hi;  // band range of the run (1..16 unless the band_lo / band_hi options say otherwise)
    if (b_lo > b_hi || (!two_phase && (b_lo != 1 || b_hi != NBND))) return RRTMG_LW_ERR_ARG;  // band ranges: two-phase path only
    if (G.timing) {
        for (auto& e : G.ev)
            if (!e) CK(cudaEventCreate(&e));
        for (auto& e : G.evb)
            if (!e) CK(cudaEventCreate(&e));
    }
    for (long c0 = 0; c0 < ncol; c0 += chunk) {
        RunCfg r{};
        r.ncol = (int)((ncol - c0) < chunk ? (ncol - c0) : chunk);
        r.c0 = c0;
        r.nlay = nlay;
        r.idrv = idrv;
        r.inflag = inflag; r.iceflag = iceflag; r.liqflag = liqflag;
        r.icld = icld; r.irng = irng; r.permuteseed = permuteseed;
        r.fp32 = fp32 ? 1 : 0;
        r.fld_block = 1;
        r.istart = b_lo;
        Scratch S = G.S;
        S.err_base = G.err_col0;
        S.dbg_col = (G.dbg && G.dbg_col >= c0 && G.dbg_col < c0 + r.ncol) ? (int)(G.dbg_col - c0) : -1;
        if (S.dbg_col >= 0) G.dbg_nlay = nlay;
        if (G.timing) CK(cudaEventRecord(G.ev[0], st));
        int nstage = 0;  // timed stages between ev[1] and ev[2]
        // McICA mask of the chunk (KISS per-column streams, or the single Mersenne-Twister stream of the call)
        auto gen_masks = [&](cudaStream_t q) -> int {
            if (irng == 0) {
                launch_mcica_kiss(in, S, r, q); G.launches++;
                return RRTMG_LW_OK;
            }
            // one global stream over (sub-column, column, layer): position of this chunk inside the whole call
            const long mt_tot = G.mt_total ? G.mt_total : (G.mt_opt_total ? G.mt_opt_total : (long)ncol);
            const long mt_b0 = G.mt_total ? G.mt_base : G.mt_opt_base;
            if (mt_b0 + ncol > mt_tot) return RRTMG_LW_ERR_ARG;
            const int nl = launch_mcica_mt(in, S, r, mt_tot, mt_b0 + c0, q);
            if (nl < 0) return RRTMG_LW_ERR_CUDA;
            G.launches += nl;
            return RRTMG_LW_OK;
        };
        if (two_phase) {
            // Kernel graph of a pass (arrows = stream order or events):
            //   colprep -> taumol ----------------------+-> sweep kernels (one flux slot each, any order) -> finalize
            //   McICA masks -> (colprep) cloud optics --+
            // With concurrency on (and no per-stage timing) the mask generator and the cloud optics run on a side stream
            // beside taumol, and the sweep kernels are dealt over the caller's stream and the side streams.
            const bool conc = G.concurrency != 0 && !G.timing && S.dbg_col < 0;
            if (conc && (rc = ensure_side_streams()) != 0) return rc;
            // The sweep kernels are dealt over four streams only for passes of up to 393 216 columns: there the overlap of
            // their tails pays (131 072 columns: 190.4 -> 177.9 ms per 1 M columns), while four sweep kernels interleaved
            // on every SM cost 4 ms per 1 M columns at 524 288 (173.5 -> 177.5; profiles/r02_stream_schedule_experiment.log).
            // concurrency 2 / 3 force one half of the schedule (masks beside taumol only / sweeps over streams only).
            const bool conc_masks = conc && G.concurrency != 3;
            const bool conc_sweeps = conc && (G.concurrency >= 3 || (G.concurrency == 1 && r.ncol <= 393216));
            const int sweep_streams = G.concurrency == 4 ? 2 : (G.concurrency == 5 ? 3 : Ctx::NSIDE + 1);  // 4 / 5: dev knobs
            cudaStream_t qm = conc_masks ? G.side[0] : st;  // masks + cloud optics
            if (conc) {
                CK(cudaEventRecord(G.e_fork, st));
                for (auto& q : G.side) CK(cudaStreamWaitEvent(q, G.e_fork, 0));
            }
            launch_colprep(in, S, r, st); G.launches++;
            if (conc) CK(cudaEventRecord(G.e_colprep, st));
            // masks + cloud optics beside taumol: the work-queue kernel goes first (its one CTA per SM is resident when
            // taumol's CTAs arrive and take the rest of the registers), a full-grid launch after taumol takes what is left
            bool beside = conc && cld == CLD_MASK && irng == 0 && G.mcica_beside != 0;
            if (beside) {
                static const bool noprio = env_int("RRTMG_LW_BESIDE_NOPRIO", 0) != 0;
                cudaStream_t qk = noprio ? qm : G.hi;
                CK(cudaStreamWaitEvent(qk, G.e_colprep, 0));  // the fused cloud optics read colprep's secdiff
                if (launch_mcica_kiss_work(in, S, r, G.mcica_beside, qk)) { G.launches++; CK(cudaEventRecord(G.e_kw, qk)); }
                else beside = false;
            }
            if (cld == CLD_MASK && !beside) {
                if ((rc = gen_masks(qm)) != RRTMG_LW_OK) return rc;
                if (conc) CK(cudaStreamWaitEvent(qm, G.e_colprep, 0));
                launch_cldoptics(in, S, r, qm); G.launches++;
                if (conc) CK(cudaEventRecord(G.e_cld, qm));
            }
            if (cld == CLD_ARRAYS) { launch_cldarrays(in, S, r, st); G.launches += 2; }  // after colprep: needs secdiff
            if (cld == CLD_GRID) { launch_gridcloud_prep(in, S, r, icld != 1, st); G.launches++; }  // needs colprep's secdiff
            if (G.timing) CK(cudaEventRecord(G.ev[1], st));
            if (G.timing) CK(cudaEventRecord(G.evb[nstage++], st));
            G.launches += launch_taumol(in, S, r, st);
            if (conc) {
                CK(cudaEventRecord(G.e_t, st));
                for (auto& q : G.side) CK(cudaStreamWaitEvent(q, G.e_t, 0));
                if (beside) {
                    // side[1] is past taumol now: the items the resident CTAs have not taken, at full occupancy
                    launch_mcica_kiss_work(in, S, r, 0, G.side[1]); G.launches++;
                    CK(cudaEventRecord(G.e_mop, G.side[1]));
                    CK(cudaStreamWaitEvent(st, G.e_kw, 0));
                    CK(cudaStreamWaitEvent(st, G.e_mop, 0));
                    for (auto& q : G.side) {
                        CK(cudaStreamWaitEvent(q, G.e_kw, 0));
                        CK(cudaStreamWaitEvent(q, G.e_mop, 0));
                    }
                } else if (cld == CLD_MASK) {
                    CK(cudaStreamWaitEvent(st, G.e_cld, 0));
                    for (int k = 1; k < Ctx::NSIDE; ++k) CK(cudaStreamWaitEvent(G.side[k], G.e_cld, 0));
                }
            }
            unsigned slots = 0;
            int nsweep = 0;
            auto sweep = [&](int b0, int nb) -> bool {
                cudaStream_t q = st;
                if (conc_sweeps) { const int k = nsweep % sweep_streams; q = k == 0 ? st : G.side[k - 1]; }
                ++nsweep;
                if (G.timing) CK(cudaEventRecord(G.evb[nstage++], st));
                slots |= 1u << (b0 - 1);
                G.launches++;
                if (cld == CLD_GRID) return launch_gridcloud_sweep(b0, icld != 1, idrv, in, S, r, G.sweep_tab, q);
                return launch_sweep(b0, nb, cld, idrv, in, out, S, r, q);
            };
            if (G.sweep_groups && cld != CLD_GRID && b_lo == 1 && b_hi == NBND) {
                for (const auto& g : kSweepGroups)
                    if (!sweep(g[0], g[1])) return RRTMG_LW_ERR_ARG;
            } else {
                // longest kernels first when they are dealt over several streams (bands by g-point count).  Bands 13-16
                // (4 + 2 + 2 + 2 g-points) go as one kernel when the run holds all four: a band's sweep has a fixed cost per
                // column (temperatures, mask words, its flux slot) that a 2-g-point band does not amortise.
                const int order[NBND] = {3, 5, 4, 2, 7, 9, 1, 6, 8, 11, 12, 10, 13, 14, 15, 16};
                const bool tail4 = G.tail_group != 0 && cld != CLD_GRID && b_lo <= 13 && b_hi == NBND;
                for (int k = 0; k < NBND; ++k) {
                    const int b = conc_sweeps ? order[k] : k + 1;
                    if (b < b_lo || b > b_hi) continue;
                    if (tail4 && b >= 13) {
                        if (b == 13 && !sweep(13, 4)) return RRTMG_LW_ERR_ARG;
                        continue;
                    }
                    if (!sweep(b, 1)) return RRTMG_LW_ERR_ARG;
                }
            }
            if (G.timing) CK(cudaEventRecord(G.evb[nstage], st));
            if (conc)
                for (int k = 0; k < Ctx::NSIDE; ++k) {
                    CK(cudaEventRecord(G.e_side[k], G.side[k]));
                    CK(cudaStreamWaitEvent(st, G.e_side[k], 0));
                }
            if (G.timing) CK(cudaEventRecord(G.ev[2], st));
            launch_finalize_slots(in, out, S, r, cld != CLD_NONE, slots, st);
            G.launches++;
        } else {
            launch_prep(in, S, r, st); G.launches++;
            if (cld == CLD_MASK && (rc = gen_masks(st)) != RRTMG_LW_OK) return rc;
            if (cld == CLD_ARRAYS) { launch_cldflags(in, S, r, st); G.launches += 2; }
            if (G.timing) CK(cudaEventRecord(G.ev[1], st));
            ColInF inf{};
            if (fp32) { G.launches += launch_cvt_inputs(in, S, r, cld != CLD_NONE, inflag, &inf, st); }
            for (int b = 1; b <= NBND; ++b) {
                if (G.timing) CK(cudaEventRecord(G.evb[nstage++], st));
                if (fp32 ? !launch_band32(b, cld, idrv, inf, out, S, r, st) : !launch_band(b, cld, idrv, in, out, S, r, st))
                    return RRTMG_LW_ERR_ARG;
                G.launches++;
            }
            if (G.timing) CK(cudaEventRecord(G.evb[nstage], st));
            if (G.timing) CK(cudaEventRecord(G.ev[2], st));
            launch_finalize(in, out, S, r, cld != CLD_NONE, st);
            G.launches++;
        }
        if (G.timing) {
            CK(cudaEventRecord(G.ev[3], st));
            CK(cudaEventSynchronize(G.ev[3]));
            float a = 0.f, b = 0.f;
            CK(cudaEventElapsedTime(&a, G.ev[1], G.ev[2]));
            CK(cudaEventElapsedTime(&b, G.ev[0], G.ev[3]));
            G.ms_band += a;
            G.ms_all += b;
            G.nstage = nstage;
            for (int k = 0; k < nstage; ++k) {
                float m = 0.f;
                CK(cudaEventElapsedTime(&m, G.evb[k], G.evb[k + 1]));
                G.ms_bands[k] += m;
            }
        }
        CK(cudaGetLastError());
    }
    return RRTMG_LW_OK;
}

// staging helper: carve nrows x (w * P) doubles from the arena, copy the chunk's columns in
struct Stager {
    double* base;
    size_t used = 0;
    long P;
    cudaStream_t st;
    double* carve(size_t n) { double* p = base + used; used += n; return p; }
    // src element (e, i, k) at e + w*(i + ld*k); copies columns [c0, c0+nc)
    int in(const double* src, double** dst, int w, long ld, long c0, int nc, long nrows) {
        double* d = carve((size_t)w * P * nrows);
        *dst = d;
        CK(cudaMemcpy2DAsync(d, (size_t)w * P * sizeof(double), src + (size_t)w * c0, (size_t)w * ld * sizeof(double),
                             (size_t)w * nc * sizeof(double), nrows, cudaMemcpyHostToDevice, st));
        return 0;
    }
    int out(double* dsthost, const double* d, int w, long ld, long c0, int nc, long nrows) {
        CK(cudaMemcpy2DAsync(dsthost + (size_t)w * c0, (size_t)w * ld * sizeof(double), d, (size_t)w * P * sizeof(double),
                             (size_t)w * nc * sizeof(double), nrows, cudaMemcpyDeviceToHost, st));
        return 0;
    }
};

int ensure_pipeline() {
    if (G.s_in) return 0;
    CK(cudaStreamCreateWithFlags(&G.s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&G.s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CK(cudaEventCreateWithFlags(&G.e_in[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&G.e_done[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&G.e_out[k], cudaEventDisableTiming));
    }
    return 0;
}

struct HostOut {
    double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc, *duflx_dt, *duflxc_dt;
};

// Host-pointer calls: the column range is cut into passes of G.host_chunk columns that alternate between
// two staging arenas; pass k+1 is copied in (stream s_in) while pass k computes (the caller's stream) and
// pass k-1 is copied out (stream s_out), so the PCIe transfers of a large call overlap the kernels.
// stage_in(sg, in, c0, nc) carves the device copies of one pass and enqueues their H2D copies.
template <class StageIn>
int host_pipeline(Mode mode, int ncol, int ld, int nlay, int icld, int idrv, int inflag, int iceflag, int liqflag,
                  int permuteseed, int irng, size_t per_col, StageIn stage_in, const HostOut& ho, cudaStream_t st) {
    int rc = ensure_pipeline();
    if (rc) return rc;
    int hc = G.host_chunk < G.chunk ? G.host_chunk : G.chunk;
    const int chunk = hc < ncol ? hc : ncol;
    const long P = ((chunk + 31) / 32) * 32;
    const size_t half = per_col * P + 64;
    rc = ensure_stage(2 * half);
    if (rc) return rc;
    const int saved_chunk = G.chunk;
    // every exit, error or not, first drains the three streams: the caller may free or reuse its (pinned) arrays
    // as soon as the call returns
    auto drain = [&](int code) {
        cudaStreamSynchronize(G.s_in);
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(G.s_out);
        return code;
    };
    int k = 0;
    for (long c0 = 0; c0 < ncol; c0 += chunk, ++k) {
        const int a = k & 1;
        const int nc = (int)((ncol - c0) < chunk ? (ncol - c0) : chunk);
        if (k >= 2) CK(cudaStreamWaitEvent(G.s_in, G.e_out[a], 0));  // arena a: outputs of pass k-2 have left
        Stager sg{G.stage + (size_t)a * half, 0, P, G.s_in};
        ColIn in{};
        rc = stage_in(sg, in, c0, nc);
        if (rc) return drain(rc);
        in.ld = P; in.nlay = nlay;
        ColOut out{};
        out.ld = P;
        out.uflx = sg.carve((size_t)P * (nlay + 1)); out.dflx = sg.carve((size_t)P * (nlay + 1));
        out.uflxc = sg.carve((size_t)P * (nlay + 1)); out.dflxc = sg.carve((size_t)P * (nlay + 1));
        out.duflx_dt = sg.carve((size_t)P * (nlay + 1)); out.duflxc_dt = sg.carve((size_t)P * (nlay + 1));
        out.hr = sg.carve((size_t)P * nlay); out.hrc = sg.carve((size_t)P * nlay);
        if (cudaEventRecord(G.e_in[a], G.s_in) != cudaSuccess || cudaStreamWaitEvent(st, G.e_in[a], 0) != cudaSuccess)
            return drain(RRTMG_LW_ERR_CUDA);
        const long dbg_save = G.dbg_col;
        if (G.dbg) G.dbg_col = dbg_save - c0;  // chunk-relative for the staged pass
        G.chunk = chunk;
        G.err_col0 = c0;  // a staged pass sees its columns as 0 .. nc-1; input errors are reported with the call's column
        G.mt_total = G.mt_opt_total ? G.mt_opt_total : (long)ncol; G.mt_base = G.mt_opt_base + c0;
        rc = run_device(mode, nc, nlay, icld, idrv, inflag, iceflag, liqflag, permuteseed, irng, in, out, st);
        G.mt_total = 0; G.mt_base = 0;
        G.err_col0 = 0;
        G.chunk = saved_chunk;
        G.dbg_col = dbg_save;
        if (rc) return drain(rc);
        if (cudaEventRecord(G.e_done[a], st) != cudaSuccess || cudaStreamWaitEvent(G.s_out, G.e_done[a], 0) != cudaSuccess)
            return drain(RRTMG_LW_ERR_CUDA);
        sg.st = G.s_out;
        if (sg.out(ho.uflx, out.uflx, 1, ld, c0, nc, nlay + 1) || sg.out(ho.dflx, out.dflx, 1, ld, c0, nc, nlay + 1) ||
            sg.out(ho.uflxc, out.uflxc, 1, ld, c0, nc, nlay + 1) || sg.out(ho.dflxc, out.dflxc, 1, ld, c0, nc, nlay + 1) ||
            sg.out(ho.hr, out.hr, 1, ld, c0, nc, nlay) || sg.out(ho.hrc, out.hrc, 1, ld, c0, nc, nlay))
            return drain(RRTMG_LW_ERR_CUDA);
        if (idrv == 1 && (sg.out(ho.duflx_dt, out.duflx_dt, 1, ld, c0, nc, nlay + 1) ||
                          sg.out(ho.duflxc_dt, out.duflxc_dt, 1, ld, c0, nc, nlay + 1)))
            return drain(RRTMG_LW_ERR_CUDA);
        if (cudaEventRecord(G.e_out[a], G.s_out) != cudaSuccess) return drain(RRTMG_LW_ERR_CUDA);
    }
    drain(RRTMG_LW_OK);
    return decode_status(nullptr, nullptr);
}

void reset_timing() {
    G.ms_band = 0.f;
    G.ms_all = 0.f;
    G.nstage = 0;
    for (auto& m : G.ms_bands) m = 0.f;
}

bool check_common(int ncol, int ld, int nlay) { return ncol >= 1 && ld >= ncol && nlay >= 1 && nlay <= LW_MAXLAY; }

}  // namespace
}  // namespace lw

using namespace lw;

// Array-by-array staging of the unreduced k-data (what a Fortran shim has at hand after the
// reference's lw_kgb01..16 filled modules rrlw_kg01..16); rrtmg_lw_b200_ini(.., NULL, 0, ..) then
// consumes the staged blob.
static std::vector<double> g_kstage;
static std::vector<char> g_khave;

extern "C" int rrtmg_lw_b200_kdata_put(int band, const char* name, const double* data, long n) {
    LW_ENTER();
    if (!name || !data) return RRTMG_LW_ERR_ARG;
    if (g_kstage.empty()) {
        g_kstage.assign(RRTMG_LW_KDATA_DOUBLES, 0.0);
        g_khave.assign(RRTMG_LW_NKARRAYS, 0);
    }
    size_t off = 0;
    for (int k = 0; k < RRTMG_LW_NKARRAYS; ++k) {
        const rrtmg_lw_karray_desc& d = RRTMG_LW_KARRAYS[k];
        const size_t cnt = (size_t)16 * d.lead;
        if (d.band == band && std::strcmp(d.name, name) == 0) {
            if ((size_t)n != cnt) return RRTMG_LW_ERR_ARG;
            std::memcpy(g_kstage.data() + off, data, cnt * sizeof(double));
            g_khave[k] = 1;
            return RRTMG_LW_OK;
        }
        off += cnt;
    }
    return RRTMG_LW_ERR_ARG;  // unknown (band, name)
}

static int upload_tables();  // G.H -> the current device

// host-only half of rrtmg_lw_ini: the reduced tables as a versioned binary image (no CUDA device needed)
extern "C" long rrtmg_lw_b200_tables_build(double cpdair, const double* kdata, long nkdata, int default_real_kind,
                                           void* out, long cap) {
    if (!kdata || nkdata != RRTMG_LW_KDATA_DOUBLES) return -RRTMG_LW_ERR_ARG;
    if (default_real_kind != 4 && default_real_kind != 8) return -RRTMG_LW_ERR_ARG;
    if (out && cap < 0) return -RRTMG_LW_ERR_ARG;
    HostTables H;
    build_host_tables(cpdair, kStaticTables, kdata, default_real_kind, H);
    return host_tables_write(H, out, out ? cap : 0);
}

// rrtmg_lw_ini from such an image: no k-data, no reduction at start-up
extern "C" int rrtmg_lw_b200_ini_tables(const void* blob, long nbytes) {
    LW_ENTER();
    HostTables H;
    if (!host_tables_read(blob, nbytes, H)) return RRTMG_LW_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::fprintf(stderr, "rrtmg_lw_b200: no CUDA device; this library has no CPU path\n");
        return RRTMG_LW_ERR_CUDA;
    }
    G.H = std::move(H);
    return upload_tables();
}

extern "C" int rrtmg_lw_b200_ini(double cpdair, const double* kdata, long nkdata, int default_real_kind) {
    LW_ENTER();
    if (!kdata && nkdata == 0) {
        if (g_khave.empty()) return RRTMG_LW_ERR_ARG;
        for (char h : g_khave)
            if (!h) return RRTMG_LW_ERR_ARG;  // an array was never supplied
        kdata = g_kstage.data();
        nkdata = RRTMG_LW_KDATA_DOUBLES;
    }
    if (!kdata || nkdata != RRTMG_LW_KDATA_DOUBLES) return RRTMG_LW_ERR_ARG;
    if (default_real_kind != 4 && default_real_kind != 8) return RRTMG_LW_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::fprintf(stderr, "rrtmg_lw_b200: no CUDA device; this library has no CPU path\n");
        return RRTMG_LW_ERR_CUDA;
    }
    static_assert(sizeof(kStaticTables) / sizeof(double) == 9110, "static table blob size");
    build_host_tables(cpdair, kStaticTables, kdata, default_real_kind, G.H);
    return upload_tables();
}

static int upload_tables() {
    HostTables& H = G.H;
    auto up = [&](double** d, const std::vector<double>& h) -> int {
        if (*d) cudaFree(*d);
        *d = nullptr;
        CK(cudaMalloc((void**)d, h.size() * sizeof(double)));
        CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    };
    if (up(&G.d_arena, H.arena)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_arena_p, H.arena_p)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_totplnk, H.totplnk)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_totplnkderiv, H.totplnkderiv)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_tau, H.tau_tbl)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_absice2, H.absice2)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_absice3, H.absice3)) return RRTMG_LW_ERR_CUDA;
    if (up(&G.d_absliq1, H.absliq1)) return RRTMG_LW_ERR_CUDA;
    std::vector<double> et(2 * (NTBL + 1));
    for (int i = 0; i <= NTBL; ++i) { et[2 * i] = H.exp_tbl[i]; et[2 * i + 1] = H.tfn_tbl[i]; }
    if (up(&G.d_exptfn, et)) return RRTMG_LW_ERR_CUDA;

    // jp-indexed tables + chi ratios (setcoef.f90:338-351, 392-396: chi_mls(a,jp)/chi_mls(b,jp))
    GTab gt;
    std::memcpy(gt.preflog, H.ct.preflog, sizeof(gt.preflog));
    std::memcpy(gt.tref, H.ct.tref, sizeof(gt.tref));
    std::memcpy(gt.chi, H.ct.chi_mls, sizeof(gt.chi));
    auto chi = [&](int sp, int lev) { return H.ct.chi_mls[(sp - 1) + 7 * (lev - 1)]; };
    const int num[6] = {1, 1, 1, 1, 4, 3}, den[6] = {2, 3, 4, 6, 2, 2};
    for (int k = 0; k < 6; ++k)
        for (int jp = 1; jp <= 59; ++jp) gt.rat[k][jp - 1] = chi(num[k], jp) / chi(den[k], jp);
    if (!G.d_gtab) CK(cudaMalloc((void**)&G.d_gtab, sizeof(GTab)));
    CK(cudaMemcpy(G.d_gtab, &gt, sizeof(GTab), cudaMemcpyHostToDevice));

    // refrat_* constants of taumol (taumol.f90:504-513, 812-815, 1080-1086, 1442-1445, 1839-1842, 2242,
    // 2454-2460, 2770-2773, 2991)
    double rr[NBND][5];
    std::memset(rr, 0, sizeof(rr));
    rr[2][0] = chi(1, 9) / chi(2, 9);   rr[2][1] = chi(1, 13) / chi(2, 13); rr[2][2] = chi(1, 3) / chi(2, 3); rr[2][3] = chi(1, 13) / chi(2, 13);
    rr[3][0] = chi(1, 11) / chi(2, 11); rr[3][1] = chi(3, 13) / chi(2, 13);
    rr[4][0] = chi(1, 5) / chi(2, 5);   rr[4][1] = chi(3, 43) / chi(2, 43); rr[4][2] = chi(1, 7) / chi(2, 7);
    rr[6][0] = chi(1, 3) / chi(3, 3);   rr[6][2] = chi(1, 3) / chi(3, 3);
    rr[8][0] = chi(1, 9) / chi(6, 9);   rr[8][2] = chi(1, 3) / chi(6, 3);
    rr[11][0] = chi(1, 10) / chi(2, 10);
    rr[12][0] = chi(1, 5) / chi(4, 5);  rr[12][2] = chi(1, 1) / chi(4, 1);  rr[12][4] = chi(1, 3) / chi(4, 3);
    rr[14][0] = chi(4, 1) / chi(2, 1);  rr[14][2] = chi(4, 1) / chi(2, 1);
    rr[15][0] = chi(1, 6) / chi(6, 6);

    // band pointers into the device arena
    for (int b = 0; b < NBND; ++b) {
        auto P = [&](int slot) -> const double* {
            return H.off[b][slot] == (size_t)-1 ? nullptr : G.d_arena + H.off[b][slot];
        };
        BandTab& T = H.ct.band[b];
        T.ka = P(0); T.kb = P(1); T.selfref = P(2); T.forref = P(3); T.fracrefa = P(4); T.fracrefb = P(5);
        for (int m = 0; m < 5; ++m) T.m[m] = P(6 + m);
        T.v[0] = P(11); T.v[1] = P(12);
        auto Q = [&](int slot) -> const double* {
            return H.offp[b][slot] == (size_t)-1 ? nullptr : G.d_arena_p + H.offp[b][slot];
        };
        BandTab& U = H.ct.bandp[b];
        U.ka = Q(0); U.kb = Q(1); U.selfref = Q(2); U.forref = Q(3); U.fracrefa = Q(4); U.fracrefb = Q(5);
        for (int m = 0; m < 5; ++m) U.m[m] = Q(6 + m);
        U.v[0] = Q(11); U.v[1] = Q(12);
    }
    {
        SweepTab& tb = G.sweep_tab;
        tb.exptfn = reinterpret_cast<const double2*>(G.d_exptfn);
        tb.tau_tbl = G.d_tau; tb.totplnk = G.d_totplnk; tb.totplnkderiv = G.d_totplnkderiv;
        for (int b = 0; b < NBND; ++b) {
            tb.fracrefa[b] = H.ct.bandp[b].fracrefa; tb.fracrefb[b] = H.ct.bandp[b].fracrefb;
            tb.delwave[b] = H.ct.delwave[b];
        }
        tb.bpade = H.ct.bpade; tb.fluxfac = H.ct.fluxfac;
    }
    DevPtrs dp;
    dp.g = G.d_gtab; dp.totplnk = G.d_totplnk; dp.totplnkderiv = G.d_totplnkderiv;
    dp.exptfn = reinterpret_cast<const double2*>(G.d_exptfn); dp.tau_tbl = G.d_tau;
    dp.absice2 = G.d_absice2; dp.absice3 = G.d_absice3; dp.absliq1 = G.d_absliq1;
    // ---- float copies for the optional FP32 mode (set_option("precision", 32))
    for (float* p : G.f32_allocs) cudaFree(p);
    G.f32_allocs.clear();
    auto upf = [&](const double* h, size_t n, const float** d) -> int {
        std::vector<float> f(n);
        for (size_t k = 0; k < n; ++k) f[k] = (float)h[k];
        float* p = nullptr;
        CK(cudaMalloc((void**)&p, n * sizeof(float)));
        CK(cudaMemcpy(p, f.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        G.f32_allocs.push_back(p);
        *d = p;
        return 0;
    };
    const float* f_arena = nullptr;
    const float* f_et = nullptr;
    const float* f_gt = nullptr;
    if (upf(H.arena_p.data(), H.arena_p.size(), &f_arena) || upf(H.totplnk.data(), H.totplnk.size(), &dp.totplnk_f) ||
        upf(H.totplnkderiv.data(), H.totplnkderiv.size(), &dp.totplnkderiv_f) ||
        upf(H.tau_tbl.data(), H.tau_tbl.size(), &dp.tau_tbl_f) || upf(H.absice2.data(), H.absice2.size(), &dp.absice2_f) ||
        upf(H.absice3.data(), H.absice3.size(), &dp.absice3_f) || upf(H.absliq1.data(), H.absliq1.size(), &dp.absliq1_f) ||
        upf(et.data(), et.size(), &f_et) ||
        upf(reinterpret_cast<const double*>(&gt), sizeof(GTab) / sizeof(double), &f_gt))
        return RRTMG_LW_ERR_CUDA;
    static_assert(sizeof(GTabF) / sizeof(float) == sizeof(GTab) / sizeof(double), "GTabF mirrors GTab");
    dp.exptfn_f = reinterpret_cast<const float2*>(f_et);
    dp.gf = reinterpret_cast<const GTabF*>(f_gt);
    ConstTabF cf{};
    for (int b = 0; b < NBND; ++b) {
        auto Q = [&](int slot) -> const float* {
            return H.offp[b][slot] == (size_t)-1 ? nullptr : f_arena + H.offp[b][slot];
        };
        BandTabF& U = cf.bandp[b];
        U.ka = Q(0); U.kb = Q(1); U.selfref = Q(2); U.forref = Q(3); U.fracrefa = Q(4); U.fracrefb = Q(5);
        for (int m = 0; m < 5; ++m) U.m[m] = Q(6 + m);
        U.v[0] = Q(11); U.v[1] = Q(12);
        cf.band[b] = U;
        cf.delwave[b] = (float)H.ct.delwave[b];
        cf.a0[b] = (float)H.ct.a0[b]; cf.a1[b] = (float)H.ct.a1[b]; cf.a2[b] = (float)H.ct.a2[b];
    }
    cf.fluxfac = (float)H.ct.fluxfac; cf.oneminus = (float)H.ct.oneminus; cf.bpade = (float)H.ct.bpade;
    cf.e20_b3up = (float)H.ct.e20_b3up; cf.e20_b7 = (float)H.ct.e20_b7;
    for (int k = 0; k < 7; ++k) cf.b4scale[k] = (float)H.ct.b4scale[k];
    cf.co2ref_b13 = (float)H.ct.co2ref_b13;
    for (int k = 0; k < 2; ++k) cf.absice0[k] = (float)H.ct.absice0[k];
    for (int k = 0; k < 10; ++k) cf.absice1[k] = (float)H.ct.absice1[k];
    cf.absliq0 = (float)H.ct.absliq0;
    CK(upload_constants_kernels(H.ct, dp, rr));
    CK(upload_constants_tu0(H.ct, dp, rr, &cf));
    CK(upload_constants_tu1(H.ct, dp, rr, &cf));
    CK(upload_constants_tu2(H.ct, dp, rr, &cf));
    CK(upload_constants_tu3(H.ct, dp, rr, &cf));
    CK(upload_constants_tu4(H.ct, dp, rr, &cf));
    CK(upload_constants_tu5(H.ct, dp, rr, &cf));
    CK(upload_constants_stu0(H.ct, dp, rr, &cf));
    CK(upload_constants_stu1(H.ct, dp, rr, &cf));
    CK(upload_constants_stu2(H.ct, dp, rr, &cf));
    CK(upload_constants_stu3(H.ct, dp, rr, &cf));
    CK(upload_constants_stu4(H.ct, dp, rr, &cf));
    CK(upload_constants_stu5(H.ct, dp, rr, &cf));
    CK(upload_constants_stu6(H.ct, dp, rr, &cf));
    CK(cudaDeviceSynchronize());
    G.inited = true;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_rrtmg_lw(int ncol, int ld, int nlay, int* icld, int idrv, const double* play,
                                      const double* plev, const double* tlay, const double* tlev, const double* tsfc,
                                      const double* h2ovmr, const double* o3vmr, const double* co2vmr,
                                      const double* ch4vmr, const double* n2ovmr, const double* o2vmr,
                                      const double* cfc11vmr, const double* cfc12vmr, const double* cfc22vmr,
                                      const double* ccl4vmr, const double* emis, int inflglw, int iceflglw, int liqflglw,
                                      const double* cldfmcl, const double* taucmcl, const double* ciwpmcl,
                                      const double* clwpmcl, const double* reicmcl, const double* relqmcl,
                                      const double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc,
                                      double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt, int memspace,
                                      void* stream) {
    LW_ENTER();
    if (!G.inited) return RRTMG_LW_ERR_NOT_INIT;
    if (!icld || !check_common(ncol, ld, nlay)) return RRTMG_LW_ERR_ARG;
    if (*icld < 0 || *icld > 3) *icld = 2;  // rad.f90:470
    if (idrv != 0 && idrv != 1) return RRTMG_LW_ERR_ARG;
    if (idrv == 1 && (!duflx_dt || !duflxc_dt)) return RRTMG_LW_ERR_ARG;
    const bool cloudy = *icld >= 1;
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !ch4vmr || !n2ovmr || !o2vmr ||
        !cfc11vmr || !cfc12vmr || !cfc22vmr || !ccl4vmr || !emis || !tauaer || !uflx || !dflx || !hr || !uflxc ||
        !dflxc || !hrc)
        return RRTMG_LW_ERR_ARG;
    if (cloudy && (!cldfmcl || !taucmcl || !ciwpmcl || !clwpmcl || !reicmcl || !relqmcl)) return RRTMG_LW_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    reset_timing();
    if (memspace == RRTMG_LW_DEVICE) {
        ColIn in{};
        in.play = play; in.plev = plev; in.tlay = tlay; in.tlev = tlev; in.tsfc = tsfc; in.h2o = h2ovmr; in.o3 = o3vmr;
        in.co2 = co2vmr; in.ch4 = ch4vmr; in.n2o = n2ovmr; in.o2 = o2vmr; in.cfc11 = cfc11vmr; in.cfc12 = cfc12vmr;
        in.cfc22 = cfc22vmr; in.ccl4 = ccl4vmr; in.emis = emis; in.tauaer = tauaer;
        in.cldfmcl = cldfmcl; in.taucmcl = taucmcl; in.ciwpmcl = ciwpmcl; in.clwpmcl = clwpmcl;
        in.reicmcl = reicmcl; in.relqmcl = relqmcl;
        in.ld = ld; in.nlay = nlay;
        ColOut out{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, G.out_ld > 0 ? G.out_ld : (long)ld};
        return run_device(MODE_STRICT, ncol, nlay, *icld, idrv, inflglw, iceflglw, liqflglw, 0, 0, in, out, st);
    }
    // host pointers: staged, pipelined passes
    const size_t per_col = (size_t)12 * nlay + 2 * (nlay + 1) + 1 + 16 + 16 * (size_t)nlay +
                           (cloudy ? (size_t)4 * 140 * nlay + 2 * nlay : 0) + 6 * (size_t)(nlay + 1) + 2 * (size_t)nlay;
    auto stage_in = [&](Stager& sg, ColIn& in, long c0, int nc) -> int {
        double* d;
#define IN2(field, src, rows) if (sg.in(src, &d, 1, ld, c0, nc, rows)) return RRTMG_LW_ERR_CUDA; in.field = d;
        IN2(play, play, nlay) IN2(plev, plev, nlay + 1) IN2(tlay, tlay, nlay) IN2(tlev, tlev, nlay + 1) IN2(tsfc, tsfc, 1)
        IN2(h2o, h2ovmr, nlay) IN2(o3, o3vmr, nlay) IN2(co2, co2vmr, nlay) IN2(ch4, ch4vmr, nlay) IN2(n2o, n2ovmr, nlay)
        IN2(o2, o2vmr, nlay) IN2(cfc11, cfc11vmr, nlay) IN2(cfc12, cfc12vmr, nlay) IN2(cfc22, cfc22vmr, nlay)
        IN2(ccl4, ccl4vmr, nlay) IN2(emis, emis, 16) IN2(tauaer, tauaer, (long)16 * nlay)
        if (cloudy) {
#define IN3(field, src) if (sg.in(src, &d, 140, ld, c0, nc, nlay)) return RRTMG_LW_ERR_CUDA; in.field = d;
            IN3(cldfmcl, cldfmcl) IN3(taucmcl, taucmcl) IN3(ciwpmcl, ciwpmcl) IN3(clwpmcl, clwpmcl)
#undef IN3
            IN2(reicmcl, reicmcl, nlay) IN2(relqmcl, relqmcl, nlay)
        }
#undef IN2
        return 0;
    };
    const HostOut ho{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt};
    return host_pipeline(MODE_STRICT, ncol, ld, nlay, *icld, idrv, inflglw, iceflglw, liqflglw, 0, 0, per_col, stage_in, ho, st);
}

extern "C" int rrtmg_lw_b200_rrtmg_lw_mcica(int ncol, int ld, int nlay, int icld_overlap, int idrv, int permuteseed,
                                            int irng, const double* play, const double* plev, const double* tlay,
                                            const double* tlev, const double* tsfc, const double* h2ovmr,
                                            const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                                            const double* n2ovmr, const double* o2vmr, const double* cfc11vmr,
                                            const double* cfc12vmr, const double* cfc22vmr, const double* ccl4vmr,
                                            const double* emis, int inflglw, int iceflglw, int liqflglw,
                                            const double* cldfrac, const double* ciwp, const double* clwp,
                                            const double* rei, const double* rel, const double* tauc,
                                            const double* alpha, const double* tauaer, double* uflx, double* dflx,
                                            double* hr, double* uflxc, double* dflxc, double* hrc, double* duflx_dt,
                                            double* duflxc_dt, int memspace, void* stream) {
    LW_ENTER();
    if (!G.inited) return RRTMG_LW_ERR_NOT_INIT;
    if (!check_common(ncol, ld, nlay)) return RRTMG_LW_ERR_ARG;
    if (icld_overlap < 0 || icld_overlap > 5) return RRTMG_LW_ERR_FLAG;  // 'MCICA_SUBCOL_LW: INVALID ICLD'
    if (idrv != 0 && idrv != 1) return RRTMG_LW_ERR_ARG;
    if (idrv == 1 && (!duflx_dt || !duflxc_dt)) return RRTMG_LW_ERR_ARG;
    const bool cloudy = icld_overlap >= 1;
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !ch4vmr || !n2ovmr || !o2vmr ||
        !cfc11vmr || !cfc12vmr || !cfc22vmr || !ccl4vmr || !emis || !tauaer || !uflx || !dflx || !hr || !uflxc ||
        !dflxc || !hrc)
        return RRTMG_LW_ERR_ARG;
    if (cloudy && (!cldfrac || !ciwp || !clwp || !rei || !rel || !tauc)) return RRTMG_LW_ERR_ARG;
    if (icld_overlap >= 4 && !alpha) return RRTMG_LW_ERR_ARG;
    if (irng != 0) irng = 1;
    cudaStream_t st = (cudaStream_t)stream;
    reset_timing();
    if (memspace == RRTMG_LW_DEVICE) {
        ColIn in{};
        in.play = play; in.plev = plev; in.tlay = tlay; in.tlev = tlev; in.tsfc = tsfc; in.h2o = h2ovmr; in.o3 = o3vmr;
        in.co2 = co2vmr; in.ch4 = ch4vmr; in.n2o = n2ovmr; in.o2 = o2vmr; in.cfc11 = cfc11vmr; in.cfc12 = cfc12vmr;
        in.cfc22 = cfc22vmr; in.ccl4 = ccl4vmr; in.emis = emis; in.tauaer = tauaer;
        in.cldfrac = cldfrac; in.ciwp = ciwp; in.clwp = clwp; in.rei = rei; in.rel = rel; in.tauc = tauc; in.alpha = alpha;
        in.ld = ld; in.nlay = nlay;
        ColOut out{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, G.out_ld > 0 ? G.out_ld : (long)ld};
        return run_device(MODE_FUSED, ncol, nlay, icld_overlap, idrv, inflglw, iceflglw, liqflglw, permuteseed, irng, in,
                          out, st);
    }
    const size_t per_col = (size_t)12 * nlay + 2 * (nlay + 1) + 1 + 16 + 16 * (size_t)nlay +
                           (cloudy ? (size_t)6 * nlay + 16 * (size_t)nlay : 0) + 6 * (size_t)(nlay + 1) + 2 * (size_t)nlay;
    auto stage_in = [&](Stager& sg, ColIn& in, long c0, int nc) -> int {
        double* d;
#define IN2(field, src, rows) if (sg.in(src, &d, 1, ld, c0, nc, rows)) return RRTMG_LW_ERR_CUDA; in.field = d;
        IN2(play, play, nlay) IN2(plev, plev, nlay + 1) IN2(tlay, tlay, nlay) IN2(tlev, tlev, nlay + 1) IN2(tsfc, tsfc, 1)
        IN2(h2o, h2ovmr, nlay) IN2(o3, o3vmr, nlay) IN2(co2, co2vmr, nlay) IN2(ch4, ch4vmr, nlay) IN2(n2o, n2ovmr, nlay)
        IN2(o2, o2vmr, nlay) IN2(cfc11, cfc11vmr, nlay) IN2(cfc12, cfc12vmr, nlay) IN2(cfc22, cfc22vmr, nlay)
        IN2(ccl4, ccl4vmr, nlay) IN2(emis, emis, 16) IN2(tauaer, tauaer, (long)16 * nlay)
        if (cloudy) {
            IN2(cldfrac, cldfrac, nlay) IN2(ciwp, ciwp, nlay) IN2(clwp, clwp, nlay) IN2(rei, rei, nlay) IN2(rel, rel, nlay)
            if (alpha) { IN2(alpha, alpha, nlay) }
            if (sg.in(tauc, &d, 16, ld, c0, nc, nlay)) return RRTMG_LW_ERR_CUDA;
            in.tauc = d;
        }
#undef IN2
        return 0;
    };
    const HostOut ho{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt};
    return host_pipeline(MODE_FUSED, ncol, ld, nlay, icld_overlap, idrv, inflglw, iceflglw, liqflglw, permuteseed, irng,
                         per_col, stage_in, ho, st);
}

// rrtmg_lw of src/rrtmg_lw_rad.nomcica.f90:99-108: grid-mean cloud arrays, cldprop + rtrn (icld = 1) / rtrnmr (else)
extern "C" int rrtmg_lw_b200_rrtmg_lw_nomcica(int ncol, int ld, int nlay, int* icld, int idrv, const double* play,
                                              const double* plev, const double* tlay, const double* tlev,
                                              const double* tsfc, const double* h2ovmr, const double* o3vmr,
                                              const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                                              const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                                              const double* cfc22vmr, const double* ccl4vmr, const double* emis,
                                              int inflglw, int iceflglw, int liqflglw, const double* cldfr,
                                              const double* taucld, const double* cicewp, const double* cliqwp,
                                              const double* reice, const double* reliq, const double* tauaer,
                                              double* uflx, double* dflx, double* hr, double* uflxc, double* dflxc,
                                              double* hrc, double* duflx_dt, double* duflxc_dt, int memspace,
                                              void* stream) {
    LW_ENTER();
    if (!G.inited) return RRTMG_LW_ERR_NOT_INIT;
    if (!icld || !check_common(ncol, ld, nlay)) return RRTMG_LW_ERR_ARG;
    if (*icld < 0 || *icld > 3) *icld = 2;  // rad.nomcica.f90:458
    if (idrv != 0 && idrv != 1) return RRTMG_LW_ERR_ARG;
    if (idrv == 1 && (!duflx_dt || !duflxc_dt)) return RRTMG_LW_ERR_ARG;
    const bool cloudy = *icld >= 1;
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !ch4vmr || !n2ovmr || !o2vmr ||
        !cfc11vmr || !cfc12vmr || !cfc22vmr || !ccl4vmr || !emis || !tauaer || !uflx || !dflx || !hr || !uflxc ||
        !dflxc || !hrc)
        return RRTMG_LW_ERR_ARG;
    if (cloudy && (!cldfr || !taucld || !cicewp || !cliqwp || !reice || !reliq)) return RRTMG_LW_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    reset_timing();
    if (memspace == RRTMG_LW_DEVICE) {
        ColIn in{};
        in.play = play; in.plev = plev; in.tlay = tlay; in.tlev = tlev; in.tsfc = tsfc; in.h2o = h2ovmr; in.o3 = o3vmr;
        in.co2 = co2vmr; in.ch4 = ch4vmr; in.n2o = n2ovmr; in.o2 = o2vmr; in.cfc11 = cfc11vmr; in.cfc12 = cfc12vmr;
        in.cfc22 = cfc22vmr; in.ccl4 = ccl4vmr; in.emis = emis; in.tauaer = tauaer;
        in.cldfrac = cldfr; in.ciwp = cicewp; in.clwp = cliqwp; in.rei = reice; in.rel = reliq; in.tauc = taucld;
        in.ld = ld; in.nlay = nlay;
        ColOut out{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, G.out_ld > 0 ? G.out_ld : (long)ld};
        return run_device(MODE_GRID, ncol, nlay, *icld, idrv, inflglw, iceflglw, liqflglw, 0, 0, in, out, st);
    }
    const size_t per_col = (size_t)12 * nlay + 2 * (nlay + 1) + 1 + 16 + 16 * (size_t)nlay +
                           (cloudy ? (size_t)5 * nlay + 16 * (size_t)nlay : 0) + 6 * (size_t)(nlay + 1) + 2 * (size_t)nlay;
    auto stage_in = [&](Stager& sg, ColIn& in, long c0, int nc) -> int {
        double* d;
#define IN2(field, src, rows) if (sg.in(src, &d, 1, ld, c0, nc, rows)) return RRTMG_LW_ERR_CUDA; in.field = d;
        IN2(play, play, nlay) IN2(plev, plev, nlay + 1) IN2(tlay, tlay, nlay) IN2(tlev, tlev, nlay + 1) IN2(tsfc, tsfc, 1)
        IN2(h2o, h2ovmr, nlay) IN2(o3, o3vmr, nlay) IN2(co2, co2vmr, nlay) IN2(ch4, ch4vmr, nlay) IN2(n2o, n2ovmr, nlay)
        IN2(o2, o2vmr, nlay) IN2(cfc11, cfc11vmr, nlay) IN2(cfc12, cfc12vmr, nlay) IN2(cfc22, cfc22vmr, nlay)
        IN2(ccl4, ccl4vmr, nlay) IN2(emis, emis, 16) IN2(tauaer, tauaer, (long)16 * nlay)
        if (cloudy) {
            IN2(cldfrac, cldfr, nlay) IN2(ciwp, cicewp, nlay) IN2(clwp, cliqwp, nlay) IN2(rei, reice, nlay)
            IN2(rel, reliq, nlay)
            if (sg.in(taucld, &d, 16, ld, c0, nc, nlay)) return RRTMG_LW_ERR_CUDA;
            in.tauc = d;
        }
#undef IN2
        return 0;
    };
    const HostOut ho{uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt};
    return host_pipeline(MODE_GRID, ncol, ld, nlay, *icld, idrv, inflglw, iceflglw, liqflglw, 0, 0, per_col, stage_in, ho, st);
}

extern "C" int rrtmg_lw_b200_mcica_subcol_lw(int iplon, int ncol, int ld, int nlay, int icld, int permuteseed, int* irng,
                                             const double* play, const double* cldfrac, const double* ciwp,
                                             const double* clwp, const double* rei, const double* rel,
                                             const double* tauc, const double* alpha, double* cldfmcl, double* ciwpmcl,
                                             double* clwpmcl, double* reicmcl, double* relqmcl, double* taucmcl,
                                             int memspace, void* stream) {
    (void)iplon;
    LW_ENTER();
    if (!G.inited) return RRTMG_LW_ERR_NOT_INIT;
    if (icld == 0) return RRTMG_LW_OK;  // mcica:257
    if (icld < 0 || icld > 5) return RRTMG_LW_ERR_FLAG;
    if (!irng || !check_common(ncol, ld, nlay) || (nlay < 4 && *irng == 0)) return RRTMG_LW_ERR_ARG;
    if (!play || !cldfrac || !ciwp || !clwp || !rei || !rel || !tauc || !cldfmcl || !ciwpmcl || !clwpmcl || !reicmcl ||
        !relqmcl || !taucmcl)
        return RRTMG_LW_ERR_ARG;
    if (icld >= 4 && !alpha) return RRTMG_LW_ERR_ARG;
    if (*irng != 0) *irng = 1;
    cudaStream_t st = (cudaStream_t)stream;
    // host pointers: the staged pass holds the four (140, ncol, nlay) output arrays, so it follows host_chunk
    int cap = G.chunk;
    if (memspace == RRTMG_LW_HOST && G.host_chunk < cap) cap = G.host_chunk;
    const int chunk = cap < ncol ? cap : ncol;
    int rc = ensure_scratch(chunk, nlay, 0);
    if (rc) return rc;
    const long P = ((chunk + 31) / 32) * 32;
    if (memspace == RRTMG_LW_HOST) {
        rc = ensure_stage(((size_t)7 * nlay + 16 * (size_t)nlay + 4 * 140 * (size_t)nlay + 2 * (size_t)nlay) * P + 64);
        if (rc) return rc;
    }
    for (long c0 = 0; c0 < ncol; c0 += chunk) {
        RunCfg r{};
        r.ncol = (int)((ncol - c0) < chunk ? (ncol - c0) : chunk);
        r.nlay = nlay; r.icld = icld; r.irng = *irng; r.permuteseed = permuteseed;
        ColIn in{};
        in.nlay = nlay;
        auto generate = [&]() -> int {
            if (*irng == 0) { launch_mcica_kiss(in, G.S, r, st); G.launches++; return RRTMG_LW_OK; }
            if (G.mt_opt_base + ncol > (G.mt_opt_total ? G.mt_opt_total : (long)ncol)) return RRTMG_LW_ERR_ARG;
            const int nl = launch_mcica_mt(in, G.S, r, G.mt_opt_total ? G.mt_opt_total : (long)ncol, G.mt_opt_base + c0, st);
            if (nl < 0) return RRTMG_LW_ERR_CUDA;
            G.launches += nl;
            return RRTMG_LW_OK;
        };
        if (memspace == RRTMG_LW_DEVICE) {
            r.c0 = c0;
            in.play = play; in.cldfrac = cldfrac; in.ciwp = ciwp; in.clwp = clwp; in.rei = rei; in.rel = rel;
            in.tauc = tauc; in.alpha = alpha; in.ld = ld;
            if ((rc = generate()) != RRTMG_LW_OK) return rc;
            launch_materialize(in, G.S, r, cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl, st);
            G.launches++;
        } else {
            r.c0 = 0;
            Stager sg{G.stage, 0, P, st};
            double* d;
#define IN2(field, src) if (sg.in(src, &d, 1, ld, c0, r.ncol, nlay)) return RRTMG_LW_ERR_CUDA; in.field = d;
            IN2(play, play) IN2(cldfrac, cldfrac) IN2(ciwp, ciwp) IN2(clwp, clwp) IN2(rei, rei) IN2(rel, rel)
            if (alpha) { IN2(alpha, alpha) }
#undef IN2
            if (sg.in(tauc, &d, 16, ld, c0, r.ncol, nlay)) return RRTMG_LW_ERR_CUDA;
            in.tauc = d;
            in.ld = P;
            double* o_cld = sg.carve((size_t)140 * P * nlay); double* o_ciwp = sg.carve((size_t)140 * P * nlay);
            double* o_clwp = sg.carve((size_t)140 * P * nlay); double* o_tauc = sg.carve((size_t)140 * P * nlay);
            double* o_rei = sg.carve((size_t)P * nlay); double* o_rel = sg.carve((size_t)P * nlay);
            if ((rc = generate()) != RRTMG_LW_OK) return rc;
            launch_materialize(in, G.S, r, o_cld, o_ciwp, o_clwp, o_rei, o_rel, o_tauc, st);
            G.launches++;
            if (sg.out(cldfmcl, o_cld, 140, ld, c0, r.ncol, nlay) || sg.out(ciwpmcl, o_ciwp, 140, ld, c0, r.ncol, nlay) ||
                sg.out(clwpmcl, o_clwp, 140, ld, c0, r.ncol, nlay) || sg.out(taucmcl, o_tauc, 140, ld, c0, r.ncol, nlay) ||
                sg.out(reicmcl, o_rei, 1, ld, c0, r.ncol, nlay) || sg.out(relqmcl, o_rel, 1, ld, c0, r.ncol, nlay))
                return RRTMG_LW_ERR_CUDA;
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaGetLastError());
    }
    if (memspace == RRTMG_LW_HOST) {
        // the generator itself only STOPs on the KISS pressure ordering (mcica:465); radius checks belong to cldprmc
        int rcs = decode_status(nullptr, nullptr);
        return rcs == RRTMG_LW_ERR_KISS_PRESSURE ? rcs : RRTMG_LW_OK;
    }
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_get_alpha(int iplon, int ncol, int ld, int nlayers, int icld, int idcor, double decorr_con,
                                       const double* dz, const double* lat, int juldat, const double* cldfrac,
                                       double* alpha, int memspace, void* stream) {
    (void)iplon;
    LW_ENTER();
    if (!G.inited) return RRTMG_LW_ERR_NOT_INIT;
    if (!check_common(ncol, ld, nlayers) || !dz || !lat || !cldfrac || !alpha) return RRTMG_LW_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int drk = G.H.ct.drk;
    auto dkl = [&](double x) { return drk == 4 ? (double)(float)x : x; };
    const double am1 = dkl(1.4315), am2 = dkl(2.1219), am4 = dkl(-25.584);  // mcica:121-123 default-kind literals
    if (memspace == RRTMG_LW_DEVICE) {
        launch_get_alpha(ncol, ld, nlayers, icld, idcor, decorr_con, dz, lat, juldat, cldfrac, alpha, am1, am2, am4, st);
        G.launches++;
        CK(cudaGetLastError());
        return RRTMG_LW_OK;
    }
    const size_t n2 = (size_t)ld * nlayers;
    int rc = ensure_stage(3 * n2 + ld + 64);
    if (rc) return rc;
    double *d_dz = G.stage, *d_cf = G.stage + n2, *d_al = G.stage + 2 * n2, *d_lat = G.stage + 3 * n2;
    CK(cudaMemcpyAsync(d_dz, dz, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cf, cldfrac, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_al, alpha, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lat, lat, (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_get_alpha(ncol, ld, nlayers, icld, idcor, decorr_con, d_dz, d_lat, juldat, d_cf, d_al, am1, am2, am4, st);
    G.launches++;
    CK(cudaMemcpyAsync(alpha, d_al, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_sync(int* col, int* lay) {
    LW_ENTER();
    if (cudaDeviceSynchronize() != cudaSuccess) return RRTMG_LW_ERR_CUDA;
    return decode_status(col, lay);
}

extern "C" int rrtmg_lw_b200_last_error(int* col, int* lay) {
    LW_ENTER();
    if (col) *col = (int)G.last_err_col;
    if (lay) *lay = G.last_err_lay;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_debug_capture(int col_plus1) {
    LW_ENTER();
    G.dbg = col_plus1 > 0;
    G.dbg_col = col_plus1 - 1;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_debug_stages(int icol, int* ints6, double* facs, double* taug, double* fracs, int* itr,
                                          double* scal, unsigned char* mask) {
    LW_ENTER();
    if (!G.inited || !G.dbg || G.dbg_nlay <= 0) return RRTMG_LW_ERR_NOT_INIT;
    if (icol != G.dbg_col) return RRTMG_LW_ERR_ARG;
    const int nlay = G.dbg_nlay;
    CK(cudaDeviceSynchronize());
    if (ints6) CK(cudaMemcpy(ints6, G.S.dbg_ints, sizeof(int) * 6 * nlay, cudaMemcpyDeviceToHost));
    if (facs) CK(cudaMemcpy(facs, G.S.dbg_facs, sizeof(double) * 4 * nlay, cudaMemcpyDeviceToHost));
    if (taug) CK(cudaMemcpy(taug, G.S.dbg_taug, sizeof(double) * NGPT * nlay, cudaMemcpyDeviceToHost));
    if (fracs) CK(cudaMemcpy(fracs, G.S.dbg_fracs, sizeof(double) * NGPT * nlay, cudaMemcpyDeviceToHost));
    if (itr) CK(cudaMemcpy(itr, G.S.dbg_itr, sizeof(int) * 2 * NGPT * nlay, cudaMemcpyDeviceToHost));
    if (scal) CK(cudaMemcpy(scal, G.S.dbg_scal, sizeof(double) * 2, cudaMemcpyDeviceToHost));
    if (mask) {
        // only meaningful for a single-chunk call; chunk-relative column = icol % chunk
        const long P = G.S.pitch;
        const long ic = icol % G.chunk;
        std::vector<unsigned> w((size_t)5 * nlay);
        for (int l = 0; l < nlay; ++l)
            for (int k = 0; k < 5; ++k)
                CK(cudaMemcpy(&w[(size_t)l * 5 + k], G.S.mask + ((size_t)l * 5 + k) * P + ic, sizeof(unsigned),
                              cudaMemcpyDeviceToHost));
        for (int l = 0; l < nlay; ++l)
            for (int g = 0; g < NGPT; ++g) mask[(size_t)g * nlay + l] = (w[(size_t)l * 5 + (g >> 5)] >> (g & 31)) & 1u;
    }
    return RRTMG_LW_OK;
}

extern "C" long rrtmg_lw_b200_get_table(const char* name, double* out, long cap) {
    LW_ENTER();
    if (!G.inited || !name) return 0;
    const std::string n(name);
    const double* src = nullptr;
    long cnt = 0;
    if (n == "tau_tbl") { src = G.H.tau_tbl.data(); cnt = NTBL + 1; }
    else if (n == "exp_tbl") { src = G.H.exp_tbl.data(); cnt = NTBL + 1; }
    else if (n == "tfn_tbl") { src = G.H.tfn_tbl.data(); cnt = NTBL + 1; }
    else if (n == "rwgt") { src = G.H.rwgt; cnt = 256; }
    else if (n == "heatfac") { src = &G.H.ct.heatfac; cnt = 1; }   // rrlw_con, set by rrtmg_lw_ini (init.f90:298)
    else if (n == "fluxfac") { src = &G.H.ct.fluxfac; cnt = 1; }
    else {
        for (auto& kv : G.H.named)
            if (kv.first == n) { src = kv.second.data(); cnt = (long)kv.second.size(); break; }
    }
    if (src && out)
        for (long i = 0; i < cnt && i < cap; ++i) out[i] = src[i];
    return cnt;
}

extern "C" int rrtmg_lw_b200_set_device(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (device < -1 || device >= LW_MAX_DEVICES) return RRTMG_LW_ERR_ARG;
    if (device >= 0) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) return RRTMG_LW_ERR_CUDA;
        if (cudaSetDevice(device) != cudaSuccess) return RRTMG_LW_ERR_CUDA;
    }
    g_explicit = device;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_set_chunk(int max_columns) {
    LW_ENTER();
    if (max_columns < 32) return RRTMG_LW_ERR_ARG;
    G.chunk = max_columns;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_set_option(const char* name, int value) {
    LW_ENTER();
    if (!name) return RRTMG_LW_ERR_ARG;
    if (std::strcmp(name, "band_kernel") == 0) {
        if (value != 2 && value != 4) return RRTMG_LW_ERR_ARG;
        G.band_kernel = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "release_scratch") == 0) {
        // give the library-owned scratch back to the device (it is re-allocated by the next call that needs it)
        cudaDeviceSynchronize();
        free_scratch();
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "out_ld") == 0) {
        if (value < 0) return RRTMG_LW_ERR_ARG;
        G.out_ld = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "tail_group") == 0) {
        if (value != 0 && value != 1) return RRTMG_LW_ERR_ARG;
        G.tail_group = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "kiss_sparse") == 0) {
        if (value != 0 && value != 1) return RRTMG_LW_ERR_ARG;
        set_kiss_sparse(value);
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "concurrency") == 0) {
        if (value < 0 || value > 5) return RRTMG_LW_ERR_ARG;
        G.concurrency = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "mcica_beside") == 0) {
        if (value < 0 || value > 4) return RRTMG_LW_ERR_ARG;
        G.mcica_beside = value;  // resident work-queue CTAs per SM (0 = off)
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "sweep_groups") == 0) {
        if (value != 0 && value != 1) return RRTMG_LW_ERR_ARG;
        G.sweep_groups = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "band_lo") == 0 || std::strcmp(name, "band_hi") == 0) {
        // band range istart..iend of the following runs (two-phase path): only those bands enter the flux sums
        if (value < 1 || value > NBND) return RRTMG_LW_ERR_ARG;
        (name[5] == 'l' ? G.band_lo : G.band_hi) = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "precision") == 0) {
        // 64 (default) or 32: float band kernels (fluxes within 0.05 W m-2, heating rates within 0.01 K/day of FP64)
        if (value != 64 && value != 32) return RRTMG_LW_ERR_ARG;
        G.precision = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "mt_global_ncol") == 0) {
        // irng = 1 only: the columns of the following calls are a block of a batch of `value` columns (0 = off)
        if (value < 0) return RRTMG_LW_ERR_ARG;
        G.mt_opt_total = value;
        if (value == 0) G.mt_opt_base = 0;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "mt_col_offset") == 0) {
        if (value < 0) return RRTMG_LW_ERR_ARG;
        G.mt_opt_base = value;
        return RRTMG_LW_OK;
    }
    if (std::strcmp(name, "host_chunk") == 0) {
        if (value < 32) return RRTMG_LW_ERR_ARG;
        G.host_chunk = value;
        return RRTMG_LW_OK;
    }
    return RRTMG_LW_ERR_ARG;
}

extern "C" long rrtmg_lw_b200_launch_count(int reset) {
    LW_ENTER();
    const long v = G.launches;
    if (reset) G.launches = 0;
    return v;
}

extern "C" int rrtmg_lw_b200_set_timing(int enable) {
    LW_ENTER();
    G.timing = enable != 0;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_last_timing(float* ms_band, float* ms_all) {
    LW_ENTER();
    if (ms_band) *ms_band = G.ms_band;
    if (ms_all) *ms_all = G.ms_all;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_last_band_timing(float* ms16) {
    LW_ENTER();
    if (!ms16) return RRTMG_LW_ERR_ARG;
    for (int k = 0; k < NBND; ++k) ms16[k] = G.ms_bands[k];
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_last_stage_timing(float* ms, int cap) {
    LW_ENTER();
    if (!ms || cap < 0) return -1;
    for (int k = 0; k < G.nstage && k < cap; ++k) ms[k] = G.ms_bands[k];
    return G.nstage;
}

extern "C" const char* rrtmg_lw_b200_strerror(int code) {
    switch (code) {
        case RRTMG_LW_OK: return "ok";
        case RRTMG_LW_ERR_FLAG: return "INVALID ICLD / INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA";
        case RRTMG_LW_ERR_ICE_RADIUS: return "ICE RADIUS OUT OF BOUNDS";
        case RRTMG_LW_ERR_LIQ_RADIUS: return "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS";
        case RRTMG_LW_ERR_KISS_PRESSURE: return "MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS.";
        case RRTMG_LW_ERR_CUDA: return "CUDA runtime error or no CUDA device (no CPU fallback exists)";
        case RRTMG_LW_ERR_NOT_INIT: return "rrtmg_lw_b200_ini has not been called";
        case RRTMG_LW_ERR_ARG: return "invalid argument";
        default: return "unknown error";
    }
}

extern "C" int rrtmg_lw_b200_fp64_peak(double* tflops) {
    LW_ENTER();
    if (!tflops) return RRTMG_LW_ERR_ARG;
    return measure_fp64_peak(tflops, 5) ? RRTMG_LW_ERR_CUDA : RRTMG_LW_OK;
}

extern "C" void rrtmg_lw_b200_kiss_work_stats(unsigned long long* out8) { kiss_work_stats(out8); }
extern "C" int rrtmg_lw_b200_kiss_jump_selftest(void) { return kiss_jump_selftest(); }
extern "C" int rrtmg_lw_b200_mt_jump_selftest(void) { return mt_jump_selftest(); }

extern "C" const char* rrtmg_lw_b200_version(void) { return "rrtmg_lw_b200 0.1 (sm_100a)"; }

extern "C" void rrtmg_lw_b200_finalize(void) {
    LW_ENTER();
    free_scratch();
    free_mt_buffers();
    for (float* p : G.f32_allocs) cudaFree(p);
    G.f32_allocs.clear();
    if (G.stage) cudaFree(G.stage);
    G.stage = nullptr; G.stage_doubles = 0;
    for (double** p : {&G.d_arena, &G.d_arena_p, &G.d_totplnk, &G.d_totplnkderiv, &G.d_exptfn, &G.d_tau, &G.d_absice2, &G.d_absice3,
                       &G.d_absliq1}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    if (G.d_gtab) cudaFree(G.d_gtab);
    G.d_gtab = nullptr;
    if (G.s_in) {
        cudaStreamDestroy(G.s_in); cudaStreamDestroy(G.s_out);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(G.e_in[k]); cudaEventDestroy(G.e_done[k]); cudaEventDestroy(G.e_out[k]); }
        G.s_in = G.s_out = nullptr;
    }
    for (auto& q : G.side) { if (q) cudaStreamDestroy(q); q = nullptr; }
    if (G.hi) { cudaStreamDestroy(G.hi); G.hi = nullptr; }
    for (cudaEvent_t* e : {&G.e_fork, &G.e_colprep, &G.e_t, &G.e_cld, &G.e_kw, &G.e_mop}) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
    for (auto& e : G.e_side) { if (e) cudaEventDestroy(e); e = nullptr; }
    G.inited = false;
}
