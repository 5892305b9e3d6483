// holo_realize.cu -- K3 (realised GWB) and K4 (loudest-source / background split) for sm_100a.
//
// Replaces the reference's sequential loops
//   _sam_poisson_gwb                         holodeck/cyutils.pyx:862-897
//   _loudest_hc_from_sorted                  holodeck/cyutils.pyx:1266-1344
//   _loudest_hc_and_par_from_sorted          holodeck/cyutils.pyx:1409-1538
//   _loudest_hc_and_par_from_sorted_redz     holodeck/cyutils.pyx:1615-1767
//   _ss_bg_hc / _ss_bg_hc_and_par            holodeck/cyutils.pyx:935-1178
//
// Parallel decomposition (see DESIGN.md section "K3/K4"):
//   * one THREAD per realization, one CTA per (cell chunk, group of 4 frequencies, realization tile);
//     all lanes of a warp look at the same (cell,f) element, so the per-element sampler set-up is
//     staged once per CTA in shared memory (non-empty elements only: 83% of the grid has N == 0)
//     and the sampler class branch is warp-uniform;
//   * the (M,q,z) reduction lives in per-thread registers -> per-chunk partial sums in HBM ->
//     a fixed-order final reduction (bit-reproducible; no floating-point atomics);
//   * the loudest-L selection only ever needs the first few *occupied* cells in rank order.  A prep
//     pass finds, per frequency, the rank K_f at which the expected number of occupied cells reaches
//     L + margin; occupied cells with rank < K_f are appended to a small per-(f,r) event bucket and
//     resolved (sorted by rank, slots handed out with multiplicity) by one warp per (f,r); everything
//     else is summed straight into the background.  If a bucket overflows or the head turns out to
//     be too short the call reports HOLO_ERR_OVERFLOW and the host retries with a larger margin.
#include <cuda_runtime.h>

#include <cstdlib>

#include "holo_api.cuh"
#include <string.h>

#include "holo_rng.cuh"

namespace holo {

enum {
    V_GWB = 0,
    V_LOUD_PLAIN = HOLO_LOUDEST_PLAIN,
    V_LOUD_PAR = HOLO_LOUDEST_PAR,
    V_LOUD_PAR_REDZ = HOLO_LOUDEST_PAR_REDZ,
    V_SSBG = 4,
    V_SSBG_PAR = 5,
};

__host__ __device__ constexpr int nacc_of(int variant) {
    return variant == V_LOUD_PAR_REDZ ? 8 : ((variant == V_LOUD_PAR || variant == V_SSBG_PAR) ? 4 : 1);
}
__host__ __device__ constexpr bool has_events(int v) {
    return v == V_LOUD_PLAIN || v == V_LOUD_PAR || v == V_LOUD_PAR_REDZ;
}
__host__ __device__ constexpr bool has_max(int v) { return v == V_SSBG || v == V_SSBG_PAR; }

constexpr int RZ_THREADS = 256;      // threads per CTA (128 when R <= 128); each thread carries RPT realization slots
constexpr int NSCAN = 256;           // cells scanned per pass (max), in rounds of one cell per thread
constexpr int FGROUP = 4;            // frequencies per CTA (4 doubles = one 32 B sector per cell)
// 32-bit CDF thresholds per pass (dynamic shared memory).  A larger pool means fewer, longer passes where nearly every
// element is a table (the per-pass staging / barrier / build latency is paid less often): 40 KB instead of 24 KB takes
// 6 % off the R = 1000 draws.  It is what still lets two 256-thread CTAs share an SM; the plain GWB kernel keeps 24 KB
// because its two-slot form (256 < R <= 512: the eccentric harmonic slabs at R = 500) runs THREE CTAs per SM with it
// and loses 22 % with two.  The pool size fixes the pass boundaries, so it depends on the variant only, never on R.
__host__ __device__ constexpr int pool_entries_of(int variant) {
    return (variant == HOLO_LOUDEST_PLAIN || variant == HOLO_LOUDEST_PAR || variant == HOLO_LOUDEST_PAR_REDZ) ? 10240 : 6144;
}
#ifndef HOLO_GROUP_RESERVE
#define HOLO_GROUP_RESERVE 288
#endif
constexpr int GROUP_RESERVE = HOLO_GROUP_RESERVE;   // head of the pool: CDF table of the pass's superposition group
#ifndef HOLO_GROUP_MAX_LAM
#define HOLO_GROUP_MAX_LAM 0.25
#endif
constexpr double GROUP_MAX_LAM = HOLO_GROUP_MAX_LAM;   // elements below this expectation value are drawn as one Poisson process
constexpr int CLS_GROUP = 6;         // (continues the CLS_* enum of holo_rng.cuh) member of the superposition group
constexpr int STREAM_GWB = 1, STREAM_LOUD = 2, STREAM_SSBG = 3, STREAM_BULK = 4;
static_assert(FGROUP * (TABLE_WMAX + 2) + GROUP_RESERVE <= pool_entries_of(0), "one cell must always fit the table pool");

// Debug instrumentation (never in the product build): -DHOLO_PHASE_CLOCKS makes thread 0 of every CTA add the
// clock64 cycles it spends in each phase of a pass to g_phase_clk (read with holo_debug_phase_clocks).
#ifdef HOLO_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[8];
__device__ unsigned long long g_item_phase[16384][8];   // the same per (chunk, frequency group) item
__device__ unsigned long long g_cta_span[4][16384];   // globaltimer at start / end, SM id and work item of each CTA
__device__ __forceinline__ unsigned long long holo_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned holo_smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
#define HOLO_CTA_ITEM(item)                                                                          \
    if (threadIdx.x == 0) {                                                                          \
        const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);         \
        if (lin < 16384u) g_cta_span[3][lin] = (unsigned long long)(item);                           \
    }
#define HOLO_CTA_SPAN(which)                                                                         \
    if (threadIdx.x == 0) {                                                                          \
        const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);         \
        if (lin < 16384u) {                                                                          \
            g_cta_span[which][lin] = holo_globaltimer();                                             \
            g_cta_span[2][lin] = holo_smid();                                                        \
        }                                                                                            \
    }
#define HOLO_PHASE_DECL long long ph_last = clock64();
#define HOLO_PHASE_MARK(i)                                                          \
    if (threadIdx.x == 0) {                                                         \
        const long long ph_now = clock64();                                         \
        atomicAdd(&g_phase_clk[i], (unsigned long long)(ph_now - ph_last));         \
        const unsigned ph_lin = blockIdx.x + gridDim.x * blockIdx.y;                \
        const unsigned ph_item = a.order ? (unsigned)a.order[ph_lin] : ph_lin;      \
        if (ph_item < 16384u) atomicAdd(&g_item_phase[ph_item][i], (unsigned long long)(ph_now - ph_last)); \
        ph_last = ph_now;                                                           \
    }
#else
#define HOLO_PHASE_DECL
#define HOLO_PHASE_MARK(i)
#define HOLO_CTA_SPAN(which)
#define HOLO_CTA_ITEM(item)
#endif

struct Event {
    int rank;
    int cell;
    double n;
};

struct RealizeArgs {
    const double* number;     // (ncell, F)
    const double* h2fdf;      // (ncell, F)
    const int32_t* rank;      // (ncell,)  rank of each cell (events variants)
    const int32_t* kf;        // (F,)      head length per frequency
    const double* mt;         // (Mb,)
    const double* mr;         // (Qb,)
    const double* rz;         // (Zb,)
    const double* redz_final; // (ncell, F)
    const double* dcom_final;
    const double* sepa;
    const double* angs;
    const double* counts;     // (R, F, ncell) or NULL
    double* partial;          // (nchunk, F, NACC, R)
    double* pmax;             // (nchunk, F, R)        ss_bg variants
    int32_t* pidx;            // (nchunk, F, R)
    Event* events;            // (F, R, cap)
    int32_t* evcount;         // (F, R)
    int64_t ncell;
    int64_t chunk;            // cells per CTA
    int Qb, Zb, F, R, cap;
    int64_t r0;
    uint32_t k0, k1;
    double thresh;
    // Philox counters use GLOBAL column indices so that a grid processed in column slabs (eccentric harmonics)
    // draws independently in every slab: frequency group fg + fg_key0, element index cell*F_key + f + f_key0
    int fg_key0, f_key0, F_key;
    // Fused background slots (holo_loudest only): local realizations [R, R + Rg) draw an independent realised GWB
    // (stream STREAM_GWB, seed k0g/k1g, global index r0g + (r - R)) from the same staged records and tables -- the
    // work of a pass that does not depend on the realization is paid once for both products.
    const int32_t* order;     // (nchunk * nfg,) launch order of the (chunk, frequency group) items, heaviest first; or NULL
    int Rg;
    int64_t r0g;
    uint32_t k0g, k1g;
    double* partial_g;        // (nchunk, F, Rg)
};

// One staged ELEMENT (cell, frequency slot) with a non-zero expectation value.
struct Rec {
    double h;         // h2fdf
    double lam;       // expectation value
    int cell;
    uint32_t meta;    // fi [0:2) | CLS_* [2:5) | head [5] | first-of-cell [6] | floor(log2 W) [8:12) | W [12:24)
    uint32_t kmin;    // TABLE: first count of the tabulated window
    uint32_t toff;    // TABLE: pool index of threshold 0 (sentinels at toff-1 and toff+W)
};
static_assert(sizeof(Rec) == 32, "Rec is read with two 16 B shared-memory loads");
constexpr uint32_t META_HEAD = 1u << 5, META_FIRST = 1u << 6;

// records staged per pass; the parameter variants carry 3 (7) more doubles per record
__host__ __device__ constexpr int nrec_of(int nacc) { return nacc == 1 ? 512 : 256; }

// realizations carried by one thread: the per-CTA work that does not depend on the realization (staging,
// CDF tables) is shared by blockDim * rpt realizations; bounded by the accumulator registers
// The two main variants (realised GWB, loudest split without parameters) pick the number of lock-step slots from
// the realization count: 4 slots (one CTA then carries 1024 realizations: a single staging pass at R = 1000),
// 2 or 1 for small R so that no slot is dead.  Results do not depend on the choice (draws are keyed on the
// global realization index, pass boundaries on the cells only).
__host__ __device__ constexpr bool flexible_rpt(int variant) { return variant == V_GWB || variant == V_LOUD_PLAIN; }
__host__ __device__ constexpr int rpt_of(int variant) { return variant == V_SSBG ? 2 : 1; }       // fixed-slot variants
inline int rpt_for(int variant, int R) {
    if (!flexible_rpt(variant)) return rpt_of(variant);
    return R > 2 * RZ_THREADS ? 4 : (R > RZ_THREADS ? 2 : 1);
}
__host__ __device__ constexpr int min_ctas_of(int variant, int rpt, int threads) {
    return threads <= 128 ? 4 : (flexible_rpt(variant) ? (rpt >= 4 ? 2 : 3) : 2);
}
// Small realization counts run 128-thread CTAs (four per SM instead of two): the work per pass that does not depend
// on R is a chain of dependent latencies, so more resident CTAs is what hides it.  The staging scan then takes the
// 256-cell window in two rounds; pass boundaries, records and tables are the same as with 256 threads.
inline int threads_for(int rpt, int R) { return (rpt == 1 && R <= 128) ? 128 : RZ_THREADS; }

// The FGROUP consecutive frequencies of one cell are one aligned 32 B sector when F % 4 == 0: read them
// with two 16 B loads instead of four strided 8 B loads.
__device__ __forceinline__ void load_group(const double* __restrict__ base, int nf, bool vec, double (&out)[FGROUP]) {
    if (vec) {
        const double2 lo = __ldg(reinterpret_cast<const double2*>(base));
        const double2 hi = __ldg(reinterpret_cast<const double2*>(base) + 1);
        out[0] = lo.x; out[1] = lo.y; out[2] = hi.x; out[3] = hi.y;
    } else {
#pragma unroll
        for (int fi = 0; fi < FGROUP; ++fi) out[fi] = (fi < nf) ? base[fi] : 0.0;
    }
}

// An occupied head cell of the loudest variants: append (rank, cell, n) to the (f, r) event bucket.  Out of
// line: rare (a few dozen per (f, r)) and it keeps the hot loops small.
static __device__ __noinline__ void push_event(const RealizeArgs& a, int f, int r, int cell, double n) {
#ifdef HOLO_NO_PUSH
    return;      // (profiling experiment only: results are wrong)
#endif
    const int slot = atomicAdd(&a.evcount[(int64_t)f * a.R + r], 1);
    if (slot < a.cap) {
        Event ev;
        ev.rank = a.rank[cell];
        ev.cell = cell;
        ev.n = n;
        a.events[((int64_t)f * a.R + r) * a.cap + slot] = ev;
    }
}

// Fold one draw `n` of element `rec` (frequency slot FI) into the thread's accumulators or, for occupied
// head cells of the loudest variants, into the event bucket.
template <int VARIANT, int FI>
__device__ __forceinline__ void fold_static(const RealizeArgs& a, const Rec& rec, const double* w3, const double* w4,
                                            int f0, int r, double n, double (&acc)[FGROUP][nacc_of(VARIANT)],
                                            double (&vmax)[FGROUP], int (&imax)[FGROUP], bool gslot) {
    constexpr int NACC = nacc_of(VARIANT);
    const double cur = rec.h;
    if (VARIANT == V_GWB) {
        acc[FI][0] += n * cur;                                              // pyx:891, 895
    } else if (has_max(VARIANT)) {
        // `if (cur > max and num > 0)` walking cells in natural order (pyx:993, 1134): the first cell
        // holding the maximum wins.  Cells are not visited in natural order here, hence the index tie-break.
        if (n > 0.0 && (cur > vmax[FI] || (cur == vmax[FI] && cur > 0.0 && rec.cell < imax[FI]))) {
            vmax[FI] = cur;
            imax[FI] = rec.cell;
        }
        const double nc = n * cur;
        acc[FI][0] += nc;                                                   // pyx:998, 1139
        if (NACC > 1) {
#pragma unroll
            for (int k = 1; k < 4; ++k) acc[FI][k < NACC ? k : 0] += nc * w3[k - 1];
        }
    } else {
        // `if num < 1: continue` (pyx:1333, 1490, 1727): counts are integers (or > 1e10), so an empty draw adds zero
        if ((rec.meta & META_HEAD) && !gslot) {
            if (n >= 1.0) push_event(a, f0 + FI, r, rec.cell, n);
        } else {
            const double nc = n * cur;
            acc[FI][0] += nc;                                               // pyx:1342, 1505, 1745
            if (NACC > 1) {
#pragma unroll
                for (int k = 1; k < 4; ++k) acc[FI][k < NACC ? k : 0] += nc * w3[k - 1];
            }
            if (NACC > 4) {
#pragma unroll
                for (int k = 4; k < 8; ++k) acc[FI][k < NACC ? k : 0] += nc * w4[k - 4];
            }
        }
    }
}

template <int VARIANT>
__device__ __forceinline__ void fold_rec(const RealizeArgs& a, const Rec& rec, const double* w3, const double* w4,
                                         int f0, int r, double n, double (&acc)[FGROUP][nacc_of(VARIANT)],
                                         double (&vmax)[FGROUP], int (&imax)[FGROUP], bool gslot = false) {
    switch (rec.meta & 3u) {
        case 0: fold_static<VARIANT, 0>(a, rec, w3, w4, f0, r, n, acc, vmax, imax, gslot); break;
        case 1: fold_static<VARIANT, 1>(a, rec, w3, w4, f0, r, n, acc, vmax, imax, gslot); break;
        case 2: fold_static<VARIANT, 2>(a, rec, w3, w4, f0, r, n, acc, vmax, imax, gslot); break;
        default: fold_static<VARIANT, 3>(a, rec, w3, w4, f0, r, n, acc, vmax, imax, gslot); break;
    }
}

// Accumulators in shared memory (variants with one sum per frequency and no running maximum): the frequency slot
// is a run-time index, so the fold is load / fma / store on a thread-private column -- no four-way dispatch, and
// sixteen registers fewer in the draw loops.  `col` = &s_acc[0][u][tid]; consecutive frequency slots are
// SACC_STRIDE doubles apart.
template <int VARIANT>
__device__ __forceinline__ void fold_sacc(const RealizeArgs& a, const Rec& rec, int f0, int r, double n, double* col,
                                          int stride, bool gslot = false) {
    const int fi = (int)(rec.meta & 3u);
    if (has_events(VARIANT) && (rec.meta & META_HEAD) && !gslot) {
        if (n >= 1.0) push_event(a, f0 + fi, r, rec.cell, n);           // pyx:1333-1341
    } else {
        col[fi * stride] += n * rec.h;                                  // pyx:891-895, 1342
    }
}

// SW cooperating lanes tabulate the CDF of Poisson(lam) over its window as 32-bit thresholds (see holo_rng.cuh);
// t[-1] = 0 and t[W] = 2^32-1 are sentinels for the ambiguity test of the draw.  The 32/SW sub-warps of a warp build
// DIFFERENT tables at the same time (`sl` = lane within the sub-warp; every lane of the warp must call, inactive
// sub-warps with active = false): most tables have fewer than 64 entries, so the cost of a table is the one pmf
// evaluation per lane (log, exp, Stirling) and the scan, not the entries -- narrow sub-warps cut that latency 4x.
template <int SW>
static __device__ __noinline__ void build_table_sub(double lam, uint32_t* t, int kmin, int W, int sl, bool active) {
    const int seg = (W + SW - 1) / SW;
    int j0 = sl * seg;
    if (j0 > W) j0 = W;
    int j1 = j0 + seg;
    if (j1 > W) j1 = W;
    if (!active) {
        j0 = j1 = 0;
        lam = 1.0;
    }
    const double ln_lam = log(lam), inv_lam = 1.0 / lam;
    double ptop;
    double incl = table_segment_mass(lam, ln_lam, inv_lam, kmin, j0, j1, &ptop);
#pragma unroll
    for (int off = 1; off < SW; off <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, off, SW);
        if (sl >= off) incl += v;
    }
    table_segment_write(t, incl, ptop, inv_lam, kmin, j0, j1);
    if (active && sl == 0) {
        t[-1] = 0u;
        t[W] = 0xFFFFFFFFu;
    }
}
constexpr int BUILD_SW = 8;      // lanes per element table

// Stage the next run of cells of a pass (see the kernel): thread <-> cell, elements compacted in (cell, frequency)
// order into `s_rec` (main records from 0 up, group members from NREC-1 down).  Returns the number of cells
// consumed; *s_tot / *s_totlam receive the packed record counts and the group's total expectation value.
// The window is always NSCAN cells and the record / pool limits are fixed, so the pass boundaries -- and with
// them the membership of the superposition groups -- never depend on how the realizations are tiled or on the
// CTA size (the scans carry over from one round of THREADS cells to the next in cell order).
template <int VARIANT, int THREADS, int NSCAN_T = NSCAN>
static __device__ __forceinline__ int stage_pass(const RealizeArgs& a, int64_t cb, int64_t c_hi, int f0, int nf, Rec* s_rec,
                                              double* s_w3, double* s_w4, double* s_gcum, unsigned long long* s_wsum,
                                              double* s_wlam, unsigned long long* s_tot, double* s_totlam
#ifdef HOLO_PHASE_CLOCKS
                                              , long long& ph_last
#endif
                                              ) {
    constexpr int NACC = nacc_of(VARIANT);
    constexpr int NREC = nrec_of(NACC);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = THREADS / 32;
    const bool supplied = a.counts != nullptr;
    const bool vec4 = (nf == FGROUP) && ((a.F & 3) == 0);   // 32 B-aligned frequency groups
    __syncthreads();   // previous pass fully consumed
    HOLO_PHASE_MARK(4)   // waiting for the slowest thread of the previous pass
    int ncons = 0;
    unsigned long long carry = 0;
    double carry_lam = 0.0;
    for (int sbase = 0; sbase < NSCAN_T; sbase += THREADS) {    // rounds of THREADS cells
        const int64_t c = cb + sbase + tid;
        const bool inrange = (c < c_hi);
        unsigned clsw = 0;            // CLS_* byte per frequency slot
        unsigned nmain_t = 0, ngrp_t = 0, need_t = 0;
        double glam_t = 0.0;
        double lam4[FGROUP], h4[FGROUP];
        if (inrange) {
            load_group(a.number + c * a.F + f0, nf, vec4, lam4);
            load_group(a.h2fdf + c * a.F + f0, nf, vec4, h4);
#pragma unroll
            for (int fi = 0; fi < FGROUP; ++fi) {
                if (fi >= nf) continue;
                int cls = classify_draw(lam4[fi], a.thresh);
                if (VARIANT == V_LOUD_PAR_REDZ && h4[fi] == 0.0) cls = CLS_EMPTY;   // pyx:1727
                if (cls == CLS_EMPTY) continue;
                if (supplied) {
                    cls = CLS_SMALL;                                         // every count is read from `counts`
                } else if (cls != CLS_NORMAL) {
                    if (lam4[fi] < GROUP_MAX_LAM) cls = CLS_GROUP;
                    else if (lam4[fi] <= TABLE_MAX_LAM) cls = CLS_TABLE;
                    else cls = CLS_PTRS;
                }
                if (cls == CLS_GROUP) {
                    ++ngrp_t;
                    glam_t += lam4[fi];
                } else {
                    ++nmain_t;
                    if (cls == CLS_TABLE) need_t += (unsigned)table_spec(lam4[fi]).W + 2u;
                }
                clsw |= (unsigned)cls << (8 * fi);
            }
        }
        // block-wide inclusive scans in cell order: (main | group << 11 | pool entries << 22), group's lambda
        unsigned long long incl = (unsigned long long)nmain_t | ((unsigned long long)ngrp_t << 11) |
                                  ((unsigned long long)need_t << 22);
        double glam = glam_t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
            const double w = __shfl_up_sync(0xffffffffu, glam, off);
            if (lane >= off) { incl += v; glam += w; }
        }
        if (lane == 31) { s_wsum[warp] = incl; s_wlam[warp] = glam; }
        __syncthreads();
        unsigned long long round_tot = carry;
        double round_lam = carry_lam;
        {
            unsigned long long pre = carry;
            double prel = carry_lam;
            for (int w = 0; w < nwarp; ++w) {
                if (w == warp) { incl += pre; glam = prel + glam; }
                pre += s_wsum[w];
                prel += s_wlam[w];
            }
            round_tot = pre;
            round_lam = prel;
        }
        const int main_incl = (int)(incl & 2047u), grp_incl = (int)((incl >> 11) & 2047u);
        const int need_incl = (int)(incl >> 22);
        const bool taken = inrange && (main_incl + grp_incl <= NREC) &&
                           (need_incl <= pool_entries_of(VARIANT) - GROUP_RESERVE);        // a prefix of the run
        const int ntaken = __syncthreads_count(taken);
        if (ntaken > 0 && tid == ntaken - 1) { *s_tot = incl; *s_totlam = glam; }
        if (taken && clsw != 0) {
            int mi = main_incl - (int)nmain_t;                   // next main record
            int gi = grp_incl - (int)ngrp_t;                     // next group member
            double gc = glam - glam_t;                           // cumulative lambda before this cell
            uint32_t toff = (uint32_t)(GROUP_RESERVE + need_incl - (int)need_t);
            int rk = 0;
            if (has_events(VARIANT)) rk = a.rank[c];
            bool first = true;
#pragma unroll
            for (int fi = 0; fi < FGROUP; ++fi) {
                const int cls = (clsw >> (8 * fi)) & 0xff;
                if (cls == CLS_EMPTY) continue;
                Rec rec;
                rec.h = h4[fi];
                rec.lam = lam4[fi];
                rec.cell = (int)c;
                rec.meta = (uint32_t)fi | ((uint32_t)cls << 2);
                rec.kmin = 0;
                rec.toff = 0;
                if (has_events(VARIANT) && rk < a.kf[f0 + fi]) rec.meta |= META_HEAD;
                int slot;
                if (cls == CLS_GROUP) {
                    gc += lam4[fi];
                    s_gcum[gi] = gc;
                    slot = NREC - 1 - gi;
                    ++gi;
                } else {
                    if (cls == CLS_TABLE) {
                        const TableSpec ts = table_spec(lam4[fi]);
                        rec.kmin = (uint32_t)ts.kmin;
                        rec.toff = toff + 1u;
                        rec.meta |= ((uint32_t)(31 - __clz(ts.W)) << 8) | ((uint32_t)ts.W << 12);
                        toff += (uint32_t)ts.W + 2u;
                    }
                    if (first && cls != CLS_PTRS) {
                        rec.meta |= META_FIRST;
                        first = false;
                    }
                    slot = mi;
                    ++mi;
                }
                s_rec[slot] = rec;
                if (NACC > 4) {
                    const int64_t o = c * a.F + f0 + fi;
                    s_w4[slot * 4 + 0] = a.redz_final[o];
                    s_w4[slot * 4 + 1] = a.dcom_final[o];
                    s_w4[slot * 4 + 2] = a.sepa[o];
                    s_w4[slot * 4 + 3] = a.angs[o];
                }
                if (NACC > 1) {
                    const int zz = (int)(c % a.Zb);
                    const int64_t mq = c / a.Zb;
                    s_w3[slot * 3 + 0] = a.mt[(int)(mq / a.Qb)];
                    s_w3[slot * 3 + 1] = a.mr[(int)(mq % a.Qb)];
                    s_w3[slot * 3 + 2] = a.rz[zz];
                }
            }
        }
        ncons += ntaken;
        carry = round_tot;
        carry_lam = round_lam;
        if (ntaken < THREADS) break;
    }
    return ncons;
}

template <int VARIANT, int RPT, int THREADS, bool FUSED>
__global__ void __launch_bounds__(THREADS, min_ctas_of(VARIANT, RPT, THREADS))
realize_kernel(RealizeArgs a) {
    static_assert(!FUSED || has_events(VARIANT), "fused background slots exist in the loudest variants only");
    constexpr int NACC = nacc_of(VARIANT);
    constexpr int NREC = nrec_of(NACC);
    __shared__ __align__(16) Rec s_rec[NREC];           // main records grow from 0, group records from NREC-1 down
    __shared__ double s_w3[NACC > 1 ? NREC : 1][3];     // mt, mr, rz of the record's cell (parameter variants)
    __shared__ double s_w4[NACC > 4 ? NREC : 1][4];     // redz_final, dcom, sepa, angs of the element
    __shared__ double s_gcum[NREC];                     // inclusive cumulative expectation of the group members
    __shared__ unsigned short s_plist[NREC];            // main records of class PTRS
    __shared__ unsigned long long s_wsum[RZ_THREADS / 32];
    __shared__ double s_wlam[RZ_THREADS / 32];
    __shared__ unsigned long long s_tot;
    __shared__ double s_totlam;
    __shared__ int s_np;
    __shared__ int s_gspec[3];                          // kmin, W, lg of the group's table
    constexpr int POOL_ENTRIES = pool_entries_of(VARIANT);
    extern __shared__ __align__(16) uint32_t s_pool[];  // POOL_ENTRIES thresholds, then s_acc / s_words (below)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarp = blockDim.x >> 5;
    // CTAs start in launch-index order (x fastest).  The work of an item -- a (cell chunk, frequency group) pair --
    // ranges over three orders of magnitude (low masses at the lowest frequencies hold nearly all the tables), so the
    // items are handed out heaviest first (`order`, see cost_kernel): without it a handful of 9 ms CTAs that start
    // late leave most SMs idle for the last quarter of the kernel.
    int fg = blockIdx.x, chunk_id = blockIdx.y;
    if (a.order != nullptr) {
        const int item = a.order[blockIdx.y * gridDim.x + blockIdx.x];
        fg = item % (int)gridDim.x;
        chunk_id = item / (int)gridDim.x;
    }
    const int f0 = fg * FGROUP;
    const uint32_t fgk = (uint32_t)(fg + a.fg_key0);   // global frequency-group index (Philox counter)
    const int r_first = blockIdx.z * (blockDim.x * RPT) + tid;      // local realization of slot t = 0
    const int64_t c_lo = (int64_t)chunk_id * a.chunk;
    int64_t c_hi = c_lo + a.chunk;
    if (c_hi > a.ncell) c_hi = a.ncell;
    const int nf = (a.F - f0) < FGROUP ? (a.F - f0) : FGROUP;
    const bool supplied = a.counts != nullptr;
    const bool vec4 = (nf == FGROUP) && ((a.F & 3) == 0);   // 32 B-aligned frequency groups

    constexpr bool SACC = (NACC == 1) && !has_max(VARIANT);       // accumulators in shared memory (see fold_sacc)
    constexpr int SACC_STRIDE = RPT * THREADS;
    // dynamic shared memory: [pool | accumulators (SACC) | each thread's Philox block of the current cell, per slot]
    typedef double AccArr[SACC ? RPT : 1][SACC ? THREADS : 1];
    typedef uint32_t WordArr[RPT][THREADS];
    AccArr* s_acc = reinterpret_cast<AccArr*>(s_pool + POOL_ENTRIES);                       // [FGROUP][RPT][THREADS]
    WordArr* s_words = reinterpret_cast<WordArr*>(reinterpret_cast<double*>(s_pool + POOL_ENTRIES) +
                                                  (SACC ? FGROUP * RPT * THREADS : 0));      // [4][RPT][THREADS]
    if (SACC) {
#pragma unroll
        for (int fi = 0; fi < FGROUP; ++fi)
#pragma unroll
            for (int t = 0; t < RPT; ++t) s_acc[SACC ? fi : 0][SACC ? t : 0][SACC ? tid : 0] = 0.0;
    }
    double acc[RPT][FGROUP][NACC];
    double vmax[RPT][FGROUP];
    int imax[RPT][FGROUP];
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
#pragma unroll
        for (int fi = 0; fi < FGROUP; ++fi) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[t][fi][k] = 0.0;
            vmax[t][fi] = 0.0;
            imax[t][fi] = -1;
        }
    }

    DrawKey key;
    key.k0 = a.k0; key.k1 = a.k1;
    key.real = 0;
    key.stream = has_events(VARIANT) ? STREAM_LOUD : (has_max(VARIANT) ? STREAM_SSBG : STREAM_GWB);
    const uint32_t real_first = (uint32_t)(a.r0 + r_first);
    const int Rtot = a.R + (FUSED ? a.Rg : 0);       // live local realizations (incl. fused GWB slots)
    // slot u of this thread: is it a fused background slot, and the Philox key of its realization
    auto is_gslot = [&](int u) -> bool { return FUSED && (r_first + u * THREADS >= a.R); };
    auto set_key = [&](int u) {
        if (is_gslot(u)) {
            key.k0 = a.k0g; key.k1 = a.k1g; key.stream = STREAM_GWB;
            key.real = (uint32_t)(a.r0g + (r_first + u * THREADS - a.R));
        } else {
            if (FUSED) {
                key.k0 = a.k0; key.k1 = a.k1;
                key.stream = has_events(VARIANT) ? STREAM_LOUD : (has_max(VARIANT) ? STREAM_SSBG : STREAM_GWB);
            }
            key.real = real_first + (uint32_t)u * THREADS;
        }
    };

    int64_t cb = c_lo;
    HOLO_CTA_SPAN(0)
    HOLO_CTA_ITEM(chunk_id * (int)gridDim.x + fg)
    HOLO_PHASE_DECL
    while (cb < c_hi) {
        // ---- stage the next run of cells: thread <-> cell, elements compacted in (cell, frequency) order.  The run
        //      ends where the record buffer (NREC elements) or the table pool is full.
        const int ncons = stage_pass<VARIANT, THREADS>(a, cb, c_hi, f0, nf, s_rec, &s_w3[0][0], &s_w4[0][0], s_gcum, s_wsum, s_wlam, &s_tot,
                                              &s_totlam
#ifdef HOLO_PHASE_CLOCKS
                                              , ph_last
#endif
                                              );
        const uint32_t pass_id = (uint32_t)cb;        // first cell of the run: names the pass in the Philox counter
        cb += ncons;
        __syncthreads();
        HOLO_PHASE_MARK(0)   // staging
        const int nmain = (int)(s_tot & 2047u), ngrp = (int)((s_tot >> 11) & 2047u);
        const double glam_tot = s_totlam;
        // ---- per pass set-up shared by all realizations: CDF tables (one warp per table) and the PTRS list
        if (!supplied) {
            constexpr int NSUB = 32 / BUILD_SW;
            const int sub = lane / BUILD_SW, sl = lane % BUILD_SW;
            for (int base = warp * NSUB; base < nmain; base += nwarp * NSUB) {      // (warp-uniform trip count)
                const int i = base + sub;
                const bool act = (i < nmain) && (((s_rec[i < nmain ? i : 0].meta >> 2) & 7u) == CLS_TABLE);
                const Rec& rec = s_rec[act ? i : 0];
                build_table_sub<BUILD_SW>(rec.lam, s_pool + rec.toff, (int)rec.kmin, (int)((rec.meta >> 12) & 4095u), sl, act);
            }
            if (warp == nwarp - 1 && ngrp > 0) {
                const TableSpec ts = table_spec(glam_tot);
                build_table_sub<32>(glam_tot, s_pool + 1, ts.kmin, ts.W, lane, true);
                if (lane == 0) { s_gspec[0] = ts.kmin; s_gspec[1] = ts.W; s_gspec[2] = 31 - __clz(ts.W); }
            }
            if (warp == 0) {
                int np = 0;
                for (int base = 0; base < nmain; base += 32) {
                    const int i = base + lane;
                    const bool is = (i < nmain) && (((s_rec[i].meta >> 2) & 7u) == CLS_PTRS);
                    const unsigned bal = __ballot_sync(0xffffffffu, is);
                    if (is) {
                        const int j = np + __popc(bal & ((1u << lane) - 1u));
                        s_plist[j] = (unsigned short)i;
                        // set-up of the transformed rejection shared by all realizations (prep_draw's a0, a1): b goes
                        // to the free tail of s_gcum (group members fill it from the front, and ngrp + nmain <= NREC),
                        // vr into the record's unused table fields
                        FPrep pp;
                        prep_draw(s_rec[i].lam, a.thresh, pp);
                        s_gcum[NREC - 1 - j] = pp.a0;
                        *reinterpret_cast<double*>(&s_rec[i].kmin) = pp.a1;
                    }
                    np += __popc(bal);
                }
                if (lane == 0) s_np = np;
            }
        }
        __syncthreads();
        HOLO_PHASE_MARK(1)   // table builds

        // The RPT realization slots of a thread advance in lock-step through the records: they share the record
        // decode and the rung dispatch, and give the scheduler RPT independent dependency chains.
        if (supplied) {
            // ---- supplied-count mode: no random numbers at all
            for (int i = 0; i < nmain; ++i) {
                const Rec rec = s_rec[i];
                const int f = f0 + (int)(rec.meta & 3u);
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    const int r = r_first + u * (int)blockDim.x;
                    if (r >= a.R) continue;
                    const double n = a.counts[((int64_t)r * a.F + f) * a.ncell + rec.cell];
                    if (SACC) fold_sacc<VARIANT>(a, rec, f0, r, n, &s_acc[0][SACC ? u : 0][SACC ? tid : 0], SACC_STRIDE);
                    else fold_rec<VARIANT>(a, rec, s_w3[NACC > 1 ? i : 0], s_w4[NACC > 4 ? i : 0], f0, r, n, acc[u], vmax[u], imax[u]);
                }
            }
            continue;
        }
        if (r_first >= Rtot) continue;     // no live slot in this thread (the barriers are at the loop top)

        // ---- phase A, lock-step: every thread walks the main records; the (up to four) draws of a cell consume
        //      the words of one Philox block per slot, in record order
        {
            int ord = 0;
            for (int i = 0; i < nmain; ++i) {
                const Rec rec = s_rec[i];
                const uint32_t meta = rec.meta;
                const uint32_t cls = (meta >> 2) & 7u;
                if (meta & META_FIRST) {
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        set_key(u);
                        const Philox4 hi = group_bits(key, (uint32_t)rec.cell, fgk, PURPOSE_GROUP_HI);
                        s_words[0][u][tid] = hi.v[0]; s_words[1][u][tid] = hi.v[1];     // (thread-private columns:
                        s_words[2][u][tid] = hi.v[2]; s_words[3][u][tid] = hi.v[3];     //  no synchronisation needed)
                    }
                    ord = 0;
                }
                if (cls == CLS_PTRS) continue;
                double n[RPT];
                if (cls == CLS_TABLE) {
                    uint32_t word[RPT], q[RPT];
#pragma unroll
                    for (int u = 0; u < RPT; ++u) word[u] = s_words[ord][u][tid];     // next word of the cell's block
                    const int W = (int)((meta >> 12) & 4095u), lg = (int)((meta >> 8) & 15u);
                    table_ladder_n<RPT>(s_pool, rec.toff, W, lg, word, q);
                    // The word alone decides all but ~W 2^-31 of the draws.  Plain GWB variant: the slots' tests are
                    // evaluated together (independent loads) behind ONE branch -- 10.06 -> 9.61 ms at R = 1000.  The
                    // loudest variants sit at the 128-register cap and are 0.5 % faster with the tests taken slot by slot.
                    if constexpr (VARIANT == V_GWB) {
                        bool amb[RPT], amb_any = false;
#pragma unroll
                        for (int u = 0; u < RPT; ++u) {
                            n[u] = (double)(int)(rec.kmin + (q[u] - rec.toff));
                            amb[u] = table_ambiguous(s_pool, rec.toff, W, q[u], word[u]);
                            amb_any |= amb[u];
                        }
                        if (amb_any) {
#pragma unroll
                            for (int u = 0; u < RPT; ++u) {
                                if (!amb[u]) continue;
                                set_key(u);
                                n[u] = table_resolve_keyed(rec.lam, s_pool + rec.toff, (int)rec.kmin, W, (int)(q[u] - rec.toff),
                                                           word[u], key, (uint32_t)rec.cell, fgk, ord);
                            }
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < RPT; ++u) {
                            n[u] = (double)(int)(rec.kmin + (q[u] - rec.toff));
                            if (table_ambiguous(s_pool, rec.toff, W, q[u], word[u])) {
                                set_key(u);
                                n[u] = table_resolve_keyed(rec.lam, s_pool + rec.toff, (int)rec.kmin, W, (int)(q[u] - rec.toff),
                                                           word[u], key, (uint32_t)rec.cell, fgk, ord);
                            }
                        }
                    }
                    ++ord;
                } else {
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        set_key(u);
                        n[u] = draw_normal_lam(rec.lam, key, (uint64_t)(uint32_t)rec.cell * (uint64_t)a.F_key +
                                                             (uint64_t)(a.f_key0 + f0 + (int)(meta & 3u)));
                    }
                }
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    const int r = r_first + u * (int)blockDim.x;
                    if (r >= Rtot) continue;
                    if (SACC) fold_sacc<VARIANT>(a, rec, f0, r, n[u], &s_acc[0][SACC ? u : 0][SACC ? tid : 0], SACC_STRIDE, is_gslot(u));
                    else fold_rec<VARIANT>(a, rec, s_w3[NACC > 1 ? i : 0], s_w4[NACC > 4 ? i : 0], f0, r, n[u], acc[u], vmax[u], imax[u], is_gslot(u));
                }
            }
        }

        HOLO_PHASE_MARK(2)   // phase A (TABLE / NORMAL draws)
        // The two divergent stages take the slots one after the other in a run-time loop (one copy of the code);
        // their sums go through `tacc` and are merged into the slot's accumulators with static register indices.
        const int np = s_np;
        if constexpr (SACC && RPT > 1) {
            // ---- All RPT slots of a thread advance through the two divergent stages in LOCK-STEP, as in phase A: the
            //      Philox blocks, the searches and the proposals of the slots are independent dependency chains the
            //      scheduler can interleave (taken one slot after the other these stages ran at 39 % issue-slot use and
            //      held 23 % of a CTA's cycles).  Every slot draws exactly the random numbers it drew before.
            bool live[RPT];
            DrawKey kk[RPT];
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                live[u] = (r_first + u * THREADS) < Rtot;
                set_key(u);
                kk[u] = key;
            }
            if (ngrp > 0 && live[0]) {
                // the superposition group: one Poisson process of rate Lambda = sum lam_k per slot; each of its events
                // belongs to member k with probability lam_k / Lambda (see draw_group in holo_rng.cuh)
                const int gk = s_gspec[0], gW = s_gspec[1], glg = s_gspec[2];
                uint32_t w0[RPT], w1[RPT], qq[RPT];
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    const Philox4 gb = philox4x32_10(pass_id, fgk | ((uint32_t)PURPOSE_PASS_COUNT << 28), kk[u].real,
                                                     kk[u].stream << 24, kk[u].k0, kk[u].k1);
                    w0[u] = gb.v[0];
                    w1[u] = gb.v[1];
                }
                table_ladder_n<RPT>(s_pool, 1u, gW, glg, w0, qq);
                int nev[RPT], nmax = 0;
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    double ne = (double)(gk + (int)(qq[u] - 1u));
                    if (table_ambiguous(s_pool, 1u, gW, qq[u], w0[u]))
                        ne = table_resolve(glam_tot, s_pool + 1, gk, gW, (int)(qq[u] - 1u), w0[u], w1[u]);
                    nev[u] = live[u] ? (int)ne : 0;
                    nmax = nev[u] > nmax ? nev[u] : nmax;
                }
                // events of HEAD members are parked (six 10-bit record indices per 64-bit word) and appended to the
                // buckets outside the event loop (see the single-slot form below)
                unsigned long long pend[RPT];
                int npend[RPT];
#pragma unroll
                for (int u = 0; u < RPT; ++u) { pend[u] = 0ull; npend[u] = 0; }
                uint32_t odd2[RPT], odd3[RPT];          // words 2, 3 of the current pick block: the odd event's uniform
                for (int ev = 0; ev < nmax; ++ev) {
                    double v[RPT];
                    if ((ev & 1) == 0) {
#pragma unroll
                        for (int u = 0; u < RPT; ++u) {
                            const Philox4 pb = philox4x32_10(pass_id, fgk | ((uint32_t)PURPOSE_PASS_PICK << 28), kk[u].real,
                                                             (kk[u].stream << 24) | (uint32_t)(ev >> 1), kk[u].k0, kk[u].k1);
                            v[u] = u53(pb.v[0], pb.v[1]) * glam_tot;
                            odd2[u] = pb.v[2];
                            odd3[u] = pb.v[3];
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < RPT; ++u) v[u] = u53(odd2[u], odd3[u]) * glam_tot;
                    }
                    int base[RPT];
#pragma unroll
                    for (int u = 0; u < RPT; ++u) base[u] = 0;
                    int len = ngrp;                     // member index = #{j : gcum[j] <= v}: one loop for all slots
                    while (len > 1) {
                        const int half = len >> 1;
#pragma unroll
                        for (int u = 0; u < RPT; ++u)
                            if (s_gcum[base[u] + half - 1] <= v[u]) base[u] += half;
                        len -= half;
                    }
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        if (s_gcum[base[u]] <= v[u] && base[u] < ngrp - 1) base[u] += 1;
                        if (ev < nev[u]) {
                            const int slot = NREC - 1 - base[u];
                            const Rec rec = s_rec[slot];
                            const bool gslot = is_gslot(u);
                            if (has_events(VARIANT) && (rec.meta & META_HEAD) && !gslot) {
                                if (npend[u] == 6) {
                                    const int r = r_first + u * THREADS;
                                    while (npend[u] > 0) {
                                        const Rec pr = s_rec[(int)(pend[u] & 1023ull)];
                                        push_event(a, f0 + (int)(pr.meta & 3u), r, pr.cell, 1.0);
                                        pend[u] >>= 10;
                                        --npend[u];
                                    }
                                }
                                pend[u] = (pend[u] << 10) | (unsigned long long)slot;
                                ++npend[u];
                            } else {
                                fold_sacc<VARIANT>(a, rec, f0, r_first + u * THREADS, 1.0, &s_acc[0][SACC ? u : 0][SACC ? tid : 0],
                                                   SACC_STRIDE, gslot);
                            }
                        }
                    }
                }
                if (has_events(VARIANT)) {
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        const int r = r_first + u * THREADS;
                        while (npend[u] > 0) {
                            const Rec pr = s_rec[(int)(pend[u] & 1023ull)];
                            push_event(a, f0 + (int)(pr.meta & 3u), r, pr.cell, 1.0);
                            pend[u] >>= 10;
                            --npend[u];
                        }
                    }
                }
            }
            // ---- PTRS list: every slot walks it at its own pace (a rejected proposal delays only its own slot of its
            //      own lane), but the proposals of the RPT slots -- a Philox block and twenty flops each -- are
            //      computed side by side; only the rare undecided ones take the out-of-line exact test
            if (np > 0 && live[0]) {
                int it[RPT];
                uint32_t trial[RPT];
#pragma unroll
                for (int u = 0; u < RPT; ++u) { it[u] = live[u] ? 0 : np; trial[u] = 0u; }
                for (;;) {
                    bool any = false;
#pragma unroll
                    for (int u = 0; u < RPT; ++u) any |= (it[u] < np);
                    if (!any) break;
                    double kq[RPT], usq[RPT], Vq[RPT];
                    int dec[RPT], ri[RPT];
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        const int j = it[u] < np ? it[u] : np - 1;          // (finished slots compute a discarded proposal)
                        ri[u] = s_plist[j];
                        const double lam = s_rec[ri[u]].lam;
                        const uint32_t meta = s_rec[ri[u]].meta;
                        const double b = s_gcum[NREC - 1 - j];
                        const double vr = *reinterpret_cast<const double*>(&s_rec[ri[u]].kmin);
                        const uint64_t idx = (uint64_t)(uint32_t)s_rec[ri[u]].cell * (uint64_t)a.F_key +
                                             (uint64_t)(a.f_key0 + f0 + (int)(meta & 3u));
                        dec[u] = ptrs_propose(lam, b, vr, element_bits(kk[u], idx, trial[u]), &kq[u], &usq[u], &Vq[u]);
                    }
#pragma unroll
                    for (int u = 0; u < RPT; ++u) {
                        if (it[u] >= np) continue;
                        bool ok = dec[u] > 0;
                        if (dec[u] == 0) {
                            const int j = it[u];
                            ok = ptrs_decide(s_rec[ri[u]].lam, s_gcum[NREC - 1 - j], usq[u], Vq[u], kq[u]);
                        }
                        if (ok) {
                            fold_sacc<VARIANT>(a, s_rec[ri[u]], f0, r_first + u * THREADS, kq[u], &s_acc[0][SACC ? u : 0][SACC ? tid : 0],
                                               SACC_STRIDE, is_gslot(u));
                            ++it[u];
                            trial[u] = 0u;
                        } else {
                            ++trial[u];
                        }
                    }
                }
            }
        } else if (ngrp > 0 || np > 0) {
#pragma unroll 1
            for (int u = 0; u < RPT; ++u) {
                const int r = r_first + u * (int)blockDim.x;
                if (r >= Rtot) break;
                set_key(u);
                const bool gslot = is_gslot(u);
                double tacc[FGROUP][NACC];
                double tvmax[FGROUP];
                int timax[FGROUP];
#pragma unroll
                for (int fi = 0; fi < FGROUP; ++fi) {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) tacc[fi][k] = 0.0;
                    tvmax[fi] = 0.0;
                    timax[fi] = -1;
                }
                // ---- the superposition group: all elements with lam < GROUP_MAX_LAM of this pass are one Poisson
                //      process of rate Lambda = sum lam_k; each of its N events belongs to member k with
                //      probability lam_k / Lambda (exact: superposition / thinning of Poisson processes)
                if (ngrp > 0) {
                    // Events of HEAD members are parked (six 10-bit record indices in one 64-bit word) and appended to
                    // the buckets after the event loop.  Appending inside the loop makes every turn of the loop in which
                    // ANY lane holds such an event wait for a global atomic's round trip: passes whose group is made of
                    // head cells ran ten times longer (measured), and those ~100 CTAs were the tail of the kernel.
                    unsigned long long pend = 0ull;
                    int npend = 0;
                    auto flush = [&]() {
                        while (npend > 0) {
                            const Rec rec = s_rec[(int)(pend & 1023ull)];
                            push_event(a, f0 + (int)(rec.meta & 3u), r, rec.cell, 1.0);
                            pend >>= 10;
                            --npend;
                        }
                    };
                    draw_group(s_pool, 1u, s_gspec[0], s_gspec[1], s_gspec[2], glam_tot, s_gcum, ngrp, pass_id, fgk,
                               key, [&](int member) {
                                   const int slot = NREC - 1 - member;
                                   const Rec rec = s_rec[slot];
                                   if (has_events(VARIANT) && (rec.meta & META_HEAD) && !gslot) {
                                       if (npend == 6) flush();
                                       pend = (pend << 10) | (unsigned long long)slot;
                                       ++npend;
                                   } else if (SACC) {
                                       fold_sacc<VARIANT>(a, rec, f0, r, 1.0, &s_acc[0][0][SACC ? tid : 0] + u * THREADS, SACC_STRIDE, gslot);
                                   } else {
                                       fold_rec<VARIANT>(a, rec, s_w3[NACC > 1 ? slot : 0], s_w4[NACC > 4 ? slot : 0], f0, r, 1.0,
                                                         tacc, tvmax, timax, gslot);
                                   }
                               });
                    if (has_events(VARIANT)) flush();
                }
                // ---- phase B, lane-decoupled: each lane walks the PTRS list at its own pace (one rejection trial
                //      per loop turn), so a rejected proposal delays only its own lane
                int it = 0;
                uint32_t trial = 0;
                FPrep pp;
                while (it < np) {
                    const int i = s_plist[it];
                    const Rec rec = s_rec[i];
                    if (trial == 0) prep_draw(rec.lam, a.thresh, pp);
                    const uint64_t idx = (uint64_t)(uint32_t)rec.cell * (uint64_t)a.F_key + (uint64_t)(a.f_key0 + f0 + (int)(rec.meta & 3u));
                    double k;
                    const bool ok = ptrs_trial(pp, element_bits(key, idx, trial), &k);
                    if (ok) {
                        if (SACC) fold_sacc<VARIANT>(a, rec, f0, r, k, &s_acc[0][0][SACC ? tid : 0] + u * THREADS, SACC_STRIDE, gslot);
                        else fold_rec<VARIANT>(a, rec, s_w3[NACC > 1 ? i : 0], s_w4[NACC > 4 ? i : 0], f0, r, k, tacc, tvmax, timax, gslot);
                        ++it;
                        trial = 0;
                    } else {
                        ++trial;
                    }
                }
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    if (!SACC && j == u) {
#pragma unroll
                        for (int fi = 0; fi < FGROUP; ++fi) {
#pragma unroll
                            for (int k = 0; k < NACC; ++k) acc[j][fi][k] += tacc[fi][k];
                            if (has_max(VARIANT) && timax[fi] >= 0 &&
                                (tvmax[fi] > vmax[j][fi] || (tvmax[fi] == vmax[j][fi] && timax[fi] < imax[j][fi]))) {
                                vmax[j][fi] = tvmax[fi];
                                imax[j][fi] = timax[fi];
                            }
                        }
                    }
                }
            }
        }
        HOLO_PHASE_MARK(3)   // superposition group + PTRS
    }
    HOLO_CTA_SPAN(1)

#pragma unroll
    for (int t = 0; t < RPT; ++t) {
        const int r = r_first + t * (int)blockDim.x;
        if (r >= Rtot) continue;
        if (FUSED && r >= a.R) {      // fused background slot: one sum per frequency
#pragma unroll
            for (int fi = 0; fi < FGROUP; ++fi) {
                if (fi >= nf) break;
                a.partial_g[((int64_t)chunk_id * a.F + f0 + fi) * a.Rg + (r - a.R)] =
                    SACC ? s_acc[SACC ? fi : 0][SACC ? t : 0][SACC ? tid : 0] : acc[t][fi][0];
            }
            continue;
        }
#pragma unroll
        for (int fi = 0; fi < FGROUP; ++fi) {
            if (fi >= nf) break;
            const int f = f0 + fi;
            int64_t pb = ((int64_t)chunk_id * a.F + f) * NACC;
#pragma unroll
            for (int k = 0; k < NACC; ++k)
                a.partial[(pb + k) * a.R + r] = SACC ? s_acc[SACC ? fi : 0][SACC ? t : 0][SACC ? tid : 0] : acc[t][fi][k];
            if (has_max(VARIANT)) {
                int64_t mb = ((int64_t)chunk_id * a.F + f) * a.R + r;
                a.pmax[mb] = vmax[t][fi];
                a.pidx[mb] = imax[t][fi];
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// QUAD form of the realization kernel for the parameter variant V_LOUD_PAR_REDZ (cyutils.pyx:1615-1767), the one
// `sam.gwb(params=True)` and `librarian.run_model` use.
//
// The variant folds every draw into EIGHT sums per frequency (hc2 and hc2 times M, q, z, z_final, d_c, a, theta).  In
// the general kernel above a thread owns one realization and all four frequencies of its CTA: 32 fp64 accumulators =
// 64 registers per realization slot, so it runs ONE slot per thread (one dependency chain, 1.1 KB of spills, four
// realization tiles at R = 1000 that each re-stage and re-build everything): 3.5x slower than the plain split.
// Here a CTA works on ONE frequency and a thread owns a QUAD of four CONSECUTIVE realizations:
//   * 4 slots x 8 sums = the same 64 accumulator registers now carry four lock-step dependency chains for the whole
//     chunk, and one CTA carries 1024 realizations -- a single tile, one staging / table build per pass, at R = 1000;
//   * the Philox block of a table draw is keyed on (element, quad): its four words are the four realizations of the
//     quad, so the generator still costs a quarter of a block per draw;
//   * the record's eight weights h, h M, ..., h theta are formed once per thread and shared by the four slots;
//   * the staging window is 1024 cells (one frequency per cell: the same ~170 non-empty elements per pass as the
//     256-cell window of the four-frequency kernels).
// [A first form kept the four frequencies per CTA and walked them one after the other, adding the 32 sums of each
//  frequency sub-pass to global partial sums: a quarter of its time went into those read-modify-writes and another
//  quarter into replaying the pass's superposition group once per frequency.]
// Realizations of a quad: global index 4 Q + u, Q = (r0 >> 2) + local quad; slots outside [r0, r0 + R) are masked, so
// any partition of the realizations over launches / GPUs gives identical numbers.
// -------------------------------------------------------------------------------------------------
constexpr int QUAD_NSCAN = 1024;

// words `sub .. sub + N - 1` of a quad's Philox block (sub is a multiple of N)
template <int N>
__device__ __forceinline__ void quad_words(const Philox4& b, int sub, uint32_t (&w)[N]) {
    if (N == 4) {
#pragma unroll
        for (int u = 0; u < N; ++u) w[u] = b.v[u];
    } else if (N == 2) {
        w[0] = sub ? b.v[2] : b.v[0];
        w[N - 1] = sub ? b.v[3] : b.v[1];
    } else {
        w[0] = sub == 0 ? b.v[0] : (sub == 1 ? b.v[1] : (sub == 2 ? b.v[2] : b.v[3]));
    }
}

// QUAD = realizations of a quad carried by ONE thread: 4 (a CTA of 256 threads carries 1024 realizations), or -- small
// realization counts: library samples -- 2 or 1, so that the draws still spread over the warps of the CTA; the 4 / QUAD
// threads of a quad then each compute the quad's Philox block and use their own words of it (the random numbers, and
// with them every result, do not depend on QUAD).
#ifndef HOLO_QUAD_PLAIN_CTAS
#define HOLO_QUAD_PLAIN_CTAS 3
#endif
// resident CTAs per SM the register allocation is asked to allow: one-sum variants 3 (80 registers), the
// parameter variants 2 when a thread carries four slots of 4 or 8 sums
__host__ __device__ constexpr int quad_min_ctas(int variant, int quad) {
    return nacc_of(variant) == 1 ? HOLO_QUAD_PLAIN_CTAS : ((nacc_of(variant) == 4 ? quad < 4 : quad == 1) ? 3 : 2);
}

template <int VARIANT, int QUAD, int THREADS, bool FUSED>
__global__ void __launch_bounds__(THREADS, quad_min_ctas(VARIANT, QUAD))
realize_quad_kernel(RealizeArgs a) {
    static_assert(!has_max(VARIANT), "the arg-max variants (ss_bg_hc) use the general kernel");
    static_assert(!FUSED || has_events(VARIANT), "fused background slots exist in the loudest variants only");
    constexpr int NACC = nacc_of(VARIANT);
    constexpr int NREC = nrec_of(NACC);
    __shared__ __align__(16) Rec s_rec[NREC];
    __shared__ double s_w3[NACC > 1 ? NREC : 1][3];
    __shared__ double s_w4[NACC > 4 ? NREC : 1][4];
    __shared__ double s_gcum[NREC];
    __shared__ unsigned short s_plist[NREC];
    __shared__ unsigned long long s_wsum[RZ_THREADS / 32];
    __shared__ double s_wlam[RZ_THREADS / 32];
    __shared__ unsigned long long s_tot;
    __shared__ double s_totlam;
    __shared__ int s_np;
    __shared__ int s_gspec[3];
    extern __shared__ __align__(16) uint32_t s_pool[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = THREADS / 32;
    int f = blockIdx.x, chunk_id = blockIdx.y;          // ONE frequency per CTA
    if (a.order != nullptr) {
        const int item = a.order[blockIdx.y * gridDim.x + blockIdx.x];
        f = item % (int)gridDim.x;
        chunk_id = item / (int)gridDim.x;
    }
    const uint32_t fk = (uint32_t)(f + a.f_key0);       // global frequency index (Philox counter)
    const int64_t c_lo = (int64_t)chunk_id * a.chunk;
    int64_t c_hi = c_lo + a.chunk;
    if (c_hi > a.ncell) c_hi = a.ncell;
    const bool supplied = a.counts != nullptr;

    // ---- the thread's quad (and, for QUAD < 4, its part of it: slots `sub * QUAD + u`)
    constexpr int TPQ = 4 / QUAD;
    const int tg = blockIdx.z * THREADS + tid;
    const int q = tg / TPQ, sub = (tg % TPQ) * QUAD;
    const int offL = (int)(a.r0 & 3), nqL = (a.R + offL + 3) >> 2;
    const int offG = FUSED ? (int)(a.r0g & 3) : 0, nqG = FUSED ? ((a.Rg + offG + 3) >> 2) : 0;
    const bool gq = FUSED && q >= nqL;                  // a quad of fused background (GWB) realizations
    const bool liveq = q < nqL + nqG;
    const int qq = gq ? q - nqL : q;
    const int rbase = 4 * qq - (gq ? offG : offL) + sub;   // local realization of slot 0 (slots with r < 0 or >= Rme are masked)
    const int Rme = gq ? a.Rg : a.R;
    const uint32_t quad_global = (uint32_t)(((gq ? a.r0g : a.r0) >> 2) + qq);
    const uint32_t k0 = gq ? a.k0g : a.k0, k1 = gq ? a.k1g : a.k1;
    const uint32_t stream = (gq || !has_events(VARIANT)) ? (uint32_t)STREAM_GWB : (uint32_t)STREAM_LOUD;
    bool valid[QUAD];
    DrawKey kk[QUAD];
#pragma unroll
    for (int u = 0; u < QUAD; ++u) {
        valid[u] = liveq && (rbase + u >= 0) && (rbase + u < Rme);
        kk[u].k0 = k0; kk[u].k1 = k1; kk[u].stream = stream;
        kk[u].real = 4u * quad_global + (uint32_t)(sub + u);
    }
    double acc[QUAD][NACC];
#pragma unroll
    for (int u = 0; u < QUAD; ++u)
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[u][k] = 0.0;

    int64_t cb = c_lo;
    HOLO_PHASE_DECL
    while (cb < c_hi) {
        const int ncons = stage_pass<VARIANT, THREADS, QUAD_NSCAN>(a, cb, c_hi, f, 1, s_rec, &s_w3[0][0], &s_w4[0][0], s_gcum, s_wsum,
                                                                   s_wlam, &s_tot, &s_totlam
#ifdef HOLO_PHASE_CLOCKS
                                                                   , ph_last
#endif
                                                                   );
        const uint32_t pass_id = (uint32_t)cb;
        cb += ncons;
        __syncthreads();
        HOLO_PHASE_MARK(0)
        const int nmain = (int)(s_tot & 2047u), ngrp = (int)((s_tot >> 11) & 2047u);
        const double glam_tot = s_totlam;
        if (!supplied) {
            constexpr int NSUB = 32 / BUILD_SW;
            const int sub = lane / BUILD_SW, sl = lane % BUILD_SW;
            for (int base = warp * NSUB; base < nmain; base += nwarp * NSUB) {
                const int i = base + sub;
                const bool act = (i < nmain) && (((s_rec[i < nmain ? i : 0].meta >> 2) & 7u) == CLS_TABLE);
                const Rec& rec = s_rec[act ? i : 0];
                build_table_sub<BUILD_SW>(rec.lam, s_pool + rec.toff, (int)rec.kmin, (int)((rec.meta >> 12) & 4095u), sl, act);
            }
            if (warp == nwarp - 1 && ngrp > 0) {
                const TableSpec ts = table_spec(glam_tot);
                build_table_sub<32>(glam_tot, s_pool + 1, ts.kmin, ts.W, lane, true);
                if (lane == 0) { s_gspec[0] = ts.kmin; s_gspec[1] = ts.W; s_gspec[2] = 31 - __clz(ts.W); }
            }
            if (warp == 0) {
                int np = 0;         // PTRS records with prep_draw's b / vr (see the general kernel)
                for (int base = 0; base < nmain; base += 32) {
                    const int i = base + lane;
                    const bool is = (i < nmain) && (((s_rec[i].meta >> 2) & 7u) == CLS_PTRS);
                    const unsigned bal = __ballot_sync(0xffffffffu, is);
                    if (is) {
                        const int j = np + __popc(bal & ((1u << lane) - 1u));
                        s_plist[j] = (unsigned short)i;
                        FPrep pp;
                        prep_draw(s_rec[i].lam, a.thresh, pp);
                        s_gcum[NREC - 1 - j] = pp.a0;
                        *reinterpret_cast<double*>(&s_rec[i].kmin) = pp.a1;
                    }
                    np += __popc(bal);
                }
                if (lane == 0) s_np = np;
            }
        }
        __syncthreads();
        HOLO_PHASE_MARK(1)
        if (!liveq) continue;
        // ---- phase A: the main records, four slots in lock-step
        for (int i = 0; i < nmain; ++i) {
            const Rec rec = s_rec[i];
            const uint32_t meta = rec.meta;
            const uint32_t cls = (meta >> 2) & 7u;
            if (cls == CLS_PTRS) continue;
            double n[QUAD];
            if (supplied) {
#pragma unroll
                for (int u = 0; u < QUAD; ++u)
                    n[u] = valid[u] ? a.counts[((int64_t)(rbase + u) * a.F + f) * a.ncell + rec.cell] : 0.0;
            } else if (cls == CLS_TABLE) {
                const Philox4 blk = quad_bits(k0, k1, stream, (uint32_t)rec.cell, fk, 0u, quad_global, PURPOSE_QUAD_HI);
                uint32_t word[QUAD];
                quad_words<QUAD>(blk, sub, word);
                uint32_t qi[QUAD];
                const int W = (int)((meta >> 12) & 4095u), lg = (int)((meta >> 8) & 15u);
                table_ladder_n<QUAD>(s_pool, rec.toff, W, lg, word, qi);
                bool amb = false;
#pragma unroll
                for (int u = 0; u < QUAD; ++u) {
                    n[u] = (double)(int)(rec.kmin + (qi[u] - rec.toff));
                    amb |= table_ambiguous(s_pool, rec.toff, W, qi[u], word[u]);
                }
                if (amb) {      // (prob. ~ 12 W 2^-32 per record: the low halves of the quad's uniforms)
                    const Philox4 lo = quad_bits(k0, k1, stream, (uint32_t)rec.cell, fk, 0u, quad_global, PURPOSE_QUAD_LO);
                    uint32_t low[QUAD];
                    quad_words<QUAD>(lo, sub, low);
#pragma unroll
                    for (int u = 0; u < QUAD; ++u)
                        if (table_ambiguous(s_pool, rec.toff, W, qi[u], word[u]))
                            n[u] = table_resolve(rec.lam, s_pool + rec.toff, (int)rec.kmin, W, (int)(qi[u] - rec.toff), word[u], low[u]);
                }
            } else {        // CLS_NORMAL
                const uint64_t idx = (uint64_t)(uint32_t)rec.cell * (uint64_t)a.F_key + (uint64_t)fk;
#pragma unroll
                for (int u = 0; u < QUAD; ++u) n[u] = draw_normal_lam(rec.lam, kk[u], idx);
            }
            // the eight weights of the record (shared by the four slots) and the fold of its draws
            const double h = rec.h;
            double hw[NACC];
            hw[0] = h;
            if (NACC > 1) {
#pragma unroll
                for (int k = 0; k < 3; ++k) hw[(1 + k) < NACC ? 1 + k : 0] = h * s_w3[NACC > 1 ? i : 0][k];
            }
            if (NACC > 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) hw[(4 + k) < NACC ? 4 + k : 0] = h * s_w4[NACC > 4 ? i : 0][k];
            }
            const bool head = has_events(VARIANT) && (meta & META_HEAD) && !gq;
#pragma unroll
            for (int u = 0; u < QUAD; ++u) {
                if (!valid[u]) continue;
                if (head) {
                    if (n[u] >= 1.0) push_event(a, f, rbase + u, rec.cell, n[u]);       // pyx:1727-1742
                } else {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) acc[u][k] += n[u] * hw[k];          // pyx:1745-1752
                }
            }
        }
        HOLO_PHASE_MARK(2)
        if (supplied) continue;
        // ---- the pass's superposition group (see draw_group in holo_rng.cuh), four slots in lock-step
        if (ngrp > 0) {
            const int gk = s_gspec[0], gW = s_gspec[1], glg = s_gspec[2];
            uint32_t w0[QUAD], w1[QUAD], qc[QUAD];
#pragma unroll
            for (int u = 0; u < QUAD; ++u) {
                const Philox4 gb = philox4x32_10(pass_id, fk | ((uint32_t)PURPOSE_PASS_COUNT << 28), kk[u].real, stream << 24, k0, k1);
                w0[u] = gb.v[0];
                w1[u] = gb.v[1];
            }
            table_ladder_n<QUAD>(s_pool, 1u, gW, glg, w0, qc);
            int nev[QUAD], nmax = 0;
#pragma unroll
            for (int u = 0; u < QUAD; ++u) {
                double ne = (double)(gk + (int)(qc[u] - 1u));
                if (table_ambiguous(s_pool, 1u, gW, qc[u], w0[u]))
                    ne = table_resolve(glam_tot, s_pool + 1, gk, gW, (int)(qc[u] - 1u), w0[u], w1[u]);
                nev[u] = valid[u] ? (int)ne : 0;
                nmax = nev[u] > nmax ? nev[u] : nmax;
            }
            uint32_t odd2[QUAD], odd3[QUAD];
            for (int ev = 0; ev < nmax; ++ev) {
                double v[QUAD];
                if ((ev & 1) == 0) {
#pragma unroll
                    for (int u = 0; u < QUAD; ++u) {
                        const Philox4 pb = philox4x32_10(pass_id, fk | ((uint32_t)PURPOSE_PASS_PICK << 28), kk[u].real,
                                                         (stream << 24) | (uint32_t)(ev >> 1), k0, k1);
                        v[u] = u53(pb.v[0], pb.v[1]) * glam_tot;
                        odd2[u] = pb.v[2];
                        odd3[u] = pb.v[3];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < QUAD; ++u) v[u] = u53(odd2[u], odd3[u]) * glam_tot;
                }
                int base[QUAD];
#pragma unroll
                for (int u = 0; u < QUAD; ++u) base[u] = 0;
                int len = ngrp;
                while (len > 1) {
                    const int half = len >> 1;
#pragma unroll
                    for (int u = 0; u < QUAD; ++u)
                        if (s_gcum[base[u] + half - 1] <= v[u]) base[u] += half;
                    len -= half;
                }
#pragma unroll
                for (int u = 0; u < QUAD; ++u) {
                    if (s_gcum[base[u]] <= v[u] && base[u] < ngrp - 1) base[u] += 1;
                    if (ev >= nev[u]) continue;
                    const int slot = NREC - 1 - base[u];
                    const Rec rec = s_rec[slot];
                    if (has_events(VARIANT) && (rec.meta & META_HEAD) && !gq) {
                        push_event(a, f, rbase + u, rec.cell, 1.0);
                    } else {
                        const double h = rec.h;
                        acc[u][0] += h;
                        if (NACC > 1) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) acc[u][(1 + k) < NACC ? 1 + k : 0] += h * s_w3[NACC > 1 ? slot : 0][k];
                        }
                        if (NACC > 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) acc[u][(4 + k) < NACC ? 4 + k : 0] += h * s_w4[NACC > 4 ? slot : 0][k];
                        }
                    }
                }
            }
        }
        // ---- the PTRS records: every slot at its own pace, proposals side by side
        const int np = s_np;
        if (np > 0) {
            int it[QUAD];
            uint32_t trial[QUAD];
#pragma unroll
            for (int u = 0; u < QUAD; ++u) { it[u] = valid[u] ? 0 : np; trial[u] = 0u; }
            for (;;) {
                bool any = false;
#pragma unroll
                for (int u = 0; u < QUAD; ++u) any |= (it[u] < np);
                if (!any) break;
                double kq[QUAD], usq[QUAD], Vq[QUAD];
                int dec[QUAD], ri[QUAD];
#pragma unroll
                for (int u = 0; u < QUAD; ++u) {
                    const int j = it[u] < np ? it[u] : np - 1;
                    ri[u] = s_plist[j];
                    const double lam = s_rec[ri[u]].lam;
                    const double b = s_gcum[NREC - 1 - j];
                    const double vr = *reinterpret_cast<const double*>(&s_rec[ri[u]].kmin);
                    const uint64_t idx = (uint64_t)(uint32_t)s_rec[ri[u]].cell * (uint64_t)a.F_key + (uint64_t)fk;
                    dec[u] = ptrs_propose(lam, b, vr, element_bits(kk[u], idx, trial[u]), &kq[u], &usq[u], &Vq[u]);
                }
#pragma unroll
                for (int u = 0; u < QUAD; ++u) {
                    if (it[u] >= np) continue;
                    bool ok = dec[u] > 0;
                    if (dec[u] == 0) ok = ptrs_decide(s_rec[ri[u]].lam, s_gcum[NREC - 1 - it[u]], usq[u], Vq[u], kq[u]);
                    if (ok) {
                        const Rec rec = s_rec[ri[u]];
                        if (has_events(VARIANT) && (rec.meta & META_HEAD) && !gq) {
                            if (kq[u] >= 1.0) push_event(a, f, rbase + u, rec.cell, kq[u]);
                        } else {
                            const double nh = kq[u] * rec.h;
                            acc[u][0] += nh;
                            if (NACC > 1) {
#pragma unroll
                                for (int k = 0; k < 3; ++k) acc[u][(1 + k) < NACC ? 1 + k : 0] += nh * s_w3[NACC > 1 ? ri[u] : 0][k];
                            }
                            if (NACC > 4) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[u][(4 + k) < NACC ? 4 + k : 0] += nh * s_w4[NACC > 4 ? ri[u] : 0][k];
                            }
                        }
                        ++it[u];
                        trial[u] = 0u;
                    } else {
                        ++trial[u];
                    }
                }
            }
        }
        HOLO_PHASE_MARK(3)
    }
    // ---- the chunk's sums of this frequency
#pragma unroll
    for (int u = 0; u < QUAD; ++u) {
        if (!valid[u]) continue;
        const int r = rbase + u;
        if (gq) {
            a.partial_g[((int64_t)chunk_id * a.F + f) * a.Rg + r] = acc[u][0];
        } else {
            double* dst = a.partial + (((int64_t)chunk_id * a.F + f) * NACC) * a.R + r;
#pragma unroll
            for (int k = 0; k < NACC; ++k) dst[(int64_t)k * a.R] = acc[u][k];
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Launch order of the realization kernel: longest-processing-time-first list scheduling.
// cost_kernel estimates the work of every (chunk, frequency group) item from the sampler classes of its elements
// (a table draw = 16 units, a PTRS draw = 64, a member of a superposition group = 1: measured CTA durations follow
// this to R^2 = 0.998, and the makespan is insensitive to the weights) and histograms the items over COST_BINS
// logarithmic cost classes; order_kernel scatters them heaviest class first.  The order within a class is left to
// the atomics: it only decides WHEN an item runs, never what it computes.
// -------------------------------------------------------------------------------------------------
constexpr int COST_BINS = 256;     // 8 classes per octave

__device__ __forceinline__ int cost_bin(uint32_t c) {     // c >= 1
    const int lz = 31 - __clz(c);
    const uint32_t sub = lz >= 3 ? ((c >> (lz - 3)) & 7u) : ((c << (3 - lz)) & 7u);
    return lz * 8 + (int)sub;
}

__global__ void __launch_bounds__(256)
cost_kernel(const double* __restrict__ number, int64_t ncell, int F, int64_t chunk, int nfg, int fgroup, double thresh,
            const int32_t* __restrict__ rank, const int32_t* __restrict__ kf, double head_weight,
            uint32_t* __restrict__ cost, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_cost[];     // (nfg,)
    for (int i = threadIdx.x; i < nfg; i += blockDim.x) s_cost[i] = 0u;
    __syncthreads();
    const int64_t c_lo = (int64_t)blockIdx.x * chunk;
    int64_t c_hi = c_lo + chunk;
    if (c_hi > ncell) c_hi = ncell;
    const int64_t n = (c_hi - c_lo) * F;
    const double* base = number + c_lo * F;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
        const double lam = base[e];
        if (!(lam > 0.0)) continue;
        uint32_t w = lam < GROUP_MAX_LAM ? 1u : ((lam <= TABLE_MAX_LAM || lam > thresh) ? 16u : 64u);
        const int f = (int)(e % F);
        // an occupied head cell of the loudest variants appends an event per realization (a global atomic and a
        // dependent 16 B store, out of line).  Measured: the ~100 items that hold the occupied head cells run 2-3 ms
        // longer than their tables predict (0.6 ms per expected event per realization at 1024 realizations per CTA);
        // left unweighted they start late and ARE the tail of the kernel.  Overweighting only starts them earlier.
        if (rank != nullptr && rank[c_lo + e / F] < kf[f]) w += 32u + (uint32_t)(head_weight * (lam >= 1.0 ? 1.0 : lam));
        atomicAdd(&s_cost[f / fgroup], w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nfg; i += blockDim.x) {
        const uint32_t c = s_cost[i] + 1u;
        cost[(int64_t)blockIdx.x * nfg + i] = c;
        atomicAdd(&hist[cost_bin(c)], 1u);
    }
}

__global__ void __launch_bounds__(256)
order_kernel(const uint32_t* __restrict__ cost, const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, int n,
             int32_t* __restrict__ order) {
    __shared__ uint32_t s_first[COST_BINS];     // first position of each class, heaviest class first
    if (threadIdx.x == 0) {
        uint32_t run = 0u;
        for (int b = COST_BINS - 1; b >= 0; --b) {
            s_first[b] = run;
            run += hist[b];
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = cost_bin(cost[i]);
    order[s_first[b] + atomicAdd(&cursor[b], 1u)] = i;
}

// -------------------------------------------------------------------------------------------------
// Head preparation for the loudest variants
// -------------------------------------------------------------------------------------------------
__global__ void rank_inverse_kernel(const int32_t* __restrict__ order, int64_t ncell,
                                    int32_t* __restrict__ rank) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < ncell;
         p += (int64_t)gridDim.x * blockDim.x)
        rank[order[p]] = (int32_t)p;
}

constexpr int HEAD_ROWS = 128;    // rank positions per block = granularity of the head cut K_f

// bsum[blk][f] = sum over rank positions of the block of P(cell occupied) at frequency f
__global__ void __launch_bounds__(256)
head_sum_kernel(const double* __restrict__ number, const double* __restrict__ h2fdf,
                const int32_t* __restrict__ order, int64_t ncell, int F, double thresh, int need_h,
                int supplied, double* __restrict__ bsum) {
    extern __shared__ double s_sum[];
    for (int f = threadIdx.x; f < F; f += blockDim.x) s_sum[f] = 0.0;
    __syncthreads();
    int64_t p0 = (int64_t)blockIdx.x * HEAD_ROWS;
    int64_t nel = (int64_t)HEAD_ROWS * F;
    for (int64_t i = threadIdx.x; i < nel; i += blockDim.x) {
        int64_t p = p0 + i / F;
        int f = (int)(i % F);
        if (p >= ncell) break;
        int64_t c = order[p];
        double lam = number[c * F + f];
        bool elig = (lam > 0.0);
        if (need_h) elig = elig && (h2fdf[c * F + f] != 0.0);
        (void)supplied;   // supplied counts are assumed to be draws of `number`: same head estimate
        double pocc = 0.0;
        if (elig) pocc = (lam > thresh) ? 1.0 : -expm1(-lam);
        if (pocc > 0.0) atomicAdd(&s_sum[f], pocc);
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) bsum[(int64_t)blockIdx.x * F + f] = s_sum[f];
}

// One CTA per frequency: threads sum contiguous segments of the per-block occupancy sums, a block scan (thread order)
// locates the segment in which the cumulative sum crosses `target`, and that thread walks its segment (fixed order).
constexpr int HEAD_CUT_THREADS = 256;
__global__ void __launch_bounds__(HEAD_CUT_THREADS)
head_cut_kernel(const double* __restrict__ bsum, int nblk, int F, int64_t ncell, double target, int32_t* __restrict__ kf) {
    __shared__ double s_incl[HEAD_CUT_THREADS];
    __shared__ double s_wtot[HEAD_CUT_THREADS / 32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = (nblk + HEAD_CUT_THREADS - 1) / HEAD_CUT_THREADS;
    const int b0 = min(nblk, tid * seg), b1 = min(nblk, b0 + seg);
    double mine = 0.0;
    for (int b = b0; b < b1; ++b) mine += bsum[(int64_t)b * F + f];
    double incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    double before = 0.0;
    for (int w = 0; w < warp; ++w) before += s_wtot[w];
    incl += before;
    s_incl[tid] = incl;
    __syncthreads();
    const double prev = tid > 0 ? s_incl[tid - 1] : 0.0;
    const bool crosses = (incl >= target) && (prev < target);     // at most one thread: s_incl is non-decreasing
    if (crosses) {
        int64_t k = (int64_t)b1 * HEAD_ROWS;
        double cum = prev;
        for (int b = b0; b < b1; ++b) {
            cum += bsum[(int64_t)b * F + f];
            if (cum >= target) { k = (int64_t)(b + 1) * HEAD_ROWS; break; }
        }
        if (k > ncell) k = ncell;
        kf[f] = (int32_t)k;
    }
    if (!__syncthreads_or(crosses) && tid == 0) kf[f] = (int32_t)ncell;
}

// -------------------------------------------------------------------------------------------------
// Resolver: one warp per (f, r) hands out the L slots in rank order (with multiplicity) and sums
// what is left of the head into the background.
// -------------------------------------------------------------------------------------------------
struct ResolveArgs {
    const Event* events;
    const int32_t* evcount;
    const int32_t* kf;
    const double* h2fdf;
    const double* mt;
    const double* mr;
    const double* rz;
    const double* redz_final;
    const double* dcom_final;
    const double* sepa;
    const double* angs;
    double* hc2ss;    // (F,R,L)
    double* sspar;    // (4,F,R,L)
    double* lspar;    // (3,F,R)
    int64_t* ssidx;   // (3,F,R,L)
    double* rem;      // (F, NACC, R)
    int32_t* flags;   // [0]: bucket overflow, [1]: head too short
    int64_t ncell;
    int Qb, Zb, F, R, L, cap;
};

constexpr int RES_WARPS = 4;

template <int VARIANT>
__global__ void __launch_bounds__(RES_WARPS * 32)
resolve_kernel(ResolveArgs a) {
    constexpr int NACC = nacc_of(VARIANT);
    extern __shared__ unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t fr = (int64_t)blockIdx.x * RES_WARPS + warp;
    if (fr >= (int64_t)a.F * a.R) return;
    const int f = (int)(fr / a.R);
    const int r = (int)(fr % a.R);
    Event* raw = reinterpret_cast<Event*>(s_raw) + (size_t)warp * 2 * a.cap;
    Event* ev = raw + a.cap;

    int cnt = a.evcount[fr];
    if (cnt > a.cap) {
        if (lane == 0) atomicOr(&a.flags[0], 1);
        cnt = a.cap;
    }
    for (int i = lane; i < cnt; i += 32) raw[i] = a.events[fr * a.cap + i];
    __syncwarp();
    // The bucket was filled with atomics, i.e. in arbitrary order: sort it by rank (ranks are unique)
    // so that everything below is bit-reproducible.  cnt is ~L + margin, a counting sort is plenty.
    for (int i = lane; i < cnt; i += 32) {
        int rk = raw[i].rank, pos = 0;
        // (a cell that drew n = 2 through the superposition group shows up as two events of n = 1)
        for (int j = 0; j < cnt; ++j) pos += (raw[j].rank < rk || (raw[j].rank == rk && j < i)) ? 1 : 0;
        ev[pos] = raw[i];
    }
    __syncwarp();

    double rem[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) rem[k] = 0.0;
    double ls[4] = {0.0, 0.0, 0.0, 0.0};
    int ll = 0;
    int next = 0;   // first event that did not (fully) go into the loudest slots

    // walk the head in rank order until the L slots are full
    while (ll < a.L && next < cnt) {
        Event e = ev[next];
        ++next;
        const int64_t c = e.cell;
        const double cur = a.h2fdf[c * a.F + f];
        int zz = (int)(c % a.Zb);
        int64_t mq = c / a.Zb;
        int qq = (int)(mq % a.Qb);
        int mm = (int)(mq / a.Qb);
        double num = e.n;
        // `while (ll < L) and (num > 0)`  pyx:1338-1341, 1493-1503, 1730-1742
        while (ll < a.L && num > 0.0) {
            if (lane == 0) {
                int64_t o = fr * a.L + ll;
                a.hc2ss[o] = cur;
                if (VARIANT == V_LOUD_PAR) {
                    int64_t st = (int64_t)a.F * a.R * a.L;
                    a.ssidx[o] = mm; a.ssidx[st + o] = qq; a.ssidx[2 * st + o] = zz;
                }
                if (VARIANT == V_LOUD_PAR_REDZ) {
                    int64_t st = (int64_t)a.F * a.R * a.L;
                    a.sspar[o] = a.mt[mm];
                    a.sspar[st + o] = a.mr[qq];
                    a.sspar[2 * st + o] = a.rz[zz];
                    a.sspar[3 * st + o] = a.redz_final[c * a.F + f];
                }
            }
            if (VARIANT == V_LOUD_PAR) {
                ls[0] += cur; ls[1] += cur * a.mt[mm]; ls[2] += cur * a.mr[qq]; ls[3] += cur * a.rz[zz];
            }
            num -= 1.0;
            ll += 1;
        }
        // what is left of this cell goes to the background (pyx:1342, 1505-1508, 1745-1752)
        double nc = num * cur;
        rem[0] += nc;
        if (NACC > 1) {
            rem[1] += nc * a.mt[mm];
            if (NACC > 2) { rem[2] += nc * a.mr[qq]; rem[3] += nc * a.rz[zz]; }
            if (NACC > 4) {
                rem[4] += nc * a.redz_final[c * a.F + f];
                rem[5] += nc * a.dcom_final[c * a.F + f];
                rem[6] += nc * a.sepa[c * a.F + f];
                rem[7] += nc * a.angs[c * a.F + f];
            }
        }
    }
    // unfilled slots stay zero (outputs are zero-initialised by the host wrapper)
    if (ll < a.L && a.kf[f] < a.ncell && lane == 0) atomicOr(&a.flags[1], 1);

    // the rest of the head is pure background: lane-strided sums in rank order + fixed butterfly
    double part[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) part[k] = 0.0;
    for (int i = next + lane; i < cnt; i += 32) {
        Event e = ev[i];
        const int64_t c = e.cell;
        double nc = e.n * a.h2fdf[c * a.F + f];
        part[0] += nc;
        if (NACC > 1) {
            int zz = (int)(c % a.Zb);
            int64_t mq = c / a.Zb;
            int qq = (int)(mq % a.Qb);
            int mm = (int)(mq / a.Qb);
            part[1] += nc * a.mt[mm];
            if (NACC > 2) { part[2] += nc * a.mr[qq]; part[3] += nc * a.rz[zz]; }
            if (NACC > 4) {
                part[4] += nc * a.redz_final[c * a.F + f];
                part[5] += nc * a.dcom_final[c * a.F + f];
                part[6] += nc * a.sepa[c * a.F + f];
                part[7] += nc * a.angs[c * a.F + f];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        double v = part[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) a.rem[((int64_t)f * NACC + k) * a.R + r] = rem[k] + v;
    }
    if (VARIANT == V_LOUD_PAR && lane == 0) {
        int64_t st = (int64_t)a.F * a.R;
        a.lspar[fr] = ls[1] / ls[0];              // pyx:1513-1515
        a.lspar[st + fr] = ls[2] / ls[0];
        a.lspar[2 * st + fr] = ls[3] / ls[0];
    }
}

// -------------------------------------------------------------------------------------------------
// Final fixed-order reduction over chunks
// -------------------------------------------------------------------------------------------------
struct FinalArgs {
    const double* partial;   // (nchunk, F, NACC, R)
    const double* rem;       // (F, NACC, R) or NULL
    const double* pmax;
    const int32_t* pidx;
    const double* h2fdf;
    const double* mt;
    const double* mr;
    const double* rz;
    double* out0;            // gwb / hc2bg (F,R)
    double* bgpar;           // (NACC-1, F, R)
    double* hc2ss;           // ss_bg: (F,R)
    double* sspar;           // ss_bg_par: (3,F,R)
    int64_t* ssidx;          // ss_bg: (3,F,R)
    int32_t* flags;          // [2]: ss_bg_par found no source
    int nchunk, Qb, Zb, F, R;
};

constexpr int FIN_R = 32;      // realizations per block (one 256 B row of `partial` per load)
constexpr int FIN_SEG = 8;     // chunk segments summed concurrently, then combined in segment order

template <int VARIANT>
__global__ void __launch_bounds__(FIN_R * FIN_SEG)
final_kernel(FinalArgs a) {
    constexpr int NACC = nacc_of(VARIANT);
    __shared__ double s_sum[FIN_SEG][NACC][FIN_R];
    __shared__ double s_max[FIN_SEG][FIN_R];
    __shared__ int s_idx[FIN_SEG][FIN_R];
    const int rl = threadIdx.x % FIN_R, seg = threadIdx.x / FIN_R;
    const int f = blockIdx.y;
    const int r = blockIdx.x * FIN_R + rl;
    const bool live = r < a.R;
    const int per = (a.nchunk + FIN_SEG - 1) / FIN_SEG;
    const int ch0 = seg * per, ch1 = min(a.nchunk, ch0 + per);
    double s[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) s[k] = 0.0;
    double vmax = 0.0;
    int imax = -1;
    if (live) {
        for (int ch = ch0; ch < ch1; ++ch) {
            int64_t pb = ((int64_t)ch * a.F + f) * NACC;
#pragma unroll
            for (int k = 0; k < NACC; ++k) s[k] += a.partial[(pb + k) * a.R + r];
            if (has_max(VARIANT)) {
                int64_t mb = ((int64_t)ch * a.F + f) * a.R + r;
                double v = a.pmax[mb];
                if (v > vmax) { vmax = v; imax = a.pidx[mb]; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) s_sum[seg][k][rl] = s[k];
    s_max[seg][rl] = vmax;
    s_idx[seg][rl] = imax;
    __syncthreads();
    if (seg != 0 || !live) return;
    for (int g = 1; g < FIN_SEG; ++g) {          // fixed order: chunks ascending
#pragma unroll
        for (int k = 0; k < NACC; ++k) s[k] += s_sum[g][k][rl];
        if (has_max(VARIANT) && s_max[g][rl] > vmax) { vmax = s_max[g][rl]; imax = s_idx[g][rl]; }
    }
    const int64_t fr = (int64_t)f * a.R + r;
    if (a.rem) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) s[k] += a.rem[((int64_t)f * NACC + k) * a.R + r];
    }
    int64_t st = (int64_t)a.F * a.R;
    if (has_max(VARIANT)) {
        int zz = -1, qq = -1, mm = -1;
        if (imax >= 0) {
            zz = imax % a.Zb;
            int mq = imax / a.Zb;
            qq = mq % a.Qb;
            mm = mq / a.Qb;
        }
        a.hc2ss[fr] = vmax;                         // pyx:1003-1008, 1145-1165
        a.out0[fr] = s[0] - vmax;
        a.ssidx[fr] = mm; a.ssidx[st + fr] = qq; a.ssidx[2 * st + fr] = zz;
        if (VARIANT == V_SSBG_PAR) {
            if (imax < 0) {
                atomicOr(&a.flags[2], 1);           // pyx:1157-1158 bare `raise`
            } else {
                double hm = a.h2fdf[(int64_t)imax * a.F + f];
                double den = s[0] - vmax;
                a.bgpar[fr] = (s[1] - hm * a.mt[mm]) / den;
                a.bgpar[st + fr] = (s[2] - hm * a.mr[qq]) / den;
                a.bgpar[2 * st + fr] = (s[3] - hm * a.rz[zz]) / den;
                a.sspar[fr] = a.mt[mm];
                a.sspar[st + fr] = a.mr[qq];
                a.sspar[2 * st + fr] = a.rz[zz];
            }
        }
    } else {
        a.out0[fr] = s[0];
        if (NACC > 1) {
#pragma unroll
            for (int k = 1; k < NACC; ++k) a.bgpar[(k - 1) * st + fr] = s[k] / s[0];   // pyx:1509-1512, 1761-1767
        }
    }
}

// -------------------------------------------------------------------------------------------------
// poisson_as_needed (gravwaves.py:666-691): elementwise, normal branch floored
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bulk_poisson_kernel(const double* __restrict__ lam, int64_t n, uint32_t k0, uint32_t k1,
                    uint32_t stream_id, double thresh, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double l = lam[i];
        DrawKey key;
        key.k0 = k0; key.k1 = k1;
        key.real = stream_id; key.stream = STREAM_BULK;
        double v = draw_element(l, thresh, key, (uint64_t)i);
        if (l > thresh) v = floor(v);
        out[i] = v;
    }
}

// -------------------------------------------------------------------------------------------------
// host-side planning
// -------------------------------------------------------------------------------------------------
struct Plan {
    int threads, ntiles, nfg, nchunk, rpt;
    int64_t chunk;
};

// Which variants run the quad kernel (one frequency per CTA): the parameter variants.  The one-sum variants (realised
// GWB, plain loudest split) stay on the general four-frequency kernel: measured on B200 at the named grid the quad
// form is no faster for them at R = 1000 (10.07 vs 10.06 ms) and slower at small R (R = 100: 4.8 vs 3.5 ms) -- their
// accumulators already live in shared memory and four frequencies share a staging pass.  -DHOLO_PLAIN_QUAD builds
// the quad form for them too (A/B comparison).
static constexpr bool uses_quad_kernel(int variant) {
#ifdef HOLO_PLAIN_QUAD
    return !has_max(variant);
#else
    return variant == V_LOUD_PAR_REDZ || variant == V_LOUD_PAR;
#endif
}

static Plan make_plan(int64_t ncell, int F, int R, int variant) {
    Plan p;
    const int rpt = rpt_for(variant, R);
    p.rpt = rpt;
    p.threads = threads_for(rpt, R);
    p.ntiles = (R + p.threads * rpt - 1) / (p.threads * rpt);
    // frequency items per chunk: groups of FGROUP, or -- quad kernel of the parameter variant -- single frequencies
    p.nfg = uses_quad_kernel(variant) ? F : (F + FGROUP - 1) / FGROUP;
    // The cell chunking must not depend on R (or on the realization tiling): per-chunk partial sums are
    // combined in a fixed order, so a fixed chunking makes hc2 bit-identical however the realizations
    // are partitioned over launches / GPUs (partials cost nchunk*F*NACC*R*8 bytes: 328 MB ... 2.6 GB at R = 1000).
    // 1024 chunks: the heaviest (chunk, frequency group) item bounds the kernel from below -- 9.2 ms of a 10.3 ms
    // launch with 512 chunks at R = 1000, and MORE than the balanced time at R = 100 (4.2 vs 3.2 ms).  1024 halves
    // it (R = 100: 4.3 -> 3.5 ms); 2048 gains nothing more and costs partial-sum traffic at R = 1000.
    // The quad kernels take ONE frequency per item: the same number of items (and the same work per item) needs a
    // quarter of the chunks, and their 11-accumulator partial sums -- nchunk*F*11*R*8 bytes, 3.6 GB at R = 1000 with
    // 1024 chunks -- shrink with it.  Measured (loud+par L=5, R = 1000 / R = 100, ms): 1024 chunks 16.84 / 5.97,
    // 512: 16.12 / 5.91, 256: 15.75 / 5.92, 128: 15.90 / 6.18  (-DHOLO_QUAD_NCHUNK=... for sweeps).
#ifndef HOLO_QUAD_NCHUNK
#define HOLO_QUAD_NCHUNK 256
#endif
    int64_t nchunk = uses_quad_kernel(variant) ? HOLO_QUAD_NCHUNK : 1024;
    int64_t chunk = (ncell + nchunk - 1) / nchunk;
    chunk = ((chunk + 63) / 64) * 64;
    if (chunk < 64) chunk = 64;
    p.chunk = chunk;
    p.nchunk = (int)((ncell + chunk - 1) / chunk);
    if (p.nchunk < 1) p.nchunk = 1;
    return p;
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

static int auto_cap(int L, double margin) {
    double target = L + margin;
    int cap = 64;
    while (cap < 2.0 * target + 64.0) cap *= 2;
    return cap;
}

static double auto_margin(int L) { return 8.0 * sqrt((double)L) + 24.0; }

struct Workspace {
    unsigned char* base;
    int64_t used, size;
    void* take(int64_t bytes) {
        void* p = base ? base + used : nullptr;
        used += align256(bytes);
        return p;
    }
};

struct Layout {
    double* partial; double* pmax; int32_t* pidx; Event* events; int32_t* evcount; double* rem;
    int32_t* flags; int32_t* kf; int32_t* rank; double* bsum;
    int32_t* order; uint32_t* cost; uint32_t* hist;   // launch order of the realization kernel (hist: 2 x COST_BINS)
    int64_t total;
};

static Layout carve(void* ws, int variant, int64_t ncell, int F, int R, int cap, const Plan& p) {
    Workspace w{(unsigned char*)ws, 0, 0};
    int nacc = nacc_of(variant);
    Layout l{};
    l.flags = (int32_t*)w.take(4 * sizeof(int32_t));
    l.partial = (double*)w.take(sizeof(double) * (int64_t)p.nchunk * F * nacc * R);
    if (has_max(variant)) {
        l.pmax = (double*)w.take(sizeof(double) * (int64_t)p.nchunk * F * R);
        l.pidx = (int32_t*)w.take(sizeof(int32_t) * (int64_t)p.nchunk * F * R);
    }
    if (has_events(variant)) {
        l.events = (Event*)w.take(sizeof(Event) * (int64_t)F * R * cap);
        l.evcount = (int32_t*)w.take(sizeof(int32_t) * (int64_t)F * R);
        l.rem = (double*)w.take(sizeof(double) * (int64_t)F * nacc * R);
        l.kf = (int32_t*)w.take(sizeof(int32_t) * F);
        l.rank = (int32_t*)w.take(sizeof(int32_t) * ncell);
        int nblk = (int)((ncell + HEAD_ROWS - 1) / HEAD_ROWS);
        l.bsum = (double*)w.take(sizeof(double) * (int64_t)nblk * F);
    }
    const int64_t nitem = (int64_t)p.nchunk * p.nfg;
    l.order = (int32_t*)w.take(sizeof(int32_t) * nitem);
    l.cost = (uint32_t*)w.take(sizeof(uint32_t) * nitem);
    l.hist = (uint32_t*)w.take(sizeof(uint32_t) * 2 * COST_BINS);
    l.total = w.used;
    return l;
}

// fills l.order (heaviest items first) on `st`
static int plan_launch_order(const double* number, int64_t ncell, int F, double thresh, const int32_t* rank,
                             const int32_t* kf, int R, const Plan& p, const Layout& l, cudaStream_t st, int fgroup = FGROUP) {
    const int per_cta = R < p.threads * p.rpt ? R : p.threads * p.rpt;     // realizations one CTA carries
    const double head_weight = 8.0 * per_cta;
    const int n = p.nchunk * p.nfg;
    HOLO_CUDA(cudaMemsetAsync(l.hist, 0, sizeof(uint32_t) * 2 * COST_BINS, st));
    cost_kernel<<<p.nchunk, 256, sizeof(uint32_t) * p.nfg, st>>>(number, ncell, F, p.chunk, p.nfg, fgroup, thresh, rank, kf, head_weight, l.cost, l.hist);
    order_kernel<<<(n + 255) / 256, 256, 0, st>>>(l.cost, l.hist, l.hist + COST_BINS, n, l.order);
    holo::count_launches(2);
    return holo_check_launch("realization launch order");
}

template <int VARIANT, int RPT, int THREADS, bool FUSED>
static int launch_realize_fused(const RealizeArgs& ra, const Plan& p, cudaStream_t st) {
    dim3 grid(p.nfg, p.nchunk, p.ntiles);
    const bool sacc = nacc_of(VARIANT) == 1 && !has_max(VARIANT);
    const size_t pool_bytes = sizeof(uint32_t) * pool_entries_of(VARIANT) + (sacc ? sizeof(double) * FGROUP * RPT * THREADS : 0) +
                              sizeof(uint32_t) * 4 * RPT * THREADS;
    // static + dynamic shared memory may exceed the 48 KB default.  Function attributes are per DEVICE (each device has
    // its own copy of the loaded function): remember the opt-in per device ordinal, not per process.
    static bool attr_set[64] = {};
    int dev = 0;
    HOLO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        HOLO_CUDA(cudaFuncSetAttribute(realize_kernel<VARIANT, RPT, THREADS, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_bytes));
        HOLO_CUDA(cudaFuncSetAttribute(realize_kernel<VARIANT, RPT, THREADS, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    realize_kernel<VARIANT, RPT, THREADS, FUSED><<<grid, THREADS, pool_bytes, st>>>(ra); holo::count_launches(1);
    return holo_check_launch("realize_kernel");
}

template <int VARIANT, int RPT, int THREADS>
static int launch_realize_rpt(const RealizeArgs& ra, const Plan& p, cudaStream_t st) {
    if (has_events(VARIANT) && ra.Rg > 0) return launch_realize_fused<VARIANT, RPT, THREADS, has_events(VARIANT)>(ra, p, st);
    return launch_realize_fused<VARIANT, RPT, THREADS, false>(ra, p, st);
}

template <int VARIANT, int QN, bool FUSED>
static int launch_quad_fused(const RealizeArgs& ra, const Plan& p, int nq, cudaStream_t st) {
    const int nthreads = nq * (4 / QN);
    dim3 grid(p.nfg, p.nchunk, (nthreads + RZ_THREADS - 1) / RZ_THREADS);
    const size_t pool_bytes = sizeof(uint32_t) * pool_entries_of(VARIANT);
    static bool attr_set[64] = {};
    int dev = 0;
    HOLO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        HOLO_CUDA(cudaFuncSetAttribute(realize_quad_kernel<VARIANT, QN, RZ_THREADS, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_bytes));
        HOLO_CUDA(cudaFuncSetAttribute(realize_quad_kernel<VARIANT, QN, RZ_THREADS, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    realize_quad_kernel<VARIANT, QN, RZ_THREADS, FUSED><<<grid, RZ_THREADS, pool_bytes, st>>>(ra); holo::count_launches(1);
    return holo_check_launch("realize_quad_kernel");
}

// quads of realizations a launch of the quad kernel needs (loudest realizations, then the fused GWB ones)
static int quad_count(const RealizeArgs& ra) {
    const int nqL = (ra.R + (int)(ra.r0 & 3) + 3) >> 2;
    const int nqG = ra.Rg > 0 ? ((ra.Rg + (int)(ra.r0g & 3) + 3) >> 2) : 0;
    return nqL + nqG;
}

template <int VARIANT, int QN>
static int launch_quad_n(const RealizeArgs& ra, const Plan& p, int nq, cudaStream_t st) {
    if (has_events(VARIANT) && ra.Rg > 0) return launch_quad_fused<VARIANT, QN, has_events(VARIANT)>(ra, p, nq, st);
    return launch_quad_fused<VARIANT, QN, false>(ra, p, nq, st);
}

template <int VARIANT>
static int launch_quad(const RealizeArgs& ra, const Plan& p, cudaStream_t st) {
    const int nq = quad_count(ra);
    // realizations per thread from the realization count: a 256-thread CTA carries up to 1024 / 512 / 256 realizations
    if (4 * nq > 2 * RZ_THREADS) return launch_quad_n<VARIANT, 4>(ra, p, nq, st);
    if (4 * nq > RZ_THREADS) return launch_quad_n<VARIANT, 2>(ra, p, nq, st);
    return launch_quad_n<VARIANT, 1>(ra, p, nq, st);
}

template <int VARIANT>
static int launch_realize(const RealizeArgs& ra, const Plan& p, cudaStream_t st) {
    if (flexible_rpt(VARIANT)) {
        if (p.rpt == 4) return launch_realize_rpt<VARIANT, flexible_rpt(VARIANT) ? 4 : rpt_of(VARIANT), RZ_THREADS>(ra, p, st);
        if (p.rpt == 2) return launch_realize_rpt<VARIANT, flexible_rpt(VARIANT) ? 2 : rpt_of(VARIANT), RZ_THREADS>(ra, p, st);
    }
    if (rpt_of(VARIANT) == 1 && p.threads == 128) return launch_realize_rpt<VARIANT, 1, 128>(ra, p, st);
    return launch_realize_rpt<VARIANT, rpt_of(VARIANT), RZ_THREADS>(ra, p, st);
}

}  // namespace holo

using namespace holo;

#ifdef HOLO_PHASE_CLOCKS
extern "C" int holo_debug_phase_clocks(unsigned long long* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, holo::g_phase_clk, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyToSymbol(holo::g_phase_clk, z, sizeof(z));
    }
    return 0;
}
extern "C" int holo_debug_item_phases(unsigned long long* out16384x8, int reset) {
    int rc = (int)cudaMemcpyFromSymbol(out16384x8, holo::g_item_phase, sizeof(unsigned long long) * 16384 * 8);
    if (reset) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, holo::g_item_phase);
        cudaMemset(p, 0, sizeof(unsigned long long) * 16384 * 8);
    }
    return rc;
}
extern "C" int holo_debug_cta_spans(unsigned long long* out4x16384) {
    return (int)cudaMemcpyFromSymbol(out4x16384, holo::g_cta_span, sizeof(unsigned long long) * 4 * 16384);
}
#endif

extern "C" {

int64_t holo_realize_workspace_bytes(int kind, int64_t ncell, int F, int R) {
    if (ncell <= 0 || F <= 0 || R <= 0) return 256;
    int variant = kind == HOLO_REALIZE_SSBG_PAR ? V_SSBG_PAR : (kind == HOLO_REALIZE_SSBG ? V_SSBG : V_GWB);
    Plan p = make_plan(ncell, F, R, variant);
    Layout l = carve(nullptr, variant, ncell, F, R, 0, p);
    return l.total;
}

int64_t holo_loudest_workspace_bytes(int variant, int64_t ncell, int F, int R, int L, int bucket_cap) {
    if (ncell <= 0 || F <= 0 || R <= 0) return 256;
    Plan p = make_plan(ncell, F, R, variant);
    int cap = bucket_cap > 0 ? bucket_cap : auto_cap(L, auto_margin(L));
    Layout l = carve(nullptr, variant, ncell, F, R, cap, p);
    return l.total;
}

}  // extern "C"

namespace holo {
// sam_poisson_gwb on columns [key_col0, key_col0 + F) of a wider grid of `key_cols` columns (key_col0 % 4 == 0):
// the Philox counters use the global column index, so slabs of one grid draw independently.
int realize_gwb_columns(const double* number, const double* h2fdf, int64_t ncell, int F, int R, int64_t r0,
                        uint64_t seed, double normal_threshold, const double* counts, int key_col0, int key_cols,
                        double* gwb, void* workspace, int64_t workspace_bytes, void* stream) {
    HOLO_REQUIRE(number && h2fdf && gwb && workspace, "holo_sam_poisson_gwb: NULL argument");
    HOLO_REQUIRE(ncell > 0 && F > 0 && R > 0, "holo_sam_poisson_gwb: bad shape");
    HOLO_REQUIRE(ncell < 2147483647LL, "holo_sam_poisson_gwb: too many cells");
    cudaStream_t st = (cudaStream_t)stream;
    Plan p = make_plan(ncell, F, R, V_GWB);
    Layout l = carve(workspace, V_GWB, ncell, F, R, 0, p);
    HOLO_REQUIRE(l.total <= workspace_bytes, "holo_sam_poisson_gwb: workspace too small");
    RealizeArgs ra{};
    ra.number = number; ra.h2fdf = h2fdf; ra.counts = counts; ra.partial = l.partial;
    ra.ncell = ncell; ra.chunk = p.chunk; ra.Qb = 1; ra.Zb = 1; ra.F = F; ra.R = R; ra.cap = 0;
    ra.r0 = r0; ra.k0 = (uint32_t)seed; ra.k1 = (uint32_t)(seed >> 32);
    ra.thresh = (double)(int64_t)normal_threshold;                       // `long thresh`, pyx:855, 863
    ra.fg_key0 = key_col0 / FGROUP; ra.f_key0 = key_col0; ra.F_key = key_cols > 0 ? key_cols : F;
    StageTimer timer(st);
    timer.mark();
    int rc = plan_launch_order(number, ncell, F, ra.thresh, nullptr, nullptr, R, p, l, st, uses_quad_kernel(V_GWB) ? 1 : FGROUP);
    if (rc) return rc;
    ra.order = l.order;
    rc = uses_quad_kernel(V_GWB) ? launch_quad<V_GWB>(ra, p, st) : launch_realize<V_GWB>(ra, p, st);
    if (rc) return rc;
    timer.mark();
    FinalArgs fa{};
    fa.partial = l.partial; fa.out0 = gwb; fa.nchunk = p.nchunk; fa.Qb = 1; fa.Zb = 1; fa.F = F; fa.R = R;
    int64_t nfr = (int64_t)F * R;
    final_kernel<V_GWB><<<dim3((R + FIN_R - 1) / FIN_R, F), FIN_R * FIN_SEG, 0, st>>>(fa); holo::count_launches(1);
    timer.mark();
    timer.finish();
    return holo_check_launch("holo_sam_poisson_gwb");
}
}  // namespace holo

extern "C" {

int holo_sam_poisson_gwb(const double* number, const double* h2fdf, int64_t ncell, int F, int R,
                         int64_t r0, uint64_t seed, double normal_threshold, const double* counts,
                         double* gwb, void* workspace, int64_t workspace_bytes, void* stream) {
    return holo::realize_gwb_columns(number, h2fdf, ncell, F, R, r0, seed, normal_threshold, counts, 0, 0, gwb, workspace,
                                     workspace_bytes, stream);
}

int holo_loudest(const holo_loudest_args* g, void* stream) {
    HOLO_REQUIRE(g, "holo_loudest: NULL args");
    const int v = g->variant;
    HOLO_REQUIRE(v == V_LOUD_PLAIN || v == V_LOUD_PAR || v == V_LOUD_PAR_REDZ, "holo_loudest: bad variant");
    HOLO_REQUIRE(g->number && g->h2fdf && g->order && g->hc2ss && g->hc2bg && g->workspace,
                 "holo_loudest: NULL argument");
    HOLO_REQUIRE(g->Mb > 0 && g->Qb > 0 && g->Zb > 0 && g->F > 0 && g->R > 0 && g->L > 0,
                 "holo_loudest: bad shape");
    if (v != V_LOUD_PLAIN) HOLO_REQUIRE(g->mt && g->mr && g->rz && g->bgpar, "holo_loudest: NULL parameter arrays");
    if (v == V_LOUD_PAR) HOLO_REQUIRE(g->lspar && g->ssidx, "holo_loudest: NULL lspar/ssidx");
    if (v == V_LOUD_PAR_REDZ)
        HOLO_REQUIRE(g->redz_final && g->dcom_final && g->sepa && g->angs && g->sspar,
                     "holo_loudest: NULL redz arrays");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncell = (int64_t)g->Mb * g->Qb * g->Zb;
    HOLO_REQUIRE(ncell < 2147483647LL, "holo_loudest: too many cells");
    const int F = g->F, R = g->R, L = g->L;
    // fused realised GWB: Rg extra background-only realization slots share the staged records and tables
    const int Rg = g->gwb ? g->gwb_R : 0;
    HOLO_REQUIRE(Rg >= 0 && (Rg == 0 || !g->counts), "holo_loudest: fused gwb needs gwb_R > 0 and no supplied counts");
    Plan p = make_plan(ncell, F, R + Rg, v);            // slots per CTA / tiles from the total slot count
    double margin = g->head_margin > 0 ? g->head_margin : auto_margin(L);
    int cap = g->bucket_cap > 0 ? g->bucket_cap : auto_cap(L, margin);
    Layout l = carve(g->workspace, v, ncell, F, R, cap, p);
    double* partial_g = nullptr;
    int64_t ws_need = l.total;
    if (Rg > 0) {       // tail of the workspace: holo_loudest_workspace_bytes(..) + holo_realize_workspace_bytes(GWB, .., Rg)
        partial_g = reinterpret_cast<double*>(static_cast<unsigned char*>(g->workspace) + l.total);
        ws_need += align256((int64_t)sizeof(double) * p.nchunk * F * Rg);
    }
    HOLO_REQUIRE(ws_need <= g->workspace_bytes, "holo_loudest: workspace too small");
    size_t res_smem = (size_t)RES_WARPS * 2 * cap * sizeof(Event);
    HOLO_REQUIRE(res_smem <= 200 * 1024, "holo_loudest: bucket_cap too large");

    const int nacc = nacc_of(v);
    const int64_t nfr = (int64_t)F * R;
    HOLO_CUDA(cudaMemsetAsync(l.flags, 0, 4 * sizeof(int32_t), st));
    HOLO_CUDA(cudaMemsetAsync(l.evcount, 0, sizeof(int32_t) * nfr, st));
    HOLO_CUDA(cudaMemsetAsync(g->hc2ss, 0, sizeof(double) * nfr * L, st));
    if (v == V_LOUD_PAR) HOLO_CUDA(cudaMemsetAsync(g->ssidx, 0, sizeof(int64_t) * 3 * nfr * L, st));
    if (v == V_LOUD_PAR_REDZ) HOLO_CUDA(cudaMemsetAsync(g->sspar, 0, sizeof(double) * 4 * nfr * L, st));

    // ---- head preparation
    StageTimer timer(st);
    timer.mark();
    rank_inverse_kernel<<<(int)((ncell + 255) / 256 > 4736 ? 4736 : (ncell + 255) / 256), 256, 0, st>>>(
        g->order, ncell, l.rank); holo::count_launches(1);
    int nblk = (int)((ncell + HEAD_ROWS - 1) / HEAD_ROWS);
    head_sum_kernel<<<nblk, 256, sizeof(double) * F, st>>>(
        g->number, g->h2fdf, g->order, ncell, F, (double)(int64_t)g->normal_threshold,
        v == V_LOUD_PAR_REDZ ? 1 : 0, g->counts ? 1 : 0, l.bsum); holo::count_launches(1);
    head_cut_kernel<<<F, HEAD_CUT_THREADS, 0, st>>>(l.bsum, nblk, F, ncell, (double)L + margin, l.kf); holo::count_launches(1);
    int rc = holo_check_launch("holo_loudest: head preparation");
    if (rc) return rc;

    timer.mark();
    // ---- draws
    RealizeArgs ra{};
    ra.number = g->number; ra.h2fdf = g->h2fdf; ra.rank = l.rank; ra.kf = l.kf;
    ra.mt = g->mt; ra.mr = g->mr; ra.rz = g->rz;
    ra.redz_final = g->redz_final; ra.dcom_final = g->dcom_final; ra.sepa = g->sepa; ra.angs = g->angs;
    ra.counts = g->counts; ra.partial = l.partial; ra.events = l.events; ra.evcount = l.evcount;
    ra.ncell = ncell; ra.chunk = p.chunk; ra.Qb = g->Qb; ra.Zb = g->Zb; ra.F = F; ra.R = R; ra.cap = cap;
    ra.r0 = g->r0; ra.k0 = (uint32_t)g->seed; ra.k1 = (uint32_t)(g->seed >> 32);
    ra.thresh = (double)(int64_t)g->normal_threshold;
    ra.fg_key0 = 0; ra.f_key0 = 0; ra.F_key = F;
    ra.Rg = Rg; ra.r0g = g->gwb_r0; ra.k0g = (uint32_t)g->gwb_seed; ra.k1g = (uint32_t)(g->gwb_seed >> 32);
    ra.partial_g = partial_g;
    rc = plan_launch_order(g->number, ncell, F, ra.thresh, l.rank, l.kf, R, p, l, st, uses_quad_kernel(v) ? 1 : FGROUP);
    if (rc) return rc;
    ra.order = l.order;
    if (v == V_LOUD_PLAIN) rc = uses_quad_kernel(V_LOUD_PLAIN) ? launch_quad<V_LOUD_PLAIN>(ra, p, st) : launch_realize<V_LOUD_PLAIN>(ra, p, st);
    else if (v == V_LOUD_PAR) rc = launch_quad<V_LOUD_PAR>(ra, p, st);
    else rc = launch_quad<V_LOUD_PAR_REDZ>(ra, p, st);
    if (rc) return rc;

    timer.mark();
    // ---- resolve the head, then reduce
    ResolveArgs rs{};
    rs.events = l.events; rs.evcount = l.evcount; rs.kf = l.kf; rs.h2fdf = g->h2fdf;
    rs.mt = g->mt; rs.mr = g->mr; rs.rz = g->rz; rs.redz_final = g->redz_final;
    rs.dcom_final = g->dcom_final; rs.sepa = g->sepa; rs.angs = g->angs;
    rs.hc2ss = g->hc2ss; rs.sspar = g->sspar; rs.lspar = g->lspar; rs.ssidx = g->ssidx;
    rs.rem = l.rem; rs.flags = l.flags; rs.ncell = ncell;
    rs.Qb = g->Qb; rs.Zb = g->Zb; rs.F = F; rs.R = R; rs.L = L; rs.cap = cap;
    int rblocks = (int)((nfr + RES_WARPS - 1) / RES_WARPS);
    if (v == V_LOUD_PLAIN) {
        if (res_smem > 48 * 1024) HOLO_CUDA(cudaFuncSetAttribute(resolve_kernel<V_LOUD_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
        resolve_kernel<V_LOUD_PLAIN><<<rblocks, RES_WARPS * 32, res_smem, st>>>(rs); holo::count_launches(1);
    } else if (v == V_LOUD_PAR) {
        if (res_smem > 48 * 1024) HOLO_CUDA(cudaFuncSetAttribute(resolve_kernel<V_LOUD_PAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
        resolve_kernel<V_LOUD_PAR><<<rblocks, RES_WARPS * 32, res_smem, st>>>(rs); holo::count_launches(1);
    } else {
        if (res_smem > 48 * 1024) HOLO_CUDA(cudaFuncSetAttribute(resolve_kernel<V_LOUD_PAR_REDZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
        resolve_kernel<V_LOUD_PAR_REDZ><<<rblocks, RES_WARPS * 32, res_smem, st>>>(rs); holo::count_launches(1);
    }
    rc = holo_check_launch("holo_loudest: resolve");
    if (rc) return rc;
    timer.mark();

    FinalArgs fa{};
    fa.partial = l.partial; fa.rem = l.rem; fa.out0 = g->hc2bg; fa.bgpar = g->bgpar;
    fa.nchunk = p.nchunk; fa.Qb = g->Qb; fa.Zb = g->Zb; fa.F = F; fa.R = R;
    const dim3 fblocks((R + FIN_R - 1) / FIN_R, F);
    if (v == V_LOUD_PLAIN) final_kernel<V_LOUD_PLAIN><<<fblocks, FIN_R * FIN_SEG, 0, st>>>(fa);
    else if (v == V_LOUD_PAR) final_kernel<V_LOUD_PAR><<<fblocks, FIN_R * FIN_SEG, 0, st>>>(fa);
    else final_kernel<V_LOUD_PAR_REDZ><<<fblocks, FIN_R * FIN_SEG, 0, st>>>(fa);
    holo::count_launches(1);
    if (Rg > 0) {       // the fused background slots: same fixed-order chunk reduction as holo_sam_poisson_gwb
        FinalArgs fg{};
        fg.partial = partial_g; fg.out0 = g->gwb; fg.nchunk = p.nchunk; fg.Qb = 1; fg.Zb = 1; fg.F = F; fg.R = Rg;
        const dim3 gblocks((Rg + FIN_R - 1) / FIN_R, F);
        final_kernel<V_GWB><<<gblocks, FIN_R * FIN_SEG, 0, st>>>(fg); holo::count_launches(1);
    }
    rc = holo_check_launch("holo_loudest: final");
    if (rc) return rc;
    timer.mark();
    timer.finish();

    if (g->defer_check) return HOLO_OK;      // the caller reads flags[0..1] from the head of the workspace later
    int32_t flags[4] = {0, 0, 0, 0};
    HOLO_CUDA(cudaMemcpyAsync(flags, l.flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    HOLO_CUDA(cudaStreamSynchronize(st));
    if (flags[0] || flags[1]) {
        set_error("holo_loudest: %s (bucket_cap=%d, head_margin=%g): retry with larger values",
                  flags[0] ? "event bucket overflow" : "head too short to fill all loudest slots", cap, margin);
        return HOLO_ERR_OVERFLOW;
    }
    return HOLO_OK;
}

int holo_ss_bg_hc(const double* number, const double* h2fdf, int Mb, int Qb, int Zb, int F, int R,
                  int64_t r0, uint64_t seed, double normal_threshold, const double* counts,
                  const double* mt, const double* mr, const double* rz, double* hc2ss, double* hc2bg,
                  int64_t* ssidx, double* bgpar, double* sspar, void* workspace,
                  int64_t workspace_bytes, void* stream) {
    HOLO_REQUIRE(number && h2fdf && hc2ss && hc2bg && ssidx && workspace, "holo_ss_bg_hc: NULL argument");
    HOLO_REQUIRE(Mb > 0 && Qb > 0 && Zb > 0 && F > 0 && R > 0, "holo_ss_bg_hc: bad shape");
    const bool par = (bgpar != nullptr) || (sspar != nullptr);
    if (par) HOLO_REQUIRE(mt && mr && rz && bgpar && sspar, "holo_ss_bg_hc: NULL parameter arrays");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncell = (int64_t)Mb * Qb * Zb;
    HOLO_REQUIRE(ncell < 2147483647LL, "holo_ss_bg_hc: too many cells");
    const int v = par ? V_SSBG_PAR : V_SSBG;
    Plan p = make_plan(ncell, F, R, v);
    Layout l = carve(workspace, v, ncell, F, R, 0, p);
    HOLO_REQUIRE(l.total <= workspace_bytes, "holo_ss_bg_hc: workspace too small");
    HOLO_CUDA(cudaMemsetAsync(l.flags, 0, 4 * sizeof(int32_t), st));
    RealizeArgs ra{};
    ra.number = number; ra.h2fdf = h2fdf; ra.mt = mt; ra.mr = mr; ra.rz = rz; ra.counts = counts;
    ra.partial = l.partial; ra.pmax = l.pmax; ra.pidx = l.pidx;
    ra.ncell = ncell; ra.chunk = p.chunk; ra.Qb = Qb; ra.Zb = Zb; ra.F = F; ra.R = R;
    ra.r0 = r0; ra.k0 = (uint32_t)seed; ra.k1 = (uint32_t)(seed >> 32);
    ra.thresh = (double)(int64_t)normal_threshold;
    ra.fg_key0 = 0; ra.f_key0 = 0; ra.F_key = F;
    int rc = plan_launch_order(number, ncell, F, ra.thresh, nullptr, nullptr, R, p, l, st);
    if (rc) return rc;
    ra.order = l.order;
    rc = par ? launch_realize<V_SSBG_PAR>(ra, p, st) : launch_realize<V_SSBG>(ra, p, st);
    if (rc) return rc;
    FinalArgs fa{};
    fa.partial = l.partial; fa.pmax = l.pmax; fa.pidx = l.pidx; fa.h2fdf = h2fdf;
    fa.mt = mt; fa.mr = mr; fa.rz = rz; fa.out0 = hc2bg; fa.bgpar = bgpar; fa.hc2ss = hc2ss;
    fa.sspar = sspar; fa.ssidx = ssidx; fa.flags = l.flags;
    fa.nchunk = p.nchunk; fa.Qb = Qb; fa.Zb = Zb; fa.F = F; fa.R = R;
    int64_t nfr = (int64_t)F * R;
    const dim3 fblocks((R + FIN_R - 1) / FIN_R, F);
    if (par) final_kernel<V_SSBG_PAR><<<fblocks, FIN_R * FIN_SEG, 0, st>>>(fa);
    else final_kernel<V_SSBG><<<fblocks, FIN_R * FIN_SEG, 0, st>>>(fa);
    holo::count_launches(1);
    rc = holo_check_launch("holo_ss_bg_hc");
    if (rc) return rc;
    if (par) {
        int32_t flags[4] = {0, 0, 0, 0};
        HOLO_CUDA(cudaMemcpyAsync(flags, l.flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        HOLO_CUDA(cudaStreamSynchronize(st));
        if (flags[2]) {
            set_error("holo_ss_bg_hc: no single source found at some (frequency, realization)");
            return HOLO_ERR_OVERFLOW;
        }
    }
    return HOLO_OK;
}

int holo_poisson_as_needed(const double* lam, int64_t n, uint64_t seed, uint64_t stream_id,
                           double normal_threshold, double* out, void* stream) {
    HOLO_REQUIRE(lam && out && n >= 0, "holo_poisson_as_needed: bad argument");
    if (n == 0) return HOLO_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    bulk_poisson_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        lam, n, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)stream_id, normal_threshold, out); holo::count_launches(1);
    return holo_check_launch("holo_poisson_as_needed");
}

}  // extern "C"
