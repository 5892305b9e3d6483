// holo_scatter.cu -- K6: M-Mbulge scatter of the binary density on the device.
//
// Replaces the per-redshift scipy pipeline of `add_scatter_to_masses` (holodeck/sams/sam.py:1291-1394):
//   (1) sp.interpolate.CloughTocher2DInterpolator on the (log10 m1, log10 m2) images of the (mtot, mrat) grid
//       -> values on a regular G x G grid (G = refine*M), NaN / negative values replaced by the
//       NearestNDInterpolator value                                                      sam.py:1362-1380
//   (2) utils._scatter_with_weights along both axes (two dense G x G products)           sam.py:1383-1384
//       -- plain DGEMMs: done with cuBLAS (torch.matmul) by the host driver, not here
//   (3) sp.interpolate.RegularGridInterpolator(method='linear') back to the grid points   sam.py:1387-1389
//
// Everything that depends on the geometry only (Delaunay triangulation, point location, barycentric
// coordinates, nearest vertex, the 2x2 matrices of the gradient estimator, the level schedule) is
// computed once per (mtot, mrat) grid by the host driver (holodeck_b200/sams/scatter.py) and cached; these
// kernels do the data-dependent arithmetic for all Z redshift slices at once.  The arithmetic restates
// scipy/interpolate/interpnd.pyx (`_estimate_gradients_2d_global`, `_clough_tocher_2d_single`; scipy is a
// third-party dependency of the reference, the restatement is pinned against the installed scipy 1.18.1
// by tests/test_scatter.py) operation for operation, so the result agrees with the reference to rounding.
//
// Layouts: density slices are (npts, Z) with z fastest (= the reference's (M, Q, Z) array); gradients are
// (npts, 2, Z); the regular grid is (G, G, Z) with z fastest, so every access below is coalesced along z and
// the two scatter products are single (batched) DGEMMs on contiguous operands.
#include <cuda_runtime.h>

#include <cstdlib>

#include "holo_api.cuh"
#include "holo_common.cuh"

namespace holo {

constexpr int GS_EDGES = 8;                       // neighbours of one vertex handled per step (mean degree ~6)
constexpr int GS_SLOTS = 32;                      // vertices of a level processed per step
constexpr int GS_LANES = 4;                       // threads cooperating on one vertex, two edges each
constexpr int GS_THREADS = GS_SLOTS * GS_LANES;   // consumer threads (a producer warp comes on top)
constexpr int GS_MAX_STAGE = 4;                   // depth of the step ring in shared memory

// -------------------------------------------------------------------------------------------------
// K6a: gradients at the triangulation vertices by scipy's Gauss-Seidel sweeps (interpnd.pyx,
// `_estimate_gradients_2d_global`, maxiter = 400, tol = 1e-6), one CTA per redshift slice.
//
// A sweep visits the vertices in index order and uses the freshest neighbour gradients.  Vertices that are
// not connected do not interact within a sweep, so the sweep is executed level by level of the dependency
// graph (level(v) = 1 + max level of its lower-numbered neighbours): every vertex still sees exactly the
// values the sequential sweep would show it -- same iterates up to rounding, same iteration count -- but a
// level's vertices run in parallel, and the neighbours of a vertex are summed by GS_LANES cooperating lanes
// (two neighbours each).
//
// The sweep is a chain of ~450 dependent steps (441 levels at the named grid), 1..9 sweeps per slice: what a step
// costs is latency.  In round 1 that was four dependent L2 round trips per step (order -> indptr -> indices -> edge),
// 5.3 ms per launch.  All of that is geometry, so the host flattens it into a STEP PROGRAM -- one fixed-size record
// per step -- which the kernel streams through a ring of shared-memory stages with TMA bulk copies (`cp.async.bulk`
// completing on an mbarrier).  How a step is organised now, with what each change bought at the named grid (9 sweeps):
//   * records streamed by TMA, one thread issuing between steps, 8 lanes per vertex             5.3  -> 1.93 ms
//   * a PRODUCER WARP owns the ring: it polls a shared counter of consumed records (published by consumer thread 0
//     after each step's barrier), issues the copies as far ahead as the ring allows and drains them at the end.  The
//     issue path (mbarrier.arrive.expect_tx + cp.async.bulk, ~280 cycles of one warp) left the step's critical path;
//   * each consumer copies ITS fields of the next record into registers one step ahead -- plain loads with no
//     dependants, issued while the current step's sums are in flight -- so a step starts at the gather of the
//     neighbour gradients instead of record -> neighbour index -> gradient; the ring's mbarrier is polled
//     (test_wait, ~150 cycles to answer) at the top of the step and its answer used after the sums          -> 1.78 ms
//   * FOUR lanes per vertex with two edges each (same association as eight lanes with an xor-butterfly: neighbour l
//     with l + 4, then xor 2, xor 1): one shuffle level less, 4 consumer warps instead of 8 at the barrier; the edge
//     terms are branch-free (an absent edge points at the vertex itself with zero coefficients), so a thread's two
//     edges are independent fp64 chains that overlap; thread index, barrier and ring addresses live in registers
//     as 32-bit shared addresses (ptxas re-derived them from S2R SR_TID / SR_CgaCtaId inside the step)      -> 1.54 ms
//   * the two record sets swap roles every step (no register copies); the records hold 2 ex, 2 ey and MINUS the
//     inverse normal matrix (exact rescalings: two fp64 levels less)                                         -> 1.24 ms
//   * scipy's convergence measure (max over vertices of change / max(1, |r0|, |r1|)) is tracked as a fraction per lane,
//     divided once per sweep, and the bookkeeping of a step's update is done branch-free at the top of the NEXT step,
//     interleaved with its edge terms, instead of between the store and the barrier                          -> 1.05 ms
//   * plane-ordered records: a thread's fields lie at 16 t of each plane (seven conflict-free 16-byte loads off one
//     base register instead of ten loads with 2-way conflicts)                                               -> 1.01 ms
// What bounds a step now (~490 cycles) is one warp per scheduler issuing ~130 dependent instructions: measured on
// this GPU a dependent DADD/DMUL costs 8.4 cycles, a 64-bit shuffle + DADD 35.5, a 128-thread named barrier 33, store
// -> barrier -> dependent load 74 (build/ubench in the notes).  [Tried and slower: one warp per slice with __syncwarp
// between levels (2.3 ms); eight lanes per vertex with the producer warp (1.9 ms).  Ring depth 2 / 3 / 4: identical,
// no step ever finds its record not landed.]
// -------------------------------------------------------------------------------------------------
// Debug instrumentation (profiling build only, `make phase`): thread 0 of every CTA adds the cycles it spends in each
// segment of a step to g_gs_clk (read with holo_debug_gs_clocks; profiles/gs_clocks.py).
#ifdef HOLO_GS_CLOCKS
__device__ unsigned long long g_gs_clk[8];

#define GS_CLK_DECL long long gs_last = clock64();
#define GS_CLK_MARK(i)                                                              \
    if (tid == 0) {                                                                 \
        const long long gs_now = clock64();                                         \
        atomicAdd(&g_gs_clk[i], (unsigned long long)(gs_now - gs_last));            \
        gs_last = gs_now;                                                           \
    }
#else
#define GS_CLK_DECL
#define GS_CLK_MARK(i)
#endif

struct GsRec {                       // one step: <= GS_SLOTS vertices of one level, <= GS_EDGES neighbours each.
    // Plane-ordered: what consumer thread t = 4 slot + sub needs lies at 16 t of each plane (conflict-free 16-byte
    // loads off one base register).  The thread owns edges `sub` (a) and `sub + 4` (b) of the vertex in `slot`.
    double e[4][GS_THREADS][2];      // a: (2 ex, 2 ey), a: (ex, ey)/L^3, b: (2 ex, 2 ey), b: (ex, ey)/L^3
    double qinv[2][GS_SLOTS][2];     // the two rows of MINUS the inverse of the vertex's 2x2 normal matrix
    int ids[GS_THREADS][4];          // a's neighbour, b's neighbour, the vertex (-1: empty slot), step flags
                                     // (an absent edge: the vertex itself -- vertex 0 in an empty slot -- with e = 0;
                                     //  flags bit 0: first round of these vertices, bit 1: last round, apply the update)
};
static_assert(sizeof(GsRec) % 16 == 0, "bulk copies move multiples of 16 bytes");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "HOLO_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra HOLO_MBAR_DONE;\n\t"
        "bra HOLO_MBAR_WAIT;\n\t"
        "HOLO_MBAR_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "HOLO_MBAR_WAIT_A:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra HOLO_MBAR_DONE_A;\n\t"
        "bra HOLO_MBAR_WAIT_A;\n\t"
        "HOLO_MBAR_DONE_A:\n\t"
        "}" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ bool mbar_test_addr(uint32_t bar, uint32_t phase) {      // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t phase) {      // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// What a consumer thread needs from a step record, copied into registers ONE STEP AHEAD (plain loads with no
// dependants, issued while the current step's sums are in flight), so that a step's critical path starts at the
// gather of the neighbour gradients instead of at record -> neighbour index -> gradient.
struct GsFields {
    int flags, ip, ipa, ipb;      // step flags, the vertex, the neighbours of this thread's two edges
    double2 a01, a23, b01, b23;   // 2 ex, 2 ey | ex/L^3, ey/L^3 of edges `sub` and `sub + 4`
    double2 qa, qb;               // minus the inverse normal matrix of the vertex
};

__device__ __forceinline__ void gs_consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GS_THREADS) : "memory"); }

// shared-memory accesses by 32-bit shared-window address (kept in registers: ptxas otherwise rebuilds the window base
// from SR_CgaCtaId -- an S2R of tens of cycles -- inside the step)
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_d(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_d2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_i(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i(uint32_t a, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// one edge's contribution  (6 (f1 - f2) - 2 df2) (ex, ey) / L^3  to the vertex's right-hand side, or nothing.
// Branch-free on purpose: the two edges of a thread are independent fp64 chains that must overlap.
__device__ __forceinline__ double2 gs_edge(double2 e01, double2 e23, double2 y, double f1, double f2) {
    const double df2x2 = -e01.x * y.x - e01.y * y.y;
    const double num = 6 * (f1 - f2) - df2x2;
    return make_double2(num * e23.x, num * e23.y);
}

__global__ void __launch_bounds__(GS_THREADS + 32)
ct_gradients_kernel(int npts, int Z, const GsRec* __restrict__ prog, int nsteps, int nstage,
                    const double* __restrict__ data /* (npts, Z) */, int maxiter, double tol,
                    double* __restrict__ grad /* (npts, 2, Z) */, int* __restrict__ niter) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    GsRec* ring = reinterpret_cast<GsRec*>(s_raw);                                   // (nstage)
    double* s_f = reinterpret_cast<double*>(s_raw + (size_t)nstage * sizeof(GsRec)); // (npts)     data of this slice
    double* s_y = s_f + ((npts + 1) & ~1);                                           // (npts, 2)  current gradients
    __shared__ uint64_t s_full[GS_MAX_STAGE];
    __shared__ double s_err[GS_THREADS / 32];
    __shared__ int s_done;
    __shared__ int s_loaded;       // records every consumer has copied into registers: their ring stages are free
    const int z = blockIdx.x;
    // (thread index and barrier address through opaque moves: ptxas would otherwise re-read %tid / the CTA's shared
    //  window base -- S2R / S2UR, tens of cycles each -- several times per step)
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    uint32_t full0, ring0, loaded0;      // (declared before s_f / s_y are initialised below; addresses only)
    asm volatile("mov.u32 %0, %1;" : "=r"(full0) : "r"(smem_u32(&s_full[0])));
    asm volatile("mov.u32 %0, %1;" : "=r"(ring0) : "r"(smem_u32(ring)));
    asm volatile("mov.u32 %0, %1;" : "=r"(loaded0) : "r"(smem_u32(&s_loaded)));
    uint32_t f0, y0;
    asm volatile("mov.u32 %0, %1;" : "=r"(f0) : "r"(smem_u32(s_f)));
    asm volatile("mov.u32 %0, %1;" : "=r"(y0) : "r"(smem_u32(s_y)));
    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) mbar_init(&s_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_done = 0;
        s_loaded = 0;
    }
    for (int i = tid; i < npts; i += GS_THREADS + 32) {
        s_f[i] = data[(int64_t)i * Z + z];
        s_y[2 * i] = 0.0;
        s_y[2 * i + 1] = 0.0;
    }
    __syncthreads();

    if (tid >= GS_THREADS) {
        // ---- producer warp: one lane streams the step program through the ring, as far ahead as the ring allows ----
        if (tid != GS_THREADS) return;
        const long long total = (long long)maxiter * nsteps;      // (speculatively into the next sweep: it stops at s_done)
        long long p = 0;                                           // records issued
        int pst = 0, prec = 0;
        volatile int* v_loaded = &s_loaded;
        volatile int* v_done = &s_done;
        bool stop = false;
        while (p < total && !stop) {
            while (p >= (long long)*v_loaded + nstage) {          // stage `pst` still holds a record somebody reads
                if (*v_done) { stop = true; break; }
                __nanosleep(32);
            }
            if (stop || *v_done) break;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // consumers' reads before the async write
            mbar_expect_tx(&s_full[pst], (uint32_t)sizeof(GsRec));
            bulk_g2s(&ring[pst], &prog[prec], (uint32_t)sizeof(GsRec), &s_full[pst]);
            if (++pst == nstage) pst = 0;
            if (++prec == nsteps) prec = 0;
            ++p;
        }
        // every copy issued must have landed before the CTA's shared memory goes away: the last fill of each stage
        const long long first = p > nstage ? p - nstage : 0;
        for (long long q = first; q < p; ++q) mbar_wait(&s_full[(int)(q % nstage)], (uint32_t)((q / nstage) & 1));
        return;
    }

    // ---- consumers: 4 threads per vertex, two edges each ----
    const int sub = tid % GS_LANES, slot = tid / GS_LANES;
    int st = 0, nloaded = 0;
    uint32_t ph = 0;
    // shared addresses of this thread's fields in stage `st`: tbase + plane offset, qbase + row offset
    const uint32_t t_off = 16u * (uint32_t)tid, q_off = (uint32_t)offsetof(GsRec, qinv) + 16u * (uint32_t)slot;
    constexpr uint32_t PLANE = 16u * GS_THREADS, IDS = (uint32_t)offsetof(GsRec, ids);
    uint32_t rbase = ring0;                   // shared address of stage `st`
    auto fetch = [&](GsFields& n, bool ready) {       // copy this thread's part of the next record into registers
        if (!ready) mbar_wait_addr(full0 + 8u * (uint32_t)st, ph);
        const uint32_t tb = rbase + t_off, qb = rbase + q_off;
        const int4 ids = lds_i4(tb + IDS);
        n.ipa = ids.x; n.ipb = ids.y; n.ip = ids.z; n.flags = ids.w;
        n.a01 = lds_d2(tb);
        n.a23 = lds_d2(tb + PLANE);
        n.b01 = lds_d2(tb + 2u * PLANE);
        n.b23 = lds_d2(tb + 3u * PLANE);
        n.qa = lds_d2(qb);
        n.qb = lds_d2(qb + 16u * GS_SLOTS);
        rbase += (uint32_t)sizeof(GsRec);
        if (++st == nstage) { st = 0; ph ^= 1u; rbase = ring0; }
    };
    GsFields fa_, fb_;                         // the record being processed and the one being fetched swap roles
    fetch(fa_, false);
    gs_consumer_barrier();
    if (tid == 0) sts_i(loaded0, ++nloaded);
    int converged = 0;
    GS_CLK_DECL
    // scipy's convergence measure is max over vertices of change / max(1, |r0|, |r1|): each updating lane tracks its
    // largest quotient as a fraction bc / bd (compared by cross-multiplication) and divides once per sweep.  The
    // bookkeeping of a step's update (r0p, r1p replacing oyp) is done at the top of the NEXT step, branch-free in
    // the same basic block as the edge terms, so that it fills their latency instead of delaying the barrier.
    double sa0, sa1, sb0, sb1, bc, bd, r0p, r1p;
    double2 oyp;
    bool updp;
    auto account = [&]() {
        const double pc = fmax(fabs(oyp.x - r0p), fabs(oyp.y - r1p));
        const double pd = fmax(1.0, fmax(fabs(r0p), fabs(r1p)));
        const bool gt = updp && (pc * bd > bc * pd);                            // change/den > bc/bd
        bc = gt ? pc : bc;
        bd = gt ? pd : bd;
    };
    int it = 0;
    // one step: `c` is processed, `n` receives the following record.  Critical path: neighbour gradients -> edge terms
    // -> 4-lane sum -> 2x2 solve -> store -> barrier
    auto step = [&](const GsFields& c, GsFields& n, bool more) {
        const uint32_t jv = (uint32_t)max(c.ip, 0);
        // (an absent edge points at the vertex itself with zero coefficients: it adds exactly nothing)
        const double2 ya = lds_d2(y0 + 16u * (uint32_t)c.ipa), yb = lds_d2(y0 + 16u * (uint32_t)c.ipb);
        const double f1 = lds_d(f0 + 8u * jv), fa = lds_d(f0 + 8u * (uint32_t)c.ipa), fb = lds_d(f0 + 8u * (uint32_t)c.ipb);
        const double2 oy = lds_d2(y0 + 16u * jv);
        const bool ready = more ? mbar_test_addr(full0 + 8u * (uint32_t)st, ph) : false;   // asked now, needed later
        GS_CLK_MARK(0)
        account();
        const double2 pa = gs_edge(c.a01, c.a23, ya, f1, fa), pb = gs_edge(c.b01, c.b23, yb, f1, fb);
        if (c.flags & 1) {                  // (block-uniform) first round of these vertices: the sums start here
            sa0 = pa.x; sa1 = pa.y; sb0 = pb.x; sb1 = pb.y;
        } else {                            // a vertex with more than 8 neighbours adds up over several steps
            sa0 += pa.x; sa1 += pa.y; sb0 += pb.x; sb1 += pb.y;
        }
        GS_CLK_MARK(1)
        // under the sums' latency: the next record's fields (speculatively the next sweep's first) -- independent loads
#ifdef HOLO_GS_CLOCKS
        if (tid == 0 && more && !ready) atomicAdd(&g_gs_clk[6], 1ull);          // the next record had not landed yet
#endif
        if (more) fetch(n, ready);
        GS_CLK_MARK(2)
        updp = false;
        if (c.flags & 2) {                                      // (block-uniform: the shuffles are full-warp)
            // same association as eight lanes with an xor-butterfly: (s_l + s_{l+4}), then xor 2, xor 1
            double t0 = sa0 + sb0, t1 = sa1 + sb1;
#pragma unroll
            for (int off = GS_LANES / 2; off > 0; off >>= 1) {
                t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                t1 += __shfl_xor_sync(0xffffffffu, t1, off);
            }
            r0p = c.qa.x * t0 + c.qa.y * t1;                                   // = -(Q^-1 s): the new gradient
            r1p = c.qb.x * t0 + c.qb.y * t1;                                   // (every lane of the group has the sums)
            oyp = oy;
            updp = (sub == 0) && (c.ip >= 0);
            if (updp) sts_d2(y0 + 16u * jv, r0p, r1p);
        }
        GS_CLK_MARK(3)
        gs_consumer_barrier();      // the step's gradients are visible
        if (tid == 0) sts_i(loaded0, ++nloaded);                               // ... and the next record's stage is free
        GS_CLK_MARK(4)
    };
    for (; it < maxiter; ++it) {
        sa0 = sa1 = sb0 = sb1 = 0.0;
        bc = 0.0; bd = 1.0; r0p = r1p = 0.0; oyp = make_double2(0.0, 0.0); updp = false;
        const bool again = it + 1 < maxiter;
        int s = 0;
        for (; s + 2 < nsteps; s += 2) {
            step(fa_, fb_, true);
            step(fb_, fa_, true);
        }
        if (s + 2 == nsteps) {
            step(fa_, fb_, true);
            step(fb_, fa_, again);
        } else {                                                // odd step count: one register copy per sweep
            step(fa_, fb_, again);
            fa_ = fb_;
        }
        account();                                              // the last step's update
        double err = bc / bd;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, off));
        if ((tid & 31) == 0) s_err[tid >> 5] = err;
        gs_consumer_barrier();
        if (tid == 0) {
            double e = s_err[0];
            for (int w = 1; w < GS_THREADS / 32; ++w) e = fmax(e, s_err[w]);
            if (e < tol) *(volatile int*)&s_done = it + 1;
        }
        gs_consumer_barrier();
        GS_CLK_MARK(5)
        if (*(volatile int*)&s_done) { converged = s_done; break; }
    }
    if (tid == 0 && !converged) *(volatile int*)&s_done = -1;      // maxiter sweeps done: the producer stops as well
    for (int i = tid; i < npts; i += GS_THREADS) {
        grad[((int64_t)i * 2) * Z + z] = s_y[2 * i];
        grad[((int64_t)i * 2 + 1) * Z + z] = s_y[2 * i + 1];
    }
    if (tid == 0 && niter) niter[z] = converged;   // 0: not converged within maxiter (scipy warns and goes on)
}

// -------------------------------------------------------------------------------------------------
// K6b: Clough-Tocher values on the regular grid + nearest-vertex fill of bad values.
// One CTA per grid point, threads over z.  `_clough_tocher_2d_single` of interpnd.pyx.
// -------------------------------------------------------------------------------------------------
struct CtPoint {       // geometry of one regular-grid point (host-computed)
    int simplex;       // -1: outside the convex hull (scipy returns NaN -> nearest fill)
    int nearest;       // index of the nearest data point (NearestNDInterpolator)
    int v[3];          // vertices of the simplex
    int pad;
    double b[3];       // barycentric coordinates
    double e[6];       // e12x, e12y, e23x, e23y, e31x, e31y
    double g[3];       // affine-invariant edge parameters (neighbour centroids)
};

// Flat over (point, z), z fastest: consecutive threads write consecutive doubles, no lane idles on Z = 101, and the 88 %
// of the regular grid that lies outside the hull (nearest fill) reads 8 bytes of its point's geometry, not the record.
__global__ void __launch_bounds__(256)
ct_eval_kernel(int64_t ngrid, int Z, const CtPoint* __restrict__ geo, const double* __restrict__ data /* (npts, Z) */,
               const double* __restrict__ grad /* (npts, 2, Z) */, double* __restrict__ out /* (ngrid, Z) */,
               int* __restrict__ flags /* [0]: a bad value survived the nearest fill */) {
    const uint32_t total = (uint32_t)(ngrid * Z), stride = gridDim.x * blockDim.x;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const uint32_t p = idx / (uint32_t)Z;
        const int z = (int)(idx - p * (uint32_t)Z);
        const int2 head = *reinterpret_cast<const int2*>(&geo[p]);      // simplex, nearest
        CtPoint gp;
        gp.simplex = head.x;
        gp.nearest = head.y;
        if (gp.simplex >= 0) gp = geo[p];
        double w = nan("");
        if (gp.simplex >= 0) {
            const double f1 = data[(int64_t)gp.v[0] * Z + z], f2 = data[(int64_t)gp.v[1] * Z + z],
                         f3 = data[(int64_t)gp.v[2] * Z + z];
            const double d0x = grad[((int64_t)gp.v[0] * 2) * Z + z], d0y = grad[((int64_t)gp.v[0] * 2 + 1) * Z + z];
            const double d1x = grad[((int64_t)gp.v[1] * 2) * Z + z], d1y = grad[((int64_t)gp.v[1] * 2 + 1) * Z + z];
            const double d2x = grad[((int64_t)gp.v[2] * 2) * Z + z], d2y = grad[((int64_t)gp.v[2] * 2 + 1) * Z + z];
            const double e12x = gp.e[0], e12y = gp.e[1], e23x = gp.e[2], e23y = gp.e[3], e31x = gp.e[4], e31y = gp.e[5];
            const double df12 = +(d0x * e12x + d0y * e12y);
            const double df21 = -(d1x * e12x + d1y * e12y);
            const double df23 = +(d1x * e23x + d1y * e23y);
            const double df32 = -(d2x * e23x + d2y * e23y);
            const double df31 = +(d2x * e31x + d2y * e31y);
            const double df13 = -(d0x * e31x + d0y * e31y);
            const double c3000 = f1;
            const double c2100 = (df12 + 3 * c3000) / 3;
            const double c2010 = (df13 + 3 * c3000) / 3;
            const double c0300 = f2;
            const double c1200 = (df21 + 3 * c0300) / 3;
            const double c0210 = (df23 + 3 * c0300) / 3;
            const double c0030 = f3;
            const double c1020 = (df31 + 3 * c0030) / 3;
            const double c0120 = (df32 + 3 * c0030) / 3;
            const double c2001 = (c2100 + c2010 + c3000) / 3;
            const double c0201 = (c1200 + c0300 + c0210) / 3;
            const double c0021 = (c1020 + c0120 + c0030) / 3;
            const double c0111 = (gp.g[0] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) +
                                  (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2;
            const double c1011 = (gp.g[1] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) +
                                  (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2;
            const double c1101 = (gp.g[2] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) +
                                  (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2;
            const double c1002 = (c1101 + c1011 + c2001) / 3;
            const double c0102 = (c1101 + c0111 + c0201) / 3;
            const double c0012 = (c1011 + c0111 + c0021) / 3;
            const double c0003 = (c1002 + c0102 + c0012) / 3;
            double minval = gp.b[0];
            if (gp.b[1] < minval) minval = gp.b[1];
            if (gp.b[2] < minval) minval = gp.b[2];
            const double b1 = gp.b[0] - minval, b2 = gp.b[1] - minval, b3 = gp.b[2] - minval, b4 = 3 * minval;
            w = (b1 * b1 * b1 * c3000 + 3 * b1 * b1 * b2 * c2100 + 3 * b1 * b1 * b3 * c2010 + 3 * b1 * b1 * b4 * c2001 +
                 3 * b1 * b2 * b2 * c1200 + 6 * b1 * b2 * b4 * c1101 + 3 * b1 * b3 * b3 * c1020 + 6 * b1 * b3 * b4 * c1011 +
                 3 * b1 * b4 * b4 * c1002 + b2 * b2 * b2 * c0300 + 3 * b2 * b2 * b3 * c0210 + 3 * b2 * b2 * b4 * c0201 +
                 3 * b2 * b3 * b3 * c0120 + 6 * b2 * b3 * b4 * c0111 + 3 * b2 * b4 * b4 * c0102 + b3 * b3 * b3 * c0030 +
                 3 * b3 * b3 * b4 * c0021 + 3 * b3 * b4 * b4 * c0012 + b4 * b4 * b4 * c0003);
        }
        if (isnan(w) || w < 0.0) {                                        // sam.py:1370-1375
            w = data[(int64_t)gp.nearest * Z + z];
            if (isnan(w) || w < 0.0) atomicOr(&flags[0], 1);              // sam.py:1376-1380 raises
        }
        out[idx] = w;
    }
}

// -------------------------------------------------------------------------------------------------
// K6c: bilinear interpolation from the regular grid back to the data points (RegularGridInterpolator,
// method='linear').  One CTA per data point, threads over z.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bilinear_back_kernel(int npts, int G, int Z, const int* __restrict__ i0, const int* __restrict__ i1,
                     const double* __restrict__ y0, const double* __restrict__ y1,
                     const double* __restrict__ grid /* (G, G, Z) */, double* __restrict__ out /* (npts, Z) */) {
    const int p = blockIdx.x;
    const int a = i0[p], b = i1[p];
    const double u = y0[p], v = y1[p];
    const int64_t o00 = ((int64_t)a * G + b) * Z, o01 = o00 + Z, o10 = o00 + (int64_t)G * Z, o11 = o10 + Z;
    for (int z = threadIdx.x; z < Z; z += blockDim.x) {
        out[(int64_t)p * Z + z] = grid[o00 + z] * (1 - u) * (1 - v) + grid[o01 + z] * (1 - u) * v +
                                  grid[o10 + z] * u * (1 - v) + grid[o11 + z] * u * v;
    }
}

}  // namespace holo

using namespace holo;

extern "C" {

int holo_scatter_step_bytes(void) { return (int)sizeof(GsRec); }

int holo_scatter_gradients(int npts, int Z, const void* program, int nsteps, const double* data, int maxiter, double tol,
                           double* grad, int* niter, void* stream) {
    HOLO_REQUIRE(program && data && grad, "holo_scatter_gradients: NULL argument");
    HOLO_REQUIRE(npts > 0 && Z > 0 && nsteps > 0 && maxiter > 0, "holo_scatter_gradients: bad shape");
    HOLO_REQUIRE((reinterpret_cast<uintptr_t>(program) & 15) == 0, "holo_scatter_gradients: program must be 16-byte aligned");
    // the deepest step ring that still fits next to the slice's data and gradients
    const size_t fixed = sizeof(double) * (3 * (size_t)npts + 1);   // data (padded to even) + gradients
    int nstage = GS_MAX_STAGE;
#ifdef HOLO_GS_CLOCKS
    if (const char* env = getenv("HOLO_GS_STAGES")) nstage = atoi(env) < GS_MAX_STAGE ? atoi(env) : GS_MAX_STAGE;
#endif
    while (nstage > 2 && fixed + (size_t)nstage * sizeof(GsRec) > 226 * 1024) --nstage;
    const size_t smem = fixed + (size_t)nstage * sizeof(GsRec);
    HOLO_REQUIRE(smem <= 226 * 1024, "holo_scatter_gradients: too many grid points for shared memory (M*Q <= 8700)");
    HOLO_CUDA(cudaFuncSetAttribute(ct_gradients_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ct_gradients_kernel<<<Z, GS_THREADS + 32, smem, (cudaStream_t)stream>>>(npts, Z, (const GsRec*)program, nsteps, nstage, data, maxiter,
                                                                        tol, grad, niter);
    holo::count_launches(1);
    return holo_check_launch("holo_scatter_gradients");
}

int holo_scatter_ct_eval(int64_t ngrid, int Z, const void* geo, const double* data, const double* grad, double* out,
                         int* flags, void* stream) {
    HOLO_REQUIRE(geo && data && grad && out && flags, "holo_scatter_ct_eval: NULL argument");
    HOLO_REQUIRE(ngrid > 0 && Z > 0 && ngrid * Z < 4294967295LL, "holo_scatter_ct_eval: bad shape");
    HOLO_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), (cudaStream_t)stream));
    const int64_t want = (ngrid * Z + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);        // a few resident waves, grid-stride
    ct_eval_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ngrid, Z, (const CtPoint*)geo, data, grad, out, flags);
    holo::count_launches(1);
    return holo_check_launch("holo_scatter_ct_eval");
}

int holo_scatter_bilinear(int npts, int G, int Z, const int* i0, const int* i1, const double* y0, const double* y1,
                          const double* grid, double* out, void* stream) {
    HOLO_REQUIRE(i0 && i1 && y0 && y1 && grid && out, "holo_scatter_bilinear: NULL argument");
    HOLO_REQUIRE(npts > 0 && G > 1 && Z > 0, "holo_scatter_bilinear: bad shape");
    bilinear_back_kernel<<<npts, 128, 0, (cudaStream_t)stream>>>(npts, G, Z, i0, i1, y0, y1, grid, out);
    holo::count_launches(1);
    return holo_check_launch("holo_scatter_bilinear");
}

#ifdef HOLO_GS_CLOCKS
int holo_debug_gs_clocks(unsigned long long* out, int reset) {
    HOLO_CUDA(cudaMemcpyFromSymbol(out, holo::g_gs_clk, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        HOLO_CUDA(cudaMemcpyToSymbol(holo::g_gs_clk, z, sizeof(z)));
    }
    return 0;
}
#endif

int holo_scatter_geo_bytes(void) { return (int)sizeof(CtPoint); }

}  // extern "C"
