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

#include "holo_api.cuh"
#include "holo_common.cuh"

namespace holo {

constexpr int GS_LANES = 8;                       // lanes cooperating on one vertex (mean degree ~6)
constexpr int GS_THREADS = 256;
constexpr int GS_SLOTS = GS_THREADS / GS_LANES;   // vertices of a level processed per step
constexpr int GS_MAX_STAGE = 4;                   // depth of the step ring in shared memory

// -------------------------------------------------------------------------------------------------
// K6a: gradients at the triangulation vertices by scipy's Gauss-Seidel sweeps (interpnd.pyx,
// `_estimate_gradients_2d_global`, maxiter = 400, tol = 1e-6), one CTA per redshift slice.
//
// A sweep visits the vertices in index order and uses the freshest neighbour gradients.  Vertices that are
// not connected do not interact within a sweep, so the sweep is executed level by level of the dependency
// graph (level(v) = 1 + max level of its lower-numbered neighbours): every vertex still sees exactly the
// values the sequential sweep would show it -- same iterates up to rounding, same iteration count -- but a
// level's vertices run in parallel, and the neighbours of a vertex are summed by GS_LANES cooperating lanes.
//
// The sweep is a chain of ~450 dependent steps (441 levels at the named grid), 1..9 sweeps per slice: what a step
// costs is latency, and in round 1 that latency was four dependent L2 round trips per step (order -> indptr ->
// indices -> edge), 5.3 ms per launch.  All of that is geometry.  The host now flattens it into a STEP PROGRAM:
// one fixed-size record per step holding, per thread, its neighbour index and edge quotients and, per vertex slot,
// the vertex index and its inverse normal matrix.  A record is one contiguous 10 KB block, so the kernel streams the
// program through a ring of shared-memory stages with TMA bulk copies (`cp.async.bulk` completing on an mbarrier),
// issued GS_MAX_STAGE-1 steps ahead by one thread: a step then touches shared memory only.  The barrier of the NEXT
// stage is polled (non-blocking) while the current step computes, and the division of scipy's convergence measure
// is taken once per sweep (each lane tracks its largest change/scale as a fraction).  Measured on B200 at the named
// grid (9 sweeps): 5.3 ms -> 2.0 ms.  [A one-warp-per-slice variant -- one thread per vertex, no shuffles, __syncwarp
// between levels -- was slower (2.3 ms): one warp keeps a single scheduler's FP64 pipe busy, four share the work.]
// -------------------------------------------------------------------------------------------------
struct GsRec {                       // one step: <= GS_SLOTS vertices of one level, <= GS_LANES neighbours each
    double e[GS_THREADS][4];         // per thread: ex, ey, ex/L^3, ey/L^3 of its edge
    double qinv[GS_SLOTS][4];        // per vertex slot: inverse of the 2x2 normal matrix, row-major
    int nb[GS_THREADS];              // per thread: neighbour vertex, -1 = none
    int vip[GS_SLOTS];               // per vertex slot: vertex, -1 = none
    int hdr[4];                      // [0] bit 0: first round of its vertices (clear the sums), bit 1: last round (update)
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

__global__ void __launch_bounds__(GS_THREADS)
ct_gradients_kernel(int npts, int Z, const GsRec* __restrict__ prog, int nsteps, int nstage,
                    const double* __restrict__ data /* (npts, Z) */, int maxiter, double tol,
                    double* __restrict__ grad /* (npts, 2, Z) */, int* __restrict__ niter) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    GsRec* ring = reinterpret_cast<GsRec*>(s_raw);                                   // (nstage)
    double* s_f = reinterpret_cast<double*>(s_raw + (size_t)nstage * sizeof(GsRec)); // (npts)     data of this slice
    double* s_y = s_f + npts;                                                        // (npts, 2)  current gradients
    __shared__ uint64_t s_full[GS_MAX_STAGE];
    __shared__ double s_err[GS_THREADS / 32];
    __shared__ int s_done;
    const int z = blockIdx.x, tid = threadIdx.x;
    const int sub = tid % GS_LANES, slot = tid / GS_LANES;
    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) mbar_init(&s_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_done = 0;
    }
    for (int i = tid; i < npts; i += GS_THREADS) {
        s_f[i] = data[(int64_t)i * Z + z];
        s_y[2 * i] = 0.0;
        s_y[2 * i + 1] = 0.0;
    }
    __syncthreads();
    // Ring bookkeeping without divisions: the consumer walks (stage `st`, parity `ph`), the producer -- nstage-1 steps
    // ahead -- (stage `pst`, program record `prec`); `ahead` counts the copies in flight beyond the consumer.
    int st = 0, pst = 0, prec = 0, ahead = 0;
    uint32_t ph = 0;
    int psweeps = maxiter;                                     // sweeps the producer may still stream
    auto issue = [&]() {                                       // (thread 0) stream the next record into stage `pst`
        mbar_expect_tx(&s_full[pst], (uint32_t)sizeof(GsRec));
        bulk_g2s(&ring[pst], &prog[prec], (uint32_t)sizeof(GsRec), &s_full[pst]);
        if (++pst == nstage) pst = 0;
        if (++prec == nsteps) { prec = 0; --psweeps; }
        ++ahead;
    };
    if (tid == 0)
        while (ahead < nstage - 1 && psweeps > 0) issue();
    int converged = 0;
    bool ready = false;            // has the barrier of the stage about to be consumed been seen complete already?
    for (int it = 0; it < maxiter; ++it) {
        double s0 = 0.0, s1 = 0.0;
        // scipy's convergence measure is max over vertices of change / max(1, |r0|, |r1|): each updating lane tracks
        // its largest quotient as a fraction (compared by cross-multiplication) and divides once per sweep
        double bc = 0.0, bd = 1.0;
        for (int s = 0; s < nsteps; ++s) {
            // the stage of the previous step was released by the barrier that ended it: refill it
            if (tid == 0 && psweeps > 0) issue();
            if (!ready) mbar_wait(&s_full[st], ph);
            const GsRec& r = ring[st];
            if (++st == nstage) { st = 0; ph ^= 1u; }
            if (tid == 0) --ahead;
            // poll the NEXT stage now: the answer is back long before the next step asks for it
            ready = ((s + 1 < nsteps) || (it + 1 < maxiter)) ? mbar_test(&s_full[st], ph) : false;
            const int flags = r.hdr[0];
            if (flags & 1) { s0 = 0.0; s1 = 0.0; }
            const int ip = r.vip[slot];
            const int ip2 = r.nb[tid];
            if (ip2 >= 0) {
                const double f1 = s_f[ip];
                const double2 e01 = *reinterpret_cast<const double2*>(&r.e[tid][0]);
                const double2 e23 = *reinterpret_cast<const double2*>(&r.e[tid][2]);
                const double df2 = -e01.x * s_y[2 * ip2] - e01.y * s_y[2 * ip2 + 1];
                const double num = 6 * (f1 - s_f[ip2]) - 2 * df2;
                s0 += num * e23.x;
                s1 += num * e23.y;
            }
            if (flags & 2) {                                    // (block-uniform: the shuffles are full-warp)
                double t0 = s0, t1 = s1;
#pragma unroll
                for (int off = GS_LANES / 2; off > 0; off >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, off);
                }
                if (ip >= 0 && sub == 0) {
                    const double2 qa = *reinterpret_cast<const double2*>(&r.qinv[slot][0]);
                    const double2 qb = *reinterpret_cast<const double2*>(&r.qinv[slot][2]);
                    const double r0 = qa.x * t0 + qa.y * t1;
                    const double r1 = qb.x * t0 + qb.y * t1;
                    const double change = fmax(fabs(s_y[2 * ip] + r0), fabs(s_y[2 * ip + 1] + r1));
                    s_y[2 * ip] = -r0;
                    s_y[2 * ip + 1] = -r1;
                    const double den = fmax(1.0, fmax(fabs(r0), fabs(r1)));
                    if (change * bd > bc * den) { bc = change; bd = den; }      // change/den > bc/bd
                }
            }
            __syncthreads();      // the step's gradients are visible; its ring stage is free
        }
        double err = bc / bd;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, off));
        if ((tid & 31) == 0) s_err[tid >> 5] = err;
        __syncthreads();
        if (tid == 0) {
            double e = s_err[0];
            for (int w = 1; w < GS_THREADS / 32; ++w) e = fmax(e, s_err[w]);
            if (e < tol) s_done = it + 1;
        }
        __syncthreads();
        if (s_done) { converged = s_done; break; }
    }
    // copies that were issued ahead but never consumed must land before the CTA may exit
    if (tid == 0) {
        while (ahead > 0) {
            mbar_wait(&s_full[st], ph);
            if (++st == nstage) { st = 0; ph ^= 1u; }
            --ahead;
        }
    }
    for (int i = tid; i < npts; i += GS_THREADS) {
        grad[((int64_t)i * 2) * Z + z] = s_y[2 * i];
        grad[((int64_t)i * 2 + 1) * Z + z] = s_y[2 * i + 1];
    }
    if (tid == 0 && niter) niter[z] = converged;   // 0: not converged within maxiter (scipy warns and goes on)
    __syncthreads();                               // (thread 0's waits above precede every thread's exit)
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

__global__ void __launch_bounds__(128)
ct_eval_kernel(int64_t ngrid, int Z, const CtPoint* __restrict__ geo, const double* __restrict__ data /* (npts, Z) */,
               const double* __restrict__ grad /* (npts, 2, Z) */, double* __restrict__ out /* (ngrid, Z) */,
               int* __restrict__ flags /* [0]: a bad value survived the nearest fill */) {
    const int64_t p = blockIdx.x;
    const CtPoint gp = geo[p];
    for (int z = threadIdx.x; z < Z; z += blockDim.x) {
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
        out[p * Z + z] = w;
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
    const size_t fixed = sizeof(double) * 3 * (size_t)npts;
    int nstage = GS_MAX_STAGE;
    while (nstage > 2 && fixed + (size_t)nstage * sizeof(GsRec) > 226 * 1024) --nstage;
    const size_t smem = fixed + (size_t)nstage * sizeof(GsRec);
    HOLO_REQUIRE(smem <= 226 * 1024, "holo_scatter_gradients: too many grid points for shared memory (M*Q <= 8700)");
    HOLO_CUDA(cudaFuncSetAttribute(ct_gradients_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ct_gradients_kernel<<<Z, GS_THREADS, smem, (cudaStream_t)stream>>>(npts, Z, (const GsRec*)program, nsteps, nstage, data, maxiter,
                                                                        tol, grad, niter);
    holo::count_launches(1);
    return holo_check_launch("holo_scatter_gradients");
}

int holo_scatter_ct_eval(int64_t ngrid, int Z, const void* geo, const double* data, const double* grad, double* out,
                         int* flags, void* stream) {
    HOLO_REQUIRE(geo && data && grad && out && flags, "holo_scatter_ct_eval: NULL argument");
    HOLO_REQUIRE(ngrid > 0 && ngrid < 2147483647LL && Z > 0, "holo_scatter_ct_eval: bad shape");
    HOLO_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), (cudaStream_t)stream));
    ct_eval_kernel<<<(int)ngrid, 128, 0, (cudaStream_t)stream>>>(ngrid, Z, (const CtPoint*)geo, data, grad, out, flags);
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

int holo_scatter_geo_bytes(void) { return (int)sizeof(CtPoint); }

}  // extern "C"
