// holo_eccen.cu -- K5: eccentric-SAM GWB harmonic sum for sm_100a.
//
// Replaces   _sam_calc_gwb_single_eccen            holodeck/cyutils.pyx:370-597   -> gwb (F,H)
//            _sam_calc_gwb_single_eccen_discrete   holodeck/cyutils.pyx:609-851   -> gwb (F,H,R)
//
// The reference loops z -> M (descending) -> q -> (f,n) sorted by f/n, evaluating two scipy `jv`
// Bessel functions per innermost iteration although the eccentricity e(f_r), g(n,e), F(e) and the
// separation depend on (z, M, f, n) only.  Here:
//   E1  per (z,M): the q-sum of everything that multiplies the (f,n)-dependent factor (continuous case);
//   E2  per (f,n): one CTA strides over the (z,M) pairs, evaluates A(z,M,f,n) (Bessel by Miller's
//       backward recurrence, then the reference's three upward steps) and reduces A*B in a fixed order.
//   E3  (discrete) per (f,n): the same A-terms are staged per tile of (z,M) pairs, then one THREAD per
//       realization walks the tile's (z,M,q) cells drawing Poisson(number_term) (Philox, holo_rng.cuh).
// The monotone `ecc_idx` walk of the reference (pyx:536-565) always starts at or below the bracketing
// index (its hint comes from a larger mass / lower frequency), so a stateless bracket search is
// equivalent; "continue" (below the track) and "break" (above the track, frequencies ascending) both
// mean "this (f,n) gets nothing from this (z,M)".
#include <cuda_runtime.h>

#include "holo_api.cuh"
#include "holo_eccen_math.cuh"
#include "holo_rng.cuh"

namespace holo {

__global__ void ecc_prefactor_kernel(const double* __restrict__ sepa_evo, int E, double* __restrict__ frst_pref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) frst_pref[i] = ((1.0 / (2.0 * CY_PI)) * sqrt(CY_NWTG)) / pow(sepa_evo[i], 1.5);   // pyx:441, 448
}

// E1: B(z,M) = sum_q weight * hterm_pref / (m1 m2)        (pyx:488-533 hoisted out of the (f,n) loop)
__global__ void ecc_qsum_kernel(EccGeom g, EccConsts cc, double* __restrict__ bsum /* (Z,M) */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.Z * g.M) return;
    const int kk = i / g.M, ii = i % g.M;
    const double four_pi_c_mpc = 4 * CY_PI * (CY_SPLC / CY_MPC);
    double kw, kdx, iw, idx_;
    trapz_grid_weight(kk, g.Z, g.redz, &kw, &kdx);
    trapz_grid_weight(ii, g.M, g.mtot_log10, &iw, &idx_);
    const double zterm = 1.0 + g.redz[kk];
    const double dc_mpc = g.dcom[kk];
    const double dc_cm = dc_mpc * CY_MPC;
    const double dc_term = four_pi_c_mpc * pow(dc_mpc, 2.0);
    const double mt = pow(10.0, g.mtot_log10[ii]);
    const double weight_ik = idx_ * kdx / (iw * kw);
    double acc = 0.0;
    for (int jj = 0; jj < g.Q; ++jj) {
        double jw, jdx;
        trapz_grid_weight(jj, g.Q, g.mrat, &jw, &jdx);
        const double weight = weight_ik * (jdx / jw);
        const double q = g.mrat[jj];
        const double m1 = mt / (1.0 + q);
        const double m2 = mt - m1;
        const double mchirp = mt * pow(q, 3.0 / 5.0) / pow(1 + q, 6.0 / 5.0);
        double hterm_pref = g.ndens[((int64_t)ii * g.Q + jj) * g.Z + kk] * dc_term * zterm;
        hterm_pref *= pow(cc.gw_src_const * mchirp * pow(2.0 * mchirp, 2.0 / 3.0) / dc_cm, 2.0);
        acc += weight * hterm_pref / (m1 * m2);
    }
    bsum[i] = acc;
}

constexpr int ECC_THREADS = 128;

// E2: gwb[f, n-1] = sum_{z,M} A(z,M,f,n) * B(z,M)
__global__ void __launch_bounds__(ECC_THREADS)
ecc_sum_kernel(EccGeom g, EccConsts cc, const double* __restrict__ frst_pref, const double* __restrict__ bsum,
               double* __restrict__ gwb) {
    __shared__ double red[ECC_THREADS];
    const int ff = blockIdx.x / g.H;
    const int nh = blockIdx.x % g.H + 1;
    const int npair = g.Z * g.M;
    double acc = 0.0;
    for (int p = threadIdx.x; p < npair; p += ECC_THREADS) {
        const double b = bsum[p];
        if (b == 0.0) continue;
        const int kk = p / g.M, ii = p % g.M;
        double afac, tf, hf;
        if (eccen_factor(g, cc, frst_pref, kk, ii, ff, nh, &afac, &tf, &hf)) acc += afac * b;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = ECC_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) gwb[(int64_t)ff * g.H + (nh - 1)] = red[0];
}

// E3 (discrete): gwb[f,n,r] = sum_cells Poisson(number_term) * hterm / weight  (pyx:783-839) is the realised-GWB
// sum of the SAM path with (f,n) pairs as "frequencies": the (cell, column) expectation values and strain factors are
// tabulated for a slab of columns and handed to the realization kernel (shared CDF tables, superposition groups).
//   E3a  per (z,M,q):  c1 = ndens * 4 pi c d_c^2 (1+z) * volume / (m1 m2),  c2 = hterm_pref / weight
//   E3b  per (z,M) x column: taufac, hfac (Bessel) once, then for all q:  number = c1 * taufac,  h = c2 * hfac
__global__ void ecc_cell_consts_kernel(EccGeom g, EccConsts cc, double* __restrict__ c1, double* __restrict__ c2) {
    const int64_t n = (int64_t)g.M * g.Q * g.Z;
    const double four_pi_c_mpc = 4 * CY_PI * (CY_SPLC / CY_MPC);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(i % g.Z);
        const int64_t mq = i / g.Z;
        const int jj = (int)(mq % g.Q), ii = (int)(mq / g.Q);
        double kw, kdx, iw, idx_, jw, jdx;
        trapz_grid_weight(kk, g.Z, g.redz, &kw, &kdx);
        trapz_grid_weight(ii, g.M, g.mtot_log10, &iw, &idx_);
        trapz_grid_weight(jj, g.Q, g.mrat, &jw, &jdx);
        const double zterm = 1.0 + g.redz[kk];
        const double dc_mpc = g.dcom[kk];
        const double dc_cm = dc_mpc * CY_MPC;
        const double dc_term = four_pi_c_mpc * pow(dc_mpc, 2.0);
        const double mt = pow(10.0, g.mtot_log10[ii]);
        const double volume = idx_ * kdx * jdx;                               // pyx:783-784
        const double weight = iw * kw * jw;
        const double q = g.mrat[jj];
        const double m1 = mt / (1.0 + q);
        const double m2 = mt - m1;
        const double mchirp = mt * pow(q, 3.0 / 5.0) / pow(1 + q, 6.0 / 5.0);
        const double nd = g.ndens[i];
        const double number_term_pref = nd * dc_term * zterm;                 // pyx:799
        const double hterm_pref = pow(cc.gw_src_const * mchirp * pow(2.0 * mchirp, 2.0 / 3.0) / dc_cm, 2.0);
        c1[i] = (nd > 0.0) ? number_term_pref / (m1 * m2) * volume : 0.0;      // pyx:826, 829
        c2[i] = hterm_pref / weight;                                          // pyx:832, 839
    }
}

__global__ void __launch_bounds__(128)
ecc_slab_kernel(EccGeom g, EccConsts cc, const double* __restrict__ frst_pref, const double* __restrict__ c1,
                const double* __restrict__ c2, int col0, int ncol, double* __restrict__ number /* (ncell, ncol) */,
                double* __restrict__ hval) {
    const int p = blockIdx.x;                    // (z, M) pair
    const int kk = p / g.M, ii = p % g.M;
    for (int cl = threadIdx.x; cl < ncol; cl += blockDim.x) {
        const int fh = col0 + cl;
        double afac, tf = 0.0, hf = 0.0;
        const bool ok = (fh < g.F * g.H) && eccen_factor(g, cc, frst_pref, kk, ii, fh / g.H, fh % g.H + 1, &afac, &tf, &hf);
        for (int jj = 0; jj < g.Q; ++jj) {
            const int64_t cell = ((int64_t)ii * g.Q + jj) * g.Z + kk;
            const double lam = ok ? c1[cell] * tf : 0.0;
            number[cell * ncol + cl] = lam > 0.0 ? lam : 0.0;
            hval[cell * ncol + cl] = ok ? c2[cell] * hf : 0.0;
        }
    }
}

static int ecc_slab_columns(int64_t ncell, int ncols_total) {
    // columns per slab: a multiple of 4, two (ncell, cols) fp64 arrays within ~6 GB
    int64_t cols = (int64_t)6.0e9 / (16 * ncell);
    cols = (cols / 4) * 4;
    if (cols < 4) cols = 4;
    const int64_t all = ((int64_t)ncols_total + 3) / 4 * 4;
    return (int)(cols < all ? cols : all);
}

}  // namespace holo

using namespace holo;

extern "C" {

int64_t holo_realize_workspace_bytes(int kind, int64_t ncell, int F, int R);

int64_t holo_eccen_workspace_bytes(int M, int Q, int Z, int F, int nharms, int nreals) {
    int64_t base = 256 + 8 * ((int64_t)Z * M + 4096);
    if (nreals > 0) {
        const int64_t ncell = (int64_t)M * Q * Z;
        const int cols = ecc_slab_columns(ncell, F * nharms);
        base += 2 * 8 * ncell + 256;                                   // c1, c2
        base += 2 * 8 * ncell * cols + 512;                            // number, hval slabs
        base += holo_realize_workspace_bytes(0, ncell, cols, nreals) + 256;
    }
    return base;
}

int holo_sam_calc_gwb_single_eccen(holo_cy_consts cyc, double gw_src_const, const double* ndens,
                                   const double* mtot_log10, const double* mrat, const double* redz,
                                   const double* dcom_mpc, const double* gwfobs, const double* sepa_evo,
                                   const double* eccen_evo, int M, int Q, int Z, int F, int E, int nharms,
                                   int nreals, int64_t r0, uint64_t seed, double* gwb, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    HOLO_REQUIRE(ndens && mtot_log10 && mrat && redz && dcom_mpc && gwfobs && sepa_evo && eccen_evo && gwb && workspace,
                 "holo_sam_calc_gwb_single_eccen: NULL argument");
    HOLO_REQUIRE(M > 1 && Q > 1 && Z > 1 && F > 0 && E > 1 && nharms > 0 && nreals >= 0,
                 "holo_sam_calc_gwb_single_eccen: bad shape");
    HOLO_REQUIRE(E <= 4096, "holo_sam_calc_gwb_single_eccen: evolution track longer than 4096 steps");
    HOLO_REQUIRE(workspace_bytes >= holo_eccen_workspace_bytes(M, Q, Z, F, nharms, nreals),
                 "holo_sam_calc_gwb_single_eccen: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* frst_pref = (double*)workspace;
    double* bsum = frst_pref + 4096;
    EccGeom g{ndens, mtot_log10, mrat, redz, dcom_mpc, gwfobs, sepa_evo, eccen_evo, M, Q, Z, F, E, nharms};
    EccConsts cc{cyc.gw_dadt_sep_const, gw_src_const};
    ecc_prefactor_kernel<<<(E + 127) / 128, 128, 0, st>>>(sepa_evo, E, frst_pref); holo::count_launches(1);
    if (nreals == 0) {
        ecc_qsum_kernel<<<(Z * M + 127) / 128, 128, 0, st>>>(g, cc, bsum); holo::count_launches(1);
        ecc_sum_kernel<<<F * nharms, ECC_THREADS, 0, st>>>(g, cc, frst_pref, bsum, gwb); holo::count_launches(1);
    } else {
        const int64_t ncell = (int64_t)M * Q * Z;
        const int FH = F * nharms;
        const int cols = ecc_slab_columns(ncell, FH);
        auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
        unsigned char* wp = (unsigned char*)workspace + align(8 * ((int64_t)Z * M + 4096));
        double* c1 = (double*)wp; wp += align(8 * ncell);
        double* c2 = (double*)wp; wp += align(8 * ncell);
        double* number = (double*)wp; wp += align(8 * ncell * cols);
        double* hval = (double*)wp; wp += align(8 * ncell * cols);
        const int64_t rws = (int64_t)((unsigned char*)workspace + workspace_bytes - wp);
        ecc_cell_consts_kernel<<<148 * 8, 256, 0, st>>>(g, cc, c1, c2); holo::count_launches(1);
        for (int col0 = 0; col0 < FH; col0 += cols) {
            // the last slab keeps the full width (columns beyond F*H are empty) so that every slab is 4-aligned
            ecc_slab_kernel<<<Z * M, 128, 0, st>>>(g, cc, frst_pref, c1, c2, col0, cols, number, hval); holo::count_launches(1);
            int rc = holo_check_launch("holo_sam_calc_gwb_single_eccen: slab");
            if (rc) return rc;
            const int live = (FH - col0) < cols ? (FH - col0) : cols;
            if (live == cols) {
                rc = holo::realize_gwb_columns(number, hval, ncell, cols, nreals, r0, seed, 9.0e18, nullptr, col0, FH,
                                               gwb + (int64_t)col0 * nreals, wp, rws, st);
            } else {
                // ragged tail: realise the full-width slab into scratch rows, copy the live rows out
                double* tail = hval;   // reuse after the launch below has consumed it: stream-ordered
                rc = holo::realize_gwb_columns(number, hval, ncell, cols, nreals, r0, seed, 9.0e18, nullptr, col0, FH,
                                               number /* (cols, R) fits: cols*R <= ncell*cols */, wp, rws, st);
                (void)tail;
                if (!rc) HOLO_CUDA(cudaMemcpyAsync(gwb + (int64_t)col0 * nreals, number, sizeof(double) * (int64_t)live * nreals,
                                                    cudaMemcpyDeviceToDevice, st));
            }
            if (rc) return rc;
        }
    }
    return holo_check_launch("holo_sam_calc_gwb_single_eccen");
}

}  // extern "C"
