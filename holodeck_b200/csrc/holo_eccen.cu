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

// E3 (discrete): one CTA per (f, n, realization tile); threads = realizations.
constexpr int ECC_TILE = 128;   // (z,M) pairs staged per pass

__global__ void __launch_bounds__(ECC_THREADS)
ecc_discrete_kernel(EccGeom g, EccConsts cc, const double* __restrict__ frst_pref, int R, int64_t r0,
                    uint32_t k0, uint32_t k1, double* __restrict__ gwb /* (F,H,R) */) {
    __shared__ double s_tau[ECC_TILE], s_h[ECC_TILE];
    __shared__ int s_pair[ECC_TILE];
    __shared__ int s_n;
    const int fh = blockIdx.x;
    const int ff = fh / g.H;
    const int nh = fh % g.H + 1;
    const int r = blockIdx.y * ECC_THREADS + threadIdx.x;
    const bool live = r < R;
    const int npair = g.Z * g.M;
    const double four_pi_c_mpc = 4 * CY_PI * (CY_SPLC / CY_MPC);
    DrawKey key;
    key.k0 = k0; key.k1 = k1; key.real = (uint32_t)(r0 + r); key.stream = 5;
    double acc = 0.0;
    for (int p0 = 0; p0 < npair; p0 += ECC_TILE) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        // stage the (z,M) pairs of this tile whose track covers (f,n): order within the tile is fixed by
        // a serial compaction (thread 0), so the accumulation order is reproducible
        double afac, tf = 0.0, hf = 0.0;
        bool ok = false;
        const int p = p0 + threadIdx.x;
        if (threadIdx.x < ECC_TILE && p < npair)
            ok = eccen_factor(g, cc, frst_pref, p / g.M, p % g.M, ff, nh, &afac, &tf, &hf);
        s_tau[threadIdx.x] = tf;
        s_h[threadIdx.x] = hf;
        s_pair[threadIdx.x] = ok ? p : -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            for (int i = 0; i < ECC_TILE; ++i) {
                if (s_pair[i] >= 0) {
                    const int pp = s_pair[i];
                    const double a = s_tau[i], b = s_h[i];
                    s_pair[n] = pp; s_tau[n] = a; s_h[n] = b;
                    ++n;
                }
            }
            s_n = n;
        }
        __syncthreads();
        if (!live) continue;
        for (int i = 0; i < s_n; ++i) {
            const int pp = s_pair[i];
            const int kk = pp / g.M, ii = pp % g.M;
            double kw, kdx, iw, idx_;
            trapz_grid_weight(kk, g.Z, g.redz, &kw, &kdx);
            trapz_grid_weight(ii, g.M, g.mtot_log10, &iw, &idx_);
            const double zterm = 1.0 + g.redz[kk];
            const double dc_mpc = g.dcom[kk];
            const double dc_cm = dc_mpc * CY_MPC;
            const double dc_term = four_pi_c_mpc * pow(dc_mpc, 2.0);
            const double mt = pow(10.0, g.mtot_log10[ii]);
            const double volume_ik = idx_ * kdx;                             // pyx:783-784
            const double weight_ik = iw * kw;
            for (int jj = 0; jj < g.Q; ++jj) {
                const double nd = g.ndens[((int64_t)ii * g.Q + jj) * g.Z + kk];
                if (!(nd > 0.0)) continue;                                    // Poisson(0) = 0
                double jw, jdx;
                trapz_grid_weight(jj, g.Q, g.mrat, &jw, &jdx);
                const double volume = volume_ik * jdx;
                const double weight = weight_ik * jw;
                const double q = g.mrat[jj];
                const double m1 = mt / (1.0 + q);
                const double m2 = mt - m1;
                const double mchirp = mt * pow(q, 3.0 / 5.0) / pow(1 + q, 6.0 / 5.0);
                const double number_term_pref = nd * dc_term * zterm;          // pyx:799
                const double hterm_pref = pow(cc.gw_src_const * mchirp * pow(2.0 * mchirp, 2.0 / 3.0) / dc_cm, 2.0);
                const double tau = s_tau[i] / (m1 * m2);                       // pyx:826
                const double number_term = number_term_pref * tau * volume;    // pyx:829
                const double hterm = hterm_pref * s_h[i];                      // pyx:832
                const uint64_t idx = ((uint64_t)(((int64_t)ii * g.Q + jj) * g.Z + kk)) * (uint64_t)(g.F * g.H) + (uint64_t)fh;
                const double num = draw_element(number_term, 1.0e300, key, idx);
                acc += hterm * num / weight;                                   // pyx:839
            }
        }
    }
    if (live) gwb[(int64_t)fh * R + r] = acc;
}

}  // namespace holo

using namespace holo;

extern "C" {

int64_t holo_eccen_workspace_bytes(int M, int Q, int Z, int F, int nharms, int nreals) {
    (void)Q; (void)F; (void)nharms; (void)nreals;
    return 256 + 8 * ((int64_t)Z * M + 4096);
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
        dim3 grid(F * nharms, (nreals + ECC_THREADS - 1) / ECC_THREADS);
        ecc_discrete_kernel<<<grid, ECC_THREADS, 0, st>>>(g, cc, frst_pref, nreals, r0, (uint32_t)seed,
                                                          (uint32_t)(seed >> 32), gwb); holo::count_launches(1);
    }
    return holo_check_launch("holo_sam_calc_gwb_single_eccen");
}

}  // extern "C"
