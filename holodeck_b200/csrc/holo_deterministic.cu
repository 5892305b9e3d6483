// holo_deterministic.cu -- K0 (density), K1a (2PL norm), K1b/K1c (dynamic binary number),
// K2 (bin integration), K2b (strain) for sm_100a.  Compiled with -fmad=false so that the fp64
// arithmetic rounds like the reference's gcc -O2 x86-64 build (no FMA contraction); these kernels
// are HBM- or latency-bound, not FMA-bound.
#include <cuda_runtime.h>

#include "holo_api.cuh"
#include "holo_math.cuh"

namespace holo {

static inline CyConsts to_cc(const holo_cy_consts& c) {
    CyConsts cc;
    cc.gw_dadt_sep_const = c.gw_dadt_sep_const;
    cc.kepler_const_freq = c.kepler_const_freq;
    cc.kepler_const_sepa = c.kepler_const_sepa;
    cc.four_pi_c_over_mpc = c.four_pi_c_over_mpc;
    return cc;
}

// -------------------------------------------------------------------------------------------------
// K0
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
density_kernel(const double* __restrict__ mtot, const double* __restrict__ mrat,
               const double* __restrict__ redz, const double* __restrict__ age_z,
               const double* __restrict__ dtdz_z, int M, int Q, int Z, holo_sam_params par,
               const double* __restrict__ bf_tab, double* __restrict__ dens, double* __restrict__ gmt_time,
               double* __restrict__ redz_prime) {
    int64_t n = (int64_t)M * Q * Z;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int kk = (int)(i % Z);
        int64_t mq = i / Z;
        int jj = (int)(mq % Q);
        int ii = (int)(mq / Q);
        DensityOut o = density_point(par, bf_tab, mtot[ii], mrat[jj], redz[kk], age_z[kk], dtdz_z[kk]);
        dens[i] = o.dens;
        if (gmt_time) gmt_time[i] = o.gmt_time;
        if (redz_prime) redz_prime[i] = o.redz_prime;
    }
}

__global__ void zero_stalled_kernel(double* __restrict__ dens, const double* __restrict__ zp, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (zp[i] < 0.0) dens[i] = 0.0;
    }
}

// -------------------------------------------------------------------------------------------------
// K1a: one warp per (M,q); the nsteps separation steps are spread over the lanes.
// -------------------------------------------------------------------------------------------------
struct NormTrack {
    const double* sepa;
    const double* p1;
    const double* p2;
    const double* gw;
    int nsteps;
};

// get_binary_lifetime_2pwl (sam_cyutils.pyx:360-398) minus `target`; every lane returns the same bits.
__device__ __forceinline__ double lifetime_minus_target(const NormTrack& t, double norm_log10,
                                                        double target, int lane) {
    double norm = pow(10.0, norm_log10);
    double part = 0.0;
    for (int s = lane; s < t.nsteps; s += 32) {
        double dl = (-norm * t.p1[s]) / t.p2[s] + t.gw[s];
        double dr = (-norm * t.p1[s + 1]) / t.p2[s + 1] + t.gw[s + 1];
        part += 2.0 * (t.sepa[s + 1] - t.sepa[s]) / (dl + dr);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    return part - target;
}

struct LifetimeFn {
    NormTrack t;
    double target;
    int lane;
    __device__ __forceinline__ double operator()(double x) const {
        return lifetime_minus_target(t, x, target, lane);
    }
};

constexpr int NORM_WARPS = 4;

// mode 0: solve for norm_log10 (out = root); mode 1: out = lifetime(norm_in[i]) (target = 0)
__global__ void __launch_bounds__(NORM_WARPS * 32)
norm_2pwl_kernel(CyConsts cc, double target_time, const double* __restrict__ mtot,
                 const double* __restrict__ mrat, int N, double sepa_init_log10, double rchar,
                 double gamma_inner, double gamma_outer, int nsteps, int mode,
                 const double* __restrict__ norm_in, double* __restrict__ out) {
    extern __shared__ double smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * NORM_WARPS + warp;
    if (i >= N) return;
    int ne = nsteps + 1;
    double* sepa = smem + (size_t)warp * 4 * ne;
    double* p1 = sepa + ne;
    double* p2 = p1 + ne;
    double* gw = p2 + ne;
    double mt = mtot[i], mr = mrat[i];
    double risco_log10 = log10(3.0 * CY_SCHW * mt);                      // pyx:363
    double dx = (sepa_init_log10 - risco_log10) / nsteps;               // pyx:368
    // the running `sepa_log10 -= dx` of the reference (pyx:381) is kept serial so it rounds identically:
    // lane 0 walks the chain once into shared memory, then the lanes share the transcendental work
    if (lane == 0) {
        double slog = sepa_init_log10;
        for (int k = 0; k < ne; ++k) {
            sepa[k] = slog;
            slog -= dx;
        }
    }
    __syncwarp();
    for (int k = lane; k < ne; k += 32) {
        const double sp = pow(10.0, sepa[k]);
        const double xx = sp / rchar;
        sepa[k] = sp;
        p1[k] = pow(1.0 + xx, -gamma_outer + gamma_inner);              // pyx:242
        p2[k] = pow(xx, gamma_inner - 1.0);
        gw[k] = hard_gw(cc, mt, mr, sp);
    }
    __syncwarp();
    NormTrack t{sepa, p1, p2, gw, nsteps};
    double res;
    if (mode == 0) {
        LifetimeFn fn{t, target_time, lane};
        res = brentq(fn, -20.0, 20.0, 1e-3, 1e-5, 100);                  // pyx:332-349
    }
    else res = lifetime_minus_target(t, norm_in[i], 0.0, lane);
    if (lane == 0) out[i] = res;
}

__global__ void hard_func_kernel(CyConsts cc, const double* __restrict__ mtot,
                                 const double* __restrict__ mrat, const double* __restrict__ sepa,
                                 const double* __restrict__ norm, double rchar, double gi, double go,
                                 int64_t N, double* __restrict__ dadt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x)
        dadt[i] = hard_func_2pwl_gw(cc, mtot[i], mrat[i], sepa[i], norm[i], rchar, gi, go);
}

// -------------------------------------------------------------------------------------------------
// K1b: one CTA per (M,q).  Phase 1 builds the evolution track in shared memory (the two running
// sums of the reference, `sepa_log10 -= dx` and `time_evo += dt`, are kept serial so they round
// identically); phase 2 solves every (z,f) pair independently by bisection over the track.
// -------------------------------------------------------------------------------------------------
constexpr int DBN_THREADS = 128;

__global__ void __launch_bounds__(DBN_THREADS)
dbn_2pwl_kernel(CyConsts cc, const double* __restrict__ fobs, int F, double sepa_init_log10,
                int nsteps, const double* __restrict__ hard_norm, double rchar, double gamma_inner,
                double gamma_outer, const double* __restrict__ nden, const double* __restrict__ mtot,
                const double* __restrict__ mrat, const double* __restrict__ redz,
                const double* __restrict__ gmt_time, int M, int Q, int Z,
                const double* __restrict__ grid_z, const double* __restrict__ grid_dcom,
                const double* __restrict__ grid_age, int n_interp, double* __restrict__ redz_final,
                double* __restrict__ diff_num) {
    extern __shared__ double smem[];
    int ne = nsteps + 1;
    double* s_sepa = smem;               // (ne) : first log10(sepa), then sepa
    double* s_dadt = s_sepa + ne;        // (ne)
    double* s_frst = s_dadt + ne;        // (ne)
    double* s_tevo = s_frst + ne;        // (ne)
    double* s_dt = s_tevo + ne;          // (ne)
    double* s_tage = s_dt + ne;          // (n_interp)
    double* s_gz = s_tage + n_interp;
    double* s_gdc = s_gz + n_interp;
    double* s_zage = s_gdc + n_interp;   // (Z)
    double* s_fobs = s_zage + Z;         // (F)

    int mq = blockIdx.x;
    int ii = mq / Q, jj = mq % Q;
    int tid = threadIdx.x;
    double mt = mtot[ii], mr = mrat[jj];
    double norm = hard_norm[mq];

    for (int i = tid; i < n_interp; i += DBN_THREADS) {
        s_tage[i] = grid_age[i];
        s_gz[i] = grid_z[i];
        s_gdc[i] = grid_dcom[i];
    }
    for (int i = tid; i < F; i += DBN_THREADS) s_fobs[i] = fobs[i];
    __shared__ MqConsts s_mqc;
    if (tid == 32) s_mqc = mq_consts(cc, mt, mr);                        // (three pow calls: once per CTA, not per thread)
    if (tid == 0) {
        double risco = 3.0 * CY_SCHW * mt;                               // pyx:609
        double dx = (sepa_init_log10 - log10(risco)) / nsteps;          // pyx:610
        double slog = sepa_init_log10;
        s_sepa[0] = slog;
        for (int k = 1; k < ne; ++k) {
            slog -= dx;                                                  // pyx:636
            s_sepa[k] = slog;
        }
    }
    __syncthreads();
    // ages of the SAM redshift edges from the (decreasing-z) table, pyx:590-601
    for (int k = tid; k < Z; k += DBN_THREADS) {
        double zz = redz[k];
        int idx = bracket_decreasing(n_interp, zz, s_gz);
        s_zage[k] = interp_at_index(idx, zz, s_gz, s_tage);
    }
    for (int k = tid; k < ne; k += DBN_THREADS) {
        double sp = pow(10.0, s_sepa[k]);                                // pyx:621, 637
        s_dadt[k] = hard_func_2pwl_gw(cc, mt, mr, sp, norm, rchar, gamma_inner, gamma_outer);
        s_frst[k] = kepler_freq_from_sepa(cc, mt, sp);
        s_sepa[k] = sp;   // each thread only touches its own k: in-place is safe
    }
    __syncthreads();
    for (int k = tid; k < nsteps; k += DBN_THREADS)
        s_dt[k] = 2.0 * (s_sepa[k + 1] - s_sepa[k]) / (s_dadt[k] + s_dadt[k + 1]);   // pyx:648
    __syncthreads();
    if (tid == 0) {
        double tevo = 0.0;
        s_tevo[0] = 0.0;
        for (int k = 0; k < nsteps; ++k) {
            tevo += s_dt[k];                                             // pyx:652
            s_tevo[k + 1] = tevo;
        }
    }
    __syncthreads();

    Track2pwl t;
    t.frst = s_frst; t.tevo = s_tevo; t.dt = s_dt; t.nsteps = nsteps;
    t.tage = s_tage; t.gz = s_gz; t.gdc = s_gdc; t.n_interp = n_interp;
    t.age_universe = s_tage[n_interp - 1];                               // pyx:579

    const MqConsts mqc = s_mqc;
    int64_t base = (int64_t)mq * Z;
    int nzf = Z * F;
    // phase 2: one thread per redshift finds, for every target frequency, the first step whose right edge reaches
    // it -- bisection for the lowest frequency, then a merge walk (both sequences ascend); when `fobs` is not
    // ascending every frequency is bisected on its own.
    unsigned short* s_lo = reinterpret_cast<unsigned short*>(s_fobs + F);   // (Z, F)
    for (int kk = tid; kk < Z; kk += DBN_THREADS) {
        const double gmt = gmt_time[base + kk], az = s_zage[kk];
        int lo = 0, hint = 0, fr_lo = -1;
        double fprev = 0.0, fr = 0.0;
        for (int ff = 0; ff < F; ++ff) {
            const double ft = s_fobs[ff];
            if (ff == 0 || ft < fprev) {
                lo = dbn_2pwl_first_step(t, gmt, az, ft);
                // age-table bracket of the step the walk resumes from (the walk only moves it up from here)
                hint = bracket_increasing(n_interp, s_tevo[(lo < nsteps ? lo : nsteps - 1) + 1] + gmt + az, s_tage);
                fr_lo = -1;
            } else {
                lo = dbn_2pwl_next_step_walk(t, gmt, az, ft, lo, hint, fr_lo, fr);
            }
            fprev = ft;
            s_lo[kk * F + ff] = (unsigned short)lo;
        }
    }
    __syncthreads();
    // phase 3: every (z,f) pair evaluates its (one or two) candidate steps; f is the fastest index, so the
    // two stores of a warp are fully coalesced
    for (int idx = tid; idx < nzf; idx += DBN_THREADS) {
        int kk = idx / F, ff = idx - kk * F;
        double rz = -1.0, dn = 0.0;                                      // pyx:460-461
        double gmt = gmt_time[base + kk];
        double nd = nden[base + kk];
        dbn_2pwl_cell_from(cc, mqc, t, norm, rchar, gamma_inner, gamma_outer, nd, gmt, s_zage[kk],
                           s_fobs[ff], (int)s_lo[idx], &rz, &dn);
        int64_t o = (base + kk) * F + ff;
        redz_final[o] = rz;
        diff_num[o] = dn;
    }
}

// -------------------------------------------------------------------------------------------------
// K1c
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dbn_gw_kernel(CyConsts cc, const double* __restrict__ fobs, int F, const double* __restrict__ nden,
              const double* __restrict__ mtot, const double* __restrict__ mrat,
              const double* __restrict__ redz_prime, int M, int Q, int Z,
              const double* __restrict__ grid_z, const double* __restrict__ grid_dcom, int n_interp,
              double* __restrict__ redz_final, double* __restrict__ diff_num) {
    int64_t n = (int64_t)M * Q * Z * F;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int ff = (int)(i % F);
        int64_t c = i / F;
        int64_t mq = c / Z;
        int jj = (int)(mq % Q);
        int ii = (int)(mq / Q);
        double rz, dn;
        dbn_gw_cell(cc, mtot[ii], mrat[jj], nden[c], redz_prime[c], fobs, ff, grid_z, grid_dcom,
                    n_interp, &rz, &dn);
        redz_final[i] = rz;
        diff_num[i] = dn;
    }
}

// -------------------------------------------------------------------------------------------------
// K2 / K2b / fused
// -------------------------------------------------------------------------------------------------
struct BinGeom {
    int M, Q, Z, F;   // edge counts
};

// One CTA per (m,q) bin; its threads sweep the (z,f) plane, f fastest (coalesced 8-corner stencils).
// All index arithmetic is 32-bit; only the final address is 64-bit.
template <bool DO_NUM, bool DO_STRAIN>
__global__ void __launch_bounds__(256)
bin_kernel(BinGeom g, GLTable gl, double hubble_distance, double om0, double gw_src_const, double nwtg,
           const double* __restrict__ log10_mtot, const double* __restrict__ mrat,
           const double* __restrict__ redz, const double* __restrict__ dln_freq,
           const double* __restrict__ dnum, const double* __restrict__ redz_final,
           const double* __restrict__ rz_mid, const double* __restrict__ mt_mid,
           const double* __restrict__ mr_mid, const double* __restrict__ fc,
           const double* __restrict__ fc_over_df, double* __restrict__ numb,
           double* __restrict__ h2fdf, double* __restrict__ zmid, double* __restrict__ dcom,
           double* __restrict__ sepa, double* __restrict__ angs, int32_t* __restrict__ bad_redz,
           const double* __restrict__ dc_tab, int dc_n, double dc_inv_h) {
    const int Qb = g.Q - 1, Zb = g.Z - 1, F = g.F;
    const int mm = blockIdx.x / Qb, qq = blockIdx.x - mm * Qb;
    bool any_bad = false;
    const int64_t sZ = F, sQ = (int64_t)g.Z * F, sM = (int64_t)g.Q * g.Z * F;
    const bool want_par = (zmid != nullptr) || (dcom != nullptr) || (sepa != nullptr) || (angs != nullptr);
    const int64_t in0 = mm * sM + qq * sQ;                      // first edge element of this (m,q)
    const int64_t out0 = (int64_t)blockIdx.x * Zb * F;         // first bin element of this (m,q)
    double dmdq = 0.0, mc = 0.0, mtm = 0.0;
    if (DO_NUM) {
        const double dm = log10_mtot[mm + 1] - log10_mtot[mm];          // pyx:195
        dmdq = dm * (mrat[qq + 1] - mrat[qq]);                          // pyx:198
    }
    if (DO_STRAIN) {
        mtm = mt_mid[mm];
        mc = chirp_mass_mtmr(mtm, mr_mid[qq]);
    }
    const int nzf = Zb * F;
    for (int zf = threadIdx.x; zf < nzf; zf += blockDim.x) {
        const int zz = zf / F, ff = zf - zz * F;
        const int64_t base = in0 + zz * sZ + ff;
        const int64_t i = out0 + zf;
        if (DO_NUM) {
            const double dmdqdz = dmdq * (redz[zz + 1] - redz[zz]);      // pyx:201
            numb[i] = integrate_bin(dnum, sM, sQ, sZ, base, dmdqdz, dln_freq[ff]);
        }
        if (DO_STRAIN) {
            bool bad = false;
            const double zc = redz_final ? corner_mean_redz(redz_final, sM, sQ, sZ, base, bad_redz ? &bad : nullptr) : rz_mid[zz];
            any_bad |= bad;
            const StrainOut o = strain_cell(gl, hubble_distance, om0, gw_src_const, nwtg, zc, mc, mtm, fc[ff],
                                            fc_over_df[ff], want_par, dc_tab, dc_n, dc_inv_h);
            h2fdf[i] = o.h2fdf;
            if (zmid) zmid[i] = o.zmid;
            if (dcom) dcom[i] = o.dcom;
            if (sepa) sepa[i] = o.sepa;
            if (angs) angs[i] = o.angs;
        }
    }
    if (DO_STRAIN && any_bad) atomicOr(bad_redz, 1);
}

// hc2[f] = sum_cells number*h2fdf : two deterministic passes
constexpr int EXP_ROWS = 8;
__global__ void __launch_bounds__(32 * EXP_ROWS)
expect_partial_kernel(const double* __restrict__ number, const double* __restrict__ h2fdf,
                      int64_t ncell, int F, int64_t cells_per_block, double* __restrict__ partial) {
    __shared__ double red[EXP_ROWS][33];
    int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    int64_t c0 = blockIdx.x * cells_per_block;
    int64_t c1 = c0 + cells_per_block;
    if (c1 > ncell) c1 = ncell;
    for (int f0 = 0; f0 < F; f0 += 32) {
        int f = f0 + x;
        double acc = 0.0;
        if (f < F)
            for (int64_t c = c0 + y; c < c1; c += EXP_ROWS) acc += number[c * F + f] * h2fdf[c * F + f];
        red[y][x] = acc;
        __syncthreads();
        if (y == 0 && f < F) {
            double s = 0.0;
            for (int r = 0; r < EXP_ROWS; ++r) s += red[r][x];
            partial[(int64_t)blockIdx.x * F + f] = s;
        }
        __syncthreads();
    }
}

__global__ void expect_final_kernel(const double* __restrict__ partial, int nblk, int F,
                                    double* __restrict__ hc2) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * F + f];
    hc2[f] = s;
}

// -------------------------------------------------------------------------------------------------
// K7: lib_tools._calc_model_details (librarian/lib_tools.py:904-939): per total-mass bin and frequency, the
// strain-weighted and the plain number of binaries histogrammed over their FINAL redshift (cell-centre mean of
// redz_final, sentinels included, as utils.midpoints does; bins = the SAM redshift edges, last bin closed;
// scipy.stats.binned_statistic(..., statistic='sum')).
// One CTA per (mass bin, group of 4 frequencies); thread <-> redshift bin.  The cells of one q row are staged
// (bin index + the two values), then every thread gathers the cells that fall in ITS bin in (q, z) order:
// no floating-point atomics, the sums are bit-reproducible.
// -------------------------------------------------------------------------------------------------
constexpr int DET_FG = 4;

__global__ void __launch_bounds__(128)
details_hist_kernel(BinGeom g, const double* __restrict__ redz_edges, const double* __restrict__ redz_final,
                    const double* __restrict__ number, const double* __restrict__ h2fdf,
                    double* __restrict__ gwb_hist /* (M-1, Z-1, F) */, double* __restrict__ num_hist) {
    extern __shared__ unsigned char s_raw[];
    const int Qb = g.Q - 1, Zb = g.Z - 1, F = g.F;
    double* s_edges = reinterpret_cast<double*>(s_raw);           // (Z)
    double* s_hv = s_edges + g.Z;                                 // (DET_FG, Zb) hc2 * number
    double* s_nv = s_hv + DET_FG * Zb;                            // (DET_FG, Zb) number
    int* s_bin = reinterpret_cast<int*>(s_nv + DET_FG * Zb);      // (DET_FG, Zb) bin index or -1
    const int mm = blockIdx.x, f0 = blockIdx.y * DET_FG;
    const int nf = (F - f0) < DET_FG ? (F - f0) : DET_FG;
    const int64_t sZ = F, sQ = (int64_t)g.Z * F, sM = (int64_t)g.Q * g.Z * F;
    for (int i = threadIdx.x; i < g.Z; i += blockDim.x) s_edges[i] = redz_edges[i];
    double accg[DET_FG], accn[DET_FG];      // this thread's bin (threadIdx.x + k*blockDim.x handled by the outer loop)
    for (int b0 = 0; b0 < Zb; b0 += blockDim.x) {
        const int mybin = b0 + threadIdx.x;
#pragma unroll
        for (int fi = 0; fi < DET_FG; ++fi) { accg[fi] = 0.0; accn[fi] = 0.0; }
        for (int qq = 0; qq < Qb; ++qq) {
            __syncthreads();
            for (int zz = threadIdx.x; zz < Zb; zz += blockDim.x) {
                const int64_t base = mm * sM + qq * sQ + zz * sZ + f0;
                const int64_t cell = (((int64_t)mm * Qb + qq) * Zb + zz) * F + f0;
                for (int fi = 0; fi < nf; ++fi) {
                    const double zc = corner_mean_redz(redz_final, sM, sQ, sZ, base + fi);
                    int bin = -1;
                    if (zc >= s_edges[0] && zc <= s_edges[g.Z - 1]) {
                        int lo = 0, hi = g.Z - 1;                 // largest lo with edges[lo] <= zc
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (s_edges[mid] <= zc) lo = mid; else hi = mid;
                        }
                        bin = lo;                                 // zc == last edge lands in the last bin
                    }
                    const double nv = number[cell + fi];
                    s_bin[fi * Zb + zz] = bin;
                    s_nv[fi * Zb + zz] = nv;
                    s_hv[fi * Zb + zz] = h2fdf[cell + fi] * nv;
                }
            }
            __syncthreads();
            if (mybin < Zb) {
                for (int fi = 0; fi < nf; ++fi) {
                    for (int zz = 0; zz < Zb; ++zz) {
                        if (s_bin[fi * Zb + zz] == mybin) {
                            accg[fi] += s_hv[fi * Zb + zz];
                            accn[fi] += s_nv[fi * Zb + zz];
                        }
                    }
                }
            }
        }
        if (mybin < Zb) {
            for (int fi = 0; fi < nf; ++fi) {
                const int64_t o = ((int64_t)mm * Zb + mybin) * F + f0 + fi;
                gwb_hist[o] = accg[fi];
                num_hist[o] = accn[fi];
            }
        }
    }
}

static int grid_for(int64_t n, int threads) {
    int64_t blocks = (n + threads - 1) / threads;
    int64_t cap = (int64_t)148 * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace holo

using namespace holo;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int holo_sam_density(const double* mtot, const double* mrat, const double* redz, const double* age_z,
                     const double* dtdz_z, int M, int Q, int Z, const holo_sam_params* par, const double* bf_tables,
                     double* dens, double* gmt_time, double* redz_prime, void* stream) {
    HOLO_REQUIRE(mtot && mrat && redz && age_z && dtdz_z && par && dens, "holo_sam_density: NULL argument");
    HOLO_REQUIRE(M > 0 && Q > 0 && Z > 0, "holo_sam_density: bad shape");
    HOLO_REQUIRE(par->bf_kind == 0 || par->bf_kind == 1, "holo_sam_density: unknown bulge-fraction kind");
    if (par->bf_kind == 0) HOLO_REQUIRE(par->mmb[3] > 0.0, "holo_sam_density: bulge fraction must be > 0");
    else HOLO_REQUIRE(bf_tables && par->bf_n > 0 && par->bf[0] > 0.0 && par->bf[1] >= par->bf[0] && par->bf[2] > 0.0,
                      "holo_sam_density: BF_Sigmoid needs its spline tables and parameters");
    int64_t n = (int64_t)M * Q * Z;
    density_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        mtot, mrat, redz, age_z, dtdz_z, M, Q, Z, *par, bf_tables, dens, gmt_time, redz_prime); holo::count_launches(1);
    return holo_check_launch("holo_sam_density");
}

int holo_zero_stalled(double* dens, const double* redz_prime, int64_t n, void* stream) {
    HOLO_REQUIRE(dens && redz_prime && n >= 0, "holo_zero_stalled: bad argument");
    if (n == 0) return HOLO_OK;
    zero_stalled_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dens, redz_prime, n); holo::count_launches(1);
    return holo_check_launch("holo_zero_stalled");
}

static int launch_norm(holo_cy_consts cc, double target_time, const double* mtot, const double* mrat,
                       int N, double sepa_init, double rchar, double gi, double go, int nsteps,
                       int mode, const double* norm_in, double* out, void* stream, const char* who) {
    HOLO_REQUIRE(mtot && mrat && out && N >= 0 && nsteps > 0, who);
    HOLO_REQUIRE(sepa_init > 0 && rchar > 0, who);
    if (N == 0) return HOLO_OK;
    size_t smem = (size_t)NORM_WARPS * 4 * (nsteps + 1) * sizeof(double);
    HOLO_REQUIRE(smem <= 200 * 1024, "2PL norm: num_steps too large for shared memory");
    if (smem > 48 * 1024)
        HOLO_CUDA(cudaFuncSetAttribute(norm_2pwl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (N + NORM_WARPS - 1) / NORM_WARPS;
    norm_2pwl_kernel<<<blocks, NORM_WARPS * 32, smem, (cudaStream_t)stream>>>(
        to_cc(cc), target_time, mtot, mrat, N, log10(sepa_init), rchar, gi, go, nsteps, mode, norm_in, out); holo::count_launches(1);
    return holo_check_launch(who);
}

int holo_find_2pwl_hardening_norm(holo_cy_consts cc, double target_time, const double* mtot,
                                  const double* mrat, int N, double sepa_init, double rchar,
                                  double gamma_inner, double gamma_outer, int nsteps,
                                  double* norm_log10, void* stream) {
    return launch_norm(cc, target_time, mtot, mrat, N, sepa_init, rchar, gamma_inner, gamma_outer,
                       nsteps, 0, nullptr, norm_log10, stream, "holo_find_2pwl_hardening_norm: bad argument");
}

int holo_binary_lifetime_2pwl(holo_cy_consts cc, const double* norm_log10, const double* mtot,
                              const double* mrat, int N, double sepa_init, double rchar,
                              double gamma_inner, double gamma_outer, int nsteps, double* lifetime,
                              void* stream) {
    HOLO_REQUIRE(norm_log10, "holo_binary_lifetime_2pwl: NULL norm");
    return launch_norm(cc, 0.0, mtot, mrat, N, sepa_init, rchar, gamma_inner, gamma_outer, nsteps, 1,
                       norm_log10, lifetime, stream, "holo_binary_lifetime_2pwl: bad argument");
}

int holo_hard_func_2pwl_gw(holo_cy_consts cc, const double* mtot, const double* mrat,
                           const double* sepa, const double* norm, double rchar, double gamma_inner,
                           double gamma_outer, int64_t N, double* dadt, void* stream) {
    HOLO_REQUIRE(mtot && mrat && sepa && norm && dadt && N >= 0, "holo_hard_func_2pwl_gw: bad argument");
    if (N == 0) return HOLO_OK;
    hard_func_kernel<<<grid_for(N, 256), 256, 0, (cudaStream_t)stream>>>(
        to_cc(cc), mtot, mrat, sepa, norm, rchar, gamma_inner, gamma_outer, N, dadt); holo::count_launches(1);
    return holo_check_launch("holo_hard_func_2pwl_gw");
}

int holo_dbn_2pwl(holo_cy_consts cc, const double* fobs_orb, int F, double sepa_init, int num_steps,
                  const double* hard_norm, double rchar, double gamma_inner, double gamma_outer,
                  const double* nden, const double* mtot, const double* mrat, const double* redz,
                  const double* gmt_time, int M, int Q, int Z, const double* grid_z,
                  const double* grid_dcom, const double* grid_age, int n_interp, double* redz_final,
                  double* diff_num, void* stream) {
    HOLO_REQUIRE(fobs_orb && hard_norm && nden && mtot && mrat && redz && gmt_time && grid_z &&
                 grid_dcom && grid_age && redz_final && diff_num, "holo_dbn_2pwl: NULL argument");
    HOLO_REQUIRE(M > 0 && Q > 0 && Z > 0 && F > 0 && num_steps > 0 && n_interp >= 2, "holo_dbn_2pwl: bad shape");
    HOLO_REQUIRE(num_steps < 65535, "holo_dbn_2pwl: num_steps must fit 16 bits");
    size_t smem = ((size_t)5 * (num_steps + 1) + 3 * (size_t)n_interp + Z + F) * sizeof(double) +
                  (size_t)Z * F * sizeof(unsigned short);
    HOLO_REQUIRE(smem <= 200 * 1024, "holo_dbn_2pwl: num_steps / table too large for shared memory");
    if (smem > 48 * 1024)
        HOLO_CUDA(cudaFuncSetAttribute(dbn_2pwl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dbn_2pwl_kernel<<<M * Q, DBN_THREADS, smem, (cudaStream_t)stream>>>(
        to_cc(cc), fobs_orb, F, log10(sepa_init), num_steps, hard_norm, rchar, gamma_inner,
        gamma_outer, nden, mtot, mrat, redz, gmt_time, M, Q, Z, grid_z, grid_dcom, grid_age, n_interp,
        redz_final, diff_num); holo::count_launches(1);
    return holo_check_launch("holo_dbn_2pwl");
}

int holo_dbn_gw(holo_cy_consts cc, const double* fobs_orb, int F, const double* nden,
                const double* mtot, const double* mrat, const double* redz, const double* redz_prime,
                int M, int Q, int Z, const double* grid_z, const double* grid_dcom, int n_interp,
                double* redz_final, double* diff_num, void* stream) {
    (void)redz;
    HOLO_REQUIRE(fobs_orb && nden && mtot && mrat && redz_prime && grid_z && grid_dcom && redz_final &&
                 diff_num, "holo_dbn_gw: NULL argument");
    HOLO_REQUIRE(M > 0 && Q > 0 && Z > 0 && F > 0 && n_interp >= 2, "holo_dbn_gw: bad shape");
    int64_t n = (int64_t)M * Q * Z * F;
    dbn_gw_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        to_cc(cc), fobs_orb, F, nden, mtot, mrat, redz_prime, M, Q, Z, grid_z, grid_dcom, n_interp,
        redz_final, diff_num); holo::count_launches(1);
    return holo_check_launch("holo_dbn_gw");
}

int holo_integrate_differential_number_3dx1d(const double* log10_mtot, const double* mrat,
                                             const double* redz, const double* dln_freq,
                                             const double* dnum, double* numb, int M, int Q, int Z,
                                             int F, void* stream) {
    HOLO_REQUIRE(log10_mtot && mrat && redz && dln_freq && dnum && numb, "holo_integrate: NULL argument");
    HOLO_REQUIRE(M > 0 && Q > 0 && Z > 0 && F >= 0, "holo_integrate: bad shape");
    int64_t n = (int64_t)(M - 1) * (Q - 1) * (Z - 1) * F;
    if (n <= 0) return HOLO_OK;
    BinGeom g{M, Q, Z, F};
    GLTable gl{};
    bin_kernel<true, false><<<(M - 1) * (Q - 1), 256, 0, (cudaStream_t)stream>>>(
        g, gl, 0, 0, 0, 0, log10_mtot, mrat, redz, dln_freq, dnum, nullptr, nullptr, nullptr, nullptr,
        nullptr, nullptr, numb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.0); holo::count_launches(1);
    return holo_check_launch("holo_integrate_differential_number_3dx1d");
}

static GLTable to_gl(const holo_cosmo_params* c) {
    GLTable gl;
    for (int i = 0; i < GL_ORDER; ++i) { gl.x[i] = c->gl_x[i]; gl.w[i] = c->gl_w[i]; }
    return gl;
}

int holo_char_strain_sq(const holo_cosmo_params* cosmo, double gw_src_const, double nwtg,
                        const double* redz_final, const double* rz_mid, const double* mt_mid,
                        const double* mr_mid, const double* fc, const double* fc_over_df, int M,
                        int Q, int Z, int F, double* h2fdf, double* zmid, double* dcom, double* sepa,
                        double* angs, int32_t* bad_redz, const double* dc_table, int dc_n, double dc_wmax,
                        void* stream) {
    HOLO_REQUIRE(dc_table == nullptr || (dc_n > 0 && dc_wmax > 0.0), "holo_char_strain_sq: bad distance table");
    HOLO_REQUIRE(cosmo && (redz_final || rz_mid) && mt_mid && mr_mid && fc && fc_over_df && h2fdf,
                 "holo_char_strain_sq: NULL argument");
    HOLO_REQUIRE(M > 1 && Q > 1 && Z > 1 && F > 0, "holo_char_strain_sq: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    BinGeom g{M, Q, Z, F};
    bin_kernel<false, true><<<(M - 1) * (Q - 1), 256, 0, st>>>(
        g, to_gl(cosmo), cosmo->hubble_distance, cosmo->om0, gw_src_const, nwtg, nullptr, nullptr,
        nullptr, nullptr, nullptr, redz_final, rz_mid, mt_mid, mr_mid, fc, fc_over_df, nullptr, h2fdf,
        zmid, dcom, sepa, angs, bad_redz, dc_table, dc_n, dc_table ? dc_n / dc_wmax : 0.0); holo::count_launches(1);
    return holo_check_launch("holo_char_strain_sq");
}

int holo_integrate_and_strain(const holo_cosmo_params* cosmo, double gw_src_const, double nwtg,
                              const double* log10_mtot, const double* mrat, const double* redz,
                              const double* dln_freq, const double* dnum, const double* redz_final,
                              const double* mt_mid, const double* mr_mid, const double* fc,
                              const double* fc_over_df, int M, int Q, int Z, int F, double* numb,
                              double* h2fdf, double* zmid, double* dcom, double* sepa, double* angs,
                              int32_t* bad_redz, const double* dc_table, int dc_n, double dc_wmax, void* stream) {
    HOLO_REQUIRE(dc_table == nullptr || (dc_n > 0 && dc_wmax > 0.0), "holo_integrate_and_strain: bad distance table");
    HOLO_REQUIRE(cosmo && log10_mtot && mrat && redz && dln_freq && dnum && redz_final && mt_mid &&
                 mr_mid && fc && fc_over_df && numb && h2fdf, "holo_integrate_and_strain: NULL argument");
    HOLO_REQUIRE(M > 1 && Q > 1 && Z > 1 && F > 0, "holo_integrate_and_strain: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    BinGeom g{M, Q, Z, F};
    bin_kernel<true, true><<<(M - 1) * (Q - 1), 256, 0, st>>>(
        g, to_gl(cosmo), cosmo->hubble_distance, cosmo->om0, gw_src_const, nwtg, log10_mtot, mrat, redz,
        dln_freq, dnum, redz_final, nullptr, mt_mid, mr_mid, fc, fc_over_df, numb, h2fdf, zmid, dcom,
        sepa, angs, bad_redz, dc_table, dc_n, dc_table ? dc_n / dc_wmax : 0.0); holo::count_launches(1);
    return holo_check_launch("holo_integrate_and_strain");
}

int64_t holo_gwb_expectation_workspace_bytes(int F) {
    return (int64_t)sizeof(double) * (int64_t)(148 * 4) * (F > 0 ? F : 1);
}

int holo_gwb_expectation(const double* number, const double* h2fdf, int64_t ncell, int F, double* hc2,
                         void* workspace, int64_t workspace_bytes, void* stream) {
    HOLO_REQUIRE(number && h2fdf && hc2 && workspace && ncell >= 0 && F > 0, "holo_gwb_expectation: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int nblk = 148 * 4;
    int64_t cpb = (ncell + nblk - 1) / nblk;
    if (cpb < 1) cpb = 1;
    nblk = (int)((ncell + cpb - 1) / cpb);
    if (nblk < 1) nblk = 1;
    // per-block partial sums live in the CALLER's workspace: the library itself never allocates device memory
    HOLO_REQUIRE((int64_t)sizeof(double) * nblk * F <= workspace_bytes, "holo_gwb_expectation: workspace too small");
    double* partial = static_cast<double*>(workspace);
    expect_partial_kernel<<<nblk, 32 * EXP_ROWS, 0, st>>>(number, h2fdf, ncell, F, cpb, partial); holo::count_launches(1);
    expect_final_kernel<<<(F + 63) / 64, 64, 0, st>>>(partial, nblk, F, hc2); holo::count_launches(1);
    return holo_check_launch("holo_gwb_expectation");
}

int holo_model_details_hist(const double* redz_edges, const double* redz_final, const double* number,
                            const double* h2fdf, int M, int Q, int Z, int F, double* gwb_hist, double* num_hist,
                            void* stream) {
    HOLO_REQUIRE(redz_edges && redz_final && number && h2fdf && gwb_hist && num_hist, "holo_model_details_hist: NULL argument");
    HOLO_REQUIRE(M > 1 && Q > 1 && Z > 1 && F > 0, "holo_model_details_hist: bad shape");
    BinGeom g{M, Q, Z, F};
    const size_t smem = sizeof(double) * ((size_t)Z + 2 * DET_FG * (size_t)(Z - 1)) + sizeof(int) * DET_FG * (size_t)(Z - 1);
    HOLO_REQUIRE(smem <= 48 * 1024, "holo_model_details_hist: too many redshift bins for shared memory");
    dim3 grid(M - 1, (F + DET_FG - 1) / DET_FG);
    details_hist_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(g, redz_edges, redz_final, number, h2fdf, gwb_hist, num_hist);
    holo::count_launches(1);
    return holo_check_launch("holo_model_details_hist");
}

}  // extern "C"
