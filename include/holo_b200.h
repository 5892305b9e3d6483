/* holo_b200.h -- C ABI of libholo_b200.so: B200 (sm_100a) kernels for holodeck's SAM GW-background path.
 *
 * Drop-in boundary.  The reference has no FFI table: its "operator API" is
 * the set of module-level callables of its two Cython extension modules,
 *     holodeck/sams/sam_cyutils.pyx   and   holodeck/cyutils.pyx ,
 * built by the reference's setup.py:32-73.  Each entry point below replaces one of those Cython
 * `cdef` loops (file:line cited per function).  The thin Python shims that keep the reference's
 * call signatures live in holodeck_b200/sams/sam_cyutils.py and holodeck_b200/cyutils.py and bind
 * these symbols with ctypes (see INTEGRATION.md for the stub a holodeck maintainer would add).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (cudaMalloc'ed by the caller) unless the name ends in
 *    `_host`; the library never allocates or frees caller-visible memory (scratch space is passed in);
 *  - all floating point data is IEEE double, C-contiguous, shapes in comments use the reference's
 *    names: M,Q,Z = number of mtot/mrat/redz grid EDGES, F = number of frequency bins,
 *    R = realizations, L = loudest sources;
 *  - `stream` is a cudaStream_t (pass NULL for the legacy default stream); calls are asynchronous
 *    with respect to the host unless noted;
 *  - every function returns 0 on success, non-zero on failure; `holo_last_error()` then returns a
 *    static, thread-local message.  There is no CPU fallback: without a usable sm_100 device the
 *    calls fail with HOLO_ERR_CUDA.
 */
#ifndef HOLO_B200_H
#define HOLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOLO_OK 0
#define HOLO_ERR_ARG 1
#define HOLO_ERR_CUDA 2
#define HOLO_ERR_OVERFLOW 3 /* loudest-source event buckets overflowed or head too short: retry bigger */

#define HOLO_ABI_VERSION 3

int holo_abi_version(void);
const char* holo_last_error(void);
/* Number of CUDA devices visible (<=0: none). */
int holo_device_count(void);
/* Running count of kernels this library has launched in this process (bench.py `gpu_launches`). */
int64_t holo_launch_count(void);
/* Optional per-kernel timing of the multi-kernel entry points (holo_loudest, holo_sam_poisson_gwb):
 * when enabled they record CUDA events around their internal kernels on the caller's stream and
 * synchronise before returning.  holo_get_profile copies the last call's stage durations [ms] --
 * holo_loudest: {head preparation, draw kernel, resolve, final reduce}; holo_sam_poisson_gwb:
 * {draw kernel, final reduce} -- and returns how many entries are valid.  Used by bench.py only. */
void holo_set_profiling(int on);
int holo_get_profile(double* ms, int n);

/* Constants the reference computes once at import with libm (sam_cyutils.pyx:36-42).  The host
 * computes them the same way (libm pow/sqrt) and passes them in, so device code sees identical bits. */
typedef struct {
    double gw_dadt_sep_const;  /* -64 G^3/(5 c^5)        */
    double kepler_const_freq;  /* sqrt(G)/(2 pi)         */
    double kepler_const_sepa;  /* G^(1/3)/(2 pi)^(2/3)   */
    double four_pi_c_over_mpc; /* 4 pi c / Mpc           */
} holo_cy_consts;

/* Flat-LCDM parameters + Gauss-Legendre nodes for the comoving-distance quadrature (the host twin
 * is holodeck_b200/cosmology.py; replaces astropy `cosmo.comoving_distance`, gravwaves.py:718). */
#define HOLO_GL_ORDER 16
typedef struct {
    double hubble_distance; /* c/H0 [cm] */
    double hubble_time;     /* 1/H0 [s]  */
    double om0;
    double age_universe;    /* [s] */
    double gl_x[HOLO_GL_ORDER];
    double gl_w[HOLO_GL_ORDER];
} holo_cosmo_params;

/* ---------------------------------------------------------------------------------------------
 * K0  static binary density.  Replaces the numpy chain of
 *     Semi_Analytic_Model.static_binary_density   holodeck/sams/sam.py:280-365
 * with the component callables fused in: GSMF_Schechter / GSMF_Double_Schechter
 * (sams/components.py:132-172, 315-329), GPF_Power_Law (:556-583), GMT_Power_Law (:646-675) +
 * zprime (:620-626) + utils.redz_after (utils.py:1772-1808), GMR_Illustris (:452-482),
 * MMBulge_Standard + BF_Constant / BF_Sigmoid (host_relations.py:745-771, 720-743, 483-512, 190-195, 198-331).
 * Outputs are the PRE-scatter density, the galaxy-merger time and z' (all (M,Q,Z)).
 * `holo_zero_stalled` then applies sam.py:392-394.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    int gsmf_kind;      /* 0 = GSMF_Schechter, 1 = GSMF_Double_Schechter */
    int use_gmr;        /* 1: GMR_Illustris gives the merger rate; 0: GPF/GMT */
    int has_gmt;        /* a GMT instance exists (z' and stalling are computed) */
    int gsmf_uses_mtot; /* sam.py:57-59 module switches */
    int gpf_uses_mtot;
    int gmt_uses_mtot;
    int bf_kind;        /* bulge fraction: 0 = BF_Constant (mmb[3]), 1 = BF_Sigmoid (bf[], spline tables) */
    int bf_n;           /* BF_Sigmoid: pieces per spline table */
    double gsmf[12];    /* kind 0: phi0, phiz, mchar0[g], mcharz, alpha0, alphaz
                           kind 1: phi1[3], phi2[3], log10_mstar[3], alpha1, alpha2, MSOL */
    double gpf[6];      /* frac_norm, mref[g], malpha, zbeta, qgamma, max_frac */
    double gmt[5];      /* time_norm[s], mref[g], malpha, zbeta, qgamma */
    double gmr[11];     /* norm0[1/s], normz, malpha0, malphaz, mdelta0, mdeltaz, qgamma0, qgammaz,
                           qgammam, mref[g], mref_delta[g] */
    double mmb[4];      /* mamp[g], mplaw, mref[g], bulge_frac */
    double hubble_time; /* [s] */
    double om0;
    double age_universe; /* [s]: utils._AGE_UNIVERSE_GYR * GYR, utils.py:46 */
    double bf[4];       /* BF_Sigmoid (host_relations.py:198-331): frac_lo, frac_hi, mstar_char[g], width_dex */
} holo_sam_params;

/* `bf_tables` (device; NULL unless bf_kind == 1): the two quadratic interpolants BF_Sigmoid inverts its relation with
 * (scipy interp1d(kind='quadratic') of mstar(mbulge) and dmstar/dmbulge(mbulge), host_relations.py:276-283) as
 * piecewise polynomials, each [breaks (bf_n + 1) | c0 (bf_n) | c1 (bf_n) | c2 (bf_n)], value = c0 dx^2 + c1 dx + c2. */
int holo_sam_density(const double* mtot, const double* mrat, const double* redz,
                     const double* age_z /* (Z,) cosmo.age(redz) [s] */,
                     const double* dtdz_z /* (Z,) cosmo.dtdz(redz) [s] */,
                     int M, int Q, int Z, const holo_sam_params* par_host, const double* bf_tables,
                     double* dens, double* gmt_time /* may be NULL if !has_gmt */,
                     double* redz_prime /* may be NULL if !has_gmt */, void* stream);

/* dens[i] = 0 where redz_prime[i] < 0   (sam.py:328, 392-394) */
int holo_zero_stalled(double* dens, const double* redz_prime, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1a  Fixed_Time_2PL_SAM normalisation.  Replaces
 *      find_2pwl_hardening_norm / _get_hardening_norm_2pwl / get_binary_lifetime_2pwl
 *      holodeck/sams/sam_cyutils.pyx:289-398  (scipy `brentq`, xtol=1e-3, rtol=1e-5, maxiter=100,
 *      bracket [-20, 20] in log10-norm).
 * --------------------------------------------------------------------------------------------- */
int holo_find_2pwl_hardening_norm(holo_cy_consts cc, double target_time, const double* mtot,
                                  const double* mrat, int N, double sepa_init, double rchar,
                                  double gamma_inner, double gamma_outer, int nsteps,
                                  double* norm_log10, void* stream);

/* integrate_binary_evolution_2pwl (sam_cyutils.pyx:401-413), vectorised over N binaries:
 * lifetime[i] = sum over nsteps of trapezoid dt, for the given log10-norm. */
int holo_binary_lifetime_2pwl(holo_cy_consts cc, const double* norm_log10, const double* mtot,
                              const double* mrat, int N, double sepa_init, double rchar,
                              double gamma_inner, double gamma_outer, int nsteps, double* lifetime,
                              void* stream);

/* hard_func_2pwl_gw (sam_cyutils.pyx:256-286) on N flattened, pre-broadcast elements. */
int holo_hard_func_2pwl_gw(holo_cy_consts cc, const double* mtot, const double* mrat,
                           const double* sepa, const double* norm, double rchar, double gamma_inner,
                           double gamma_outer, int64_t N, double* dadt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1b  _dynamic_binary_number_at_fobs_2pwl   holodeck/sams/sam_cyutils.pyx:510-781
 * K1c  _dynamic_binary_number_at_fobs_gw     holodeck/sams/sam_cyutils.pyx:788-899
 * Outputs redz_final, diff_num are (M,Q,Z,F); every element is written (sentinels -1 / 0 included).
 * grid_z is DECREASING and ends at 0, grid_age INCREASING (cosmo._grid_z/_grid_dcom/_grid_age).
 * --------------------------------------------------------------------------------------------- */
int holo_dbn_2pwl(holo_cy_consts cc, const double* fobs_orb, int F, double sepa_init, int num_steps,
                  const double* hard_norm /* (M,Q) */, double rchar, double gamma_inner,
                  double gamma_outer, const double* nden /* (M,Q,Z) */, const double* mtot,
                  const double* mrat, const double* redz, const double* gmt_time /* (M,Q,Z) */,
                  int M, int Q, int Z, const double* grid_z, const double* grid_dcom,
                  const double* grid_age, int n_interp, double* redz_final, double* diff_num,
                  void* stream);

int holo_dbn_gw(holo_cy_consts cc, const double* fobs_orb, int F, const double* nden,
                const double* mtot, const double* mrat, const double* redz,
                const double* redz_prime /* (M,Q,Z) */, int M, int Q, int Z, const double* grid_z,
                const double* grid_dcom, int n_interp, double* redz_final, double* diff_num,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2   _integrate_differential_number_3dx1d   holodeck/sams/sam_cyutils.pyx:170-216
 *      dnum (M,Q,Z,F) -> numb (M-1,Q-1,Z-1,F); log10_mtot (M,), mrat (Q,), redz (Z,), dln_freq (F,)
 * --------------------------------------------------------------------------------------------- */
int holo_integrate_differential_number_3dx1d(const double* log10_mtot, const double* mrat,
                                             const double* redz, const double* dln_freq,
                                             const double* dnum, double* numb, int M, int Q, int Z,
                                             int F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2b  char_strain_sq_from_bin_edges_redz     holodeck/gravwaves.py:694-725
 *      (and the params=True glue of ss_gws_redz, holodeck/single_sources.py:112-139).
 *  redz_final (M,Q,Z,F) at grid edges -> h2fdf (M-1,Q-1,Z-1,F).
 *  mt_mid (M-1,), mr_mid (Q-1,) are bin-centre masses / ratios; fc (F,) bin-centre orbital
 *  frequencies; fc_over_df (F,) = fc/diff(edges).
 *  Optional outputs (NULL to skip), each (M-1,Q-1,Z-1,F): zmid (cell-centre redshift, -1 where <=0),
 *  dcom (cm, +inf where z<=0), sepa (cm), angs (rad).
 *  `gw_src_const`/`nwtg` are the astropy-valued constants of utils.py:40 / constants.py:23.
 *  If `redz_final` is NULL, `rz_mid` (Z-1,) is used for all cells: char_strain_sq_from_bin_edges,
 *  gravwaves.py:760-783.
 *  `dc_table` (optional, device): the comoving distance as a table instead of a 16-point quadrature per cell:
 *  2 (dc_n + 1) doubles, pairs (R_i, h R'_i) at the uniform nodes w_i = i dc_wmax / dc_n of
 *  R(w) = (1/w) int_{1-w}^{1} 2 ds / sqrt(Om0 + (1-Om0) s^6),  w = 1 - (1+z)^(-1/2),  d_c = hubble_distance w R(w)
 *  (cubic Hermite; the quadrature remains the fallback beyond the table).
 *  `bad_redz` (optional, one zero-initialised int32 on the device) is set to 1 when some redz_final value is
 *  negative but not the -1 sentinel -- the input check of single_sources.py:95-99, made while the values are
 *  being read instead of in separate passes over the grid.
 * --------------------------------------------------------------------------------------------- */
int holo_char_strain_sq(const holo_cosmo_params* cosmo_host, double gw_src_const, double nwtg,
                        const double* redz_final, const double* rz_mid, const double* mt_mid,
                        const double* mr_mid, const double* fc, const double* fc_over_df, int M,
                        int Q, int Z, int F, double* h2fdf, double* zmid, double* dcom, double* sepa,
                        double* angs, int32_t* bad_redz, const double* dc_table, int dc_n, double dc_wmax,
                        void* stream);

/* Fused K2 + K2b for the sam.gwb pipeline: one pass over (diff_num, redz_final) producing
 * (number, h2fdf [, zmid, dcom, sepa, angs]). */
int holo_integrate_and_strain(const holo_cosmo_params* cosmo_host, double gw_src_const, double nwtg,
                              const double* log10_mtot, const double* mrat, const double* redz,
                              const double* dln_freq, const double* dnum, const double* redz_final,
                              const double* mt_mid, const double* mr_mid, const double* fc,
                              const double* fc_over_df, int M, int Q, int Z, int F, double* numb,
                              double* h2fdf, double* zmid, double* dcom, double* sepa, double* angs,
                              int32_t* bad_redz, const double* dc_table, int dc_n, double dc_wmax,
                              void* stream);

/* hc2[f] = sum_{m,q,z} number * h2fdf   (realize=False branch, gravwaves.py:481-485, 557-561) */
int64_t holo_gwb_expectation_workspace_bytes(int F);
int holo_gwb_expectation(const double* number, const double* h2fdf, int64_t ncell, int F,
                         double* hc2 /* (F,) */, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3/K4  realised GWB and loudest-source split.
 *
 *  Random numbers: counter-based Philox4x32-10 keyed on `seed`; the counter is
 *  (flat (cell,f) index, global realization index, stream id, retry) so results do not depend on how
 *  realizations are partitioned over launches / GPUs (`r0` = global index of local realization 0).
 *  If `counts` is non-NULL no random numbers are drawn: the kernel reads the supplied draw for
 *  (local realization r, frequency f, cell c) at counts[((int64)r*F + f)*ncell + c]  (doubles, so
 *  that un-floored normal-approximation draws can be supplied too) -- "supplied-count mode".
 *
 *  number, h2fdf are (ncell, F) with ncell = (M-1)(Q-1)(Z-1) in the reference's C order.
 * --------------------------------------------------------------------------------------------- */

/* _sam_poisson_gwb   holodeck/cyutils.pyx:862-897 :  gwb (F,R) */
int holo_sam_poisson_gwb(const double* number, const double* h2fdf, int64_t ncell, int F, int R,
                         int64_t r0, uint64_t seed, double normal_threshold, const double* counts,
                         double* gwb, void* workspace, int64_t workspace_bytes, void* stream);

/* Variants of the loudest-source kernels. */
#define HOLO_LOUDEST_PLAIN 1     /* _loudest_hc_from_sorted              cyutils.pyx:1266-1344 */
#define HOLO_LOUDEST_PAR 2       /* _loudest_hc_and_par_from_sorted      cyutils.pyx:1409-1538 */
#define HOLO_LOUDEST_PAR_REDZ 3  /* _loudest_hc_and_par_from_sorted_redz cyutils.pyx:1615-1767 */

typedef struct {
    int variant;            /* HOLO_LOUDEST_* */
    int Mb, Qb, Zb, F;      /* BIN counts (M-1, Q-1, Z-1) and frequencies */
    int R, L;
    int64_t r0;
    uint64_t seed;
    double normal_threshold;
    const double* number;   /* (Mb,Qb,Zb,F) */
    const double* h2fdf;    /* (Mb,Qb,Zb,F) */
    const int32_t* order;   /* (ncell,) flat cell index (m*Qb+q)*Zb+z, loudest first == msort/qsort/zsort */
    const double* mt;       /* (Mb,) variants 2,3 */
    const double* mr;       /* (Qb,) */
    const double* rz;       /* (Zb,) */
    const double* redz_final; /* (Mb,Qb,Zb,F) variant 3 */
    const double* dcom_final;
    const double* sepa;
    const double* angs;
    const double* counts;   /* supplied-count mode, see above; else NULL */
    /* outputs */
    double* hc2ss;          /* (F,R,L) */
    double* hc2bg;          /* (F,R) */
    double* sspar;          /* variant 3: (4,F,R,L) */
    double* bgpar;          /* variant 2: (3,F,R); variant 3: (7,F,R) */
    double* lspar;          /* variant 2: (3,F,R) */
    int64_t* ssidx;         /* variant 2: (3,F,R,L) */
    /* scratch */
    void* workspace;
    int64_t workspace_bytes;
    int bucket_cap;         /* event-bucket capacity per (f,r); 0 = choose automatically */
    double head_margin;     /* expected occupied head cells beyond L; <=0 = automatic */
    /* Optional fused product (ABI version 2).  lib_tools.run_model (librarian/lib_tools.py:801-832) asks for the
     * loudest split AND an independently drawn realised GWB of the same grid; when `gwb` is non-NULL the same
     * pass over the grid also draws `gwb_R` background-only realizations (own seed / global offset, stream of
     * holo_sam_poisson_gwb) and writes gwb (F, gwb_R).  The workspace then needs
     * holo_loudest_workspace_bytes(..) + holo_realize_workspace_bytes(HOLO_REALIZE_GWB, ncell, F, gwb_R) bytes.
     * Not available in supplied-count mode. */
    double* gwb;
    int gwb_R;
    int64_t gwb_r0;
    uint64_t gwb_seed;
    /* Deferred overflow check (ABI version 3).  By default `holo_loudest` reads its two overflow flags (event bucket
     * overflow / head too short) back and is therefore synchronous.  With `defer_check` != 0 it returns as soon as the
     * kernels are enqueued; the caller reads `int32 flags[2]` from the FIRST 8 bytes of `workspace` once the stream has
     * got there and, if either is set, repeats the call with a larger `head_margin` / `bucket_cap`.  This lets a
     * driver prepare and enqueue the next model while this one is still drawing (librarian/gen_lib.py). */
    int defer_check;
} holo_loudest_args;

/* Bytes of scratch `holo_loudest` needs for these sizes (bucket_cap 0 = automatic). */
int64_t holo_loudest_workspace_bytes(int variant, int64_t ncell, int F, int R, int L, int bucket_cap);
/* Synchronous with respect to `stream` (it checks an overflow flag before returning) unless `defer_check` is set. */
int holo_loudest(const holo_loudest_args* args_host, void* stream);

/* _ss_bg_hc (cyutils.pyx:935-1014) and _ss_bg_hc_and_par (:1017-1178): L=1, arg-max by value.
 * ssidx (3,F,R) int64; bgpar/sspar (3,F,R) or NULL. */
int holo_ss_bg_hc(const double* number, const double* h2fdf, int Mb, int Qb, int Zb, int F, int R,
                  int64_t r0, uint64_t seed, double normal_threshold, const double* counts,
                  const double* mt, const double* mr, const double* rz, double* hc2ss, double* hc2bg,
                  int64_t* ssidx, double* bgpar, double* sspar, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* Scratch bytes for holo_sam_poisson_gwb (kind 0) / holo_ss_bg_hc without (4) or with (5) parameters. */
#define HOLO_REALIZE_GWB 0
#define HOLO_REALIZE_SSBG 4
#define HOLO_REALIZE_SSBG_PAR 5
int64_t holo_realize_workspace_bytes(int kind, int64_t ncell, int F, int R);

/* poisson_as_needed (gravwaves.py:666-691) / Realizer_SAM bulk draws: out[c] ~ Poisson(lam[c])
 * (floor(Normal) above the threshold).  Flat arrays of n elements. */
int holo_poisson_as_needed(const double* lam, int64_t n, uint64_t seed, uint64_t stream_id,
                           double normal_threshold, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  eccentric harmonic sum.
 *     _sam_calc_gwb_single_eccen           holodeck/cyutils.pyx:370-597   -> gwb (F,H)
 *     _sam_calc_gwb_single_eccen_discrete  holodeck/cyutils.pyx:609-851   -> gwb (F,H,R)
 * --------------------------------------------------------------------------------------------- */
int holo_sam_calc_gwb_single_eccen(holo_cy_consts cc /* cyutils.pyx:47 GW_DADT_SEP_CONST */,
                                   double gw_src_const /* cyutils.pyx:48, libm-evaluated by the host */,
                                   const double* ndens /* (M,Q,Z) */, const double* mtot_log10,
                                   const double* mrat, const double* redz, const double* dcom_mpc,
                                   const double* gwfobs, const double* sepa_evo,
                                   const double* eccen_evo, int M, int Q, int Z, int F, int E,
                                   int nharms, int nreals /* 0: continuous */, int64_t r0,
                                   uint64_t seed, double* gwb, void* workspace,
                                   int64_t workspace_bytes, void* stream);
int64_t holo_eccen_workspace_bytes(int M, int Q, int Z, int F, int nharms, int nreals);

/* ---------------------------------------------------------------------------------------------
 * K7  Library "details" (SURVEY 8f N4).  Replaces the 2 x F x (1 + (M-1)) scipy.stats.binned_statistic calls of
 *      lib_tools._calc_model_details  holodeck/librarian/lib_tools.py:904-939:
 *      gwb_hist[m, b, f] = sum over (q, z) cells of mass bin m whose cell-centre final redshift falls in
 *      redshift bin b of  h2fdf * number ;  num_hist the same for `number`  (bins [e_b, e_b+1), last closed).
 *      The marginals of lib_tools.py:876-900 are plain axis sums (host driver, torch.sum).
 * ------------------------------------------------------------------------------------------- */
int holo_model_details_hist(const double* redz_edges /* (Z,) */, const double* redz_final /* (M,Q,Z,F) */,
                            const double* number /* (M-1,Q-1,Z-1,F) */, const double* h2fdf,
                            int M, int Q, int Z, int F, double* gwb_hist /* (M-1,Z-1,F) */,
                            double* num_hist, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  M-Mbulge scatter of the binary density.  Replaces the per-redshift scipy pipeline of
 *      add_scatter_to_masses  holodeck/sams/sam.py:1291-1394 (with utils.py:416-488):
 *      CloughTocher2DInterpolator + NearestNDInterpolator fill -> two dense scatter products ->
 *      RegularGridInterpolator(linear).  The geometry (Delaunay triangulation of the (log10 m1, log10 m2)
 *      images of the grid, point location, level schedule of the Gauss-Seidel sweep) is data-independent and
 *      prepared once per grid by the host driver (holodeck_b200/sams/scatter.py); the scatter products are plain
 *      DGEMMs (cuBLAS) between `holo_scatter_ct_eval` and `holo_scatter_bilinear`.
 *      Layouts: density (npts = M*Q, Z) z fastest; gradients (npts, 2, Z); regular grid (G, G, Z).
 * ------------------------------------------------------------------------------------------- */

/* gradients at the triangulation vertices: scipy interpnd `_estimate_gradients_2d_global` (Gauss-Seidel,
 * maxiter / tol as CloughTocher2DInterpolator: 400, 1e-6), all Z slices at once.
 * `program` (16-byte aligned, nsteps records of holo_scatter_step_bytes() bytes): the level schedule of the sweep
 * flattened by the host into one fixed-size, plane-ordered record per step (32 vertices of one level, 8 neighbours
 * each in Delaunay.vertex_neighbor_vertices order; 128 consumer threads t = 4 slot + sub, each owning neighbours `sub`
 * and `sub + 4` of the vertex in `slot`):
 *     double e[4][128][2]   a: (2 ex, 2 ey) | a: (ex, ey)/L^3 | b: (2 ex, 2 ey) | b: (ex, ey)/L^3
 *     double qinv[2][32][2] the two rows of MINUS the inverse of the vertex's 2x2 normal matrix
 *     int    ids[128][4]    a's neighbour, b's neighbour, the vertex (-1: empty slot), flags (bit 0: first round of the
 *                           step's vertices, bit 1: last round -> apply the update)
 * An absent edge points at the vertex itself (vertex 0 in an empty slot) with zero coefficients.  The kernel streams the
 * records through shared memory with TMA bulk copies (cp.async.bulk + mbarrier) issued by a producer warp.
 * niter (Z,) (may be NULL): sweeps used, 0 = not converged.  The regular-grid values the back-interpolation does not
 * read need not be computed: the host driver evaluates the scatter product for the needed row/column blocks only. */
int holo_scatter_step_bytes(void);
int holo_scatter_gradients(int npts, int Z, const void* program, int nsteps, const double* data, int maxiter,
                           double tol, double* grad, int* niter, void* stream);

/* Clough-Tocher values at `ngrid` regular-grid points (geometry records of holo_scatter_geo_bytes() bytes
 * each: simplex, nearest vertex, vertices, barycentric coordinates, edge vectors, g[3]); NaN (outside the
 * hull) and negative values take the nearest vertex's value (sam.py:1370-1375); flags[0] != 0 if a bad
 * value survives (sam.py:1376-1380 raises ValueError). */
int holo_scatter_ct_eval(int64_t ngrid, int Z, const void* geo, const double* data, const double* grad,
                         double* out, int* flags, void* stream);

/* RegularGridInterpolator(method='linear') from the (G, G, Z) grid back to the npts data points:
 * i0/i1 lower cell indices, y0/y1 normalised distances. */
int holo_scatter_bilinear(int npts, int G, int Z, const int* i0, const int* i1, const double* y0,
                          const double* y1, const double* grid, double* out, void* stream);

int holo_scatter_geo_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* HOLO_B200_H */
