/* pyflation_b200.h -- C ABI of the B200-native first-order hot path of ihuston/pyflation.
 *
 * The library replaces, for CanonicalBackground / CanonicalFirstOrder models, what the
 * reference does in pure Python/NumPy inside
 *     pyflation/rk4.py:58-138          rkdriver_tsix   (the time loop, NaN prefill, IC scatter)
 *     pyflation/rk4.py:27-55           rk4stepks       (one classical RK4 step)
 *     pyflation/cosmomodels.py:551-571 CanonicalBackground.derivs
 *     pyflation/cosmomodels.py:625-675 CanonicalFirstOrder.derivs
 *     pyflation/cmpotentials.py:16-1144  the 18 potential functions (U, dU, d2U)
 *     pyflation/analysis/adiabatic.py:153-285  Pr / scaled_Pr (pf_pr_spectrum)
 *     pyflation/cosmomodels.py:678-971 CanonicalSecondOrder / RampedSecondOrder.derivs (pf_so_run)
 * The reference has no FFI of its own (a Python callable `derivs` is passed to the solver),
 * so the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md; the host-side
 * mirror of the reference's modules that does this lives in pyflation_b200/{rk4,cosmomodels}.py.
 *
 * Conventions
 *   - plain pointers and sizes only; `*_dev` pointers are device memory of the CURRENT CUDA
 *     device, `*_host` pointers are host memory.  The caller owns every buffer; the library
 *     allocates nothing that outlives a call and keeps no state except a thread-local error
 *     string.  `stream` is a cudaStream_t passed as void* (NULL = default stream).  All
 *     launches are asynchronous on `stream`.
 *   - complex128 buffers are arrays of (re, im) double pairs, NumPy layout.
 *   - trajectory layout is the reference's: yresult[t][var][k], k innermost (rk4.py:95-100),
 *     var order  phi_0, phi'_0, ..., phi_{nf-1}, phi'_{nf-1}, H, dphi_00, dphi'_00, dphi_01, ...
 *     (cosmomodels.py:613-623);  nvar = 2 nf^2 + 2 nf + 1.
 *   - return value: 0 = OK; < 0 = argument error (PF_E_*), mirrors rk4.SimRunError /
 *     cosmomodels.ModelError on the Python side; > 0 = cudaError_t.  pf_last_error() gives text.
 */
#ifndef PYFLATION_B200_H
#define PYFLATION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 2
#define PF_MAX_NFIELDS 8          /* compiled instantiations (state in registers): nf = 1..8 */
#define PF_MAX_NFIELDS_RT 64      /* runtime-nf kernels: nflation, the one potential of the reference
                                     defined for any nfields (cmpotentials.py:791-846), nf = 9..64 */
#define PF_MAX_PARAMS 8

/* potential ids, in the order of pyflation/cmpotentials.py */
enum pf_potential {
    PF_MSQPHISQ = 0,          /* cmpotentials.py:16    params: mass */
    PF_LAMBDAPHI4 = 1,        /* :74    lambda */
    PF_LINDE = 2,             /* :131   mass, lambda */
    PF_HYBRID2AND4 = 3,       /* :201   mass, lambda */
    PF_PHI2OVER3 = 4,         /* :271   sigma */
    PF_MSQPHISQ_WITHV0 = 5,   /* :328   mass, V0 */
    PF_STEP_POTENTIAL = 6,    /* :391   mass, c, d, phi_s */
    PF_BUMP_POTENTIAL = 7,    /* :468   mass, c, d, phi_b */
    PF_RESONANCE = 8,         /* :545   mass, c, d */
    PF_BUMP_NOTHIRDDERIV = 9, /* :621   mass, c, d, phi_b */
    PF_HYBRIDQUADRATIC = 10,  /* :696   m1, m2 */
    PF_RIDGE_TWOFIELD = 11,   /* :744   g, m, V0 */
    PF_NFLATION = 12,         /* :791   mass  (nfields comes from pf_model.nfields) */
    PF_QUARTICTWOFIELD = 13,  /* :848   m1, m2, l1, l2 */
    PF_HYBRIDQUARTIC = 14,    /* :896   lambda, v, mu, phi_c */
    PF_INFLECTION = 15,       /* :963   lambda, g, r, m */
    PF_HILLTOPAXION = 16,     /* :1026  Lambda, f, m */
    PF_PRODUCTEXPONENTIAL = 17, /* :1088 lambda, V_0 */
    PF_NUM_POTENTIALS = 18
};

enum pf_error {
    PF_OK = 0,
    PF_E_ARG = -1,            /* null pointer / negative size / bad window */
    PF_E_POTENTIAL = -2,      /* unknown potential id, or nfields not valid for it */
    PF_E_NFIELDS = -3,        /* nfields outside 1..PF_MAX_NFIELDS (nflation: 1..PF_MAX_NFIELDS_RT) */
    PF_E_STEP = -4,           /* h is NaN or zero   (rk4.py:63 "Need to specify h.") */
    PF_E_START = -5           /* start index outside the step range (rk4.py:74) */
};

/* What cosmomodels.CosmologicalModel.__init__ resolves from (potential_func, pot_params,
 * nfields) (cosmomodels.py:135-154) plus CanonicalFirstOrder's ainit (:592-596). */
typedef struct pf_model {
    int32_t potential_id;
    int32_t nfields;
    double params[PF_MAX_PARAMS];     /* order given per potential above */
    double ainit;                     /* a(t) = ainit * exp(t); unused by pf_bg_run */
} pf_model;

/* output layouts of pf_fo_run / pf_fo_solve */
#define PF_LAYOUT_FULL 0        /* every row of yresult: [rows][nvar][nk] */
#define PF_LAYOUT_PERT 1        /* perturbation rows only: [rows][2 nf^2][nk]; the background rows
                                   are identical for every mode, so a host-bound result ships
                                   them once (pf_host_fill_background) instead of nk times */

/* flags written by pf_fo_run into *flags_dev (bitwise OR) */
#define PF_FLAG_BG_MISMATCH 1   /* a column's ystart background rows are not on column 0's trajectory */
#define PF_FLAG_TIMEOUT 2       /* a fused launch gave up waiting for its producer / lock-step partners
                                   (~20 s); the result is invalid -- rerun with PF_FO_NOFUSE=1 */

int pf_version(void);
const char *pf_last_error(void);

/* ---- potential table (seam 2: cosmomodels.py:136-139 resolves potential_func by name) ---- */
int pf_potential_id(const char *name);             /* -1 if unknown */
const char *pf_potential_name(int id);             /* NULL if unknown */
int pf_potential_nfields(int id);                  /* 1, 2, or 0 = any (nflation) */
int pf_potential_nparams(int id);
const char *pf_potential_param_name(int id, int i);
double pf_potential_param_default(int id, int i);

/* Evaluate (U, dU, d2U) on the device at n points; y_dev is [n][2nf+1] real background
 * vectors.  Replaces cmpotentials.<name>(y, params)[0:3].  d2U_dev may be NULL. */
int pf_potential_eval(const pf_model *m, int64_t n, const double *y_dev, double *U_dev,
                      double *dU_dev, double *d2U_dev, void *stream);

/* ---- K8: one right-hand-side evaluation (API completeness; not used by the solvers) ----
 * dydx = derivs(y, t) for y c128 [rows][ncol]: first_order = 1 -> CanonicalFirstOrder.derivs
 * (cosmomodels.py:625-675, rows = nvar, k_dev[ncol] required); first_order = 0 ->
 * CanonicalBackground.derivs (:551-571, rows = 2nf+1).  The potential is evaluated on column 0. */
int pf_derivs(const pf_model *m, int first_order, int64_t ncol, const double *y_dev, double t,
              const double *k_dev, double *dydx_dev, void *stream);

/* Sum of n <= 128 host doubles in the order the kernels use for sums over fields -- NumPy's
 * float64 add-reduce (eight running sums, a tree, the remainder in order): np.sum in
 * cmpotentials.py:837 and cosmomodels.py:569.  NaN for n outside 0..128. */
double pf_np_sum(const double *a_host, int n);

/* ---- sizes ---- */
int64_t pf_number_of_steps(double simtstart, double tend, double h);  /* rk4.py:73 */
int64_t pf_nvar(int nfields);                      /* 2 nf^2 + 2 nf + 1 */
int64_t pf_rec_doubles(int nfields);               /* doubles per step record (below) */
int64_t pf_rec_bg_offset(int nfields);             /* column of the background row inside a record */

/* ---- K1: background RK4 (replaces rkdriver_tsix + CanonicalBackground.derivs) ----
 * Integrates the real background y = (phi_i, phi'_i, H) from row n0 (state y0_host) to row
 * T-1 with step h.  bg_dev[T][2nf+1]: rows < n0 are NaN, row n0 = y0.  stagey_dev (may be
 * NULL) [T][4][2nf+1] receives the four RK4 stage arguments of every step n -> n+1. */
int pf_bg_run(const pf_model *m, int64_t T, double h, int64_t n0, const double *y0_host,
              double *bg_dev, double *stagey_dev, void *stream);

/* ---- K1b: per-step records consumed by pf_fo_run ----
 * rec_dev[T][pf_rec_doubles(nf)]:  for each RK4 stage s=0..3 of the step n -> n+1:
 * A = U/H^2,  R2 = 1/(a H)^2,  C[i][l] = (V_il + phi'_i V_l + V_i phi'_l + phi'_i phi'_l U)/H^2
 * (cosmomodels.py:663-674), a = ainit exp(t_stage), t_stage = simtstart + n h (+ h/2, + h/2,
 * + h) (rk4.py:27-55,116-118); then the bg row n (2nf+1 doubles, at column
 * pf_rec_bg_offset(nf)); padded to an even length.  Rows n < n0 are NaN.
 * nflation with nf >= 5 stores the factors of C instead of the dense matrix -- its Hessian is
 * m^2 I (cmpotentials.py:840), so C = d I + (p g^T + g p^T + U p p^T)/H^2 with p = phi',
 * g = dU: a stage slot is (A, R2, d = m^2/H^2, U, 1/H^2, 0, p[nf], g[nf]), and the mode kernel
 * needs 4 nf + 3 multiply-adds per column for the coupling instead of nf^2.  The records are an
 * opaque contract between pf_stage_table and pf_fo_run of the same library build. */
int pf_stage_table(const pf_model *m, int64_t T, double simtstart, double h, int64_t n0,
                   const double *bg_dev, const double *stagey_dev, double *rec_dev,
                   void *stream);

/* ---- K2/K3: first-order perturbation RK4 over a batch of k modes ----
 * Replaces rkdriver_tsix + rk4stepks + CanonicalFirstOrder.derivs for rows [t0, t1).
 *   k_dev[nk], tsix_dev[nk] (start row per mode), ystart_dev c128 [nvar][nk] (row tsix_k of
 *   column k is exactly ystart[:,k]; rows before it are NaN+NaNj; rows after it are RK4).
 *   state_dev c128 [2 nf^2][nk]: perturbation state at row t0 for modes already started
 *   (read when t0 > tsix_k), overwritten with the state at row t1 (for windowed runs).
 *   yout_dev c128: row (t - t0)/out_every of the window for every t in [t0,t1) with
 *   (t - t0) % out_every == 0, each row [nvar][nk]; out_every = 0 stores no trajectory.
 *   out_pitch: modes per output row, 0 = nk.  A k-sharded run passes the width of the
 *   assembled result and yout_dev = &assembled[0][0][k0] (k0 = first column of this slab): the
 *   kernel then writes its columns straight into the assembled [rows][nvar][out_pitch] array,
 *   which may live on another GPU of the box (peer memory, see pf_peer_*), i.e. the k-gather
 *   rides on the kernel's own stores over NVLink instead of a separate collective.
 *   flags_dev: int32[4] (may be NULL): [0] is OR-ed with PF_FLAG_*; [3] is scratch for the
 *   lock-step barrier of full-trajectory launches (zeroed by the call; with NULL the launch
 *   runs without lock-step).  pf_fo_solve's ws_dev + 1 has this layout. */
int pf_fo_run(const pf_model *m, int64_t nk, const double *k_dev, const int64_t *tsix_dev,
              const double *ystart_dev, const double *rec_dev, int64_t T, double h,
              int64_t t0, int64_t t1, double *state_dev, double *yout_dev, int64_t out_every,
              int32_t layout, int64_t out_pitch, int32_t *flags_dev, void *stream);

/* ---- one-call solve of the first window [0, t1): K1 + K1b + K2 ----
 * What rk4.rkdriver_tsix does for a CanonicalFirstOrder model in one call.  For single-field
 * runs with full-trajectory output (out_every == 1) this is ONE fused launch: a producer CTA
 * integrates the background and publishes the step records while the other CTAs integrate the
 * modes (they acquire the published count before each record chunk), so K1 hides behind K2.
 * Otherwise it issues pf_bg_run, pf_stage_table, pf_fo_run back to back.
 *   n0, y0_host  : column 0's start row and its background values (real parts of ystart[0:2nf+1, 0])
 *   scratch_dev  : pf_solve_scratch_doubles(nf, T) doubles; on return its first
 *                  T*pf_rec_doubles(nf) doubles are the records, valid as rec_dev for
 *                  pf_fo_run on later windows [t0 > 0, t1')
 *   ws_dev       : int32[8], zeroed by the call; ws_dev[1] holds the PF_FLAG_* bits afterwards */
int64_t pf_solve_scratch_doubles(int nfields, int64_t T);
int pf_fo_solve(const pf_model *m, int64_t nk, const double *k_dev, const int64_t *tsix_dev,
                const double *ystart_dev, int64_t n0, const double *y0_host, int64_t T,
                double simtstart, double h, int64_t t1, double *scratch_dev, double *state_dev,
                double *yout_dev, int64_t out_every, int32_t layout, int64_t out_pitch,
                int32_t *ws_dev, void *stream);

/* ---- K6 / K7: initial-condition glue on the device (FOCanonicalTwoStage.setfoics) ----
 * pf_crossing_rows: rows_dev[k] = first row t with cum_dev[t] > k, cum_dev[T] the running
 * maximum of cq*ainit*exp(t)*H(t) over the background rows before the end of inflation
 * (cosmomodels.py:1074-1131); rows_dev[k] = T when there is none.  nbad_dev (may be NULL) is
 * int32[2]: [0] += modes that never cross (ModelError at :1109), [1] += modes that cross at
 * row 0 (ModelError at :1335).
 * pf_bunch_davies: ystart_dev c128 [nvar][nk] from the background rows bg_dev[T][2nf+1]:
 * background rows of the start row, diagonal mode-matrix entries in the Bunch-Davies state
 * (cosmomodels.py:1443-1492); phase = 0 drops exp(-i k eta) (FONoPhase, :2083-2132);
 * suppress_ix >= 0 zeroes that field's mode (FOSuppressOneField, :2144-2197), -1 = none. */
int pf_crossing_rows(int64_t nk, const double *k_dev, int64_t T, const double *cum_dev,
                     int64_t *rows_dev, int32_t *nbad_dev, void *stream);
int pf_bunch_davies(int nfields, int64_t nk, const double *k_dev, const int64_t *rows_dev,
                    const double *bg_dev, double simtstart, double h, double ainit, double etainit,
                    int phase, int suppress_ix, double *ystart_dev, void *stream);

/* ---- K5: curvature power spectrum of trajectory rows (adiabatic.py:153-285) ----
 * y_dev c128 [nrows][nvar][nk] -> Pr_dev [nrows][nk];  if k_dev != NULL the result is the
 * scaled spectrum k^3/(2 pi^2) P_R (utilities.py:12-37). */
int pf_pr_spectrum(int nfields, int64_t nrows, int64_t nk, const double *y_dev,
                   const double *k_dev, double *Pr_dev, void *stream);
/* The same for a PF_LAYOUT_PERT window pert_dev c128 [nrows][2 nf^2][nk] holding rows
 * t0 .. t0+nrows-1: phi'_I is read from the background row inside the step records
 * (rec_dev[t][rec_pitch], background row at column bg_off).  Lets a host-bound solve reduce
 * P_R on each device window before it is recycled, instead of re-uploading the trajectory. */
int pf_pr_spectrum_pert(int nfields, int64_t nrows, int64_t nk, const double *pert_dev,
                        const double *rec_dev, int64_t rec_pitch, int64_t bg_off, int64_t t0,
                        const double *k_dev, double *Pr_dev, void *stream);

/* ---- host-bound results: ship the perturbation rows over PCIe, rebuild the rest on the host ----
 * pf_copy_rows_d2h: asynchronous 2-D device->host copy (height rows of width_bytes each, with
 * pitches), used to drop a PF_LAYOUT_PERT window into rows [2nf+1, nvar) of the host yresult.
 * pf_host_fill_background: host threads write rows [0, 2nf+1) of yresult[t0:t1] -- NaN+NaNj
 * before a mode's start row, ystart[:,k] at it (bit for bit, rk4.py:106-107), (bg[t][v], 0)
 * after it.  bg_host is [T][bg_pitch] doubles with the background row at column bg_off
 * (pass the step records: bg_pitch = pf_rec_doubles(nf), bg_off = 4(2+nf^2)).  pert_nan != 0
 * also writes the NaN prefix of the perturbation rows (t < tsix_k), so that the device->host
 * copies may skip columns that have not started.  Streaming (non-temporal) stores; nthreads <= 0
 * picks the hardware concurrency. */
int pf_copy_rows_d2h(void *dst_host, size_t dst_pitch, const void *src_dev, size_t src_pitch,
                     size_t width_bytes, size_t height, void *stream);
/* A whole device window window_dev c128 [nt][wrows][nk] into yresult_host c128 [T][nvar][nk] at
 * variable rows [row0, row0 + wrows) of time steps [t0, t0 + nt), columns [0, kcols) only: ONE
 * asynchronous 3-D copy (instead of one pitched copy per variable row), no pitch limit. */
int pf_copy_window_d2h(void *yresult_host, int64_t nvar, int64_t nk, int64_t row0, int64_t t0,
                       const void *window_dev, int64_t wrows, int64_t kcols, int64_t nt, void *stream);
int pf_host_fill_background(double *yout_host, int nfields, int64_t nk, int64_t t0, int64_t t1,
                            const double *bg_host, int64_t bg_pitch, int64_t bg_off,
                            const int64_t *tsix_host, const double *ystart_host, int pert_nan,
                            int nthreads);

/* ---- gather helper for the k-sharded multi-GPU layout ----
 * Interleave a rank's slab [nrows][nvar][nk_r] into the global layout [nrows][nvar][nk] at
 * column offset k0 (the on-GPU permute after an NCCL all-gather of slabs). */
int pf_scatter_slab(int64_t nrows_times_nvar, int64_t nk_r, int64_t nk, int64_t k0,
                    const double *slab_dev, double *full_dev, void *stream);

/* ---- K9: second-order perturbation RK4 over a batch of k modes ----
 * Replaces rkdriver_tsix + rk4stepks driving CanonicalSecondOrder.derivs (cosmomodels.py:708-766)
 * or CanonicalRampedSecondOrder.derivs (:893-971) as SOCanonicalThreeStage.runso sets them up
 * (:1680-1790), for rows [t0, t1) of the second-order time grid (step h = 2 x the first-order step).
 * The state per mode is the real 4-vector (Re d2, Re d2', Im d2, Im d2'):
 *     y0' = y1,  y1' = -( (3-eps) y1 + (k/(aH))^2 y0 + (V''/H^2 - 3 phi'^2) y0 + Re S )   (same for y2, y3, Im S)
 *   coef_dev[T][3][PF_SO_COEF]: for step n -> n+1 and its three stage times t = simtstart + n h,
 *     + h/2, + h (rk4.py:27-55, 116-118):  3 - eps,  1/(a H)^2,  V''/H^2 - 3 phi'^2,
 *     fotix = around((t - fo_simtstart)/fo_step) (row of the first-order time grid the reference
 *     reads at this stage, :724; an int64 in this 8-byte word),  t,  t/h (compared with the start row by
 *     the ramp, :935-937);  eps, H, phi', V'' taken at row fotix of the first-order / background
 *     result.  16-byte aligned.
 *   source_dev c128 [source_rows][source_pitch]: the source term on the first-order time grid,
 *     column k of this batch at column k; NULL = no source (homogeneous equation).  Every fotix
 *     in coef_dev must be < source_rows (the caller checks; the reference raises IndexError).
 *   tsix_dev[nk]: second-order start rows; fo_tsix_dev[nk] (may be NULL): first-order start rows --
 *     a stage whose fotix lies before a mode's first-order start reads NaN there in the reference
 *     and poisons the mode; ystart_dev [4][nk].
 *   ramp_host (may be NULL) = {a, b, c, d, e}: S is multiplied by (tanh(a (t - tstart_k - b)) + c)/d
 *     while |t - tstart_k - b| < e, and by 0 at the stage with t/h == the model's tstartindex_k
 *     (:925-944): ramp_tsix_dev[nk], NULL = tsix_dev (they differ only when the solver is called
 *     with other start rows than the model's, e.g. rk4stepks);
 *     tstart_dev[nk] = first-order start times (required with a ramp).
 *   state_dev [4][nk], yout_dev [rows][4][out_pitch or nk] float64, out_every: as pf_fo_run. */
#define PF_SO_COEF 6
int pf_so_run(int64_t nk, const double *k_dev, const int64_t *tsix_dev, const int64_t *fo_tsix_dev,
              const double *ystart_dev, const double *coef_dev, const double *source_dev,
              int64_t source_rows, int64_t source_pitch, const double *tstart_dev,
              const int64_t *ramp_tsix_dev, const double *ramp_host, int64_t T, double h,
              int64_t t0, int64_t t1,
              double *state_dev, double *yout_dev, int64_t out_every, int64_t out_pitch,
              void *stream);
/* One right-hand side dydx_dev[4][nk] = derivs(y_dev[4][nk], t): coef_dev[PF_SO_COEF] of that
 * time, source_row_dev c128 [nk] = source[fotix] (NULL: CanonicalHomogeneousSecondOrder.derivs,
 * :797-848); tsix_dev may be NULL without a ramp. */
int pf_so_derivs(int64_t nk, const double *k_dev, const int64_t *tsix_dev, const int64_t *fo_tsix_dev,
                 const double *y_dev, const double *coef_dev, const double *source_row_dev,
                 const double *tstart_dev, const double *ramp_host, double *dydx_dev, void *stream);

/* ---- measured fp64 peak (roofline denominator of the compute-bound configurations) ----
 * Runs a DFMA saturation kernel (8 independent chains per thread, 8 x 256 threads per SM) on the
 * default stream of the current device, synchronously, and returns the best of 3 timed launches
 * in TFLOP/s (FMA = 2 flops).  scratch_dev: at least 8 * 256 * (number of SMs) doubles. */
int pf_fp64_peak_probe(int iters, double *scratch_dev, int64_t scratch_doubles, double *tflops_out);

/* ---- peer memory: the destination of a fused k-gather (one process per GPU, one box) ----
 * The rank that assembles a result allocates it with pf_peer_alloc (a dedicated cudaMalloc, so
 * that the whole allocation can be exported), exports a 64-byte handle (cudaIpcGetMemHandle),
 * ships it to the other ranks by any means (torch.distributed in pyflation_b200/parallel.py);
 * they map it with pf_peer_open and pass the mapped pointer (+ their column offset) as yout_dev
 * with out_pitch = total width.  Stores travel over NVLink / NVSwitch.  The writers must
 * synchronise their stream and the ranks must meet at a barrier before the owner reads. */
#define PF_PEER_HANDLE_BYTES 64
int pf_peer_alloc(size_t bytes, void **ptr_dev);
int pf_peer_free(void *ptr_dev);
int pf_peer_export(const void *ptr_dev, unsigned char handle[PF_PEER_HANDLE_BYTES]);
int pf_peer_open(const unsigned char handle[PF_PEER_HANDLE_BYTES], void **ptr_dev);
int pf_peer_close(void *ptr_dev);

#ifdef __cplusplus
}
#endif
#endif /* PYFLATION_B200_H */
