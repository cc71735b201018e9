/*
 * upol_b200.h - C ABI of the B200-native U_ind path (libupol_b200.so).
 *
 * Drop-in boundary for the induced-dipole (Drude oscillator) energy, its gradient with respect
 * to the Drude displacements and the minimiser over them, i.e. the functions of the reference's
 * U_pol/polarization.py:19-188.  Plain pointers and sizes only; no torch / Python types.
 *
 * Conventions (reference semantics that survive):
 *   - units nm / e / kJ mol^-1; all floating-point data is IEEE fp64 at the boundary;
 *   - a "site" is one core particle of the reference's (nmol, natoms) layout flattened to
 *     i = m*natoms + a; sites of one molecule are contiguous (mol_id non-decreasing);
 *   - d_i = r_core_i - r_shell_i (util.py:40), so the shell sits at r_i - d_i;
 *   - k_i already contains ONE_4PI_EPS0 (util.py:267-268); q_shell_i = k_i = 0 on
 *     non-polarisable sites, whose gradient is exactly zero;
 *   - same-molecule pairs are excluded from the plain Coulomb sum (polarization.py:90-91) and
 *     interact only through the Thole-screened Drude-dipole pairs listed in (pair_a, pair_b,
 *     pair_u) (polarization.py:94-105);
 *   - zero-distance terms contribute nothing (polarization.py:33-42).
 *
 * Every function returns 0 on success, <0 on a CUDA/NCCL/argument error (see
 * upol_error_string) and >0 for numerical flags of the minimiser (UPOL_FLAG_*).
 * "_dev" entry points take DEVICE pointers, are asynchronous and ordered on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).  "_host" entry points take
 * HOST pointers, do the host<->device copies themselves and return after the result is on
 * the host.
 */
#ifndef UPOL_B200_H
#define UPOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct upol_system upol_system;   /* one geometry + parameters resident in HBM */

/* arithmetic modes */
#define UPOL_FP64  0   /* fp64 pair math (reference runs JAX x64, polarization.py:196)        */
#define UPOL_MIXED 1   /* fp32 pair math in difference form, fp64 accumulation (energy and field) */

/* minimiser methods */
#define UPOL_LBFGS 0   /* limited-memory BFGS, H0 = gamma * inverse molecule-block Hessian, derivative-driven line search */
#define UPOL_CG    1   /* the same with history 1: block-preconditioned memoryless BFGS (= PR-type conjugate gradient) */

/* error codes (<0) and numerical flags (>0) */
#define UPOL_OK               0
#define UPOL_ERR_CUDA        -1
#define UPOL_ERR_ARG         -2
#define UPOL_ERR_NCCL        -3
#define UPOL_ERR_ALLOC       -4
#define UPOL_FLAG_NOT_CONVERGED 1
#define UPOL_FLAG_NONFINITE     2
#define UPOL_FLAG_LINESEARCH    4
#define UPOL_FLAG_COMM          8   /* a peer did not deliver within the time-out (P2P exchange) */

/* layout of the energy vector written by the energy/gradient entry points (8 doubles) */
#define UPOL_E_UIND    0   /* (U_coul - U_coul_static) + U_self      polarization.py:149       */
#define UPOL_E_INTER   1   /* inter-molecular part of (U_coul - U_coul_static); the static
                              potential is subtracted row by row before the sum (conditioning) */
#define UPOL_E_INTRA   2   /* Thole-screened intra-molecular part     polarization.py:94-105   */
#define UPOL_E_SELF    3   /* 1/2 sum k d^2                            polarization.py:44-48    */
#define UPOL_E_GNORM2  4   /* ||dU/dd||_2^2                                                     */
#define UPOL_E_STATIC  5   /* informational: the d-independent remainder of the static
                              subtraction (:50-61,:82) summed on its own; ALREADY inside E_INTER */
#define UPOL_E_SIZE    8

const char *upol_error_string(int code);
int upol_version(void);

/* Replaces set_constants() (polarization.py:19-25): 138.93545764438198 kJ mol^-1 nm e^-2. */
double upol_one_4pi_eps0(void);

/* Number of CUDA devices visible to the library (0 if none / driver missing). */
int upol_device_count(void);

/*
 * Build a system on `device` from HOST arrays.
 * Replaces the array construction of util.py:45-279 (get_Rij_Dij, get_QiQj, get_pol_params)
 * with compact O(N) inputs instead of the dense (nmol,nmol,natoms,natoms[,3]) tensors.
 *   r_core[nsite*3], q_core[nsite], q_shell[nsite], k[nsite], mol_id[nsite] (non-decreasing)
 *   screened pairs: npair unordered site pairs (pair_a[p], pair_b[p]) of one molecule with
 *   Thole parameter pair_u[p] = thole/(alpha_a alpha_b)^(1/6)  (util.py:240-245)
 */
int upol_system_create(upol_system **out, int device, int64_t nsite,
                       const double *r_core, const double *q_core, const double *q_shell,
                       const double *k, const int32_t *mol_id,
                       int64_t npair, const int32_t *pair_a, const int32_t *pair_b,
                       const double *pair_u);
int upol_system_destroy(upol_system *sys);

/* New geometry, same topology (HOST pointer r_core[nsite*3]); invalidates the static part. */
int upol_system_set_positions(upol_system *sys, const double *r_core, void *stream);
/* Same, DEVICE pointer. */
int upol_system_set_positions_dev(upol_system *sys, const double *r_core_dev, void *stream);

int64_t upol_system_nsite(const upol_system *sys);
int64_t upol_system_ndrude(const upol_system *sys);
/* Point-charge interactions one gradient evaluation executes: N_D*[(N-n_a)+(N_D-n_D)] summed
 * exactly over rows (I_inter); `static_part` != 0 returns the count of the
 * d-independent pass instead. */
int64_t upol_system_interactions(const upol_system *sys, int static_part);

/*
 * Row sharding for multi-GPU runs: this process evaluates only the Drude rows of
 * molecules [mol_lo, mol_hi); all columns stay replicated.  Energies and gradients returned
 * are then the partial sums / row slices of that range.
 */
int upol_system_set_row_range(upol_system *sys, int64_t mol_lo, int64_t mol_hi);

/*
 * Uind + dUind/dDij.  Replaces Uind (polarization.py:111-151) and the implicit
 * jax.grad(Uind, argnums=1) taken inside jaxopt.BFGS (polarization.py:172-179).
 *   d[nsite*3]      Drude displacements, reference layout (values on non-polarisable sites are
 *                   ignored, as in the reference where they multiply a zero shell charge)
 *   energy[UPOL_E_SIZE]  may be NULL (gradient only: field pass, no potentials)
 *   grad[nsite*3]        may be NULL (energy only)
 * precision selects the pair arithmetic.  UPOL_FP64: fp64 pair terms, one pass that also forms the
 * d-independent static potential of the rows when it is not cached for the current geometry.  UPOL_MIXED:
 * ONE fp32 pass for the energy AND the gradient, no fp64 pair kernel and no static pass: U_ind is the
 * small difference U_coul - U_coul_static of two large sums, so the pair terms are evaluated in the
 * cancellation-free difference form f(A) - f(R) = (|R|^2-|A|^2)/(|A||R|(|A|+|R|)) on exact fixed-point
 * coordinate differences, with fp64 accumulation per 128-column tile (csrc/pair_diff.cu); energies and
 * gradients agree with UPOL_FP64 to 1e-5 relative (measured ~1e-8 / 2e-6 on liquid boxes).  In mixed mode
 * energy[UPOL_E_STATIC] is not formed (0).  Zero-distance rule of all kernels: an inter-molecular 1/r term
 * is dropped when r < 1e-5 nm (the reference: r == 0 exactly, polarization.py:33-42).
 */
int upol_uind_energy_grad_dev(upol_system *sys, const double *d, int precision,
                              double *energy, double *grad, void *stream);
int upol_uind_energy_grad_host(upol_system *sys, const double *d, int precision,
                               double *energy, double *grad);

/*
 * Sharded host path for row-sharded runs (after upol_system_attach_comm; every rank calls it with the same
 * arguments' meaning).  Each rank holds the full host arrays but moves only ITS site slice across PCIe:
 *   r_core[nsite*3]  new geometry, or NULL to keep the current one; only [site_lo, site_hi) is uploaded and the
 *                    peers' slices arrive over NVLink (peer stores after upol_system_p2p_attach, NCCL otherwise)
 *   d[nsite*3]       same rule
 *   energy[UPOL_E_SIZE]  the TOTAL over all ranks, identical bits on every rank (may be NULL)
 *   grad[nsite*3]    only the entries [3*site_lo, 3*site_hi) are written - the rows this rank evaluated; the
 *                    caller (or rank 0) assembles the slices.  May be NULL.
 *   site_lo/site_hi  out (may be NULL): the rank's site range (its molecule block)
 * With one rank it is upol_system_set_positions + upol_uind_energy_grad_host in one call.
 */
int upol_uind_energy_grad_host_sharded(upol_system *sys, const double *r_core, const double *d, int precision,
                                       double *energy, double *grad, int64_t *site_lo, int64_t *site_hi);

/*
 * dUind/dr_core at fixed displacements, dudr[nsite*3] (kJ mol^-1 nm^-1); the force on core i is
 * -dudr[i].  NEW relative to the reference (SURVEY 8f "next" #3): U_pol has no such function; it
 * is what an MD driver needs to move the cores on the U_ind surface.  fp64.  Sharded runs compute the
 * sites of the rank's molecule block and all-gather the result.
 */
int upol_uind_core_gradient_dev(upol_system *sys, const double *d, double *dudr, void *stream);

/*
 * d-independent Coulomb sums needed to rebuild the reference's separate Ucoul / Ucoul_static
 * values (polarization.py:50-61 and the qc*qc term of :82-86):
 *   out[0] = K sum_{i<j, inter} qc_i qc_j / R_ij
 *   out[1] = K sum_{i<j, inter} (qc+qs)_i (qc+qs)_j / R_ij      ( = Ucoul_static )
 */
int upol_coulomb_static_dev(upol_system *sys, double *out2, void *stream);

/* Uself = 1/2 sum_i k_i |d_i|^2 on plain DEVICE arrays d[n*3], k[n] (Uself, polarization.py:44-48); out[1].
 * The same quantity is the UPOL_E_SELF component of an energy evaluation. */
int upol_uself_dev(const double *d, const double *k, int64_t n, double *out, void *stream);

/* Thole screening function on a dense array, S = 1-(1+u r/2)exp(-u r), r = |Rij|
 * (make_Sij, polarization.py:27-31).  rij[n*3], u[n] (or u[1] when u_is_scalar), out[n]. */
int upol_make_sij_dev(const double *rij, const double *u, int u_is_scalar, int64_t n,
                      double *out, void *stream);

typedef struct upol_min_options {
    double tol;          /* stop when ||grad||_2 <= tol       (polarization.py:178: 1e-4)      */
    int32_t maxiter;     /* jaxopt default 500                                                   */
    int32_t method;      /* UPOL_LBFGS | UPOL_CG                                                 */
    int32_t precision;   /* UPOL_FP64 | UPOL_MIXED (mixed: fp32 field until ||g|| <= 1e-4 ||g(d=0)||,
                            final iterations and the returned energy in fp64)                    */
    int32_t history;     /* L-BFGS memory (default 8)                                            */
    int32_t check_every; /* host polls the device-side convergence flag every this many
                            iterations (default 4); no other host round-trip happens            */
    int32_t reserved;
} upol_min_options;

typedef struct upol_min_result {
    double energy[UPOL_E_SIZE];
    double gnorm;        /* final ||grad||_2                                                     */
    int32_t iterations;
    int32_t evaluations; /* gradient evaluations                                                 */
    int32_t flags;       /* UPOL_FLAG_*                                                          */
    int32_t reserved;
} upol_min_result;

void upol_min_options_default(upol_min_options *opts);

/*
 * Minimise Uind over the Drude displacements.  Replaces drudeOpt (polarization.py:154-188).
 *   d[nsite*3]  in: starting displacements (reference passes zeros, util.py:73-82);
 *               out: minimiser
 * The whole iteration (energy/gradient passes, dot products, line search, update, and for
 * sharded runs the collectives) runs on the device; the host only enqueues work and polls a
 * flag.  `_dev` returns after the final poll (it has to, to know when to stop enqueuing).
 */
int upol_minimize_dev(upol_system *sys, double *d, const upol_min_options *opts,
                      upol_min_result *result, void *stream);
int upol_minimize_host(upol_system *sys, double *d, const upol_min_options *opts,
                       upol_min_result *result);

/*
 * Multi-GPU plumbing (one process per GPU).  The library talks to NCCL itself so that the
 * minimiser loop stays on one stream: the caller obtains an id on rank 0
 * (upol_comm_unique_id), ships the UPOL_COMM_ID_BYTES to every rank by its own means
 * (torch.distributed in the Python host layer) and attaches the communicator to the system.
 * After that, energy/gradient calls all-reduce the energy vector and all-gather the gradient,
 * and the minimiser performs one all-gather of d and one all-reduce of a scalar pack per
 * iteration - through NCCL, or without any NCCL call once the peer mapping below is in place.
 */
#define UPOL_COMM_ID_BYTES 128
int upol_comm_unique_id(unsigned char *id_bytes);
int upol_system_attach_comm(upol_system *sys, const unsigned char *id_bytes, int rank, int nranks);

/*
 * Peer-memory exchange (optional, after upol_system_attach_comm).  Each rank exports CUDA-IPC handles of its
 * exchange buffers (displacements, gradient rows, positions, mailboxes, arrival flags: UPOL_P2P_EXPORT_BYTES),
 * the caller all-gathers the blobs (rank order) and hands them back.  From then on neither a minimiser iteration
 * nor an energy/gradient call uses NCCL: the minimiser's update kernel stores its slice of d straight into every
 * peer's vector over NVLink and raises the peers' arrival flags, the partial inner products travel the same way;
 * the finalize kernel of an evaluation stores its gradient rows and the rank's energy partials into every peer
 * (fused compute + all-gather); sums are taken in rank order, so every rank holds the same bits.
 */
#define UPOL_P2P_EXPORT_BYTES 512
int upol_system_p2p_export(upol_system *sys, unsigned char *blob);
int upol_system_p2p_attach(upol_system *sys, const unsigned char *all_blobs);
/* Back to NCCL for the minimiser loop (every rank must call it, e.g. when a peer could not map). */
int upol_system_p2p_disable(upol_system *sys);

/*
 * Batched independent small systems (SAPT-style dimer scan, replaces the subprocess-per-dimer
 * loop of wrapper.py:265-316 around drudeOpt): nbatch systems sharing one topology
 * (nsite sites each), one thread block per system, dense BFGS in shared memory.
 *   r_core[nbatch*nsite*3]; d[nbatch*nsite*3] in/out; energy[nbatch*UPOL_E_SIZE];
 *   energy[...][6] holds the UPOL_FLAG_* bits of the system; iterations[nbatch] may be NULL.
 *   DEVICE pointers except the topology arrays (HOST).  At most 32 Drude sites per system (one
 *   lane per Drude row).  Returns after the batch has finished.
 */
int upol_batched_minimize_dev(int device, int64_t nbatch, int64_t nsite,
                              const double *r_core_dev, double *d_dev,
                              const double *q_core, const double *q_shell, const double *k,
                              const int32_t *mol_id,
                              int64_t npair, const int32_t *pair_a, const int32_t *pair_b,
                              const double *pair_u,
                              double tol, int32_t maxiter,
                              double *energy_dev, int32_t *iterations_dev, void *stream);

/*
 * Kernel timing hook for roofline reporting (bench.py): when enabled, every energy/gradient
 * evaluation brackets its main pair kernel with CUDA events on the launching stream.
 * upol_system_kernel_time synchronises those events and returns the summed duration in ms and
 * the number of launches since the last call (and resets the counters).  Up to 256 launches are
 * kept between two calls.
 */
int upol_system_set_kernel_timing(upol_system *sys, int enable);
/* Kernels of this library launched on behalf of `sys` since it was created (energy/gradient evaluations,
 * geometry uploads, minimiser slots; counted where they are enqueued).  bench.py reports the difference
 * over its timed region as gpu_launches. */
int64_t upol_system_launch_count(const upol_system *sys);
int upol_system_kernel_time(upol_system *sys, double *ms_total, int32_t *launches);

/* Periodic boundaries - an EXTENSION, not a replacement of reference code: shehan807/U_pol is gas
 * phase only (benchmarks/OpenMM/water/calc_energy.py:80); SURVEY 8f lists this as "next".  With a box
 * set, U_ind and its gradient (and therefore upol_minimize_*) are those of the orthorhombic periodic
 * system, by direct Ewald summation with tin-foil boundary conditions and the reference's
 * intra-molecular exclusions (definition and checker: oracle/ewald_dense.py).  box = 3 edge lengths in
 * nm (NULL: back to open boundaries); kappa <= 0 picks 13 / min(box) (the real-space sum visits the
 * nearest image only, so kappa * min(box) >= 12 is required); nmax = 0 keeps every reciprocal vector
 * with exp(-k^2/4 kappa^2) >= exp(-41.5), nmax > 0 the cube |n| <= nmax.  Molecules must be whole
 * (not wrapped atom by atom) and smaller than half the box (checked on the device, UPOL_ERR_ARG
 * otherwise).  fp64; row shards work as in the open case; upol_uind_core_gradient_dev gives the
 * periodic dU_ind/dr_core; the separate static values and the mixed field are open-boundary only
 * (UPOL_ERR_ARG). */
int upol_system_set_periodic(upol_system *sys, const double *box, double kappa, int32_t nmax);
int upol_system_periodic_info(upol_system *sys, double *box3, double *kappa, int32_t *n_kvectors);

/* Debug aid (compute-sanitizer stand-in): with UPOL_GUARD=1 in the environment every device buffer of the library
 * carries guard zones; this returns the number of damaged ones over all live buffers (0 = clean, -1 = guard mode off). */
int upol_debug_check_guards(void);
int upol_debug_guard_selftest(void);   /* 1 = a deliberate one-byte overrun was detected */

/* FMA-pipe peak microbenchmarks used as roofline denominators: returns the
 * measured FLOP/s of a register-resident FMA chain on the current device: use_fp64 = 1 fp64 (DFMA),
 * 0 fp32 (FFMA), 2 packed fp32 pairs (FFMA2, fma.rn.f32x2). */
int upol_measure_fma_peak(int device, int use_fp64, double *flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* UPOL_B200_H */
