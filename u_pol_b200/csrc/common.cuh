// Internal declarations shared by the translation units of libupol_b200.so.
// Not part of the ABI (see include/upol_b200.h for that).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/upol_b200.h"

namespace upol {

// Coulomb constant, U_pol/polarization.py:19-25 (same literals, same operation order).
constexpr double kPi = 3.14159265358979323846;
constexpr double kECharge = 1.602176634e-19;
constexpr double kAvogadro = 6.02214076e23;
constexpr double kEps0 = 1e-6 * 8.8541878128e-12 / (kECharge * kECharge * kAvogadro);
constexpr double kCoulomb = 1 / (4 * kPi * kEps0);

// Pair-kernel tiling.  A column tile is kTileCols point charges staged in shared memory;
// a thread block owns kBlock * ROWS rows (rows live in registers).
constexpr int kBlock = 128;
constexpr int kTileCols = 128;
constexpr int kRowsPerThread = 2;
constexpr int kRowTile = kBlock * kRowsPerThread;
constexpr int kPartComps = 6;   // phi_core, phi_shell, Fx, Fy, Fz (+1: the core-force pass stores two fields)

#define UPOL_CUDA(expr)                                   \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) {                          \
            upol::set_last_cuda_error(_e, __FILE__, __LINE__); \
            return UPOL_ERR_CUDA;                         \
        }                                                 \
    } while (0)

void set_last_cuda_error(cudaError_t e, const char *file, int line);

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// device allocations of the library (guard.cu): plain cudaMalloc / cudaFree unless UPOL_GUARD=1
cudaError_t dmalloc(void **p, size_t bytes);
cudaError_t dfree(void *p);
bool guard_enabled();

// Arguments of one pair-kernel launch (struct passed by value).
struct PairArgs {
    // rows
    int n_rows;                 // rows to process
    int row0;                   // first row (global row index; shard offset)
    const double *rx, *ry, *rz; // row positions, global row index
    const int *exA_lo, *exA_len;  // excluded column interval inside section A (per row)
    const int *exB_lo, *exB_len;  // excluded column interval inside section B (per row, already offset)
    // columns: concatenated, each section padded to a multiple of kTileCols
    const void *cols;           // double4* (fp64): x, y, z, q;  int4* (mixed, sections P|N): fixed-point x, y, z + float q bits
    const void *cols_b;         // optional: section B lives in a separate array (else it follows A)
    int n_tiles_a;              // tiles in section A (cores)
    int n_tiles;                // total tiles (A then B)
    // output
    int n_split;                // column splits (gridDim.y)
    double *part;               // [n_split][kPartComps][part_stride], indexed by LOCAL row
    int64_t part_stride;
    // mixed mode only: box centre and 1/grid of the fixed-point coordinates
    double cx, cy, cz, inv_grid;
    // mixed mode (pair_diff.cu): row displacements [nd*3] (compact), per-site -2 d_j | q_shell (float4) and
    // |d_j|^2 (float) of section P
    const double *dcomp;
    const void *cols_d;
    const float *cols_d2;
    // fused static pass of the fp64 kernel: qc_j + qs_j/2 per core column [na_pad]
    const double *col_qeff;
    // device-side gate (minimiser): the kernel returns at once when (*gate & gate_skip_mask) != 0
    const int *gate;
    int gate_skip_mask;
};

// Peer-memory exchange of one evaluation (multi-GPU, after upol_system_p2p_attach): finalize_rows_kernel stores
// the gradient rows of its shard and the rank's energy partials straight into every peer (NVLink stores), fences,
// and its last block raises the rank's arrival epoch on every peer; eval_xchg_wait_kernel waits for all epochs and
// adds the energy partials in rank order (every rank gets the same bits).  No NCCL call, no separate all-gather.
struct EvalXchg {
    int on, rank, nranks, epoch;
    int push_grad;          // 0: only the energy partials travel (sharded host path: every rank keeps its own rows)
    double **peer_g;        // every rank's xg (self included)
    double **peer_e;        // every rank's xe
    int **peer_flags;       // every rank's xflags2
    int64_t g_off;          // (epoch & 1) * 3 * nd
    int e_off;              // (epoch & 1) * nranks * UPOL_E_SIZE
};

// gate bits written by the minimiser's control kernel
constexpr int kGateDone = 1, kGatePhase64 = 2, kGatePhaseMixed = 4;
constexpr int kMaxPack = 17 * 8;   // (max history + 1) x values per slot of the minimiser's scalar pack

int launch_pair_fp64(const PairArgs &a, bool with_pot, bool with_grad, bool fuse_static, cudaStream_t st);
int launch_pair_diff(const PairArgs &a, bool with_pot, cudaStream_t st);
int diff_rows_per_block(bool with_pot);
int launch_pair_fp64_two_fields(const PairArgs &a, cudaStream_t st);
int fp64_rows_per_block();

}  // namespace upol

// ------------------------------------------------------------------------------------------
// The system object behind the opaque handle.
struct upol_system {
    int device = 0;
    int64_t nsite = 0, nmol = 0, nd = 0, npair_ordered = 0;
    int64_t nd_pad = 0;           // rows padded to kRowTile
    int64_t na_pad = 0, nb_pad = 0;  // column sections padded to kTileCols
    int max_split = 0;

    // host-side topology
    std::vector<int32_t> h_mol, h_drude_site, h_site_row;
    std::vector<int32_t> h_mol_site0, h_mol_nsite, h_mol_row0, h_mol_nrow;
    std::vector<double> h_qc, h_qs, h_k;

    // device: per-site parameters and geometry
    double *r = nullptr;          // [nsite*3] core positions
    double *qc = nullptr, *qs = nullptr, *k = nullptr;   // [nsite]
    // device: per-Drude-row (compact)
    int *row_site = nullptr;      // [nd] site of the row
    int *site_row = nullptr;      // [nsite] Drude row of a site, -1 if not polarisable
    double *row_qs = nullptr, *row_k = nullptr;           // [nd]
    int *exA_lo = nullptr, *exA_len = nullptr, *exB_lo = nullptr, *exB_len = nullptr;   // [nd_pad]
    double *rowx = nullptr, *rowy = nullptr, *rowz = nullptr;     // [nd_pad] shell positions
    double *srowx = nullptr, *srowy = nullptr, *srowz = nullptr;  // [nd_pad] Drude core positions
    // device: columns
    double4 *cols = nullptr;      // [na_pad + nb_pad]  cores (x,y,z,qc) | shells (x,y,z,qs)
    // mixed mode, difference form (pair_diff.cu): section P = polarisable sites in Drude-row order (nb_pad
    // slots), section N = the other sites (nn_pad slots)
    int4 *cols_x = nullptr;       // [nb_pad + nn_pad] fixed-point core position, float bits of q_core
    float4 *cols_d = nullptr;     // [nb_pad] -2 d_j in grid units, q_shell
    float *cols_d2 = nullptr;     // [nb_pad] |d_j|^2 in grid units^2
    int *site_nslot = nullptr;    // [nsite] slot of a non-polarisable site in section N, -1 otherwise
    int *exP_lo = nullptr, *exP_len = nullptr, *exN_lo = nullptr, *exN_len = nullptr;   // [nd_pad]
    int64_t nn_pad = 0;
    double4 *cols_static = nullptr;  // [na_pad] cores, q = qc + qs/2 (static columns of the core-force pass and of the Ewald static pass)
    double *col_qeff = nullptr;      // [na_pad] qc + qs/2 alone (fused static pass of the fp64 pair kernel)
    double centre[3] = {0, 0, 0};
    double inv_grid = 1.0;        // fixed-point grid: 1/grid, grid = 2^-k nm so the box fits +-2^30
    // Thole screened pairs as CSR over Drude rows (ordered pairs)
    int *th_ptr = nullptr;        // [nd+1]
    int *th_row = nullptr;        // partner row
    double *th_u = nullptr;
    // work buffers
    double *part = nullptr;       // [max_split][kPartComps][nd_pad]
    size_t part_elems = 0;        // doubles allocated for part
    double *dc = nullptr;         // [nd*3] compact displacements
    double *gc = nullptr;         // [nd*3] compact gradient
    double *eblock = nullptr;     // per-block energy partials
    int n_eblock = 0;
    int *ticket = nullptr;        // [4] device: block-completion counter of the finalize kernel (self-resetting)
    int64_t launches = 0;         // kernels of this library launched for this system (bench.py's gpu_launches)
    double *energy = nullptr;     // [UPOL_E_SIZE] device result of the last evaluation
    double *e_static = nullptr;   // [2] device: d-independent part summed over rows (informational)
    double *phi_static = nullptr; // [nd_pad] device: per-row static potential, subtracted row by row
    // core-force pass (allocated on first use): Drude-core columns (r_i, qs_i), site-row tables, scratch
    double4 *cols_dcore = nullptr;   // [nb_pad]
    double *site_x = nullptr, *site_y = nullptr, *site_z = nullptr;   // [ns_pad]
    int *exS_lo = nullptr, *exS_len = nullptr, *exS2_lo = nullptr;    // [ns_pad] excluded Drude rows of the site's molecule
    double *cf_tmp = nullptr;        // [6*nd + 6*nsite] folded fields
    int64_t ns_pad = 0;
    bool static_valid = false;
    int64_t geom_gen = 0;         // bumped by every geometry upload (captured graphs hold the box centre / grid by value)

    // row shard (Drude rows of molecules [mol_lo, mol_hi))
    int64_t mol_lo = 0, mol_hi = 0, row_lo = 0, row_hi = 0;

    // pinned host staging + device scratch for the *_host entry points (kept between calls)
    double *pin = nullptr;
    size_t pin_bytes = 0;
    double *hscratch = nullptr;   // device: d_full | g_full | energy
    size_t hscratch_elems = 0;

    // communicator (NCCL), opaque here
    void *comm = nullptr;
    int rank = 0, nranks = 1;

    // peer-memory exchange of the minimiser (NVLink P2P through CUDA IPC, see comm.cu): every rank
    // maps every other rank's displacement vector, scalar-pack mailbox and arrival flags
    bool p2p = false;
    double *xchg_pack = nullptr;      // [nranks][kMaxPack] mailbox: rank r's partial inner products
    int *xchg_flags = nullptr;        // [2*nranks + 1] arrival epochs: packs, d slices; +1 block counter
    double **d_peer_x = nullptr;      // device table [nranks]: peers' dc (self included)
    double **d_peer_pack = nullptr;   // device table [nranks]: peers' xchg_pack
    int **d_peer_flags = nullptr;     // device table [nranks]: peers' xchg_flags
    std::vector<void *> ipc_opened;   // mapped peer pointers (closed at destroy)
    int p2p_epoch = 0;                // last epoch handed out (monotonic across minimisations)
    // the same for the energy/gradient entry points and the sharded host path: finalize_rows_kernel stores its
    // gradient rows and the rank's energy partials into every peer, push_slices_kernel the rank's slices of r and d
    double *xg = nullptr;             // [2][3*nd] gradient rows of all ranks, double-buffered by epoch parity
    double *xe = nullptr;             // [2][nranks][UPOL_E_SIZE] energy partials of every rank, same parity rule
    double *xd = nullptr;             // [3*nsite] displacements in the (nsite,3) layout (sharded host path)
    int *xflags2 = nullptr;           // [3][nranks] arrival epochs (evaluation results | r slices | d slices) + [4] block counters
    double **d_peer_g = nullptr, **d_peer_e = nullptr, **d_peer_r = nullptr, **d_peer_d = nullptr;   // device tables [nranks]
    int **d_peer_flags2 = nullptr;
    int epoch_eval = 0, epoch_push = 0;

    // minimiser workspace (allocated on first use)
    void *min_ws = nullptr;

    // periodic boundaries: Ewald state (ewald.cu), null = open boundaries
    void *ewald = nullptr;

    // optional timing of the main pair kernel (bench.py roofline)
    bool timing = false;
    std::vector<cudaEvent_t> tev;   // start/stop pairs
    int tev_used = 0;
};

namespace upol {

// evaluation pipeline pieces (system.cu / finalize_kernels.cu)
// what one evaluation computes
struct EvalOpts {
    int precision = UPOL_FP64;     // arithmetic of the FIELD; the potential is always fp64
    bool want_energy = true;
    bool want_grad = true;
    bool reduce_ranks = true;      // all-reduce the energy vector over ranks (sharded runs)
    const int *gate = nullptr;     // minimiser gate word (see kGate*)
    bool phase_switch = false;     // minimiser: launch both field kernels, gate picks one
    const double *d_full = nullptr; // displacements in the (nsite,3) layout: gathered into s->dc by the prepare kernel
    double *g_full = nullptr;       // gradient also in the (nsite,3) layout, written by the finalize kernel (single rank, all rows)
    bool shells_ready = false;      // shell rows / columns already match d (minimiser: written by its update kernel)
    EvalXchg gx = {0, 0, 1, 0, 0, nullptr, nullptr, nullptr, 0, 0};   // peer-memory exchange instead of NCCL (reduce_ranks)
};
EvalXchg make_eval_xchg(upol_system *s, bool push_grad);   // hands out the next evaluation epoch
int sys_eval_compact(upol_system *s, const double *d_compact, const EvalOpts &o,
                     double *energy_dev, double *grad_compact, cudaStream_t st);
int sys_static_part(upol_system *s, cudaStream_t st);
int sys_prepare_shells(upol_system *s, const double *d_compact, cudaStream_t st);
int sys_gather_d(upol_system *s, const double *d_full, double *d_compact, cudaStream_t st);
int sys_scatter_g(upol_system *s, const double *g_compact, double *g_full, cudaStream_t st);
int sys_scatter_g_range(upol_system *s, const double *g_compact, double *g_full, int64_t site_lo, int64_t site_hi, cudaStream_t st);
int sys_scatter_d(upol_system *s, const double *d_compact, double *d_full, cudaStream_t st);
int sys_ensure_pinned(upol_system *s, size_t bytes);
int sys_ensure_hscratch(upol_system *s, size_t elems);
int choose_split(const upol_system *s, int64_t n_rows, int n_tiles, int rows_per_block);
int sys_upload_positions(upol_system *s, const double *r_dev_or_null, cudaStream_t st);
void set_box(upol_system *s, const double *lo, const double *hi);
int sys_box_from_device(upol_system *s, cudaStream_t st);
int sys_uself(const double *d, const double *k, int64_t n, double *out, cudaStream_t st);
int sys_make_sij(const double *rij, const double *u, int u_is_scalar, int64_t n, double *out, cudaStream_t st);
int sys_coulomb_static(upol_system *s, double *out2, cudaStream_t st);
int sys_core_forces(upol_system *s, const double *d_compact, double *dudr_full, cudaStream_t st);
const char *last_error_message();

// collectives (comm.cu)
int comm_allreduce_sum(upol_system *s, double *buf, int64_t n, cudaStream_t st);
int comm_allgather_rows(upol_system *s, double *vec3, cudaStream_t st);  // [nd*3] in place by shard
int comm_allgather_sites(upol_system *s, double *vec3, cudaStream_t st); // [nsite*3] in place by shard
int comm_site_range(const upol_system *s, int rank, int64_t *site_lo, int64_t *site_hi);
int p2p_allgather_r_d(upol_system *s, bool with_r, cudaStream_t st);     // site slices of s->r and s->xd by peer stores
void comm_destroy(upol_system *s);
void p2p_destroy(upol_system *s);

// periodic boundaries (ewald.cu)
int ewald_setup(upol_system *s, const double *box, double kappa, int nmax);
int ewald_info(const upol_system *s, double *box, double *kappa, int *nk);
void ewald_free(upol_system *s);
void ewald_invalidate(upol_system *s);
int ewald_static_part(upol_system *s, PairArgs a, int *n_slots, cudaStream_t st);
int ewald_dynamic_part(upol_system *s, PairArgs a, int *n_slots, cudaStream_t st);
int ewald_field_part(upol_system *s, PairArgs a, int which, int *n_slots, cudaStream_t st);

// minimiser (minimize.cu)
int minimize_compact(upol_system *s, double *d_compact, const upol_min_options *o,
                     upol_min_result *res, cudaStream_t st);
void minimize_free(upol_system *s);

}  // namespace upol
