/*
 * bfe2.h -- C ABI of libbfe2.so, the B200 (sm_100a) implementation of BeamFE2's numeric hot path.
 *
 * The reference (jkbgbr/BeamFE2) is pure Python and has no FFI layer; its operator boundary is the
 * Solver contract  Solver(structure).solve() -> bool  (BeamFE2/Solver.py:16-21) which reads
 * Structure.K_with_BC / M_with_masses / condense / load_vector (BeamFE2/Structure.py:183-236) and
 * writes structure.results[...] (BeamFE2/Results.py:87-173).  The entry points below are what a
 * ctypes binding placed inside those solve() methods would call; each one
 * cites the reference code it replaces.  All file:line citations are relative to the reference root.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a CUDA DEVICE pointer to contiguous
 *     float64 (or int32 / uint8 where stated) owned by the caller; outputs are fully overwritten.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync inside.
 *   - return value: 0 = ok, negative = bad argument / CUDA failure (text via bfe2_last_error()).
 *     Numerical failure is reported PER STRUCTURE in info[B]:
 *        0  ok
 *        k>0  non-positive Cholesky pivot of the condensed K at column k (1-based)
 *             == the reference's "stiffness matrix is singular" (Structure.py:325-354)
 *        -1  condensed mass matrix not positive definite / all-zero (Solver.py:34-35 returns False)
 *        -2  Jacobi did not converge in BFE2_MAX_SWEEPS sweeps / the reduced matrix broke down
 *        -3  (BFE2_INFO_RETRY) internal: the two-ended modal engine declined the structure; the Jacobi engine
 *            repairs it inside the same call, so a caller never sees this value
 *     k > 0 invalidates every output of that structure (NaN).  -1 / -2 invalidate only the MODAL outputs
 *     (lam, phi = NaN): K factored fine, so u / endforces / reactions of that structure are valid.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with -100.
 *
 * Layout of the per-variant tensors (B = number of structure variants sharing one topology plan,
 * C = number of load cases, nN nodes, nE beams, sumdof = 3 nN, n = number of free DOFs, hbw = half
 * bandwidth of the condensed matrices in free-DOF numbering):
 *     coords      [B, nN, 2]      node coordinates (Node.py:31-37)
 *     props       [B, nE, 4]      E, A, I, rho per beam (HermitianBeam_2D.py:42-67)
 *     nodal_mass  [B, nN, 2]      mx, my per node or NULL (Structure.py:53-74, 289-313)
 *     f_nodal     [B, C, sumdof]  directly applied nodal loads (Structure.py:76-97, 258-287)
 *     r_local     [B, C, nE, 6]   per beam: sum of the LOCAL fixed-end vectors of its internal loads
 *                                 (Loads.py reactions_asvector, :202-208, 264-270, 333-342, 409-416,
 *                                 474-484) or NULL
 *     Kb, Mb      [B, hbw+1, n]   lower band storage, Xb[d][j] = X[j+d, j] of the condensed matrix
 *     u           [B, C, sumdof]  displacements, exact zeros at supported DOFs
 *     endforces   [B, C, nE, 6]   local end forces Ke u_loc - r_local (HermitianBeam_2D.py:336-351)
 *     reactions   [B, C, sumdof]  Results.py:92-127
 *     lam         [B, n]          all eigenvalues, ascending (Solver.py:39)
 *     phi         [B, n_modes, sumdof]  first n_modes M-normalised shapes, zeros at supports
 *                                 (Solver.py:54-65); sign: the largest-magnitude entry is positive
 */
#ifndef BFE2_H
#define BFE2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFE2_VERSION 100          /* 0.1.0 */
#define BFE2_MAX_SWEEPS 40
#define BFE2_MAX_JACOBI_N 116     /* largest n_free of the all-modes path (registers up to 64, shared memory above) */

typedef struct bfe2_plan bfe2_plan;

/* plan integer arrays retrievable with bfe2_plan_get_array (for bit-exact DOF-map checks) */
enum bfe2_plan_array {
    BFE2_ARR_CONN = 0,          /* [nE*2]       0-based node indices (helpers.py:28-31 minus the *3) */
    BFE2_ARR_FREE_OF_POS = 1,   /* [sumdof]     free index of a global position, -1 if supported   */
    BFE2_ARR_POS_OF_FREE = 2,   /* [n]          global position of each free DOF (what np.delete keeps,
                                                Structure.py:183-191)                               */
    BFE2_ARR_FIXED = 3,         /* [nF]         == Structure.positions_to_eliminate (:164-181)       */
    BFE2_ARR_SLOT_PTR = 4,      /* [(hbw+1)*n+1] CSR pointers of the band-slot scatter plan          */
    BFE2_ARR_CONTRIB = 5,       /* [n_contrib]  e*36 + r*6 + c, beam-list order inside every slot    */
    BFE2_ARR_NODE_PTR = 6,      /* [nN+1]       CSR pointers node -> incident (beam*2 + end)         */
    BFE2_ARR_NODE_INC = 7       /* [2*nE]                                                            */
};

/* ---- library ------------------------------------------------------------------------------ */
int         bfe2_version(void);
const char* bfe2_last_error(void);                 /* thread-local, never NULL */
int         bfe2_device_count(void);               /* 0 when no usable CUDA device */

/* ---- topology plan (replaces the DOF bookkeeping of Structure.py:154-191, 379-393 and the block
 *      indexing of helpers.py:7-40; pure host code, no CUDA call until the first launch) -------- */
int  bfe2_plan_create(bfe2_plan** out, int32_t n_nodes, int32_t n_elem,
                      const int32_t* conn /*[nE,2] 0-based*/,
                      int32_t n_fixed, const int32_t* fixed_pos /* any order, no duplicates */,
                      const uint8_t* lumped /*[nE] or NULL*/);
/* flags: BFE2_PLAN_REORDER = internal bandwidth-reducing (reverse Cuthill-McKee) numbering of the free DOFs for
 * frames that are not numbered as chains (e.g. Examples/Chopra_105.py:35-45); inputs and outputs keep the
 * reference's DOF positions, only the band storage [B, hbw+1, n] is in the internal order. */
#define BFE2_PLAN_REORDER 1
int  bfe2_plan_create_ex(bfe2_plan** out, int32_t n_nodes, int32_t n_elem, const int32_t* conn,
                         int32_t n_fixed, const int32_t* fixed_pos, const uint8_t* lumped, int32_t flags);
void bfe2_plan_destroy(bfe2_plan* plan);
int  bfe2_plan_dims(const bfe2_plan* plan, int32_t* n_nodes, int32_t* n_elem, int32_t* sumdof,
                    int32_t* n_free, int32_t* hbw, int32_t* n_contrib);
int  bfe2_plan_get_array(const bfe2_plan* plan, int which, int32_t* out, int64_t capacity);
/* bytes of dynamic shared memory per structure the fused / modal kernels need for this plan */
size_t bfe2_plan_smem_bytes(const bfe2_plan* plan, int32_t n_loadcases);
/* caller-owned scratch needed by the entry points below for a batch of B: 0.  Structures whose band does not fit
 * the shared memory of an SM take the static path through a stream-ordered workspace the library allocates and
 * frees on the caller's stream (cudaMallocAsync / cudaFreeAsync, <= 256 MB); nothing persistent is allocated. */
size_t bfe2_workspace_bytes(const bfe2_plan* plan, int64_t B);

/* ---- K1: element kernel (HermitianBeam_2D.py:98-120, 225-249, 284-297; helpers.py:43-69;
 *      matrix_in_global :173-175).  SoA inputs of nE_total elements, outputs [nE_total, 6, 6]. --- */
int bfe2_element_km(int64_t n_elem_total,
                    const double* xi, const double* yi, const double* xj, const double* yj,
                    const double* E, const double* A, const double* I, const double* rho,
                    const uint8_t* lumped /*or NULL*/, double* Kg, double* Mg, void* stream);

/* ---- K2: banded assembly with support condensation and nodal masses
 *      (helpers.py:7-40; Structure.py:183-209) ------------------------------------------------- */
int bfe2_assemble_banded(bfe2_plan* plan, int64_t B, const double* coords, const double* props,
                         const double* nodal_mass /*or NULL*/, double* Kb, double* Mb /*or NULL*/,
                         void* stream);

/* ---- K3: banded Cholesky + forward/back substitution, end forces, reactions
 *      (Solver.py:76-87; Results.py:48-127; HermitianBeam_2D.py:336-351).  No size limit: narrow bands run
 *      the 16-lanes-per-structure kernel, wide ones a warp per structure, bands beyond shared memory the
 *      HBM-workspace kernel. ---------------------------------------------------------------------- */
int bfe2_static_solve(bfe2_plan* plan, int64_t B, int32_t C, const double* Kb,
                      const double* coords, const double* props,
                      const double* f_nodal, const double* r_local /*or NULL*/,
                      double* u, double* endforces /*or NULL*/, double* reactions /*or NULL*/,
                      int32_t* info, void* stream);

/* ---- K4: generalized symmetric eigenproblem K phi = lam M phi, n_free <= BFE2_MAX_JACOBI_N: standard form
 *      C = Lm^-1 K Lm^-T, pivoted Cholesky C = Z Z^T, batched one-sided Jacobi on Z (Solver.py:29-69); ALL
 *      eigenvalues, the first n_modes shapes.  sweeps[B] (or NULL) receives the sweep count.  Larger
 *      structures: -3 (use the first-k route, beamfe2_b200/frame_modes.py). ------------------------- */
int bfe2_modal_solve(bfe2_plan* plan, int64_t B, int32_t n_modes, const double* Kb, const double* Mb,
                     double* lam, double* phi /*or NULL when n_modes==0*/, int32_t* info,
                     int32_t* sweeps /*or NULL*/, void* stream);

/* ---- fused K1->K4: nothing but inputs and outputs touches HBM.  C may be 0 (modal only, the
 *      static outputs may then be NULL); do_modal = 0 skips the eigenproblem. -------------------- */
int bfe2_solve_fused(bfe2_plan* plan, int64_t B, int32_t C, int32_t do_modal, int32_t n_modes,
                     const double* coords, const double* props, const double* nodal_mass,
                     const double* f_nodal, const double* r_local,
                     double* u, double* endforces, double* reactions,
                     double* lam, double* phi, int32_t* info, int32_t* sweeps, void* stream);

/* ---- modal engine of the fused path (process-wide setting):
 *      BFE2_ENGINE_JACOBI     (default) one-sided Jacobi on the pivoted-Cholesky factor, any n_free <= BFE2_MAX_JACOBI_N
 *      BFE2_ENGINE_TWO_ENDED  for n_free <= 64 and n_modes <= 16: Householder tridiagonalisation of C = Lm^-1 K Lm^-T
 *                             AND of C^-1, Sturm bisection -- every eigenvalue from the end of the spectrum it is
 *                             closer to, so the small ones keep their relative accuracy, which the plain LAPACK route
 *                             (scipy's dsygvd included) does not --, inverse iteration + back-transformation for the
 *                             shapes (bfe2_te.cuh); a structure it declines is repaired by the Jacobi engine in the same
 *                             call.  Both engines meet the 1e-10 bar against the 40-digit truth fixture; on B200 the
 *                             Jacobi engine is 2.5x faster at n_free = 57 (profiles/r2_te_engine.md), hence the default. */
#define BFE2_ENGINE_JACOBI    0
#define BFE2_ENGINE_TWO_ENDED 1
#define BFE2_INFO_RETRY   (-3)
int bfe2_set_modal_engine(int32_t engine);
int bfe2_get_modal_engine(void);

/* ---- static-only requests (do_modal = 0) of bfe2_solve_fused run the 16-lanes-per-structure group kernel.  With the
 *      environment variable BFE2_STATIC_TILE=1 (read once per process) batches of >= 8192 structures
 *      (BFE2_STATIC_TILE_MIN=<batch> moves the threshold) with half bandwidth <= 5 whose tile fits twice into an SM run
 *      the one-lane-per-structure tile kernel instead (bfe2_tile.cu) -- same results to rounding, same speed
 *      (profiles/r2_static_tile.md), kept as a measured experiment.  The counter says how many launches of the tile
 *      kernel this process has made. */
int64_t bfe2_static_tile_launches(void);

/* ---- the same fused path on PACKED RECORDS with compact load / result layouts: the end-to-end entry for data
 *      that lives in HOST memory (one copy per direction per chunk; only numbers that carry information travel).
 *      A pattern is created once per (plan, load layout):
 *        f_index [nnz_f]  dense indices lc*sumdof + pos of the entries of f_nodal that may be non-zero
 *                         (Structure.add_nodal_load touches a handful of positions, Structure.py:258-287)
 *        r_index [nnz_r]  dense indices (lc*nE + e)*6 + k into r_local (Structure.add_internal_loads, :244-248)
 *      (HOST pointers; entries outside the pattern are zero for every structure of the batch).
 *      Input record  [w_in]  = coords [nN,2] | props [nE,4] | nodal_mass [nN,2] | f values [nnz_f] | r values [nnz_r]
 *      Output record [w_out] = u [C, n] on the FREE DOFs (the reference re-inserts the zero rows on the host,
 *                              Solver.py:57-65) | endforces [C, nE, 6] | reactions [C, n_fixed] at the supported
 *                              positions in positions_to_eliminate order (all Results.py:129-143 ever reads) |
 *                              lam [n] | phi [n_modes, n] on the free DOFs | info, sweeps (two int32 in one slot)
 *      bfe2_pattern_layout reports the widths and the offsets (in doubles) of the 5 input / 6 output fields.
 *      in_rec [B, w_in], out_rec [B, w_out] are DEVICE pointers.  Results are bit-identical to bfe2_solve_fused. ---- */
typedef struct bfe2_pattern bfe2_pattern;
int  bfe2_pattern_create(bfe2_pattern** out, bfe2_plan* plan, int32_t C, int32_t do_modal, int32_t n_modes,
                         int32_t nnz_f, const int32_t* f_index, int32_t nnz_r, const int32_t* r_index);
void bfe2_pattern_destroy(bfe2_pattern* pattern);
int  bfe2_pattern_layout(const bfe2_pattern* pattern, int32_t* in_width, int32_t* out_width,
                         int32_t* in_offsets /*[5] or NULL*/, int32_t* out_offsets /*[6] or NULL*/);
int  bfe2_solve_fused_packed(bfe2_plan* plan, bfe2_pattern* pattern, int64_t B, const double* in_rec,
                             double* out_rec, void* stream);

/* ---- small dense generalized symmetric eigenproblem A q = theta B q, A, B [Bt, n, n] row-major, B SPD
 *      (the Rayleigh-Ritz step of the subspace iteration below; same Cholesky-reduction + Jacobi kernel).
 *      V [Bt, n, n]: row k = k-th eigenvector, V B V^T = I.  workspace: 2*Bt*n*n doubles. ------------- */
int bfe2_plan_create_dense(bfe2_plan** out, int32_t n);
int bfe2_sym_eig_dense(bfe2_plan* dense_plan, int64_t Bt, const double* A, const double* Bm, double* lam,
                       double* V, int32_t* info, int32_t* sweeps, double* workspace, void* stream);

/* ---- one LARGE chain structure (one beam, 1e5 elements).  The reference cannot
 *      run this size at all (helpers.py:17 allocates a dense sumdof^2 matrix); the CPU comparators are
 *      scipy.linalg.solveh_banded and scipy.sparse.linalg.eigsh(sigma=0).
 *      coords [n_elem+1, 2], props [n_elem, 4] are DEVICE pointers, fixed_mask [3 (n_elem+1)] a HOST
 *      pointer (1 = supported DOF, Structure.py:164-181).  The handle owns the block-tridiagonal K, M and
 *      the partitioned Cholesky factor of K (device memory, freed by bfe2_beam_destroy).
 *      Multi-vectors are [n_dof, nrhs] row-major (the nrhs values of a DOF are contiguous).
 *      Precision: cond(K) ~ 16 n^4 (1.6e21 at n = 1e5), so the element matrices, the factorisation and the
 *      substitutions run in double-double (BFE2_BEAM_DD, the default of bfe2_beam_create; the inputs are exact
 *      FP64 numbers, eps * cond stays ~2e-11); BFE2_BEAM_FP64 instantiates the same kernels in plain FP64.
 *      Partitioning (n_partitions): the chain is cut into partitions separated by single nodes, factorised
 *      independently; the separator system is again a block tridiagonal chain.  > 0: ONE level with that many
 *      partitions, the separator system eliminated sequentially.  < 0: TWO levels with -n_partitions partitions on the
 *      first -- the separator system is cut the same way once more (~sqrt of it), only ITS separator system is
 *      sequential.  0: automatic -- one level of ~sqrt(N) partitions below 4096 nodes, two levels above with the three
 *      sequential chains of a solve (interiors, second-level interiors, last system) ~cbrt(N) steps each.
 *      A handle owns scratch for the low words of a running solve: use it from ONE stream at a time. ---- */
typedef struct bfe2_beam bfe2_beam;
#define BFE2_BEAM_FP64 0
#define BFE2_BEAM_DD   1
int  bfe2_beam_create(bfe2_beam** out, int64_t n_elem, const double* coords, const double* props,
                      const uint8_t* fixed_mask, int32_t n_partitions /*0 = auto*/, void* stream);
int  bfe2_beam_create_ex(bfe2_beam** out, int64_t n_elem, const double* coords, const double* props,
                         const uint8_t* fixed_mask, int32_t n_partitions /*0 = auto*/, int32_t precision,
                         void* stream);
void bfe2_beam_destroy(bfe2_beam* beam);
int  bfe2_beam_dims(const bfe2_beam* beam, int64_t* n_nodes, int64_t* n_dof, int32_t* n_partitions);
/* U = K^-1 F (Solver.py:76-87 on the condensed system; supported DOFs come out 0 when F is 0 there) */
int  bfe2_beam_solve(bfe2_beam* beam, int32_t nrhs, const double* F, double* U, void* stream);
/* the same with the low words of the double-double right-hand side / solution exposed: F + F_lo in (F_lo may be
 * NULL), U + U_lo out (U_lo may be NULL; all zero for a BFE2_BEAM_FP64 handle) */
int  bfe2_beam_solve_dd(bfe2_beam* beam, int32_t nrhs, const double* F, const double* F_lo /*or NULL*/,
                        double* U, double* U_lo /*or NULL*/, void* stream);
/* Y = K X (which = 0) or Y = M X (which = 1); accumulated in the handle's precision, rounded once */
int  bfe2_beam_matvec(bfe2_beam* beam, int32_t which, int32_t nrhs, const double* X, double* Y, void* stream);
/* R = F - K (U + U_lo), evaluated in the handle's precision and rounded once (U_lo may be NULL) */
int  bfe2_beam_residual(bfe2_beam* beam, int32_t nrhs, const double* F, const double* U,
                        const double* U_lo /*or NULL*/, double* R, void* stream);
/* introspection: the stored K (which = 0) or M (1) rounded to FP64, out [n_nodes, 2, 3, 3] = diagonal block
 * (k, k) and sub-diagonal block (k+1, k) of every node */
int  bfe2_beam_export_blocks(bfe2_beam* beam, int32_t which, double* out, void* stream);
/* FP64 tensor-core (DMMA) contractions for blocks of exactly 64 vectors:
 *   C [64,64] = X^T Y  for X, Y [n, 64];   Xout [n,64] = Z [n,64] * Q [64,64] */
size_t bfe2_gram_workspace_bytes(int64_t n);
int  bfe2_gram64(int64_t n, const double* X, const double* Y, double* C, double* workspace, void* stream);
int  bfe2_update64(int64_t n, const double* Z, const double* Q, double* Xout, void* stream);

/* ---- post-processing next to the hot path.
 *   internal actions: N, V, M at npts equidistant positions of every beam from the local end forces
 *     (HermitianBeam_2D.py:398-431) plus the add-on curves of uniform beam loads q_axial / q_perp
 *     [B, C, nE] or NULL (Loads.py:214-225, 277-289); actions [B, C, nE, 3, npts] and/or
 *     maxabs [B, C, nE, 3] (either may be NULL).
 *   modal participation: gamma [B, n_modes, 2] = phi_k^T M r_x, phi_k^T M r_y (Examples/modal_masses.py:40-49;
 *     effective modal mass = gamma^2).  `plan` and Mb are those of the UNSUPPORTED structure (plan created
 *     with n_fixed = 0, Mb from bfe2_assemble_banded on it): the reference's arithmetic uses the full M, whose
 *     columns at the supports couple the ground motion into the free DOFs. */
int bfe2_internal_actions(bfe2_plan* plan, int64_t B, int32_t C, int32_t npts, const double* coords,
                          const double* endforces, const double* q_axial, const double* q_perp,
                          double* actions, double* maxabs, void* stream);
/*   ... plus the add-on curves of up to n_conc CONCENTRATED loads per beam (Loads.py:347-365 perpendicular force,
 *     :421-432 axial force, :489-502 moment): conc [B, C, nE, n_conc, 3] = (kind, value, normalised position), kind
 *     one of BFE2_LOAD_* (0 = empty slot).  Sampled values use the reference's own comparisons at xi == position;
 *     maxabs also takes both one-sided limits at every load position (the curves jump / kink there). */
#define BFE2_LOAD_NONE        0
#define BFE2_LOAD_CONC_PERP   1
#define BFE2_LOAD_CONC_AXIAL  2
#define BFE2_LOAD_CONC_MOMENT 3
int bfe2_internal_actions_ex(bfe2_plan* plan, int64_t B, int32_t C, int32_t npts, const double* coords,
                             const double* endforces, const double* q_axial, const double* q_perp,
                             int32_t n_conc, const double* conc /*or NULL*/,
                             double* actions, double* maxabs, void* stream);
int bfe2_modal_participation(bfe2_plan* plan, int64_t B, int32_t n_modes, const double* Mb, const double* phi,
                             double* gamma, void* stream);

/* ---- Y [n, nrhs] = A X for a symmetric band matrix in the lower-band layout of bfe2_assemble_banded ([hbw+1, n],
 *      ONE structure) and a row-major multi-vector: the K X / M X products of the first-k modal route of large frames
 *      (beamfe2_b200/frame_modes.py; the reference forms dense products, Solver.py:39). */
int bfe2_band_matvec(int32_t n, int32_t hbw, int32_t nrhs, const double* band, const double* X, double* Y,
                     void* stream);
/*      ... and the batched forms of the three block products (band [B, hbw+1, n]; X, Y, Z, Xout [B, n, nrhs = 64];
 *      C, Q [B, 64, 64]; B <= 65535) for the first-k modes of MANY frames at once
 *      (beamfe2_b200/frame_modes.BatchedFrameModes: 116 < n_free). */
int bfe2_band_matvec_batched(int32_t n, int32_t hbw, int32_t nrhs, int64_t B, const double* band, const double* X,
                             double* Y, void* stream);
int bfe2_gram64_batched(int64_t n, int64_t B, const double* X, const double* Y, double* C, void* stream);
int bfe2_update64_batched(int64_t n, int64_t B, const double* Z, const double* Q, double* Xout, void* stream);

/* ---- geometric stiffness and linear buckling (HermitianBeam_2D.py:261-282 _Ke_geom,
 *      Structure.py:211-214 K_geom, helpers.py:22-27 geometrical branch, Solver.py:90-180 BucklingSolver -- unfinished
 *      in the reference, README.md:27).  Kgb [B, hbw+1, n]: condensed band of K_g for the axial forces
 *      axial [B, nE] (tension positive; the static end forces give N_e = endforces[.., e, 3]).  The buckling pencil
 *      (K + lambda K_g) phi = 0 is solved with bfe2_modal_solve on (K_g + sigma K, K): see beamfe2_b200/batch.py. */
int bfe2_assemble_geometric(bfe2_plan* plan, int64_t B, const double* coords, const double* axial, double* Kgb,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BFE2_H */
