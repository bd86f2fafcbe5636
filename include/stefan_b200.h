/* stefan_b200.h -- C ABI of the B200-native StefanProblem hot path.
 *
 * One shared library (libstefan_b200.so, CUDA kernels built for sm_100a) behind
 * plain-C entry points: pointers and sizes only, int status returns (0 = ok, never
 * throws across the boundary; stefan_last_error() gives the message), opaque
 * handles, caller-owned DEVICE buffers unless the name ends in _host, an explicit
 * cudaStream_t passed as void*.  No hidden global state; one handle per (GPU,
 * host thread).  There is NO CPU fallback: every compute entry point needs a
 * CUDA device.
 *
 * The reference (ArjunNarayanan/StefanProblem) has no FFI of its own; the seam
 * these entry points replace is its Julia call surface, cited per entry as
 * file:line under /root/reference.  The Julia `ccall` module that keeps that
 * call surface is julia/StefanB200.jl; its Python twin (used by all tests,
 * because no Julia runtime exists in the build image) is stefanproblem_b200/.
 *
 * Dof order is the reference's: cells y-fastest (cell = i*ny + j), nodes
 * consecutive per cell, eta-node fastest (a = ix*(p+1) + iy), dof = node*dpn + c.
 * All reals are fp64.
 */
#ifndef STEFAN_B200_H
#define STEFAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  STEFAN_OK = 0,
  STEFAN_BAD_ARGUMENT = 1,
  STEFAN_CUDA_ERROR = 2,
  STEFAN_UNSUPPORTED = 3,
  STEFAN_OUT_OF_RANGE = 4,
  STEFAN_CAPACITY = 5
};

const char* stefan_last_error(void);
int stefan_version(void);

/* ------------------------------------------------------------------ tables
 * 1-D reference-element matrices the uniform-mesh operators factor into
 * (host only, no GPU needed): mass sum_q w N N', stiffness sum_q w N' N'',
 * advection sum_q w N' N', and N'(-1), N'(+1).  Row-major (order+1)^2 / (order+1).
 * Restates PolynomialBasis.LagrangeTensorProductBasis + tensor_product_quadrature
 * as used by IP_gradient_operator.jl:1-9 and LDG_div_operator.jl:1-11. */
int stefan_reference_tables(int32_t order, int32_t numqp, double* mass, double* stiff,
                            double* adv, double* dleft, double* dright,
                            double* qpoints, double* qweights);

/* ------------------------------------------------------------------- IP-DG
 * Symmetric interior-penalty operator for -div(k grad T) on an uncut uniform
 * nx x ny mesh of Q_order cells, two phases (phase +1 -> k1, -1 -> k2).
 *
 * `terms` selects which of the reference's assemble_* passes are present
 * (STEFAN_IP_* below, OR-ed); STEFAN_IP_REFERENCE_SYSTEM is what
 * assemble_interior_penalty_linear_system! (IP_assembly.jl:1-80) assembles on an
 * uncut mesh, where faces between unlike phases are NOT coupled
 * (IP_flux_operator.jl:181).  The INTERFACE_* and ROBIN bits add the reference's
 * face-level interface terms (assemble_face_flux_operator! IP_flux_operator.jl:35-118,
 * assemble_face_penalty_operator! IP_penalty_operator.jl:21-65,
 * assemble_face_robin_operator! IP_robin_interface_operator.jl:1-45) on the
 * mesh faces between unlike phases, side 1 = the +1 cell.
 * variant 1 = the legacy boundary form that reproduces the stored
 * uncut_mesh_tests/IP_convergence.csv (boundary operator -M', rhs penalty only). */
enum {
  STEFAN_IP_VOLUME = 1,               /* assemble_gradient_operator!            IP_gradient_operator.jl:29-57 */
  STEFAN_IP_INTERELEMENT_FLUX = 2,    /* assemble_interelement_flux_operator!   IP_flux_operator.jl:202-256 */
  STEFAN_IP_INTERELEMENT_PENALTY = 4, /* assemble_interelement_penalty_operator! IP_penalty_operator.jl:134-181 */
  STEFAN_IP_INTERFACE_FLUX = 8,
  STEFAN_IP_INTERFACE_PENALTY = 16,
  STEFAN_IP_BOUNDARY_FLUX = 32,       /* assemble_boundary_flux_operator!       IP_flux_operator.jl:393-444 */
  STEFAN_IP_BOUNDARY_PENALTY = 64,    /* assemble_boundary_penalty_operator!    IP_penalty_operator.jl:286-329 */
  STEFAN_IP_ROBIN = 128,              /* assemble_robin_operator!               IP_robin_interface_operator.jl:76-99 */
  STEFAN_IP_REFERENCE_SYSTEM = 1 | 2 | 4 | 32 | 64,
  STEFAN_IP_ALL_TERMS = 255
};

typedef struct {
  int32_t nx, ny;      /* cells of the local slab */
  double hx, hy;       /* cell widths */
  int32_t order;       /* 1..3 */
  int32_t numqp;       /* Gauss points per direction (reference: order+1) */
  double k1, k2;       /* conductivities of phase +1 / -1 */
  double penalty;      /* penaltyfactor / min element size (set by the caller, as in the reference) */
  double lambda, Tm;   /* Robin coefficient and melting temperature */
  int32_t variant;     /* 0 current, 1 legacy_boundary */
  int32_t terms;       /* STEFAN_IP_* */
} stefan_ipdg_params;

typedef struct stefan_ipdg stefan_ipdg;

/* phase: host int8[nx*ny] (+1/-1; NULL = all +1).  phase_lo / phase_hi: host
 * int8[ny], the phases of the neighbouring slab's adjacent cell column when the
 * mesh is partitioned by cell columns across GPUs (NULL = domain boundary). */
int stefan_ipdg_create(const stefan_ipdg_params* params, const int8_t* phase,
                       const int8_t* phase_lo, const int8_t* phase_hi, stefan_ipdg** out);
int stefan_ipdg_destroy(stefan_ipdg* h);
int64_t stefan_ipdg_ndofs(const stefan_ipdg* h);

/* y = A x (replaces sparse_operator(sysmatrix, mesh, 1) * x).  x_lo / x_hi: the
 * neighbouring slabs' adjacent cell columns (ny*NF doubles each; local or peer
 * device memory), required iff the handle has phase_lo / phase_hi.
 * algo 0 = production kernel (TMA march kernel, csrc/ipdg_march.cuh), 1 = simple
 * one-thread-per-cell kernel (independent second opinion used by the tests). */
int stefan_ipdg_apply(stefan_ipdg* h, const double* x, double* y, const double* x_lo,
                      const double* x_hi, int32_t algo, void* stream);

/* Rows of the cell columns [col_begin, col_end) only.  Columns [1, nx-1) never read
 * x_lo / x_hi, so a column-slab partition can run them while the halo exchange is in
 * flight and finish with the two edge columns. */
int stefan_ipdg_apply_columns(stefan_ipdg* h, const double* x, double* y, const double* x_lo,
                              const double* x_hi, int32_t col_begin, int32_t col_end, int32_t algo,
                              void* stream);

/* Right-hand side of assemble_interior_penalty_rhs! (IP_assembly.jl:82-115) plus, when
 * the handle has the ROBIN term, assemble_robin_source! (face-level form) -- on the
 * device.  The reference calls its sourceterm(x) / boundaryfunc(x) closures at the
 * physical quadrature points; here the caller passes their samples:
 *   f_qp  [ncells][numqp*numqp]  at the cell quadrature points (x-point outer, y-point
 *         inner; for two phases the host samples f1 / f2 by cell phase), or NULL
 *   g_qp  [bottom nx*numqp][right ny*numqp][top nx*numqp][left ny*numqp] at the face
 *         quadrature points of the boundary faces, or NULL
 *   b     [ndofs] is overwritten.
 * Slab handles (phase_lo / phase_hi given): f_qp and the bottom / top parts of g_qp are the
 * slab's own; the left / right parts are read only where the slab touches the domain boundary. */
int stefan_ipdg_rhs(stefan_ipdg* h, const double* f_qp, const double* g_qp, double* b, void* stream);

/* The COO seam: the (row, col, val) triplets of the assembled operator, one dense
 * NF x NF block per cell and per existing neighbour (duplicates already summed, explicit
 * zeros kept), column-major inside a block like vec(M).  Feed them to
 * sparse(rows, cols, vals, N, N) exactly as CutCellDG.sparse_operator does.
 * index_base 1 gives Julia indices. */
int64_t stefan_ipdg_coo_size(const stefan_ipdg* h);
int stefan_ipdg_assemble_coo(stefan_ipdg* h, int64_t* rows, int64_t* cols, double* vals,
                             int64_t capacity, int32_t index_base, int64_t* nnz_out, void* stream);

/* Interface sums that drive the front velocity, over the mesh faces between a +1 cell
 * (side 1) and a -1 cell (side 2), n = outward normal of side 1:
 *   sums[0] = length, sums[1] = int lambda (Tm - {T}), sums[2] = int (k1 dn T1 - k2 dn T2),
 *   sums[3] = int {T}.
 * Face-integrated form of the point samples of IP_robin_interface_flux_velocity.jl:213-278
 * (values / gradients as IP_postprocess.jl:1-52).  Faces are owned by their +1 cell, so the
 * per-slab sums of a column partition add up with one 4-double all-reduce.  sums: device. */
int stefan_ipdg_interface_sums(stefan_ipdg* h, const double* u, const double* u_lo, const double* u_hi,
                               double* sums_dev, void* stream);

/* u <- u + dt Minv (b - y) with y = A u from stefan_ipdg_apply and M the block-diagonal DG
 * mass matrix: explicit Euler step of M du/dt = b - A u (b may be NULL).  No reference
 * implementation exists (the reference only solves steady problems); defined in
 * oracle/timeloop.py. */
int stefan_ipdg_euler_update(stefan_ipdg* h, double* u, const double* y, const double* b, double dt,
                             void* stream);

/* Solve A x = b by preconditioned conjugate gradients on the device (replaces the sparse
 * direct solve `matrix \ rhs` of the reference's drivers, e.g.
 * uncut_mesh_tests/test_uncut_IP_convergence.jl:90, for the symmetric `current` variant).
 * x: initial guess in, solution out.  preconditioner 0 = none, 1 = block Jacobi (inverse of
 * every cell's NF x NF diagonal block).  The residual is read back every check_every
 * iterations (<= 0: 10).  Synchronous with respect to `stream`. */
typedef struct {
  int32_t max_iterations;
  double rel_tolerance;   /* stop when ||b - A x|| <= rel_tolerance ||b|| */
  int32_t preconditioner;
  int32_t check_every;
} stefan_cg_params;
typedef struct {
  int32_t iterations;
  double residual_norm, rhs_norm;
  int32_t converged;
} stefan_cg_result;
int stefan_ipdg_cg_solve(stefan_ipdg* h, const double* b, double* x, const stefan_cg_params* params,
                         stefan_cg_result* result, void* stream);

/* The same solve across column slabs, one process per GPU: every rank calls it with its slab handle
 * (created with phase_lo / phase_hi) and its slab of b and x.  Two callbacks carry what crosses the slabs:
 *   exchange(ctx, v, stream): fill the caller's halo buffers halo_lo / halo_hi (one cell column each, layout of
 *     stefan_ipdg_apply's x_lo / x_hi) with the neighbours' edge columns of the slab vector v, ordered on `stream`
 *     (NCCL send / recv, or peer copies); called before every operator application;
 *   allreduce(ctx, dev, count, stream): sum `count` device doubles over the ranks in place, ordered on `stream`
 *     (the r.z, r.r and p.Ap sums; everything else in the iteration is rank-local).
 * Both return 0 on success.  `stream` inside the callbacks is the solver's own stream, not the caller's.  The
 * residual every rank reads back is the all-reduced one, so all ranks stop at the same iteration.  Null callbacks
 * on a handle without neighbours: identical to stefan_ipdg_cg_solve (without its CUDA-graph batching). */
typedef int (*stefan_halo_fn)(void* ctx, const double* v, void* stream);
typedef int (*stefan_allreduce_fn)(void* ctx, double* dev, int32_t count, void* stream);
int stefan_ipdg_cg_solve_slab(stefan_ipdg* h, const double* b, double* x, const stefan_cg_params* params,
                              stefan_cg_result* result, const double* halo_lo, const double* halo_hi,
                              stefan_halo_fn exchange, stefan_allreduce_fn allreduce, void* ctx, void* stream);

/* mesh_L2_error (useful_routines.jl:95-133) on the device: out2[0] = ||u_h - u_ex||_L2,
 * out2[1] = ||u_ex||_L2 with the cell quadrature of the handle; uex_qp [ncells][numqp*numqp]
 * are the exact solution's samples at the cell quadrature points (layout of f_qp). */
int stefan_ipdg_l2_error(stefan_ipdg* h, const double* u, const double* uex_qp, double* out2_dev, void* stream);

/* Same operation with HOST buffers (pinned memory recommended): host->device
 * copies, kernels and device->host copies of `nslabs` column slabs are pipelined
 * on three streams inside the call (nslabs <= 0: default).  Synchronous. */
int stefan_ipdg_apply_host(stefan_ipdg* h, const double* x_host, double* y_host, int32_t nslabs);

/* -------------------------------------------------------------------- LDG
 * Local-DG mixed form, 3 dofs per node (T, q_x, q_y), of
 * assemble_LDG_linear_system! (StefanDG/LocalDG/LDG_assembly.jl:1-91) on an uncut
 * uniform mesh: divergence + mass + gradient cell operators, inter-element scalar and
 * vector fluxes with the directional switch V0, boundary vector flux.  As in the
 * reference, V0 itself is used as `beta` in the vector flux (LDG_assembly.jl:44-53) and
 * faces between unlike phases are not coupled by the reference's drivers (they couple phases
 * only inside cut cells).  aligned_interface != 0 (extension, SURVEY.md 8(d)(ii), the LDG analogue of the
 * STEFAN_IP_INTERFACE_* terms): the reference's face-level functions
 * assemble_face_scalar_flux_operator! / assemble_face_vector_flux_operator! are also applied to the
 * faces between uncut cells of unlike phase, in both orientations, with alpha = interfacepenalty and
 * the neighbour's conductivity on the neighbour's flux -- what LDG_scalar_flux_operator.jl:241-283 and
 * LDG_vector_flux_operator.jl:318-355 do with the two sides of a cut cell's interface. */
typedef struct {
  int32_t nx, ny;
  double hx, hy;
  int32_t order, numqp;
  double k1, k2;
  double interiorpenalty, negboundarypenalty, posboundarypenalty;
  double V0x, V0y;
  double interfacepenalty;      /* used only with aligned_interface */
  int32_t aligned_interface;
} stefan_ldg_params;

typedef struct stefan_ldg stefan_ldg;
int stefan_ldg_create(const stefan_ldg_params* params, const int8_t* phase, const int8_t* phase_lo,
                      const int8_t* phase_hi, stefan_ldg** out);
int stefan_ldg_destroy(stefan_ldg* h);
int64_t stefan_ldg_ndofs(const stefan_ldg* h);
/* y = A x (replaces sparse_operator(sysmatrix, mesh, 3) * x); dof = 3*node + component.
 * x_lo / x_hi: the neighbouring slabs' adjacent cell columns (3*ny*NF doubles each). */
int stefan_ldg_apply(stefan_ldg* h, const double* x, double* y, const double* x_lo,
                     const double* x_hi, void* stream);
/* Same with an explicit kernel choice: 0 = what stefan_ldg_apply does (the march kernel;
 * order 3 below 1 M cells: the plain kernel), 1 = plain one-thread-per-cell kernel (the
 * independent second opinion of the tests), 2 = march kernel (csrc/ldg_march.cuh). */
int stefan_ldg_apply_algo(stefan_ldg* h, const double* x, double* y, const double* x_lo,
                          const double* x_hi, int32_t algo, void* stream);

/* `matrix \ rhs` for the local-DG system on the device (replaces the sparse direct solve of
 * uncut_mesh_tests/test_uncut_LDG_convergence.jl:96): q is eliminated through the block-diagonal cell
 * mass matrix (LDG_mass_operator.jl:1-25) and BiCGStab runs on the Schur complement in T; one Schur
 * application = two launches of the apply kernel around a per-cell M^-1.  All iteration scalars stay on
 * the device; the residual is read back every params->check_every iterations (<= 0: 10);
 * params->preconditioner is ignored (none).  rhs, sol: 3 dofs per node, interleaved.  result->residual_norm
 * is the TRUE residual norm of the Schur system at exit, rhs_norm the norm of its right-hand side. */
int stefan_ldg_solve(stefan_ldg* h, const double* rhs, double* sol, const stefan_cg_params* params,
                     stefan_cg_result* result, void* stream);

/* ------------------------------------------------------------------- DG1D
 * 1-D two-phase interior-penalty DG with a Robin interface on a DGMesh1D
 * (StefanRobinIC/DG1D/dgmesh_1d.jl:1-51): nelmts1 uniform cells on [xL, interfacepoint]
 * (label 1, k1), nelmts2 on [interfacepoint, xR] (label 2, k2).  `terms` = which passes
 * were assembled: */
enum {
  STEFAN_DG1D_GRADIENT = 1,          /* assemble_gradient_operator!          assemble_gradient_operator.jl:25-48 */
  STEFAN_DG1D_FLUX = 2,              /* assemble_flux_operator!              assemble_flux_operator.jl:93-120 */
  STEFAN_DG1D_PENALTY = 4,           /* assemble_penalty_operator!           assemble_penalty_operator.jl:52-70 */
  STEFAN_DG1D_ROBIN = 8,             /* assemble_interface_robin_operator!   assemble_interface_robin_operator.jl:46-69 */
  STEFAN_DG1D_BOUNDARY_FLUX = 16,    /* assemble_boundary_flux_operator!     assemble_flux_operator.jl:122-173 */
  STEFAN_DG1D_BOUNDARY_PENALTY = 32, /* assemble_boundary_penalty_operator!  assemble_penalty_operator.jl:72-102 */
  STEFAN_DG1D_ALL_TERMS = 63
};
typedef struct {
  int32_t nelmts1, nelmts2;
  double xL, xR, interfacepoint;
  int32_t order, numqp;
  double k1, k2, penalty, lambda, Tm;
  int32_t terms;
} stefan_dg1d_params;
typedef struct stefan_dg1d stefan_dg1d;
int stefan_dg1d_create(const stefan_dg1d_params* params, stefan_dg1d** out);
int stefan_dg1d_destroy(stefan_dg1d* h);
int64_t stefan_dg1d_ndofs(const stefan_dg1d* h);
/* y = A x (replaces DG1D.sparse_operator(sysmatrix, mesh, 1) * x, assembly.jl:1-5). */
int stefan_dg1d_apply(stefan_dg1d* h, const double* x, double* y, void* stream);
/* b = two-phase source (f_qp [ncells][numqp] samples of rhsfunc1 / rhsfunc2 at the cell
 * quadrature points, or NULL) + Robin source (if the ROBIN term is set) + boundary rhs for
 * Dirichlet values TL, TR (if with_boundary): assemble_source_term.jl:26-61,
 * assemble_interface_robin_source.jl:19-43, assemble_boundary_conditions.jl:1-44. */
int stefan_dg1d_rhs(stefan_dg1d* h, const double* f_qp, double TL, double TR, int32_t with_boundary,
                    double* b, void* stream);

/* The COO seams of the LDG and DG1D systems (as stefan_ipdg_assemble_coo: dense blocks per
 * cell and per existing neighbour, column-major inside a block, feed them to sparse()). */
int64_t stefan_ldg_coo_size(const stefan_ldg* h);
int stefan_ldg_assemble_coo(stefan_ldg* h, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                            int32_t index_base, int64_t* nnz_out, void* stream);
int64_t stefan_dg1d_coo_size(const stefan_dg1d* h);
int stefan_dg1d_assemble_coo(stefan_dg1d* h, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                             int32_t index_base, int64_t* nnz_out, void* stream);

/* Right-hand side of assemble_LDG_rhs! (LDG_assembly.jl:93-131); f_qp / g_qp sampled as for
 * stefan_ipdg_rhs; b [3*nnodes] is overwritten. */
int stefan_ldg_rhs(stefan_ldg* h, const double* f_qp, const double* g_qp, double* b, void* stream);

/* --------------------------------------------------------------------- FD
 * Two-phase 1-D finite-difference Laplace operator of module FiniteDifference
 * (StefanFD/finite-difference.jl).  phaseid: host int8[numnodes] as produced by
 * phase_id (:92-103; 1 where phi < 0, else 2).  stefan_fd_create classifies every
 * row once, following assemble_laplace_operator! (:71-85) -- central / backward /
 * forward second derivative, or no row -- and, if `dirichlet`, the identity rows
 * of assemble_dirichlet_boundary_condition! (:87-90).  Where the reference would
 * index phaseid out of bounds (interface within 3 nodes of an end) it returns
 * STEFAN_OUT_OF_RANGE.  phase_lo / phase_hi: host int8[3], the neighbouring slab's
 * three adjacent nodes when the line is partitioned across GPUs (else NULL). */
typedef struct stefan_fd stefan_fd;
int stefan_fd_create(int64_t numnodes, double stepsize, const int8_t* phaseid, int32_t dirichlet,
                     const int8_t* phase_lo, const int8_t* phase_hi, stefan_fd** out);
int stefan_fd_destroy(stefan_fd* h);
/* assemble_interface_continuity!(matrix, r, s, nodeid) (:111-116); nodeid 0-based. */
int stefan_fd_add_interface_continuity(stefan_fd* h, double r, double s, int64_t nodeid);
/* y = K u (replaces sparse_operator(matrix, ndofs) * u, :105-109).  u_lo / u_hi: the
 * neighbouring slabs' 3 adjacent nodes (device), iff phase_lo / phase_hi were given. */
int stefan_fd_apply(stefan_fd* h, const double* u, double* y, const double* u_lo,
                    const double* u_hi, void* stream);
/* Fused explicit heat step (no reference implementation; defined in oracle/timeloop.py):
 * unew = u + dt * kappa(phase) * (K u) on the Laplace rows, unew = u on all other rows. */
int stefan_fd_step(stefan_fd* h, const double* u, double* unew, const double* u_lo,
                   const double* u_hi, double dt, double kappa1, double kappa2, void* stream);
/* The COO triplets the reference appends (0-based, ordered by row), on the device. */
int64_t stefan_fd_nnz(stefan_fd* h);
int stefan_fd_assemble_coo(stefan_fd* h, int64_t* rows, int64_t* cols, double* vals,
                           int64_t capacity, int64_t* nnz_out, void* stream);

/* ------------------------------------------------------- element-block path
 * The reference's element-local assembly for ARBITRARY per-cell / per-face quadratures (what
 * a cut-cell mesh produces: CellQuadratures, FaceQuadratures, InterfaceQuadratures), and the
 * application of the resulting dense blocks.  A "cell" is a solution cell (a cut cell
 * contributes one per phase, with the interface as an extra face slot between the two).
 * For every cell: gradient_operator (IP_gradient_operator.jl:1-9); for every face slot, seen
 * from the cell (side 1, outward normal): the (1,1) and (1,2) blocks of
 * assemble_face_flux_operator! (IP_flux_operator.jl:35-118), assemble_face_penalty_operator!
 * (IP_penalty_operator.jl:21-65) and assemble_face_robin_operator!
 * (IP_robin_interface_operator.jl:1-45); boundary slots: -(M + M') + pen P
 * (IP_flux_operator.jl:330-357, IP_penalty_operator.jl:239-254).  All arrays are DEVICE arrays. */
typedef struct {
  int32_t ncells;   /* solution cells */
  int32_t order;    /* 1..3 */
  int32_t nqv;      /* volume quadrature points per cell (pad with weight 0) */
  int32_t nslots;   /* face slots per cell, <= 16 */
  int32_t nqf;      /* quadrature points per face slot (pad with weight 0) */
} stefan_blocks_dims;
typedef struct {
  const double* vol_pts;       /* [ncells][nqv][2]  reference coordinates */
  const double* vol_wts;       /* [ncells][nqv]     quadrature weights (detjac is applied inside) */
  const double* cell_jac;      /* [ncells][2]       jacobian (hx/2, hy/2) */
  const double* cell_k;        /* [ncells]          conductivity */
  const int32_t* slot_nbr;     /* [ncells][nslots]  neighbour cell; -1 boundary face; -2 empty slot */
  const double* face_pts;      /* [ncells][nslots][nqf][2]  own reference coordinates */
  const double* face_nbr_pts;  /* [ncells][nslots][nqf][2]  the neighbour's reference coordinates */
  const double* face_wts;      /* [ncells][nslots][nqf]     weight x scale area */
  const double* face_normals;  /* [ncells][nslots][nqf][2]  pointing out of the cell */
  const double* slot_pen;      /* [ncells][nslots]  penalty */
  const double* slot_lambda;   /* [ncells][nslots]  Robin coefficient (0: no Robin term) */
} stefan_blocks_data;
typedef struct stefan_blocks stefan_blocks;
int stefan_blocks_create(const stefan_blocks_dims* dims, stefan_blocks** out);
int stefan_blocks_destroy(stefan_blocks* h);
int64_t stefan_blocks_ndofs(const stefan_blocks* h);
/* Boundary form used by stefan_blocks_assemble / stefan_blocks_rhs: 0 = current (boundary block -(M + M') + pen P,
 * rhs pen sum V g w - sum k (G.n) g w; IP_flux_operator.jl:330-357, IP_source_term.jl:131-159), 1 = the legacy
 * boundary form (-M' + pen P, rhs pen sum V g w) that the reference's stored cut-cell CSVs of the interior-penalty
 * method were produced with (as stefan_ipdg_params.variant).  Call before stefan_blocks_assemble. */
int stefan_blocks_set_variant(stefan_blocks* h, int32_t variant);
/* fills the blocks [ncells][1 + nslots][NF][NF] (diag, then one per slot) */
int stefan_blocks_assemble(stefan_blocks* h, const stefan_blocks_data* data, void* stream);
/* y = A x with the assembled blocks; 16 + 8 (1 + nslots) NF bytes of HBM traffic per dof */
int stefan_blocks_apply(stefan_blocks* h, const double* x, double* y, void* stream);
/* the COO seam (explicit zeros for empty / boundary slots, as a dense block per slot) */
int64_t stefan_blocks_coo_size(const stefan_blocks* h);
int stefan_blocks_assemble_coo(stefan_blocks* h, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                               int32_t index_base, int64_t* nnz_out, void* stream);
/* Right-hand side for the same data (what assemble_interior_penalty_rhs! + assemble_two_phase_source! +
 * assemble_robin_source! add; IP_source_term.jl:97-159, IP_robin_interface_source.jl:1-25): per solution cell
 *   sum V f detjac w  +  boundary slots: sum (pen V - k (G.n)) g w  +  slots with lambda != 0: 1/2 lambda Tm sum V w.
 * f_qp [ncells][nqv]: source at the volume points (NULL: none); g_qp [ncells][nslots][nqf]: boundary data at the
 * points of the boundary slots (NULL: none).  rhs: [ncells * NF] device doubles. */
int stefan_blocks_rhs(stefan_blocks* h, const stefan_blocks_data* data, const double* f_qp, const double* g_qp,
                      double Tm, double* rhs, void* stream);
/* mesh_L2_error / integral_norm_on_mesh (useful_routines.jl:95-133, :203-221) over the volume points of all
 * solution cells: out2[0] = ||u_h - u_ex||, out2[1] = ||u_ex||; uex_qp [ncells][nqv]. */
int stefan_blocks_l2_error(stefan_blocks* h, const stefan_blocks_data* data, const double* u, const double* uex_qp,
                           double* out2_dev, void* stream);
/* `matrix \ rhs` on the device: conjugate gradients over stefan_blocks_apply (the cut-cell IP systems are
 * symmetric: `@assert issymmetric`, IP_test_cylindrical_bc_convergence.jl:88), block-Jacobi preconditioner
 * (params->preconditioner 1: inverse of every solution cell's diagonal block), scalars on the device. */
int stefan_blocks_cg_solve(stefan_blocks* h, const double* b, double* x, const stefan_cg_params* params,
                           stefan_cg_result* result, void* stream);
/* interface_L2_error / integral_norm_on_interface (useful_routines.jl:239-274, :288-307) on the face slots
 * selected by slot_select [ncells][nslots] (non-zero = include; e.g. the interface slots of the solution cells
 * of one level-set sign): out2[0] = sqrt(sum (u_h - u_ex)^2 sa w), out2[1] = sqrt(sum u_ex^2 sa w);
 * uex_fq [ncells][nslots][nqf] are the exact values at the slots' points. */
int stefan_blocks_interface_l2_error(stefan_blocks* h, const stefan_blocks_data* data, const int8_t* slot_select,
                                     const double* u, const double* uex_fq, double* out2_dev, void* stream);

/* ------------------------------------------ local DG on element blocks (cut cells; SURVEY 8(a) LDG interface drivers)
 * The local-DG twin of stefan_blocks_*: `assemble_LDG_linear_system!` (LDG_assembly.jl:1-91) for arbitrary per-cell /
 * per-face quadratures in the same `stefan_blocks_data` layout -- divergence_operator (LDG_div_operator.jl:1-11),
 * mass_matrix (LDG_mass_operator.jl:1-25), gradient_operator (LDG_gradient_operator.jl:1-11) from the volume points;
 * assemble_face_scalar_flux_operator! (LDG_scalar_flux_operator.jl:25-107) and assemble_face_vector_flux_operator!
 * (LDG_vector_flux_operator.jl:47-154) on every neighbour slot, which is what the interelement drivers (:111-237,
 * :158-300) AND the cut-cell interface drivers (LDG_scalar_flux_operator.jl:241-309,
 * LDG_vector_flux_operator.jl:318-386) call; assemble_boundary_face_vector_flux_operator! (:391-439) on boundary
 * slots.  3 dofs per node, interleaved (T, q_x, q_y): ndofs = ncells * 3 NF, blocks are 3 NF x 3 NF.
 * `slot_pen` is the slot's penalty (interior / interface penalty; on a boundary slot negboundarypenalty or
 * posboundarypenalty, selected by the producer from sign(V0 . n) as the reference does); `slot_lambda` is ignored.
 * (V0x, V0y): the unit vector V0 (edge switch of the scalar flux; n . V0 in the vector flux). */
typedef struct stefan_ldgblocks stefan_ldgblocks;
int stefan_ldgblocks_create(const stefan_blocks_dims* dims, stefan_ldgblocks** out);
int stefan_ldgblocks_destroy(stefan_ldgblocks* h);
int64_t stefan_ldgblocks_ndofs(const stefan_ldgblocks* h);
int stefan_ldgblocks_assemble(stefan_ldgblocks* h, const stefan_blocks_data* data, double V0x, double V0y, void* stream);
int stefan_ldgblocks_apply(stefan_ldgblocks* h, const double* x, double* y, void* stream);
/* assemble_LDG_rhs! (LDG_assembly.jl:93-131) for the same data: T rows sum V f detjac w (LDG_source_term.jl:1-60); on
 * boundary slots q rows sum n V g w (:64-198) and T rows + slot_pen sum V g w (:203-321).  f_qp [ncells][nqv] or NULL,
 * g_qp [ncells][nslots][nqf] or NULL; rhs [ncells * 3 NF]. */
int stefan_ldgblocks_rhs(stefan_ldgblocks* h, const stefan_blocks_data* data, const double* f_qp, const double* g_qp,
                         double* rhs, void* stream);
int64_t stefan_ldgblocks_coo_size(const stefan_ldgblocks* h);
int stefan_ldgblocks_assemble_coo(stefan_ldgblocks* h, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                                  int32_t index_base, int64_t* nnz_out, void* stream);
/* `matrix \ rhs` (cylindrical_bc_tests/LDG_test_cylindrical_bc_convergence.jl:157-159) on the device: q is eliminated
 * through every solution cell's own mass block (inverted at assemble time) and BiCGStab runs on the Schur complement
 * in T, as stefan_ldg_solve does on uniform meshes.  result as for stefan_ldg_solve. */
int stefan_ldgblocks_solve(stefan_ldgblocks* h, const double* rhs, double* sol, const stefan_cg_params* params,
                           stefan_cg_result* result, void* stream);

/* ------------------------------------------------- point evaluation and front velocity
 * InteriorPenalty.interpolate_at_reference_points (IP_postprocess.jl:30-52) and gradient_at_reference_points
 * (:1-28) as device gathers: nodalvalues holds NF = (order + 1)^2 dofs per (solution) cell, refpoints [numpts][2]
 * are reference coordinates in the cells refcellids [numpts] (0-based; the caller picks the solution cell of the
 * wanted level-set sign, which is what `nodal_connectivity(mesh, levelsetsign, cellid)` does).  vals [numpts];
 * grads [numpts][2] physical gradients (transform_gradient with jacobian (jx, jy) = (hx/2, hy/2)). */
int stefan_interpolate_at_reference_points(int32_t order, const double* nodalvalues, const double* refpoints,
                                           const int32_t* refcellids, int64_t numpts, double* vals, void* stream);
int stefan_gradient_at_reference_points(int32_t order, const double* nodalvalues, const double* refpoints,
                                        const int32_t* refcellids, int64_t numpts, double jx, double jy, double* grads,
                                        void* stream);
/* The velocity formulas of robin_interface_velocity/IP_robin_interface_flux_velocity.jl:213-278 at numpts interface
 * points: velocity_s = -lambda (Tm - vals_s), meanvelocity = their mean, gvelocity = k2 dnT2 - k1 dnT1 with
 * dnT_s = grads_s . normals.  Any output may be NULL; gvelocity needs grads1, grads2, normals ([numpts][2]). */
int stefan_front_velocity(int64_t numpts, double lambda, double Tm, double k1, double k2, const double* vals1,
                          const double* vals2, const double* grads1, const double* grads2, const double* normals,
                          double* velocity1, double* velocity2, double* meanvelocity, double* gvelocity, void* stream);
/* CutCellDG.interpolate_at_reference_points(G, ndofs, ...) for a field with several dofs per node (e.g. the LDG
 * solution: T, q_x, q_y interleaved): `ncomponents` (<= 3) consecutive components from `first_component` on;
 * vals [numpts][ncomponents]. */
int stefan_interpolate_field_at_reference_points(int32_t order, const double* nodalvalues, int32_t dofspernode,
                                                 int32_t first_component, int32_t ncomponents, const double* refpoints,
                                                 const int32_t* refcellids, int64_t numpts, double* vals, void* stream);
/* LocalDG/cylindrical-bc-flux/lagrange_levelset_LDG_interface_flux.jl:266-307: with grads_s = q of side s at the
 * interface points ([numpts][2]) and the interface normals: normalflux_s = k_s grads_s . n, and the LDG numerical flux
 * 1/2 (k1 grads1 + k2 grads2) - beta (normalflux2 - normalflux1), beta = edge_switch(V0, n) (LDG_utilities.jl:1-4),
 * dotted with n.  Any output may be NULL. */
int stefan_ldg_interface_flux(int64_t numpts, double k1, double k2, double V0x, double V0y, const double* grads1,
                              const double* grads2, const double* normals, double* normalflux1, double* normalflux2,
                              double* numericalnormalflux, void* stream);
/* mesh_L2_error / integral_norm_on_mesh (StefanDG/useful_routines.jl:95-133, :203-221) and uniform_mesh_L2_error /
 * integral_norm_on_uniform_mesh (StefanRobinIC/useful_routines.jl:52-78) for any field on an uncut mesh: dim 2
 * (NF = (order+1)^2 nodes, numqp^2 tensor Gauss points per cell, x-point outer) or dim 1; `dofspernode` (<= 3)
 * interleaved dofs per node (1: IP / DG1D, 3: LDG).  detjac: constant cell det-jacobian, or per-cell values in
 * cell_detjac [ncells] (DG1D: two element sizes).  uex_qp [ncells][points][dofspernode].  out_dev [2 * dofspernode]:
 * per component (||u_h - u_ex||, ||u_ex||). */
int stefan_mesh_l2_error(int32_t dim, int32_t order, int32_t numqp, int64_t ncells, double detjac,
                         const double* cell_detjac, int32_t dofspernode, const double* u, const double* uex_qp,
                         double* out_dev, void* stream);

/* ------------------------------------------------------------- peer memory
 * Column-slab partition across the GPUs of one NVLink domain, one process per GPU (the
 * reference is single-process; this is the multi-GPU extension of SURVEY.md 8(e)).  A slab
 * vector allocated here can be mapped by the neighbouring ranks, whose apply kernels then
 * read its edge column directly (pass the mapped address as x_lo / x_hi): the halo exchange
 * becomes TMA copies over NVLink inside the apply kernel.  The 64-byte handle travels
 * between the processes by any means the host program has. */
typedef struct { unsigned char bytes[64]; } stefan_ipc_handle;
int stefan_peer_alloc(int64_t nbytes, void** dev_ptr, stefan_ipc_handle* handle); /* zero-filled */
int stefan_peer_free(void* dev_ptr);
int stefan_peer_open(const stefan_ipc_handle* handle, void** peer_ptr);
int stefan_peer_close(void* peer_ptr);
/* Epoch flags (uint64 in peer memory): signal = "everything this stream wrote before is
 * final for `epoch`" (release, system scope); wait holds the stream until *peer_flag >=
 * epoch (acquire), gives up after timeout_ms and then sets *status_dev (device int32, may
 * be NULL) to 1 instead of hanging. */
int stefan_peer_signal(uint64_t* flag, uint64_t epoch, void* stream);
int stefan_peer_wait(const uint64_t* peer_flag, uint64_t epoch, int32_t timeout_ms, int32_t* status_dev,
                     void* stream);

/* The same ordering fused into the apply kernel (one launch per application, nothing else
 * on the stream): the kernel publishes *my_ready = epoch when it starts, the warps that fetch
 * a neighbour's edge column (x_lo / x_hi = mapped peer addresses) first wait for *lo_ready /
 * *hi_ready >= epoch, and the last CTA to finish publishes *my_done = epoch.  A rank that is
 * about to overwrite its vector waits for its neighbours' my_done flags (stefan_peer_wait). */
typedef struct {
  uint64_t* my_ready;
  uint64_t* my_done;
  const uint64_t* lo_ready;   /* neighbours' my_ready flags in peer memory; NULL where there is none */
  const uint64_t* hi_ready;
  uint64_t epoch;             /* > 0, increasing (ignored, but still > 0, when epoch_dev is given) */
  int32_t timeout_ms;
  int32_t* status;            /* device int32 set to 1 if a wait timed out; may be NULL */
  /* Optional device-resident epoch (local memory, initialised to the last completed epoch, e.g. 0): the kernel
   * works with *epoch_dev + 1 and its last CTA stores that value back.  The launch then carries no
   * per-step host state, so a step captured in a CUDA graph advances the protocol on every replay. */
  uint64_t* epoch_dev;
  /* Optional push-style mailboxes: addresses in the NEIGHBOURS' memory (low, high; NULL = none) where the kernel
   * also stores its ready value (at start) / done value (by its last CTA).  lo_ready / hi_ready may then point at
   * the local slots the neighbours post into, so that the waiting warps poll local memory instead of reading a flag
   * over NVLink (a remote poll is a 1-2 us round trip). */
  uint64_t* post_ready[2];
  uint64_t* post_done[2];
  /* chained != 0: the vector read in epoch k is final as soon as the kernel of epoch k - 1 is complete (a static
   * vector, or the two-buffer loop of stefan_ipdg_step).  lo_ready / hi_ready must then point at DONE-type flags
   * (the neighbours' my_done or the mailboxes they post their done value into) and the kernel waits for epoch - 1:
   * the neighbours' kernels of this epoch need not have started, which takes the launch gap and the flag's flight
   * time off the critical path. */
  int32_t chained;
} stefan_peer_sync;
int stefan_ipdg_apply_peer(stefan_ipdg* h, const double* x, double* y, const double* x_lo, const double* x_hi,
                           const stefan_peer_sync* sync, void* stream);

/* The fused explicit time step (the reference assembles a steady operator and never
 * time-steps, so the definition is oracle/timeloop.py::ip_euler_run):
 *     unew = u + dt * Minv * (b - A u),   M = block-diagonal DG mass matrix (CutCellDG.mass_matrix),
 * in ONE pass of the apply kernel: u and b are read once and unew is written once (24 B per dof; apply +
 * stefan_ipdg_euler_update move 48).  b may be NULL.  u != unew.  u_lo / u_hi as in stefan_ipdg_apply;
 * sync != NULL selects the peer-halo form of stefan_ipdg_apply_peer (u_lo / u_hi then point into the
 * neighbours' u in peer memory; the kernel publishes sync->my_ready = epoch when it starts -- which also tells
 * the neighbours that this rank is done reading THEIR previous vector -- and its edge tiles wait for the
 * neighbours' flag of the same epoch before they read the halo columns and before they overwrite the edge
 * columns of unew, so two buffers alternating between u and unew need no other ordering). */
int stefan_ipdg_step(stefan_ipdg* h, const double* u, double* unew, const double* b, double dt, const double* u_lo,
                     const double* u_hi, const stefan_peer_sync* sync, void* stream);
/* stefan_ipdg_interface_sums with the neighbours' edge columns in peer memory, ordered inside the one launch: the
 * threads that read a halo column first wait until the flag sync->lo_ready / sync->hi_ready points at has reached
 * sync->epoch -- point them at DONE-type flags (the neighbour's kernel that wrote its vector is complete). */
int stefan_ipdg_interface_sums_peer(stefan_ipdg* h, const double* u, const double* u_lo, const double* u_hi,
                                    const stefan_peer_sync* sync, double* sums_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STEFAN_B200_H */
