/* knpemi.h - C ABI of libknpemi.so, the B200 (sm_100a) implementation of the
 * per-time-step hot path of adajel/KNP-EMI-DG.
 *
 * The reference has no C/FFI plug-in interface: its hot path is Python calling
 * dolfin `assemble`, PETSc `KSP.solve` and numbalsoda `lsoda`.
 * Each entry point below therefore cites the reference *call site* it
 * replaces (paths relative to the reference checkout).  The Python package
 * `knpemidg` (knp-emi-dg_b200/knpemidg) binds these with ctypes and mirrors
 * the reference's Solver / MembraneModel API on top of them.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error;
 *     knp_last_error() returns the message of the last failing call.
 *   - all pointers are HOST pointers (plain C arrays) unless the name ends in
 *     `_dev`; the library owns all device memory.
 *   - a context is bound to one CUDA device and is not thread-safe.
 *   - d = geometric dimension (2|3), nd = d+1 DG-P1 dofs per cell per field,
 *     global dof of a scalar field = nd*cell + local_vertex.
 *   - N = number of ion species, the last one is eliminated
 *     (src/knpemidg/solver.py:69).
 */
#ifndef KNPEMI_H
#define KNPEMI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct knp_ctx knp_ctx;

/* ---- library / context ------------------------------------------------- */
const char* knp_last_error(void);
int knp_version(void);
/* 1 when built for the GPU (the product), 0 for the host-emulation build that
 * exists only for the CPU test-suite (tests/emu). */
int knp_is_cuda_build(void);
int knp_ctx_create(int device, knp_ctx** out);
int knp_ctx_destroy(knp_ctx* ctx);
int knp_sync(knp_ctx* ctx);

/* ---- mesh + tags: replaces Solver.setup_domain (solver.py:85-121), the
 * dolfin facet<->cell connectivity, interface_normal (utils.py:61-85) and the
 * ODE-point selection dlt_dof_extraction.get_indices (:18-48).
 *   coords[nv*d], cell_verts[nc*nd], cell_region[nc] (dense rank of the cell
 *   tag; order preserving, so "lower tag = ECS side" is kept),
 *   facet_cells[nf*2] (second = -1 on the boundary), facet_tag[nf],
 *   mem_tags[n_mem_tags] = facet tags that carry a membrane (ODE model tags,
 *   or the MMS interface tags).  Membrane rows are the interior facets with
 *   such a tag in ascending facet index.
 * Also computes, once, everything about the mesh that the per-step kernels would otherwise re-derive: P1 gradients,
 * volumes, diameters and the per-cell-side table of facet area, 1/avg(h), unit normal and the neighbour's normal
 * derivatives (what UFL's FacetNormal / CellDiameter / avg() deliver at every assemble() of solver.py:477-479, 730-731). */
int knp_mesh_set(knp_ctx* ctx, int d, int64_t nc, int64_t nv, const double* coords,
                 const int32_t* cell_verts, const int32_t* cell_region,
                 int64_t nf, const int32_t* facet_cells, const int32_t* facet_tag,
                 int n_mem_tags, const int32_t* mem_tags);
/* info[0]=d, [1]=nc, [2]=n (=nd*nc), [3]=nm, [4]=nnz of the scalar CSR export,
 * [5]=number of SIP (tag-0 interior) facets, [6]=block slots per cell (nd+1) */
int knp_mesh_info(knp_ctx* ctx, int64_t info[8]);
/* membrane rows: facet index, ICS-side cell ('minus' of n_g), ECS-side cell
 * ('plus'), facet tag; each array has nm entries (any may be NULL). */
int knp_membrane_table(knp_ctx* ctx, int32_t* facet, int32_t* cell_i, int32_t* cell_e,
                       int32_t* tag);

/* ---- parameters: replaces Solver.setup_parameters (solver.py:124-154) and
 * the constants captured by the forms (:275-278, 538-541).
 *   z[N]; D[N*ntags] (make_global, :1244-1258); rho[ntags];
 *   C_sub[(N-1)*ntags] only for the manufactured-solution mode (mms != 0);
 *   splitting: 1 = solve_system_active (:1042), 0 = passive (:958). */
int knp_params_set(knp_ctx* ctx, double F, double R, double T, double C_M, double C_phi,
                   double dt, double tau_emi, double tau_knp, double Lp, int N, const double* z,
                   int ntags, const double* D, const double* rho, const double* C_sub,
                   int splitting, int mms);

/* ---- fields (dolfin Functions in the reference) ------------------------- */
enum {
  KNP_F_C = 0,      /* idx 0..N-1: c_prev_k of the solved ions, N-1 = eliminated ion  [n]   */
  KNP_F_CN = 1,     /* idx 0..N-2: c_prev_n (solver.py:597)                           [n]   */
  KNP_F_PHI = 2,    /* potential                                                      [n]   */
  KNP_F_PHIM = 3,   /* phi_M_prev_PDE on membrane rows (solver.py:211-214)            [nm]  */
  KNP_F_ICH = 4,    /* idx 0..N-1: I_ch_k on membrane rows (solver.py:251-259)        [nm]  */
  KNP_F_NERNST = 5, /* idx 0..N-1: ion['E'] on membrane rows (solver.py:299, 827)     [nm]  */
  KNP_F_RHS_EMI = 6,/* assembled L_emi                                                [n]   */
  KNP_F_RHS_KNP = 7,/* idx 0..N-2: assembled L_knp block                              [n]   */
  KNP_F_LOAD_EMI = 8,/* extra load vector added to L_emi (MMS terms :365-374)         [n]   */
  KNP_F_LOAD_KNP = 9 /* idx 0..N-2: extra load added to L_knp (f_source :599, MMS)    [n]   */
};
int knp_field_set(knp_ctx* ctx, int which, int idx, const double* src, int64_t count);
int knp_field_get(knp_ctx* ctx, int which, int idx, double* dst, int64_t count);

/* ---- assembly: replaces assemble(a_emi), assemble(L_emi), assemble(B_emi)
 * (solver.py:452-453, 477-479) and assemble(A_knp), assemble(L_knp)
 * (:710, 730-731); forms at :289-395 and :550-657. */
int knp_assemble_emi(knp_ctx* ctx);
int knp_assemble_knp(knp_ctx* ctx);
/* scalar CSR copy for the parity harness; which: 0 = A_emi, 1 = B_emi,
 * 2+k = A_knp of solved ion k.  rowptr[n+1], col[nnz], val[nnz]. */
int knp_matrix_export(knp_ctx* ctx, int which, int64_t* rowptr, int32_t* col, double* val);
/* y = M x with the device SpMV kernel (test hook) */
int knp_spmv(knp_ctx* ctx, int which, const double* x, double* y);

/* ---- linear solves: replaces setup_solver_emi/solve_emi KSP (CG + AMG(B),
 * solver.py:425-444, 502-529) and setup_solver_knp/solve_knp (GMRES(30) +
 * AMG(A), :684-701, 767-789).  pc: 0 = element block-Jacobi, 1 = AMG. */
int knp_amg_setup(knp_ctx* ctx, double theta, int max_levels, int coarse_size);
int knp_amg_info(knp_ctx* ctx, int64_t* nlevels, int64_t* rows, int64_t* nnz, int cap);
int knp_solver_options(knp_ctx* ctx, int pc, int nu_pre, int nu_post, int gamma,
                       double omega, int gmres_restart, int knp_min_it);
int knp_solve_emi(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid);
int knp_solve_knp(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid);

/* ---- post-step updates: replaces solver.py:809-842 (c_prev <- c, phi_M
 * facet mean of phi_i - phi_e, Nernst potentials, eliminated ion) and the
 * pcws_constant_project calls (utils.py:100-124).  KNP_POST_ALL is the update at the end of a regular
 * time step and includes c_prev_n <- c (solver.py:810): a c_prev_n written separately through
 * KNP_F_CN is dropped by it. */
enum { KNP_POST_ELIMINATED = 1, KNP_POST_PHIM = 2, KNP_POST_NERNST = 4, KNP_POST_ALL = 7 };
int knp_post_step(knp_ctx* ctx, int what);
/* facet mean of the one-sided trace of a field on the membrane rows;
 * side 0 = plus (ECS, utils.py:87), 1 = minus (ICS, utils.py:94). out[nm]. */
int knp_facet_trace(knp_ctx* ctx, int which, int idx, int side, double* out);

/* ---- membrane ODEs: replaces MembraneModel (membrane.py:7-184) and its
 * per-facet numbalsoda call (membrane.py:84-119).
 *   rows[nrows]: membrane-row index of each ODE point; states[nrows*ns],
 *   params[nrows*np] row-major as in the reference (membrane.py:29-41). */
int knp_model_count(void);
const char* knp_model_name(int model_id);
int knp_model_dims(int model_id, int* ns, int* np);
int knp_membrane_register(knp_ctx* ctx, int model_id, int64_t nrows, const int32_t* rows,
                          const double* states, const double* params, int* handle);
int knp_membrane_states_get(knp_ctx* ctx, int handle, double* states);
int knp_membrane_states_set(knp_ctx* ctx, int handle, const double* states);
int knp_membrane_params_get(knp_ctx* ctx, int handle, double* params);
int knp_membrane_params_set(knp_ctx* ctx, int handle, const double* params);
/* PDE->ODE links executed at the start of every knp_ode_step
 * (solver.py:1094-1101): parameter column `col` <- source.
 *   kind 0: membrane-row field (which, idx) e.g. KNP_F_NERNST
 *   kind 1: facet mean of the `side` trace of cell field (which, idx)
 *           (the update_ode hook, e.g. examples/idealized-geometries/run_2D.py:38-50) */
int knp_membrane_link(knp_ctx* ctx, int handle, int col, int kind, int which, int idx, int side);
/* ODE->PDE: after the step phi_M <- states[:,v_col] (solver.py:1108) and
 * I_ch[ion] <- params[:,col] (solver.py:1111-1113). */
int knp_membrane_outputs(knp_ctx* ctx, int handle, int v_col, int n_ion, const int32_t* ich_cols);
/* stimulus (membrane.py:92, 102-104): rows with mask!=0 get params[:,col]=value
 * at every step. */
int knp_membrane_stimulus(knp_ctx* ctx, int handle, const uint8_t* mask, int ncols,
                          const int32_t* cols, const double* values);
/* one step_lsoda: gather (set_v: also states[:,V] <- phi_M, solver.py:1094),
 * integrate t0 -> t0+dt with relative tolerance rtol, scatter.  Explicit Dormand-Prince 5(4); facets
 * whose step turns out stability-limited (stiff models: LSODA would switch to BDF, membrane.py:108-112)
 * finish the interval with an L-stable Rosenbrock pair. */
int knp_ode_step(knp_ctx* ctx, int handle, double t0, double dt, double rtol, double atol,
                 int set_v, int64_t* stats /* [3]: max steps, total rhs evals, facets on the stiff path; may be NULL */);

/* ---- multi-GPU: cell-partitioned mesh, one context (= one process, one GPU) per part.
 * Replaces what the reference gets from dolfin's distributed mesh + PETSc's MPI
 * matrices/vectors when run_*.py is started under mpirun (ghosted DG facet
 * integrals, src/knpemidg/solver.py:16 `parameters['ghost_mode'] = 'shared_vertex'`, MPI bbox reductions :387-388;
 * VecScatter inside MatMult and MPI_Allreduce inside VecDot/VecNorm of KSPSolve,
 * solver.py:509, 771).
 *
 * The mesh passed to knp_mesh_set is then the LOCAL mesh: the nc_owned cells of this
 * part first, followed by its ghost cells (the face neighbours owned by other parts),
 * grouped by owning rank in the order of neigh_rank[]; only facets with at least one
 * owned cell are listed.  Rows are assembled and solved for owned cells only; ghost
 * values of every field/vector are refreshed by halo exchanges.
 *   send_ptr[nneigh+1], send_cells[]: owned cells whose dofs go to neighbour i,
 *     in the order in which that rank numbers them as ghosts;
 *   recv_ptr[nneigh+1]: ghost cells received from neighbour i are the local cells
 *     nc_owned + recv_ptr[i] .. nc_owned + recv_ptr[i+1].
 * Call after knp_mesh_set and before knp_params_set. */
int knp_dist_set(knp_ctx* ctx, int rank, int world, int64_t nc_owned, int nneigh,
                 const int32_t* neigh_rank, const int64_t* send_ptr, const int32_t* send_cells,
                 const int64_t* recv_ptr);
/* transport 1 (product): NCCL send/recv + allreduce over NVLink on the context's
 * stream.  Rank 0 obtains an id with knp_nccl_unique_id, the host program hands the
 * 128 bytes to the other ranks (torch.distributed broadcast in the Python layer),
 * every rank calls knp_dist_init_nccl (collective).  The call also tries to map every
 * rank's exchange arena with CUDA IPC; where that works the halo exchanges and the small
 * allreduces run as the library's own peer-memory kernels (one launch each, NVLink
 * stores + flags) and NCCL remains for setup and large reductions.  KNP_P2P=0 in the
 * environment keeps everything on NCCL. */
int knp_nccl_unique_id(char out[128]);
int knp_dist_init_nccl(knp_ctx* ctx, const char uid[128]);
/* transport 2 (host-emulation build only, for the CPU test-suite under gloo): the
 * exchanges are delegated to host callbacks.
 *   exchange: send[send_off[i]..send_off[i+1]) goes to ranks[i], the message from
 *     ranks[i] lands in recv[recv_off[i]..recv_off[i+1]);
 *   allreduce: in-place sum of n doubles over all ranks (same bits on every rank). */
typedef int (*knp_exchange_fn)(void* user, int nneigh, const int32_t* ranks, const double* send,
                               const int64_t* send_off, double* recv, const int64_t* recv_off);
typedef int (*knp_allreduce_fn)(void* user, double* buf, int64_t n);
int knp_dist_set_callbacks(knp_ctx* ctx, knp_exchange_fn exchange, knp_allreduce_fn allreduce,
                           void* user);
/* info[0]=rank, [1]=world, [2]=owned cells, [3]=ghost cells, [4]=neighbours,
 * [5]=halo exchanges issued so far, [6]=allreduces issued so far, [7]=how many of those
 * were served by the library's own peer-memory (NVLink, CUDA IPC) kernels instead of NCCL */
int knp_dist_info(knp_ctx* ctx, int64_t info[8]);
/* refresh the ghost values of a cell field from their owners (collective) */
int knp_field_halo(knp_ctx* ctx, int which, int idx);

/* ---- timers (solver.py:77-81): seconds accumulated since the last reset in
 * out[0..5] = emi assembly, emi solve, knp assembly, knp solve, ode, post. */
int knp_timers_get(knp_ctx* ctx, double* out, int reset);

/* ---- measurement hooks (bench.py) ----------------------------------------
 * knp_launch_count: kernels launched by this library since it was loaded.
 * knp_timer_start/stop: CUDA events recorded on the context's stream (the
 *   library does not launch on torch's current stream, so torch.cuda.Event
 *   cannot see its kernels); stop returns the elapsed milliseconds.
 * knp_bench_kernel: `reps` back-to-back launches of one hot kernel on the
 *   current state, timed with CUDA events; *ms = average per launch and
 *   *bytes = algorithmic bytes one launch moves.
 *   kernel ids: 0 block-ELL SpMV (A_emi), 1 EMI assembly (pre-pass + cells),
 *   2 KNP assembly (all solved ions), 3 block-Jacobi sweep (B_emi);
 *   multi-GPU exchanges (collective; 0 ms on a single part): 4 DG halo of one field,
 *   5 allreduce of 4 Krylov scalars, 6 all-gather of the replicated AMG level's rhs. */
long long knp_launch_count(void);
int knp_timer_start(knp_ctx* ctx);
int knp_timer_stop(knp_ctx* ctx, double* ms);
int knp_bench_kernel(knp_ctx* ctx, int kernel, int reps, double* ms, double* bytes);

#ifdef __cplusplus
}
#endif
#endif
