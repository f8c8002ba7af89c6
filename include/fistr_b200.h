/*
 * fistr_b200 — C-ABI of the B200-native linear-static hot path for FrontISTR / HEC-MW.
 *
 * Every entry point is what a Fortran `interface ... bind(C)` (ISO_C_BINDING) stub binds:
 * plain pointers and sizes, scalars by value, arrays exactly as hecmwST_matrix /
 * hecmwST_local_mesh hold them (1-based node ids, 0:NP prefix arrays, row-major 3x3 blocks).
 * Host pointers are borrowed for the duration of the call only.  One process drives one GPU;
 * like the reference's preconditioner modules (module-level SAVE state) the library keeps ONE
 * live matrix / factorisation per process.
 *
 * There is NO CPU fallback: every compute entry point returns FISTR_B200_ERR_CUDA (-2) if no
 * CUDA device is usable.  Return value convention: 0 success, >0 HEC-MW solver code
 * (hecmw1/src/solver/init/hecmw_solve_error.f90:9-15), <0 library error
 * (fistr_b200_last_error() gives the text).
 *
 * Citations are file:line in the reference tree (ricosjp/FrontISTR).
 */
#ifndef FISTR_B200_H
#define FISTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FISTR_B200_OK 0
#define FISTR_B200_ERR_ARG (-1)
#define FISTR_B200_ERR_CUDA (-2)
#define FISTR_B200_ERR_COMM (-3)
#define FISTR_B200_ERR_STATE (-4)
#define FISTR_B200_ERR_CONNECTIVITY (-5) /* "cannot find connectivity", hecmw_mat_ass.f90:86-90 */

/* ---- life cycle ------------------------------------------------------------------------- */
/* replaces nothing in the reference (there is no accelerator boundary today); called once after
 * hecmw_init (hecmw1/src/common/hecmw_util_f.F90:478-520). */
int fistr_b200_init(int device_ordinal);
int fistr_b200_finalize(void);
const char *fistr_b200_last_error(void);
/* Integer knobs and read-outs (unknown keys read as -1).
 *   set: "spmv_variant" 0|1|2 (2), "sell_sigma" rows of the SELL-C-sigma sorting window (0 = off, -1 = 1024 if the natural order pads > 10 %), "spmv_blocks_per_sm" (8),
 *        "cg_fused" 0|1 (1), "ssor_fused" 0|1 SSOR x 1 sweep in 2 ncolor + 3 launches per iteration (0), "cg_check_interval" iterations per launch batch, "time_ax" 0|1, "asm_gather" 0|1 (1),
 *        "asm_probe" development switches of the assembly kernel (0), "ic_inverse" 0 | 1 | 2 | -2 pick by self-test + timing (0),
 *        "ic_pick_min_elems" IC block size that starts the kernel selection (65536), "p2p" 0 NCCL | 1 peer memory if
 *        every rank can map it | 2 peer memory or fail (1), "p2p_loopback" 0|1 one rank: the peer-memory kernels against the
 *        rank's own arena (their first run on a single GPU; set before mat_structure) (0), "comm_zhalo" 0|1 NCCL transport: z halo + (r.r, r.z) partials in one
 *        grouped send/recv, 2 communication steps per CG iteration instead of 3; 2 = also on one rank (0), "matcon_device" 0|1 (0), "bilu_chain_rows" (256),
 *        "launch_count_reset"
 *   get: "ic_inverse_active", "ic_selftest_ppb", "ic_selftest_ppb_ldlt", "ic_time_us_smem|reg|ldlt", "ic_time_elems",
 *        "bilu_levels", "bilu_launches_per_apply", "update_kernel_us", "nodal_stress_kernel_us", "p2p_active",
 *        "comm_steps" NCCL steps (grouped send/recv or all-reduce) issued since init */
int fistr_b200_set_option(const char *key, int64_t value);
int64_t fistr_b200_get_option(const char *key);

/* ---- communication: replaces the MPI layer under hecmw_update_3_R / hecmw_allreduce_R1 ------
 * (hecmw1/src/solver/communication/hecmw_comm_f.F90:346-393,669-694; hecmw_solver_SR_33.F90:42-124).
 * Rank 0 obtains an id (NCCL unique id, 128 bytes) and broadcasts it with the host's own
 * transport (MPI_Bcast in fistr1, torch.distributed in the Python harness). */
int fistr_b200_comm_get_unique_id(void *id128);
int fistr_b200_comm_init(const void *id128, int my_rank, int petot);
int fistr_b200_comm_finalize(void);

/* ---- hecmw_mat_con (hecmw1/src/solver/matrix/hecmw_mat_con.f90:23-268) -----------------------
 * count builds the node graph and keeps it; fill copies it into caller-allocated arrays
 * indexL/indexU(0:NP), itemL(NPL), itemU(NPU) — sorted ascending per row, 1-based.
 * Built on the host by default (usable without a GPU); option "matcon_device" = 1 builds it on the
 * device (sort of node-pair keys), same arrays. */
int fistr_b200_mat_con_count(int n_node, int nn_internal, int n_elem_type, const int *elem_type_index,
                             const int *elem_type_item, const int *elem_node_index,
                             const int *elem_node_item, int *N, int *NP, int *NPL, int *NPU);
int fistr_b200_mat_con_fill(int *indexL, int *itemL, int *indexU, int *itemU);

/* ---- matrix residency ---------------------------------------------------------------------
 * Declares the structure of THE matrix (hecmwST_matrix, hecmw_util_f.F90:433-468) and the halo
 * tables of the subdomain (hecmwST_local_mesh comm fields, :304-312).  Called implicitly by
 * solve/matvec/stiffmatrix when the structure pointers or sizes change or Iarray(98)=1. */
int fistr_b200_mat_structure(int N, int NP, int NPL, int NPU, const int *indexL, const int *itemL,
                             const int *indexU, const int *itemU, int n_neighbor_pe,
                             const int *neighbor_pe, const int *import_index, const int *import_item,
                             const int *export_index, const int *export_item);
/* values/RHS between host arrays and the device-resident copy (any pointer may be NULL = skip) */
int fistr_b200_mat_upload(const double *D, const double *AL, const double *AU, const double *B);
int fistr_b200_mat_download(double *D, double *AL, double *AU, double *B, double *X);
/* Page-lock / release a host array the caller will pass repeatedly (hecMAT%D/AL/AU/B/X after
 * hecMAT_init, fistr1/src/lib/m_fstr.f90:807-857) so the borrowed-pointer copies run at full PCIe
 * rate.  Optional: unregistered (pageable) arrays work, only slower. */
int fistr_b200_host_register(void *ptr, int64_t bytes);
int fistr_b200_host_unregister(void *ptr);

/* ---- fstr_StiffMatrix + hecmw_mat_ass_elem ---------------------------------------------------
 * (fistr1/src/analysis/static/fstr_StiffMatrix.f90:18-212; STF_C3D8IC static_LIB_3dIC.f90:21-215;
 *  STF_C3 static_LIB_3d.f90:47-205; STF_C3D8Bbar static_LIB_C3D8.f90:23-198;
 *  hecmw_mat_ass_elem hecmw_mat_ass.f90:31-134).  Element types 361 (elemopt361 = 1 FI, 2 B-bar,
 *  3 IC) and 342, linear elastic (nlgeom_flag == INFINITE).  The matrix structure must have been
 *  declared.  D/AL/AU: host destinations in the hecmwST_matrix layout, or NULL to leave the
 *  assembled matrix on the device only.  Zeroes the matrix first (hecmw_mat_clear). */
int fistr_b200_stiffmatrix_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                              const int *elem_type_index, const int *elem_type_item,
                              const int *elem_node_index, const int *elem_node_item,
                              const int *section_ID, int n_sect, const int *sect_elemopt361,
                              const double *sect_E, const double *sect_nu, double *D, double *AL,
                              double *AU, double *time_ms /* [2]: integrate+scatter kernel, total */);

/* ---- geometrically nonlinear static analysis (`!STATIC, TYPE=NLGEOM` / `!SOLUTION, TYPE=NLSTATIC`), total
 * Lagrangian, SURVEY 8(f) rank 2.  The Newton loop itself (fistr1/src/analysis/static/fstr_solve_NonLinear.f90:
 * 29-167) stays in the caller; per Newton iteration it calls
 *   nl_stiffmatrix_3d  = fstr_StiffMatrix with nlgeom_flag == TOTALLAG (fstr_StiffMatrix.f90:60-150; STF_C3
 *                        static_LIB_3d.f90:47-205, STF_C3D8Bbar static_LIB_C3D8.f90:23-198): tangent stiffness at
 *                        u = unode + dunode incl. the initial-stress matrix of the stored Gauss-point stress,
 *   mat_ass_bc, solve_33 (above), and
 *   nl_update_3d       = fstr_UpdateNewton (fstr_Update.f90:25-293; UPDATE_C3 static_LIB_3d.f90:516-837,
 *                        Update_C3D8Bbar static_LIB_C3D8.f90:203-552): Green-Lagrange strain and 2nd Piola-Kirchhoff
 *                        stress of every Gauss point (kept on the device, [element][8][6]) and QFORCE(3*n_node).
 * Element types 361 (elemopt361 = 1 FI, 2 B-bar; 3 IC is rejected as in the reference) and 342.  Materials per
 * section: sect_mtype 0 ELASTIC (sect_E, sect_nu), 1 MOONEYRIVLIN (sect_hyper[3*s..] = C10, C01, D;
 * lib/physics/Hyperelastic.f90:221-286).  sect_mtype / sect_hyper may be NULL (all elastic). */
int fistr_b200_nl_state_reset(int n_elem);                          /* zero strain / stress (start of the analysis) */
int fistr_b200_nl_state_download(double *strain, double *stress);   /* [n_elem][8][6] each, either may be NULL */
int fistr_b200_nl_stiffmatrix_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                                 const int *elem_type_index, const int *elem_type_item,
                                 const int *elem_node_index, const int *elem_node_item, const int *section_ID,
                                 int n_sect, const int *sect_elemopt361, const double *sect_E, const double *sect_nu,
                                 const int *sect_mtype, const double *sect_hyper, const double *u, double *D,
                                 double *AL, double *AU, double *time_ms);
int fistr_b200_nl_update_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                            const int *elem_type_index, const int *elem_type_item, const int *elem_node_index,
                            const int *elem_node_item, const int *section_ID, int n_sect,
                            const int *sect_elemopt361, const double *sect_E, const double *sect_nu,
                            const int *sect_mtype, const double *sect_hyper, const double *unode,
                            const double *dunode, double *QFORCE);

/* ---- strain / stress recovery of the linear static analysis, SURVEY 8(f) rank 2 -----------------------------------
 * update_3d       = fstr_UpdateNewton with nlgeom_flag == INFINITE (fistr1/src/analysis/static/fstr_Update.f90:25-293):
 *                   UpdateST_C3D8IC (lib/static_LIB_3dIC.f90:220-455) for 361 elemopt361 = 3, UPDATE_C3
 *                   (lib/static_LIB_3d.f90:516-837) for 361 FI and 342, Update_C3D8Bbar (lib/static_LIB_C3D8.f90:203-552):
 *                   strain and stress of every Gauss point at unode + dunode (kept on the device, [element][8][6], the
 *                   same state the nl_* entry points use) and QFORCE(3*n_node), summed per node in ascending element
 *                   order.  sect_alpha / temperature / temperature0 (all NULL: no !TEMPERATURE): isotropic thermal strain
 *                   alpha (T - ref_temp) - alpha (T0 - ref_temp).  QFORCE may be NULL.
 * nodal_stress_3d = fstr_NodalStress3D (fistr1/src/analysis/static/fstr_NodalStress.f90:13-272) on the mesh and
 *                   Gauss-point state of the last update_3d / nl_update_3d call: NodalStress_INV3 extrapolation
 *                   (:337-400), nodal averaging, ElementStress_C3, von Mises.  STRAIN / STRESS (6*n_node),
 *                   MISES (n_node), ESTRAIN / ESTRESS (6*n_elem), EMISES (n_elem); any pointer may be NULL. */
int fistr_b200_update_3d(int n_node, const double *node, int n_elem, int n_elem_type, const int *elem_type_index,
                         const int *elem_type_item, const int *elem_node_index, const int *elem_node_item,
                         const int *section_ID, int n_sect, const int *sect_elemopt361, const double *sect_E,
                         const double *sect_nu, const double *sect_alpha, const double *unode, const double *dunode,
                         const double *temperature, const double *temperature0, double ref_temp, double *QFORCE);
int fistr_b200_nodal_stress_3d(double *STRAIN, double *STRESS, double *MISES, double *ESTRAIN, double *ESTRESS,
                               double *EMISES);

/* ---- hecmw_mat_ass_bc (hecmw_mat_ass.f90:292-429), batched over the constrained DOFs that
 * fstr_AddBC (fistr1/src/analysis/static/fstr_AddBC.f90:54-109) visits.  Operates on the
 * device-resident matrix and RHS. */
int fistr_b200_mat_ass_bc(int nbc, const int *inode, const int *idof, const double *rhs);

/* ---- hecmw_solve (hecmw1/src/solver/hecmw_solver.f90:9-78 -> hecmw_solve_iterative
 * iterative/hecmw_solver_Iterative.f90:13-210 -> hecmw_solve_CG iterative/hecmw_solver_CG.f90:19-312).
 * Reads ITER/METHOD/METHOD2/PRECOND/iterPREmax/SCALING/ITERLOG/TIMELOG/USEJAD/NCOLOR_IN and flags 96/97/98 from Iarray,
 * RESID/SIGMA_DIAG from Rarray (slot map hecmw_matrix_misc.f90:89-121); writes X(1:3*NP) including
 * halo, Iarray(81/82/96/97/98).  D/AL/AU/B == NULL means "use the device-resident copy".
 *   METHOD  : 1 (CG) only; anything else returns 1001 (HECMW_SOLVER_ERROR_INCONS_PC, the reference's code for an
 *             undefined METHOD, Iterative.f90:136-138).
 *   PRECOND : 1, 2 SSOR | 3 DIAG | 10 BILU(0); others 1001.
 *   SIGMA_DIAG < 0 ("auto", Iterative.f90:72-77): starts at 1.0; after a divergence (3002 / 3003) with PRECOND 10..19 the
 *             reference's loop adds 0.1 and calls CG again while SIGMA_DIAG < 2, each call starting from the X the last
 *             one left (Iterative.f90:141-156) - reproduced as the reference behaves, incl. the fact that
 *             hecmw_precond_BILU_33_setup keeps the first factorisation (BILU_33.f90:49-57).  get_option
 *             "auto_sigma_retries" = calls repeated by the last solve.
 *   METHOD2 : > 1 asks for a switch to BiCGSTAB / GMRES / GPBiCG after CG diverged (Iterative.f90:150-153): those solvers
 *             are outside this path, the call then returns 1001.  Not looked at unless CG diverges.
 *   SCALING : Iarray(7) != 0: hecmw_solver_scaling_fw_33 / _bk_33 (las/hecmw_solver_scaling_33.f90) around the CG loop,
 *             in place on the device-resident matrix / B / X, scale vector halo-exchanged; host D/AL/AU given by the
 *             caller are not written back (the reference scales and un-scales hecMAT%D/AL/AU in place, which leaves
 *             them within an ulp or two of their values).
 *   USEJAD  : Iarray(33) != 0 returns 1001: the library always multiplies from its own private SELL copy.
 * Collective over all ranks of the communicator.  time_out[5] = set-up, solver, comm, matvec seconds
 * (the TIMELOG summary, Iterative.f90:192-208; comm / matvec / precond = CUDA-event times of the sampled iterations x
 * iterations, measured when TIMELOG or the "time_ax" option asks; TIMELOG = 2 (VERBOSE) prints max / min / avg / std dev
 * of set-up and solver time over the ranks and reports the maxima, CG.f90:132-145,296-310) and
 * [4] = CUDA-event time of the whole call on the library stream, H2D/D2H copies included. */
int fistr_b200_solve_33(int N, int NP, int NPL, int NPU, const double *D, const double *AL,
                        const double *AU, const int *indexL, const int *itemL, const int *indexU,
                        const int *itemU, const double *B, double *X, int *Iarray, double *Rarray,
                        int n_neighbor_pe, const int *neighbor_pe, const int *import_index,
                        const int *import_item, const int *export_index, const int *export_item,
                        int my_rank, int petot, int *iter_out, double *resid_out, double *time_out);
/* hecmw_mat_dump + hecmw_mat_dump_rhs (what = 1) / hecmw_mat_dump_solution (what = 2) of the DEVICE-RESIDENT system
 * (hecmw1/src/solver/matrix/hecmw_matrix_dump.f90:30-372): text files dump_matrix_<call>_<rank>.{mm|csr|bsr, rhs, sol} in
 * dir (NULL or "" = working directory), all NP rows, values in Fortran's E20.12E3 format one per line.  dump_type 1
 * MatrixMarket, 2 scalar CSR, 3 block CSR.  numcall <= 0: the library's own count of solve calls (the reference's
 * NumCall).  fistr_b200_solve_33 calls it by itself when Iarray(31) (DUMPTYPE) != 0 - before the CG loop and, for the
 * solution, after the final residual - writing into $FISTR_B200_DUMP_DIR or the working directory; with Iarray(32)
 * (DUMPEXIT) != 0 it then returns FISTR_B200_ERR_STATE without solving (the reference stops the program there). */
int fistr_b200_mat_dump(int dump_type, int what, int numcall, const char *dir);
/* host-only: v in Fortran's E20.12E3 edit descriptor as the dump files carry it; out receives 20 characters + 0 */
int fistr_b200_debug_format_e20(double v, char *out);

/* residual history of the last solve (RESID per iteration, what ITERLOG=YES prints, CG.f90:245) */
int fistr_b200_solve_history(double *hist, int max_len);
/* CUDA-event profile of the last solve (needs the "time_ax" option or TIMELOG): mean seconds per sampled
 * iteration of [p-update incl. halo push, SpMV, tail kernel incl. dot-product reductions, whole iteration] */
int fistr_b200_solve_profile(double *seconds4);
/* final true relative residual of the last solve (hecmw_rel_resid_L2, Iterative.f90:166) */
int fistr_b200_solve_resid2(double *resid2);

/* ---- hecmw_matvec_33 (hecmw1/src/solver/las/hecmw_solver_las_33.f90:50-74,135-351):
 * halo exchange of X then Y(1:3N) = A X.  Uses the device-resident matrix; X is read for 3*NP
 * values (halo part overwritten by the exchange when petot>1), Y receives 3*N values. */
int fistr_b200_matvec_33(const double *X, double *Y);
/* ---- hecmw_matresid (las/hecmw_solver_las.f90:105-122, las_33.f90:358-380) and hecmw_rel_resid_L2 (las.f90:129-158):
 * halo exchange of X, R(1:3N) = B - A X, *rel_resid = sqrt(r.r / b.b) over all ranks (b.b == 0 counts as 1).
 * B == NULL uses the device-resident right-hand side; R and rel_resid may be NULL.  Collective. */
int fistr_b200_matresid_33(const double *X, const double *B, double *R, double *rel_resid);
/* SpMV micro-benchmark on the resident matrix: nrep launches, x_i = sin(i); returns mean ms/launch */
int fistr_b200_matvec_bench(int nrep, int flush_l2, double *ms_mean, double *ms_min);

/* ---- hecmw_precond_33_{setup,apply,clear} (precond/33/hecmw_precond_33.f90:27-115;
 * DIAG (precond = 3) precond/33/hecmw_precond_DIAG_33.f90:27-159; SSOR (1, 2) precond/33/hecmw_precond_SSOR_33.f90:55-467;
 * BILU(0) (10) precond/33/hecmw_precond_BILU_33.f90:35-362 - a zero pivot block, where the reference stops the program
 * (:1509), returns 2001; 11 / 12 (ILU(1), ILU(2)) return 1001) */
int fistr_b200_precond_33_setup(int precond, double sigma_diag, int ncolor_in);
int fistr_b200_precond_33_apply(double *ZP /* 3*NP in/out */);
int fistr_b200_precond_33_clear(void);
int fistr_b200_precond_33_info(int *ncolor, int *perm /* N or NULL */);

/* ---- hecmw_update_3_R (communication/hecmw_comm_f.F90:669-694) on a host vector */
int fistr_b200_update_3_R(double *X /* 3*NP in/out */);

/* ---- hecmw_part1 in-process: node-based RCB (hecmw1/tools/partitioner/hecmw_partition.c:1570-1660),
 * axes[i] in {0,1,2} per bisection level (the 'x,y,z' line of hecmw_part_ctrl.dat); host-only.  At every level each
 * domain hands the ceil(count/2) nodes with the smallest coordinate to the new domain, as the reference does; nodes with
 * EQUAL coordinates are ordered by node index (the reference leaves that to its unstable sort), so the map equals
 * hecmw_part1's wherever the coordinates along the cut axis are distinct and is a valid balanced bisection otherwise. */
int fistr_b200_part_rcb(int n_node, const double *node, int n_div, const int *axes, int *domain);

/* METHOD=PMETIS (method 1, METIS_PartGraphRecursive) / KMETIS (method 2, METIS_PartGraphKway) of hecmw_part1, node-based
 * (hecmw1/tools/partitioner/hecmw_partition.c:2140-2300, 2727-2850; node graph of element edges as built by
 * hecmw_mesh_edge_info.c + create_node_graph :1715-1830; unit weights, default options).  domain[k] in 0..n_domain-1;
 * any n_domain >= 1.  Connectivity as everywhere in this header (FrontISTR-internal node order).  METIS 5 is linked from
 * the CUDA toolkit's libmetis_static.a when the build finds it: fistr_b200_part_metis_available() says whether; without it
 * the call fails with FISTR_B200_ERR_STATE (the reference's partitioner without -DHECMW_PART_WITH_METIS refuses too,
 * hecmw_part_get_control.c:449-460).  Identical to the real hecmw_part1 linked to the same METIS (tests/test_partition.py).
 * host-only */
int fistr_b200_part_metis(int method, int n_domain, int n_node, int n_elem_type, const int *elem_type_index,
                          const int *elem_type_item, const int *elem_node_index, const int *elem_node_item, int *domain,
                          long long *edgecut /* may be NULL */);
int fistr_b200_part_metis_available(void);

/* ... and the distributed mesh of ONE rank from that node->domain map, exactly as hecmw_part1 writes it
 * (hecmw1/src/common/hecmw_io_dist.c:1991-2284: internal nodes first, then external; elements touching an
 * internal node; neighbour / import / export tables).  _count builds and returns
 * sizes[8] = {n_node, nn_internal, n_elem, n_elem_type, len(elem_node_item), n_neighbor_pe, n_import, n_export};
 * _fill copies into caller arrays: global_node (0-based global index per local node), node_ID (2*n_node),
 * elem_type_index (n_elem_type+1), elem_type_item, elem_node_index (n_elem+1), elem_node_item (1-based local ids),
 * global_elem (0-based), neighbor_pe, import_index/export_index (n_neighbor_pe+1), import_item, export_item. */
int fistr_b200_part_local_count(int n_node, int n_elem_type, const int *elem_type_index, const int *elem_type_item,
                                const int *elem_node_index, const int *elem_node_item, const int *domain, int rank,
                                int nparts, int *sizes);
int fistr_b200_part_local_fill(int *global_node, int *node_ID, int *elem_type_index, int *elem_type_item,
                               int *elem_node_index, int *elem_node_item, int *global_elem, int *neighbor_pe,
                               int *import_index, int *import_item, int *export_index, int *export_item);

/* host-only (no GPU): the multicolour ordering (RCM + greedy colouring, hecmw_matrix_ordering_CM.f90 /
 * hecmw_matrix_ordering_MC.f90) the device SSOR uses; perm[new-1]=old (1-based), color_index has ncolor+1 entries */
int fistr_b200_debug_ssor_order(int N, const int *indexL, const int *itemL, const int *indexU, const int *itemU,
                                int ncolor_in, int *perm, int *ncolor, int *color_index);

/* host-only: layout of the private sliced-ELL SpMV copy (row -> slice/lane, slice widths) for `sigma`
 * ("sell_sigma" option; <= 1 natural order) and `variant` (1 or 2).  rowperm[32*nslice], slice_ptr[nslice+1]. */
int fistr_b200_debug_sell_layout(int N, const int *indexL, const int *indexU, int sigma, int variant, int *rowperm,
                                 long long *slice_ptr);

/* ---- instrumentation */
int64_t fistr_b200_launch_count(void);   /* kernels launched by this library since init/reset */
int fistr_b200_device_info(char *name, int name_len, int *sm_count, int64_t *mem_bytes);
/* FP64 FMA micro-benchmark (roofline denominator for element integration); TFLOP/s */
int fistr_b200_fp64_peak(double *tflops);
/* HBM copy micro-benchmark (cross-check of MEASURED_PEAKS.json); GB/s read+write */
int fistr_b200_hbm_copy_peak(double *gbs);

#ifdef __cplusplus
}
#endif
#endif
