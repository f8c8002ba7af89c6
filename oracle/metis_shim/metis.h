/* ORACLE build recipe (test infrastructure).  The reference's partitioner is written against a 32-bit-idx_t METIS 5
 * ("int *xadj", hecmw1/tools/partitioner/hecmw_partition.c:2140-2300) and no METIS ships with it.  The only
 * METIS in this image is the CUDA toolkit's libmetis_static.a, built with 64-bit idx_t.  This header declares the part
 * of the METIS 5 API the partitioner calls with 32-bit idx_t; metis32_adapter.c implements it on top of the archive
 * (whose two entry points are renamed with objcopy in oracle/_ref so that the names do not clash).  Used ONLY to build
 * oracle/_ref/hecmw_part1_metis, the checker of the product's METHOD=KMETIS/PMETIS. */
#ifndef ORACLE_METIS_SHIM_H
#define ORACLE_METIS_SHIM_H
#define METIS_VER_MAJOR 5
#define METIS_VER_MINOR 1
typedef int idx_t;
typedef float real_t;
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt,
                        idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part);
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt,
                             idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part);
#endif
