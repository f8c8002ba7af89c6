/* ORACLE build recipe (test infrastructure): 32-bit-idx_t front of the two METIS 5 entry points hecmw_part1 calls, on top
 * of the 64-bit-idx_t archive of the CUDA toolkit (entry points renamed FB64_* by oracle/build_ref.sh).  The partitioner
 * passes NULL for vsize / adjwgt / tpwgts / ubvec / options; vertex weights (contact modes only) are converted. */
#include <stdint.h>
#include <stdlib.h>

#include "metis.h"

int FB64_METIS_PartGraphKway(int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, void *, void *,
                             int64_t *, int64_t *, int64_t *);
int FB64_METIS_PartGraphRecursive(int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, void *, void *,
                                  int64_t *, int64_t *, int64_t *);

static int64_t *widen(const idx_t *a, size_t n) {
  int64_t *w = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) w[i] = a[i];
  return w;
}

static int call(int kway, idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *nparts, idx_t *objval, idx_t *part) {
  const size_t n = (size_t)*nvtxs;
  int64_t nv = *nvtxs, nc = *ncon, np = *nparts, cut = 0;
  int64_t *x = widen(xadj, n + 1), *a = widen(adjncy, (size_t)xadj[n]), *w = vwgt ? widen(vwgt, n * (size_t)*ncon) : NULL;
  int64_t *p = (int64_t *)calloc(n ? n : 1, sizeof(int64_t));
  const int rc = kway ? FB64_METIS_PartGraphKway(&nv, &nc, x, a, w, NULL, NULL, &np, NULL, NULL, NULL, &cut, p)
                      : FB64_METIS_PartGraphRecursive(&nv, &nc, x, a, w, NULL, NULL, &np, NULL, NULL, NULL, &cut, p);
  for (size_t i = 0; i < n; i++) part[i] = (idx_t)p[i];
  *objval = (idx_t)cut;
  free(x); free(a); free(w); free(p);
  return rc;
}

int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt,
                        idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part) {
  (void)vsize; (void)adjwgt; (void)tpwgts; (void)ubvec; (void)options;
  return call(1, nvtxs, ncon, xadj, adjncy, vwgt, nparts, objval, part);
}
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt,
                             idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part) {
  (void)vsize; (void)adjwgt; (void)tpwgts; (void)ubvec; (void)options;
  return call(0, nvtxs, ncon, xadj, adjncy, vwgt, nparts, objval, part);
}
