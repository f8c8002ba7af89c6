// METHOD=PMETIS / KMETIS of hecmw_part1, node-based (hecmw1/tools/partitioner/hecmw_partition.c: metis_partition_nb
// :2923-2950 -> metis_partition_nb_contact_dist :2727-2850 with no contact pair, pmetis_/kmetis_interface :2140-2300):
// the NODE GRAPH of the mesh's element edges is handed to METIS 5 with unit weights and default options, and the part of
// every vertex is the domain of the node.  north_star: "the repo's own METIS partitioner".
//
// The graph (what decides the result) follows the reference exactly:
//   * edges per element from the per-type tables of tools/partitioner/hecmw_mesh_edge_info.c, elements in mesh order,
//     every undirected edge numbered at its FIRST appearance (HECMW_mesh_hsort_edge, hecmw_mesh_hash_sort.c:655-702);
//   * the adjacency list of a node holds its neighbours in edge order (create_node_graph_link_list, :1715-1758),
//     0-based in the arrays given to METIS (create_node_graph_compress, :1761-1778).
// hecmw_part1 sees second-order connectivity in HEC-MW FILE order while this library holds it in FrontISTR's internal
// order (fistr1/src/common/hecmw2fstr_connect_conv.c:17-19): the tables below are written in file positions and mapped.
//
// METIS itself is not part of the reference tree; the library links the METIS 5 archive that ships with the
// CUDA toolkit (libmetis_static.a, 64-bit idx_t) when the build finds it (FB_WITH_METIS).  Without it the entry returns
// an error, as hecmw_part1 built without -DHECMW_PART_WITH_METIS does (hecmw_part_get_control.c:449-460).
#pragma once
#include <cstdint>
#include <vector>

#include "matcon.hpp"

#ifdef FB_WITH_METIS
extern "C" {
int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                        int64_t *adjwgt, int64_t *nparts, void *tpwgts, void *ubvec, int64_t *options, int64_t *objval,
                        int64_t *part);
int METIS_PartGraphRecursive(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                             int64_t *adjwgt, int64_t *nparts, void *tpwgts, void *ubvec, int64_t *options, int64_t *objval,
                             int64_t *part);
}
#endif

namespace fb {

struct PartEdgeTable {
  int n_edge;
  const signed char (*edge)[2];     // node positions in HEC-MW file order
  const signed char *file_to_int;   // file position -> position in the library's (FrontISTR-internal) connectivity, or null
};

inline const PartEdgeTable *part_edge_table(int etype) {
  static const signed char tet1[][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  static const signed char tet2[][2] = {{0, 4}, {4, 1}, {1, 5}, {5, 2}, {2, 6}, {6, 0}, {0, 7}, {7, 3}, {1, 8}, {8, 3}, {2, 9}, {9, 3}};
  static const signed char pri1[][2] = {{0, 1}, {1, 2}, {2, 0}, {3, 4}, {4, 5}, {5, 3}, {0, 3}, {1, 4}, {2, 5}};
  static const signed char pri2[][2] = {{0, 8}, {8, 1}, {1, 6}, {6, 2}, {2, 7}, {7, 0}, {3, 11}, {11, 4}, {4, 9}, {9, 5}, {5, 10}, {10, 3},
                                        {0, 12}, {12, 3}, {1, 13}, {13, 4}, {2, 14}, {14, 5}};
  static const signed char hex1[][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  static const signed char hex2[][2] = {{0, 8}, {8, 1}, {1, 9}, {9, 2}, {2, 10}, {10, 3}, {3, 11}, {11, 0}, {4, 12}, {12, 5}, {5, 13}, {13, 6},
                                        {6, 14}, {14, 7}, {7, 15}, {15, 4}, {0, 16}, {16, 4}, {1, 17}, {17, 5}, {2, 18}, {18, 6}, {3, 19}, {19, 7}};
  // internal[j] = file[T[j] - 1] (Table342, Table352)  =>  file position p sits at internal position j with T[j] - 1 == p
  static const signed char f2i_342[10] = {0, 1, 2, 3, 5, 6, 4, 7, 8, 9};
  static const signed char f2i_352[15] = {0, 1, 2, 3, 4, 5, 7, 8, 6, 10, 11, 9, 12, 13, 14};
  static const PartEdgeTable T341 = {6, tet1, nullptr}, T342 = {12, tet2, f2i_342}, T351 = {9, pri1, nullptr},
                             T352 = {18, pri2, f2i_352}, T361 = {12, hex1, nullptr}, T362 = {24, hex2, nullptr};
  switch (etype) {
    case 341: return &T341;
    case 342: return &T342;
    case 351: return &T351;
    case 352: return &T352;
    case 361: return &T361;
    case 362: return &T362;
    default: return nullptr;
  }
}

// xadj (n_node + 1) / adjncy of the reference's node graph, 0-based.  -1: unsupported element type, -2: bad node id
inline int part_node_graph(int n_node, int n_elem_type, const int *elem_type_index, const int *elem_type_item,
                           const int *elem_node_index, const int *elem_node_item, std::vector<int64_t> &xadj,
                           std::vector<int64_t> &adjncy) {
  // pass 1: undirected edges in order of first appearance; open-addressing set of (min, max) keys
  size_t cand = 0;
  for (int it = 0; it < n_elem_type; it++) {
    const PartEdgeTable *T = part_edge_table(elem_type_item[it]);
    if (!T) return -1;
    cand += (size_t)(elem_type_index[it + 1] - elem_type_index[it]) * T->n_edge;
  }
  size_t cap = 64;
  while (cap < 2 * cand + 16) cap <<= 1;
  std::vector<uint64_t> table(cap, ~0ull);
  std::vector<int> e1, e2;
  e1.reserve(cand / 3 + 16); e2.reserve(cand / 3 + 16);
  for (int it = 0; it < n_elem_type; it++) {
    const PartEdgeTable *T = part_edge_table(elem_type_item[it]);
    for (int e = elem_type_index[it]; e < elem_type_index[it + 1]; e++) {
      const int *c = elem_node_item + elem_node_index[e];
      for (int k = 0; k < T->n_edge; k++) {
        int pa = T->edge[k][0], pb = T->edge[k][1];
        if (T->file_to_int) { pa = T->file_to_int[pa]; pb = T->file_to_int[pb]; }
        const int a = c[pa], b = c[pb];
        if (a < 1 || a > n_node || b < 1 || b > n_node) return -2;
        const uint64_t key = ((uint64_t)(uint32_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
        size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 17) & (cap - 1);
        while (table[h] != ~0ull && table[h] != key) h = (h + 1) & (cap - 1);
        if (table[h] == key) continue;
        table[h] = key;
        e1.push_back(a); e2.push_back(b);
      }
    }
  }
  // pass 2: adjacency lists in edge order
  xadj.assign((size_t)n_node + 1, 0);
  for (size_t k = 0; k < e1.size(); k++) { xadj[e1[k]]++; xadj[e2[k]]++; }
  for (int i = 0; i < n_node; i++) xadj[i + 1] += xadj[i];
  adjncy.assign((size_t)xadj[n_node], 0);
  std::vector<int64_t> fill(xadj.begin(), xadj.end() - 1);
  for (size_t k = 0; k < e1.size(); k++) {
    adjncy[fill[e1[k] - 1]++] = e2[k] - 1;
    adjncy[fill[e2[k] - 1]++] = e1[k] - 1;
  }
  return 0;
}

// method: 1 = PMETIS (METIS_PartGraphRecursive), 2 = KMETIS (METIS_PartGraphKway).  -3: METIS not linked, -4: METIS error
inline int part_metis(int method, int n_domain, int n_node, int n_elem_type, const int *elem_type_index,
                      const int *elem_type_item, const int *elem_node_index, const int *elem_node_item, int *domain,
                      long long *edgecut) {
  if (edgecut) *edgecut = 0;
  if (n_domain <= 1 || n_node == 0) {
    for (int i = 0; i < n_node; i++) domain[i] = 0;
    return 0;
  }
#ifdef FB_WITH_METIS
  std::vector<int64_t> xadj, adjncy;
  const int rc = part_node_graph(n_node, n_elem_type, elem_type_index, elem_type_item, elem_node_index, elem_node_item, xadj, adjncy);
  if (rc) return rc;
  int64_t nv = n_node, ncon = 1, nparts = n_domain, cut = 0;
  std::vector<int64_t> part((size_t)n_node, 0);
  const int ok = (method == 1)
                     ? METIS_PartGraphRecursive(&nv, &ncon, xadj.data(), adjncy.data(), nullptr, nullptr, nullptr, &nparts, nullptr, nullptr, nullptr, &cut, part.data())
                     : METIS_PartGraphKway(&nv, &ncon, xadj.data(), adjncy.data(), nullptr, nullptr, nullptr, &nparts, nullptr, nullptr, nullptr, &cut, part.data());
  if (ok != 1) return -4;   // METIS_OK == 1
  for (int i = 0; i < n_node; i++) {
    if (part[i] < 0 || part[i] >= n_domain) return -4;
    domain[i] = (int)part[i];
  }
  if (edgecut) *edgecut = cut;
  return 0;
#else
  (void)method; (void)n_elem_type; (void)elem_type_index; (void)elem_type_item; (void)elem_node_index; (void)elem_node_item;
  return -3;
#endif
}

}  // namespace fb
