// In-process node-based RCB domain decomposition with the tie behaviour of the reference's
// hecmw_part1 (hecmw1/tools/partitioner/hecmw_partition.c: rcb_partition :1570-1660 and its
// quick_sort :1469-1566) so that a synthetic cube can be split for N GPUs without writing a
// multi-GB .msh and running the external tool.  SURVEY §8(f) rank 4 ("hecmw_part1 in-process").
//
// At bisection level i every existing domain j < 2^i is sorted along the requested axis and the
// lower half (ceil(count/2) nodes in sorted order) moves to domain j + 2^i.  Which nodes of a tied
// coordinate plane move is decided by the (unstable) sort, so the sort is restated step for step.
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

namespace fb {

inline int rcb_quick_sort(int n, double *arr, int *brr, std::vector<int> &istack) {
  const int kLower = 50;   // QSORT_LOWER, hecmw_partition.c:64
  int jstack = 0, l = 0, ir = n - 1;
  const int nstack = (int)istack.size() - 1;
  for (;;) {
    if (ir - l < kLower) {           // straight insertion on small ranges
      for (int j = l + 1; j <= ir; j++) {
        const double a = arr[j];
        const int b = brr[j];
        int i = j - 1;
        for (; i >= l; i--) {
          if (arr[i] <= a) break;
          arr[i + 1] = arr[i];
          brr[i + 1] = brr[i];
        }
        arr[i + 1] = a;
        brr[i + 1] = b;
      }
      if (!jstack) return 0;
      ir = istack[jstack];
      l = istack[jstack - 1];
      jstack -= 2;
    } else {                         // median-of-three partitioning
      const int k = (l + ir) >> 1;
      std::swap(arr[k], arr[l + 1]); std::swap(brr[k], brr[l + 1]);
      if (arr[l] > arr[ir]) { std::swap(arr[l], arr[ir]); std::swap(brr[l], brr[ir]); }
      if (arr[l + 1] > arr[ir]) { std::swap(arr[l + 1], arr[ir]); std::swap(brr[l + 1], brr[ir]); }
      if (arr[l] > arr[l + 1]) { std::swap(arr[l], arr[l + 1]); std::swap(brr[l], brr[l + 1]); }
      int i = l + 1, j = ir;
      const double a = arr[l + 1];
      const int b = brr[l + 1];
      for (;;) {
        do i++; while (arr[i] < a);
        do j--; while (arr[j] > a);
        if (j < i) break;
        std::swap(arr[i], arr[j]); std::swap(brr[i], brr[j]);
      }
      arr[l + 1] = arr[j]; arr[j] = a;
      brr[l + 1] = brr[j]; brr[j] = b;
      jstack += 2;
      if (jstack > nstack) return -1;
      if (ir - i + 1 >= j - l) { istack[jstack] = ir; istack[jstack - 1] = i; ir = j - 1; }
      else { istack[jstack] = j - 1; istack[jstack - 1] = l; l = i; }
    }
  }
}

// axes[i] in {0,1,2}; domain[k] receives 0..2^n_div-1
inline int rcb_partition(int n, const double *coord, int n_div, const int *axes, int *domain) {
  std::vector<int> id(n), stack((size_t)n + 2);
  std::vector<double> value(n);
  for (int k = 0; k < n; k++) domain[k] = 0;
  for (int i = 0; i < n_div; i++) {
    if (axes[i] < 0 || axes[i] > 2) return -2;
    const int nd = 1 << i;
    for (int j = 0; j < nd; j++) {
      int counter = 0;
      for (int k = 0; k < n; k++)
        if (domain[k] == j) { id[counter] = k; value[counter] = coord[3 * (size_t)k + axes[i]]; counter++; }
      if (counter == 0) continue;
      if (rcb_quick_sort(counter, value.data(), id.data(), stack)) return -1;
      for (int k = 0; k < counter * 0.5; k++) domain[id[k]] = j + nd;   // 'k < counter * F_1_2'
    }
  }
  return 0;
}


// ---- the distributed mesh of ONE rank, as hecmw_part1 writes it (hecmw1/tools/partitioner/hecmw_partition.c;
// on-disk form hecmw1/src/common/hecmw_io_dist.c:1888-2284), node-based partition with depth-1 overlap:
//   * internal nodes (owned by `rank`) first in ascending global id, then the external (halo) nodes in ascending
//     global id; node_ID = (local id in the owning rank, owning rank);
//   * the elements that touch at least one internal node, in global order inside each element-type block;
//   * neighbour ranks ascending; per neighbour the import list (my external nodes it owns) and the export list
//     (my internal nodes it holds as external = my nodes in elements that also have a node it owns), both in
//     ascending global id.
// Identity with the real hecmw_part1 is pinned by tests/test_partition.py (fixtures made with the reference binary under oracle/_ref).
struct LocalMesh {
  int n_node = 0, nn_internal = 0, n_elem = 0;
  std::vector<int> global_node;        // n_node: 0-based global node index of every local node
  std::vector<int> node_ID;            // 2*n_node: (local id in owner (1-based), owner rank)
  std::vector<int> elem_type_index, elem_type_item;
  std::vector<int> elem_node_index, elem_node_item;   // 1-based LOCAL node ids
  std::vector<int> global_elem;        // n_elem: 0-based global element index
  std::vector<int> neighbor_pe, import_index, import_item, export_index, export_item;

  int build(int n_node_g, int n_elem_type, const int *elem_type_index_g, const int *elem_type_item_g,
            const int *elem_node_index_g, const int *elem_node_item_g, const int *domain, int rank, int nparts,
            int (*nn_of)(int)) {
    if (rank < 0 || rank >= nparts) return -1;
    // owner-local numbering: position among the owner's nodes in ascending global id
    std::vector<int> count(nparts, 0), local_in_owner(n_node_g);
    for (int g = 0; g < n_node_g; g++) {
      if (domain[g] < 0 || domain[g] >= nparts) return -2;
      local_in_owner[g] = ++count[domain[g]];
    }
    nn_internal = count[rank];
    // element selection + the set of nodes they use
    std::vector<char> used(n_node_g, 0);
    elem_type_index.assign(1, 0); elem_type_item.clear(); global_elem.clear();
    std::vector<int> kept_nn;
    for (int it = 0; it < n_elem_type; it++) {
      const int nn = nn_of(elem_type_item_g[it]);
      if (nn <= 0) return -3;
      int kept = 0;
      for (int e = elem_type_index_g[it]; e < elem_type_index_g[it + 1]; e++) {
        const int *c = elem_node_item_g + elem_node_index_g[e];
        bool mine = false;
        for (int j = 0; j < nn; j++) mine = mine || (domain[c[j] - 1] == rank);
        if (!mine) continue;
        for (int j = 0; j < nn; j++) used[c[j] - 1] = 1;
        global_elem.push_back(e);
        kept_nn.push_back(nn);
        kept++;
      }
      if (kept) { elem_type_item.push_back(elem_type_item_g[it]); elem_type_index.push_back(elem_type_index.back() + kept); }
    }
    n_elem = (int)global_elem.size();
    // local numbering: internal ascending, then external ascending
    global_node.clear();
    for (int g = 0; g < n_node_g; g++) if (domain[g] == rank) global_node.push_back(g);
    for (int g = 0; g < n_node_g; g++) if (used[g] && domain[g] != rank) global_node.push_back(g);
    n_node = (int)global_node.size();
    std::vector<int> g2l(n_node_g, 0);
    node_ID.resize(2 * (size_t)n_node);
    for (int l = 0; l < n_node; l++) {
      const int g = global_node[l];
      g2l[g] = l + 1;
      node_ID[2 * l] = local_in_owner[g];
      node_ID[2 * l + 1] = domain[g];
    }
    elem_node_index.assign(1, 0); elem_node_item.clear();
    for (int k = 0; k < n_elem; k++) {
      const int *c = elem_node_item_g + elem_node_index_g[global_elem[k]];
      for (int j = 0; j < kept_nn[k]; j++) elem_node_item.push_back(g2l[c[j] - 1]);
      elem_node_index.push_back((int)elem_node_item.size());
    }
    // export candidates: (other owner q, my node) for every kept element holding nodes of both
    std::vector<std::pair<int, int>> exp_pairs;
    for (int k = 0; k < n_elem; k++) {
      const int *c = elem_node_item_g + elem_node_index_g[global_elem[k]];
      const int nn = kept_nn[k];
      int others[32], n_others = 0;
      for (int j = 0; j < nn; j++) {
        const int q = domain[c[j] - 1];
        if (q == rank) continue;
        bool seen = false;
        for (int t = 0; t < n_others; t++) seen = seen || (others[t] == q);
        if (!seen && n_others < 32) others[n_others++] = q;
      }
      if (!n_others) continue;
      for (int j = 0; j < nn; j++)
        if (domain[c[j] - 1] == rank)
          for (int t = 0; t < n_others; t++) exp_pairs.emplace_back(others[t], c[j] - 1);
    }
    std::sort(exp_pairs.begin(), exp_pairs.end());
    exp_pairs.erase(std::unique(exp_pairs.begin(), exp_pairs.end()), exp_pairs.end());
    // neighbours = owners of my external nodes  U  ranks I export to
    std::vector<char> is_nb(nparts, 0);
    for (int l = nn_internal; l < n_node; l++) is_nb[domain[global_node[l]]] = 1;
    for (const auto &pr : exp_pairs) is_nb[pr.first] = 1;
    neighbor_pe.clear();
    for (int q = 0; q < nparts; q++) if (is_nb[q] && q != rank) neighbor_pe.push_back(q);
    import_index.assign(1, 0); export_index.assign(1, 0); import_item.clear(); export_item.clear();
    size_t ep = 0;
    for (int q : neighbor_pe) {
      for (int l = nn_internal; l < n_node; l++)
        if (domain[global_node[l]] == q) import_item.push_back(l + 1);
      import_index.push_back((int)import_item.size());
      while (ep < exp_pairs.size() && exp_pairs[ep].first < q) ep++;
      for (; ep < exp_pairs.size() && exp_pairs[ep].first == q; ep++) export_item.push_back(g2l[exp_pairs[ep].second]);
      export_index.push_back((int)export_item.size());
    }
    return 0;
  }
};

}  // namespace fb
