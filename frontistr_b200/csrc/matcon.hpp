// hecmw_mat_con (hecmw1/src/solver/matrix/hecmw_mat_con.f90:23-268): node-node graph from the
// element connectivity -> indexL/itemL/indexU/itemU, rows sorted ascending, over all NP rows.
//
// The reference grows fixed-width tables with linear-probe insertion and restarts the whole
// pass on overflow, then bubble-sorts every row; the RESULT is simply "sorted unique neighbours,
// split at the diagonal", which is what every later routine binary-searches.  Here it is built
// from a node->element adjacency with per-row sort/unique, threaded over row ranges.
#pragma once
#include <algorithm>
#include <thread>
#include <vector>

namespace fb {

inline int etype_max_node(int ic_type) {
  switch (ic_type) {
    case 341: return 4;
    case 342: return 10;
    case 351: return 6;
    case 352: return 15;
    case 361: return 8;
    case 362: return 20;
    default: return -1;
  }
}

struct MatCon {
  int N = 0, NP = 0;
  std::vector<int> indexL, indexU, itemL, itemU;

  int build(int n_node, int nn_internal, int n_elem_type, const int *elem_type_index,
            const int *elem_type_item, const int *elem_node_index, const int *elem_node_item) {
    NP = n_node; N = nn_internal;
    const int n_elem = elem_type_index[n_elem_type];
    std::vector<int> enn(n_elem, 0);
    for (int it = 0; it < n_elem_type; it++) {
      int nn = etype_max_node(elem_type_item[it]);
      if (nn < 0) return -1;
      for (int e = elem_type_index[it]; e < elem_type_index[it + 1]; e++) enn[e] = nn;
    }
    // node -> element CSR
    std::vector<long long> nptr((size_t)NP + 1, 0);
    for (int e = 0; e < n_elem; e++) {
      const int *c = elem_node_item + elem_node_index[e];
      for (int a = 0; a < enn[e]; a++) {
        if (c[a] < 1 || c[a] > NP) return -2;
        nptr[c[a]]++;
      }
    }
    for (int i = 0; i < NP; i++) nptr[i + 1] += nptr[i];
    std::vector<int> nelem((size_t)nptr[NP]);
    {
      std::vector<long long> pos(nptr.begin(), nptr.end() - 1);
      for (int e = 0; e < n_elem; e++) {
        const int *c = elem_node_item + elem_node_index[e];
        for (int a = 0; a < enn[e]; a++) nelem[(size_t)pos[c[a] - 1]++] = e;
      }
    }
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 64u));
    if (NP < 4096) nt = 1;
    std::vector<std::vector<int>> tl(nt), tu(nt);
    std::vector<int> cntL((size_t)NP, 0), cntU((size_t)NP, 0);
    auto work = [&](int t) {
      const int r0 = (int)((long long)NP * t / nt), r1 = (int)((long long)NP * (t + 1) / nt);
      std::vector<int> buf;
      for (int i = r0; i < r1; i++) {
        buf.clear();
        for (long long k = nptr[i]; k < nptr[i + 1]; k++) {
          const int e = nelem[(size_t)k];
          const int *c = elem_node_item + elem_node_index[e];
          for (int a = 0; a < enn[e]; a++) buf.push_back(c[a]);
        }
        std::sort(buf.begin(), buf.end());
        buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
        for (int v : buf) {
          if (v < i + 1) { tl[t].push_back(v); cntL[i]++; }
          else if (v > i + 1) { tu[t].push_back(v); cntU[i]++; }
        }
      }
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(work, t);
      for (auto &x : th) x.join();
    }
    indexL.assign((size_t)NP + 1, 0); indexU.assign((size_t)NP + 1, 0);
    long long sl = 0, su = 0;
    for (int i = 0; i < NP; i++) {
      sl += cntL[i]; su += cntU[i];
      if (sl > 2147483647LL / 9 || su > 2147483647LL / 9) return -3;  // kint=4 limit, SURVEY F10
      indexL[i + 1] = (int)sl; indexU[i + 1] = (int)su;
    }
    itemL.resize((size_t)sl); itemU.resize((size_t)su);
    size_t ol = 0, ou = 0;
    for (int t = 0; t < nt; t++) {
      std::copy(tl[t].begin(), tl[t].end(), itemL.begin() + ol); ol += tl[t].size();
      std::copy(tu[t].begin(), tu[t].end(), itemU.begin() + ou); ou += tu[t].size();
    }
    return 0;
  }
};

}  // namespace fb
