// In-process node-based RCB domain decomposition with the tie behaviour of the reference's
// hecmw_part1 (hecmw1/tools/partitioner/hecmw_partition.c: rcb_partition :1570-1660 and its
// quick_sort :1469-1566) so that a synthetic cube can be split for N GPUs without writing a
// multi-GB .msh and running the external tool.  SURVEY §8(f) rank 4 ("hecmw_part1 in-process").
//
// At bisection level i every existing domain j < 2^i is sorted along the requested axis and the
// lower half (ceil(count/2) nodes in sorted order) moves to domain j + 2^i.  Which nodes of a tied
// coordinate plane move is decided by the (unstable) sort, so the sort is restated step for step.
#pragma once
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

}  // namespace fb
