// !SOLVER DUMPTYPE=MM|CSR|BSR (+ DUMPEXIT): hecmw_mat_dump / hecmw_mat_dump_rhs / hecmw_mat_dump_solution
// (hecmw1/src/solver/matrix/hecmw_matrix_dump.f90:30-372), called by hecmw_solve_iterative before the Krylov loop
// (Iterative.f90:112) and after the final residual (:172).  Host code: text files `dump_matrix_<call>_<rank>.{mm,csr,bsr,
// rhs,sol}` in the working directory, all NP rows, numbers in Fortran's E20.12E3 edit descriptor, one per line.  Besides
// being the reference's debugging aid, the BSR dump is the format in which a genuine fistr1 run can hand its K to a
// comparison with the matrix assembled here.  No reference output of it exists in the tree: the layout is
// restated from the WRITE statements, the number format from the Fortran standard ("parity unpinned").
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace fb {

// Fortran E20.12E3: [-]0.dddddddddddd E+eee, right-justified in 20 columns
inline void fortran_e20_12e3(double v, char out[24]) {
  if (std::isnan(v)) { snprintf(out, 24, "%20s", "NaN"); return; }
  if (std::isinf(v)) { snprintf(out, 24, "%20s", v < 0 ? "-Infinity" : "Infinity"); return; }
  char buf[40];
  snprintf(buf, sizeof buf, "%.11e", std::fabs(v));      // d.ddddddddddde[+-]xx
  char digits[16];
  digits[0] = buf[0];
  memcpy(digits + 1, buf + 2, 11);
  digits[12] = 0;
  int e10 = atoi(strchr(buf, 'e') + 1);
  int ex = (v == 0.0) ? 0 : e10 + 1;
  const bool neg = std::signbit(v) && v != 0.0;
  const unsigned eabs = (unsigned)(ex < 0 ? -ex : ex) % 1000u;   // |exponent| of a double is < 400
  snprintf(out, 24, "%c0.%.12sE%c%03u", neg ? '-' : ' ', digits, ex < 0 ? '-' : '+', eabs);
}

struct MatDumpView {   // host pointers, hecmwST_matrix layout
  int N, NP;
  const int *indexL, *itemL, *indexU, *itemU;
  const double *D, *AL, *AU, *B, *X;
};

inline std::string mat_dump_name(const char *dir, int numcall, int rank, const char *ext) {
  char f[96];
  snprintf(f, sizeof f, "dump_matrix_%d_%d%s", numcall, rank, ext);
  std::string p = (dir && *dir) ? std::string(dir) + "/" : std::string();
  return p + f;
}

inline int mat_dump_vector(const std::string &path, const double *v, size_t n) {
  FILE *f = fopen(path.c_str(), "w");
  if (!f) return -1;
  char s[24];
  for (size_t i = 0; i < n; i++) { fortran_e20_12e3(v[i], s); fputs(s, f); fputc('\n', f); }
  fclose(f);
  return 0;
}

// type 1 MatrixMarket coordinate, 2 scalar CSR, 3 block CSR; rows 1..NP like the reference ("n = hecMAT%NP")
inline int mat_dump_matrix(const std::string &path, int type, const MatDumpView &M) {
  FILE *f = fopen(path.c_str(), "w");
  if (!f) return -1;
  const int n = M.NP, nd = 3;
  const long long nblk = (long long)n + M.indexL[n] + M.indexU[n];
  char s[24];
  auto val = [&](double v) { fortran_e20_12e3(v, s); fputs(s, f); fputc('\n', f); };
  if (type == 1) {
    fputs("%%MatrixMarket matrix coordinate real general\n", f);
    fprintf(f, "%d %d %lld\n", n * nd, M.NP * nd, 9 * nblk);
    for (int i = 1; i <= n; i++)
      for (int id = 1; id <= nd; id++) {
        const int ii = (i - 1) * nd + id;
        for (int j = M.indexL[i - 1]; j < M.indexL[i]; j++)
          for (int jd = 1; jd <= nd; jd++) { fortran_e20_12e3(M.AL[9 * (size_t)j + 3 * (id - 1) + jd - 1], s); fprintf(f, "%d %d %s\n", ii, (M.itemL[j] - 1) * nd + jd, s); }
        for (int jd = 1; jd <= nd; jd++) { fortran_e20_12e3(M.D[9 * (size_t)(i - 1) + 3 * (id - 1) + jd - 1], s); fprintf(f, "%d %d %s\n", ii, (i - 1) * nd + jd, s); }
        for (int j = M.indexU[i - 1]; j < M.indexU[i]; j++)
          for (int jd = 1; jd <= nd; jd++) { fortran_e20_12e3(M.AU[9 * (size_t)j + 3 * (id - 1) + jd - 1], s); fprintf(f, "%d %d %s\n", ii, (M.itemU[j] - 1) * nd + jd, s); }
      }
  } else if (type == 2) {
    fputs("%%CSR matrix real general\n% nrow ncol nnonzero\n", f);
    fprintf(f, "%d %d %lld\n", n * nd, M.NP * nd, 9 * nblk);
    fputs("% index(0:nrow)\n0\n", f);
    long long idx = 0;
    for (int i = 1; i <= n; i++) {
      const long long nnz1 = (long long)nd * ((M.indexL[i] - M.indexL[i - 1]) + 1 + (M.indexU[i] - M.indexU[i - 1]));
      for (int id = 0; id < nd; id++) { idx += nnz1; fprintf(f, "%lld\n", idx); }
    }
    fputs("% item(1:nnonzero)\n", f);
    for (int i = 1; i <= n; i++)
      for (int id = 1; id <= nd; id++) {
        for (int j = M.indexL[i - 1]; j < M.indexL[i]; j++)
          for (int jd = 1; jd <= nd; jd++) fprintf(f, "%d\n", (M.itemL[j] - 1) * nd + jd);
        for (int jd = 1; jd <= nd; jd++) fprintf(f, "%d\n", (i - 1) * nd + jd);
        for (int j = M.indexU[i - 1]; j < M.indexU[i]; j++)
          for (int jd = 1; jd <= nd; jd++) fprintf(f, "%d\n", (M.itemU[j] - 1) * nd + jd);
      }
    fputs("% value(1:nnonzero)\n", f);
    for (int i = 1; i <= n; i++)
      for (int id = 1; id <= nd; id++) {
        for (int j = M.indexL[i - 1]; j < M.indexL[i]; j++)
          for (int jd = 0; jd < nd; jd++) val(M.AL[9 * (size_t)j + 3 * (id - 1) + jd]);
        for (int jd = 0; jd < nd; jd++) val(M.D[9 * (size_t)(i - 1) + 3 * (id - 1) + jd]);
        for (int j = M.indexU[i - 1]; j < M.indexU[i]; j++)
          for (int jd = 0; jd < nd; jd++) val(M.AU[9 * (size_t)j + 3 * (id - 1) + jd]);
      }
  } else {
    fputs("%%Block-CSR matrix real general\n% nrow ncol nnonzero ndof\n", f);
    fprintf(f, "%d %d %lld %d\n", n, M.NP, nblk, nd);
    fputs("% index(0:nrow)\n0\n", f);
    long long idx = 0;
    for (int i = 1; i <= n; i++) { idx += (M.indexL[i] - M.indexL[i - 1]) + 1 + (M.indexU[i] - M.indexU[i - 1]); fprintf(f, "%lld\n", idx); }
    fputs("% item(1:nnonzero)\n", f);
    for (int i = 1; i <= n; i++) {
      for (int j = M.indexL[i - 1]; j < M.indexL[i]; j++) fprintf(f, "%d\n", M.itemL[j]);
      fprintf(f, "%d\n", i);
      for (int j = M.indexU[i - 1]; j < M.indexU[i]; j++) fprintf(f, "%d\n", M.itemU[j]);
    }
    fputs("% value(1:nnonzero*ndof*ndof)\n", f);
    for (int i = 1; i <= n; i++) {
      for (int j = M.indexL[i - 1]; j < M.indexL[i]; j++) for (int e = 0; e < 9; e++) val(M.AL[9 * (size_t)j + e]);
      for (int e = 0; e < 9; e++) val(M.D[9 * (size_t)(i - 1) + e]);
      for (int j = M.indexU[i - 1]; j < M.indexU[i]; j++) for (int e = 0; e < 9; e++) val(M.AU[9 * (size_t)j + e]);
    }
  }
  fclose(f);
  return 0;
}

}  // namespace fb
