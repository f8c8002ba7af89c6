/* Plain-C consumer of include/fistr_b200.h: proves the header is valid C (not only C++), that the library links
 * with nothing but -lfistr_b200 from a C program, and exercises the host-only entry points (no GPU needed):
 * hecmw_mat_con and the hecmw_part1 RCB bisection.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "fistr_b200.h"

int main(void) {
  /* two hexahedra sharing a face: 12 nodes, node order 1-4 bottom CCW, 5-8 top */
  const double node[36] = {0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1,
                           1, 1, 1, 0, 1, 1, 2, 0, 0, 2, 1, 0, 2, 0, 1, 2, 1, 1};
  const int elem_type_index[2] = {0, 2}, elem_type_item[1] = {361};
  const int elem_node_index[3] = {0, 8, 16};
  const int elem_node_item[16] = {1, 2, 3, 4, 5, 6, 7, 8, 2, 9, 10, 3, 6, 11, 12, 7};
  int N, NP, NPL, NPU;
  int rc = fistr_b200_mat_con_count(12, 12, 1, elem_type_index, elem_type_item, elem_node_index, elem_node_item, &N,
                                    &NP, &NPL, &NPU);
  if (rc != FISTR_B200_OK) { printf("mat_con_count rc=%d %s\n", rc, fistr_b200_last_error()); return 1; }
  int indexL[13], indexU[13], itemL[64], itemU[64];
  rc = fistr_b200_mat_con_fill(indexL, itemL, indexU, itemU);
  if (rc != FISTR_B200_OK) return 2;
  /* 8 nodes see 8 nodes (7 neighbours), the 4 shared ones see all 12 (11 neighbours): 8*7 + 4*11 = 100 */
  if (N != 12 || NP != 12 || NPL + NPU != 100 || NPL != NPU) { printf("graph %d %d %d %d\n", N, NP, NPL, NPU); return 3; }
  const int axes[1] = {0};
  int domain[12];
  rc = fistr_b200_part_rcb(12, node, 1, axes, domain);
  if (rc != FISTR_B200_OK) return 4;
  int n0 = 0;
  for (int i = 0; i < 12; i++) n0 += (domain[i] == 0);
  if (n0 != 6) { printf("rcb split %d\n", n0); return 5; }
  /* METHOD=KMETIS of hecmw_part1 (when the build found the CUDA toolkit's METIS archive): 3 domains, every node owned */
  if (fistr_b200_part_metis_available()) {
    long long cut = 0;
    int cnt[3] = {0, 0, 0};
    rc = fistr_b200_part_metis(2, 3, 12, 1, elem_type_index, elem_type_item, elem_node_index, elem_node_item, domain, &cut);
    if (rc != FISTR_B200_OK) { printf("part_metis rc=%d (%s)\n", rc, fistr_b200_last_error()); return 6; }
    for (int i = 0; i < 12; i++) { if (domain[i] < 0 || domain[i] > 2) return 7; cnt[domain[i]]++; }
    if (cnt[0] + cnt[1] + cnt[2] != 12 || cut <= 0) return 8;
  }
  /* compute entry points must fail loudly without a device, never fall back to the CPU */
  rc = fistr_b200_init(0);
  printf("c_consumer ok: NPL=%d NPU=%d init rc=%d (%s)\n", NPL, NPU, rc, rc ? fistr_b200_last_error() : "gpu present");
  if (rc == FISTR_B200_OK) fistr_b200_finalize();
  return 0;
}
