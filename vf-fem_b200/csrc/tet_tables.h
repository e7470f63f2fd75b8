// Gather-friendly copies of the tetrahedral mesh tables for the thread-per-node assembly kernels
// (assembly.cu), built once per engine on the host.  The generic tables are structure-of-arrays
// (cells[a * ne + e], xyz[c * nn + node]) and the CSR slot of a neighbour is found by scanning the
// node's column list: per (node, cell) visit that is 4 + 12 scattered 32-byte sectors plus four
// dependent scans.  Here a cell's connectivity is one 16-byte record, a node's coordinates one
// 32-byte record, and the four slots of every (node, cell) pair are precomputed (8 bits each).
#pragma once

#include <cstdint>
#include <vector>

namespace vf {

// cells4: (ne, 4) node ids; xyz4: (nn, 4) x, y, z, 0; slots: per entry t of n2e, byte c = CSR slot
// of local node c of the pair's cell in the block row of the pair's node.  Returns false (tables
// unusable, callers keep the generic path) if a row has more than 255 blocks or is inconsistent.
inline bool build_tet_gather_tables(int nn, int ne, const double* xyz, const int32_t* cells,
                                    const int32_t* brptr, const int32_t* bcol,
                                    const int32_t* n2e_ptr, const int32_t* n2e,
                                    std::vector<int32_t>& cells4, std::vector<double>& xyz4,
                                    std::vector<uint32_t>& slots) {
  cells4.resize(4 * (size_t)ne);
  for (int e = 0; e < ne; ++e)
    for (int a = 0; a < 4; ++a) cells4[4 * (size_t)e + a] = cells[(size_t)a * ne + e];
  xyz4.assign(4 * (size_t)nn, 0.0);
  for (int i = 0; i < nn; ++i)
    for (int c = 0; c < 3; ++c) xyz4[4 * (size_t)i + c] = xyz[(size_t)c * nn + i];
  slots.resize((size_t)n2e_ptr[nn]);
  for (int i = 0; i < nn; ++i) {
    const int b0 = brptr[i], deg = brptr[i + 1] - b0;
    if (deg > 255) return false;
    for (int t = n2e_ptr[i]; t < n2e_ptr[i + 1]; ++t) {
      const int e = n2e[t] >> 2;
      uint32_t packed = 0;
      for (int c = 0; c < 4; ++c) {
        const int node = cells4[4 * (size_t)e + c];
        int k = 0;
        while (k < deg && bcol[b0 + k] != node) ++k;
        if (k == deg) return false;
        packed |= (uint32_t)k << (8 * c);
      }
      slots[t] = packed;
    }
  }
  return true;
}

}  // namespace vf
