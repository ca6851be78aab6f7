// Native flow-map builder ("next" row 3 of SURVEY.md §8f): what the reference obtains from rcmtool
// (un-vendored Rust crate) through CmaUtils::get_transition_matrix, get_cumulative_probabilities and
// get_diag_transition (documentation/pages/03_cma.md:26-63; call sites apps/libs/simulation/src/
// simulation.cpp:99-137), restated from the documentation:
//   * transition matrix M (COO): off-diagonals F(i,j), diagonal -sum_k F(i,k)            (03_cma.md:33-53)
//   * out_flows[i] = sum_k F(i,k)                              (diag_transition, positive, domain.cpp:66-67)
//   * neighbours of i in order of appearance, padded with i itself (test_transport.cpp:97-114)
//   * cumulative probability rows: running sum of P(I_j|L) = F(i,j) / sum_k F(i,k), non-decreasing,
//     last real entry exactly 1, padding 1; a compartment without out-flow has an all-zero row
//     (03_cma.md:55-63, test_transport.cpp:50-78)
// Host-only code (no device work): the arrays feed bmc_domain_update / bmc_liquid_set_transition.
#include <cstdint>
#include <vector>

#include "../../include/bmc.h"

extern "C" int bmc_cma_build(uint64_t n, uint64_t nnz, const uint64_t* from, const uint64_t* to, const double* flow,
                             uint64_t n_cols, uint64_t* n_cols_out, uint64_t* neighbors, double* cumulative_probability,
                             double* out_flows, uint64_t* tr_rows, uint64_t* tr_cols, double* tr_vals) {
  if (n == 0 || (nnz && (!from || !to || !flow))) return BMC_ERR_INVALID;
  std::vector<uint64_t> degree(n, 0);
  for (uint64_t e = 0; e < nnz; ++e) {
    if (from[e] >= n || to[e] >= n) return BMC_ERR_RANGE;
    if (!(flow[e] >= 0.0)) return BMC_ERR_INVALID;  // flows are non-negative (probability_leaving.hpp:42)
    if (from[e] != to[e] && flow[e] > 0.0) degree[from[e]]++;
  }
  uint64_t max_deg = 1;
  for (uint64_t i = 0; i < n; ++i) if (degree[i] > max_deg) max_deg = degree[i];
  const uint64_t m = n_cols ? n_cols : max_deg;
  if (n_cols_out) *n_cols_out = m;
  if (m < max_deg) return BMC_ERR_RANGE;
  if (!neighbors && !cumulative_probability && !out_flows && !tr_rows) return BMC_OK;  // size query
  std::vector<double> out(n, 0.0);
  for (uint64_t e = 0; e < nnz; ++e) if (from[e] != to[e]) out[from[e]] += flow[e];
  if (out_flows) for (uint64_t i = 0; i < n; ++i) out_flows[i] = out[i];
  if (neighbors && cumulative_probability) {
    std::vector<uint64_t> fill(n, 0);
    std::vector<double> run(n, 0.0);
    for (uint64_t i = 0; i < n; ++i)
      for (uint64_t k = 0; k < m; ++k) { neighbors[i * m + k] = i; cumulative_probability[i * m + k] = out[i] > 0.0 ? 1.0 : 0.0; }
    for (uint64_t e = 0; e < nnz; ++e) {
      const uint64_t i = from[e];
      if (i == to[e] || !(flow[e] > 0.0)) continue;
      const uint64_t k = fill[i]++;
      run[i] += flow[e];
      neighbors[i * m + k] = to[e];
      double c = run[i] / out[i];
      if (c > 1.0 || fill[i] == degree[i]) c = 1.0;  // the last real neighbour closes the CDF exactly
      cumulative_probability[i * m + k] = c;
    }
  }
  if (tr_rows && tr_cols && tr_vals) {  // nnz_offdiag + n triplets: off-diagonals in input order, then the diagonal
    uint64_t t = 0;
    for (uint64_t e = 0; e < nnz; ++e)
      if (from[e] != to[e] && flow[e] > 0.0) { tr_rows[t] = from[e]; tr_cols[t] = to[e]; tr_vals[t] = flow[e]; ++t; }
    for (uint64_t i = 0; i < n; ++i) { tr_rows[t] = i; tr_cols[t] = i; tr_vals[t] = -out[i]; ++t; }
  }
  return BMC_OK;
}
