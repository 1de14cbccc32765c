// ordering.cu -- host-side elimination orderings for the reduced KKT system [dI + G'W^-2 G, A'; A, -dI].
// scpb_order_rcm: reverse Cuthill-McKee on the KKT graph (variables coupled through a cone row -- or through any row of
// the same second-order cone, whose W^-2 block is dense --, equality rows coupled to their variables), every equality
// row moved just behind the last of its variables.  It is what a caller without stage information uses (the
// MathOptInterface shim, julia/SCPToolboxB200.jl; scptoolbox.jl_b200/ordering.py has the same heuristic on scipy for
// the Python host); the SCP templates supply the better stage-wise order themselves (ordering.py: stage_order).
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>
#include "../../include/scpb.h"

extern "C" int32_t scpb_order_rcm(int32_t n, int32_t p, int32_t m, const int32_t *A_rowptr, const int32_t *A_colind,
                                  const int32_t *G_rowptr, const int32_t *G_colind, int32_t l, int32_t nsoc,
                                  const int32_t *soc_dims, int32_t *perm)
{
    if (n <= 0 || p < 0 || m < 0 || l < 0 || l > m || nsoc < 0 || !perm || (p > 0 && (!A_rowptr || !A_colind)) ||
        (m > 0 && (!G_rowptr || !G_colind)) || (nsoc > 0 && !soc_dims))
        return SCPB_ERR_ARG;
    long long tot = l;
    for (int k = 0; k < nsoc; k++) { if (soc_dims[k] <= 0) return SCPB_ERR_ARG; tot += soc_dims[k]; }
    if (tot != m) return SCPB_ERR_ARG;
    const int nk = n + p;
    std::vector<std::vector<int>> adj(nk);
    auto clique = [&](std::vector<int> &cols) {
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        for (size_t a = 0; a < cols.size(); a++)
            for (size_t b = a + 1; b < cols.size(); b++) { adj[cols[a]].push_back(cols[b]); adj[cols[b]].push_back(cols[a]); }
    };
    std::vector<int> cols;
    for (int r = 0; r < l; r++) {
        cols.assign(G_colind + G_rowptr[r], G_colind + G_rowptr[r + 1]);
        for (int c : cols) if (c < 0 || c >= n) return SCPB_ERR_ARG;
        clique(cols);
    }
    int row = l;
    for (int k = 0; k < nsoc; k++) {
        cols.clear();
        for (int r = row; r < row + soc_dims[k]; r++)
            for (int q = G_rowptr[r]; q < G_rowptr[r + 1]; q++) {
                if (G_colind[q] < 0 || G_colind[q] >= n) return SCPB_ERR_ARG;
                cols.push_back(G_colind[q]);
            }
        clique(cols);
        row += soc_dims[k];
    }
    for (int r = 0; r < p; r++)
        for (int q = A_rowptr[r]; q < A_rowptr[r + 1]; q++) {
            const int c = A_colind[q];
            if (c < 0 || c >= n) return SCPB_ERR_ARG;
            adj[n + r].push_back(c); adj[c].push_back(n + r);
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    // Cuthill-McKee per connected component, started from a node of minimum degree, neighbours by increasing degree
    std::vector<int> order; order.reserve(nk);
    std::vector<char> seen(nk, 0);
    std::vector<int> by_deg(nk);
    std::iota(by_deg.begin(), by_deg.end(), 0);
    std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return adj[a].size() < adj[b].size(); });
    std::vector<int> nb;
    for (int s : by_deg) {
        if (seen[s]) continue;
        size_t head = order.size();
        order.push_back(s); seen[s] = 1;
        while (head < order.size()) {
            const int v = order[head++];
            nb.clear();
            for (int w : adj[v]) if (!seen[w]) { nb.push_back(w); seen[w] = 1; }
            std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[a].size() < adj[b].size(); });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    std::vector<double> key(nk);
    std::vector<int> pos(nk);
    for (int k = 0; k < nk; k++) pos[order[k]] = k;
    for (int v = 0; v < nk; v++) key[v] = pos[v];
    for (int r = 0; r < p; r++)        // an equality row is eliminated after the last of its variables (its pivot is -delta
        for (int q = A_rowptr[r]; q < A_rowptr[r + 1]; q++)   // until then: quasi-definite ordering)
            key[n + r] = std::max(key[n + r], pos[A_colind[q]] + 0.5);
    std::vector<int> idx(nk);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return key[a] < key[b]; });
    for (int k = 0; k < nk; k++) perm[k] = idx[k];
    return SCPB_OK;
}
