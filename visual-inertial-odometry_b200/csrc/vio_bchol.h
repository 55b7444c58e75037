// vio_bchol.h — host-side symbolic factorisation for the block-sparse Cholesky of the reduced camera system
// (S + lambda I) = L L^T on the 6x6 BSR pattern of vio_pack.h, natural (creation) order of the pose blocks.
// Pure C++ (no CUDA).  For a camera chain / ring the pattern is a block band plus the wrap-around border, whose fill
// stays inside band + border; other patterns work too, the fill is whatever the natural order gives (capped).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

struct BcholSymbolic {
    int nb = 0;
    long long nnzL = 0;                 // blocks of L incl. the diagonal
    std::vector<int> colptr;            // [nb+1] block column j holds blocks colptr[j] .. colptr[j+1]; the first is the diagonal
    std::vector<int> rowidx;            // [nnzL] block row of every stored block, ascending inside a column
    std::vector<long long> a_to_l;      // per BSR block k of S (row >= col only, else -1): index of the L block it initialises
    std::vector<long long> upd_ptr;     // [nb+1] per column: range of its update pairs
    std::vector<int> upd_a, upd_b;      // per pair: local indices (1-based positions inside the column) of the two source blocks
    std::vector<long long> upd_dst;     // per pair: index of the target block  (row of a, column = row of b)
    bool ok = false;
};

// rowptr/col: symmetric BSR pattern (full).  max_blocks: cap on nnz(L).
inline bool bchol_symbolic(int nb, const std::vector<int> &rowptr, const std::vector<int> &col, long long max_blocks, BcholSymbolic &Y) {
    Y = BcholSymbolic();
    Y.nb = nb;
    // struct(j) = { i > j : A_ij != 0 }  U  ( struct(c) \ {j} for every child c of j in the elimination tree )
    std::vector<std::vector<int>> st(nb);
    std::vector<std::vector<int>> children(nb);
    for (int j = 0; j < nb; ++j) {
        std::vector<int> s;
        for (int k = rowptr[j]; k < rowptr[j + 1]; ++k)
            if (col[k] > j) s.push_back(col[k]);  // symmetric pattern: row j's columns > j are column j's rows > j
        std::sort(s.begin(), s.end());
        for (int c : children[j]) {
            std::vector<int> merged;
            merged.reserve(s.size() + st[c].size());
            std::set_union(s.begin(), s.end(), st[c].begin() + 1, st[c].end(), std::back_inserter(merged));  // st[c][0] == j
            s.swap(merged);
        }
        s.erase(std::unique(s.begin(), s.end()), s.end());
        st[j].swap(s);
        if (!st[j].empty()) children[st[j][0]].push_back(j);
        Y.nnzL += 1 + (long long)st[j].size();
        if (Y.nnzL > max_blocks) return false;
    }
    Y.colptr.assign(nb + 1, 0);
    Y.rowidx.reserve(Y.nnzL);
    for (int j = 0; j < nb; ++j) {
        Y.rowidx.push_back(j);
        Y.rowidx.insert(Y.rowidx.end(), st[j].begin(), st[j].end());
        Y.colptr[j + 1] = (int)Y.rowidx.size();
    }
    auto find = [&](int i, int j) -> long long {  // index of block (i, j), i >= j
        auto b = Y.rowidx.begin() + Y.colptr[j], e = Y.rowidx.begin() + Y.colptr[j + 1];
        auto it = std::lower_bound(b, e, i);
        return (it != e && *it == i) ? (long long)(it - Y.rowidx.begin()) : -1;
    };
    Y.a_to_l.assign(col.size(), -1);
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            if (col[k] <= i) Y.a_to_l[k] = find(i, col[k]);
    Y.upd_ptr.assign(nb + 1, 0);
    for (int j = 0; j < nb; ++j) {
        const long long c = Y.colptr[j + 1] - Y.colptr[j] - 1;
        Y.upd_ptr[j + 1] = Y.upd_ptr[j] + c * (c + 1) / 2;
    }
    if (Y.upd_ptr[nb] > 64LL * 1000 * 1000) return false;
    Y.upd_a.resize(Y.upd_ptr[nb]); Y.upd_b.resize(Y.upd_ptr[nb]); Y.upd_dst.resize(Y.upd_ptr[nb]);
    for (int j = 0; j < nb; ++j) {
        const int c = Y.colptr[j + 1] - Y.colptr[j] - 1;
        long long q = Y.upd_ptr[j];
        for (int b = 1; b <= c; ++b)
            for (int a = b; a <= c; ++a, ++q) {
                Y.upd_a[q] = a; Y.upd_b[q] = b;
                Y.upd_dst[q] = find(Y.rowidx[Y.colptr[j] + a], Y.rowidx[Y.colptr[j] + b]);
                if (Y.upd_dst[q] < 0) return false;  // cannot happen for a valid symbolic factorisation
            }
    }
    Y.ok = true;
    return true;
}
