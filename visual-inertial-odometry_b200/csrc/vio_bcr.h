// vio_bcr.h — host-side plan of the BLOCK CYCLIC REDUCTION solver for the reduced camera system of a camera CHAIN / RING
// (VIO_SOLVER_BCR): (S + lambda I) x = b exactly, replacing S.ldlt().solve (A17/src/backend/problem.cc:434-440) on the
// block-sparse S of configs 4/5.  Pure C++ (no CUDA); the device side is vio_bcr.cuh, a CPU interpreter of the same
// plan lives in tests/host_emul.cu.
//
// Structure exploited: cameras that co-observe landmarks are close in creation order, so S is block-banded (half
// bandwidth w pose blocks) plus the wrap-around corner of a closed loop.  Grouping w consecutive pose blocks into one
// NODE (a dense M x M super-block, M = 6 w) makes S block-TRIDIAGONAL and CYCLIC over n = NB / w nodes:
//     D_i (node i with itself),  E_i (node i with node i+1 mod n).
// Cyclic reduction eliminates every other node of the cycle per level (a nested-dissection Cholesky):
//     eliminate k with neighbours l, r:   D_k = L L^T,  U = L^-T,
//         W_l = U^T A[k,l],  W_r = U^T A[k,r],  y_k = U^T b_k
//         D_l -= W_l^T W_l,  D_r -= W_r^T W_r,  A[l,r] = -W_l^T W_r,  b_l -= W_l^T y_k,  b_r -= W_r^T y_k
//     back-substitution (reverse order):  x_k = U (y_k - W_l x_l - W_r x_r)
// log2(n) levels, every level a set of independent dense M x M operations -> the plan is a list of ITEMS (one per active
// node and level, then one per node for the back-substitution) in an order in which every item depends only on earlier
// ones; a persistent kernel hands them out through an atomic counter and synchronises through per-item flags.
// Pose blocks without any off-diagonal block (fixed vertices: zero rows + lambda on the diagonal,
// A17/src/backend/problem.cc:325,340,434-436) are not part of the chain; they are solved on their own.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#define BCR_ELIM 1     // the node is eliminated at this level (factorise, form W_l / W_r / y)
#define BCR_MERGE 2    // two-node cycle: both couplings join the same neighbour, A[k,r] += A[k,l], no left neighbour
#define BCR_BACKSUB 4  // back-substitution item
#define BCR_MAX_M 72   // five M x M operand tiles must fit in shared memory (5 x 41.5 KB)
#define BCR_DST_DIAG (1LL << 62)

struct BcrItem {
    int node, kind;
    // per side (0 = left, 1 = right): D_node -= W^T W, b_node -= W^T y[upd_node] with W = pool slot upd_slot (rows: the
    // neighbour eliminated one level earlier); -1 = none.  When the side's coupling is a product, c?_a == upd_slot.
    int upd_slot[2], upd_node[2];
    // couplings of an eliminated node, rows = this node.  mode 0: none, 1: materialised in slot `a` (transpose on load if
    // `b` != 0), 2: product  -pool[a]^T pool[b]
    int cl_mode, cl_a, cl_b, cr_mode, cr_a, cr_b;
    int cl_slot, cr_slot;  // where W_l / W_r go (-1 = none)
    int left, right;       // neighbour nodes at elimination time (-1 = none)
    int dep[6];            // items that must be complete first (-1 = none)
    int pad[2];
};
static_assert(sizeof(BcrItem) == 96, "BcrItem layout is shared with the device");

struct BcrPlan {
    bool ok = false;
    int nb = 0;          // pose blocks of S
    int n = 0;           // nodes
    int w = 0;           // half bandwidth in pose blocks
    int mb = 0, M = 0;   // pose blocks per node (padded), node dimension 6 * mb
    int ld = 0;          // row stride of a tile (doubles): M, or M + 4 where M would put the rows of a 4-row operand fragment on the
                         // same shared-memory banks (the DMMA fragment loads read 4 rows x 8 columns); a tile is M x ld
    int n_slots = 0;     // M x M tiles in the pool: [0, n) = D / U of the nodes, then couplings / W tiles
    int n_elim_items = 0;
    std::vector<int> blk_node, blk_loc;   // [nb] node and position inside the node of every pose block (-1: isolated)
    std::vector<int> node_size;           // [n] pose blocks per node (<= mb; the rest of the tile is identity padding)
    std::vector<int> iso;                 // isolated pose blocks
    std::vector<long long> dst;           // [nnzb] per BSR block: offset of its (0,0) element in the pool (| BCR_DST_DIAG for a
                                          // diagonal block, which receives lambda), or -1 (not loaded)
    std::vector<BcrItem> items;           // elimination items, then back-substitution items
    std::vector<int> level_of_item;       // diagnostics
    int n_levels = 0;
};

// rowptr/col: symmetric 6x6 BSR pattern (both triangles, diagonal present).  Returns plan.ok = false when the pattern is
// not a (cyclic) block band narrow enough for BCR_MAX_M.
inline void bcr_plan(int nb, const std::vector<int> &rowptr, const std::vector<int> &col, BcrPlan &Y) {
    Y = BcrPlan();
    Y.nb = nb;
    if (nb <= 0 || (int)rowptr.size() != nb + 1) return;
    // ---- chain = pose blocks with at least one off-diagonal block, in creation order
    std::vector<int> pos(nb, -1), chain;
    for (int i = 0; i < nb; ++i) {
        bool off = false;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) off |= col[k] != i;
        if (off) { pos[i] = (int)chain.size(); chain.push_back(i); }
        else Y.iso.push_back(i);
    }
    const int nc = (int)chain.size();
    if (nc < 3) return;
    // ---- cyclic half bandwidth
    int w = 1;
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];
            if (j == i) continue;
            if (pos[i] < 0 || pos[j] < 0) return;  // asymmetric pattern
            const int d = std::abs(pos[i] - pos[j]);
            w = std::max(w, std::min(d, nc - d));
        }
    const int n = nc / w;
    if (n < 3) return;
    const int base = nc / n, rem = nc % n;
    int mb = base + (rem > 0 ? 1 : 0);
    mb += mb & 1;  // even: the tile dimension 6 mb is a multiple of 4 (4x4 register tiles, 16-byte shared loads)
    if (6 * mb > BCR_MAX_M) return;
    Y.n = n; Y.w = w; Y.mb = mb; Y.M = 6 * mb;
    Y.ld = (Y.M % 16 == 4 || Y.M % 16 == 12) ? Y.M : Y.M + 4;
    Y.blk_node.assign(nb, -1); Y.blk_loc.assign(nb, -1); Y.node_size.assign(n, 0);
    {
        int c = 0;
        for (int a = 0; a < n; ++a) {
            const int sz = base + (a < rem ? 1 : 0);
            Y.node_size[a] = sz;
            for (int q = 0; q < sz; ++q, ++c) { Y.blk_node[chain[c]] = a; Y.blk_loc[chain[c]] = q; }
        }
    }
    // ---- the pattern must be cyclic block-tridiagonal over the nodes
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];
            if (j == i) continue;
            const int a = Y.blk_node[i], b = Y.blk_node[j];
            const int d = (b - a + n) % n;
            if (!(d == 0 || d == 1 || d == n - 1)) return;
        }
    const int M = Y.ld;  // row stride of the tiles (the loader map below only needs the stride)
    const long long MM = (long long)Y.M * Y.ld;
    // ---- level schedule
    struct Coup { int mode, a, b, rows; };  // mode 1: materialised slot a (rows = node `rows`); mode 2: product of eliminated node a
    struct ElimInfo { int item = -1, wl = -1, wr = -1; };
    std::vector<ElimInfo> einfo(n);
    std::vector<int> last_item(n, -1);  // last item that wrote D / b of the node
    std::vector<int> act(n);
    for (int a = 0; a < n; ++a) act[a] = a;
    std::vector<Coup> coup(n);
    int next_slot = n;
    for (int p = 0; p < n; ++p) {
        // level-0 couplings are written by the loader with rows = the node that is eliminated first (odd position)
        const int rows = (p & 1) ? act[p] : ((p + 1 < n) ? act[p + 1] : act[p]);
        coup[p] = {1, next_slot++, 0, rows};
    }
    std::vector<int> lvl0_slot_rows(n);
    for (int p = 0; p < n; ++p) lvl0_slot_rows[p] = coup[p].rows;
    auto new_item = [&](int node, int kind, int level) -> BcrItem & {
        BcrItem it;
        it.node = node; it.kind = kind;
        it.upd_slot[0] = it.upd_slot[1] = it.upd_node[0] = it.upd_node[1] = -1;
        it.cl_mode = it.cr_mode = 0; it.cl_a = it.cl_b = it.cr_a = it.cr_b = -1;
        it.cl_slot = it.cr_slot = -1; it.left = it.right = -1;
        for (int &d : it.dep) d = -1;
        it.pad[0] = it.pad[1] = 0;
        Y.items.push_back(it);
        Y.level_of_item.push_back(level);
        return Y.items.back();
    };
    bool dep_overflow = false;
    auto add_dep = [&](BcrItem &it, int d) {
        if (d < 0) return;
        for (int &x : it.dep) {
            if (x == d) return;
            if (x < 0) { x = d; return; }
        }
        dep_overflow = true;
    };
    std::vector<int> prev_elim_left, prev_elim_right;  // per position of the CURRENT level: eliminated neighbours of the previous level
    prev_elim_left.assign(n, -1); prev_elim_right.assign(n, -1);
    int level = 0;
    std::vector<int> elim_order;  // nodes in elimination order (for the back-substitution)
    for (;;) {
        const int nl = (int)act.size();
        // which positions are eliminated at this level
        auto is_elim = [&](int p) { return nl == 1 ? true : (p & 1) != 0; };
        // eliminated items first (they are the critical path), then the kept nodes' updates
        for (int pass = 0; pass < 2; ++pass)
            for (int p = 0; p < nl; ++p) {
                const bool el = is_elim(p);
                if (el != (pass == 0)) continue;
                const int k = act[p];
                const int eL = prev_elim_left[p], eR = prev_elim_right[p];
                if (!el && eL < 0 && eR < 0) continue;  // nothing to do for a kept node at level 0
                BcrItem &it = new_item(k, el ? BCR_ELIM : 0, level);
                const int me = (int)Y.items.size() - 1;
                add_dep(it, last_item[k]);
                if (eL >= 0) {  // side 0: k is the right neighbour of eL
                    it.upd_slot[0] = einfo[eL].wr; it.upd_node[0] = eL;
                    add_dep(it, einfo[eL].item);
                }
                if (eR >= 0 && eR != eL) {  // side 1: k is the left neighbour of eR
                    it.upd_slot[1] = einfo[eR].wl; it.upd_node[1] = eR;
                    add_dep(it, einfo[eR].item);
                }
                last_item[k] = me;
                if (!el) continue;
                elim_order.push_back(k);
                einfo[k].item = me;
                if (nl == 1) continue;  // the last node: no couplings
                const int pl = (p - 1 + nl) % nl, pr = (p + 1) % nl;
                const Coup &cL = coup[pl], &cR = coup[p];
                auto fill = [&](const Coup &c, bool is_left, int &mode, int &a, int &b) {
                    if (c.mode == 1) { mode = 1; a = c.a; b = c.rows == k ? 0 : 1; }
                    else {
                        mode = 2;
                        // left coupling A[k,l] = -W_r(e)^T W_l(e) ; right coupling A[k,r] = -W_l(e)^T W_r(e)
                        a = is_left ? einfo[c.a].wr : einfo[c.a].wl;
                        b = is_left ? einfo[c.a].wl : einfo[c.a].wr;
                        add_dep(Y.items[me], einfo[c.a].item);
                    }
                };
                BcrItem &e = Y.items[me];
                fill(cL, true, e.cl_mode, e.cl_a, e.cl_b);
                fill(cR, false, e.cr_mode, e.cr_a, e.cr_b);
                e.left = act[pl]; e.right = act[pr];
                // W tiles: a materialised coupling is overwritten in place, a product needs a fresh tile
                e.cl_slot = cL.mode == 1 ? cL.a : next_slot++;
                e.cr_slot = cR.mode == 1 ? cR.a : next_slot++;
                if (nl == 2) {  // both couplings join the same neighbour
                    e.kind |= BCR_MERGE;
                    e.left = -1;
                    e.cl_slot = -1;
                }
                einfo[k].wl = e.cl_slot; einfo[k].wr = e.cr_slot;
            }
        if (nl == 1) break;
        // next level: even positions stay; couplings between consecutive survivors
        std::vector<int> act2;
        std::vector<Coup> coup2;
        std::vector<int> pel, per;
        for (int p = 0; p < nl; p += 2) {
            act2.push_back(act[p]);
            if (p + 1 < nl) coup2.push_back({2, act[p + 1], 0, -1});  // product over the eliminated node act[p+1]
            else coup2.push_back(coup[p]);                             // odd level size: (last, first) carried over
            // eliminated neighbours of this survivor at the level just processed
            int l = -1, r = -1;
            if (p + 1 < nl) r = act[p + 1];
            if (p >= 1) l = act[p - 1];
            else if ((nl & 1) == 0) l = act[nl - 1];
            pel.push_back(l); per.push_back(r);
        }
        if (act2.size() == 1) coup2.clear();
        act.swap(act2); coup.swap(coup2);
        prev_elim_left.swap(pel); prev_elim_right.swap(per);
        ++level;
    }
    Y.n_levels = level + 1;
    Y.n_elim_items = (int)Y.items.size();
    Y.n_slots = next_slot;
    // ---- back-substitution: reverse elimination order
    std::vector<int> bs_item(n, -1);
    for (int q = (int)elim_order.size() - 1; q >= 0; --q) {
        const int k = elim_order[q];
        const BcrItem e = Y.items[einfo[k].item];
        BcrItem &it = new_item(k, BCR_BACKSUB, level + 1 + ((int)elim_order.size() - 1 - q));
        it.left = e.left; it.right = e.right; it.cl_slot = e.cl_slot; it.cr_slot = e.cr_slot;
        add_dep(it, einfo[k].item);
        if (e.left >= 0) add_dep(it, bs_item[e.left]);
        if (e.right >= 0) add_dep(it, bs_item[e.right]);
        bs_item[k] = (int)Y.items.size() - 1;
    }
    // ---- loader map: BSR block -> pool offset
    Y.dst.assign(col.size(), -1);
    for (int i = 0; i < nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];
            const int a = Y.blk_node[i], b = Y.blk_node[j];
            if (a < 0 || b < 0) continue;  // isolated block: solved separately
            if (a == b) {
                Y.dst[k] = ((long long)a * MM + (long long)(6 * Y.blk_loc[i]) * M + 6 * Y.blk_loc[j]) | (i == j ? BCR_DST_DIAG : 0);
                continue;
            }
            const int p = ((b - a + n) % n == 1) ? a : b;  // coupling between nodes p and p+1
            if (lvl0_slot_rows[p] != a) continue;          // the twin block (j, i) is the one stored
            Y.dst[k] = (long long)(n + p) * MM + (long long)(6 * Y.blk_loc[i]) * M + 6 * Y.blk_loc[j];
        }
    if (dep_overflow) return;
    // every dependency must point backwards (deadlock freedom of the in-order work queue)
    for (size_t q = 0; q < Y.items.size(); ++q)
        for (int d : Y.items[q].dep)
            if (d >= (int)q) return;
    Y.ok = true;
}
