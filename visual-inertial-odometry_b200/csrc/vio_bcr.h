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
#define BCR_EXPORT 8   // open chain only: the coupling left between the two pinned end nodes, -pool[cl_a]^T pool[cl_b], written to
                       // tile cl_slot of the EXPORT pool (the interface system of the multi-GPU solve)
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

// The level schedule over n nodes, independent of what the nodes hold.
//   cyclic (open_chain = false): nodes 0..n-1 form a cycle, coupling p joins nodes p and p+1 (mod n); every node is
//       eliminated (n >= 2).
//   open chain (open_chain = true): nodes 0..n-1 form a chain whose two END nodes are PINNED: they are never eliminated,
//       they only collect the Schur updates of the interior; the schedule ends with an EXPORT item for the coupling left
//       between them.  This is one rank's part of the multi-GPU solve: the ends are interface nodes shared with the
//       neighbouring ranks (n >= 3).
// Pool slots: [0, n) = D / U of the nodes, [n, n + #couplings) = level-0 couplings, then the W tiles of later levels.
struct BcrSched {
    bool ok = false;
    int n = 0;
    bool open_chain = false;
    int n_slots = 0, n_elim_items = 0, n_levels = 0;
    std::vector<int> lvl0_rows;          // per level-0 coupling p: the node whose rows the loader must store (the other is columns)
    std::vector<BcrItem> items;          // elimination (+ export) items, then back-substitution items
    std::vector<int> level_of_item;      // diagnostics
};

// export_rows_left: orientation of the exported coupling tile (rows = end node 0, else rows = end node n-1); export_tile:
// its tile index in the export pool
inline void bcr_schedule(int n, bool open_chain, bool export_rows_left, int export_tile, BcrSched &Y) {
    Y = BcrSched();
    Y.n = n; Y.open_chain = open_chain;
    if (n < (open_chain ? 3 : 2)) return;
    struct Coup { int mode, a, b, rows; };  // mode 1: materialised slot a (rows = node `rows`); mode 2: product of eliminated node a
    struct ElimInfo { int item = -1, wl = -1, wr = -1; };
    std::vector<ElimInfo> einfo(n);
    std::vector<int> last_item(n, -1);  // last item that wrote D / b of the node
    std::vector<int> act(n);
    for (int a = 0; a < n; ++a) act[a] = a;
    const int nc0 = open_chain ? n - 1 : n;
    std::vector<Coup> coup(nc0);
    int next_slot = n;
    auto elim_at = [&](int p, int nl) { return open_chain ? ((p & 1) != 0 && p != nl - 1) : (nl == 1 ? true : (p & 1) != 0); };
    for (int p = 0; p < nc0; ++p) {
        // level-0 couplings are written by the loader with rows = the end that is eliminated first (odd position)
        int rows;
        if (elim_at(p, n)) rows = act[p];
        else if (p + 1 < n && elim_at(p + 1, n)) rows = act[p + 1];
        else rows = act[p];
        coup[p] = {1, next_slot++, 0, rows};
    }
    Y.lvl0_rows.resize(nc0);
    for (int p = 0; p < nc0; ++p) Y.lvl0_rows[p] = coup[p].rows;
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
    std::vector<int> prev_elim_left(n, -1), prev_elim_right(n, -1);  // per position of the CURRENT level: eliminated neighbours of the previous level
    int level = 0;
    std::vector<int> elim_order;  // nodes in elimination order (for the back-substitution)
    for (;;) {
        const int nl = (int)act.size();
        // eliminated items first (they are the critical path), then the kept nodes' updates
        for (int pass = 0; pass < 2; ++pass)
            for (int p = 0; p < nl; ++p) {
                const bool el = elim_at(p, nl);
                if (el != (pass == 0)) continue;
                const int k = act[p];
                const int eL = prev_elim_left[p], eR = prev_elim_right[p];
                if (!el && eL < 0 && eR < 0) continue;  // nothing to do for a kept node
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
                if (nl == 1) continue;  // the last node of a cycle: no couplings
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
                if (!open_chain && nl == 2) {  // two-node cycle: both couplings join the same neighbour
                    e.kind |= BCR_MERGE;
                    e.left = -1;
                    e.cl_slot = -1;
                }
                einfo[k].wl = e.cl_slot; einfo[k].wr = e.cr_slot;
            }
        if (!open_chain && nl == 1) break;
        if (open_chain && nl == 2) {
            // the two pinned ends are left: export the coupling between them
            const Coup &c = coup[0];
            if (c.mode != 2) return;  // cannot happen with an interior (n >= 3)
            BcrItem &it = new_item(act[0], BCR_EXPORT, level);
            it.cl_mode = 2;
            it.cl_a = export_rows_left ? einfo[c.a].wl : einfo[c.a].wr;
            it.cl_b = export_rows_left ? einfo[c.a].wr : einfo[c.a].wl;
            it.cl_slot = export_tile;
            add_dep(it, einfo[c.a].item);
            break;
        }
        // next level: the survivors in order; the coupling to the next survivor is a product over the eliminated node in
        // between, or carried over when the two were already neighbours
        std::vector<int> act2, pel, per;
        std::vector<Coup> coup2;
        for (int p = 0; p < nl; ++p) {
            if (elim_at(p, nl)) continue;
            act2.push_back(act[p]);
            const bool has_right = !open_chain || p + 1 < nl;
            const int pr = (p + 1) % nl, pl = (p - 1 + nl) % nl;
            const bool has_left = !open_chain || p >= 1;
            int l = -1, r = -1;
            if (has_right) {
                if (elim_at(pr, nl)) { r = act[pr]; coup2.push_back({2, act[pr], 0, -1}); }
                else coup2.push_back(coup[p]);
            }
            if (has_left && elim_at(pl, nl)) l = act[pl];
            pel.push_back(l); per.push_back(r);
        }
        if (!open_chain && act2.size() == 1) coup2.clear();
        act.swap(act2); coup.swap(coup2);
        prev_elim_left.swap(pel); prev_elim_right.swap(per);
        ++level;
    }
    Y.n_levels = level + 1;
    Y.n_elim_items = (int)Y.items.size();
    Y.n_slots = next_slot;
    // ---- back-substitution: reverse elimination order (the pinned ends of an open chain get their x from outside)
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
    if (dep_overflow) return;
    // every dependency must point backwards (deadlock freedom of the in-order work queue)
    for (size_t q = 0; q < Y.items.size(); ++q)
        for (int d : Y.items[q].dep)
            if (d >= (int)q) return;
    Y.ok = true;
}

struct BcrPlan {
    bool ok = false;
    int nb = 0;          // pose blocks of S
    int n = 0;           // nodes
    int w = 0;           // half bandwidth in pose blocks
    int mb = 0, M = 0;   // pose blocks per node (padded), node dimension 6 * mb
    int ld = 0;          // row stride of a tile (doubles): M, or M + 4 where M would put the rows of a 4-row operand fragment on the
                         // same shared-memory banks (the DMMA fragment loads read 4 rows x 8 columns); a tile is M x ld
    int n_slots = 0;     // M x ld tiles in the pool: [0, n) = D / U of the nodes, then couplings / W tiles
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

// node partition of the pose blocks: plan.ok = false when the pattern is not a (cyclic) block band narrow enough
inline void bcr_partition(int nb, const std::vector<int> &rowptr, const std::vector<int> &col, BcrPlan &Y) {
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
    mb += mb & 1;  // even: the tile dimension 6 mb is a multiple of 4
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
    Y.ok = true;
}

// rowptr/col: symmetric 6x6 BSR pattern (both triangles, diagonal present).  Returns plan.ok = false when the pattern is
// not a (cyclic) block band narrow enough for BCR_MAX_M.
inline void bcr_plan(int nb, const std::vector<int> &rowptr, const std::vector<int> &col, BcrPlan &Y) {
    bcr_partition(nb, rowptr, col, Y);
    if (!Y.ok) return;
    Y.ok = false;
    const int n = Y.n, M = Y.ld;  // M: row stride of the tiles
    const long long MM = (long long)Y.M * Y.ld;
    BcrSched S;
    bcr_schedule(n, false, false, -1, S);
    if (!S.ok) return;
    Y.items.swap(S.items); Y.level_of_item.swap(S.level_of_item);
    Y.n_levels = S.n_levels; Y.n_elim_items = S.n_elim_items; Y.n_slots = S.n_slots;
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
            if (S.lvl0_rows[p] != a) continue;              // the twin block (j, i) is the one stored
            Y.dst[k] = (long long)(n + p) * MM + (long long)(6 * Y.blk_loc[i]) * M + 6 * Y.blk_loc[j];
        }
    Y.ok = true;
}

// ---- multi-GPU: rank r of W owns the nodes [lo_r, lo_{r+1}) of the cycle.  Its first node is an INTERFACE node; the rest,
// together with the next rank's interface node, is an open chain it eliminates on its own (bcr_schedule, open chain).
// What is left is a cyclic system over the W interface nodes, summed over the ranks and solved redundantly by everybody.
struct BcrDistPlan {
    bool ok = false;
    int rank = 0, world = 1;
    int lo = 0, m = 0;                 // first owned node, number of owned nodes; local node j = global node (lo + j) mod n, j = 0..m
    BcrSched local, iface;             // local open chain over m + 1 nodes ; cyclic interface system over `world` nodes
    std::vector<long long> dst;        // [nnzb] loader map into the LOCAL pool (lambda only on owned nodes)
    std::vector<int> blk_lnode;        // [nb] local node of every pose block, -1 when outside [lo, lo + m]
    std::vector<int> node_lo;          // [world + 1] first node of every rank
};
inline int bcr_rank_lo(int n, int world, int r) { return (int)((long long)n * r / world); }
inline void bcr_dist_plan(const BcrPlan &P, const std::vector<int> &rowptr, const std::vector<int> &col, int rank, int world, BcrDistPlan &D) {
    D = BcrDistPlan();
    D.rank = rank; D.world = world;
    const int n = P.n;
    if (!P.ok || world < 2 || n < 2 * world) return;  // every rank needs its interface node and at least one interior node
    D.node_lo.resize(world + 1);
    for (int r = 0; r <= world; ++r) D.node_lo[r] = bcr_rank_lo(n, world, r);
    D.lo = D.node_lo[rank]; D.m = D.node_lo[rank + 1] - D.lo;
    const int m = D.m, M = P.ld;
    const long long MM = (long long)P.M * P.ld;
    bcr_schedule(world, false, false, -1, D.iface);
    if (!D.iface.ok) return;
    // the exported coupling is interface coupling `rank` (between interface nodes rank and rank + 1): rows as its loader wants
    const bool rows_left = D.iface.lvl0_rows[rank] == rank;
    bcr_schedule(m + 1, true, rows_left, world + rank, D.local);
    if (!D.local.ok) return;
    D.blk_lnode.assign(P.nb, -1);
    for (int i = 0; i < P.nb; ++i) {
        if (P.blk_node[i] < 0) continue;
        const int j = (P.blk_node[i] - D.lo + n) % n;
        if (j <= m) D.blk_lnode[i] = j;
    }
    D.dst.assign(col.size(), -1);
    for (int i = 0; i < P.nb; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];
            const int a = D.blk_lnode[i], b = D.blk_lnode[j];
            if (a < 0 || b < 0) continue;
            if (a == b) {
                // lambda belongs to the owner of a node: the next rank's interface node (local m) only collects this rank's share
                D.dst[k] = ((long long)a * MM + (long long)(6 * P.blk_loc[i]) * M + 6 * P.blk_loc[j]) | ((i == j && a < m) ? BCR_DST_DIAG : 0);
                continue;
            }
            if (std::abs(a - b) != 1) continue;   // (0, m) can only be adjacent on a 2-node ring, excluded by m >= 2
            const int p = std::min(a, b);          // local coupling between local nodes p and p+1
            if (D.local.lvl0_rows[p] != a) continue;
            D.dst[k] = (long long)(m + 1 + p) * MM + (long long)(6 * P.blk_loc[i]) * M + 6 * P.blk_loc[j];
        }
    D.ok = true;
}
