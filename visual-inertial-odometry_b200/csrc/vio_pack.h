// vio_pack.h — host-side graph packer: vio_graph (caller's arrays) -> landmark-sorted SoA + ordering +
// reduced-system sparsity pattern.  Pure C++ (no CUDA) so the same code feeds the device upload in
// vio_set_graph and the CPU emulation harness in tests/.  Mirrors Problem::SetOrdering
// (/root/reference/workspace/assignments/17-vins-initialization/vins-mono/src/backend/problem.cc:256-285):
// pose-class vertices get consecutive offsets in creation order, landmarks follow.
#pragma once
#include <chrono>
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <utility>
#include <string>
#include <thread>
#include <vector>
#include "../../include/vio_b200.h"

#include "vio_bcr.h"

// std::vector whose resize() leaves new elements uninitialised: the per-edge arrays (hundreds of MB at 10M edges) are
// written exactly once by the packer's worker threads, which then also take their first-touch page faults in parallel
// instead of one thread zero-filling everything first (that alone was a quarter of vio_set_graph).
template <class T>
struct pack_alloc : std::allocator<T> {
    template <class U> struct rebind { using other = pack_alloc<U>; };
    pack_alloc() = default;
    template <class U> pack_alloc(const pack_alloc<U> &) {}
    template <class U, class... A>
    void construct(U *q, A &&...a) {
        if constexpr (sizeof...(A) == 0) ::new ((void *)q) U;
        else ::new ((void *)q) U(std::forward<A>(a)...);
    }
};
template <class T> using pvec = std::vector<T, pack_alloc<T>>;

struct PackedGraph {
    int C = 0, NSB = 0, NB = 0, P = 0, L = 0, Lglobal = 0, storage = 1;
    bool ext_free = false;    // the extrinsic VertexPose is being estimated (4-vertex EdgeReprojection, 4th Jacobian)
    int batch = 1, Pper = 0;  // lock-step batch: `batch` stacked Pper x Pper reduced systems (dense), P = batch * Pper
    long long E = 0, nnzb = 0;
    size_t s_count = 0;
    std::vector<int> pose_off, sb_off, pose_blk, blk_off, blk_dim;
    std::vector<uint8_t> blk_fixed, pose_fixed, sb_fixed, row_fixed;
    double qic[4], tic[3];
    pvec<int> lm_global, lm_host, lm_eptr, e_pose_j;
    std::vector<uint8_t> lm_fixed, pt_fixed;  // empty = none fixed
    pvec<double> pix, piy, piz, pjx, pjy, invd;
    std::vector<int> rowptr, col, tr, diag;
    // VertexPointXYZ landmarks (caller order) and their EdgeReprojectionXYZ observations, CSR by point
    int Lx = 0;
    long long Ex = 0;
    std::vector<int> px_eptr, ex_pose;
    std::vector<double> ex_ox, ex_oy;
    // landmark groups for the grouped linearise kernel (vio_grouped.cuh)
    bool grouped_ok = false;
    int n_groups = 0, group_threads = 0;
    size_t group_smem_max = 0;
    std::vector<int> g_hdr;        // 8 ints per group: host ns lm0 nlm ell0 pair0 slot0 pad
    std::vector<int> g_slot_pose;  // per group ns entries (entry 0 = host)
    std::vector<long long> g_pairinfo;
    pvec<double> ell_pjx, ell_pjy;
    pvec<int> ell_edge;
    // multi-GPU, node-range sharding (see pack_graph): this rank keeps the landmarks hosted by the cameras of ITS nodes of the
    // block-cyclic-reduction partition, so its share of S stays inside its own nodes and the next rank's interface node
    bool shard_by_node = false;
    std::vector<int> se3_keep;   // SE3 priors this rank accumulates (indices into the caller's arrays)
};

inline int pack_fail(std::string &err, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
}

// Host threads for the O(E) passes of the packer (VIO_B200_PACK_THREADS overrides; small graphs stay on the caller's thread:
// the lock-step batch path already packs one problem per thread).
inline int pack_threads(long long work) {
    if (work < (1 << 18)) return 1;
    int t = (int)std::thread::hardware_concurrency();
    if (const char *ev = getenv("VIO_B200_PACK_THREADS")) t = atoi(ev);
    return std::max(1, std::min(t, 16));
}
template <class F>
inline void pack_parallel(int nthreads, F &&body) {  // body(thread index)
    if (nthreads <= 1) { body(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back([&body, t] { body(t); });
    body(0);
    for (auto &x : th) x.join();
}

// co-visibility pattern of the inverse-depth landmarks over the pose blocks (+ diagonal): rows[a] = blocks coupled with a
template <class EO>
inline void covis_rows_range(const vio_graph *g, const std::vector<int> &cnt, const EO &eorder, const std::vector<int> &pose_blk, int NB,
                             int l_begin, int l_end, std::vector<std::vector<int>> &rows) {
    rows.assign(NB, std::vector<int>());
    auto add = [&](int a, int b) {
        auto &r = rows[a];
        if (std::find(r.begin(), r.end(), b) == r.end()) r.push_back(b);
    };
    std::vector<int> set, last;
    for (int l = l_begin; l < l_end; ++l) {
        set.clear();
        if (cnt[l] == cnt[l + 1]) continue;
        set.push_back(pose_blk[g->rp_pose_i[eorder[cnt[l]]]]);
        for (int k = cnt[l]; k < cnt[l + 1]; ++k) set.push_back(pose_blk[g->rp_pose_j[eorder[k]]]);
        if (set == last) continue;
        for (size_t a = 0; a < set.size(); ++a)
            for (size_t b = 0; b < set.size(); ++b) add(set[a], set[b]);
        last = set;
    }
}
// landmark ranges on the packer's host threads, then the union per block row (+ the diagonal)
template <class EO>
inline void covis_rows(const vio_graph *g, const std::vector<int> &cnt, const EO &eorder, const std::vector<int> &pose_blk, int NB,
                       int Lg, std::vector<std::vector<int>> &rows) {
    const int nth = pack_threads((long long)cnt[Lg]);
    std::vector<std::vector<std::vector<int>>> part(nth);
    pack_parallel(nth, [&](int t) {
        covis_rows_range(g, cnt, eorder, pose_blk, NB, (int)((long long)Lg * t / nth), (int)((long long)Lg * (t + 1) / nth), part[t]);
    });
    rows.assign(NB, std::vector<int>());
    pack_parallel(nth, [&](int t) {
        for (int k = (int)((long long)NB * t / nth); k < (int)((long long)NB * (t + 1) / nth); ++k) {
            auto &r = rows[k];
            r.push_back(k);
            for (int u = 0; u < nth; ++u)
                for (int b : part[u][k])
                    if (std::find(r.begin(), r.end(), b) == r.end()) r.push_back(b);
        }
    });
}
template <class EO>
inline void covis_pattern(const vio_graph *g, const std::vector<int> &cnt, const EO &eorder, const std::vector<int> &pose_blk, int NB,
                          int Lg, std::vector<int> &rowptr, std::vector<int> &col) {
    std::vector<std::vector<int>> rows;
    covis_rows(g, cnt, eorder, pose_blk, NB, Lg, rows);
    rowptr.assign(NB + 1, 0);
    col.clear();
    for (int k = 0; k < NB; ++k) {
        std::sort(rows[k].begin(), rows[k].end());
        col.insert(col.end(), rows[k].begin(), rows[k].end());
        rowptr[k + 1] = (int)col.size();
    }
}

struct PackTimer {  // VIO_B200_PACK_PROFILE=1: phase times of pack_graph on stderr
    bool on;
    std::chrono::steady_clock::time_point t0;
    PackTimer() : on(getenv("VIO_B200_PACK_PROFILE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[vio_b200 pack] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

inline int pack_graph(const vio_graph *g, int shard_rank, int shard_world, PackedGraph &K, std::string &err, int batch = 1) {
    PackTimer ptm;
    const int C = g->n_pose, NSB = g->n_speedbias, Lg = g->n_landmark;
    const long long Eg = g->n_reproj;
    if (C < 0 || NSB < 0 || Lg < 0 || Eg < 0) return pack_fail(err, VIO_ERR_INVALID, "negative size");
    if (C > 0 && !g->pose) return pack_fail(err, VIO_ERR_INVALID, "pose array missing");
    if (Eg > 0 && (!g->rp_landmark || !g->rp_pose_i || !g->rp_pose_j || !g->rp_pts_i || !g->rp_pts_j))
        return pack_fail(err, VIO_ERR_INVALID, "reprojection arrays missing");
    if (Eg > 0x7fffffffLL) return pack_fail(err, VIO_ERR_UNSUPPORTED, "more than 2^31 edges per shard");
    if (Lg > 0 && !g->inv_depth) return pack_fail(err, VIO_ERR_INVALID, "inv_depth array missing");
    if (NSB > 0 && !g->speedbias) return pack_fail(err, VIO_ERR_INVALID, "speedbias array missing");
    if (g->n_se3prior < 0 || g->n_imu < 0) return pack_fail(err, VIO_ERR_INVALID, "negative size");
    if (g->n_se3prior > 0) {
        if (!g->sp_pose || !g->sp_p || !g->sp_q || !g->sp_info) return pack_fail(err, VIO_ERR_INVALID, "EdgeSE3Prior arrays missing");
        for (int k = 0; k < g->n_se3prior; ++k)
            if (g->sp_pose[k] < 0 || g->sp_pose[k] >= C) return pack_fail(err, VIO_ERR_INVALID, "se3 prior %d: pose out of range", k);
    }
    if (g->n_imu > 0 && (!g->imu_pose_i || !g->imu_pose_j || !g->imu_sb_i || !g->imu_sb_j || !g->imu_sum_dt || !g->imu_delta_p ||
                         !g->imu_delta_q || !g->imu_delta_v || !g->imu_lin_ba || !g->imu_lin_bg || !g->imu_jacobian ||
                         !g->imu_covariance))
        return pack_fail(err, VIO_ERR_INVALID, "EdgeImu arrays missing");
    // ---- ordering of the pose-class vertices (reference SetOrdering) -------------------------
    const int NB = C + NSB;
    std::vector<int> &pose_off = K.pose_off, &sb_off = K.sb_off, &pose_blk = K.pose_blk, &blk_off = K.blk_off, &blk_dim = K.blk_dim;
    std::vector<uint8_t> &blk_fixed = K.blk_fixed;
    pose_off.assign(C, -1); sb_off.assign(NSB, -1); pose_blk.assign(C, -1); blk_off.assign(NB, 0); blk_dim.assign(NB, 0); blk_fixed.assign(NB, 0);
    int P = 0;
    for (int k = 0; k < NB; ++k) {
        int ent = g->pclass_order ? g->pclass_order[k] : (k < C ? k : ~(k - C));
        if (ent >= 0) {
            if (ent >= C || pose_off[ent] >= 0) return pack_fail(err, VIO_ERR_INVALID, "bad pclass_order entry %d", k);
            pose_off[ent] = P; pose_blk[ent] = k; blk_off[k] = P; blk_dim[k] = 6;
            blk_fixed[k] = g->pose_fixed ? g->pose_fixed[ent] : 0;
            P += 6;
        } else {
            int i = ~ent;
            if (i >= NSB || sb_off[i] >= 0) return pack_fail(err, VIO_ERR_INVALID, "bad pclass_order entry %d", k);
            sb_off[i] = P; blk_off[k] = P; blk_dim[k] = 9;
            blk_fixed[k] = g->speedbias_fixed ? g->speedbias_fixed[i] : 0;
            P += 9;
        }
    }
    // ---- extrinsics -----------------------------------------------------------------------------
    K.ext_free = false;
    double *qic = K.qic, *tic = K.tic;
    if (g->ext_pose >= 0) {
        if (g->ext_pose >= C) return pack_fail(err, VIO_ERR_INVALID, "ext_pose out of range");
        // a FREE extrinsic vertex (ESTIMATE_EXTRINSIC=1) contributes the 4th Jacobian of every EdgeReprojection and couples
        // with every landmark: handled by the per-landmark kernel on unsharded, single-problem, dense handles
        K.ext_free = !(g->pose_fixed && g->pose_fixed[g->ext_pose]);
        if (K.ext_free && (shard_world > 1 || batch > 1))
            return pack_fail(err, VIO_ERR_UNSUPPORTED, "a free extrinsic vertex is not supported with sharding / lock-step batches");
        if (K.ext_free && (g->n_point > 0 || g->n_reproj_xyz > 0))
            return pack_fail(err, VIO_ERR_UNSUPPORTED, "a free extrinsic vertex cannot be combined with EdgeReprojectionXYZ (constant extrinsics)");
        const double *e = g->pose + 7 * (size_t)g->ext_pose;
        tic[0] = e[0]; tic[1] = e[1]; tic[2] = e[2];
        qic[0] = e[3]; qic[1] = e[4]; qic[2] = e[5]; qic[3] = e[6];
    } else {
        for (int k = 0; k < 4; ++k) qic[k] = g->q_ic[k];
        for (int k = 0; k < 3; ++k) tic[k] = g->t_ic[k];
    }
    // ---- landmark-sorted edge CSR (global), checks -------------------------------------------
    // (drivers add the edges of a landmark back to back, usually landmark after landmark: then rp_landmark is already
    // non-decreasing, the CSR offsets are the positions where it changes and the edge order is the identity)
    std::vector<int> cnt(Lg + 1, 0);
    pvec<int> eorder;
    eorder.resize(Eg);
    {
        const int nth = pack_threads(Eg);
        std::vector<long long> bad_l(nth, -1), bad_p(nth, -1);
        std::vector<char> sorted_t(nth, 1);
        pack_parallel(nth, [&](int t) {
            const long long ea = Eg * t / nth, eb = Eg * (t + 1) / nth;
            int prev = ea > 0 ? g->rp_landmark[ea - 1] : 0;
            for (long long e = ea; e < eb; ++e) {
                const int l = g->rp_landmark[e];
                if (l < 0 || l >= Lg) { if (bad_l[t] < 0) bad_l[t] = e; continue; }
                const int a = g->rp_pose_i[e], b = g->rp_pose_j[e];
                if (a < 0 || a >= C || b < 0 || b >= C) { if (bad_p[t] < 0) bad_p[t] = e; }
                if (l < prev) sorted_t[t] = 0;
                prev = l;
                eorder[e] = (int)e;
            }
        });
        bool sorted = true;
        for (int t = 0; t < nth; ++t) {
            if (bad_l[t] >= 0) return pack_fail(err, VIO_ERR_INVALID, "edge %lld: landmark out of range", bad_l[t]);
            if (bad_p[t] >= 0) return pack_fail(err, VIO_ERR_INVALID, "edge %lld: pose out of range", bad_p[t]);
            sorted = sorted && sorted_t[t];
        }
        if (sorted) {
            int l = 0;  // cnt[k] = first edge of landmark k
            for (long long e = 0; e < Eg; ++e) {
                const int le = g->rp_landmark[e];
                while (l < le) cnt[++l] = (int)e;
            }
            while (l < Lg) cnt[++l] = (int)Eg;
        } else {
            for (long long e = 0; e < Eg; ++e) cnt[g->rp_landmark[e] + 1]++;
            for (int l = 0; l < Lg; ++l) cnt[l + 1] += cnt[l];
            std::vector<int> cur(cnt.begin(), cnt.end() - 1);
            for (long long e = 0; e < Eg; ++e) eorder[cur[g->rp_landmark[e]]++] = (int)e;
        }
    }
    ptm.mark("validate + edge CSR order");
    // ---- which landmarks this rank keeps
    // Legacy sharding: contiguous landmark ranges balanced by edge count; the whole reduced system is all-reduced.
    // Node-range sharding (chosen when the reduced system is block-sparse, a cyclic block band the block cyclic reduction
    // covers, with at least two nodes per rank, and every landmark is observed only from its host's node and the node
    // after it): rank r keeps the landmarks hosted in its nodes; nothing outside its own nodes and the next rank's first
    // node is touched, so only the W-node interface system and the pose update cross the ranks.
    std::vector<int> lsel;
    K.shard_by_node = false;
    K.se3_keep.clear();
    for (int k = 0; k < g->n_se3prior; ++k)
        if (shard_rank == 0) K.se3_keep.push_back(k);
    if (shard_world > 1) {
        int st0 = g->storage;
        if (st0 == VIO_STORAGE_AUTO) st0 = (NSB == 0 && P > 2048 && !K.ext_free) ? VIO_STORAGE_BSR : VIO_STORAGE_DENSE;
        const bool try_nodes = st0 == VIO_STORAGE_BSR && NSB == 0 && g->n_imu == 0 && g->n_point == 0 && g->n_reproj_xyz == 0 && batch == 1 &&
                               getenv("VIO_B200_SHARD_LEGACY") == nullptr;
        if (try_nodes) {
            std::vector<int> rp0, cl0;
            covis_pattern(g, cnt, eorder, pose_blk, NB, Lg, rp0, cl0);
            BcrPlan P0;
            bcr_partition(NB, rp0, cl0, P0);
            if (P0.ok && P0.n >= 2 * shard_world) {
                const int n = P0.n;
                auto owner = [&](int node) {
                    int r = (int)(((long long)node * shard_world + shard_world - 1) / n);  // smallest r with lo_r > node, minus one
                    while (r > 0 && bcr_rank_lo(n, shard_world, r) > node) --r;
                    while (r + 1 < shard_world && bcr_rank_lo(n, shard_world, r + 1) <= node) ++r;
                    return r;
                };
                bool local_ok = true;
                for (int l = 0; l < Lg && local_ok; ++l) {
                    if (cnt[l] == cnt[l + 1]) continue;
                    const int a = P0.blk_node[pose_blk[g->rp_pose_i[eorder[cnt[l]]]]];
                    if (a < 0) { local_ok = false; break; }
                    for (int k = cnt[l]; k < cnt[l + 1]; ++k) {
                        const int b = P0.blk_node[pose_blk[g->rp_pose_j[eorder[k]]]];
                        if (b < 0 || (b - a + n) % n > 1) { local_ok = false; break; }
                    }
                }
                for (int k = 0; k < g->n_se3prior && local_ok; ++k)
                    if (P0.blk_node[pose_blk[g->sp_pose[k]]] < 0 && !(g->pose_fixed && g->pose_fixed[g->sp_pose[k]])) local_ok = false;
                if (local_ok) {
                    K.shard_by_node = true;
                    for (int l = 0; l < Lg; ++l) {
                        if (cnt[l] == cnt[l + 1]) { if (shard_rank == 0) lsel.push_back(l); continue; }
                        if (owner(P0.blk_node[pose_blk[g->rp_pose_i[eorder[cnt[l]]]]]) == shard_rank) lsel.push_back(l);
                    }
                    K.se3_keep.clear();
                    for (int k = 0; k < g->n_se3prior; ++k) {
                        const int nd = P0.blk_node[pose_blk[g->sp_pose[k]]];
                        if ((nd < 0 ? 0 : owner(nd)) == shard_rank) K.se3_keep.push_back(k);
                    }
                }
            }
        }
    }
    if (!K.shard_by_node) {
        int l_begin = 0, l_end = Lg;
        if (shard_world > 1) {
            auto cut = [&](int r) -> int {
                if (r <= 0) return 0;
                if (r >= shard_world) return Lg;
                long long target = Eg * r / shard_world;
                return (int)(std::lower_bound(cnt.begin(), cnt.end(), (int)target) - cnt.begin());
            };
            l_begin = std::min(cut(shard_rank), Lg);
            l_end = std::min(cut(shard_rank + 1), Lg);
            if (l_end < l_begin) l_end = l_begin;
        }
        lsel.resize(l_end - l_begin);
        for (int ll = 0; ll < l_end - l_begin; ++ll) lsel[ll] = l_begin + ll;
    }
    const int L = (int)lsel.size();
    long long E = 0;
    for (int l : lsel) E += cnt[l + 1] - cnt[l];
    pvec<int> &lm_host = K.lm_host, &lm_eptr = K.lm_eptr, &e_pose_j = K.e_pose_j;
    pvec<double> &pix = K.pix, &piy = K.piy, &piz = K.piz, &pjx = K.pjx, &pjy = K.pjy, &invd = K.invd;
    // uninitialised: every entry is written below (landmarks without edges get the defaults there)
    lm_host.resize(L); lm_eptr.resize((size_t)L + 1); e_pose_j.resize(E);
    pix.resize(L); piy.resize(L); piz.resize(L); pjx.resize(E); pjy.resize(E); invd.resize(L);
    K.lm_global.resize(L);
    ptm.mark("  (allocate SoA)");
    {
        bool any = false;
        if (g->landmark_fixed)
            for (int l = 0; l < Lg && !any; ++l) any = g->landmark_fixed[l] != 0;
        K.lm_fixed.assign(any ? (size_t)L : 0, 0);
    }
    // local landmark order: sorted by host pose (stable) so that landmarks sharing a host are contiguous and
    // can be grouped; landmarks without edges go last.
    std::vector<int> lorder(L);
    for (int ll = 0; ll < L; ++ll) lorder[ll] = lsel[ll];
    auto host_of = [&](int l) { return cnt[l] == cnt[l + 1] ? 0x7fffffff : g->rp_pose_i[eorder[cnt[l]]]; };
    {
        // scenes are usually generated host by host: skip the sort when the order is already non-decreasing
        std::vector<int> hkey(L);
        for (int ll = 0; ll < L; ++ll) hkey[ll] = host_of(lsel[ll]);
        if (!std::is_sorted(hkey.begin(), hkey.end())) {
            std::vector<int> idx(L);
            for (int ll = 0; ll < L; ++ll) idx[ll] = ll;
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return hkey[a] < hkey[b]; });
            for (int ll = 0; ll < L; ++ll) lorder[ll] = lsel[idx[ll]];
        }
    }
    ptm.mark("  (landmark order)");
    {
        long long ecur = 0;
        for (int ll = 0; ll < L; ++ll) { lm_eptr[ll] = (int)ecur; ecur += cnt[lorder[ll] + 1] - cnt[lorder[ll]]; }
    }
    {
        const int nth = pack_threads(E);
        std::vector<int> bad(nth, -1);
        pack_parallel(nth, [&](int t) {
            const int la = (int)((long long)L * t / nth), lb = (int)((long long)L * (t + 1) / nth);
            for (int ll = la; ll < lb; ++ll) {
                const int l = lorder[ll];
                K.lm_global[ll] = l;
                invd[ll] = g->inv_depth[l];
                if (!K.lm_fixed.empty()) K.lm_fixed[ll] = g->landmark_fixed[l] ? 1 : 0;
                int le = lm_eptr[ll];
                if (cnt[l] == cnt[l + 1]) { lm_host[ll] = 0; pix[ll] = 0.0; piy[ll] = 0.0; piz[ll] = 1.0; }
                for (int k = cnt[l]; k < cnt[l + 1]; ++k, ++le) {
                    const int e = eorder[k];
                    if (k == cnt[l]) {
                        lm_host[ll] = g->rp_pose_i[e];
                        pix[ll] = g->rp_pts_i[3 * (size_t)e]; piy[ll] = g->rp_pts_i[3 * (size_t)e + 1]; piz[ll] = g->rp_pts_i[3 * (size_t)e + 2];
                    } else if (g->rp_pose_i[e] != lm_host[ll] || g->rp_pts_i[3 * (size_t)e] != pix[ll] ||
                               g->rp_pts_i[3 * (size_t)e + 1] != piy[ll] || g->rp_pts_i[3 * (size_t)e + 2] != piz[ll]) {
                        if (bad[t] < 0) bad[t] = l;
                    }
                    e_pose_j[le] = g->rp_pose_j[e];
                    pjx[le] = g->rp_pts_j[2 * (size_t)e];
                    pjy[le] = g->rp_pts_j[2 * (size_t)e + 1];
                }
            }
        });
        for (int t = 0; t < nth; ++t)
            if (bad[t] >= 0)
                return pack_fail(err, VIO_ERR_UNSUPPORTED,
                                 "landmark %d: edges disagree on host pose / host observation (see vio_b200.h preconditions)", bad[t]);
    }
    lm_eptr[L] = (int)E;
    ptm.mark("landmark-sorted SoA copy");

    // ---- VertexPointXYZ observations, CSR by point ---------------------------------------------------
    const int Lx = g->n_point;
    const long long Ex = g->n_reproj_xyz;
    if (Lx < 0 || Ex < 0) return pack_fail(err, VIO_ERR_INVALID, "negative size");
    if (Ex > 0 && (!g->rx_point || !g->rx_pose || !g->rx_obs)) return pack_fail(err, VIO_ERR_INVALID, "EdgeReprojectionXYZ arrays missing");
    if (Lx > 0 && !g->point_xyz) return pack_fail(err, VIO_ERR_INVALID, "point_xyz missing");
    if ((Lx > 0 || Ex > 0) && (shard_world > 1 || batch > 1))
        return pack_fail(err, VIO_ERR_UNSUPPORTED, "VertexPointXYZ landmarks are not supported with sharding / lock-step batches");
    if (Ex > 0x7fffffffLL) return pack_fail(err, VIO_ERR_UNSUPPORTED, "more than 2^31 EdgeReprojectionXYZ edges");
    K.Lx = Lx; K.Ex = Ex;
    K.pt_fixed.clear();
    if (g->point_fixed)
        for (int l = 0; l < Lx; ++l)
            if (g->point_fixed[l]) { K.pt_fixed.assign(g->point_fixed, g->point_fixed + Lx); break; }
    K.px_eptr.assign((size_t)Lx + 1, 0); K.ex_pose.assign(Ex, 0); K.ex_ox.assign(Ex, 0.0); K.ex_oy.assign(Ex, 0.0);
    {
        for (long long e = 0; e < Ex; ++e) {
            const int l = g->rx_point[e], a = g->rx_pose[e];
            if (l < 0 || l >= Lx) return pack_fail(err, VIO_ERR_INVALID, "xyz edge %lld: point out of range", e);
            if (a < 0 || a >= C) return pack_fail(err, VIO_ERR_INVALID, "xyz edge %lld: pose out of range", e);
            K.px_eptr[l + 1]++;
        }
        for (int l = 0; l < Lx; ++l) K.px_eptr[l + 1] += K.px_eptr[l];
        std::vector<int> cur(K.px_eptr.begin(), K.px_eptr.end() - 1);
        for (long long e = 0; e < Ex; ++e) {
            const int k = cur[g->rx_point[e]]++;
            K.ex_pose[k] = g->rx_pose[e]; K.ex_ox[k] = g->rx_obs[2 * (size_t)e]; K.ex_oy[k] = g->rx_obs[2 * (size_t)e + 1];
        }
    }

    // ---- storage of the reduced system ----------------------------------------------------------
    int storage = g->storage;
    if (storage == VIO_STORAGE_AUTO) storage = (NSB == 0 && P > 2048 && !K.ext_free) ? VIO_STORAGE_BSR : VIO_STORAGE_DENSE;
    if (K.ext_free && storage != VIO_STORAGE_DENSE) return pack_fail(err, VIO_ERR_UNSUPPORTED, "a free extrinsic vertex needs dense storage");
    if (storage == VIO_STORAGE_BSR && (NSB != 0 || g->n_imu != 0))
        return pack_fail(err, VIO_ERR_UNSUPPORTED, "BSR storage supports 6-dof pose vertices only");
    size_t s_count;
    long long nnzb = 0;
    std::vector<int> &rowptr = K.rowptr, &col = K.col, &tr = K.tr, &diag = K.diag;
    rowptr.clear(); col.clear(); tr.clear(); diag.clear();
    if (batch < 1 || P % batch != 0) return pack_fail(err, VIO_ERR_INVALID, "batch %d does not divide P=%d", batch, P);
    const int Pper = P / batch;
    if (batch > 1 && storage != VIO_STORAGE_DENSE) return pack_fail(err, VIO_ERR_UNSUPPORTED, "lock-step batches use dense storage");
    if (storage == VIO_STORAGE_DENSE) {
        if ((size_t)P * Pper * sizeof(double) > (size_t)16 << 30) return pack_fail(err, VIO_ERR_UNSUPPORTED, "dense S too large");
        s_count = (size_t)P * Pper;
    } else {
        // global pattern (all shards must agree): co-visibility of every landmark + diagonal
        std::vector<std::vector<int>> rows;
        covis_rows(g, cnt, eorder, pose_blk, NB, Lg, rows);
        auto add = [&](int a, int b) {
            auto &r = rows[a];
            if (std::find(r.begin(), r.end(), b) == r.end()) r.push_back(b);
        };
        std::vector<int> set;
        for (int l = 0; l < Lx; ++l) {
            set.clear();
            for (int k = K.px_eptr[l]; k < K.px_eptr[l + 1]; ++k) set.push_back(pose_blk[K.ex_pose[k]]);
            for (size_t a = 0; a < set.size(); ++a)
                for (size_t b = 0; b < set.size(); ++b) add(set[a], set[b]);
        }
        rowptr.assign(NB + 1, 0);
        for (int k = 0; k < NB; ++k) {
            std::sort(rows[k].begin(), rows[k].end());
            rowptr[k + 1] = rowptr[k] + (int)rows[k].size();
        }
        nnzb = rowptr[NB];
        col.resize(nnzb); tr.resize(nnzb); diag.resize(NB);
        for (int k = 0; k < NB; ++k) std::copy(rows[k].begin(), rows[k].end(), col.begin() + rowptr[k]);
        auto find = [&](int a, int b) -> int {
            auto it = std::lower_bound(col.begin() + rowptr[a], col.begin() + rowptr[a + 1], b);
            return (int)(it - col.begin());
        };
        for (int a = 0; a < NB; ++a)
            for (int k = rowptr[a]; k < rowptr[a + 1]; ++k) {
                tr[k] = find(col[k], a);
                if (col[k] == a) diag[a] = k;
            }
        s_count = (size_t)nnzb * 36;
    }

    ptm.mark("storage pattern (BSR)");
    // ---- landmark groups (same host, <= VIO_PACK_NS_MAX pose slots, shared-memory budget) ---------------------
    {
        const int NS_MAX = 22;
        const size_t SMEM_BUDGET = 200 * 1024;
        int target_lm = 100;  // landmarks per CTA (VIO_B200_GROUP_LM overrides; tuning knob; 100 x 10 warps measured best)
        if (const char *ev = getenv("VIO_B200_GROUP_LM")) target_lm = std::max(1, atoi(ev));
        auto smem_bytes = [](int ns, int nlm) -> size_t {
            const size_t npairs = (size_t)ns * (ns + 1) / 2;
            const size_t dbl = (size_t)ns * 12 + (size_t)nlm * 17 + 1 + (size_t)nlm * ns * 6 + (size_t)nlm * (((ns - 1) * 9) | 1) +
                               2 * (size_t)nlm + 1 + (size_t)ns * 48 + npairs * 36 + (size_t)ns * 18 + (size_t)ns * 9 + npairs;
            return dbl * 8 + (3 * (size_t)ns + 2 * npairs) * 4;
        };
        auto block_off = [&](int pa, int pb, long long &off) -> bool {  // element offset of block (pa,pb) in S storage
            if (storage == VIO_STORAGE_DENSE) {
                if (pose_off[pa] / Pper != pose_off[pb] / Pper) return false;  // blocks never couple two problems of a batch
                off = (long long)pose_off[pa] * Pper + pose_off[pb] % Pper;
                return true;
            }
            const int ra = pose_blk[pa], cb = pose_blk[pb];
            auto it = std::lower_bound(col.begin() + rowptr[ra], col.begin() + rowptr[ra + 1], cb);
            if (it == col.begin() + rowptr[ra + 1] || *it != cb) return false;
            off = 36LL * (it - col.begin());
            return true;
        };
        // Groups never span two hosts, so the landmark range is cut at host boundaries into one chunk per host thread.
        // Pass 1 (per chunk): group boundaries, slot tables, pair tables, ELL sizes.  The chunks' tables are concatenated
        // (small), which fixes every group's offset into the ELL arrays.  Pass 2 (per chunk): the ELL tiles are written
        // straight into the final arrays - no per-chunk copies of the per-edge data, first-touch page faults in parallel.
        struct GroupPart {
            std::vector<int> g_hdr, g_slot_pose;
            std::vector<long long> g_pairinfo;
            size_t ell = 0;
            bool ok = true;
            int n_groups = 0, max_obs_slots = 0;
            size_t smem_max = 0;
        };
        auto plan = [&](int lbeg, int lend, GroupPart &O) {
            {
                const size_t ng = (size_t)(lend - lbeg) / (size_t)std::max(1, target_lm / 2) + 16;
                O.g_hdr.reserve(8 * ng); O.g_slot_pose.reserve(NS_MAX * ng / 2); O.g_pairinfo.reserve(ng * 70);
            }
            int l0 = lbeg;
            std::vector<int> slots;  // slot -> pose (slot 0 = host)
            std::vector<int> slot_of_pose(C, -1);
            std::vector<int> seen_stamp(C, 0);  // evaluation (a landmark may be tried twice: last of a run, first of the next) that last listed this pose
            int stamp = 0;
            while (l0 < lend && O.ok) {
                if (lm_eptr[l0] == lm_eptr[l0 + 1]) break;  // only edge-less landmarks remain
                const int host = lm_host[l0];
                slots.assign(1, host);
                slot_of_pose[host] = 0;
                int l1 = l0;
                while (l1 < lend && lm_eptr[l1] != lm_eptr[l1 + 1] && lm_host[l1] == host && (l1 - l0) < 128) {
                    // would this landmark fit?
                    ++stamp;
                    int added = 0;
                    bool bad = false;
                    for (int e = lm_eptr[l1]; e < lm_eptr[l1 + 1]; ++e) {
                        const int pj = e_pose_j[e];
                        if (pj == host || seen_stamp[pj] == stamp) { bad = true; break; }  // observer == host, or seen twice
                        seen_stamp[pj] = stamp;
                        if (slot_of_pose[pj] < 0) { slot_of_pose[pj] = (int)slots.size(); slots.push_back(pj); ++added; }
                    }
                    if (bad) { O.ok = false; break; }
                    const int ns_new = (int)slots.size();
                    if (ns_new > NS_MAX || smem_bytes(ns_new, l1 - l0 + 1) > SMEM_BUDGET) {
                        // undo this landmark's slots and close the group (a single landmark that does not fit: irregular)
                        for (int k = 0; k < added; ++k) { slot_of_pose[slots.back()] = -1; slots.pop_back(); }
                        if (l1 == l0) O.ok = false;
                        break;
                    }
                    ++l1;
                }
                for (int p2 : slots) slot_of_pose[p2] = -1;
                if (!O.ok) break;
                // split the feasible run [l0, l1) evenly into chunks of about `target_lm` landmarks
                const int nrun = l1 - l0;
                const int nchunk = (nrun + target_lm - 1) / target_lm;
                for (int ch = 0; ch < nchunk && O.ok; ++ch) {
                    const int la = l0 + (int)((long long)nrun * ch / nchunk), lb = l0 + (int)((long long)nrun * (ch + 1) / nchunk);
                    slots.assign(1, host);
                    slot_of_pose[host] = 0;
                    for (int l = la; l < lb; ++l)
                        for (int e = lm_eptr[l]; e < lm_eptr[l + 1]; ++e) {
                            const int pj = e_pose_j[e];
                            if (slot_of_pose[pj] < 0) { slot_of_pose[pj] = (int)slots.size(); slots.push_back(pj); }
                        }
                    const int ns = (int)slots.size(), nlm = lb - la;
                    const int pair0 = (int)O.g_pairinfo.size(), slot0 = (int)O.g_slot_pose.size();
                    if (O.ell + (size_t)(ns - 1) * nlm > 0x7fffffffULL) { O.ok = false; break; }
                    const int ell0 = (int)O.ell;
                    O.ell += (size_t)(ns - 1) * nlm;
                    int full = 1;
                    for (int l = la; l < lb; ++l) if (lm_eptr[l + 1] - lm_eptr[l] != ns - 1) full = 0;
                    const int hdr[8] = {host, ns, la, nlm, ell0, pair0, slot0, full};  // [7]: bit 0 full, bit 1 edges in slot order (pass 2)
                    O.g_hdr.insert(O.g_hdr.end(), hdr, hdr + 8);
                    O.g_slot_pose.insert(O.g_slot_pose.end(), slots.begin(), slots.end());
                    for (int a = 0; a < ns && O.ok; ++a)
                        for (int b = a; b < ns; ++b) {
                            const int pa = slots[a], pb = slots[b];
                            long long off = 0, info;
                            const bool fa = g->pose_fixed && g->pose_fixed[pa], fb = g->pose_fixed && g->pose_fixed[pb];
                            if (fa || fb) info = 3;
                            else if (a == b) { if (!block_off(pa, pa, off)) { O.ok = false; break; } info = (off << 2) | 2; }
                            else if (pose_off[pa] < pose_off[pb]) { if (!block_off(pa, pb, off)) { O.ok = false; break; } info = (off << 2) | 0; }
                            else { if (!block_off(pb, pa, off)) { O.ok = false; break; } info = (off << 2) | 1; }
                            O.g_pairinfo.push_back(info);
                        }
                    O.smem_max = std::max(O.smem_max, smem_bytes(ns, nlm));
                    O.max_obs_slots = std::max(O.max_obs_slots, ns - 1);
                    for (int p2 : slots) slot_of_pose[p2] = -1;
                    O.n_groups++;
                }
                l0 = l1;
            }
        };
        const int nth = pack_threads(E);
        std::vector<int> cut(nth + 1, L);
        cut[0] = 0;
        for (int t = 1; t < nth; ++t) {
            int c = std::max(cut[t - 1], (int)((long long)L * t / nth));
            while (c < L && c > 0 && lm_host[c] == lm_host[c - 1] && lm_eptr[c] != lm_eptr[c + 1]) ++c;  // next host boundary
            cut[t] = c;
        }
        std::vector<GroupPart> parts(nth);
        pack_parallel(nth, [&](int t) { if (cut[t] < cut[t + 1]) plan(cut[t], cut[t + 1], parts[t]); });
        K.grouped_ok = true;
        K.n_groups = 0; K.group_smem_max = 0;
        int max_obs_slots = 0;
        std::vector<size_t> oh(nth + 1, 0), os(nth + 1, 0), op(nth + 1, 0), ox(nth + 1, 0);
        for (int t = 0; t < nth; ++t) {
            const GroupPart &O = parts[t];
            K.grouped_ok = K.grouped_ok && O.ok;
            K.n_groups += O.n_groups; K.group_smem_max = std::max(K.group_smem_max, O.smem_max);
            max_obs_slots = std::max(max_obs_slots, O.max_obs_slots);
            oh[t + 1] = oh[t] + O.g_hdr.size(); os[t + 1] = os[t] + O.g_slot_pose.size();
            op[t + 1] = op[t] + O.g_pairinfo.size(); ox[t + 1] = ox[t] + O.ell;
        }
        if (ox[nth] > 0x7fffffffULL || op[nth] > 0x7fffffffULL) K.grouped_ok = false;
        K.g_hdr.clear(); K.g_slot_pose.clear(); K.g_pairinfo.clear(); K.ell_pjx.clear(); K.ell_pjy.clear(); K.ell_edge.clear();
        if (K.grouped_ok) {
            K.g_hdr.resize(oh[nth]); K.g_slot_pose.resize(os[nth]); K.g_pairinfo.resize(op[nth]);
            K.ell_pjx.resize(ox[nth]); K.ell_pjy.resize(ox[nth]); K.ell_edge.resize(ox[nth]);  // uninitialised: pass 2 writes every entry
            pack_parallel(nth, [&](int t) {
                const GroupPart &O = parts[t];
                std::copy(O.g_slot_pose.begin(), O.g_slot_pose.end(), K.g_slot_pose.begin() + os[t]);
                std::copy(O.g_pairinfo.begin(), O.g_pairinfo.end(), K.g_pairinfo.begin() + op[t]);
                std::vector<int> slot_of_pose(C, -1);
                const double qnan = std::numeric_limits<double>::quiet_NaN();
                for (size_t i = 0; i < O.g_hdr.size(); i += 8) {
                    int *h = &K.g_hdr[oh[t] + i];
                    for (int k = 0; k < 8; ++k) h[k] = O.g_hdr[i + k];
                    h[4] += (int)ox[t]; h[5] += (int)op[t]; h[6] += (int)os[t];
                    const int ns = h[1], la = h[2], nlm = h[3], lb = la + nlm;
                    const size_t ell0 = (size_t)h[4];
                    const int *slots = &K.g_slot_pose[h[6]];
                    for (int sl = 0; sl < ns; ++sl) slot_of_pose[slots[sl]] = sl;
                    double *ex = &K.ell_pjx[ell0], *ey = &K.ell_pjy[ell0];
                    int *ee = &K.ell_edge[ell0];
                    const size_t cnt_e = (size_t)(ns - 1) * nlm;
                    if (!(h[7] & 1)) {  // ragged group: missing observations are NaN / -1
                        for (size_t k = 0; k < cnt_e; ++k) { ex[k] = qnan; ey[k] = 0.0; ee[k] = -1; }
                    }
                    bool direct = (h[7] & 1) != 0;
                    const int e0 = lm_eptr[la];
                    for (int l = la; l < lb; ++l)
                        for (int e = lm_eptr[l]; e < lm_eptr[l + 1]; ++e) {
                            const int sl = slot_of_pose[e_pose_j[e]];
                            const size_t idx = (size_t)(sl - 1) * nlm + (l - la);
                            ex[idx] = pjx[e]; ey[idx] = pjy[e]; ee[idx] = e;
                            // every landmark's edges stored in slot order, landmarks back to back: edge(l, s) = e0 + l (ns - 1) + (s - 1),
                            // so the Schur kernel reads the group's H_lp rows as one slab (TMA bulk copy) without index loads
                            if (e != e0 + (l - la) * (ns - 1) + (sl - 1)) direct = false;
                        }
                    if (direct) h[7] |= 2;
                    for (int sl = 0; sl < ns; ++sl) slot_of_pose[slots[sl]] = -1;
                }
            });
        }
        if (K.n_groups == 0 || K.ext_free) K.grouped_ok = false;
        // one observer slot per warp, 4..10 warps (VIO_B200_GROUP_WARPS overrides; tuning knob)
        const int rounds = (max_obs_slots + 9) / 10;
        int nwarp = rounds > 0 ? (max_obs_slots + rounds - 1) / rounds : 4;
        if (nwarp < 4) nwarp = 4;
        if (nwarp > 10) nwarp = 10;
        if (const char *ev = getenv("VIO_B200_GROUP_WARPS")) nwarp = std::min(10, std::max(1, atoi(ev)));
        K.group_threads = 32 * nwarp;
    }

    ptm.mark("groups + ELL");
    K.C = C; K.NSB = NSB; K.NB = NB; K.P = P; K.L = L; K.Lglobal = Lg; K.E = E; K.storage = storage; K.nnzb = nnzb;
    K.s_count = s_count; K.batch = batch; K.Pper = Pper;
    K.pose_fixed.assign(C, 0); K.sb_fixed.assign(NSB, 0);
    if (g->pose_fixed) K.pose_fixed.assign(g->pose_fixed, g->pose_fixed + C);
    if (g->speedbias_fixed) K.sb_fixed.assign(g->speedbias_fixed, g->speedbias_fixed + NSB);
    K.row_fixed.assign(P, 0);
    for (int k = 0; k < NB; ++k)
        for (int d = 0; d < blk_dim[k]; ++d) K.row_fixed[blk_off[k] + d] = blk_fixed[k];
    for (int i = 0; i < g->n_se3prior; ++i)
        if (g->sp_pose[i] < 0 || g->sp_pose[i] >= C) return pack_fail(err, VIO_ERR_INVALID, "se3 prior %d: pose out of range", i);
    return VIO_OK;
}

// ---- lock-step batches -------------------------------------------------------------------------------------------
// Merge per-item packs (same C / NSB / pose-class order, dense storage, unsharded) into ONE packed graph whose reduced
// system is `B` stacked Pper x Pper blocks: pose-class offsets of item k are shifted by k*Pper, pose / speed-bias /
// landmark / edge indices by the item prefix sums.  `fill(k0, k1)` may be called from several threads on disjoint item
// ranges after `prepare`.
struct PackedMerge {
    const std::vector<PackedGraph> *Ks = nullptr;
    PackedGraph *M = nullptr;
    std::vector<long long> Lb, Eb, Gb, Sb, Pb, Xb;  // prefix sums: landmarks, edges, groups, slots, pairs, ELL entries
    int C = 0, NSB = 0, NBper = 0, Pper = 0;

    int prepare(const std::vector<PackedGraph> &ks, PackedGraph &m, std::string &err) {
        Ks = &ks; M = &m;
        const int B = (int)ks.size();
        C = ks[0].C; NSB = ks[0].NSB; NBper = ks[0].NB; Pper = ks[0].P;
        Lb.assign(B + 1, 0); Eb.assign(B + 1, 0); Gb.assign(B + 1, 0); Sb.assign(B + 1, 0); Pb.assign(B + 1, 0); Xb.assign(B + 1, 0);
        bool grouped = true;
        int threads = 0;
        size_t smem = 0;
        for (int k = 0; k < B; ++k) {
            const PackedGraph &K = ks[k];
            if (K.C != C || K.NSB != NSB || K.P != Pper || K.storage != VIO_STORAGE_DENSE || K.L != K.Lglobal)
                return pack_fail(err, VIO_ERR_UNSUPPORTED, "item %d: pose-class structure differs inside a lock-step batch", k);
            Lb[k + 1] = Lb[k] + K.L; Eb[k + 1] = Eb[k] + K.E;
            const bool gk = K.grouped_ok || K.E == 0;
            grouped = grouped && gk;
            const long long ng = K.grouped_ok ? K.n_groups : 0;
            Gb[k + 1] = Gb[k] + ng; Sb[k + 1] = Sb[k] + (K.grouped_ok ? (long long)K.g_slot_pose.size() : 0);
            Pb[k + 1] = Pb[k] + (K.grouped_ok ? (long long)K.g_pairinfo.size() : 0);
            Xb[k + 1] = Xb[k] + (K.grouped_ok ? (long long)K.ell_pjx.size() : 0);
            if (K.grouped_ok) { threads = std::max(threads, K.group_threads); smem = std::max(smem, K.group_smem_max); }
        }
        if (Lb[B] > 0x7fffffffLL || Eb[B] > 0x7fffffffLL || Xb[B] > 0x7fffffffLL || (long long)B * Pper > 0x7fffffffLL ||
            (long long)B * Pper * Pper > (1LL << 40))
            return pack_fail(err, VIO_ERR_UNSUPPORTED, "lock-step batch too large");
        m.C = B * C; m.NSB = B * NSB; m.NB = B * NBper; m.P = B * Pper; m.L = (int)Lb[B]; m.Lglobal = m.L; m.E = Eb[B];
        m.storage = VIO_STORAGE_DENSE; m.nnzb = 0; m.batch = B; m.Pper = Pper; m.s_count = (size_t)m.P * Pper;
        m.Lx = 0; m.Ex = 0; m.px_eptr.assign(1, 0); m.ex_pose.clear(); m.ex_ox.clear(); m.ex_oy.clear();
        for (int q = 0; q < 4; ++q) m.qic[q] = ks[0].qic[q];
        for (int q = 0; q < 3; ++q) m.tic[q] = ks[0].tic[q];
        m.pose_off.resize(m.C); m.sb_off.resize(m.NSB); m.pose_blk.resize(m.C); m.blk_off.resize(m.NB); m.blk_dim.resize(m.NB);
        m.blk_fixed.resize(m.NB); m.pose_fixed.resize(m.C); m.sb_fixed.resize(m.NSB); m.row_fixed.resize(m.P);
        {
            bool any = false;
            for (int k = 0; k < B; ++k) any = any || !ks[k].lm_fixed.empty();
            m.lm_fixed.assign(any ? (size_t)m.L : 0, 0);
            m.pt_fixed.clear();
        }
        m.lm_global.resize(m.L); m.lm_host.resize(m.L); m.lm_eptr.resize((size_t)m.L + 1); m.e_pose_j.resize(m.E);
        m.pix.resize(m.L); m.piy.resize(m.L); m.piz.resize(m.L); m.invd.resize(m.L); m.pjx.resize(m.E); m.pjy.resize(m.E);
        m.lm_eptr[m.L] = (int)m.E;
        m.rowptr.clear(); m.col.clear(); m.tr.clear(); m.diag.clear();
        m.grouped_ok = grouped && Gb[B] > 0;
        m.n_groups = m.grouped_ok ? (int)Gb[B] : 0; m.group_threads = threads; m.group_smem_max = smem;
        if (m.grouped_ok) {
            m.g_hdr.resize(8 * (size_t)Gb[B]); m.g_slot_pose.resize(Sb[B]); m.g_pairinfo.resize(Pb[B]);
            m.ell_pjx.resize(Xb[B]); m.ell_pjy.resize(Xb[B]); m.ell_edge.resize(Xb[B]);  // every entry is written by fill()
        } else {
            m.g_hdr.clear(); m.g_slot_pose.clear(); m.g_pairinfo.clear(); m.ell_pjx.clear(); m.ell_pjy.clear(); m.ell_edge.clear();
        }
        return VIO_OK;
    }

    void fill(int k0, int k1) const {
        PackedGraph &m = *M;
        for (int k = k0; k < k1; ++k) {
            const PackedGraph &K = (*Ks)[k];
            const int c0 = k * C, s0 = k * NSB, b0 = k * NBper, r0 = k * Pper;
            const int l0 = (int)Lb[k], e0 = (int)Eb[k];
            for (int i = 0; i < C; ++i) {
                m.pose_off[c0 + i] = K.pose_off[i] + r0; m.pose_blk[c0 + i] = K.pose_blk[i] + b0; m.pose_fixed[c0 + i] = K.pose_fixed[i];
            }
            for (int i = 0; i < NSB; ++i) { m.sb_off[s0 + i] = K.sb_off[i] + r0; m.sb_fixed[s0 + i] = K.sb_fixed[i]; }
            for (int i = 0; i < NBper; ++i) {
                m.blk_off[b0 + i] = K.blk_off[i] + r0; m.blk_dim[b0 + i] = K.blk_dim[i]; m.blk_fixed[b0 + i] = K.blk_fixed[i];
            }
            for (int i = 0; i < Pper; ++i) m.row_fixed[r0 + i] = K.row_fixed[i];
            for (int l = 0; l < K.L; ++l) {
                m.lm_global[l0 + l] = K.lm_global[l] + l0; m.lm_host[l0 + l] = K.lm_host[l] + c0; m.lm_eptr[l0 + l] = K.lm_eptr[l] + e0;
            }
            if (K.L && !m.lm_fixed.empty() && !K.lm_fixed.empty()) std::copy(K.lm_fixed.begin(), K.lm_fixed.end(), m.lm_fixed.begin() + l0);
            if (K.L) {
                std::copy(K.pix.begin(), K.pix.end(), m.pix.begin() + l0); std::copy(K.piy.begin(), K.piy.end(), m.piy.begin() + l0);
                std::copy(K.piz.begin(), K.piz.end(), m.piz.begin() + l0); std::copy(K.invd.begin(), K.invd.end(), m.invd.begin() + l0);
            }
            for (long long e = 0; e < K.E; ++e) m.e_pose_j[e0 + e] = K.e_pose_j[e] + c0;
            if (K.E) { std::copy(K.pjx.begin(), K.pjx.end(), m.pjx.begin() + e0); std::copy(K.pjy.begin(), K.pjy.end(), m.pjy.begin() + e0); }
            if (!m.grouped_ok || !K.grouped_ok) continue;
            const int gb = (int)Gb[k], sb = (int)Sb[k], pb = (int)Pb[k], xb = (int)Xb[k];
            for (int gi = 0; gi < K.n_groups; ++gi) {
                const int *h = &K.g_hdr[8 * (size_t)gi];
                int *o = &m.g_hdr[8 * (size_t)(gb + gi)];
                o[0] = h[0] + c0; o[1] = h[1]; o[2] = h[2] + l0; o[3] = h[3]; o[4] = h[4] + xb; o[5] = h[5] + pb; o[6] = h[6] + sb; o[7] = h[7];
            }
            for (size_t i = 0; i < K.g_slot_pose.size(); ++i) m.g_slot_pose[sb + i] = K.g_slot_pose[i] + c0;
            const long long blk_shift = (long long)r0 * Pper;  // (pose_off + k Pper) * Pper + col  =  off + k Pper^2
            for (size_t i = 0; i < K.g_pairinfo.size(); ++i) {
                const long long info = K.g_pairinfo[i];
                m.g_pairinfo[pb + i] = (info & 3) == 3 ? info : (((info >> 2) + blk_shift) << 2) | (info & 3);
            }
            std::copy(K.ell_pjx.begin(), K.ell_pjx.end(), m.ell_pjx.begin() + xb);
            std::copy(K.ell_pjy.begin(), K.ell_pjy.end(), m.ell_pjy.begin() + xb);
            for (size_t i = 0; i < K.ell_edge.size(); ++i) m.ell_edge[xb + i] = K.ell_edge[i] < 0 ? -1 : K.ell_edge[i] + e0;
        }
    }
};
