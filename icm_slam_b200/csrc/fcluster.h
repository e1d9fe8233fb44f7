// fcluster.h -- host restatement of the ONE scipy call Mapa.actualizar makes at t = 0 of pass 0
// (ICM_SLAM.py:161):   c = fcluster(linkage(pdist(obs)), dist_thr) - 1
// i.e. single linkage + flat clusters by the INCONSISTENCY criterion with depth 2 (scipy's defaults; not
// a distance cut).  scipy is a third-party dependency of the reference (pinned 1.5.4,
// requisitos.txt:21; 1.18.1 in this image) whose compiled source is not under /root/reference, so this
// follows its published algorithm (scipy/cluster/_hierarchy.pyx: mst_single_linkage, label,
// inconsistent, get_max_Rfield_for_each_cluster, cluster_monocrit) and is pinned by golden vectors
// minted from scipy itself (oracle/make_golden_fcluster.py -> tests/golden/fcluster.npz).
// n <= a few hundred points (the kept beams of one scan); everything is O(n^2) on the host.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <algorithm>
#include <vector>

namespace icm_fcluster {

struct Link { int a, b; double d; int size; };

// linkage(pdist(P), 'single'): Prim's MST from point 0, stable sort by distance, union-find relabelling
inline std::vector<Link> single_linkage(const double* px, const double* py, int n)
{
    std::vector<Link> Z;
    if (n < 2) return Z;
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<char> merged(n, 0);
    std::vector<double> D(n, inf);
    std::vector<Link> raw(n - 1);
    int x = 0;
    for (int k = 0; k < n - 1; ++k) {
        double current_min = inf;
        int y = -1;
        merged[x] = 1;
        for (int i = 0; i < n; ++i) {
            if (merged[i]) continue;
            const double dx = px[x] - px[i], dy = py[x] - py[i];
            const double dist = std::sqrt(dx * dx + dy * dy);      // pdist 'euclidean'
            if (D[i] > dist) D[i] = dist;
            if (D[i] < current_min) { y = i; current_min = D[i]; }
        }
        raw[k] = Link{x, y, current_min, 0};
        x = y;
    }
    std::vector<int> order(n - 1);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int i, int j) { return raw[i].d < raw[j].d; });   // argsort(kind='mergesort')
    // label(): cluster ids n, n+1, ... in merge order; smaller root first
    std::vector<int> parent(2 * n - 1), size(2 * n - 1, 1);
    for (int i = 0; i < 2 * n - 1; ++i) parent[i] = i;
    int next_label = n;
    auto find = [&](int v) {
        int r = v;
        while (parent[r] != r) r = parent[r];
        while (parent[v] != r) { int nx = parent[v]; parent[v] = r; v = nx; }
        return r;
    };
    Z.resize(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        const Link& l = raw[order[i]];
        int xr = find(l.a), yr = find(l.b);
        if (xr > yr) std::swap(xr, yr);
        parent[xr] = next_label; parent[yr] = next_label;
        size[next_label] = size[xr] + size[yr];
        Z[i] = Link{xr, yr, l.d, size[next_label]};
        ++next_label;
    }
    return Z;
}

// inconsistent(Z, d)[:, 3]: (height - mean) / std over the links of the subtree down to depth d
inline std::vector<double> inconsistency(const std::vector<Link>& Z, int n, int depth)
{
    const int m = n - 1;
    std::vector<double> R(m, 0.0);
    std::vector<int> curr(std::max(depth, 1) + 1);
    std::vector<char> visited(2 * n, 0);
    for (int i = 0; i < m; ++i) {
        std::fill(visited.begin(), visited.end(), 0);
        int k = 0, level_count = 0;
        double level_sum = 0.0, level_std_sum = 0.0;
        curr[0] = i;
        while (k >= 0) {
            const int root = curr[k];
            if (k < depth - 1) {
                const int lc = Z[root].a;
                if (lc >= n && !visited[lc]) { visited[lc] = 1; ++k; curr[k] = lc - n; continue; }
                const int rc = Z[root].b;
                if (rc >= n && !visited[rc]) { visited[rc] = 1; ++k; curr[k] = rc - n; continue; }
            }
            const double dist = Z[root].d;
            ++level_count;
            level_sum += dist;
            level_std_sum += dist * dist;
            --k;
        }
        const double mean = level_sum / level_count;
        double var;
        if (level_count < 2) var = (level_std_sum - (level_sum * level_sum)) / level_count;
        else var = (level_std_sum - ((level_sum * level_sum) / level_count)) / (level_count - 1);
        if (var > 0.0) {
            const double sd = std::sqrt(var);
            R[i] = (Z[i].d - mean) / sd;
        }
    }
    return R;
}

// fcluster(Z, t, criterion='inconsistent', depth=2) - 1 : labels 0 .. k-1 in scipy's numbering
inline int fcluster_inconsistent(const double* px, const double* py, int n, double t, int* labels, int depth = 2)
{
    if (n <= 0) return 0;
    if (n == 1) { labels[0] = 0; return 1; }
    const std::vector<Link> Z = single_linkage(px, py, n);
    const std::vector<double> R = inconsistency(Z, n, depth);
    const int m = n - 1;
    // get_max_Rfield_for_each_cluster: links are in merge order, children always precede their parent
    std::vector<double> MI(m);
    for (int i = 0; i < m; ++i) {
        double v = R[i];
        if (Z[i].a >= n) v = std::max(v, MI[Z[i].a - n]);
        if (Z[i].b >= n) v = std::max(v, MI[Z[i].b - n]);
        MI[i] = v;
    }
    // cluster_monocrit: depth-first, left child first; a node whose max coefficient <= t leads a cluster
    std::vector<int> curr(n), T(n, 0);
    std::vector<char> visited(2 * n, 0);
    int n_cluster = 0, leader = -1, k = 0;
    curr[0] = 2 * n - 2;
    while (k >= 0) {
        const int root = curr[k] - n;
        const int lc = Z[root].a, rc = Z[root].b;
        if (leader == -1 && MI[root] <= t) { leader = root; ++n_cluster; }
        if (lc >= n && !visited[lc]) { visited[lc] = 1; ++k; curr[k] = lc; continue; }
        if (rc >= n && !visited[rc]) { visited[rc] = 1; ++k; curr[k] = rc; continue; }
        if (lc < n) { if (leader == -1) ++n_cluster; T[lc] = n_cluster; }
        if (rc < n) { if (leader == -1) ++n_cluster; T[rc] = n_cluster; }
        if (leader == root) leader = -1;
        --k;
    }
    for (int i = 0; i < n; ++i) labels[i] = T[i] - 1;
    return n_cluster;
}

}   // namespace icm_fcluster
