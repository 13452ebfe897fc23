// bvh_build.cu — host-side builder of the 4-wide BVH (EXTENSION; see flux_bvh.cuh for the
// exactness argument).  Object-median splits on the longest axis of the centroid bounds: the tree is
// balanced, so its depth — and with it the traversal stack a thread needs — is known in advance
// (3 * levels + 1 <= BVH_STACK; the leaf size is raised until that holds).  Each 4-wide node is made by
// splitting a range in two and each half in two again.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <limits>
#include <string>
#include <system_error>
#include <thread>

#include "flux_bvh.cuh"
#include "host_slices.h"

namespace {

struct Box {
    double lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::numeric_limits<double>::infinity();
            hi[k] = -std::numeric_limits<double>::infinity();
        }
    }
    void grow(const Box &b) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
};

// What the median splits move about is 32 bytes per primitive — the split key and where the box is — not the 80 of key
// and box together: the selection passes are bound by memory traffic, and a box is looked at once, when its leaf is made.
struct Item {
    double c[3];      // centroid (split key only)
    uint32_t ref;     // (kind << 30) | index
    uint32_t slot;    // its box: Builder::boxes[slot]
};
static_assert(sizeof(Item) == 32, "Item layout");

struct Builder {
    Item *items = nullptr;        // the primitives in the tree, a view into the caller's storage
    const Box *boxes = nullptr;   // by Item::slot
    BvhBuild *out;
    uint32_t leaf_size;
    double pad;

    Box bounds(uint32_t a, uint32_t b) const {
        Box x;
        x.reset();
        for (uint32_t i = a; i < b; i++) x.grow(boxes[items[i].slot]);
        return x;
    }
    bool presplit = false;   // partition() has already put every range in its final order: split() only names the middle

    // The whole sequence of nested median splits, ahead of the node emission and on several host threads.  Where
    // a range is cut depends on its length only (split() returns a + (b - a + 1) / 2) and a range's content is
    // fixed by the splits of its ancestors, so sibling ranges can be partitioned in any order, or at once: the
    // items end up exactly where the sequential build would leave them and the emitted tree is the same bit for
    // bit (flux_bvh_hash; tests/test_host_logic.py).  1 M triangles: 0.48 s -> 0.1 s on 8 cores.
    // (Measured and dropped, r2: the ranges of the first three levels selected on several threads each — two pivots from a
    // sorted sample, the range copied out in three runs, the selection proper in the middle run only: 16 instead of 22 ms
    // for those levels, 63 -> 57 ms for all twenty; the levels below are the bulk, and they already run 16 ranges at a time.)
    void partition(uint32_t a, uint32_t b, int par_levels) {
        uint32_t cut[5];
        cut[0] = a;
        cut[4] = b;
        cut[2] = split(a, b);
        auto left = [&] { cut[1] = (cut[2] - a > leaf_size) ? split(a, cut[2]) : a; };
        auto right = [&] { cut[3] = (b - cut[2] > leaf_size) ? split(cut[2], b) : cut[2]; };
        const bool par = par_levels > 0 && b - a >= (1u << 15);
        if (par) {
            std::vector<std::thread> one;
            flux_spawn_or_run(one, left);
            right();
            for (auto &t : one) t.join();
        } else {
            left();
            right();
        }
        std::vector<std::thread> th;
        for (int q = 0; q < 4; q++) {
            const uint32_t qa = cut[q], qb = cut[q + 1];
            if (qb - qa <= leaf_size) continue;
            if (par && q < 3) flux_spawn_or_run(th, [this, qa, qb, par_levels] { partition(qa, qb, par_levels - 1); });
            else partition(qa, qb, par ? par_levels - 1 : 0);
        }
        for (auto &t : th) t.join();
    }

    // split [a,b) at the object median of the longest centroid axis; returns the middle
    uint32_t split(uint32_t a, uint32_t b) {
        if (presplit) return a + (b - a + 1) / 2;
        double lo[3], hi[3];
        for (int k = 0; k < 3; k++) lo[k] = std::numeric_limits<double>::infinity(), hi[k] = -lo[k];
        for (uint32_t i = a; i < b; i++)
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], items[i].c[k]);
                hi[k] = std::max(hi[k], items[i].c[k]);
            }
        int ax = 0;
        for (int k = 1; k < 3; k++)
            if (hi[k] - lo[k] > hi[ax] - lo[ax]) ax = k;
        const uint32_t mid = a + (b - a + 1) / 2;
        std::nth_element(items + a, items + mid, items + b, [ax](const Item &p, const Item &q) {
            return p.c[ax] < q.c[ax] || (p.c[ax] == q.c[ax] && p.ref < q.ref);
        });
        return mid;
    }
    // Leaves tile [0, n) in item order and the nodes are numbered in depth-first pre-order, so both arrays can be laid
    // out before anything is written: a leaf over [a, b) has its references at prims[a .. b), and a subtree over m items
    // has nodes_for(m) nodes — a function of m alone, like the cuts.  Subtrees are then emitted on several host threads,
    // each into its own slice (1 M triangles: 69 ms -> 15 ms on 8 cores); the arrays are the sequential build's bit for bit.
    uint32_t make_leaf(uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; i++) out->prims[i] = items[i].ref;   // kept for every leaf: the host-side checks walk it
        if (b - a == 1) {   // the primitive named in the reference itself (BVH_DIRECT): kind << 28 | index
            const uint32_t pr = items[a].ref;
            return BVH_LEAF | BVH_DIRECT | ((pr >> 30) << 28) | (pr & 0x0FFFFFFFu);
        }
        return BVH_LEAF | (a << 3) | (b - a - 1);
    }
    void set_child(uint32_t node, int slot, uint32_t ref, const Box &bx) {
        BvhNode4 &n = out->nodes[node];
        n.child[slot] = ref;
        for (int k = 0; k < 3; k++) {   // padded in f64, then rounded OUTWARD to f32
            const double lo = bx.lo[k] - pad, hi = bx.hi[k] + pad;
            float flo = (float)lo, fhi = (float)hi;
            if ((double)flo > lo) flo = std::nextafter(flo, -std::numeric_limits<float>::infinity());
            if ((double)fhi < hi) fhi = std::nextafter(fhi, std::numeric_limits<float>::infinity());
            n.lo[k][slot] = flo;
            n.hi[k][slot] = fhi;
        }
    }
    // the four quarter lengths split() and build_node() cut a range of m items into (0 = "no split": that slot stays empty)
    void quarters(uint64_t m, uint64_t q[4]) const {
        const uint64_t l = (m + 1) / 2, r = m - l;   // split(): mid = a + (b - a + 1) / 2
        q[0] = l > leaf_size ? (l + 1) / 2 : 0;
        q[1] = l > leaf_size ? l - (l + 1) / 2 : l;
        q[2] = r > leaf_size ? (r + 1) / 2 : 0;
        q[3] = r > leaf_size ? r - (r + 1) / 2 : r;
    }
    std::vector<std::pair<uint64_t, uint64_t>> count_memo;   // the recursion only ever sees a handful of distinct lengths per level
    // nodes of the subtree over m > leaf_size items (filled before the threads start; read-only afterwards)
    uint64_t nodes_for(uint64_t m) {
        for (auto &kv : count_memo)
            if (kv.first == m) return kv.second;
        uint64_t q[4], total = 1;
        quarters(m, q);
        for (uint64_t x : q)
            if (x > leaf_size) total += nodes_for(x);
        count_memo.push_back({m, total});
        return total;
    }
    uint64_t nodes_for_ro(uint64_t m) const {
        for (auto &kv : count_memo)
            if (kv.first == m) return kv.second;
        return 0;   // unreachable: nodes_for(n) has visited every length the build can meet
    }
    std::atomic<uint32_t> deepest{0};
    // builds the subtree over [a,b) (b - a > leaf_size) at node index `me` (and the indices after it); level = depth of this node
    void build_node(uint32_t a, uint32_t b, uint32_t level, uint32_t me, Box *whole, int par_levels) {
        {
            uint32_t seen = deepest.load(std::memory_order_relaxed);
            while (seen < level + 1 && !deepest.compare_exchange_weak(seen, level + 1, std::memory_order_relaxed)) {}
        }
        {
            BvhNode4 &n = out->nodes[me];
            for (int s = 0; s < 4; s++) {
                n.child[s] = BVH_EMPTY;
                for (int k = 0; k < 3; k++) {
                    n.lo[k][s] = std::numeric_limits<float>::infinity();
                    n.hi[k][s] = -std::numeric_limits<float>::infinity();
                }
            }
            for (int s = 0; s < 4; s++) n.pad[s] = 0;
        }
        uint32_t cut[5];
        cut[0] = a;
        cut[4] = b;
        cut[2] = split(a, b);
        cut[1] = (cut[2] - a > leaf_size) ? split(a, cut[2]) : a;          // a == "no split": slot 0 empty
        cut[3] = (b - cut[2] > leaf_size) ? split(cut[2], b) : cut[2];
        // the box of a subtree is the union of its children's boxes (min / max are exact, so this is the very
        // box a scan over its items gives): one pass over the items in all, not one per level
        Box bx[4];
        uint32_t ref[4];
        uint32_t next = me + 1;
        const bool par = par_levels > 0 && b - a >= (1u << 15);
        std::vector<std::thread> th;
        for (int q = 0; q < 4; q++) {
            const uint32_t qa = cut[q], qb = cut[q + 1];
            if (qa == qb) continue;
            if (qb - qa <= leaf_size) {
                bx[q] = bounds(qa, qb);
                ref[q] = make_leaf(qa, qb);
            } else {
                const uint32_t child = next;
                next += (uint32_t)nodes_for_ro(qb - qa);
                ref[q] = child;
                Box *dst = &bx[q];
                if (par && q < 3) flux_spawn_or_run(th, [this, qa, qb, level, child, dst, par_levels] { build_node(qa, qb, level + 1, child, dst, par_levels - 1); });
                else build_node(qa, qb, level + 1, child, dst, par ? par_levels - 1 : 0);
            }
        }
        for (auto &t : th) t.join();
        int slot = 0;
        Box mine;
        mine.reset();
        for (int q = 0; q < 4; q++) {
            if (cut[q] == cut[q + 1]) continue;
            mine.grow(bx[q]);
            set_child(me, slot++, ref[q], bx[q]);
        }
        if (whole) *whole = mine;
    }
};

}  // namespace

static bool build_bvh4_impl(const double *sph, const uint32_t *sph_meta, uint32_t ns, const double *tri, const uint32_t *tri_meta,
                            const double *tri_v1, const double *tri_v2, uint32_t nt, BvhBuild &out, std::string &err);

// No exception leaves the builder: its callers are extern "C" (flux_set_scene, flux_bvh_describe, flux_bvh_hash).
bool build_bvh4(const double *sph, const uint32_t *sph_meta, uint32_t ns, const double *tri, const uint32_t *tri_meta,
                const double *tri_v1, const double *tri_v2, uint32_t nt, BvhBuild &out, std::string &err) {
    try {
        return build_bvh4_impl(sph, sph_meta, ns, tri, tri_meta, tri_v1, tri_v2, nt, out, err);
    } catch (const std::exception &e) {
        err = std::string("bvh: ") + e.what();
    } catch (...) {
        err = "bvh: unknown failure";
    }
    out = BvhBuild{};
    return false;
}

static bool build_bvh4_impl(const double *sph, const uint32_t *sph_meta, uint32_t ns, const double *tri, const uint32_t *tri_meta,
                            const double *tri_v1, const double *tri_v2, uint32_t nt, BvhBuild &out, std::string &err) {
    out = BvhBuild{};
    const bool timing = std::getenv("FLUX_BVH_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto t = now();
        std::fprintf(stderr, "[bvh] %-10s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    if ((uint64_t)ns >= (1u << 30) || (uint64_t)nt >= (1u << 30) || (uint64_t)ns + nt >= (1u << 27)) {   // leaf references: 28-bit direct primitive index, 27-bit offset << 3
        err = "bvh: too many primitives";
        return false;
    }
    // ---- leaf records ----
    out.sph.resize(ns);
    for (uint32_t i = 0; i < ns; i++) {
        SphRec &s = out.sph[i];
        s.c0x = sph[(size_t)SPH_C0X * ns + i]; s.c1x = sph[(size_t)SPH_C1X * ns + i];
        s.c0y = sph[(size_t)SPH_C0Y * ns + i]; s.c1y = sph[(size_t)SPH_C1Y * ns + i];
        s.c0z = sph[(size_t)SPH_C0Z * ns + i]; s.c1z = sph[(size_t)SPH_C1Z * ns + i];
        s.cx = sph[(size_t)SPH_CX * ns + i]; s.cy = sph[(size_t)SPH_CY * ns + i]; s.cz = sph[(size_t)SPH_CZ * ns + i];
        s.rr = sph[(size_t)SPH_RR * ns + i]; s.r = sph[(size_t)SPH_R * ns + i]; s.inv = sph[(size_t)SPH_INV * ns + i];
        s.shape_id = sph_meta[i]; s.material = sph_meta[ns + i]; s.index = i; s.pad = 0;
    }
    out.tri.resize(nt);
    // independent per triangle: in slices on several host threads when there are many (host_slices.h)
    const auto in_slices = flux_in_slices;
    in_slices(nt, [&](uint32_t lo_i, uint32_t hi_i) {
    for (uint32_t i = lo_i; i < hi_i; i++) {
        TriRec &q = out.tri[i];
        q.v0x = tri[(size_t)TRI_V0X * nt + i]; q.v0y = tri[(size_t)TRI_V0Y * nt + i]; q.v0z = tri[(size_t)TRI_V0Z * nt + i];
        q.e1x = tri[(size_t)TRI_E1X * nt + i]; q.e1y = tri[(size_t)TRI_E1Y * nt + i]; q.e1z = tri[(size_t)TRI_E1Z * nt + i];
        q.e2x = tri[(size_t)TRI_E2X * nt + i]; q.e2y = tri[(size_t)TRI_E2Y * nt + i]; q.e2z = tri[(size_t)TRI_E2Z * nt + i];
        q.shape_id = tri_meta[i]; q.material = tri_meta[nt + i]; q.index = i;
        q.pad[0] = q.pad[1] = q.pad[2] = 0;
    }
    });
    lap("records");
    // ---- primitive boxes ----
    Builder B;
    B.out = &out;
    flux_raw_vector<Item> items((size_t)ns + nt);   // spheres first (those with a finite box), then the triangles
    flux_raw_vector<Box> boxes((size_t)ns + nt);    // boxes[k] belongs to the item made k-th, wherever the splits move it
    auto finite_box = [](const Box &b) {
        for (int k = 0; k < 3; k++)
            if (!std::isfinite(b.lo[k]) || !std::isfinite(b.hi[k])) return false;
        return true;
    };
    // The bounds the later steps need — of the centroids (what "large" is measured against) and of the boxes (the padding
    // scale) — are gathered in the passes that make the items, per slice; min / max are exact, so the order of the unions
    // does not matter.
    Box sph_cb, tri_cb, tri_all;
    sph_cb.reset(); tri_cb.reset(); tri_all.reset();
    size_t tri_base = 0;
    for (uint32_t i = 0; i < ns; i++) {
        Item it;
        Box bx;
        const SphRec &s = out.sph[i];
        bx.lo[0] = std::min(s.c0x, s.c1x); bx.hi[0] = std::max(s.c0x, s.c1x);   // negative radii swap corners
        bx.lo[1] = std::min(s.c0y, s.c1y); bx.hi[1] = std::max(s.c0y, s.c1y);
        bx.lo[2] = std::min(s.c0z, s.c1z); bx.hi[2] = std::max(s.c0z, s.c1z);
        it.ref = ((uint32_t)KIND_SPHERE << 30) | i;
        if (!finite_box(bx)) {  // NaN / inf geometry: keep the linear scan's verdict by testing it linearly
            out.linear.push_back(i);
            continue;
        }
        for (int k = 0; k < 3; k++) {
            it.c[k] = 0.5 * bx.lo[k] + 0.5 * bx.hi[k];
            sph_cb.lo[k] = std::min(sph_cb.lo[k], it.c[k]);
            sph_cb.hi[k] = std::max(sph_cb.hi[k], it.c[k]);
        }
        it.slot = (uint32_t)tri_base;
        boxes[tri_base] = bx;
        items[tri_base++] = it;
    }
    const size_t n_items = tri_base + nt;
    std::vector<uint32_t> first_bad(1, 0xFFFFFFFFu);
    std::mutex merge_mu;   // what the slices hand back: the first bad triangle, their partial bounds
    in_slices(nt, [&](uint32_t lo_i, uint32_t hi_i) {
    Box part_c, part_b;
    part_c.reset(); part_b.reset();
    for (uint32_t i = lo_i; i < hi_i; i++) {
        Item it;
        Box bx;
        const TriRec &q = out.tri[i];
        const double v0[3] = {q.v0x, q.v0y, q.v0z};
        for (int k = 0; k < 3; k++) {
            const double a = v0[k], b = tri_v1[3 * (size_t)i + k], c = tri_v2[3 * (size_t)i + k];
            // the device works with e1 = v1 - v0, e2 = v2 - v0: cover both the given and the reconstructed vertices
            const double e1 = (k == 0 ? q.e1x : k == 1 ? q.e1y : q.e1z), e2 = (k == 0 ? q.e2x : k == 1 ? q.e2y : q.e2z);
            bx.lo[k] = std::min({a, b, c, a + e1, a + e2});
            bx.hi[k] = std::max({a, b, c, a + e1, a + e2});
            it.c[k] = 0.5 * bx.lo[k] + 0.5 * bx.hi[k];
        }
        it.ref = ((uint32_t)KIND_TRI << 30) | i;
        it.slot = (uint32_t)(tri_base + i);
        bool finite = finite_box(bx);
        for (int k = 0; k < 3; k++)   // min / max skip a NaN or not depending on where it stands: look at the vertices themselves
            finite = finite && std::isfinite(v0[k]) && std::isfinite(tri_v1[3 * (size_t)i + k]) && std::isfinite(tri_v2[3 * (size_t)i + k]);
        if (!finite) {
            std::lock_guard<std::mutex> g(merge_mu);
            first_bad[0] = std::min(first_bad[0], i);   // the lowest index, whichever thread saw it
        }
        part_b.grow(bx);
        for (int k = 0; k < 3; k++) {
            part_c.lo[k] = std::min(part_c.lo[k], it.c[k]);
            part_c.hi[k] = std::max(part_c.hi[k], it.c[k]);
        }
        boxes[tri_base + i] = bx;
        items[tri_base + i] = it;
    }
    std::lock_guard<std::mutex> g(merge_mu);
    tri_cb.grow(part_c);
    tri_all.grow(part_b);
    });
    if (first_bad[0] != 0xFFFFFFFFu) {
        err = "bvh: triangle " + std::to_string(first_bad[0]) + " has a non-finite vertex";
        return false;
    }
    lap("boxes");
    // ---- oversized spheres go to the linear list (at most 64, largest first) ----
    // Only spheres are ever moved, and they stand first: the ones that stay close up towards the triangles and the tree's
    // items begin that many places later — no pass over, and no copy of, the million triangles behind them.
    size_t first_item = 0;
    if (tri_base > 0) {
        auto diag_of = [](const Box &b) {
            const double x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
            return std::sqrt(x * x + y * y + z * z);
        };
        // centroid bounds of everything approximate the populated region
        Box cb = sph_cb;
        cb.grow(tri_cb);
        std::vector<std::pair<double, size_t>> big;
        for (size_t k = 0; k < tri_base; k++) big.push_back({diag_of(boxes[k]), k});
        std::sort(big.begin(), big.end(), [](auto &p, auto &q) { return p.first > q.first || (p.first == q.first && p.second < q.second); });
        if (big.size() > 64) big.resize(64);
        std::vector<size_t> drop;
        const double scene = diag_of(cb);
        std::vector<std::pair<double, size_t>> large;   // the candidates: large against the populated region ...
        for (auto &p : big)
            if (p.first > 0.25 * scene) large.push_back(p);
        if (n_items > 64 && !large.empty()) {
            // ... and against 16 times the median diagonal, so that one environment sphere does not define "large".  The
            // median itself is not needed: x -> 16.0 * x is monotone, so 16 * median < d exactly when more than half of the
            // diagonals (n / 2 + 1 of them: the median is the element of rank n / 2) have 16 * diagonal < d — one counting
            // pass in slices, no array of a million diagonals to select in.
            std::vector<uint64_t> below(large.size(), 0);
            std::mutex mu;
            in_slices((uint32_t)n_items, [&](uint32_t lo_i, uint32_t hi_i) {
                std::vector<uint64_t> mine(large.size(), 0);
                for (uint32_t i = lo_i; i < hi_i; i++) {
                    const double d16 = 16.0 * diag_of(boxes[i]);
                    for (size_t c = 0; c < large.size(); c++) mine[c] += d16 < large[c].first;
                }
                std::lock_guard<std::mutex> g(mu);
                for (size_t c = 0; c < large.size(); c++) below[c] += mine[c];
            });
            for (size_t c = 0; c < large.size(); c++)
                if (below[c] >= (uint64_t)n_items / 2 + 1) drop.push_back(large[c].second);
        }
        if (!drop.empty()) {   // at most 64, all among the leading spheres; the order of the others is kept
            std::sort(drop.begin(), drop.end());
            for (size_t k : drop) out.linear.push_back(items[k].ref & 0x3FFFFFFFu);
            size_t w = tri_base, d = drop.size();
            for (size_t k = tri_base; k-- > 0;) {
                if (d > 0 && drop[d - 1] == k) {
                    d--;
                    continue;
                }
                items[--w] = items[k];
            }
            first_item = w;   // == drop.size()
        }
    }
    std::sort(out.linear.begin(), out.linear.end());
    const uint64_t n = n_items - first_item;
    if (n == 0) return true;  // tree-less: linear list only
    B.items = items.data() + first_item;
    B.boxes = boxes.data();

    Box all = tri_all;   // ... and the spheres that stayed
    {
        std::mutex mu;
        in_slices((uint32_t)(tri_base - first_item), [&](uint32_t lo_i, uint32_t hi_i) {
            Box part;
            part.reset();
            for (uint32_t i = lo_i; i < hi_i; i++) part.grow(boxes[B.items[i].slot]);
            std::lock_guard<std::mutex> g(mu);
            all.grow(part);
        });
    }
    double ext = 0.0;
    for (int k = 0; k < 3; k++) ext = std::max({ext, std::fabs(all.lo[k]), std::fabs(all.hi[k])});
    out.extent = ext;
    B.pad = BVH_PAD_REL * std::max(ext, 1e-300);

    lap("oversized");
    // ---- leaf size: smallest in {4, 8} whose (balanced) tree fits the traversal stack ----
    // where a range is cut depends on its length only, so the depth of the tree for a given leaf size is known
    // before anything is moved: levels(m) = 1 + the deepest of the four quarters that are still larger than a leaf
    auto levels_for = [](uint64_t total, uint32_t leaf) {
        std::vector<std::pair<uint64_t, uint32_t>> memo;   // the recursion only ever sees a handful of distinct lengths per level
        struct Rec {
            std::vector<std::pair<uint64_t, uint32_t>> &memo;
            uint32_t leaf;
            uint32_t operator()(uint64_t m) {
                for (auto &kv : memo)
                    if (kv.first == m) return kv.second;
                const uint64_t l = (m + 1) / 2, r = m - l;                       // split(): mid = a + (b - a + 1) / 2
                const uint64_t q[4] = {l > leaf ? (l + 1) / 2 : 0, l > leaf ? l - (l + 1) / 2 : l, r > leaf ? (r + 1) / 2 : 0,
                                       r > leaf ? r - (r + 1) / 2 : r};
                uint32_t deepest = 0;
                for (uint64_t x : q)
                    if (x > leaf) deepest = std::max(deepest, (*this)(x));
                memo.push_back({m, deepest + 1});
                return deepest + 1;
            }
        } rec{memo, leaf};
        return total <= leaf ? 1u : rec(total);
    };
    uint32_t leaf = BVH_FIRST_LEAF;
    while (3 * levels_for(n, leaf) + 1 > BVH_STACK) {
        leaf *= 2;
        if (leaf > 8) {
            err = "bvh: tree depth " + std::to_string(levels_for(n, 8)) + " exceeds the traversal stack (scene too large)";
            return false;
        }
    }
    {
        B.leaf_size = leaf;
        out.leaf_size = leaf;
        out.nodes.clear();
        out.prims.clear();
        out.depth = 0;
        out.prims.resize(n);   // every entry is written by the leaf that covers it (make_leaf)
        if (n <= leaf) {  // a single leaf under a root node
            out.nodes.resize(1);
            BvhNode4 &r = out.nodes[0];
            for (int s = 0; s < 4; s++) {
                r.child[s] = BVH_EMPTY;
                for (int k = 0; k < 3; k++) r.lo[k][s] = std::numeric_limits<float>::infinity(), r.hi[k][s] = -r.lo[k][s];
            }
            for (int s = 0; s < 4; s++) r.pad[s] = 0;
            out.depth = 1;
            B.set_child(0, 0, B.make_leaf(0, (uint32_t)n), B.bounds(0, (uint32_t)n));
        } else {
            B.presplit = false;
            lap("plan");
            B.partition(0, (uint32_t)n, 2);   // up to 4^2 ranges in flight
            lap("partition");
            B.presplit = true;
            const uint64_t n_nodes = B.nodes_for(n);
            if (n_nodes >= (1ull << 30)) {
                err = "bvh: too many nodes";
                return false;
            }
            out.nodes.resize((size_t)n_nodes);
            B.build_node(0, (uint32_t)n, 0, 0, nullptr, 2);   // up to 4^2 subtrees in flight
            out.depth = B.deepest.load();
            lap("emit");
        }
        if (out.depth != levels_for(n, leaf)) {   // levels_for() is the depth build_node() reaches, or the leaf size was chosen wrongly
            err = "bvh: internal error: planned depth " + std::to_string(levels_for(n, leaf)) + ", built depth " + std::to_string(out.depth);
            return false;
        }
    }
    return true;
}
