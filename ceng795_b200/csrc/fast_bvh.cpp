// fast_bvh.cpp — binned-SAH BVH builder for the device traversal (see fast_bvh.h).
#include "fast_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <system_error>
#include <thread>

namespace b200rt_fast {
namespace {

constexpr int kBins = 16;
int kMaxLeaf = 4;
float kTraversalCost = 1.0f;  // relative to one triangle test (swept on cfg2: profiles/r01b_tuning.txt)

struct Prim {
  float mn[3], mx[3], c[3];
  int32_t id;
};
struct Bx {
  float mn[3], mx[3];
  Bx() {
    for (int i = 0; i < 3; i++) { mn[i] = std::numeric_limits<float>::infinity(); mx[i] = -mn[i]; }
  }
  void add(const float* a, const float* b) {
    for (int i = 0; i < 3; i++) { mn[i] = std::min(mn[i], a[i]); mx[i] = std::max(mx[i], b[i]); }
  }
  void add(const Bx& o) { add(o.mn, o.mx); }
  float area() const {
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (!(dx >= 0.f)) return 0.f;
    return 2.f * (dx * dy + dy * dz + dz * dx);
  }
};

// One builder context per thread: `prims` is shared (every subtree partitions
// its own slice in place), `nodes` / `order` are the thread's own output.
struct Ctx;
struct Ctx {
  Prim* prims;
  std::vector<b200rt_bvh_node>* nodes;
  std::vector<int32_t>* order;
  float grow;
  int par_levels;  // levels below this one that may still hand their left half to another thread
};
constexpr int kParMinPrims = 1 << 14;

void put(float* mn, float* mx, const Bx& b, float g) {
  for (int i = 0; i < 3; i++) { mn[i] = b.mn[i] - g; mx[i] = b.mx[i] + g; }
}

uint32_t make_leaf(Ctx& cx, int s, int e) {
  const int n = e - s;
  const uint32_t start = (uint32_t)cx.order->size();
  for (int i = s; i < e; i++) cx.order->push_back(cx.prims[i].id);
  return B200RT_REF(B200RT_KIND_LEAF1 + (uint32_t)(n - 1), start);
}

// A ref built against arrays counted from 0, moved to where those arrays land.
uint32_t shift_ref(uint32_t ref, uint32_t node_off, uint32_t order_off) {
  if (ref == B200RT_REF_NONE) return ref;
  const uint32_t kind = B200RT_REF_KIND(ref), idx = B200RT_REF_INDEX(ref);
  return B200RT_REF(kind, idx + (kind == B200RT_KIND_FNODE ? node_off : order_off));
}
void splice(struct Ctx& cx, const std::vector<b200rt_bvh_node>& nodes, const std::vector<int32_t>& order,
            uint32_t node_off, uint32_t order_off);

// Returns the child ref of the subtree over prims[s,e); *box = its tight bounds.
uint32_t build_rec(Ctx& cx, int s, int e, Bx* box) {
  Bx b, cb;
  for (int i = s; i < e; i++) {
    b.add(cx.prims[i].mn, cx.prims[i].mx);
    cb.add(cx.prims[i].c, cx.prims[i].c);
  }
  *box = b;
  const int n = e - s;
  if (n == 1) return make_leaf(cx, s, e);

  // best binned split over the three axes
  float best_cost = std::numeric_limits<float>::infinity();
  int best_axis = -1, best_bin = -1;
  for (int ax = 0; ax < 3; ax++) {
    const float ext = cb.mx[ax] - cb.mn[ax];
    if (!(ext > 0.f)) continue;
    Bx bb[kBins];
    int cnt[kBins] = {0};
    const float scale = (float)kBins / ext;
    for (int i = s; i < e; i++) {
      int k = (int)((cx.prims[i].c[ax] - cb.mn[ax]) * scale);
      k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
      bb[k].add(cx.prims[i].mn, cx.prims[i].mx);
      cnt[k]++;
    }
    float la[kBins - 1], ra[kBins - 1];
    int ln[kBins - 1], rn[kBins - 1];
    Bx acc;
    int c = 0;
    for (int k = 0; k < kBins - 1; k++) { acc.add(bb[k]); c += cnt[k]; la[k] = acc.area(); ln[k] = c; }
    acc = Bx(); c = 0;
    for (int k = kBins - 1; k > 0; k--) { acc.add(bb[k]); c += cnt[k]; ra[k - 1] = acc.area(); rn[k - 1] = c; }
    for (int k = 0; k < kBins - 1; k++) {
      if (ln[k] == 0 || rn[k] == 0) continue;
      const float cost = la[k] * ln[k] + ra[k] * rn[k];
      if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = k; }
    }
  }
  const float pa = b.area();
  if (n <= kMaxLeaf) {
    const float leaf_cost = (float)n * pa;
    const float split_cost = kTraversalCost * pa + best_cost;
    if (best_axis < 0 || leaf_cost <= split_cost) return make_leaf(cx, s, e);
  }
  int mid;
  if (best_axis < 0) {
    mid = s + n / 2;  // coincident centroids: split by position in the array
  } else {
    const float ext = cb.mx[best_axis] - cb.mn[best_axis];
    const float scale = (float)kBins / ext;
    const float cmin = cb.mn[best_axis];
    const int ax = best_axis, bin = best_bin;
    Prim* first = cx.prims + s;
    Prim* last = cx.prims + e;
    Prim* m = std::partition(first, last, [=](const Prim& p) {
      int k = (int)((p.c[ax] - cmin) * scale);
      k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
      return k <= bin;
    });
    mid = (int)(m - cx.prims);
    if (mid == s || mid == e) mid = s + n / 2;
  }
  const uint32_t me = (uint32_t)cx.nodes->size();
  cx.nodes->push_back(b200rt_bvh_node());
  Bx lb, rb;
  uint32_t l, r;
  if (cx.par_levels > 0 && n >= kParMinPrims) {
    // The two halves are independent (disjoint slices of prims).  The left one is
    // built by another thread into its own arrays, with indices counted from 0,
    // and spliced in where the sequential recursion would have put it: directly
    // after this node and before anything the right half appends.
    std::vector<b200rt_bvh_node> sub_nodes;
    std::vector<int32_t> sub_order;
    Ctx lcx{cx.prims, &sub_nodes, &sub_order, cx.grow, cx.par_levels - 1};
    std::future<uint32_t> left_task;
    try {
      left_task = std::async(std::launch::async, [&lcx, &lb, s, mid]() { return build_rec(lcx, s, mid, &lb); });
    } catch (const std::system_error&) {  // no thread to be had: same arrays, built here
      l = build_rec(lcx, s, mid, &lb);
    }
    std::vector<b200rt_bvh_node> r_nodes;
    std::vector<int32_t> r_order;
    Ctx rcx{cx.prims, &r_nodes, &r_order, cx.grow, cx.par_levels - 1};
    r = build_rec(rcx, mid, e, &rb);
    if (left_task.valid()) l = left_task.get();
    const uint32_t lnode_off = (uint32_t)cx.nodes->size(), lorder_off = (uint32_t)cx.order->size();
    const uint32_t rnode_off = lnode_off + (uint32_t)sub_nodes.size();
    const uint32_t rorder_off = lorder_off + (uint32_t)sub_order.size();
    l = shift_ref(l, lnode_off, lorder_off);
    r = shift_ref(r, rnode_off, rorder_off);
    splice(cx, sub_nodes, sub_order, lnode_off, lorder_off);
    splice(cx, r_nodes, r_order, rnode_off, rorder_off);
  } else {
    const int keep = cx.par_levels;
    cx.par_levels = 0;
    l = build_rec(cx, s, mid, &lb);
    r = build_rec(cx, mid, e, &rb);
    cx.par_levels = keep;
  }
  b200rt_bvh_node nd;
  std::memset(&nd, 0, sizeof(nd));
  put(nd.lmin, nd.lmax, lb, cx.grow);
  put(nd.rmin, nd.rmax, rb, cx.grow);
  nd.left = l;
  nd.right = r;
  (*cx.nodes)[me] = nd;
  return B200RT_REF(B200RT_KIND_FNODE, me);
}

void splice(Ctx& cx, const std::vector<b200rt_bvh_node>& nodes, const std::vector<int32_t>& order, uint32_t node_off,
            uint32_t order_off) {
  for (b200rt_bvh_node nd : nodes) {
    nd.left = shift_ref(nd.left, node_off, order_off);
    nd.right = shift_ref(nd.right, node_off, order_off);
    cx.nodes->push_back(nd);
  }
  cx.order->insert(cx.order->end(), order.begin(), order.end());
}

}  // namespace

uint32_t build(const float* boxes6, int n, int tri_base, float grow, std::vector<b200rt_bvh_node>& nodes,
               std::vector<int32_t>& order) {
  if (n <= 0) return B200RT_REF_NONE;
  if (const char* e = std::getenv("B200RT_SAH_MAXLEAF")) kMaxLeaf = std::max(1, std::min(4, std::atoi(e)));
  if (const char* e = std::getenv("B200RT_SAH_TCOST")) kTraversalCost = (float)std::atof(e);
  std::vector<Prim> prims((size_t)n);
  int par_levels = 0;
  {
    unsigned hc = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("B200RT_BUILD_THREADS")) hc = (unsigned)std::max(1, std::atoi(e));
    while ((1u << par_levels) < hc && par_levels < 6) par_levels++;
  }
  Ctx cx{prims.data(), &nodes, &order, grow, par_levels};
  for (int i = 0; i < n; i++) {
    Prim& p = prims[i];
    for (int k = 0; k < 3; k++) {
      p.mn[k] = boxes6[6 * (size_t)i + k];
      p.mx[k] = boxes6[6 * (size_t)i + 3 + k];
      p.c[k] = 0.5f * (p.mn[k] + p.mx[k]);
    }
    p.id = tri_base + i;
  }
  Bx root;
  return build_rec(cx, 0, n, &root);
}

}  // namespace b200rt_fast
