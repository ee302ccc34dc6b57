// gj_world_build: the world layout of include/gradjune_b200.h built from the REFERENCE'S OWN arrays — unsorted int64
// [2, E] edge lists per venue type, people per group, age, sex (june_world_loader/graph_loader.py:16-39,
// network_loader.py:30-44, runner.py:65-91) — behind the C ABI, on the GPU: radix sorts, scans, gathers and
// scatters (Thrust/CUB), no Python.  It performs, in this order,
//   1. the agent renumbering (households contiguous inside their leisure cell; see grad_june/world.py::layout_order,
//      of which this is the device implementation: the two produce identical arrays, tests/test_world_build.py),
//   2. the layout tiers (RANGE for the household type, CELL for "leisure", GENERIC CSR in both orientations for
//      the rest), the one-entry-per-agent view, the work lists of the group-major passes, the CTA tiles.
// The handle owns its device copies and is immutable after the build.
//
// The same code is instantiated for thrust's HOST backend (gj_world_build_host): that build runs without a GPU
// and exists so that the CPU test-suite can compare every array with the Python builder.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <thrust/adjacent_difference.h>
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/count.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/extrema.h>
#include <thrust/fill.h>
#include <thrust/gather.h>
#include <thrust/host_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/tuple.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/scatter.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/transform.h>
#include <thrust/unique.h>

#include "gradjune_b200.h"

namespace gjw {

constexpr int64_t kRangeMaxGroup = 64;      // == world.RANGE_MAX_GROUP
constexpr int64_t kCellMinMeanAgents = 64;  // == world.CELL_MIN_MEAN_AGENTS
constexpr int64_t kCellMaxGroups = 16;      // == world.CELL_MAX_GROUPS
constexpr int64_t kNoSlot = 0xFFFFFFFFll;
constexpr int64_t kEntMulti = 0xFFFFFFFEll;

struct DeviceBackend {
  template <class T>
  using vec = thrust::device_vector<T>;
  template <class T>
  static vec<T> load(const T* p, int64_t n) {
    vec<T> v(n);
    if (n > 0) thrust::copy(thrust::device_pointer_cast(p), thrust::device_pointer_cast(p) + n, v.begin());
    return v;
  }
};
struct HostBackend {
  template <class T>
  using vec = thrust::host_vector<T>;
  template <class T>
  static vec<T> load(const T* p, int64_t n) {
    return vec<T>(p, p + n);
  }
};

using cnt = thrust::counting_iterator<int64_t>;

template <class B>
struct Ops {
  using I64 = typename B::template vec<int64_t>;
  using F32 = typename B::template vec<float>;

  static I64 arange(int64_t n, int64_t start = 0) {
    I64 v(n);
    thrust::sequence(v.begin(), v.end(), start);
    return v;
  }
  static I64 gather(const I64& src, const I64& idx) {
    I64 out(idx.size());
    thrust::gather(idx.begin(), idx.end(), src.begin(), out.begin());
    return out;
  }
  static void scatter(I64& dst, const I64& idx, const I64& vals) {
    thrust::scatter(vals.begin(), vals.end(), idx.begin(), dst.begin());
  }
  static int64_t max_of(const I64& v) { return v.empty() ? 0 : (int64_t)*thrust::max_element(v.begin(), v.end()); }
  static int64_t min_of(const I64& v) { return v.empty() ? 0 : (int64_t)*thrust::min_element(v.begin(), v.end()); }
  static int64_t sum_of(const I64& v) { return thrust::reduce(v.begin(), v.end(), (int64_t)0); }
  // counts of each value in [0, n)
  static I64 bincount(const I64& idx, int64_t n) {
    I64 s = idx;
    thrust::sort(s.begin(), s.end());
    I64 ub(n);
    thrust::upper_bound(s.begin(), s.end(), cnt(0), cnt(n), ub.begin());
    I64 out(n);
    thrust::adjacent_difference(ub.begin(), ub.end(), out.begin());
    return out;
  }
  static I64 ptr_from_counts(const I64& c) {
    I64 p(c.size() + 1, 0);
    thrust::inclusive_scan(c.begin(), c.end(), p.begin() + 1);
    return p;
  }
  static I64 argsort_stable(const I64& keys) {
    I64 k = keys;
    I64 perm = arange((int64_t)keys.size());
    thrust::stable_sort_by_key(k.begin(), k.end(), perm.begin());
    return perm;
  }
  // indices i with mask[i] != 0
  static I64 nonzero(const I64& mask) {
    I64 out(thrust::count_if(mask.begin(), mask.end(), [] __host__ __device__(int64_t m) { return m != 0; }));
    thrust::copy_if(cnt(0), cnt((int64_t)mask.size()), mask.begin(), out.begin(),
                    [] __host__ __device__(int64_t m) { return m != 0; });
    return out;
  }
  // segment id of every element of a ragged layout with these counts (= repeat_interleave(arange(len), counts))
  static I64 segments(const I64& ptr) {
    const int64_t total = ptr.empty() ? 0 : (int64_t)ptr.back();
    I64 seg(total);
    thrust::upper_bound(ptr.begin() + 1, ptr.end(), cnt(0), cnt(total), seg.begin());
    return seg;
  }
  static I64 searchsorted_left(const I64& sorted, const I64& vals) {
    I64 out(vals.size());
    thrust::lower_bound(sorted.begin(), sorted.end(), vals.begin(), vals.end(), out.begin());
    return out;
  }
  template <class F>
  static I64 map1(const I64& a, F f) {
    I64 out(a.size());
    thrust::transform(a.begin(), a.end(), out.begin(), f);
    return out;
  }
  template <class F>
  static I64 map2(const I64& a, const I64& b, F f) {
    I64 out(a.size());
    thrust::transform(a.begin(), a.end(), b.begin(), out.begin(), f);
    return out;
  }
  static I64 cat(const I64& a, const I64& b) {
    I64 out(a.size() + b.size());
    thrust::copy(a.begin(), a.end(), out.begin());
    thrust::copy(b.begin(), b.end(), out.begin() + a.size());
    return out;
  }
  static I64 select(const I64& v, const I64& mask) { return gather(v, nonzero(mask)); }
  // dense rank of every key (ascending) and the number of distinct keys
  static I64 unique_inverse(const I64& keys, int64_t* n_unique) {
    const int64_t n = (int64_t)keys.size();
    I64 k = keys;
    I64 perm = arange(n);
    thrust::sort_by_key(k.begin(), k.end(), perm.begin());
    I64 fresh(n, 0);
    if (n > 0) {
      thrust::transform(k.begin() + 1, k.end(), k.begin(), fresh.begin() + 1,
                        [] __host__ __device__(int64_t a, int64_t b) { return (int64_t)(a != b); });
      thrust::inclusive_scan(fresh.begin(), fresh.end(), fresh.begin());
    }
    I64 rank(n);
    thrust::scatter(fresh.begin(), fresh.end(), perm.begin(), rank.begin());
    *n_unique = n > 0 ? (int64_t)fresh.back() + 1 : 0;
    return rank;
  }
};

struct TypeSrc {
  std::string name;
  int64_t E, G;
};

template <class B>
struct World {
  template <class T>
  using vec = typename B::template vec<T>;
  gj_world_desc desc;
  vec<uint32_t> am_ptr, am_ent, gm_ptr, gm_agent, small_groups, chunk_group, chunk_begin, chunk_end, big_groups,
      big_part_ptr, tile_begin, tile_flags, ent1, orig_id;
  vec<int32_t> chunk_part;
  vec<float> pc;
  vec<uint8_t> cls;
  vec<uint32_t> range_slot[GJ_MAX_TYPES];
  vec<float> range_pc[GJ_MAX_TYPES];
  vec<uint32_t> tile_cell[GJ_MAX_TYPES], cell_tile_ptr[GJ_MAX_TYPES], cell_grp_ptr[GJ_MAX_TYPES], cell_grp[GJ_MAX_TYPES],
      grp_cell_ptr[GJ_MAX_TYPES], grp_cell[GJ_MAX_TYPES];
  vec<int64_t> perm;   // new -> old; empty = identity
  std::string error;
};

template <class B>
struct Builder {
  using O = Ops<B>;
  using I64 = typename O::I64;
  template <class T>
  using vec = typename B::template vec<T>;

  // uint32 bit patterns with `slack` readable elements behind (bulk copies read 16-byte granules)
  static vec<uint32_t> to_u32(const I64& v, int slack = 0) {
    vec<uint32_t> out(v.size() + slack, 0u);
    thrust::transform(v.begin(), v.end(), out.begin(), [] __host__ __device__(int64_t x) { return (uint32_t)x; });
    return out;
  }
  template <class T>
  static const T* raw(const vec<T>& v) {
    return v.empty() ? nullptr : thrust::raw_pointer_cast(v.data());
  }

  struct Cell {
    int64_t n_cells = 0;
    I64 cell_start, cell_of_agent, cell_grp_ptr, cell_grp, grp_cell_ptr, grp_cell;
  };

  // grad_june/world.py::_list_rank
  static I64 list_rank(int64_t n, const I64& src, const I64& dst, int64_t G, int64_t* n_lists) {
    I64 deg = O::bincount(src, n);
    I64 perm = O::argsort_stable(src);
    I64 s_sorted = O::gather(src, perm), d_sorted = O::gather(dst, perm);
    I64 aptr = O::ptr_from_counts(deg);
    I64 pos = O::map2(O::arange((int64_t)src.size()), O::gather(aptr, s_sorted),
                      [] __host__ __device__(int64_t i, int64_t p) { return i - p; });
    I64 rank(n, 0);
    *n_lists = 1;
    const int64_t kmax = O::max_of(deg);
    for (int64_t j = 0; j < kmax; ++j) {
      I64 gj(n, 0);
      I64 sel = O::map1(pos, [j] __host__ __device__(int64_t p) { return (int64_t)(p == j); });
      I64 at = O::nonzero(sel);
      I64 d1 = O::map1(O::gather(d_sorted, at), [] __host__ __device__(int64_t d) { return d + 1; });
      O::scatter(gj, O::gather(s_sorted, at), d1);
      I64 key = O::map2(rank, gj, [G] __host__ __device__(int64_t r, int64_t g) { return r * (G + 1) + g; });
      rank = O::unique_inverse(key, n_lists);
    }
    return rank;
  }

  // (max degree, max group size) of a type; (0, 0) without edges
  static void shape_of(int64_t n, const I64& src, const I64& dst, int64_t G, int64_t* deg_max, int64_t* size_max) {
    *deg_max = *size_max = 0;
    if (src.empty()) return;
    *deg_max = O::max_of(O::bincount(src, n));
    *size_max = O::max_of(O::bincount(dst, G));
  }

  // grad_june/world.py::tier_candidates
  static void tier_candidates(int64_t n, const std::vector<TypeSrc>& types, const std::vector<I64>& src,
                              const std::vector<I64>& dst, int* range_type, int* cell_type) {
    *range_type = *cell_type = -1;
    std::vector<int64_t> dmax(types.size()), smax(types.size());
    for (size_t t = 0; t < types.size(); ++t) shape_of(n, src[t], dst[t], types[t].G, &dmax[t], &smax[t]);
    for (size_t t = 0; t < types.size(); ++t)
      if (types[t].name == "leisure" && dmax[t] > 0 && dmax[t] <= kCellMaxGroups) *cell_type = (int)t;
    int best = -1;
    for (size_t t = 0; t < types.size(); ++t) {
      if ((int)t == *cell_type || dmax[t] != 1 || smax[t] > kRangeMaxGroup) continue;
      if (types[t].name == "household") {
        best = (int)t;
        break;
      }
      if (best < 0 || types[t].E > types[best].E) best = (int)t;
    }
    *range_type = best;
  }

  // grad_june/world.py::layout_order ; empty result = the given numbering already has the layout
  static I64 layout_order(int64_t n, const std::vector<TypeSrc>& types, const std::vector<I64>& src,
                          const std::vector<I64>& dst) {
    int rt, ct;
    tier_candidates(n, types, src, dst, &rt, &ct);
    if (n == 0 || (rt < 0 && ct < 0)) return I64();
    I64 ids = O::arange(n);
    I64 cell(n, 0);
    int64_t n_cells = 1;
    if (ct >= 0) cell = list_rank(n, src[ct], dst[ct], types[ct].G, &n_cells);
    int64_t Gh = 0;
    I64 hkey = ids, epos = ids;
    if (rt >= 0) {
      Gh = types[rt].G;
      const int64_t E = types[rt].E;
      hkey = O::map1(ids, [Gh] __host__ __device__(int64_t a) { return Gh + a; });
      O::scatter(hkey, src[rt], dst[rt]);
      epos = O::map1(ids, [E] __host__ __device__(int64_t a) { return E + a; });
      O::scatter(epos, src[rt], O::arange(E));
    }
    const int64_t n_keys = Gh + n;
    // first[h] = smallest edge position among the members of (pseudo-)household h (a reduce-by-key: scatters with
    // duplicate indices have no defined winner)
    I64 order = O::argsort_stable(epos);               // members in edge order
    I64 first(n_keys, O::max_of(epos) + 1);
    {
      I64 hk = hkey, ep = epos;
      thrust::sort_by_key(hk.begin(), hk.end(), ep.begin());
      I64 uk(n), um(n);
      auto end = thrust::reduce_by_key(hk.begin(), hk.end(), ep.begin(), uk.begin(), um.begin(),
                                       thrust::equal_to<int64_t>(), thrust::minimum<int64_t>());
      const int64_t m = end.first - uk.begin();
      uk.resize(m);
      um.resize(m);
      O::scatter(first, uk, um);
    }
    I64 hcell(n_keys, n_cells);
    I64 is_first = O::map2(epos, O::gather(first, hkey), [] __host__ __device__(int64_t e, int64_t f) { return (int64_t)(e == f); });
    I64 at = O::nonzero(is_first);
    O::scatter(hcell, O::gather(hkey, at), O::gather(cell, at));
    I64 base = order;                                  // (3): members in edge order
    I64 key = O::map2(O::gather(hcell, hkey), hkey, [n_keys] __host__ __device__(int64_t c, int64_t h) { return c * n_keys + h; });
    I64 kb = O::gather(key, base);
    I64 ord2 = O::argsort_stable(kb);
    I64 perm = O::gather(base, ord2);
    if (thrust::equal(perm.begin(), perm.end(), ids.begin())) return I64();
    return perm;
  }

  // grad_june/world.py::_try_range_tier
  static bool try_range(int64_t n, const I64& src, const I64& dst, int64_t G, const vec<float>& pc, int64_t pc_off,
                        I64* slot_out, vec<float>* rpc_out, bool* from_size) {
    const int64_t E = (int64_t)src.size();
    if (E == 0) return false;
    if (O::max_of(O::bincount(src, n)) > 1) return false;
    I64 size = O::bincount(dst, G);
    if (O::max_of(size) > kRangeMaxGroup) return false;
    I64 perm = O::argsort_stable(dst);
    I64 members = O::gather(src, perm), gid = O::gather(dst, perm);
    I64 ptr = O::ptr_from_counts(size);
    I64 p0(ptr.begin(), ptr.end() - 1);
    I64 p0c = O::map1(p0, [E] __host__ __device__(int64_t p) { return p < E - 1 ? p : E - 1; });
    I64 first = O::gather(members, p0c);
    I64 offset = O::map2(members, O::gather(first, gid), [] __host__ __device__(int64_t m, int64_t f) { return m - f; });
    I64 expect = O::map2(O::arange(E), O::gather(ptr, gid), [] __host__ __device__(int64_t i, int64_t p) { return i - p; });
    if (!thrust::equal(offset.begin(), offset.end(), expect.begin())) return false;
    I64 slot(n, kNoSlot);
    I64 val = O::map2(offset, O::gather(size, gid), [] __host__ __device__(int64_t o, int64_t s) { return (o << 16) | s; });
    O::scatter(slot, members, val);
    vec<float> rpc(n + 32, 0.0f);
    thrust::scatter(thrust::make_permutation_iterator(pc.begin() + pc_off, gid.begin()),
                    thrust::make_permutation_iterator(pc.begin() + pc_off, gid.end()), members.begin(), rpc.begin());
    {   // does `people` equal the member count (contact probability derivable from the slot's size)?
      vec<float> have(E), want(E);
      thrust::copy(thrust::make_permutation_iterator(pc.begin() + pc_off, gid.begin()),
                   thrust::make_permutation_iterator(pc.begin() + pc_off, gid.end()), have.begin());
      I64 sz = O::gather(size, gid);
      thrust::transform(sz.begin(), sz.end(), want.begin(), [] __host__ __device__(int64_t x) {
        const float v = 1.0f / (float)(x - 1);
        return fmaxf(fminf(v, 1.0f), 0.0f);
      });
      *from_size = thrust::equal(have.begin(), have.end(), want.begin());
    }
    *slot_out = slot;
    *rpc_out = rpc;
    return true;
  }

  // grad_june/world.py::_try_cell_tier
  static bool try_cell(int64_t n, const I64& src, const I64& dst, int64_t G, Cell* c) {
    const int64_t E = (int64_t)src.size();
    if (E == 0) return false;
    I64 deg = O::bincount(src, n);
    if (O::max_of(deg) > kCellMaxGroups) return false;
    I64 perm = O::argsort_stable(src);
    I64 s_sorted = O::gather(src, perm), d_sorted = O::gather(dst, perm);
    I64 aptr = O::ptr_from_counts(deg);
    I64 pos = O::map2(O::arange(E), O::gather(aptr, s_sorted), [] __host__ __device__(int64_t i, int64_t p) { return i - p; });
    I64 same_deg(n, 0);
    if (n > 1)
      thrust::transform(deg.begin() + 1, deg.end(), deg.begin(), same_deg.begin() + 1,
                        [] __host__ __device__(int64_t a, int64_t b) { return (int64_t)(a == b); });
    I64 sm1 = O::map1(s_sorted, [] __host__ __device__(int64_t s) { return s > 0 ? s - 1 : 0; });
    I64 prev_idx = O::map2(O::gather(aptr, sm1), pos, [E] __host__ __device__(int64_t p, int64_t q) {
      const int64_t v = p + q;
      return v < E - 1 ? v : E - 1;
    });
    I64 dprev = O::gather(d_sorted, prev_idx);
    I64 sd = O::gather(same_deg, s_sorted);
    I64 differs(E);
    {
      I64 ne = O::map2(d_sorted, dprev, [] __host__ __device__(int64_t a, int64_t b) { return (int64_t)(a != b); });
      differs = O::map2(ne, sd, [] __host__ __device__(int64_t a, int64_t b) { return a & b; });
    }
    // n_diff per agent: s_sorted is sorted, so a segmented sum
    I64 csum(E + 1, 0);
    thrust::inclusive_scan(differs.begin(), differs.end(), csum.begin() + 1);
    I64 lo(aptr.begin(), aptr.end() - 1), hi(aptr.begin() + 1, aptr.end());
    I64 n_diff = O::map2(O::gather(csum, hi), O::gather(csum, lo), [] __host__ __device__(int64_t a, int64_t b) { return a - b; });
    I64 boundary = O::map2(same_deg, n_diff, [] __host__ __device__(int64_t s, int64_t d) { return (int64_t)(s == 0 || d > 0); });
    const int64_t n_cells = O::sum_of(boundary);
    if (n_cells * kCellMinMeanAgents > n) return false;
    c->n_cells = n_cells;
    c->cell_start = O::nonzero(boundary);
    c->cell_of_agent = I64(n);
    thrust::inclusive_scan(boundary.begin(), boundary.end(), c->cell_of_agent.begin());
    thrust::transform(c->cell_of_agent.begin(), c->cell_of_agent.end(), c->cell_of_agent.begin(),
                      [] __host__ __device__(int64_t x) { return x - 1; });
    I64 cdeg = O::gather(deg, c->cell_start);
    c->cell_grp_ptr = O::ptr_from_counts(cdeg);
    I64 seg = O::segments(c->cell_grp_ptr);
    const int64_t total = (int64_t)seg.size();
    I64 a0 = O::gather(O::gather(aptr, c->cell_start), seg);
    I64 g0 = O::gather(c->cell_grp_ptr, seg);
    I64 idx(total);
    {
      I64 ar = O::arange(total);
      I64 w = O::map2(ar, g0, [] __host__ __device__(int64_t i, int64_t g) { return i - g; });
      idx = O::map2(a0, w, [] __host__ __device__(int64_t a, int64_t b) { return a + b; });
    }
    c->cell_grp = O::gather(d_sorted, idx);
    I64 gperm = O::argsort_stable(c->cell_grp);
    c->grp_cell = O::gather(seg, gperm);
    c->grp_cell_ptr = O::ptr_from_counts(O::bincount(c->cell_grp, G));
    return true;
  }
};

struct SrcView {   // the caller's arrays, loaded into the backend's memory
  int64_t n_agents;
  std::vector<TypeSrc> types;
};

template <class B>
static int build(const gj_world_src* s, World<B>* W) {
  using Bd = Builder<B>;
  using O = Ops<B>;
  using I64 = typename O::I64;
  const int64_t n = s->n_agents;
  const int nt = s->n_types;
  if (n < 0 || nt < 0 || nt > GJ_MAX_TYPES) {
    W->error = "n_agents / n_types out of range";
    return -1;
  }
  const int64_t small_group = GJ_SMALL_GROUP, chunk = GJ_CHUNK, scatter_max = GJ_SCATTER_MAX_GROUP;
  std::vector<TypeSrc> types(nt);
  std::vector<I64> src(nt), dst(nt);
  int64_t n_edges_total = 0;
  for (int t = 0; t < nt; ++t) {
    types[t].name = s->type_name[t] ? s->type_name[t] : "";
    types[t].E = s->n_edges[t];
    types[t].G = s->n_groups[t];
    if (types[t].E < 0 || types[t].G < 0 || types[t].G >= ((int64_t)1 << 28)) {
      W->error = "edge type " + types[t].name + ": sizes out of range";
      return -1;
    }
    src[t] = B::template load<int64_t>(s->edge_agent[t], types[t].E);
    dst[t] = B::template load<int64_t>(s->edge_group[t], types[t].E);
    if (types[t].E > 0) {
      if (O::min_of(src[t]) < 0 || O::max_of(src[t]) >= n) {
        W->error = "edge type " + types[t].name + ": agent index out of range";
        return -1;
      }
      if (O::min_of(dst[t]) < 0 || O::max_of(dst[t]) >= types[t].G) {
        W->error = "edge type " + types[t].name + ": group index out of range";
        return -1;
      }
    }
    n_edges_total += types[t].E;
  }
  if (n_edges_total >= ((int64_t)1 << 32) || n >= ((int64_t)1 << 32)) {
    W->error = "world too large for 32-bit CSR offsets";
    return -1;
  }
  I64 age = B::template load<int64_t>(s->age, n), sex = B::template load<int64_t>(s->sex, n);
  if (n > 0 && (O::min_of(age) < 0 || O::max_of(age) > 99 || O::min_of(sex) < 0 || O::max_of(sex) > 1)) {
    W->error = "age must be in [0, 99] and sex in {0, 1}";
    return -1;
  }

  // ---- 1. renumbering ---------------------------------------------------------------------------------
  I64 orig;   // id of every agent in the numbering the world was LOADED in (empty = this one)
  if (s->original_index) orig = B::template load<int64_t>(s->original_index, n);
  if (s->renumber) {
    I64 perm = Bd::layout_order(n, types, src, dst);
    if (!perm.empty()) {
      I64 inv(n);
      O::scatter(inv, perm, O::arange(n));
      for (int t = 0; t < nt; ++t) src[t] = O::gather(inv, src[t]);
      age = O::gather(age, perm);
      sex = O::gather(sex, perm);
      orig = orig.empty() ? perm : O::gather(orig, perm);
      W->perm = perm;
    }
  }

  // ---- 2. tiers ----------------------------------------------------------------------------------------
  gj_world_desc& d = W->desc;
  memset(&d, 0, sizeof(d));
  std::vector<int64_t> offs(nt + 1, 0);
  for (int t = 0; t < nt; ++t) offs[t + 1] = offs[t] + types[t].G;
  const int64_t G = offs[nt];
  if (G >= ((int64_t)1 << 31)) {
    W->error = "too many groups";
    return -1;
  }
  {   // pc = clamp(1 / (people - 1), 0, 1)   (infection_networks/base.py:64-69, fp32)
    typename B::template vec<float> pc(G);
    for (int t = 0; t < nt; ++t) {
      if (s->people_i64[t]) {
        I64 p = B::template load<int64_t>(s->people_i64[t], types[t].G);
        thrust::transform(p.begin(), p.end(), pc.begin() + offs[t], [] __host__ __device__(int64_t x) {
          const float v = 1.0f / (float)(x - 1);
          return fmaxf(fminf(v, 1.0f), 0.0f);
        });
      } else if (s->people_f32[t]) {
        auto p = B::template load<float>(s->people_f32[t], types[t].G);
        thrust::transform(p.begin(), p.end(), pc.begin() + offs[t], [] __host__ __device__(float x) {
          const float v = 1.0f / (x - 1.0f);
          return fmaxf(fminf(v, 1.0f), 0.0f);
        });
      } else {
        W->error = "edge type " + types[t].name + ": people is NULL";
        return -1;
      }
    }
    W->pc = pc;
  }
  int range_type = -1, cell_type = -1;
  if (n > 0) Bd::tier_candidates(n, types, src, dst, &range_type, &cell_type);
  std::vector<int> want(nt, GJ_TIER_GENERIC);
  for (int t = 0; t < nt; ++t) {
    if (s->want_tier) want[t] = s->want_tier[t];
    else if (t == range_type) want[t] = GJ_TIER_RANGE;
    else if (t == cell_type) want[t] = GJ_TIER_CELL;
  }
  std::vector<typename Bd::Cell> cells(nt);
  std::vector<int> tier(nt, GJ_TIER_GENERIC);
  I64 g_src, g_key, g_ent;
  for (int t = 0; t < nt; ++t) {
    if (n > 0 && want[t] == GJ_TIER_RANGE) {
      I64 slot;
      typename B::template vec<float> rpc;
      bool from_size = false;
      if (Bd::try_range(n, src[t], dst[t], types[t].G, W->pc, offs[t], &slot, &rpc, &from_size)) {
        tier[t] = GJ_TIER_RANGE;
        d.range_pc_from_size[t] = from_size ? 1 : 0;
        W->range_slot[t] = Bd::to_u32(slot, 32);
        W->range_pc[t] = rpc;
      }
    } else if (n > 0 && want[t] == GJ_TIER_CELL) {
      if (Bd::try_cell(n, src[t], dst[t], types[t].G, &cells[t])) tier[t] = GJ_TIER_CELL;
    }
    if (tier[t] == GJ_TIER_GENERIC) {
      const int64_t off = offs[t], tt = t;
      g_src = O::cat(g_src, src[t]);
      g_key = O::cat(g_key, O::map1(dst[t], [off] __host__ __device__(int64_t g) { return g + off; }));
      g_ent = O::cat(g_ent, O::map1(dst[t], [tt] __host__ __device__(int64_t g) { return g + (tt << 28); }));
    }
  }
  const int64_t E = (int64_t)g_src.size();

  // ---- generic tier: CSR in both orientations, one-entry view ---------------------------------------------
  I64 size = O::bincount(g_key, G);
  I64 gm_ptr = O::ptr_from_counts(size);
  {
    I64 perm = O::argsort_stable(g_key);   // stable: the reference's edge order inside each group
    W->gm_agent = Bd::to_u32(O::gather(g_src, perm));
  }
  I64 deg = O::bincount(g_src, n);
  I64 am_ptr = O::ptr_from_counts(deg);
  {
    I64 perm = O::argsort_stable(g_src);   // types were concatenated in order: (type, edge order) per agent
    W->am_ent = Bd::to_u32(O::gather(g_ent, perm), 32);
  }
  {
    I64 ent1(n, kNoSlot);
    I64 v = O::map2(g_key, O::gather(size, g_key), [scatter_max] __host__ __device__(int64_t k, int64_t sz) {
      return k + ((int64_t)(sz > scatter_max) << 31);
    });
    O::scatter(ent1, g_src, v);
    thrust::transform(ent1.begin(), ent1.end(), deg.begin(), ent1.begin(),
                      [] __host__ __device__(int64_t e, int64_t dg) { return dg > 1 ? kEntMulti : e; });
    W->ent1 = Bd::to_u32(ent1, 32);
  }
  {
    typename B::template vec<uint8_t> cls(n + 64, (uint8_t)0);
    thrust::transform(sex.begin(), sex.end(), age.begin(), cls.begin(),
                      [] __host__ __device__(int64_t sx, int64_t ag) { return (uint8_t)(sx * 100 + ag); });
    W->cls = cls;
  }

  // ---- work lists of the group-major passes (generic types only), giant groups first -------------------------
  I64 generic_group(G, 0);
  for (int t = 0; t < nt; ++t)
    if (tier[t] == GJ_TIER_GENERIC) thrust::fill(generic_group.begin() + offs[t], generic_group.begin() + offs[t + 1], (int64_t)1);
  I64 gids = O::arange(G);
  I64 small_mask = O::map2(size, generic_group, [small_group] __host__ __device__(int64_t sz, int64_t g) { return (int64_t)(sz <= small_group && g); });
  I64 big_giant = O::map2(size, generic_group, [small_group, scatter_max] __host__ __device__(int64_t sz, int64_t g) { return (int64_t)(sz > small_group && sz > scatter_max && g); });
  I64 big_rest = O::map2(size, generic_group, [small_group, scatter_max] __host__ __device__(int64_t sz, int64_t g) { return (int64_t)(sz > small_group && sz <= scatter_max && g); });
  I64 small = O::select(gids, small_mask);
  I64 bg_g = O::select(gids, big_giant);
  I64 bg = O::cat(bg_g, O::select(gids, big_rest));
  const int64_t n_giant_big = (int64_t)bg_g.size();
  I64 nchunk = O::map1(O::gather(size, bg), [chunk] __host__ __device__(int64_t sz) { return (sz + chunk - 1) / chunk; });
  int64_t n_giant_chunks = 0;
  if (n_giant_big > 0) n_giant_chunks = thrust::reduce(nchunk.begin(), nchunk.begin() + n_giant_big, (int64_t)0);
  I64 cfirst = O::ptr_from_counts(nchunk);
  I64 cseg = O::segments(cfirst);
  I64 chunk_group = O::gather(bg, cseg);
  const int64_t n_chunks = (int64_t)chunk_group.size();
  I64 within = O::map2(O::arange(n_chunks), O::gather(cfirst, cseg), [] __host__ __device__(int64_t i, int64_t f) { return i - f; });
  I64 chunk_begin = O::map2(O::gather(gm_ptr, chunk_group), within, [chunk] __host__ __device__(int64_t p, int64_t w) { return p + w * chunk; });
  I64 cg1 = O::map1(chunk_group, [] __host__ __device__(int64_t g) { return g + 1; });
  I64 chunk_end = O::map2(chunk_begin, O::gather(gm_ptr, cg1), [chunk] __host__ __device__(int64_t b, int64_t e) { return b + chunk < e ? b + chunk : e; });
  I64 multi_g = O::map1(nchunk, [] __host__ __device__(int64_t c) { return (int64_t)(c > 1); });
  I64 multi = O::gather(multi_g, cseg);
  I64 part_idx(n_chunks, 0);
  thrust::exclusive_scan(multi.begin(), multi.end(), part_idx.begin());
  const int64_t n_parts = O::sum_of(multi);
  {
    typename B::template vec<int32_t> chunk_part(n_chunks);
    thrust::transform(multi.begin(), multi.end(), part_idx.begin(), chunk_part.begin(),
                      [] __host__ __device__(int64_t m, int64_t i) { return m ? (int32_t)i : (int32_t)-1; });
    W->chunk_part = chunk_part;
  }
  I64 big_groups = O::select(bg, multi_g);
  I64 big_part_ptr = O::ptr_from_counts(O::select(nchunk, multi_g));

  // ---- CTA tiles: <= GJ_TILE_AGENTS consecutive agents inside one cell of every CELL type ----------------------
  I64 cuts(1, 0);
  for (int t = 0; t < nt; ++t)
    if (tier[t] == GJ_TIER_CELL) cuts = O::cat(cuts, cells[t].cell_start);
  I64 tile_begin;
  if (n > 0) {
    thrust::sort(cuts.begin(), cuts.end());
    cuts.resize(thrust::unique(cuts.begin(), cuts.end()) - cuts.begin());
    I64 seg_start = cuts;
    I64 seg_end(seg_start.size());
    thrust::copy(seg_start.begin() + 1, seg_start.end(), seg_end.begin());
    seg_end[seg_end.size() - 1] = n;
    const int64_t T = GJ_TILE_AGENTS;
    I64 ntile = O::map2(seg_end, seg_start, [T] __host__ __device__(int64_t e, int64_t b) { return (e - b + T - 1) / T; });
    I64 tfirst = O::ptr_from_counts(ntile);
    I64 tseg = O::segments(tfirst);
    const int64_t n_tiles = (int64_t)tseg.size();
    I64 tw = O::map2(O::arange(n_tiles), O::gather(tfirst, tseg), [] __host__ __device__(int64_t i, int64_t f) { return i - f; });
    tile_begin = O::map2(O::gather(seg_start, tseg), tw, [T] __host__ __device__(int64_t b, int64_t w) { return b + w * T; });
    tile_begin.push_back(n);
  } else {
    tile_begin = I64(1, 0);
  }
  const int64_t n_tiles = (int64_t)tile_begin.size() - 1;
  I64 tile_flags(n_tiles, 0);
  if (n_tiles > 0) tile_flags[0] = 1;
  int64_t cell_off = 0;
  for (int t = 0; t < nt; ++t) {
    d.type_tier[t] = tier[t];
    d.cell_off[t] = cell_off;
    if (tier[t] == GJ_TIER_RANGE) {
      d.range_slot[t] = Bd::raw(W->range_slot[t]);
      d.range_pc[t] = Bd::raw(W->range_pc[t]);
    } else if (tier[t] == GJ_TIER_CELL) {
      auto& c = cells[t];
      I64 tb0(tile_begin.begin(), tile_begin.end() - 1);
      I64 tc = O::gather(c.cell_of_agent, tb0);
      if (n_tiles > 1)
        thrust::transform(thrust::make_zip_iterator(thrust::make_tuple(tc.begin() + 1, tc.begin(), tile_flags.begin() + 1)),
                          thrust::make_zip_iterator(thrust::make_tuple(tc.end(), tc.end() - 1, tile_flags.end())),
                          tile_flags.begin() + 1, [] __host__ __device__(thrust::tuple<int64_t, int64_t, int64_t> x) {
                            return thrust::get<2>(x) | (int64_t)(thrust::get<0>(x) != thrust::get<1>(x));
                          });
      W->tile_cell[t] = Bd::to_u32(tc);
      I64 ctp = O::searchsorted_left(tb0, c.cell_start);
      ctp.push_back(n_tiles);
      W->cell_tile_ptr[t] = Bd::to_u32(ctp);
      W->cell_grp_ptr[t] = Bd::to_u32(c.cell_grp_ptr);
      W->cell_grp[t] = Bd::to_u32(c.cell_grp);
      W->grp_cell_ptr[t] = Bd::to_u32(c.grp_cell_ptr);
      W->grp_cell[t] = Bd::to_u32(c.grp_cell);
      d.n_cells[t] = c.n_cells;
      d.tile_cell[t] = Bd::raw(W->tile_cell[t]);
      d.cell_tile_ptr[t] = Bd::raw(W->cell_tile_ptr[t]);
      d.cell_grp_ptr[t] = Bd::raw(W->cell_grp_ptr[t]);
      d.cell_grp[t] = Bd::raw(W->cell_grp[t]);
      d.grp_cell_ptr[t] = Bd::raw(W->grp_cell_ptr[t]);
      d.grp_cell[t] = Bd::raw(W->grp_cell[t]);
      cell_off += c.n_cells;
    }
  }
  d.n_cells_total = cell_off;

  W->am_ptr = Bd::to_u32(am_ptr, 32);
  W->gm_ptr = Bd::to_u32(gm_ptr);
  W->small_groups = Bd::to_u32(small);
  W->chunk_group = Bd::to_u32(chunk_group);
  W->chunk_begin = Bd::to_u32(chunk_begin);
  W->chunk_end = Bd::to_u32(chunk_end);
  W->big_groups = Bd::to_u32(big_groups);
  W->big_part_ptr = Bd::to_u32(big_part_ptr);
  W->tile_begin = Bd::to_u32(tile_begin, 4);
  W->tile_flags = Bd::to_u32(tile_flags, 4);
  if (!orig.empty()) W->orig_id = Bd::to_u32(orig, 32);

  d.n_agents = n;
  d.n_groups = G;
  d.n_edges = n_edges_total;
  d.n_types = nt;
  for (int t = 0; t <= nt; ++t) d.type_group_off[t] = offs[t];
  d.am_ptr = Bd::raw(W->am_ptr);
  d.am_ent = Bd::raw(W->am_ent);
  d.gm_ptr = Bd::raw(W->gm_ptr);
  d.gm_agent = Bd::raw(W->gm_agent);
  d.pc = Bd::raw(W->pc);
  d.cls = Bd::raw(W->cls);
  d.small_groups = Bd::raw(W->small_groups);
  d.n_small = (int64_t)small.size();
  d.chunk_group = Bd::raw(W->chunk_group);
  d.chunk_begin = Bd::raw(W->chunk_begin);
  d.chunk_end = Bd::raw(W->chunk_end);
  d.chunk_part = Bd::raw(W->chunk_part);
  d.n_chunks = n_chunks;
  d.big_groups = Bd::raw(W->big_groups);
  d.big_part_ptr = Bd::raw(W->big_part_ptr);
  d.n_big = (int64_t)big_groups.size();
  d.n_parts = n_parts;
  d.n_tiles = n_tiles;
  d.tile_begin = Bd::raw(W->tile_begin);
  d.tile_flags = Bd::raw(W->tile_flags);
  d.ent1 = Bd::raw(W->ent1);
  d.n_giant_chunks = n_giant_chunks;
  d.n_giant_big = n_giant_big;
  d.dbeta_w = nullptr;
  d.orig_id = Bd::raw(W->orig_id);
  (void)E;
  return 0;
}

}  // namespace gjw

struct gj_world {
  gjw::World<gjw::DeviceBackend>* dev = nullptr;
  gjw::World<gjw::HostBackend>* host = nullptr;
  thrust::host_vector<int64_t> perm_host;   // copy for gj_world_original_index of a host build
};

static thread_local char g_world_err[512] = "";

extern "C" {

const char* gj_world_last_error(void) { return g_world_err; }

static int build_any(const gj_world_src* src, gj_world** out, bool on_device) {
  if (!src || !out) {
    snprintf(g_world_err, sizeof(g_world_err), "gj_world_build: NULL argument");
    return -1;
  }
  gj_world* w = new gj_world();
  int rc = 0;
  try {
    if (on_device) {
      w->dev = new gjw::World<gjw::DeviceBackend>();
      rc = gjw::build<gjw::DeviceBackend>(src, w->dev);
      if (rc) snprintf(g_world_err, sizeof(g_world_err), "gj_world_build: %s", w->dev->error.c_str());
    } else {
      w->host = new gjw::World<gjw::HostBackend>();
      rc = gjw::build<gjw::HostBackend>(src, w->host);
      if (rc) snprintf(g_world_err, sizeof(g_world_err), "gj_world_build: %s", w->host->error.c_str());
    }
  } catch (const std::exception& e) {
    snprintf(g_world_err, sizeof(g_world_err), "gj_world_build: %s", e.what());
    rc = -2;
  }
  if (rc) {
    delete w->dev;
    delete w->host;
    delete w;
    return rc;
  }
  *out = w;
  return 0;
}

int gj_world_build(const gj_world_src* src, gj_world** out) { return build_any(src, out, true); }
int gj_world_build_host(const gj_world_src* src, gj_world** out) { return build_any(src, out, false); }

const gj_world_desc* gj_world_descriptor(const gj_world* w) {
  if (!w) return nullptr;
  return w->dev ? &w->dev->desc : &w->host->desc;
}

const int64_t* gj_world_permutation(const gj_world* w) {
  if (!w) return nullptr;
  if (w->dev) return w->dev->perm.empty() ? nullptr : thrust::raw_pointer_cast(w->dev->perm.data());
  return w->host->perm.empty() ? nullptr : thrust::raw_pointer_cast(w->host->perm.data());
}

// copy out of (or into) a handle's arrays: any direction between host and device memory (unified addressing)
int gj_memcpy(void* dst, const void* src, int64_t bytes) {
  if (bytes <= 0) return 0;
  if (!dst || !src) {
    snprintf(g_world_err, sizeof(g_world_err), "gj_memcpy: NULL argument");
    return -1;
  }
  const cudaError_t e = cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDefault);
  if (e != cudaSuccess) {
    snprintf(g_world_err, sizeof(g_world_err), "gj_memcpy: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int gj_world_destroy(gj_world* w) {
  if (!w) return 0;
  delete w->dev;
  delete w->host;
  delete w;
  return 0;
}

}  // extern "C"
