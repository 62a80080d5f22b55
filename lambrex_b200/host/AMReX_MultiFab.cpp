// amrex-mini: device field containers, host tag boxes, and the gather-plan builders that stand
// in for AMReX's ParallelCopy family (see AMReX_MultiFab.H, AMReX_FillPatch.H).
#include <chrono>
#include <iostream>
#include <cstdlib>
#include "AMReX_MultiFab.H"

#include <cstring>
#include <functional>
#include <map>
#include <unordered_map>

#include "AMReX_FillPatch.H"

namespace amrex {

void lbx_check(int rc, const char* where) {
  if (rc != 0) Abort(std::string(where) + ": " + lbx_last_error());
}

// ----------------------------------------------------------------------------- DeviceFabArray
template <class T, int DTYPE>
DeviceFabArray<T, DTYPE>::Storage::~Storage() {
  if (mf) lbx_mf_destroy(mf);
}

template <class T, int DTYPE>
void DeviceFabArray<T, DTYPE>::define(const BoxArray& ba, const DistributionMapping& dm, int ncomp, int ngrow,
                                      Layout lay) {
  clear();
  ba_ = ba;
  dm_ = dm;
  ncomp_ = ncomp;
  ngrow_ = ngrow;
  lay_ = lay;
  if (ba.empty()) return;
  std::vector<lbx_box> vb;
  std::vector<int> slab_owner;
  int sg = ngrow;
  slabs_.clear();
  if (lay == Layout::FLAT) {
    const Box mb = ba.minimalBox();
    if (mb.numPts() != ba.numPts()) Abort("FLAT layout needs a BoxArray that tiles its minimal box");
    const int np = DistributionMapping::NProcs();
    if (np > 1) {
      if (!SlabOwnership(ba, dm, np, &slabs_)) Abort("FLAT storage in a distributed run needs one z-slab per rank (SlabOwnership)");
      for (int r = 0; r < np; ++r) slab_owner.push_back(r);
    } else {
      slabs_.assign(1, mb);
    }
    vb.resize(slabs_.size());
    for (size_t s = 0; s < slabs_.size(); ++s)
      for (int d = 0; d < 3; ++d) { vb[s].lo[d] = slabs_[s].smallEnd(d); vb[s].hi[d] = slabs_[s].bigEnd(d); }
    sg = 0;
  } else {
    vb.resize(ba.size());
    for (long i = 0; i < ba.size(); ++i)
      for (int d = 0; d < 3; ++d) { vb[i].lo[d] = ba[i].smallEnd(d); vb[i].hi[d] = ba[i].bigEnd(d); }
  }
  st_ = std::make_shared<Storage>();
  // distributed run: box i lives on rank dm[i] (BOXES storage; FLAT is single-rank only)
  const bool dist = DistributionMapping::NProcs() > 1 && lay == Layout::BOXES && dm.size() == ba.size();
  const int* owners = !slab_owner.empty() ? slab_owner.data() : dist ? dm.ProcessorMap().data() : nullptr;
  lbx_check(lbx_mf_create_dist(vb.data(), (int)vb.size(), ncomp, sg, DTYPE, owners, &st_->mf), "MultiFab::define");
  size_t bytes = 0;
  lbx_check(lbx_mf_info(st_->mf, nullptr, nullptr, nullptr, nullptr, &bytes), "MultiFab::define");
  st_->elems = bytes / sizeof(T);
  st_->offset.resize(vb.size());
  for (size_t s = 0; s < vb.size(); ++s) {
    size_t off = 0;
    lbx_check(lbx_mf_fab(st_->mf, (int)s, nullptr, nullptr, &off), "MultiFab::define");
    st_->offset[s] = off / sizeof(T);
  }
  mirror_ok_ = false;
}

template <class T, int DTYPE>
void DeviceFabArray<T, DTYPE>::clear() {
  st_.reset();
  ba_.clear();
  dm_ = DistributionMapping();
  ncomp_ = ngrow_ = 0;
  slabs_.clear();
  mirror_.clear();
  mirror_ok_ = false;
}

template <class T, int DTYPE>
lbx_fab DeviceFabArray<T, DTYPE>::fabDesc(int s) const {
  lbx_fab f;
  lbx_check(lbx_mf_fab(st_->mf, s, &f, nullptr, nullptr), "MultiFab::fabDesc");
  return f;
}

template <class T, int DTYPE>
void DeviceFabArray<T, DTYPE>::setVal(T v) {
  if (!st_) return;
  lbx_check(lbx_mf_setval(st_->mf, (double)v), "MultiFab::setVal");
  touch();
}

template <class T, int DTYPE>
const std::vector<T>& DeviceFabArray<T, DTYPE>::hostMirror() const {
  if (!mirror_ok_) {
    mirror_.resize(st_ ? st_->elems : 0);
    if (st_) {
      lbx_check(lbx_mf_download(st_->mf, mirror_.data(), mirror_.size() * sizeof(T)), "MultiFab::hostMirror");
      lbx_check(lbx_sync(), "MultiFab::hostMirror");
    }
    mirror_ok_ = true;
  }
  return mirror_;
}

template <class T, int DTYPE>
void DeviceFabArray<T, DTYPE>::upload(const std::vector<T>& host) {
  if (!st_ || host.size() != st_->elems) Abort("MultiFab::upload: size mismatch");
  lbx_check(lbx_mf_upload(st_->mf, host.data(), host.size() * sizeof(T)), "MultiFab::upload");
  lbx_check(lbx_sync(), "MultiFab::upload");
  touch();
}

template <class T, int DTYPE>
T DeviceFabArray<T, DTYPE>::hostValue(int bi, const IntVect& p, int comp) const {
  const std::vector<T>& m = hostMirror();
  int s = bi;
  if (isFlat()) {                  // the slab that holds p
    s = 0;
    while (s + 1 < (int)slabs_.size() && !slabs_[s].contains(p)) ++s;
  }
  if (!storageBox(s).contains(p)) Abort("MultiFab::hostValue: cell outside the fab");
  // the ALLOCATED fab may be wider in x than valid + ghosts (sector alignment, lbx_mf_fab)
  lbx_fab fd;
  lbx_check(lbx_mf_fab(st_->mf, s, &fd, nullptr, nullptr), "MultiFab::hostValue");
  const size_t nx = fd.n[0], ny = fd.n[1], nz = fd.n[2];
  const size_t c = (size_t)(p[0] - fd.lo[0]) + nx * ((size_t)(p[1] - fd.lo[1]) + ny * (size_t)(p[2] - fd.lo[2]));
  return m[st_->offset[s] + (size_t)comp * nx * ny * nz + c];
}

template <>
void DeviceFabArray<double, LBX_F64>::relayout(Layout lay) {
  if (lay == lay_ || !st_) { lay_ = st_ ? lay_ : lay; return; }
  MultiFab fresh(ba_, dm_, ncomp_, ngrow_, lay);
  CopyValid(fresh, *this);
  st_ = fresh.st_;
  slabs_ = fresh.slabs_;
  lay_ = lay;
  touch();
}
template <>
void DeviceFabArray<int, LBX_I32>::relayout(Layout lay) {
  if (lay != lay_) Abort("iMultiFab::relayout is not supported");
}

template class DeviceFabArray<double, LBX_F64>;
template class DeviceFabArray<int, LBX_I32>;

// ----------------------------------------------------------------------------- tags
void TagBox::fill(char v, const Box& r, bool only_clear) {
  for (int k = r.smallEnd(2); k <= r.bigEnd(2); ++k)
    for (int j = r.smallEnd(1); j <= r.bigEnd(1); ++j) {
      char* row = &d_[index(IntVect(r.smallEnd(0), j, k))];
      const int len = r.length(0);
      if (!only_clear) { std::memset(row, v, (size_t)len); continue; }
      // inside a solid tagged region the rows hold no CLEAR cell at all: libc's vectorised scan finds that out an
      // order of magnitude faster than the byte loop
      char* z = static_cast<char*>(std::memchr(row, CLEAR, (size_t)len));
      if (!z) continue;
      for (int x = (int)(z - row); x < len; ++x) row[x] = row[x] == CLEAR ? v : row[x];
    }
}

void TagBox::materialise() {
  if (!d_.empty()) return;
  d_.assign((size_t)box_.numPts(), (char)CLEAR);
  const std::vector<Op> ops = std::move(ops_);
  ops_.clear();
  for (const Op& o : ops) fill(o.value, o.region, o.only_clear);
}

char TagBox::operator()(const IntVect& p) const {
  if (!d_.empty()) return d_[index(p)];
  char v = CLEAR;                              // box mode: replay the operations on this one cell
  for (const Op& o : ops_)
    if (o.region.contains(p) && !(o.only_clear && v != CLEAR)) v = o.value;
  return v;
}

void TagBox::setVal(char v, const Box& region) {
  const Box r = region & box_;
  if (!r.ok() || !allocated()) return;
  if (d_.empty()) {
    if (v == CLEAR && ops_.empty()) return;          // nothing tagged yet: already clear
    if (v != CLEAR && ops_.size() < MAX_OPS) { ops_.push_back({r, v, false}); return; }     // stay in box mode
    materialise();
  }
  fill(v, r, false);
}

// Dilation of the SET cells inside `interior` by +-nbuf in every direction (a cube, so it is done
// run by run: a row's run [i0, i1] of SET cells marks [i0 - nbuf, i1 + nbuf] in the (2 nbuf + 1)^2
// neighbouring rows), CLEAR cells becoming BUF.  Same result as growing every SET cell on its own.
// first index >= x in row[0..n) whose byte is not 0 (CLEAR), skipping eight cells at a time
static inline int next_nonclear(const char* row, int x, int n) {
  while (x < n && (reinterpret_cast<uintptr_t>(row + x) & 7) != 0) { if (row[x]) return x; ++x; }
  while (x + 8 <= n && *reinterpret_cast<const uint64_t*>(row + x) == 0) x += 8;
  while (x < n && !row[x]) ++x;
  return x;
}

void TagBox::buffer(int nbuf, const Box& interior) {
  if (nbuf <= 0 || !allocated() || !hasStorage()) return;
  const Box in = interior & box_;
  if (!in.ok()) return;
  if (boxMode()) {
    // every operation so far wrote SET (or BUF) over a box: the dilation of a SET box's part inside `interior` is that
    // part grown by nbuf, written into CLEAR cells only.  (BUF regions of an earlier buffer() do not spread.)
    bool simple = true;
    for (const Op& o : ops_) simple = simple && (o.value == SET || o.only_clear);
    if (simple && ops_.size() * 2 <= MAX_OPS) {
      const size_t n = ops_.size();
      for (size_t q = 0; q < n; ++q) {
        if (ops_[q].value != SET) continue;
        const Box core = ops_[q].region & in;
        if (!core.ok()) continue;
        const Box grown = amrex::grow(core, nbuf) & box_;
        if (grown != core) ops_.push_back({grown, (char)BUF, true});
      }
      return;
    }
    materialise();
  }
  struct Run { int i0, i1, j, k; };
  std::vector<Run> runs;
  for (int k = in.smallEnd(2); k <= in.bigEnd(2); ++k)
    for (int j = in.smallEnd(1); j <= in.bigEnd(1); ++j) {
      const char* row = &d_[index(IntVect(in.smallEnd(0), j, k))];
      const int n = in.length(0);
      for (int x = next_nonclear(row, 0, n); x < n;) {
        if (row[x] != SET) { x = next_nonclear(row, x + 1, n); continue; }
        int y = x;
        while (y + 1 < n && row[y + 1] == SET) ++y;
        runs.push_back({in.smallEnd(0) + x, in.smallEnd(0) + y, j, k});
        x = next_nonclear(row, y + 1, n);
      }
    }
  for (const Run& r : runs) {
    const int i0 = std::max(r.i0 - nbuf, box_.smallEnd(0)), i1 = std::min(r.i1 + nbuf, box_.bigEnd(0));
    for (int k = std::max(r.k - nbuf, box_.smallEnd(2)); k <= std::min(r.k + nbuf, box_.bigEnd(2)); ++k)
      for (int j = std::max(r.j - nbuf, box_.smallEnd(1)); j <= std::min(r.j + nbuf, box_.bigEnd(1)); ++j) {
        char* row = &d_[index(IntVect(i0, j, k))];
        const int len = i1 - i0 + 1;
        char* z = static_cast<char*>(std::memchr(row, CLEAR, (size_t)len));
        if (!z) continue;
        for (int x = (int)(z - row); x < len; ++x) row[x] = row[x] == CLEAR ? (char)BUF : row[x];
      }
  }
}

TagBoxArray::TagBoxArray(const BoxArray& ba, const DistributionMapping& dm, int ngrow) {
  ba_ = ba;
  dm_ = dm;
  ncomp_ = 1;
  ngrow_ = ngrow;
  fabs_.reserve(ba.size());
  // distributed run: a rank tags (and stores tags for) its own boxes only; collate() merges
  const bool dist = DistributionMapping::NProcs() > 1 && dm.size() == ba.size();
  for (long i = 0; i < ba.size(); ++i)
    fabs_.emplace_back(amrex::grow(ba[i], ngrow), !dist || dm[i] == DistributionMapping::MyProc());
}
void TagBoxArray::setVal(const BoxArray& ba, TagBox::TagVal v) { setVal(ba.boxList(), v); }
void TagBoxArray::setVal(const BoxList& bl, TagBox::TagVal v) {
  for (TagBox& t : fabs_)
    for (const Box& b : bl) t.setVal((char)v, b);
}
void TagBoxArray::buffer(int nbuf) {
  for (long i = 0; i < size(); ++i) fabs_[i].buffer(nbuf, ba_[i]);
}
// Every non-CLEAR cell mapped into `domain` through the periodic directions, minus the cells of
// `remove`, as a sorted (z slowest), duplicate-free list.  Built on a bitmap over the bounding box
// of the tagged rows: runs of tagged cells set bit ranges, `remove` boxes clear them, one scan emits
// the points in order -- no per-cell sort, no per-cell box tests.
void TagBoxArray::collate(std::vector<TagRun>& out, const Box& domain, const std::array<int, 3>& is_per,
                          const BoxList* remove) const {
  out.clear();
  using Run = TagRun;
  std::vector<Run> runs;
  auto wrap1 = [&](int v, int d, bool& ok) {
    const int len = domain.length(d), lo = domain.smallEnd(d);
    if (v >= lo && v <= domain.bigEnd(d)) return v;
    if (!is_per[d]) { ok = false; return v; }
    return lo + (((v - lo) % len) + len) % len;
  };
  // the run [i0, i1] of row (j, k) in index space, cut at the domain faces and wrapped piece by piece
  auto emit = [&](int i0, int i1, int jj, int kk) {
    int a = i0;
    while (a <= i1) {
      const int lo = domain.smallEnd(0), len = domain.length(0);
      const int cell0 = lo + (((a - lo) % len) + len) % len;         // image of a
      const int room = domain.bigEnd(0) - cell0;                     // cells up to the face
      const int piece = std::min(i1 - a, room);
      const bool inside = a >= lo && a <= domain.bigEnd(0);
      if (inside || is_per[0]) runs.push_back({cell0, cell0 + piece, jj, kk});
      a += piece + 1;
    }
  };
  // tagged regions of box-mode TagBoxes travel as BOXES (24 bytes for a solid 32^3 patch instead of a thousand runs):
  // cut at the domain faces and wrapped piece by piece in every periodic direction
  std::vector<Box> boxes;
  auto wrap_box = [&](const Box& r) {
    std::vector<Box> cur(1, r), nxt;
    for (int d = 0; d < 3; ++d) {
      const int lo = domain.smallEnd(d), hi = domain.bigEnd(d), len = domain.length(d);
      nxt.clear();
      for (const Box& b : cur) {
        int a = b.smallEnd(d);
        while (a <= b.bigEnd(d)) {
          const int cell0 = lo + (((a - lo) % len) + len) % len;
          const int piece = std::min(b.bigEnd(d) - a, hi - cell0);
          const bool inside = a >= lo && a <= hi;
          if (inside || is_per[d]) {
            Box c = b;
            c.setSmall(d, cell0);
            c.setBig(d, cell0 + piece);
            nxt.push_back(c);
          }
          a += piece + 1;
        }
      }
      cur.swap(nxt);
    }
    boxes.insert(boxes.end(), cur.begin(), cur.end());
  };
  for (const TagBox& t : fabs_) {
    if (!t.allocated() || !t.hasStorage()) continue;
    if (t.boxMode()) {
      // every operation of a box-mode TagBox wrote a non-CLEAR value: its regions ARE the tagged cells (overlaps
      // are merged by the bitmap below)
      for (const TagBox::Op& o : t.ops()) wrap_box(o.region);
      continue;
    }
    const Box& b = t.box();
    const int n = b.length(0);
    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
      for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j) {
        bool ok = true;
        const int jj = wrap1(j, 1, ok), kk = wrap1(k, 2, ok);
        if (!ok) continue;
        const char* row = t.row(IntVect(b.smallEnd(0), j, k));
        for (int x = next_nonclear(row, 0, n); x < n;) {
          int y = x;
          while (y + 1 < n && row[y + 1] != TagBox::CLEAR) ++y;
          emit(b.smallEnd(0) + x, b.smallEnd(0) + y, jj, kk);
          x = next_nonclear(row, y + 1, n);
        }
      }
  }
  if (DistributionMapping::NProcs() > 1) {
    // every rank contributes the runs and boxes of its own TagBoxes: sizes first, then the padded lists
    const int np = DistributionMapping::NProcs();
    long long mine[2] = {(long long)runs.size(), (long long)boxes.size()};
    std::vector<long long> counts(2 * (size_t)np);
    lbx_check(lbx_par_allgather(mine, sizeof(mine), counts.data()), "TagBoxArray::collate");
    long long most_r = 0, most_b = 0;
    for (int r = 0; r < np; ++r) { most_r = std::max(most_r, counts[2 * r]); most_b = std::max(most_b, counts[2 * r + 1]); }
    if (most_r == 0 && most_b == 0) return;
    if (most_r > 0) {
      std::vector<Run> send((size_t)most_r, Run{0, -1, 0, 0}), all((size_t)most_r * np);
      std::copy(runs.begin(), runs.end(), send.begin());
      lbx_check(lbx_par_allgather(send.data(), sizeof(Run) * (size_t)most_r, all.data()), "TagBoxArray::collate");
      runs.clear();
      for (int r = 0; r < np; ++r) runs.insert(runs.end(), all.begin() + (size_t)r * most_r, all.begin() + (size_t)r * most_r + counts[2 * r]);
    }
    if (most_b > 0) {
      struct B6 { int v[6]; };
      std::vector<B6> send((size_t)most_b, B6{{0, 0, 0, -1, -1, -1}}), all((size_t)most_b * np);
      for (size_t q = 0; q < boxes.size(); ++q)
        for (int d = 0; d < 3; ++d) { send[q].v[d] = boxes[q].smallEnd(d); send[q].v[3 + d] = boxes[q].bigEnd(d); }
      lbx_check(lbx_par_allgather(send.data(), sizeof(B6) * (size_t)most_b, all.data()), "TagBoxArray::collate");
      boxes.clear();
      for (int r = 0; r < np; ++r)
        for (long long q = 0; q < counts[2 * r + 1]; ++q) {
          const B6& e = all[(size_t)r * most_b + q];
          boxes.push_back(Box(IntVect(e.v[0], e.v[1], e.v[2]), IntVect(e.v[3], e.v[4], e.v[5])));
        }
    }
  }
  if (runs.empty() && boxes.empty()) return;
  IntVect blo, bhi;
  bool have = false;
  auto widen = [&](const IntVect& lo, const IntVect& hi) {
    if (!have) { blo = lo; bhi = hi; have = true; return; }
    for (int d = 0; d < 3; ++d) { blo[d] = std::min(blo[d], lo[d]); bhi[d] = std::max(bhi[d], hi[d]); }
  };
  for (const Run& r : runs) widen(IntVect(r.i0, r.j, r.k), IntVect(r.i1, r.j, r.k));
  for (const Box& b : boxes) widen(b.smallEnd(), b.bigEnd());
  const Box bb(blo, bhi);
  const size_t wpr = ((size_t)bb.length(0) + 63) / 64, ny = (size_t)bb.length(1), nz = (size_t)bb.length(2);
  std::vector<uint64_t> bits(wpr * ny * nz, 0);
  auto range = [&](int i0, int i1, int j, int k, bool set) {          // [i0, i1] relative to blo[0]
    uint64_t* row = &bits[wpr * ((size_t)(j - blo[1]) + ny * (size_t)(k - blo[2]))];
    const size_t w0 = (size_t)i0 / 64, w1 = (size_t)i1 / 64;
    for (size_t w = w0; w <= w1; ++w) {
      uint64_t m = ~uint64_t(0);
      if (w == w0) m &= ~uint64_t(0) << (i0 % 64);
      if (w == w1) m &= ~uint64_t(0) >> (63 - (i1 % 64));
      if (set) row[w] |= m; else row[w] &= ~m;
    }
  };
  for (const Run& r : runs) range(r.i0 - blo[0], r.i1 - blo[0], r.j, r.k, true);
  for (const Box& b : boxes)
    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
      for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j) range(b.smallEnd(0) - blo[0], b.bigEnd(0) - blo[0], j, k, true);
  if (remove)
    for (const Box& q : *remove) {
      const Box r = q & bb;
      if (!r.ok()) continue;
      for (int k = r.smallEnd(2); k <= r.bigEnd(2); ++k)
        for (int j = r.smallEnd(1); j <= r.bigEnd(1); ++j) range(r.smallEnd(0) - blo[0], r.bigEnd(0) - blo[0], j, k, false);
    }
  for (size_t k = 0; k < nz; ++k)
    for (size_t j = 0; j < ny; ++j) {
      const uint64_t* row = &bits[wpr * (j + ny * k)];
      const int nbits = bb.length(0);
      int x = 0;
      while (x < nbits) {
        const uint64_t rest = row[x / 64] >> (x % 64);
        if (!rest) { x = (x / 64 + 1) * 64; continue; }
        x += __builtin_ctzll(rest);                               // first set bit at or after x
        int y = x;                                                // extend over the run of set bits
        for (;;) {
          const uint64_t inv = ~(row[y / 64] >> (y % 64));        // zero bits from y on (shifted-in bits count as zero)
          const int room = 64 - (y % 64);
          const int ones = inv ? __builtin_ctzll(inv) : 64;
          if (ones < room) { y += ones; break; }
          y += room;
          if (y >= nbits) break;
        }
        if (y > nbits) y = nbits;
        out.push_back({blo[0] + x, blo[0] + y - 1, blo[1] + (int)j, blo[2] + (int)k});
        x = y;
      }
    }
}

// ----------------------------------------------------------------------------- gather plans
namespace {

struct PlanDeleter { void operator()(lbx_plan* p) const { lbx_plan_destroy(p); } };
// Plan cache.  A plan is identified by the geometry of the MultiFabs it moves data between, so every regrid
// that changes a BoxArray creates new keys and strands the old ones (64 B per descriptor on the device:
// tens of MB per plan at 10^4 boxes).  Two bounds keep a long dynamic-AMR run from growing: (1) a
// generation sweep -- AmrCore::regrid calls PlanCacheNewGeneration() when the grids changed, and plans that
// were not used since the regrid before that are destroyed; (2) an LRU cap on the number of plans.
struct CachedPlan {
  std::unique_ptr<lbx_plan, PlanDeleter> plan;
  uint64_t stamp = 0;        // last use (monotonic counter)
  uint64_t generation = 0;   // regrid generation of the last use
};
std::map<std::string, CachedPlan> g_plans;
uint64_t g_plan_clock = 0, g_plan_generation = 0;
constexpr size_t PLAN_CACHE_CAP = 64;
std::map<std::string, MultiFab> g_coarsened;     // sum_fine_to_coarse temporaries, one per fine BoxArray

uint64_t fnv(uint64_t h, int64_t v) {
  for (int b = 0; b < 8; ++b) { h ^= (uint64_t)((v >> (8 * b)) & 0xff); h *= 1099511628211ull; }
  return h;
}
// identity of a MultiFab's storage geometry
std::string geom_key(const FabArrayBase& fa, bool flat, int storage_ng) {
  uint64_t h = 1469598103934665603ull;
  for (long i = 0; i < fa.size(); ++i)
    for (int d = 0; d < 3; ++d) { h = fnv(h, fa.box((int)i).smallEnd(d)); h = fnv(h, fa.box((int)i).bigEnd(d)); }
  // ... and of who owns the boxes: plans are built per rank from the boxes it owns, and the same BoxArray can be
  // distributed in two ways in one process (one slab per rank for a uniform run, interleaved runs for a hierarchy)
  if (DistributionMapping::NProcs() > 1 && fa.DistributionMap().size() == fa.size())
    for (long i = 0; i < fa.size(); ++i) h = fnv(h, fa.DistributionMap()[i]);
  char buf[96];
  std::snprintf(buf, sizeof(buf), "%016llx:%ld:%d:%d", (unsigned long long)h, fa.size(), flat ? 1 : 0, storage_ng);
  return buf;
}
template <class FA>
std::string gkey(const FA& m) { return geom_key(m, m.isFlat(), m.isFlat() ? 0 : m.nGrow()); }
std::string pkey(const Periodicity& p) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "p%d,%d,%d", p.period()[0], p.period()[1], p.period()[2]);
  return buf;
}

// uniform-bin spatial index over a list of boxes
class BoxHash {
 public:
  explicit BoxHash(const std::vector<Box>& boxes) : boxes_(boxes) {
    bin_ = IntVect(1);
    for (size_t i = 0; i < boxes.size(); ++i) bound_ = i ? bound_.minBox(boxes[i]) : boxes[i];
    for (const Box& b : boxes)
      for (int d = 0; d < 3; ++d) bin_[d] = std::max(bin_[d], b.length(d));
    for (size_t i = 0; i < boxes.size(); ++i) {
      const IntVect c = cell(boxes[i].smallEnd());
      map_[key(c)].push_back((int)i);
    }
  }
  // bounding box of all boxes: a query outside it is empty -- what lets the plan builders skip 26 of the 27
  // periodic shifts for every region that is not near a domain face
  bool mayIntersect(const Box& q) const { return !boxes_.empty() && q.ok() && bound_.intersects(q); }
  // indices (ascending) of boxes intersecting q
  std::vector<int> query(const Box& q) const {
    std::vector<int> r;
    if (!mayIntersect(q)) return r;
    IntVect lo = cell(q.smallEnd()), hi = cell(q.bigEnd());
    for (int d = 0; d < 3; ++d) lo[d] -= 1;      // a box registered by its low corner may start one bin earlier
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          auto it = map_.find(key(IntVect(i, j, k)));
          if (it == map_.end()) continue;
          for (int b : it->second)
            if (boxes_[b].intersects(q)) r.push_back(b);
        }
    std::sort(r.begin(), r.end());
    return r;
  }

 private:
  IntVect cell(const IntVect& p) const {
    return IntVect(IntVect::floor_div(p[0], bin_[0]), IntVect::floor_div(p[1], bin_[1]), IntVect::floor_div(p[2], bin_[2]));
  }
  static uint64_t key(const IntVect& c) {
    return ((uint64_t)(uint32_t)(c[0] + (1 << 20)) << 42) ^ ((uint64_t)(uint32_t)(c[1] + (1 << 20)) << 21) ^
           (uint64_t)(uint32_t)(c[2] + (1 << 20));
  }
  const std::vector<Box>& boxes_;
  Box bound_;
  IntVect bin_;
  std::unordered_map<uint64_t, std::vector<int>> map_;
};

int g_group = 0;    // tile group stamped on the descriptors being built (see lbx_gather.group)

lbx_gather make_desc(int dst_fab, int src_set, int src_fab, int kind, int ratio, const IntVect& shift, const Box& r,
                     double value = 0.0) {
  lbx_gather g;
  g.dst_fab = dst_fab; g.group = g_group; g.src_set = src_set; g.src_fab = src_fab; g.kind = kind; g.ratio = ratio;
  for (int d = 0; d < 3; ++d) { g.shift[d] = shift[d]; g.region.lo[d] = r.smallEnd(d); g.region.hi[d] = r.bigEnd(d); }
  g.value = value;
  return g;
}

template <class FA>
std::vector<Box> storage_valid(const FA& m) {
  std::vector<Box> v(m.numStorageFabs());
  for (int s = 0; s < m.numStorageFabs(); ++s) v[s] = m.storageValid(s);
  return v;
}

// what to tile for one destination fab: the whole wanted box, or (ghost shell only) its up to
// six slabs around the valid box, each its own tile group
BoxList tile_regions(const Box& want, const Box& valid, bool shell_only) {
  BoxList r;
  if (shell_only) boxDiff(r, want, valid);
  else r.push_back(want);
  return r;
}

lbx_plan* cached(const std::string& key0, const std::function<void(std::vector<lbx_gather>&)>& build) {
  // plans describe this rank's own destination boxes only: the ownership view is part of the identity
  const std::string key = key0 + "|" + std::to_string(DistributionMapping::NProcs()) + "." + std::to_string(DistributionMapping::MyProc());
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    it->second.stamp = ++g_plan_clock;
    it->second.generation = g_plan_generation;
    return it->second.plan.get();
  }
  std::vector<lbx_gather> descs;
  const auto T0 = std::chrono::steady_clock::now();
  build(descs);
  const auto T1 = std::chrono::steady_clock::now();
  lbx_plan* p = nullptr;
  lbx_check(lbx_plan_create(descs.data(), (int)descs.size(), &p), "gather plan");
  if (getenv("LBX_HOST_TIMING"))
    std::cerr << "  [plan " << key.substr(0, 3) << "] " << descs.size() << " descriptors: intersections "
              << std::chrono::duration<double>(T1 - T0).count() << " s, upload "
              << std::chrono::duration<double>(std::chrono::steady_clock::now() - T1).count() << " s\n";
  CachedPlan& slot = g_plans[key];
  slot.plan.reset(p);
  slot.stamp = ++g_plan_clock;
  slot.generation = g_plan_generation;
  while (g_plans.size() > PLAN_CACHE_CAP) {       // least recently used first; never the one just built
    auto victim = g_plans.end();
    for (auto q = g_plans.begin(); q != g_plans.end(); ++q)
      if (q->second.plan.get() != p && (victim == g_plans.end() || q->second.stamp < victim->second.stamp)) victim = q;
    if (victim == g_plans.end()) break;
    g_plans.erase(victim);
  }
  return p;
}

// COPY descriptors: dst fab k <- every (source box, shift) that meets its region `want`
// COPY regions of one destination fab come from disjoint sources, so their order is free:
// smallest first, because the kernel walks the list backwards and most cells sit in the
// largest region.  (ADD keeps the ParallelCopy order: it fixes the summation order.)
void by_volume(std::vector<lbx_gather>& v, size_t from) {
  std::stable_sort(v.begin() + from, v.end(), [](const lbx_gather& a, const lbx_gather& b) {
    auto vol = [](const lbx_gather& g) {
      long n = 1;
      for (int d = 0; d < 3; ++d) n *= (g.region.hi[d] - g.region.lo[d] + 1);
      return n;
    };
    return vol(a) < vol(b);
  });
}

void copy_descs(std::vector<lbx_gather>& out, int k, const Box& want, const std::vector<Box>& svalid, const BoxHash& sh,
                int src_ng, const std::vector<IntVect>& shifts, int src_set, bool skip_self, int self_index,
                bool keep_order = false) {
  const size_t from = out.size();
  // collect (i, shift) pairs, then order by source index, then shift order (ParallelCopy order)
  std::vector<std::pair<int, int>> hits;
  for (size_t si = 0; si < shifts.size(); ++si) {
    const Box q = amrex::grow(amrex::shift(want, IntVect(0) - shifts[si]), src_ng);
    for (int i : sh.query(q)) {
      if (skip_self && i == self_index && shifts[si] == IntVect(0)) continue;
      hits.emplace_back(i, (int)si);
    }
  }
  std::sort(hits.begin(), hits.end());
  for (auto& h : hits) {
    const IntVect& s = shifts[h.second];
    const Box r = amrex::shift(amrex::grow(svalid[h.first], src_ng), s) & want;
    if (r.ok()) out.push_back(make_desc(k, src_set, h.first, LBX_G_COPY, 1, IntVect(0) - s, r));
  }
  if (!keep_order) by_volume(out, from);
}

// PC descriptors: fine fab k region `want` <- coarse boxes (valid cells, periodic images)
void pc_descs(std::vector<lbx_gather>& out, int k, const Box& want, const std::vector<Box>& cvalid, const BoxHash& ch,
              const std::vector<IntVect>& cshifts, int ratio, int src_set, const Box* exclude = nullptr) {
  const size_t from = out.size();
  const Box cw = amrex::coarsen(want, ratio);
  std::vector<std::pair<int, int>> hits;
  for (size_t si = 0; si < cshifts.size(); ++si)
    for (int i : ch.query(amrex::shift(cw, IntVect(0) - cshifts[si]))) hits.emplace_back(i, (int)si);
  std::sort(hits.begin(), hits.end());
  for (auto& h : hits) {
    const IntVect& s = cshifts[h.second];
    const Box rc = amrex::shift(cvalid[h.first], s) & cw;
    if (!rc.ok()) continue;
    const Box rf = amrex::refine(rc, ratio) & want;
    if (!rf.ok()) continue;
    BoxList parts;
    if (exclude) boxDiff(parts, rf, *exclude);
    else parts.push_back(rf);
    for (const Box& part : parts) out.push_back(make_desc(k, src_set, h.first, LBX_G_PC, ratio, IntVect(0) - s, part));
  }
  by_volume(out, from);
}

}  // namespace

void ClearPlanCache() { g_plans.clear(); g_coarsened.clear(); }
size_t PlanCacheSize() { return g_plans.size(); }
void PlanCacheNewGeneration() {
  // plans last used before the PREVIOUS regrid belong to grids that no longer exist (a plan of a level
  // the regrid left alone was used in between and survives)
  for (auto q = g_plans.begin(); q != g_plans.end();) {
    if (q->second.generation + 1 < g_plan_generation + 1 && q->second.generation < g_plan_generation) q = g_plans.erase(q);
    else ++q;
  }
  ++g_plan_generation;
}

namespace {
// ghosts-only plans open every ghost slab with a source-less descriptor spanning the slab, so
// that the launch tiles ALL its cells (the fused pass needs a thread per ghost cell)
void whole_region(std::vector<lbx_gather>& d, int k, const Box& reg) {
  d.push_back(make_desc(k, 0, 0, LBX_G_NONE, 1, IntVect(0), reg));
}
void run_plan(lbx_plan* p, MultiFab& dst, const lbx_mf* s0, const lbx_mf* s1, int op, const GhostPush* push,
              const char* what) {
  if (push && push->level_step)
    lbx_check(lbx_mf_collide_stream_level(push->src_valid->mf(), dst.mf(), push->omega_s, push->omega_b, p, s0, s1, push->wa,
                                          push->crse_b ? push->crse_b->mf() : nullptr, push->wb,
                                          push->fallback ? push->fallback->mf() : nullptr),
              what);
  else if (push)
    lbx_check(lbx_mf_collide_stream_fillpatch(push->src_valid->mf(), dst.mf(), push->omega_s, push->omega_b,
                                              push->mask ? push->mask->mf() : nullptr, push->fine_val,
                                              push->zero_invalid ? 1 : 0, p, s0, s1,
                                              push->fallback ? push->fallback->mf() : nullptr),
              what);
  else
    lbx_check(lbx_plan_apply(p, dst.mf(), s0, s1, op), what);
  dst.touch();
}
}  // namespace

void ParallelCopy(MultiFab& dst, const MultiFab& src, int src_ng, int dst_ng, const Periodicity& period, bool add,
                  bool ghosts_only, const GhostPush* push) {
  if (dst.empty() || src.empty()) return;
  if (push && !ghosts_only) Abort("ParallelCopy: GhostPush needs ghosts_only");
  if (ghosts_only && (add || dst.boxArray() != src.boxArray() || dst.layout() != src.layout()))
    Abort("ParallelCopy: ghosts_only needs identical source and destination boxes");
  if (src.isFlat()) src_ng = 0;
  if (dst.isFlat()) dst_ng = 0;
  if (src_ng > src.nGrow() || dst_ng > dst.nGrow()) Abort("ParallelCopy: ghost width exceeds the MultiFab's");
  char tail[64];
  std::snprintf(tail, sizeof(tail), "|%d|%d|%d|%d", src_ng, dst_ng, add ? 1 : 0, ghosts_only ? 1 : 0);
  const std::string key = "PC|" + gkey(dst) + "|" + gkey(src) + "|" + pkey(period) + tail;
  lbx_plan* p = cached(key, [&](std::vector<lbx_gather>& d) {
    const std::vector<Box> sv = storage_valid(src);
    const BoxHash sh(sv);
    const std::vector<IntVect> shifts = period.shiftIntVect();
    for (int k = 0; k < dst.numStorageFabs(); ++k) {
      if (!dst.isLocal(k)) continue;            // another rank fills its own boxes
      g_group = 0;
      for (const Box& reg : tile_regions(amrex::grow(dst.storageValid(k), dst_ng), dst.storageValid(k), ghosts_only)) {
        if (ghosts_only) whole_region(d, k, reg);
        copy_descs(d, k, reg, sv, sh, src_ng, shifts, 0, ghosts_only, k, add);
        ++g_group;
      }
    }
    g_group = 0;
  });
  run_plan(p, dst, src.mf(), nullptr, add ? LBX_OP_ADD : LBX_OP_COPY, push, "ParallelCopy");
}

void CopyValid(MultiFab& dst, const MultiFab& src) { ParallelCopy(dst, src, 0, 0, Periodicity::NonPeriodic()); }

void FillBoundary(MultiFab& mf, const Periodicity& period) {
  if (mf.empty() || mf.isFlat() || mf.nGrow() == 0) return;    // FLAT storage has no ghost cells
  const std::string key = "FB|" + gkey(mf) + "|" + pkey(period);
  lbx_plan* p = cached(key, [&](std::vector<lbx_gather>& d) {
    const std::vector<Box> sv = storage_valid(mf);
    const BoxHash sh(sv);
    const std::vector<IntVect> shifts = period.shiftIntVect();
    for (int k = 0; k < mf.numStorageFabs(); ++k) {
      if (!mf.isLocal(k)) continue;
      g_group = 0;
      for (const Box& reg : tile_regions(mf.storageBox(k), mf.storageValid(k), true)) {
        copy_descs(d, k, reg, sv, sh, 0, shifts, 0, true, k);
        ++g_group;
      }
    }
    g_group = 0;
  });
  lbx_check(lbx_plan_apply(p, mf.mf(), mf.mf(), nullptr, LBX_OP_COPY), "FillBoundary");
  mf.touch();
  // all directions periodic and the boxes tile the whole period: every ghost cell is now a copy
  // of the valid cell that covers it
  const IntVect& pp = period.period();
  if (pp[0] > 0 && pp[1] > 0 && pp[2] > 0 && mf.boxArray().numPts() == (long)pp[0] * pp[1] * pp[2]) mf.markGhostsFresh();
}

void FillPatchSingleLevel(MultiFab& dst, const MultiFab& src, const Geometry& geom, bool ghosts_only,
                          const GhostPush* push) {
  ParallelCopy(dst, src, 0, dst.nGrow(), geom.periodicity(), false, ghosts_only, push);
}

static void two_level_fill(MultiFab& dst, const MultiFab& crse, const MultiFab* fine, const Geometry& cgeom,
                           const Geometry& fgeom, const IntVect& ratio, const char* tag, bool ghosts_only = false,
                           const GhostPush* push = nullptr) {
  if (dst.empty()) return;
  if (push && !ghosts_only) Abort("FillPatchTwoLevels: GhostPush needs ghosts_only");
  if (ghosts_only && (!fine || dst.boxArray() != fine->boxArray() || fine->isFlat()))
    Abort("FillPatchTwoLevels: ghosts_only needs identical fine source and destination boxes");
  if (ratio[0] != ratio[1] || ratio[0] != ratio[2]) Abort("anisotropic refinement ratios are not supported");
  if (dst.isFlat()) Abort("two-level fills need BOXES storage on the fine level");
  const std::string key = std::string(tag) + "|" + gkey(dst) + "|" + gkey(crse) + "|" + (fine ? gkey(*fine) : "-") + "|" +
                          pkey(cgeom.periodicity()) + "|" + pkey(fgeom.periodicity()) + "|" + std::to_string(ratio[0]) +
                          (ghosts_only ? "|g" : "|a");
  lbx_plan* p = cached(key, [&](std::vector<lbx_gather>& d) {
    const std::vector<Box> cv = storage_valid(crse);
    const BoxHash ch(cv);
    const std::vector<IntVect> cs = cgeom.periodicity().shiftIntVect();
    std::vector<Box> fv;
    if (fine) fv = storage_valid(*fine);
    const BoxHash fh(fv);
    const std::vector<IntVect> fs = fgeom.periodicity().shiftIntVect();
    for (int k = 0; k < dst.numStorageFabs(); ++k) {
      if (!dst.isLocal(k)) continue;
      g_group = 0;
      for (const Box& reg : tile_regions(dst.storageBox(k), dst.storageValid(k), ghosts_only)) {
        if (ghosts_only) whole_region(d, k, reg);
        pc_descs(d, k, reg, cv, ch, cs, ratio[0], 1);                              // coarse first ...
        if (fine) copy_descs(d, k, reg, fv, fh, 0, fs, 0, ghosts_only, k);        // ... fine data wins
        ++g_group;
      }
    }
    g_group = 0;
  });
  run_plan(p, dst, fine ? fine->mf() : nullptr, crse.mf(), LBX_OP_COPY, push, tag);
}

void FillPatchTwoLevels(MultiFab& dst, const MultiFab& crse, const MultiFab& fine, const Geometry& cgeom,
                        const Geometry& fgeom, const IntVect& ratio, bool ghosts_only, const GhostPush* push) {
  two_level_fill(dst, crse, &fine, cgeom, fgeom, ratio, "FillPatchTwoLevels", ghosts_only, push);
}

void InterpFromCoarseLevel(MultiFab& dst, const MultiFab& crse, const Geometry& cgeom, const Geometry& fgeom,
                           const IntVect& ratio) {
  two_level_fill(dst, crse, nullptr, cgeom, fgeom, ratio, "InterpFromCoarseLevel");
}

// AMReX's own two steps: (1) average every fine box, ghosts included, onto its coarsened box with
// nGrow/ratio ghost cells (one search-free kernel), (2) ParallelCopy that temporary into the coarse
// valid cells with ADD and periodic wrap (src_ng = its ghosts, dst_ng = 0): a coarse cell under
// the ghost overlap of neighbouring fine boxes receives every contribution, in ParallelCopy order.
// Optional (fused, SetSumFineToCoarseFused): ONE gather -- the ParallelCopy's descriptors with kind AVG instead of
// COPY, applied with ADD straight from the fine level: a coarse cell forms the mean of the 8 fine cells of each
// contribution in registers (amrex_avgdown's summation order) and adds them in the same list order, so the result
// is bit-identical to the two steps while the coarsened temporary is never written or read.  MEASURED SLOWER
// (profiles/r02_launches_amr_2level_128_fused_sum.csv: 282 us against 73 + 101 us at 128^3; again after the warp-level
// descriptor filter of k_plan_apply: C4 at 256^3 runs 8.9-9.4 GLUPS fused against 11.0 in two steps): a coarse cell
// under the ghost overlap of several fine boxes pays eight times the loads behind EACH contribution, so the default
// stays with the two steps.  LBX_SUM_FUSED=1 selects the one-gather form for experiments.
namespace { bool g_sum_fused = [] { const char* e = std::getenv("LBX_SUM_FUSED"); return e && e[0] == '1'; }(); }
void SetSumFineToCoarseFused(bool on) { g_sum_fused = on; }

// the coarsened image of a fine level (its boxes coarsened, cng ghost cells, the fine level's ownership): the
// temporary both restrictions average into before anything crosses to the coarse level's owners
static MultiFab& coarsened_tmp(const MultiFab& fine, int r, int cng, const char* who) {
  const std::string key = gkey(fine) + "|" + std::to_string(r) + "|" + std::to_string(cng) + "|" +
                          std::to_string(DistributionMapping::NProcs()) + "." + std::to_string(DistributionMapping::MyProc());
  if (!g_coarsened.count(key) && g_coarsened.size() >= 4) g_coarsened.clear();   // grids change at regrid: keep a few
  MultiFab& tmp = g_coarsened[key];
  if (tmp.empty()) {
    BoxList bl;
    for (long i = 0; i < fine.size(); ++i) {
      const Box cb = amrex::coarsen(fine.box((int)i), r);
      if (amrex::refine(amrex::grow(cb, cng), r) != fine.fabbox((int)i)) Abort(std::string(who) + ": fine box not aligned to the coarse grid");
      bl.push_back(cb);
    }
    tmp.define(BoxArray(bl), fine.DistributionMap(), fine.nComp(), cng);
  }
  return tmp;
}

void sum_fine_to_coarse(const MultiFab& fine, MultiFab& crse, int scomp, int ncomp, const IntVect& ratio,
                        const Geometry& cgeom, const Geometry& /*fgeom*/) {
  if (fine.empty() || crse.empty()) return;
  if (scomp != 0 || ncomp != crse.nComp()) Abort("sum_fine_to_coarse: only all components are supported");
  if (fine.isFlat()) Abort("sum_fine_to_coarse: fine level must use BOXES storage");
  const int r = ratio[0];
  if (fine.nGrow() % r != 0) Abort("sum_fine_to_coarse: fine.nGrow() must be a multiple of the ratio");
  const int cng = fine.nGrow() / r;
  if (g_sum_fused && !crse.isFlat()) {
    const Periodicity period = cgeom.periodicity();
    const std::string key = "SF|" + gkey(crse) + "|" + gkey(fine) + "|" + pkey(period) + "|" + std::to_string(r);
    lbx_plan* p = cached(key, [&](std::vector<lbx_gather>& d) {
      std::vector<Box> cf(fine.size());                   // the coarsened fine boxes (valid part)
      for (long i = 0; i < fine.size(); ++i) {
        cf[i] = amrex::coarsen(fine.box((int)i), r);
        if (amrex::refine(amrex::grow(cf[i], cng), r) != fine.fabbox((int)i)) Abort("sum_fine_to_coarse: fine box not aligned to the coarse grid");
      }
      const BoxHash sh(cf);
      const std::vector<IntVect> shifts = period.shiftIntVect();
      for (int k = 0; k < crse.numStorageFabs(); ++k) {
        if (!crse.isLocal(k)) continue;
        g_group = 0;
        const size_t from = d.size();
        copy_descs(d, k, crse.storageValid(k), cf, sh, cng, shifts, 0, false, k, true);      // ParallelCopy order (ADD)
        for (size_t q = from; q < d.size(); ++q) {        // COPY from the temporary -> AVG from the fine box itself
          d[q].kind = LBX_G_AVG;
          d[q].ratio = r;
          for (int a = 0; a < 3; ++a) d[q].shift[a] *= r;
        }
      }
      g_group = 0;
    });
    lbx_check(lbx_plan_apply(p, crse.mf(), fine.mf(), nullptr, LBX_OP_ADD), "sum_fine_to_coarse");
    crse.touch();
    return;
  }
  MultiFab& tmp = coarsened_tmp(fine, r, cng, "sum_fine_to_coarse");
  lbx_check(lbx_mf_average_down(fine.mf(), tmp.mf(), r), "sum_fine_to_coarse");
  tmp.touch();
  ParallelCopy(crse, tmp, cng, 0, cgeom.periodicity(), true);
}

// One AVG plan: coarse box k, region coarsen(fine box i) n (valid box k) <- mean of the fine cells above
// it.  Fine valid boxes are disjoint, so every coarse cell has at most one source.
void average_down(const MultiFab& fine, MultiFab& crse, int scomp, int ncomp, const IntVect& ratio) {
  if (fine.empty() || crse.empty()) return;
  if (scomp != 0 || ncomp != crse.nComp() || ncomp != fine.nComp()) Abort("average_down: only all components are supported");
  if (fine.isFlat() || crse.isFlat()) Abort("average_down: both levels must use BOXES storage");
  if (ratio[0] != ratio[1] || ratio[0] != ratio[2]) Abort("anisotropic refinement ratios are not supported");
  const int r = ratio[0];
  if (ncomp == 15 && fine.nGrow() % r == 0) {
    // the populations: average on the fine level's owner first (one search-free launch at the read roofline, the
    // same arithmetic as the AVG descriptor: identical bits), then copy the coarsened boxes -- 1/8 of the bytes --
    // to the coarse level's owners.  The AVG plan below reads the FINE boxes from wherever they live: on 8 GPUs that
    // put the whole fine level on NVLink every coarse step (16 of 36 ms of the SUBCYCLE C5 step).
    const int cng = fine.nGrow() / r;
    MultiFab& tmp = coarsened_tmp(fine, r, cng, "average_down");
    lbx_check(lbx_mf_average_down(fine.mf(), tmp.mf(), r), "average_down");
    tmp.touch();
    ParallelCopy(crse, tmp, 0, 0, Periodicity::NonPeriodic());
    return;
  }
  const std::string key = "AD|" + gkey(crse) + "|" + gkey(fine) + "|" + std::to_string(r);
  lbx_plan* p = cached(key, [&](std::vector<lbx_gather>& d) {
    std::vector<Box> cf(fine.size());
    for (long i = 0; i < fine.size(); ++i) {
      cf[i] = amrex::coarsen(fine.box((int)i), r);
      if (amrex::refine(cf[i], r) != fine.box((int)i)) Abort("average_down: fine box not aligned to the coarse grid");
    }
    const BoxHash fh(cf);
    for (int k = 0; k < (int)crse.size(); ++k) {
      if (!crse.isLocal(k)) continue;
      const Box want = crse.box(k);
      const size_t from = d.size();
      for (int i : fh.query(want)) {
        const Box reg = cf[i] & want;
        if (reg.ok()) d.push_back(make_desc(k, 0, i, LBX_G_AVG, r, IntVect(0), reg));
      }
      by_volume(d, from);
    }
  });
  lbx_check(lbx_plan_apply(p, crse.mf(), fine.mf(), nullptr, LBX_OP_COPY), "average_down");
  crse.touch();
}

void LinComb(MultiFab& dst, double a, const MultiFab& x, double b, const MultiFab& y) {
  if (dst.empty()) return;
  if (dst.boxArray() != x.boxArray() || dst.boxArray() != y.boxArray() || dst.layout() != x.layout() ||
      dst.layout() != y.layout())
    Abort("LinComb: the three MultiFabs must share boxes and layout");
  lbx_check(lbx_mf_lincomb(dst.mf(), a, x.mf(), b, y.mf()), "LinComb");
  dst.touch();
}

iMultiFab makeFineMask(const MultiFab& cmf, const BoxArray& fba, const IntVect& ratio, int crse_value, int fine_value) {
  iMultiFab mask(cmf.boxArray(), cmf.DistributionMap(), 1, cmf.nGrow());
  if (mask.empty()) return mask;
  mask.setVal(crse_value);
  std::vector<Box> cf(fba.size());
  for (long i = 0; i < fba.size(); ++i) cf[i] = amrex::coarsen(fba[i], ratio);
  const BoxHash fh(cf);
  std::vector<lbx_gather> d;
  for (int k = 0; k < (int)mask.size(); ++k) {
    if (!mask.isLocal(k)) continue;
    const Box want = mask.fabbox(k);
    for (int i : fh.query(want)) d.push_back(make_desc(k, 0, 0, LBX_G_CONST, 1, IntVect(0), cf[i] & want, (double)fine_value));
  }
  lbx_plan* p = nullptr;
  lbx_check(lbx_plan_create(d.data(), (int)d.size(), &p), "makeFineMask");
  lbx_check(lbx_plan_apply(p, mask.mf(), nullptr, nullptr, LBX_OP_COPY), "makeFineMask");
  lbx_plan_destroy(p);
  mask.touch();
  return mask;
}

}  // namespace amrex
