// amrex-mini AmrMesh/AmrCore (see AMReX_AmrCore.H).  [AMReX, unverified] throughout: restated
// from upstream semantics recorded in SURVEY.md appendix C, not from AMReX source.
#include "AMReX_AmrCore.H"
#include <chrono>
#include <cstdlib>
#include <iostream>

#include <iostream>
#include <list>

#include "AMReX_FillPatch.H"

namespace amrex {

// ----------------------------------------------------------------------------- clustering
namespace {
enum CutStatus { HoleCut = 0, SteepCut = 1, BisectCut = 2, InvalidCut = 3 };

int find_cut(const std::vector<int>& hist, CutStatus& status) {
  const int len = (int)hist.size();
  status = InvalidCut;
  if (len <= 1) return 0;
  const int mid = len / 2;
  int cut = -1;
  for (int i = 0; i < len; ++i) {
    if (hist[i] == 0) {                        // centre-most empty plane
      status = HoleCut;
      if (std::abs(cut - mid) > std::abs(i - mid)) {
        cut = i;
        if (i > mid) break;
      }
    }
  }
  if (status == HoleCut) return cut;
  std::vector<int> dh(len, 0);                 // discrete Laplacian of the signature
  for (int i = 1; i < len - 1; ++i) dh[i] = hist[i + 1] - 2 * hist[i] + hist[i - 1];
  int locmax = -1;
  for (int i = 2; i < len - 2; ++i) {          // strongest sign change, at least 2 cells from the ends
    const int ip = dh[i - 1], ic = dh[i], dif = std::abs(ip - ic);
    if (ip * ic < 0 && dif >= locmax) {
      if (dif > locmax) { status = SteepCut; cut = i; locmax = dif; }
      else if (std::abs(i - mid) < std::abs(cut - mid)) cut = i;
    }
  }
  if (locmax <= 2) { status = BisectCut; return mid; }
  return cut;
}

Box min_box(const std::vector<TagRun>& r) {
  Box b(IntVect(r[0].i0, r[0].j, r[0].k), IntVect(r[0].i1, r[0].j, r[0].k));
  for (const TagRun& t : r) b.minBox(Box(IntVect(t.i0, t.j, t.k), IntVect(t.i1, t.j, t.k)));
  return b;
}
long long cells_of(const std::vector<TagRun>& r) {
  long long n = 0;
  for (const TagRun& t : r) n += t.i1 - t.i0 + 1;
  return n;
}
}  // namespace

// Berger-Rigoutsos: split the minimal box of the tags at a hole / an inflection of the signature /
// the middle until every box is filled to `eff` [AMReX ClusterList::chop, unverified].
BoxList ClusterRuns(std::vector<TagRun> all, Real eff) {
  struct Cl { std::vector<TagRun> r; long long count; Box box; };
  BoxList out;
  if (all.empty()) return out;
  std::list<Cl> lst;
  {
    Cl c;
    c.r = std::move(all);
    c.count = cells_of(c.r);
    c.box = min_box(c.r);
    lst.push_back(std::move(c));
  }
  for (auto it = lst.begin(); it != lst.end();) {
    Cl& c = *it;
    if ((Real)c.count / (Real)c.box.numPts() >= eff) { ++it; continue; }
    CutStatus st[3], mincut = InvalidCut;
    int cut[3];
    for (int d = 0; d < 3; ++d) {
      const int lo = c.box.smallEnd(d), len = c.box.length(d);
      std::vector<int> hist(len, 0);
      if (d == 0) {
        std::vector<int> diff(len + 1, 0);
        for (const TagRun& t : c.r) { ++diff[t.i0 - lo]; --diff[t.i1 + 1 - lo]; }
        int acc = 0;
        for (int i = 0; i < len; ++i) { acc += diff[i]; hist[i] = acc; }
      } else {
        for (const TagRun& t : c.r) hist[(d == 1 ? t.j : t.k) - lo] += t.i1 - t.i0 + 1;
      }
      cut[d] = lo + find_cut(hist, st[d]);
      if (st[d] < mincut) mincut = st[d];
    }
    int dir = -1;
    for (int d = 0; d < 3; ++d)
      if (st[d] == mincut && (dir < 0 || c.box.length(d) > c.box.length(dir))) dir = d;
    std::vector<TagRun> lo_r, hi_r;
    for (const TagRun& t : c.r) {
      if (dir == 0) {
        if (t.i1 < cut[0]) lo_r.push_back(t);
        else if (t.i0 >= cut[0]) hi_r.push_back(t);
        else { lo_r.push_back({t.i0, cut[0] - 1, t.j, t.k}); hi_r.push_back({cut[0], t.i1, t.j, t.k}); }
      } else {
        ((dir == 1 ? t.j : t.k) < cut[dir] ? lo_r : hi_r).push_back(t);
      }
    }
    if (lo_r.empty() || hi_r.empty()) { ++it; continue; }   // cannot split further
    Cl h;
    h.r = std::move(hi_r);
    h.count = cells_of(h.r);
    h.box = min_box(h.r);
    lst.push_back(std::move(h));
    c.r = std::move(lo_r);
    c.count = cells_of(c.r);
    c.box = min_box(c.r);
    // the low part is examined again before moving on
  }
  for (const Cl& c : lst) out.push_back(c.box);
  return out;
}

BoxList ClusterTags(std::vector<IntVect>& tags, Real eff) {
  std::vector<TagRun> runs;
  runs.reserve(tags.size());
  for (const IntVect& p : tags) {              // consecutive cells of a row merge; any order is fine
    if (!runs.empty() && runs.back().j == p[1] && runs.back().k == p[2] && runs.back().i1 + 1 == p[0]) ++runs.back().i1;
    else runs.push_back({p[0], p[0], p[1], p[2]});
  }
  return ClusterRuns(std::move(runs), eff);
}

// ----------------------------------------------------------------------------- AmrMesh
AmrMesh::AmrMesh(const Geometry& g0, const AmrInfo& info)
    : verbose(info.verbose), max_level(info.max_level), grid_eff(info.grid_eff), n_proper(info.n_proper),
      refine_grid_layout(info.refine_grid_layout) {
  const int nlev = max_level + 1;
  auto spread = [nlev](const Vector<IntVect>& v, const IntVect& dflt) {
    Vector<IntVect> r(nlev, dflt);
    for (int i = 0; i < nlev; ++i) r[i] = v.empty() ? dflt : v[std::min<size_t>(i, v.size() - 1)];
    return r;
  };
  ref_ratio = spread(info.ref_ratio, IntVect(2));
  blocking_factor = spread(info.blocking_factor, IntVect(8));
  max_grid_size = spread(info.max_grid_size, IntVect(32));
  n_error_buf = spread(info.n_error_buf, IntVect(1));
  geom.resize(nlev);
  grids.resize(nlev);
  dmap.resize(nlev);
  geom[0] = g0;
  for (int l = 1; l < nlev; ++l) geom[l] = amrex::refine(geom[l - 1], ref_ratio[l - 1]);
}

BoxArray AmrMesh::MakeBaseGrids() const {
  const Box& dom = geom[0].Domain();
  IntVect fac(2);
  for (int d = 0; d < 3; ++d)
    if (dom.length(d) % 2 != 0) fac[d] = 1;     // odd extents are not coarsened
  BoxArray ba(amrex::coarsen(dom, fac));
  IntVect chunk;
  for (int d = 0; d < 3; ++d) chunk[d] = std::max(1, max_grid_size[0][d] / fac[d]);
  ba.maxSize(chunk);
  ba.refine(fac);
  return ba;
}

namespace {
void proj_periodic(BoxList& bl, const Box& domain, const Geometry& g) {
  const BoxList orig(bl);
  for (const Box& b : orig)
    for (const IntVect& s : g.periodicity().shiftIntVect()) {
      if (s == IntVect(0)) continue;
      const Box r = amrex::shift(b, s) & domain;
      if (r.ok()) bl.push_back(r);
    }
}
}  // namespace

void AmrMesh::MakeNewGrids(int lbase, Real time, int& new_finest, Vector<BoxArray>& new_grids) {
  const int max_crse = std::min(finest_level, max_level - 1);
  if ((int)new_grids.size() < max_crse + 2) new_grids.resize(max_crse + 2);
  // proper-nesting domains (blocking factor 1 on this path: no tag coarsening)
  Vector<BoxList> p_n(max_level), p_n_comp(max_level);
  auto TP = std::chrono::steady_clock::now();
  auto plap = [&](const char* what) {
    if (!getenv("LBX_HOST_TIMING")) return;
    auto T1 = std::chrono::steady_clock::now();
    std::cerr << "  [regrid lbase " << lbase << "] proper nesting: " << what << " " << std::chrono::duration<double>(T1 - TP).count() << " s\n";
    TP = T1;
  };
  {
    BoxList bl = grids[lbase].boxList();
    simplify(bl);
    plap("simplify(level grids)");
    p_n_comp[lbase] = complementIn(geom[lbase].Domain(), bl);
    plap("complementIn");
    simplify(p_n_comp[lbase]);
    plap("simplify(complement)");
    for (Box& b : p_n_comp[lbase]) b.grow(n_proper);
    if (geom[lbase].isAnyPeriodic()) proj_periodic(p_n_comp[lbase], geom[lbase].Domain(), geom[lbase]);
    p_n[lbase] = complementIn(geom[lbase].Domain(), p_n_comp[lbase]);
    simplify(p_n[lbase]);
    plap("grow/periodic/complement/simplify");
  }
  for (int i = lbase + 1; i <= max_crse; ++i) {
    p_n_comp[i] = p_n_comp[i - 1];
    simplify(p_n_comp[i]);
    for (Box& b : p_n_comp[i]) { b.refine(ref_ratio[i - 1]); b.grow(n_proper); }
    if (geom[i].isAnyPeriodic()) proj_periodic(p_n_comp[i], geom[i].Domain(), geom[i]);
    p_n[i] = complementIn(geom[i].Domain(), p_n_comp[i]);
    simplify(p_n[i]);
    plap("finer level");
  }
  new_finest = lbase;
  for (int levc = max_crse; levc >= lbase; --levc) {
    const int levf = levc + 1;
    const int nbuf = n_error_buf[levc][0];
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!getenv("LBX_HOST_TIMING")) return;
      auto T1 = std::chrono::steady_clock::now();
      std::cerr << "  [regrid levc " << levc << "] " << what << " " << std::chrono::duration<double>(T1 - T0).count() << " s\n";
      T0 = T1;
    };
    // AMReX's order [AMReX, unverified; ADVICE r01]: the new grids two levels up are projected down to levc
    // FIRST, the tag boxes are allocated wide enough to hold that projection (ngrow = how far grids[levc]
    // must grow to contain it; 0 under proper nesting), the projection is SET before buffering -- so it is
    // buffered like the user's tags -- and the buffer width is n_error_buf + ngrow.
    BoxList proj;
    int ngrow = 0;
    if (levf < new_finest) {
      for (const Box& b : new_grids[levf + 1].boxList()) {
        Box c = amrex::coarsen(b, ref_ratio[levf]);
        c.grow(n_proper);
        c.coarsen(ref_ratio[levc]);
        proj.push_back(c);
      }
      simplify(proj);        // thousands of fine boxes project onto a handful: the tag boxes stay in box mode
      // containment against the SIMPLIFIED level grids (a handful of boxes): grow(union, g) = union of grown boxes
      BoxList simp = grids[levc].boxList();
      simplify(simp);
      auto covered = [&](int g) {
        BoxList grown = simp;
        for (Box& b : grown) b.grow(g);
        for (const Box& c : proj) {
          bool inside = false;
          for (const Box& q : grown)
            if (q.contains(c)) { inside = true; break; }
          if (!inside && !complementIn(c, grown).empty()) return false;
        }
        return true;
      };
      while (ngrow < 64 && !covered(ngrow)) ++ngrow;
    }
    TagBoxArray tags(grids[levc], dmap[levc], nbuf + ngrow);
    lap("alloc tags");
    ErrorEst(levc, tags, time, 0);
    lap("ErrorEst");
    if (!proj.empty()) tags.setVal(proj, TagBox::SET);
    tags.buffer(nbuf + ngrow);
    lap("buffer");
    std::vector<TagRun> tagvec;
    // cells outside the proper nesting domain are dropped while collating
    tags.collate(tagvec, geom[levc].Domain(), geom[levc].isPeriodicArray(), p_n_comp[levc].empty() ? nullptr : &p_n_comp[levc]);
    lap("collate");
    lap("proper nesting filter");
    if (tagvec.empty()) continue;
    new_finest = std::max(new_finest, levf);
    BoxList clusters = ClusterRuns(std::move(tagvec), grid_eff);
    lap("ClusterTags");
    BoxList clipped;                           // ClusterList::intersect(p_n)
    for (const Box& b : clusters) {
      bool whole = false;
      for (const Box& q : p_n[levc])
        if (q.contains(b)) { whole = true; break; }
      if (whole) { clipped.push_back(b); continue; }
      for (const Box& q : p_n[levc]) {
        const Box r = b & q;
        if (r.ok()) clipped.push_back(r);
      }
    }
    simplify(clipped);
    IntVect largest;
    for (int d = 0; d < 3; ++d) largest[d] = std::max(1, max_grid_size[levf][d] / ref_ratio[levc][d]);
    maxSize(clipped, largest);
    for (Box& b : clipped) b.refine(ref_ratio[levc]);
    new_grids[levf].define(clipped);
    lap("clip/simplify/maxSize");
  }
}

void AmrMesh::MakeNewGrids(Real time) {
  finest_level = 0;
  {
    const BoxArray ba = MakeBaseGrids();
    const DistributionMapping dm = MakeDistributionMap(ba);
    MakeNewLevelFromScratch(0, time, ba, dm);
    SetBoxArray(0, ba);
    SetDistributionMap(0, dm);
  }
  if (max_level > 0) {
    Vector<BoxArray> new_grids(max_level + 1);
    new_grids[0] = grids[0];
    do {
      int new_finest;
      MakeNewGrids(finest_level, time, new_finest, new_grids);
      if (new_finest <= finest_level) break;
      finest_level = new_finest;
      const DistributionMapping dm = MakeDistributionMap(new_grids[new_finest]);
      MakeNewLevelFromScratch(new_finest, time, new_grids[new_finest], dm);
      SetBoxArray(new_finest, new_grids[new_finest]);
      SetDistributionMap(new_finest, dm);
    } while (finest_level < max_level);
  }
}

// ----------------------------------------------------------------------------- AmrCore
void AmrCore::InitFromScratch(Real time) {
  MakeNewGrids(time);
  if (verbose > 0)
    for (int l = 0; l <= finest_level; ++l)
      std::cout << "INITIAL GRIDS: level " << l << " has " << grids[l].size() << " grids, " << grids[l].numPts()
                << " cells" << std::endl;
}

void AmrCore::regrid(int lbase, Real time, bool) {
  if (lbase >= max_level) return;
  int new_finest;
  Vector<BoxArray> new_grids(finest_level + 2);
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int lev) {
    if (!getenv("LBX_HOST_TIMING")) return;
    auto T1 = std::chrono::steady_clock::now();
    std::cerr << "  [regrid lbase " << lbase << "] " << what << " " << lev << ": " << std::chrono::duration<double>(T1 - T0).count() << " s\n";
    T0 = T1;
  };
  MakeNewGrids(lbase, time, new_finest, new_grids);
  lap("MakeNewGrids", lbase);
  bool coarse_ba_changed = false, grids_changed = new_finest != finest_level;
  for (int lev = lbase + 1; lev <= new_finest; ++lev) {
    if (lev <= finest_level) {                 // an existing level
      const bool ba_changed = (new_grids[lev] != grids[lev]);
      grids_changed = grids_changed || ba_changed;
      if (ba_changed || coarse_ba_changed) {
        BoxArray level_grids = grids[lev];
        DistributionMapping level_dmap = dmap[lev];
        if (ba_changed) {
          level_grids = new_grids[lev];
          level_dmap = MakeDistributionMap(level_grids);
        }
        RemakeLevel(lev, time, level_grids, level_dmap);
        lap("RemakeLevel", lev);
        SetBoxArray(lev, level_grids);
        SetDistributionMap(lev, level_dmap);
      }
      coarse_ba_changed = ba_changed;
    } else {                                   // a new level
      const DistributionMapping new_dmap = MakeDistributionMap(new_grids[lev]);
      MakeNewLevelFromCoarse(lev, time, new_grids[lev], new_dmap);
      lap("MakeNewLevelFromCoarse", lev);
      SetBoxArray(lev, new_grids[lev]);
      SetDistributionMap(lev, new_dmap);
    }
  }
  for (int lev = new_finest + 1; lev <= finest_level; ++lev) {
    ClearLevel(lev);
    ClearBoxArray(lev);
    ClearDistributionMap(lev);
  }
  finest_level = new_finest;
  if (grids_changed) PlanCacheNewGeneration();     // gather plans of grids that no longer exist are dropped
  if (verbose > 0)
    std::cout << "REGRID: finest level " << finest_level << std::endl;
}

}  // namespace amrex
