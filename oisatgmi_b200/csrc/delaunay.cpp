// Host-side 2-D Delaunay triangulation of the pixel centres of one granule
// (plan builder v1, SURVEY.md section 7 step 8 / hard part H1).
//
// The reference triangulates with Qhull through scipy.spatial.Delaunay
// (interpolator.py:153; 0.9 s per 98,640-pixel OMI granule, the ceiling of the
// whole end-to-end path once everything else runs on the GPU).  This is a
// from-scratch radial sweep-hull construction: points are inserted in order of
// distance from the circumcentre of a seed triangle, each insertion fans new
// triangles onto the visible part of the advancing convex hull and restores the
// Delaunay property by edge flips.  ~50 ms for the same granule.
//
// Exactness.  For points in general position the Delaunay triangulation is
// unique, so the triangle SET equals Qhull's.  Every geometric decision goes
// through orient2d / incircle predicates evaluated with a floating-point filter
// and, when the filter cannot certify the sign, exactly (multi-term floating-point
// expansions, after Shewchuk 1997).  Exact zeros that make the answer ambiguous --
// co-circular quadruples met by the flip test, collinear triples on the final
// hull -- are counted and reported: then the triangulation is NOT unique, Qhull's
// own tie-breaking cannot be reproduced, and the Python layer falls back to scipy
// for that granule (regular L3 lattices always do, and are cached).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <vector>

#include "../../include/oisat.h"
#include "flip_rounds.h"

namespace {

// ------------------------------------------------------------------ predicates
struct Expansion {
  // non-overlapping, increasing magnitude; small fixed capacity is enough for
  // the degree-4 incircle polynomial of 2-term differences
  static constexpr int kCap = 1600;
  double v[kCap];
  int n = 0;
};

inline void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
inline void two_prod(double a, double b, double& p, double& e) {
  p = a * b;
  e = std::fma(a, b, -p);
}

// e + b  (grow-expansion with zero elimination)
void grow(const Expansion& e, double b, Expansion& h) {
  double q = b;
  int m = 0;
  for (int i = 0; i < e.n; ++i) {
    double s, err;
    two_sum(q, e.v[i], s, err);
    if (err != 0.0) h.v[m++] = err;
    q = s;
  }
  if (q != 0.0 || m == 0) h.v[m++] = q;
  h.n = m;
}

void add(const Expansion& a, const Expansion& b, Expansion& out) {
  static thread_local Expansion t0, t1;
  const Expansion* cur = &a;
  Expansion* nxt = &t0;
  for (int j = 0; j < b.n; ++j) {
    grow(*cur, b.v[j], *nxt);
    cur = nxt;
    nxt = (nxt == &t0) ? &t1 : &t0;
  }
  out.n = cur->n;
  for (int i = 0; i < cur->n; ++i) out.v[i] = cur->v[i];
}

// e * b  (scale-expansion with zero elimination)
void scale(const Expansion& e, double b, Expansion& h) {
  int m = 0;
  if (e.n == 0) { h.n = 0; return; }
  double q, err;
  two_prod(e.v[0], b, q, err);
  if (err != 0.0) h.v[m++] = err;
  for (int i = 1; i < e.n; ++i) {
    double p1, p0;
    two_prod(e.v[i], b, p1, p0);
    double s, r;
    two_sum(q, p0, s, r);
    if (r != 0.0) h.v[m++] = r;
    // fast two-sum: |p1| >= |s|
    q = p1 + s;
    const double t = s - (q - p1);
    if (t != 0.0) h.v[m++] = t;
  }
  if (q != 0.0 || m == 0) h.v[m++] = q;
  h.n = m;
}

void mul(const Expansion& a, const Expansion& b, Expansion& out) {
  static thread_local Expansion acc, term, tmp;
  acc.n = 1;
  acc.v[0] = 0.0;
  for (int j = 0; j < b.n; ++j) {
    scale(a, b.v[j], term);
    add(acc, term, tmp);
    acc.n = tmp.n;
    for (int i = 0; i < tmp.n; ++i) acc.v[i] = tmp.v[i];
  }
  out.n = acc.n;
  for (int i = 0; i < acc.n; ++i) out.v[i] = acc.v[i];
}

inline void diff2(double a, double b, Expansion& out) {  // a - b exactly
  double s, e;
  two_sum(a, -b, s, e);
  out.n = 0;
  if (e != 0.0) out.v[out.n++] = e;
  out.v[out.n++] = s;
}

inline void negate(Expansion& e) {
  for (int i = 0; i < e.n; ++i) e.v[i] = -e.v[i];
}

inline int sign_of(const Expansion& e) {
  const double top = e.v[e.n - 1];
  return (top > 0.0) - (top < 0.0);
}

constexpr double kEps = std::numeric_limits<double>::epsilon() / 2.0;  // 2^-53
constexpr double kOrientBound = (3.0 + 16.0 * kEps) * kEps;
constexpr double kInCircleBound = (10.0 + 96.0 * kEps) * kEps;

int orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
  static thread_local Expansion acx, acy, bcx, bcy, l, r, d;
  diff2(ax, cx, acx);
  diff2(ay, cy, acy);
  diff2(bx, cx, bcx);
  diff2(by, cy, bcy);
  mul(acx, bcy, l);
  mul(acy, bcx, r);
  negate(r);
  add(l, r, d);
  return sign_of(d);
}

// > 0 when a, b, c are in counter-clockwise order
__attribute__((always_inline)) inline int orient2d(double ax, double ay, double bx, double by, double cx, double cy) {
  const double l = (ax - cx) * (by - cy);
  const double r = (ay - cy) * (bx - cx);
  const double det = l - r;
  const double mag = std::fabs(l) + std::fabs(r);
  if (std::fabs(det) > kOrientBound * mag) return (det > 0.0) - (det < 0.0);
  return orient_exact(ax, ay, bx, by, cx, cy);
}

__attribute__((noinline)) int incircle_exact(double ax, double ay, double bx, double by, double cx, double cy, double dx,
                   double dy) {
  static thread_local Expansion adx, ady, bdx, bdy, cdx, cdy, t0, t1, t2, al, bl, cl, ab, bc, ca,
      s0, s1, det;
  diff2(ax, dx, adx); diff2(ay, dy, ady);
  diff2(bx, dx, bdx); diff2(by, dy, bdy);
  diff2(cx, dx, cdx); diff2(cy, dy, cdy);
  auto lift = [&](const Expansion& x, const Expansion& y, Expansion& out) {
    mul(x, x, t0);
    mul(y, y, t1);
    add(t0, t1, out);
  };
  auto cross = [&](const Expansion& x1, const Expansion& y1, const Expansion& x2,
                   const Expansion& y2, Expansion& out) {  // x1*y2 - x2*y1
    mul(x1, y2, t0);
    mul(x2, y1, t1);
    negate(t1);
    add(t0, t1, out);
  };
  lift(adx, ady, al);
  lift(bdx, bdy, bl);
  lift(cdx, cdy, cl);
  cross(bdx, bdy, cdx, cdy, bc);
  cross(cdx, cdy, adx, ady, ca);
  cross(adx, ady, bdx, bdy, ab);
  mul(al, bc, s0);
  mul(bl, ca, s1);
  add(s0, s1, t2);
  mul(cl, ab, s0);
  add(t2, s0, det);
  return sign_of(det);
}

// > 0 when d lies inside the circle through a, b, c (a, b, c counter-clockwise)
__attribute__((always_inline)) inline int incircle(double ax, double ay, double bx, double by, double cx, double cy, double dx,
             double dy) {
  const double adx = ax - dx, ady = ay - dy, bdx = bx - dx, bdy = by - dy, cdx = cx - dx,
               cdy = cy - dy;
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, alift = adx * adx + ady * ady;
  const double cdxady = cdx * ady, adxcdy = adx * cdy, blift = bdx * bdx + bdy * bdy;
  const double adxbdy = adx * bdy, bdxady = bdx * ady, clift = cdx * cdx + cdy * cdy;
  const double det = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) +
                     clift * (adxbdy - bdxady);
  const double permanent = (std::fabs(bdxcdy) + std::fabs(cdxbdy)) * alift +
                           (std::fabs(cdxady) + std::fabs(adxcdy)) * blift +
                           (std::fabs(adxbdy) + std::fabs(bdxady)) * clift;
  if (std::fabs(det) > kInCircleBound * permanent) return (det > 0.0) - (det < 0.0);
  return incircle_exact(ax, ay, bx, by, cx, cy, dx, dy);
}

// Edges whose fourth point lies within Qhull's floating-point tolerance of the
// circumcircle (note (c) in Builder::run).  Measured with tools/qhull_margin.py (planted near
// co-circular quadruples in OMI-shaped swaths, ~4000 trials): Qhull's triangle set differs from
// the exact one only at margins r = |incircle| / (m^2 * 2 area) <= 5.7e-15, never in 1700
// trials with 1e-14 <= r < 3e-14; the threshold 2e-14 is 3.5 times the largest disagreement
// seen.  Shared by both builders.
int64_t count_near_ties(const double* x, const double* y, int64_t n, const int32_t* tri,
                        const int32_t* half, int64_t ntri) {
  double maxabs = 0.0;
  for (int64_t i = 0; i < n; ++i) maxabs = std::max(maxabs, std::max(std::fabs(x[i]), std::fabs(y[i])));
  const double tol = 2e-14 * std::max(maxabs * maxabs, 1e-300);
  int64_t ties = 0;
  for (int64_t a = 0; a < 3 * ntri; ++a) {
    const int32_t b = half[a];
    if (b < a) continue;  // hull edge (-1) or already visited twin
    const int32_t a0 = (int32_t)(a - a % 3), b0 = b - b % 3;
    const int32_t p0 = tri[a0 + (a + 2) % 3], pr = tri[a], pl = tri[a0 + (a + 1) % 3],
                  p1 = tri[b0 + (b + 2) % 3];
    const double adx = x[p0] - x[p1], ady = y[p0] - y[p1], bdx = x[pl] - x[p1],
                 bdy = y[pl] - y[p1], cdx = x[pr] - x[p1], cdy = y[pr] - y[p1];
    const double det = (adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) +
                       (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) +
                       (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
    const double area2 = std::fabs((x[p0] - x[pr]) * (y[pl] - y[pr]) -
                                   (y[p0] - y[pr]) * (x[pl] - x[pr]));
    // det / (2*area) = signed height of p1's lift above the plane of the other three
    if (std::fabs(det) <= tol * area2) ++ties;
  }
  return ties;
}

// --------------------------------------------------------------- triangulation
struct Builder {
  const double* x;
  const double* y;
  int64_t n;
  std::vector<int32_t> tri;    // 3 per triangle
  std::vector<int32_t> half;   // opposite half-edge of each triangle edge, -1 on the hull
  std::vector<int32_t> hprev, hnext, htri, hash;
  std::vector<int32_t> stack;
  int64_t ntri = 0;
  int64_t ties = 0;
  bool near_ties = true;   // also report near-ties (note (c) in run)
  int32_t hull_start = 0;
  double cx = 0, cy = 0;
  int hash_size = 0;

  static double pseudo_angle(double dx, double dy) {
    const double p = dx / (std::fabs(dx) + std::fabs(dy));
    return (dy > 0.0 ? 3.0 - p : 1.0 + p) / 4.0;  // [0, 1)
  }
  int hash_key(double px, double py) const {
    return (int)std::floor(pseudo_angle(px - cx, py - cy) * hash_size) % hash_size;
  }
  int orient(int32_t a, int32_t b, int32_t c) {
    const int s = orient2d(x[a], y[a], x[b], y[b], x[c], y[c]);
    return s;
  }
  void link(int32_t a, int32_t b) {
    half[a] = b;
    if (b != -1) half[b] = a;
  }
  int32_t add_triangle(int32_t i0, int32_t i1, int32_t i2, int32_t a, int32_t b, int32_t c) {
    const int32_t t = (int32_t)(3 * ntri);
    tri[t] = i0; tri[t + 1] = i1; tri[t + 2] = i2;
    link(t, a); link(t + 1, b); link(t + 2, c);
    ++ntri;
    return t;
  }

  // restore the Delaunay property around half-edge a (iterative edge flipping)
  int32_t legalize(int32_t a) {
    int32_t ar = 0;
    stack.clear();
    for (;;) {
      const int32_t b = half[a];
      // a is edge (pr -> pl) of the new triangle, b its twin in the older one:
      //
      //           pl                    pl
      //          /||\                  /  \
      //       al/ || \bl            al/    \a
      //        /  ||  \              /      \
      //       /  a||b  \    flip    /___ar___\
      //     p0\   ||   /p1   =>   p0\---bl---/p1
      //        \  ||  /              \      /
      //       ar\ || /br             b\    /br
      //          \||/                  \  /
      //           pr                    pr
      const int32_t a0 = a - a % 3;
      ar = a0 + (a + 2) % 3;
      if (b == -1) {
        if (stack.empty()) break;
        a = stack.back();
        stack.pop_back();
        continue;
      }
      const int32_t b0 = b - b % 3;
      const int32_t al = a0 + (a + 1) % 3;
      const int32_t bl = b0 + (b + 2) % 3;
      const int32_t p0 = tri[ar], pr = tri[a], pl = tri[al], p1 = tri[bl];
      // triangles are stored clockwise in this construction (see the seed), so
      // (p0, pl, pr) is the counter-clockwise triple the predicate expects
      const int s = incircle(x[p0], y[p0], x[pl], y[pl], x[pr], y[pr], x[p1], y[p1]);
      if (s == 0) ++ties;
      if (s > 0) {
        tri[a] = p1;
        tri[b] = p0;
        const int32_t hbl = half[bl];
        // the flipped edge may have been on the hull: repair the hull's triangle reference
        if (hbl == -1) {
          int32_t e = hull_start;
          do {
            if (htri[e] == bl) { htri[e] = a; break; }
            e = hprev[e];
          } while (e != hull_start);
        }
        link(a, hbl);
        link(b, half[ar]);
        link(ar, bl);
        const int32_t br = b0 + (b + 1) % 3;
        stack.push_back(br);
      } else {
        if (stack.empty()) break;
        a = stack.back();
        stack.pop_back();
      }
    }
    return ar;
  }

  // returns 0 on success, <0 when no triangle exists (all points collinear / < 3 distinct)
  int run() {
    if (n < 3) return -1;
    tri.assign(3 * std::max<int64_t>(2 * n - 5, 1), 0);
    half.assign(tri.size(), -1);
    hprev.assign(n, 0); hnext.assign(n, 0); htri.assign(n, 0);
    hash_size = (int)std::ceil(std::sqrt((double)n));
    hash.assign(hash_size, -1);
    std::vector<int32_t> ids(n);
    std::iota(ids.begin(), ids.end(), 0);

    double minx = x[0], maxx = x[0], miny = y[0], maxy = y[0];
    for (int64_t i = 1; i < n; ++i) {
      minx = std::min(minx, x[i]); maxx = std::max(maxx, x[i]);
      miny = std::min(miny, y[i]); maxy = std::max(maxy, y[i]);
    }
    const double mx = 0.5 * (minx + maxx), my = 0.5 * (miny + maxy);
    auto d2 = [&](double ax, double ay, double bx, double by) {
      return (ax - bx) * (ax - bx) + (ay - by) * (ay - by);
    };
    // seed: the point nearest the bbox centre, its nearest neighbour, and the third
    // point giving the smallest circumcircle
    int32_t i0 = 0, i1 = -1, i2 = -1;
    double best = std::numeric_limits<double>::infinity();
    for (int64_t i = 0; i < n; ++i) {
      const double d = d2(mx, my, x[i], y[i]);
      if (d < best) { best = d; i0 = (int32_t)i; }
    }
    best = std::numeric_limits<double>::infinity();
    for (int64_t i = 0; i < n; ++i) {
      if (i == i0) continue;
      const double d = d2(x[i0], y[i0], x[i], y[i]);
      if (d < best && d > 0.0) { best = d; i1 = (int32_t)i; }
    }
    if (i1 < 0) return -2;
    auto circumradius2 = [&](int32_t a, int32_t b, int32_t c) {
      const double dx = x[b] - x[a], dy = y[b] - y[a], ex = x[c] - x[a], ey = y[c] - y[a];
      const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = dx * ey - dy * ex;
      if (d == 0.0) return std::numeric_limits<double>::infinity();
      const double ux = (ey * bl - dy * cl) * 0.5 / d, uy = (dx * cl - ex * bl) * 0.5 / d;
      const double r = ux * ux + uy * uy;
      return (bl > 0.0 && cl > 0.0 && r == r) ? r : std::numeric_limits<double>::infinity();
    };
    best = std::numeric_limits<double>::infinity();
    for (int64_t i = 0; i < n; ++i) {
      if (i == i0 || i == i1) continue;
      const double r = circumradius2(i0, i1, (int32_t)i);
      if (r < best) { best = r; i2 = (int32_t)i; }
    }
    if (i2 < 0 || !(best < std::numeric_limits<double>::infinity())) return -2;
    // store the seed clockwise (orient < 0): the hull list then runs clockwise and a
    // point outside sees edge e -> next(e) iff it lies to its left, orient(p, e, next(e)) > 0
    if (orient(i0, i1, i2) > 0) std::swap(i1, i2);
    if (orient(i0, i1, i2) == 0) return -2;
    {
      const double ax = x[i0], ay = y[i0];
      const double dx = x[i1] - ax, dy = y[i1] - ay, ex = x[i2] - ax, ey = y[i2] - ay;
      const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = dx * ey - dy * ex;
      cx = ax + (ey * bl - dy * cl) * 0.5 / d;
      cy = ay + (dx * cl - ex * bl) * 0.5 / d;
    }
    std::vector<double> dist(n);
    for (int64_t i = 0; i < n; ++i) dist[i] = d2(x[i], y[i], cx, cy);
    std::sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) {
      if (dist[a] != dist[b]) return dist[a] < dist[b];
      if (x[a] != x[b]) return x[a] < x[b];
      if (y[a] != y[b]) return y[a] < y[b];
      return a < b;
    });

    hull_start = i0;
    hnext[i0] = hprev[i2] = i1;
    hnext[i1] = hprev[i0] = i2;
    hnext[i2] = hprev[i1] = i0;
    htri[i0] = 0; htri[i1] = 1; htri[i2] = 2;
    hash[hash_key(x[i0], y[i0])] = i0;
    hash[hash_key(x[i1], y[i1])] = i1;
    hash[hash_key(x[i2], y[i2])] = i2;
    add_triangle(i0, i1, i2, -1, -1, -1);

    double xp = 0, yp = 0;
    bool have_prev = false;
    for (int64_t k = 0; k < n; ++k) {
      const int32_t i = ids[k];
      const double px = x[i], py = y[i];
      if (have_prev && px == xp && py == yp) continue;  // exact duplicate: not a vertex
      xp = px; yp = py; have_prev = true;
      if (i == i0 || i == i1 || i == i2) continue;
      if (px != px || py != py) continue;

      // a hull vertex near the direction of the new point
      int32_t start = 0;
      const int key = hash_key(px, py);
      for (int j = 0; j < hash_size; ++j) {
        start = hash[(key + j) % hash_size];
        if (start != -1 && start != hnext[start]) break;
      }
      start = hprev[start];
      int32_t e = start, q;
      // first hull edge (e -> q) that the point sees
      for (;;) {
        q = hnext[e];
        const int s = orient(i, e, q);
        if (s > 0) break;
        e = q;
        if (e == start) { e = -1; break; }
      }
      if (e == -1) continue;  // on/inside the hull numerically: a (near-)duplicate

      int32_t t = add_triangle(e, i, hnext[e], -1, -1, htri[e]);
      htri[i] = legalize(t + 2);
      htri[e] = t;

      // fan forward along the hull while edges stay visible
      int32_t nx = hnext[e];
      for (;;) {
        q = hnext[nx];
        const int s = orient(i, nx, q);
        if (!(s > 0)) break;
        t = add_triangle(nx, i, q, htri[i], -1, htri[nx]);
        htri[i] = legalize(t + 2);
        hnext[nx] = nx;  // removed from the hull
        nx = q;
      }
      // and backward, if the first visible edge was found at the walk's start
      if (e == start) {
        for (;;) {
          q = hprev[e];
          const int s = orient(i, q, e);
          if (!(s > 0)) break;
          t = add_triangle(q, i, e, -1, htri[e], htri[q]);
          legalize(t + 2);
          htri[q] = t;
          hnext[e] = e;  // removed from the hull
          e = q;
        }
      }
      hull_start = hprev[i] = e;
      hnext[e] = hprev[nx] = i;
      hnext[i] = nx;
      hash[hash_key(px, py)] = i;
      hash[hash_key(x[e], y[e])] = e;
    }
    // A collinear triple met while the hull advances is harmless (the point is
    // simply not "visible" from that edge; float32 swath coordinates produce such
    // triples routinely along a scan position).  What makes the result non-unique
    // is (a) co-circular quadruples, counted in legalize(), and (b) collinear
    // triples on the FINAL hull, where Qhull may or may not emit a flat facet.
    int32_t e = hull_start;
    do {
      if (orient(hprev[e], e, hnext[e]) == 0) ++ties;
      e = hnext[e];
    } while (e != hull_start);
    // (c) NEAR ties.  Qhull works in floating point with a distance tolerance on the
    // lifted paraboloid that is relative to the GLOBAL coordinate range; a quadruple
    // whose fourth point is closer to the circumcircle than that tolerance may be
    // split along the other diagonal there (verified with rational arithmetic: Qhull
    // then returns a non-Delaunay pair).  Such edges are reported as ties too, so the
    // caller can take Qhull's answer instead of the exact one.
    if (near_ties) ties += count_near_ties(x, y, n, tri.data(), half.data(), ntri);
    return 0;
  }
};

#include "delaunay_swath.inl"  // LatticeBuilder: structured-swath builder (uses the predicates above)
#include "delaunay_seed.inl"   // SeedBuilder: the lattice itself as the seed of a flip-only construction

}  // namespace

extern "C" int64_t oisat_h_delaunay_swath(const double* h_x, const double* h_y, int64_t n_rows,
                                          int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                                          int64_t* n_ties, int32_t* path) {
  if (!h_x || !h_y || !h_tri || n_rows < 1 || n_cols < 1 ||
      n_rows * n_cols > (int64_t)0x3fffffff)
    return OISAT_E_ARG;
  LatticeBuilder g;
  g.x = h_x;
  g.y = h_y;
  g.rows = n_rows;
  g.cols = n_cols;
  if (g.run() == 0) {
    const int64_t m = g.emit(h_tri, tri_capacity);
    if (m < 0) return OISAT_E_ARG;
    if (n_ties)
      *n_ties = g.ties + count_near_ties(h_x, h_y, g.n, g.tri.data(), g.half.data(), g.ntri);
    if (path) *path = 1;
    return m;
  }
  if (path) *path = 0;
  return oisat_h_delaunay(h_x, h_y, n_rows * n_cols, h_tri, tri_capacity, n_ties);
}

// As oisat_h_delaunay_swath, but the near-tie scan (count_near_ties: one pass over every
// interior edge, ~15% of the builder's time) is left to the device: the twin half-edge of
// every emitted triangle edge goes to h_half (3 per triangle, -1 on the hull) and
// *n_ties reports the exact hull ties only; oisat_near_ties finishes the report.
// *path = 1: h_half is valid.  *path = 0: the lattice builder did not apply, the general
// builder ran, *n_ties is the complete report and h_half is not written.
extern "C" int64_t oisat_h_delaunay_swath_adj(const double* h_x, const double* h_y, int64_t n_rows,
                                              int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                                              int32_t* h_half, int64_t* n_ties, int32_t* path) {
  if (!h_x || !h_y || !h_tri || !h_half || n_rows < 1 || n_cols < 1 ||
      n_rows * n_cols > (int64_t)0x3fffffff)
    return OISAT_E_ARG;
  // one builder per thread, reused: the plan pool's threads live as long as the process, and a
  // builder that keeps its 5 MB of work arrays does not map and page-fault them again for every
  // granule (2.7 of 37 ms on the build host)
  static thread_local LatticeBuilder g;
  g.x = h_x;
  g.y = h_y;
  g.rows = n_rows;
  g.cols = n_cols;
  if (g.run() == 0) {
    // compact numbering of the surviving triangles (emit() skips the former ghosts)
    std::vector<int32_t> slot((size_t)g.ntri, -1);
    int64_t m = 0;
    for (int64_t t = 0; t < g.ntri; ++t)
      if (g.tri[3 * t] >= 0) slot[(size_t)t] = (int32_t)m++;
    if (m > tri_capacity) { g.release_if_large(); return OISAT_E_ARG; }
    for (int64_t t = 0; t < g.ntri; ++t) {
      const int32_t u = slot[(size_t)t];
      if (u < 0) continue;
      for (int e = 0; e < 3; ++e) {
        h_tri[3 * u + e] = g.tri[3 * t + e];
        const int32_t h = g.half[3 * t + e];
        h_half[3 * u + e] = h < 0 ? -1 : 3 * slot[(size_t)(h / 3)] + h % 3;
      }
    }
    if (n_ties) *n_ties = g.ties;
    if (path) *path = 1;
    g.release_if_large();
    return m;
  }
  g.release_if_large();
  if (path) *path = 0;
  return oisat_h_delaunay(h_x, h_y, n_rows * n_cols, h_tri, tri_capacity, n_ties);
}

extern "C" int64_t oisat_h_delaunay(const double* h_x, const double* h_y, int64_t n,
                                    int32_t* h_tri, int64_t tri_capacity, int64_t* n_ties) {
  if (!h_x || !h_y || !h_tri || n < 3 || n > (int64_t)0x3fffffff) return OISAT_E_ARG;
  // scipy.spatial.Delaunay raises on NaN / infinite coordinates; the reference then skips
  // the granule (interpolator.py:152-155)
  for (int64_t i = 0; i < n; ++i)
    if (!(std::fabs(h_x[i]) <= 1.7e308) || !(std::fabs(h_y[i]) <= 1.7e308)) return OISAT_E_UNSUPPORTED;
  Builder b;
  b.x = h_x;
  b.y = h_y;
  b.n = n;
  const int rc = b.run();
  if (rc != 0) return OISAT_E_UNSUPPORTED;  // degenerate input: no triangle (scipy raises too)
  if (b.ntri > tri_capacity) return OISAT_E_ARG;
  std::copy(b.tri.begin(), b.tri.begin() + 3 * b.ntri, h_tri);
  if (n_ties) *n_ties = b.ties;
  return b.ntri;
}

// Lattice-seeded construction (delaunay_seed.inl), whole on the host.  flags bit 0: finish
// with Lawson's flips (the result is then the Delaunay triangulation; without it the SEED is
// returned, a valid triangulation of the convex hull); bit 1: add the near-tie scan to *n_ties
// (only meaningful with bit 0).  h_half as in oisat_h_delaunay_swath_adj.  info (may be NULL)
// receives {seeded quads, seam vertices, triangles outside the lattice, host flips, declining
// check}.  Returns the number of triangles, or 0 when the construction does not apply to this
// lattice (a check of delaunay_seed.inl declined: the caller uses oisat_h_delaunay_swath_adj).
extern "C" int64_t oisat_h_delaunay_seed(const double* h_x, const double* h_y, int64_t n_rows,
                                         int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                                         int32_t* h_half, int64_t* n_ties, int32_t flags,
                                         int64_t* info) {
  if (!h_x || !h_y || !h_tri || !h_half || n_rows < 1 || n_cols < 1 ||
      n_rows * n_cols > (int64_t)0x3fffffff)
    return OISAT_E_ARG;
  static thread_local SeedBuilder g;
  g.x = h_x;
  g.y = h_y;
  g.rows = n_rows;
  g.cols = n_cols;
  const int rc = g.seed();
  if (info) { info[0] = g.n_quads; info[1] = g.n_seam; info[2] = g.n_outside; info[3] = 0; info[4] = g.why * 100 - g.rec_fail; }
  if (rc != 0) { g.release_if_large(); return 0; }
  if (g.ntri > tri_capacity) { g.release_if_large(); return OISAT_E_ARG; }
  g.assemble();
  if (flags & 1) g.lawson();
  if (info) info[3] = g.flips;
  std::copy(g.tri.begin(), g.tri.begin() + 3 * g.ntri, h_tri);
  std::copy(g.half.begin(), g.half.begin() + 3 * g.ntri, h_half);
  if (n_ties) {
    *n_ties = g.ties;
    if ((flags & 3) == 3)
      *n_ties += count_near_ties(h_x, h_y, g.n, g.tri.data(), g.half.data(), g.ntri);
  }
  const int64_t nt = g.ntri;
  g.release_if_large();
  return nt;
}

// The host's share of the device construction (k12_flip.cu): h_qtri[(n_rows-1)*(n_cols-1)]
// receives, per lattice quad, the index of the first of its two triangles (-1: the quad is
// not part of the seed); h_otri / h_ohalf (3 per triangle, out_capacity triangles) the
// triangles outside the lattice and their twins in the numbering of the result.  info
// receives {seeded quads, seam vertices, outside triangles, sigma, declining check},
// *max_abs_coord (may be NULL) the largest |coordinate| (oisat_near_ties scales with it).
// Returns the number of triangles of the seed, 0 when the construction does not apply.
extern "C" int64_t oisat_h_delaunay_seed_parts(const double* h_x, const double* h_y, int64_t n_rows,
                                               int64_t n_cols, int32_t* h_qtri, int32_t* h_otri,
                                               int32_t* h_ohalf, int64_t out_capacity,
                                               int64_t* info, double* max_abs_coord) {
  if (!h_x || !h_y || !h_qtri || !h_otri || !h_ohalf || !info || n_rows < 1 || n_cols < 1 ||
      n_rows * n_cols > (int64_t)0x3fffffff)
    return OISAT_E_ARG;
  static thread_local SeedBuilder g;
  g.x = h_x;
  g.y = h_y;
  g.rows = n_rows;
  g.cols = n_cols;
  const int rc = g.seed();
  info[0] = g.n_quads; info[1] = g.n_seam; info[2] = g.n_outside; info[3] = g.sigma;
  info[4] = g.why * 100 - g.rec_fail;
  if (max_abs_coord) *max_abs_coord = g.maxabs;
  // (exact ties met while triangulating the SEAM do not matter: that triangulation only has
  // to be valid, not Delaunay; a tie of the final mesh leaves an edge the device's filter
  // cannot decide and is reported there)
  if (rc != 0) { g.release_if_large(); return 0; }
  if (g.n_outside > out_capacity) { g.release_if_large(); return OISAT_E_ARG; }
  std::copy(g.qtri.begin(), g.qtri.end(), h_qtri);
  std::copy(g.otri.begin(), g.otri.end(), h_otri);
  std::copy(g.ohalf.begin(), g.ohalf.end(), h_ohalf);
  const int64_t nt = g.ntri;
  g.release_if_large();
  return nt;
}

// The rounds of oisat_flip_delaunay replayed on the host, threads one after the other (the
// same per-edge steps, flip_rounds.h): what the CPU tests run, and the model the device
// kernel is compared with.  h_tri / h_half are changed in place; result as on the device:
// {rounds, flips, edges certainly not Delaunay, edges the filter cannot decide}.
extern "C" int oisat_h_flip_rounds(const double* h_x, const double* h_y, int32_t* h_tri,
                                   int32_t* h_half, int64_t n_tri, int64_t max_rounds,
                                   int64_t* result) {
  if (!h_x || !h_y || !h_tri || !h_half || !result || n_tri < 0 || max_rounds < 1 ||
      max_rounds > 65536 || 3 * n_tri > (int64_t)0x7ffffff0)
    return OISAT_E_ARG;
  std::vector<int32_t> stamp((size_t)n_tri, -1), cand((size_t)(3 * n_tri), 0),
      list0((size_t)n_tri), list1((size_t)n_tri), edges((size_t)(3 * n_tri / 2 + 1));
  std::vector<unsigned long long> owner((size_t)n_tri, 0ull);
  std::vector<unsigned int> n_listed((size_t)max_rounds, 0u), n_marked((size_t)max_rounds, 0u);
  oisat_flip::Mesh m{h_tri, h_half, 3 * n_tri, stamp.data(), cand.data(), owner.data(), 0};
  oisat_flip::Lists l{{list0.data(), list1.data()}, edges.data(), n_listed.data(), n_marked.data()};
  auto P = [&](int v, int axis) { return axis ? h_y[v] : h_x[v]; };
  struct HostOps {
    unsigned long long max(unsigned long long* p, unsigned long long v) const {
      const unsigned long long o = *p;
      if (v > o) *p = v;
      return o;
    }
    unsigned int add(unsigned int* p, unsigned int v) const { const unsigned int o = *p; *p = o + v; return o; }
    int32_t exch(int32_t* p, int32_t v) const { const int32_t o = *p; *p = v; return o; }
  } ops;
  int64_t round = 0, flips = 0;
  for (; round < max_rounds; ++round) {
    const int r = (int)round;
    if (r == 0)
      for (int64_t e = 0; e < m.n_half; ++e) oisat_flip::mark_edge(m, l, (int32_t)e, r, P, ops);
    else
      for (unsigned int i = 0; i < n_listed[(size_t)r - 1]; ++i) {
        const int32_t t = ((r & 1) ? list1 : list0)[i];
        for (int e = 0; e < 3; ++e) oisat_flip::mark_edge(m, l, 3 * t + e, r, P, ops);
      }
    for (unsigned int i = 0; i < n_marked[(size_t)r]; ++i)
      flips += oisat_flip::apply_edge(m, l, edges[i], r, ops);
    if (n_marked[(size_t)r] == 0) { ++round; break; }
  }
  int64_t bad = 0, unsure = 0;
  for (int64_t a = 0; a < m.n_half; ++a) {
    const int c = oisat_flip::check_edge(h_tri, h_half, a, P);
    bad += c & 1;
    unsure += (c >> 1) & 1;
  }
  result[0] = round; result[1] = flips; result[2] = bad; result[3] = unsure;
  return OISAT_OK;
}
