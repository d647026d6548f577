// Lattice-seeded triangulation of a swath; included by delaunay.cpp inside its anonymous
// namespace after delaunay_swath.inl (it uses LatticeBuilder and the exact predicates).
//
// The incremental builder spends its time locating and splitting: 98,640 insertions for an OMI
// granule although the lattice the pixels come in IS already a triangulation of almost the
// whole footprint -- every lattice quad whose two triangles are properly oriented -- and turning
// a triangulation into the Delaunay one needs nothing but edge flips (Lawson), which are
// independent of each other wherever they do not share a triangle: work for the device
// (k12_flip.cu).  What the lattice does not give is the rest of the convex hull: the pockets
// between the curved outline of the swath and its hull, the gap a date-line crossing tears
// into the lattice (quads that span the map are not part of the seed), folded or degenerate
// quads.  That part is small and is done here, exactly:
//
//   1. every quad is tested with the exact orientation predicate; a quad joins the seed when
//      both of its triangles turn the way the majority does and none of its edges spans more
//      than half of the x range (the tear of a wrapped coordinate);
//   2. the SEAM vertices -- every vertex that is not surrounded by four seeded quads: the
//      outline, both lips of a tear, the rim of a dropped quad, ~3.5 % of an OMI granule -- are
//      triangulated on their own with the incremental builder (a few thousand points);
//   3. that triangulation must contain every seam EDGE (a lattice edge with a seeded quad on
//      exactly one side).  Then its triangles lie either inside or outside the seeded region,
//      and a flood fill from the seam edges tells which; the outside ones close the seed to the
//      convex hull.  Anything that contradicts this picture -- a missing seam edge, two labels
//      meeting across a free edge, seeded regions that overlap (their lips would cross or
//      nest), an area that does not add up -- makes the function decline, and the caller
//      falls back to the incremental builder.
//
// The seed is a valid triangulation of the convex hull of all points (repeated points are
// not vertices, as in the other builders); flips of strictly non-Delaunay edges keep it one
// and end at the Delaunay triangulation.  lawson() does them serially on the host (the
// reference for the device rounds, and the path taken without a device).
struct SeedBuilder {
  enum { TOP = 0, LEFT = 1, RIGHT = 2, BOTTOM = 3 };
  const double* x;
  const double* y;
  int64_t rows = 0, cols = 0, n = 0;
  int sigma = 0;                         // +1: (r,c) -> (r,c+1) -> (r+1,c) is counter-clockwise
  std::vector<uint8_t> quad, seam, edge_is_seam, pad;
  std::vector<int8_t> o1, o2;
  std::vector<int32_t> qtri, seq, ord, hint, label, newidx, queue, stack, crossed;
  std::vector<int32_t> otri, ohalf;      // the triangles outside the lattice (result numbering)
  std::vector<int32_t> tri, half;        // the result: real triangles only, -1 on the hull
  LatticeBuilder sub;                    // triangulation of the seam vertices
  int64_t ntri = 0, ties = 0, flips = 0, n_quads = 0, n_seam = 0, n_outside = 0;
  int rec_fail = 0;
  double maxabs = 0.0;                   // largest |coordinate| (the near-tie scan scales with it)
  int why = 0;                           // which check declined (diagnostic)

  static int32_t next(int32_t e) { return e % 3 == 2 ? e - 2 : e + 1; }
  static int32_t prev(int32_t e) { return e % 3 == 0 ? e + 2 : e - 1; }
  bool valid(int64_t r, int64_t c) const {
    return r >= 0 && c >= 0 && r < rows - 1 && c < cols - 1 && quad[(size_t)(r * (cols - 1) + c)];
  }
  // half-edge of seeded quad (r, c) that lies on the given side of it
  int32_t slot(int64_t r, int64_t c, int side) const {
    const int32_t t0 = 3 * qtri[(size_t)(r * (cols - 1) + c)], t1 = t0 + 3;
    if (sigma > 0) return side == TOP ? t0 : side == LEFT ? t0 + 2 : side == RIGHT ? t1 : t1 + 1;
    return side == TOP ? t0 + 2 : side == LEFT ? t0 : side == RIGHT ? t1 + 2 : t1 + 1;
  }
  struct Flank { int64_t rl, cl; int sl; int64_t rr, cr; int sr; };
  // quads to the left and right of the directed lattice edge u -> v (and which side of the
  // quad the edge is); false when u and v are not lattice neighbours
  bool flank(int32_t u, int32_t v, Flank& f) const {
    const int64_t ru = u / cols, cu = u % cols, rv = v / cols, cv = v % cols;
    if (ru == rv && cv == cu + 1) f = Flank{ru, cu, TOP, ru - 1, cu, BOTTOM};
    else if (ru == rv && cv == cu - 1) f = Flank{ru - 1, cv, BOTTOM, ru, cv, TOP};
    else if (cu == cv && rv == ru + 1) f = Flank{ru, cu - 1, RIGHT, ru, cu, LEFT};
    else if (cu == cv && rv == ru - 1) f = Flank{rv, cu, LEFT, rv, cu - 1, RIGHT};
    else return false;
    if (sigma < 0) {
      std::swap(f.rl, f.rr);
      std::swap(f.cl, f.cr);
      std::swap(f.sl, f.sr);
    }
    return true;
  }
  void link(int32_t a, int32_t b) { half[a] = b; half[b] = a; }
  double area2(int32_t a, int32_t b, int32_t c) const {
    return (x[b] - x[a]) * (y[c] - y[a]) - (y[b] - y[a]) * (x[c] - x[a]);
  }

  // half-edge p -> q of the seam triangulation, -1 when there is none
  int32_t find_edge(int32_t p, int32_t q) const {
    int32_t t = sub.vtri[(size_t)p];
    if (t < 0) return -1;
    int32_t h = sub.tri[(size_t)t] == p ? t : (sub.tri[(size_t)t + 1] == p ? t + 1 : t + 2);
    const int32_t h0 = h;
    int64_t guard = 0;
    do {
      if (sub.tri[(size_t)next(h)] == q) return h;
      h = sub.half[(size_t)prev(h)];            // next edge out of p, counter-clockwise
      if (++guard > 8 * n_seam + 64) return -1;
    } while (h != h0);
    return -1;
  }

  // replace edge p -> q (half-edge a) of the seam triangulation by the other diagonal of its
  // quadrilateral; false when that quadrilateral is not strictly convex
  bool flip_sub(int32_t a) {
    const int32_t b = sub.half[(size_t)a];
    const int32_t al = next(a), ar = prev(a), bl = prev(b), br = next(b);
    const int32_t pr = sub.tri[(size_t)a], pl = sub.tri[(size_t)al], p0 = sub.tri[(size_t)ar],
                  p1 = sub.tri[(size_t)bl];
    if (pr == sub.G || pl == sub.G || p0 == sub.G || p1 == sub.G) return false;
    if (sub.orient(p1, pl, p0) <= 0 || sub.orient(p0, pr, p1) <= 0) return false;
    const int32_t hbl = sub.half[(size_t)bl], har = sub.half[(size_t)ar];
    sub.tri[(size_t)a] = p1;
    sub.tri[(size_t)b] = p0;
    const int32_t a0 = a - a % 3, b0 = b - b % 3;
    for (int e = 0; e < 3; ++e) {
      sub.vtri[(size_t)sub.tri[(size_t)(a0 + e)]] = a0;
      sub.vtri[(size_t)sub.tri[(size_t)(b0 + e)]] = b0;
    }
    sub.link(a, hbl);
    sub.link(b, har);
    sub.link(ar, bl);
    return true;
  }

  // make u - v an edge of the seam triangulation; false when that cannot be done by flips
  // (a vertex lies exactly on the segment, or the walk leaves the hull)
  bool recover(int32_t u, int32_t v) {
    if (sub.vtri[(size_t)u] < 0 || sub.vtri[(size_t)v] < 0) return (rec_fail = 1, false);   // a repeated point
    if (find_edge(u, v) >= 0) return true;
    // the triangle (u, a, b) at u that the segment enters: counter-clockwise, so a is strictly
    // right of u -> v and b strictly left
    int32_t t = sub.vtri[(size_t)u];
    int32_t h = sub.tri[(size_t)t] == u ? t : (sub.tri[(size_t)t + 1] == u ? t + 1 : t + 2);
    const int32_t h0 = h;
    int32_t cross = -1;
    int64_t guard = 0;
    do {
      const int32_t a = sub.tri[(size_t)next(h)], b = sub.tri[(size_t)prev(h)];
      if (a != sub.G && b != sub.G && sub.orient(u, v, a) < 0 && sub.orient(u, v, b) > 0) {
        cross = next(h);
        break;
      }
      h = sub.half[(size_t)prev(h)];
      if (++guard > 8 * n_seam + 64) return (rec_fail = 2, false);
    } while (h != h0);
    if (cross < 0) return (rec_fail = 3, false);
    crossed.clear();
    for (;;) {
      // `cross` is half-edge a -> b of the triangle already passed: a right of u -> v, b left
      crossed.push_back(sub.tri[(size_t)cross]);
      crossed.push_back(sub.tri[(size_t)next(cross)]);
      const int32_t g = sub.half[(size_t)cross];          // b -> a in the next triangle (b, a, c)
      const int32_t c = sub.tri[(size_t)prev(g)];
      if (c == sub.G) return (rec_fail = 4, false);
      if (c == v) break;
      const int sc = sub.orient(u, v, c);
      if (sc == 0) return (rec_fail = 5, false);
      cross = sc > 0 ? next(g) : prev(g);                  // a -> c when c is left, c -> b when right
      if ((int64_t)crossed.size() > 16 * n_seam + 64) return (rec_fail = 6, false);
    }
    size_t head = 0;
    int64_t budget = 64 * (int64_t)crossed.size() + 1024;
    while (head < crossed.size()) {
      if (--budget < 0) return (rec_fail = 7, false);
      const int32_t a = crossed[head], b = crossed[head + 1];
      head += 2;
      const int32_t e = find_edge(a, b);
      if (e < 0) return (rec_fail = 8, false);
      const int32_t p0 = sub.tri[(size_t)prev(e)], p1 = sub.tri[(size_t)prev(sub.half[(size_t)e])];
      if (!flip_sub(e)) {                                  // not convex yet: later
        crossed.push_back(a);
        crossed.push_back(b);
        continue;
      }
      if ((p0 == u && p1 == v) || (p0 == v && p1 == u)) continue;
      if (p0 == u || p0 == v || p1 == u || p1 == v) continue;   // ends on the segment: cannot cross it
      const int s0 = sub.orient(u, v, p0), s1 = sub.orient(u, v, p1);
      if (s0 == 0 || s1 == 0) return (rec_fail = 9, false);
      if (s0 != s1) {                                      // the new diagonal still crosses u - v
        if (s0 > 0) { crossed.push_back(p0); crossed.push_back(p1); }
        else { crossed.push_back(p1); crossed.push_back(p0); }
      }
    }
    return find_edge(u, v) >= 0;
  }

  // 0: tri / half hold the seed.  < 0: this construction does not apply (why = the check).
  int seed() {
    why = 0;
    if (rows < 2 || cols < 2) return why = -1;
    n = rows * cols;
    if (n > (int64_t)0x1fffffff) return why = -1;
    double xmin = x[0], xmax = x[0], amax = 0.0;
    for (int64_t i = 0; i < n; ++i) {       // (NaN fails both comparisons below)
      xmin = std::min(xmin, x[i]);
      xmax = std::max(xmax, x[i]);
      amax = std::max(amax, std::max(std::fabs(x[i]), std::fabs(y[i])));
    }
    bool finite = amax <= 1e300;
    for (int64_t i = 0; finite && i < n; i += 4096) {   // max() drops NaN operands: look for them
      double sum = 0.0;
      const int64_t e = std::min(n, i + 4096);
      for (int64_t k = i; k < e; ++k) sum += x[k] * 0.0 + y[k] * 0.0;
      finite = sum == 0.0;
    }
    if (!finite) return why = -1;
    maxabs = amax;
    const double reach = 0.5 * (xmax - xmin);
    const int64_t qc = cols - 1, nq = (rows - 1) * qc;
    // ---- 1. quads
    o1.resize((size_t)nq);
    o2.resize((size_t)nq);
    int64_t pos = 0, neg = 0;
    for (int64_t r = 0; r + 1 < rows; ++r) {
      // the floating-point filter of orient2d for a whole line of quads, no branches (the loop
      // vectorises); the few signs it cannot certify (0 here) are decided exactly below
      const double* xa = x + r * cols;
      const double* ya = y + r * cols;
      const double* xc = xa + cols;
      const double* yc = ya + cols;
      int8_t* q1 = o1.data() + r * qc;
      int8_t* q2 = o2.data() + r * qc;
      for (int64_t c = 0; c < qc; ++c) {
        // (a, b, cc) and (b, d, cc) with a = (r, c), b = (r, c + 1), cc = (r + 1, c), d = (r + 1, c + 1)
        const double l1 = (xa[c] - xc[c]) * (ya[c + 1] - yc[c]), r1 = (ya[c] - yc[c]) * (xa[c + 1] - xc[c]);
        const double l2 = (xa[c + 1] - xc[c]) * (yc[c + 1] - yc[c]), r2 = (ya[c + 1] - yc[c]) * (xc[c + 1] - xc[c]);
        const double d1 = l1 - r1, d2 = l2 - r2;
        const bool k1 = std::fabs(d1) > kOrientBound * (std::fabs(l1) + std::fabs(r1));
        const bool k2 = std::fabs(d2) > kOrientBound * (std::fabs(l2) + std::fabs(r2));
        q1[c] = (int8_t)(k1 ? ((d1 > 0.0) - (d1 < 0.0)) : 0);
        q2[c] = (int8_t)(k2 ? ((d2 > 0.0) - (d2 < 0.0)) : 0);
      }
      for (int64_t c = 0; c < qc; ++c) {
        if (q1[c] == 0) q1[c] = (int8_t)orient2d(xa[c], ya[c], xa[c + 1], ya[c + 1], xc[c], yc[c]);
        if (q2[c] == 0) q2[c] = (int8_t)orient2d(xa[c + 1], ya[c + 1], xc[c + 1], yc[c + 1], xc[c], yc[c]);
        pos += (q1[c] > 0) + (q2[c] > 0);
        neg += (q1[c] < 0) + (q2[c] < 0);
      }
    }
    sigma = pos >= neg ? 1 : -1;
    double area_lat = 0.0;
    quad.assign((size_t)nq, 0);
    qtri.assign((size_t)nq, -1);
    n_quads = 0;
    for (int64_t r = 0; r + 1 < rows; ++r)
      for (int64_t c = 0; c < qc; ++c) {
        const size_t q = (size_t)(r * qc + c);
        if (o1[q] != sigma || o2[q] != sigma) continue;
        const int64_t a = r * cols + c, b = a + 1, cc = a + cols, d = cc + 1;
        if (std::fabs(x[a] - x[b]) > reach || std::fabs(x[a] - x[cc]) > reach ||
            std::fabs(x[b] - x[d]) > reach || std::fabs(x[cc] - x[d]) > reach ||
            std::fabs(x[b] - x[cc]) > reach)
          continue;
        quad[q] = 1;
        qtri[q] = (int32_t)(2 * n_quads++);
        area_lat += std::fabs(area2((int32_t)a, (int32_t)b, (int32_t)cc)) +
                    std::fabs(area2((int32_t)b, (int32_t)d, (int32_t)cc));
      }
    if (n_quads == 0) return why = -2;
    // ---- 2. seam vertices and seam edges
    seam.assign((size_t)n, 0);
    n_seam = 0;
    int64_t n_seam_edges = 0;
    {
      // quad validity with a ring of zeros around it: pad[(r + 1) * pc + (c + 1)] = valid(r, c)
      const int64_t pc = qc + 2;
      pad.assign((size_t)((rows + 1) * pc), 0);
      for (int64_t r = 0; r + 1 < rows; ++r)
        std::copy(quad.begin() + r * qc, quad.begin() + (r + 1) * qc, pad.begin() + (r + 1) * pc + 1);
      for (int64_t r = 0; r < rows; ++r) {
        const uint8_t* up = pad.data() + r * pc;          // quads (r - 1, .)
        const uint8_t* dn = up + pc;                      // quads (r, .)
        uint8_t* sv = seam.data() + r * cols;
        int64_t ns = 0, ne = 0;
        for (int64_t c = 0; c < cols; ++c) {              // quads (., c - 1) at [c], (., c) at [c + 1]
          const uint8_t v = (uint8_t)!(up[c] & up[c + 1] & dn[c] & dn[c + 1]);
          sv[c] = v;
          ns += v;
          ne += (c + 1 < cols) & (dn[c + 1] != up[c + 1]);   // edge (r, c) - (r, c + 1)
          ne += (r + 1 < rows) & (dn[c + 1] != dn[c]);       // edge (r, c) - (r + 1, c)
        }
        n_seam += ns;
        n_seam_edges += ne;
      }
    }
    if (n_seam * 3 > n) return why = -3;        // more seam than lattice: not worth seeding
    // ---- 3. the seam vertices as a chain: outline first, then the rest line by line;
    //         inserted coarse to fine along the chain (every 2^L-th, L descending), each
    //         next to a chain neighbour inserted before it
    seq.clear();
    seq.reserve((size_t)n_seam);
    auto take = [&](int64_t r, int64_t c) {
      const size_t v = (size_t)(r * cols + c);
      if (seam[v] == 1) { seam[v] = 2; seq.push_back((int32_t)v); }
    };
    for (int64_t c = 0; c < cols; ++c) take(0, c);
    for (int64_t r = 1; r < rows; ++r) take(r, cols - 1);
    for (int64_t c = cols - 2; c >= 0; --c) take(rows - 1, c);
    for (int64_t r = rows - 2; r >= 1; --r) take(r, 0);
    // the rest line by line, one side of the map after the other: the two lips of a date-line
    // tear are a map width apart, and a chain that alternates between them makes every
    // insertion walk across the slivers that span the gap (5 ms instead of 1.4 for such a granule)
    const double xmid = 0.5 * (xmin + xmax);
    for (int side = 0; side < 2; ++side)
      for (int64_t r = 1; r + 1 < rows; ++r)
        for (int64_t c = 1; c + 1 < cols; ++c)
          if (seam[(size_t)(r * cols + c)] == 1 && (x[r * cols + c] < xmid) == (side == 0)) take(r, c);
    const int64_t m = (int64_t)seq.size();
    ord.clear();
    hint.clear();
    ord.reserve((size_t)m);
    hint.reserve((size_t)m);
    {
      int top = 0;
      while ((int64_t(1) << top) < m) ++top;
      for (int L = top; L >= 0; --L) {
        const int64_t step = int64_t(1) << L;
        for (int64_t k = 0; k < m; k += step)
          if (L == top ? k == 0 : ((k >> L) & 1)) {
            ord.push_back(seq[(size_t)k]);
            hint.push_back(k >= step ? seq[(size_t)(k - step)] : seq[0]);
          }
      }
    }
    if ((int64_t)ord.size() != m || m < 3) return why = -4;
    // ---- 4. triangulate them
    sub.x = x;
    sub.y = y;
    sub.rows = rows;
    sub.cols = cols;
    sub.n = n;
    sub.G = (int32_t)n;
    sub.ntri = sub.flips = sub.ties = 0;
    sub.stack.clear();
    sub.vtri.assign((size_t)n + 1, -1);
    sub.tri.assign((size_t)(3 * (2 * m + 8)), -1);
    sub.half.assign((size_t)(3 * (2 * m + 8)), -1);
    {
      int32_t p0 = ord[0], p1 = ord[1], p2 = -1;
      if (x[p0] == x[p1] && y[p0] == y[p1]) return why = -5;
      int64_t k2 = 2;
      for (; k2 < m; ++k2) {
        const int s = sub.orient(p0, p1, ord[(size_t)k2]);
        if (s != 0) {
          p2 = ord[(size_t)k2];
          if (s < 0) std::swap(p0, p1);
          break;
        }
      }
      if (p2 < 0) return why = -5;
      sub.begin(p0, p1, p2);
      int32_t last = p2;
      const int64_t max_steps = 8 * m + 64;
      for (int64_t k = 2; k < m; ++k) {
        if (k == k2) continue;
        const int32_t p = ord[(size_t)k];
        int32_t from = hint[(size_t)k];
        if (sub.vtri[(size_t)from] < 0) from = last;
        else {
          const double dh = (x[from] - x[p]) * (x[from] - x[p]) + (y[from] - y[p]) * (y[from] - y[p]);
          const double dl = (x[last] - x[p]) * (x[last] - x[p]) + (y[last] - y[p]) * (y[last] - y[p]);
          if (dl < dh) from = last;
        }
        if (!sub.insert(p, from, max_steps)) return why = -6;
        if (sub.vtri[(size_t)p] >= 0) last = p;
      }
      if (sub.hull_ties() != 0) return why = -6;
    }
    // ---- 4b. seam edges the seam triangulation lacks.  The outline of a swath is always there
    //          (nothing lies beyond it), but the lips of a tear are staircases whose long
    //          cross-track steps have seam vertices close by on both sides: not Delaunay edges.
    //          They are put in by flipping away the edges that cross them (Sloan 1993); the
    //          result is still a triangulation of the seam vertices, which is all the seed needs.
    for (int64_t k = 0; k < m; ++k) {            // both ends of a seam edge are seam vertices
      const int32_t u = seq[(size_t)k];
      const int64_t r = u / cols, c = u % cols;
      if (c + 1 < cols && valid(r, c) != valid(r - 1, c) && !recover(u, u + 1)) return why = -8;
      if (r + 1 < rows && valid(r, c) != valid(r, c - 1) && !recover(u, u + (int32_t)cols)) return why = -8;
    }
    // ---- 5. inside / outside
    const int64_t T = sub.ntri;
    label.assign((size_t)T, 0);
    edge_is_seam.assign((size_t)(3 * T), 0);
    queue.clear();
    int64_t found = 0;
    Flank f;
    for (int64_t t = 0; t < T; ++t) {
      const int32_t s = (int32_t)(3 * t);
      if (!sub.real(s)) continue;
      for (int e = 0; e < 3; ++e) {
        const int32_t u = sub.tri[(size_t)(s + e)], v = sub.tri[(size_t)next(s + e)];
        if (!flank(u, v, f)) continue;
        const bool vl = valid(f.rl, f.cl), vr = valid(f.rr, f.cr);
        if (vl == vr) continue;
        edge_is_seam[(size_t)(s + e)] = 1;
        const int32_t lab = vl ? 1 : 2;            // a counter-clockwise triangle is left of its edges
        if (label[(size_t)t] == 0) { label[(size_t)t] = lab; queue.push_back((int32_t)t); }
        else if (label[(size_t)t] != lab) return why = -7;
        if (vl) ++found;
      }
    }
    if (found != n_seam_edges) return why = -8;    // a seam edge is not an edge of the seam triangulation
    for (size_t qi = 0; qi < queue.size(); ++qi) {
      const int32_t t = queue[qi];
      for (int e = 0; e < 3; ++e) {
        const int32_t h = 3 * t + e;
        if (edge_is_seam[(size_t)h]) continue;
        const int32_t g = sub.half[(size_t)h];
        const int32_t gt = g / 3;
        if (!sub.real(3 * gt)) {                   // beyond the hull is outside
          if (label[(size_t)t] != 2) return why = -9;
          continue;
        }
        if (label[(size_t)gt] == 0) { label[(size_t)gt] = label[(size_t)t]; queue.push_back(gt); }
        else if (label[(size_t)gt] != label[(size_t)t]) return why = -9;
      }
    }
    double area_in = 0.0;
    n_outside = 0;
    newidx.assign((size_t)T, -1);
    for (int64_t t = 0; t < T; ++t) {
      const int32_t s = (int32_t)(3 * t);
      if (!sub.real(s)) continue;
      if (label[(size_t)t] == 0) return why = -10;
      if (label[(size_t)t] == 1)
        area_in += area2(sub.tri[(size_t)s], sub.tri[(size_t)s + 1], sub.tri[(size_t)s + 2]);
      else
        newidx[(size_t)t] = (int32_t)(2 * n_quads + n_outside++);
    }
    // ---- 6. the triangles outside the lattice, in the numbering of the result (the 2 n_quads
    //         lattice triangles come first); a seam edge is linked to the slot of its quad
    ntri = 2 * n_quads + n_outside;
    otri.assign((size_t)(3 * n_outside), -1);
    ohalf.assign((size_t)(3 * n_outside), -1);
    for (int64_t t = 0; t < T; ++t) {
      const int32_t u = newidx[(size_t)t];
      if (u < 0) continue;
      const int32_t o = u - (int32_t)(2 * n_quads);
      for (int e = 0; e < 3; ++e) {
        const int32_t h = (int32_t)(3 * t + e);
        otri[(size_t)(3 * o + e)] = sub.tri[(size_t)h];
        if (edge_is_seam[(size_t)h]) {             // the seeded quad is on the right of this edge
          flank(sub.tri[(size_t)h], sub.tri[(size_t)next(h)], f);
          ohalf[(size_t)(3 * o + e)] = slot(f.rr, f.cr, f.sr);
        } else {
          const int32_t g = sub.half[(size_t)h];
          const int32_t w = newidx[(size_t)(g / 3)];
          ohalf[(size_t)(3 * o + e)] = w < 0 ? -1 : 3 * w + g % 3;
        }
      }
    }
    if (!(std::fabs(area_in - area_lat) <= 1e-9 * std::max(area_lat, 1e-300))) return why = -11;
    // ---- 7. every point must be a vertex of the seed.  A pixel that does not sit where the
    //         lattice puts it -- inside somebody else's quad: its own four quads are folded and
    //         dropped -- is a seam vertex whose triangles are all labelled "inside" and go away
    //         with them; the seed would be a valid triangulation of the OTHER points.  Two
    //         counts catch that and anything like it: a seam vertex must be a corner of a seeded
    //         quad or of an outside triangle (and must have been inserted: a repeated point is
    //         not), and the triangles must be as many as Euler's formula allows for n vertices
    //         and h hull edges, T = 2 n - 2 - h.
    for (int64_t k = 0; k < 3 * n_outside; ++k) seam[(size_t)otri[(size_t)k]] = 3;
    int64_t hull_edges = n_seam_edges;
    for (int64_t k = 0; k < 3 * n_outside; ++k) {
      const int32_t g = ohalf[(size_t)k];
      if (g < 0) ++hull_edges;
      else if (g < 6 * n_quads) --hull_edges;     // a seam edge with an outside triangle behind it
    }
    for (int64_t k = 0; k < m; ++k) {
      const int32_t v = seq[(size_t)k];
      if (sub.vtri[(size_t)v] < 0) return why = -12;
      if (seam[(size_t)v] == 3) continue;
      const int64_t r = v / cols, c = v % cols;
      if (!(valid(r - 1, c - 1) || valid(r - 1, c) || valid(r, c - 1) || valid(r, c))) return why = -12;
    }
    if (ntri != 2 * n - 2 - hull_edges) return why = -13;
    ties = sub.ties;
    flips = 0;
    return 0;
  }

  // The whole seed on the host: lattice triangles (what oisat_seed_assemble does on the
  // device, k12_flip.cu) followed by the outside ones.
  void assemble() {
    const int64_t qc = cols - 1;
    tri.assign((size_t)(3 * ntri), -1);
    half.assign((size_t)(3 * ntri), -1);
    for (int64_t r = 0; r + 1 < rows; ++r)
      for (int64_t c = 0; c < qc; ++c) {
        if (!valid(r, c)) continue;
        const int32_t a = (int32_t)(r * cols + c), b = a + 1, cc = a + (int32_t)cols, d = cc + 1;
        const int32_t t0 = 3 * qtri[(size_t)(r * qc + c)], t1 = t0 + 3;
        if (sigma > 0) {
          tri[t0] = a; tri[t0 + 1] = b; tri[t0 + 2] = cc;
          tri[t1] = b; tri[t1 + 1] = d; tri[t1 + 2] = cc;
          link(t0 + 1, t1 + 2);
        } else {
          tri[t0] = a; tri[t0 + 1] = cc; tri[t0 + 2] = b;
          tri[t1] = b; tri[t1 + 1] = cc; tri[t1 + 2] = d;
          link(t0 + 1, t1);
        }
        if (valid(r - 1, c)) link(slot(r, c, TOP), slot(r - 1, c, BOTTOM));
        if (valid(r, c - 1)) link(slot(r, c, LEFT), slot(r, c - 1, RIGHT));
      }
    const int64_t base = 6 * n_quads;
    for (int64_t h = 0; h < 3 * n_outside; ++h) {
      tri[(size_t)(base + h)] = otri[(size_t)h];
      const int32_t g = ohalf[(size_t)h];
      half[(size_t)(base + h)] = g;
      if (g >= 0 && g < base) half[(size_t)g] = (int32_t)(base + h);
    }
  }

  // Lawson's flips, serially: every edge is suspect once, and again when a flip next to it
  // changed a triangle it belongs to
  void lawson() {
    stack.clear();
    for (int32_t a = 0; a < (int32_t)(3 * ntri); ++a)
      if (half[(size_t)a] > a) stack.push_back(a);
    while (!stack.empty()) {
      const int32_t a = stack.back();
      stack.pop_back();
      const int32_t b = half[(size_t)a];
      if (b < 0) continue;
      const int32_t al = next(a), ar = prev(a), bl = prev(b), br = next(b);
      const int32_t pr = tri[(size_t)a], pl = tri[(size_t)al], p0 = tri[(size_t)ar], p1 = tri[(size_t)bl];
      if (incircle(x[pr], y[pr], x[pl], y[pl], x[p0], y[p0], x[p1], y[p1]) <= 0) continue;
      ++flips;
      const int32_t hbl = half[(size_t)bl], har = half[(size_t)ar];
      tri[(size_t)a] = p1;
      tri[(size_t)b] = p0;
      half[(size_t)a] = hbl;
      if (hbl >= 0) half[(size_t)hbl] = a;
      half[(size_t)b] = har;
      if (har >= 0) half[(size_t)har] = b;
      link(ar, bl);
      stack.push_back(a);
      stack.push_back(al);
      stack.push_back(b);
      stack.push_back(br);
    }
  }

  void release_if_large() {
    if (tri.capacity() > (size_t)8 << 20) {
      *this = SeedBuilder();
    }
  }
};
