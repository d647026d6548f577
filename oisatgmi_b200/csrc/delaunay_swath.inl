// Lattice-order Delaunay builder for swaths; included by delaunay.cpp inside its
// anonymous namespace (it uses the exact predicates and count_near_ties defined there).
//
// A level-2 swath is a curvilinear (scan line x ground pixel) lattice and the
// reference triangulates ALL of its pixel centres (interpolator.py:131-133; bad
// pixels only poison the data, :126-128).  A radial sweep is a poor fit for such a
// long thin band: measured 56 edge flips per point on OMI-shaped granules.
//
// This builder works for any 2-D lattice of points, whatever its shape in the
// (lon, lat) plane: incremental Delaunay insertion in a coarse-to-fine lattice
// order (see run()).  Consecutive points are lattice neighbours of the current
// level, so locating each new point is a walk of a few triangles from the previous
// one, and the flips that follow stay local (~5 per point).  Date-line crossings
// are not special: the walk is simply long once per crossing.
//
// The convex hull is closed with "ghost" triangles (u, v, GHOST) so that every
// edge has a twin and insertion outside the hull is the same 1 -> 3 split as
// inside; the flip rule for edges that end at the ghost vertex removes reflex hull
// corners, which fans the new point onto every hull edge it can see.  All
// decisions use the exact predicates, so the result is the Delaunay triangulation;
// with no ties that is the same triangle set Qhull returns.
struct LatticeBuilder {
  const double* x;
  const double* y;
  int64_t rows, cols, n;
  int32_t G = 0;                          // ghost vertex id (= n)
  std::vector<int32_t> tri, half, stack;  // every half-edge has a twin while ghosts exist
  std::vector<int32_t> vtri;              // per vertex: a triangle that has it (ghost or real), -1 before insertion
  std::vector<int32_t> ord, hull_next;    // work arrays kept as members: a builder that is reused
  std::vector<uint8_t> lev, seen;         // (one per thread, see delaunay.cpp) keeps their pages
  int64_t ntri = 0, flips = 0, ties = 0;

  static int32_t next(int32_t e) { return e % 3 == 2 ? e - 2 : e + 1; }
  static int32_t prev(int32_t e) { return e % 3 == 0 ? e + 2 : e - 1; }
  int orient(int32_t a, int32_t b, int32_t c) const {
    return orient2d(x[a], y[a], x[b], y[b], x[c], y[c]);
  }
  void link(int32_t a, int32_t b) { half[a] = b; half[b] = a; }
  int32_t new_triangle(int32_t a, int32_t b, int32_t c) {
    const int32_t s = (int32_t)(3 * ntri++);
    tri[s] = a; tri[s + 1] = b; tri[s + 2] = c;
    if (a != G) vtri[a] = s;
    if (b != G) vtri[b] = s;
    if (c != G) vtri[c] = s;
    return s;
  }
  bool real(int32_t s) const { return tri[s] != G && tri[s + 1] != G && tri[s + 2] != G; }

  // a REAL triangle that has vertex v (v inserted): the walk to a nearby new point starts there
  int32_t real_triangle_of(int64_t v) const {
    const int32_t t = vtri[v];
    if (real(t)) return t;
    for (int e = 0; e < 3; ++e)          // a ghost of a hull vertex: cross its one real edge
      if (tri[t + e] != G && tri[next(t + e)] != G) {
        const int32_t h = half[t + e];
        return h - h % 3;
      }
    return t;
  }

  // restore the Delaunay property (and hull convexity) behind the edges on `stack`;
  // every stacked edge has the newly inserted point as the apex of its triangle
  void relax() {
    while (!stack.empty()) {
      const int32_t a = stack.back();
      stack.pop_back();
      const int32_t b = half[a];
      const int32_t al = next(a), ar = prev(a), bl = prev(b), br = next(b);
      const int32_t pr = tri[a], pl = tri[al], p0 = tri[ar], p1 = tri[bl];
      bool flip;
      if (pr == G) flip = orient(p0, pl, p1) < 0;          // hull path p0 -> pl -> p1 turns right
      else if (pl == G) flip = orient(p1, pr, p0) < 0;     // hull path p1 -> pr -> p0 turns right
      else if (p0 == G || p1 == G) flip = false;           // a hull edge
      else flip = incircle(x[pr], y[pr], x[pl], y[pl], x[p0], y[p0], x[p1], y[p1]) > 0;
      if (!flip) continue;
      ++flips;
      const int32_t hbl = half[bl], har = half[ar];
      tri[a] = p1;
      tri[b] = p0;
      {   // both triangles changed: keep every vertex pointing at a triangle that still has it
        const int32_t a0 = a - a % 3, b0 = b - b % 3;
        for (int e = 0; e < 3; ++e) {
          if (tri[a0 + e] != G) vtri[tri[a0 + e]] = a0;
          if (tri[b0 + e] != G) vtri[tri[b0 + e]] = b0;
        }
      }
      link(a, hbl);
      link(b, har);
      link(ar, bl);
      stack.push_back(a);
      stack.push_back(br);
    }
  }

  // split triangle s (base half-edge) at p, strictly inside it (or in a ghost)
  int32_t split3(int32_t s, int32_t p) {
    const int32_t b = tri[s + 1], c = tri[s + 2], a = tri[s];
    const int32_t hb = half[s + 1], hc = half[s + 2];
    const int32_t s2 = new_triangle(b, c, p), s3 = new_triangle(c, a, p);
    tri[s + 2] = p;                       // (a, b, p)
    vtri[p] = s;
    if (c != G) vtri[c] = s2;             // c left triangle s
    link(s2, hb);
    link(s3, hc);
    link(s + 1, s2 + 2);
    link(s2 + 1, s3 + 2);
    link(s3 + 1, s + 2);
    stack.push_back(s);
    stack.push_back(s2);
    stack.push_back(s3);
    return real(s) ? s : (real(s2) ? s2 : s3);
  }

  // p lies on half-edge e (strictly between its end points)
  int32_t split4(int32_t e, int32_t p) {
    const int32_t f = half[e];
    const int32_t en = next(e), ep = prev(e), fn = next(f), fp = prev(f);
    const int32_t a = tri[e], b = tri[en], c = tri[ep], d = tri[fp];
    const int32_t hen = half[en], hfn = half[fn];
    const int32_t n1 = new_triangle(p, b, c), n2 = new_triangle(p, a, d);
    tri[en] = p;                          // (a, p, c)
    tri[fn] = p;                          // (b, p, d)
    vtri[p] = e - e % 3;
    if (b != G) vtri[b] = n1;             // b left (a, b, c), a left (b, a, d)
    if (a != G) vtri[a] = n2;
    link(n1 + 1, hen);
    link(n1 + 2, en);
    link(n2 + 1, hfn);
    link(n2 + 2, fn);
    link(e, n2);
    link(f, n1);
    stack.push_back(ep);
    stack.push_back(n1 + 1);
    stack.push_back(fp);
    stack.push_back(n2 + 1);
    const int32_t s = e - e % 3;
    return real(s) ? s : f - f % 3;
  }

  // first triangle (p0, p1, p2 counter-clockwise) and its three ghosts
  void begin(int32_t p0, int32_t p1, int32_t p2) {
    const int32_t t = new_triangle(p0, p1, p2);
    const int32_t g1 = new_triangle(p1, p0, G), g2 = new_triangle(p2, p1, G),
                  g3 = new_triangle(p0, p2, G);
    link(t, g1); link(t + 1, g2); link(t + 2, g3);
    link(g1 + 1, g3 + 2); link(g2 + 1, g1 + 2); link(g3 + 1, g2 + 2);
  }

  // insert p: walk from a real triangle that has near_pt (an inserted vertex), split what the
  // walk ends in, restore the Delaunay property.  false when the walk does not end.
  bool insert(int32_t p, int64_t near_pt, int64_t max_steps) {
    const int32_t start = real_triangle_of(near_pt);
    int32_t s = start, from = -1;
    int64_t steps = 0;
    for (;;) {
      if (++steps > max_steps) return false;
      // s is real here
      int32_t cross = -1;
      int zeros = 0, zero_edge = -1;
      for (int e = 0; e < 3; ++e) {
        const int32_t h = s + e;
        if (h == from) continue;
        const int o = orient(tri[h], tri[next(h)], p);
        if (o < 0) { cross = h; break; }
        if (o == 0) { ++zeros; zero_edge = h; }
      }
      if (cross >= 0) {
        const int32_t t = half[cross];
        const int32_t ts = t - t % 3;
        if (!real(ts)) {                // left the hull through a visible edge
          split3(ts, p);
          break;
        }
        from = t;
        s = ts;
        continue;
      }
      if (zeros == 0) {
        split3(s, p);
      } else if (zeros == 1) {
        split4(zero_edge, p);
      } else {
        // coincides with a vertex: a repeated point, not a vertex of the triangulation
      }
      break;
    }
    relax();
    return true;
  }

  // hull: collinear triples make Qhull's answer non-unique (counted into `ties`)
  int hull_ties() {
    hull_next.assign(n, -1);
    int32_t start = -1;
    for (int64_t t = 0; t < ntri; ++t) {
      const int32_t s = (int32_t)(3 * t);
      if (real(s)) continue;
      const int32_t g = tri[s] == G ? s : (tri[s + 1] == G ? s + 1 : s + 2);
      const int32_t v = tri[next(g)], u = tri[prev(g)];   // real edge v -> u, hull edge u -> v
      hull_next[u] = v;
      start = u;
    }
    if (start < 0) return -1;
    int32_t p = start;
    int64_t nh = 0;
    do {
      const int32_t q = hull_next[p], r = hull_next[q];
      if (orient(p, q, r) == 0) ++ties;
      p = q;
      if (++nh > n) return -1;
    } while (p != start);
    return 0;
  }

  void drop_ghosts() {
    for (int64_t t = 0; t < ntri; ++t) {
      const int32_t s = (int32_t)(3 * t);
      if (real(s)) continue;
      for (int e = 0; e < 3; ++e) {
        if (half[s + e] >= 0) half[half[s + e]] = -1;
        half[s + e] = -1;
      }
      tri[s] = tri[s + 1] = tri[s + 2] = -1;
    }
  }

  // 0 on success, -1 when this builder does not apply (caller falls back)
  int run() {
    if (rows < 2 || cols < 2) return -1;
    n = rows * cols;
    if (n > (int64_t)0x1fffffff) return -1;
    ntri = flips = ties = 0;
    stack.clear();
    for (int64_t i = 0; i < n; ++i)
      if (!(std::fabs(x[i]) <= 1e300) || !(std::fabs(y[i]) <= 1e300)) return -1;
    G = (int32_t)n;
    vtri.assign(n + 1, -1);
    tri.assign(3 * (2 * n + 8), -1);
    half.assign(3 * (2 * n + 8), -1);
    stack.reserve(256);
    // Insertion order.  Plain row-by-row order is a trap: a new scan line lies just
    // outside the hull, whose nearly straight previous line it sees end to end, so
    // every point fans onto ~cols/2 hull edges that later points flip away again
    // (measured: 24 flips per point); so is inserting the outline first (a slightly
    // curved chain inserted end to end costs a quadratic number of flips).  Instead
    // the lattice is inserted coarse to fine: every 2^L-th point of every 2^L-th
    // line plus the last line / last point of each line, so that every level spans
    // the whole footprint; L descending, boustrophedon inside a level.  That behaves
    // like a randomised order (a few flips per point) yet keeps consecutive points
    // close, so the location walk from the previous point stays short.
    // The levels refine the lattice ISOTROPICALLY IN THE PLANE, not in index space: OMI pixels
    // are 13 km apart along the track and 24-150 km across it (4.7 : 1 in lon/lat), so a level
    // steps 4 x further in the row index than in the column index (ai = 2); TROPOMI's 1.7 : 1
    // gives ai = 1.  With square-ish cells at every level the intermediate triangulations look
    // like the final one and a new point costs 2.7 flips instead of 5.1 (measured on an OMI
    // granule: 502 k -> 267 k flips, insertion 56 -> 34 ms).
    int ai = 0, aj = 0;
    {
      std::vector<double> dr, dc;
      const int64_t sr = std::max<int64_t>(1, rows / 64), sc = std::max<int64_t>(1, cols / 64);
      for (int64_t i = 0; i + 1 < rows; i += sr)
        for (int64_t j = 0; j + 1 < cols; j += sc) {
          const int64_t v = i * cols + j;
          dr.push_back(std::hypot(x[v + cols] - x[v], y[v + cols] - y[v]));
          dc.push_back(std::hypot(x[v + 1] - x[v], y[v + 1] - y[v]));
        }
      if (!dr.empty()) {
        std::nth_element(dr.begin(), dr.begin() + dr.size() / 2, dr.end());
        std::nth_element(dc.begin(), dc.begin() + dc.size() / 2, dc.end());
        const double a = dr[dr.size() / 2], b = dc[dc.size() / 2];   // medians: row / column spacing
        if (a > 0.0 && b > 0.0) {
          const int k = (int)std::lround(std::log2(b / a));
          ai = std::min(std::max(k, 0), 3);
          aj = std::min(std::max(-k, 0), 3);
        }
      }
    }
    ord.clear();                          // insertion order
    lev.clear();                          // level index (0 = finest) each point enters at
    ord.reserve(n);
    lev.reserve(n);
    {
      seen.assign(n, 0);
      int cur_level = 0;
      auto push = [&](int64_t i, int64_t j) {
        const int64_t v = i * cols + j;
        if (!seen[v]) { seen[v] = 1; ord.push_back((int32_t)v); lev.push_back((uint8_t)cur_level); }
      };
      int top = 0;
      while ((int64_t(2) << top) < std::max(rows, cols)) ++top;
      std::vector<int64_t> ri, cj;
      const int lowest = -std::max(ai, aj);
      for (int L = top; L >= lowest; --L) {
        cur_level = L - lowest;
        const int64_t step = int64_t(1) << std::max(L + ai, 0);
        const int64_t stepj = int64_t(1) << std::max(L + aj, 0);
        ri.clear();
        cj.clear();
        for (int64_t i = 0; i < rows; i += step) ri.push_back(i);
        if (ri.back() != rows - 1) ri.push_back(rows - 1);
        for (int64_t j = 0; j < cols; j += stepj) cj.push_back(j);
        if (cj.back() != cols - 1) cj.push_back(cols - 1);
        bool backward = false;
        for (int64_t i : ri) {
          if (backward) { for (size_t c = cj.size(); c-- > 0;) push(i, cj[c]); }
          else { for (int64_t j : cj) push(i, j); }
          backward = !backward;
        }
      }
    }
    auto order = [&](int64_t k) { return ord[k]; };
    int32_t p0 = order(0), p1 = order(1), p2 = -1;
    if (x[p0] == x[p1] && y[p0] == y[p1]) return -1;
    int64_t k2 = 2;
    for (; k2 < n; ++k2) {
      const int s = orient(p0, p1, order(k2));
      if (s != 0) {
        p2 = order(k2);
        if (s < 0) std::swap(p0, p1);
        break;
      }
    }
    if (p2 < 0) return -1;
    begin(p0, p1, p2);
    // The walk to a new point starts at a triangle that HAS the nearest of: the previous point,
    // and the (up to four) corners of the coarser lattice cell around the new point, all
    // inserted at earlier levels (vtri is kept exact through every split and flip: 2.7 triangles
    // visited per point instead of 5.7 with a triangle remembered from insertion time).
    // Without the corners every scan line that crosses the date line costs a walk across the
    // whole plane in each direction of the boustrophedon (measured: 2x the build time for an
    // orbit along the date line).
    int32_t prev_pt = p2;
    const int64_t max_steps = 8 * n + 64;
    for (int64_t k = 2; k < n; ++k) {
      if (k == k2) continue;
      const int32_t p = order(k);
      int64_t near_pt = prev_pt;
      {
        const double px = x[p], py = y[p];
        double best = (x[prev_pt] - px) * (x[prev_pt] - px) + (y[prev_pt] - py) * (y[prev_pt] - py);
        const int Lr = (int)lev[k] - std::max(ai, aj);
        const int64_t s2 = int64_t(2) << std::max(Lr + ai, 0), s2j = int64_t(2) << std::max(Lr + aj, 0);
        const int64_t i = p / cols, j = p % cols;
        const int64_t i0 = (i / s2) * s2, j0 = (j / s2j) * s2j;
        const int64_t i1 = std::min(i0 + s2, rows - 1), j1 = std::min(j0 + s2j, cols - 1);
        const int64_t cand[4] = {i0 * cols + j0, i0 * cols + j1, i1 * cols + j0, i1 * cols + j1};
        for (int c = 0; c < 4; ++c) {
          if (vtri[cand[c]] < 0) continue;
          const double d = (x[cand[c]] - px) * (x[cand[c]] - px) + (y[cand[c]] - py) * (y[cand[c]] - py);
          if (d < best) { best = d; near_pt = cand[c]; }
        }
      }
      if (!insert(p, near_pt, max_steps)) return -1;
      if (vtri[p] >= 0) prev_pt = p;      // (a repeated point is not a vertex)
    }
    if (hull_ties() != 0) return -1;
    drop_ghosts();
    return 0;
  }

  // a reused builder keeps its arrays between calls unless they are large (TROPOMI-scale
  // granules: 130 MB per thread would stay resident)
  void release_if_large() {
    if (tri.capacity() > (size_t)8 << 20) {
      std::vector<int32_t>().swap(tri);
      std::vector<int32_t>().swap(half);
      std::vector<int32_t>().swap(vtri);
      std::vector<int32_t>().swap(ord);
      std::vector<int32_t>().swap(hull_next);
      std::vector<uint8_t>().swap(lev);
      std::vector<uint8_t>().swap(seen);
    }
  }

  int64_t emit(int32_t* out, int64_t capacity) const {
    int64_t m = 0;
    for (int64_t t = 0; t < ntri; ++t) {
      if (tri[3 * t] < 0) continue;
      if (m >= capacity) return -1;
      out[3 * m] = tri[3 * t]; out[3 * m + 1] = tri[3 * t + 1]; out[3 * m + 2] = tri[3 * t + 2];
      ++m;
    }
    return m;
  }
};
