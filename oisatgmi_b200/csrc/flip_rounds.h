// Lawson's edge flips as ROUNDS of independent work: the per-edge steps of k12_flip.cu, written
// once for the device kernels and for the host loop that replays them (oisat_h_flip_rounds,
// the CPU tests' stand-in for the kernels: same code, threads run one after the other).
//
// A round r has two steps, each a loop of independent work items:
//   mark   one item per triangle that changed in round r-1 (its list; in round 0 every
//          triangle): its three edges are tested
//          with the in-circle predicate; an edge that must go claims the four triangles its
//          flip touches -- its own two and the two neighbours whose twin pointers move -- with
//          an atomic maximum of (round, priority of the edge): the highest priority wins a
//          contested triangle;
//   apply  one item per edge marked in this round (its list): an edge that owns all four of its triangles is flipped; the two triangles it
//          rebuilt are stamped with the round (and so is the triangle of a marked edge that
//          lost a claim: it has to be looked at again).
// The marked edge of highest priority always wins everything it claims, so every round with marked edges
// flips at least one; a round without flips is the end: every edge is locally Delaunay.
//
// The predicate is the floating-point filter only.  Where the filter cannot certify the sign
// the edge stays (0 is "do not flip") and the final check counts such edges: the caller then
// takes the exact host builder for that granule (they are the near-cocircular quadruples the
// near-tie scan would send to Qhull anyway).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define OISAT_FLIP_HD __host__ __device__ __forceinline__
#else
#define OISAT_FLIP_HD inline
#endif

namespace oisat_flip {

OISAT_FLIP_HD int nxt(int e) { return e % 3 == 2 ? e - 2 : e + 1; }
OISAT_FLIP_HD int prv(int e) { return e % 3 == 0 ? e + 2 : e - 1; }

// +1: d strictly inside the circle through a, b, c (counter-clockwise); -1: strictly outside;
// 0: within the rounding error bound of the determinant (Shewchuk's stage-A bound)
OISAT_FLIP_HD int incircle_filter(double ax, double ay, double bx, double by, double cx, double cy,
                                  double dx, double dy) {
  const double adx = ax - dx, ady = ay - dy, bdx = bx - dx, bdy = by - dy, cdx = cx - dx,
               cdy = cy - dy;
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, alift = adx * adx + ady * ady;
  const double cdxady = cdx * ady, adxcdy = adx * cdy, blift = bdx * bdx + bdy * bdy;
  const double adxbdy = adx * bdy, bdxady = bdx * ady, clift = cdx * cdx + cdy * cdy;
  const double det = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) +
                     clift * (adxbdy - bdxady);
  const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * alift +
                           (fabs(cdxady) + fabs(adxcdy)) * blift +
                           (fabs(adxbdy) + fabs(bdxady)) * clift;
  const double bound = 1.1102230246251577e-15 * permanent;   // (10 + 96 eps) eps, eps = 2^-53
  if (det > bound) return 1;
  if (-det > bound) return -1;
  return 0;
}

// Priority of edge a in a round: a bijection of the 32-bit edge index (so no two edges share a
// key), scrambled.  With the index itself as the priority a row of quads whose diagonals all
// have to go is flipped one quad per round -- edge k+1 loses to edge k even when edge k itself
// lost to k-1 (426 rounds for an OMI granule); scrambled, the winners are the local maxima of
// a random field and the rounds follow the depth of the flip dependencies (97).
OISAT_FLIP_HD unsigned int scramble(unsigned int a, unsigned int round) {
  unsigned int h = (a + round * 0x632be5abu) * 0x9e3779b1u;
  h ^= h >> 15;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

OISAT_FLIP_HD unsigned long long claim_key(int round, int a) {
  return ((unsigned long long)(unsigned)(round + 1) << 32) |
         (unsigned long long)scramble((unsigned)a, (unsigned)round);
}

struct Mesh {                   // one granule; every index in these arrays is local to it
  int32_t* tri;                 // 3 per triangle, counter-clockwise
  int32_t* half;                // twin half-edge, -1 on the hull
  int64_t n_half;               // 3 * triangles
  int32_t* stamp;               // per triangle: last round that changed it (-1 at the start)
  int32_t* cand;                // per half-edge: round + 1 in which it was last marked
  unsigned long long* owner;    // per triangle: highest claim
  int32_t tri_base;             // number of its first triangle in the batch (the lists' numbering)
};

struct Lists {                  // work lists of a batch of meshes that go through the rounds together
  int32_t* tri_list[2];         // triangles to look at in round r: tri_list[r & 1]
  int32_t* edge_list;           // edges marked in the current round (at most half of the half-edges)
  unsigned int* n_listed;       // per round r: triangles listed FOR round r + 1
  unsigned int* n_marked;       // per round r: edges marked
};

// Ops: unsigned long long max(unsigned long long*, v); unsigned add(unsigned*, v) (returns the
// old value); int32_t exch(int32_t*, v) (returns the old value) -- atomics on the device
// Coords: double operator()(int vertex, int axis)

// The test of one edge in three steps, so that a device thread can have the loads of several
// edges in flight before it decides any of them (a step is a chain of dependent loads:
// twin -> vertices -> coordinates -> claims).
struct Quad {
  int32_t ia, ib;               // the edge's half-edges, ia < ib; ia < 0: nothing to test
  int32_t pr, pl, p0, p1;       // ia runs pr -> pl; p0 is opposite in ia's triangle, p1 in ib's
};

// edge e of a listed triangle (or any half-edge in round 0): who tests it, and its vertices
OISAT_FLIP_HD Quad mark_prepare(const Mesh& m, int32_t e, int round) {
  Quad q;
  q.ia = -1;
  const int32_t b = m.half[e];
  if (b < 0) return q;
  // an edge between two listed triangles is the lower half-edge's (everything is listed in round 0)
  if (b < e && (round == 0 || m.stamp[b / 3] == round - 1)) return q;
  q.ia = e < b ? e : b;
  q.ib = e < b ? b : e;
  q.pr = m.tri[q.ia];
  q.pl = m.tri[nxt(q.ia)];
  q.p0 = m.tri[prv(q.ia)];
  q.p1 = m.tri[prv(q.ib)];
  return q;
}

template <class Coords>
OISAT_FLIP_HD bool mark_decide(const Quad& q, const Coords& P) {
  return incircle_filter(P(q.pr, 0), P(q.pr, 1), P(q.pl, 0), P(q.pl, 1), P(q.p0, 0), P(q.p0, 1),
                         P(q.p1, 0), P(q.p1, 1)) > 0;
}

template <class Ops>
OISAT_FLIP_HD void mark_claim(const Mesh& m, const Lists& l, const Quad& q, int round, const Ops& ops) {
  const unsigned long long key = claim_key(round, q.ia);
  ops.max(&m.owner[q.ia / 3], key);
  ops.max(&m.owner[q.ib / 3], key);
  const int32_t n1 = m.half[prv(q.ib)], n2 = m.half[prv(q.ia)];
  if (n1 >= 0) ops.max(&m.owner[n1 / 3], key);
  if (n2 >= 0) ops.max(&m.owner[n2 / 3], key);
  m.cand[q.ia] = round + 1;
  l.edge_list[ops.add(&l.n_marked[round], 1u)] = q.ia + 3 * m.tri_base;
}

template <class Coords, class Ops>
OISAT_FLIP_HD void mark_edge(const Mesh& m, const Lists& l, int32_t e, int round, const Coords& P,
                             const Ops& ops) {
  const Quad q = mark_prepare(m, e, round);
  if (q.ia >= 0 && mark_decide(q, P)) mark_claim(m, l, q, round, ops);
}

template <class Ops>
OISAT_FLIP_HD void touch(const Mesh& m, const Lists& l, int t, int round, const Ops& ops) {
  if (ops.exch(&m.stamp[t], round) != round)
    ((round & 1) ? l.tri_list[0] : l.tri_list[1])[ops.add(&l.n_listed[round], 1u)] = t + m.tri_base;
}

// returns 1 when the marked edge ia (local number) was flipped.  Nothing of a triangle is read
// before its ownership is established: the owner of a triangle may be rewriting it in this
// very step (half[a] can even become -1 under a loser's feet).  A marked edge that loses
// lists its own triangle, so that it is tested again in the next round although nothing
// around it may have changed.
template <class Ops>
OISAT_FLIP_HD int apply_edge(const Mesh& m, const Lists& l, int32_t ia, int round, const Ops& ops) {
  const unsigned long long key = claim_key(round, ia);
  const int ta = ia / 3;
  if (m.owner[ta] != key) { touch(m, l, ta, round, ops); return 0; }
  const int32_t b = m.half[ia];
  if (m.owner[b / 3] != key) { touch(m, l, ta, round, ops); return 0; }
  const int ar = prv(ia), bl = prv(b);
  const int32_t hbl = m.half[bl], har = m.half[ar];
  if ((hbl >= 0 && m.owner[hbl / 3] != key) || (har >= 0 && m.owner[har / 3] != key)) {
    touch(m, l, ta, round, ops);
    return 0;
  }
  const int32_t p0 = m.tri[ar], p1 = m.tri[bl];
  m.tri[ia] = p1;
  m.tri[b] = p0;
  m.half[ia] = hbl;
  if (hbl >= 0) m.half[hbl] = ia;
  m.half[b] = har;
  if (har >= 0) m.half[har] = b;
  m.half[ar] = bl;
  m.half[bl] = ar;
  touch(m, l, ta, round, ops);
  touch(m, l, b / 3, round, ops);
  return 1;
}

// final check of half-edge a (the lower twin tests): bit 0 = certainly not Delaunay, bit 1 =
// the filter cannot tell
template <class Coords>
OISAT_FLIP_HD int check_edge(const int32_t* tri, const int32_t* half, int64_t a, const Coords& P) {
  const int32_t b = half[a];
  if (b < a) return 0;
  const int ia = (int)a;
  const int32_t pr = tri[ia], pl = tri[nxt(ia)], p0 = tri[prv(ia)], p1 = tri[prv(b)];
  const int s = incircle_filter(P(pr, 0), P(pr, 1), P(pl, 0), P(pl, 1), P(p0, 0), P(p0, 1), P(p1, 0), P(p1, 1));
  return s > 0 ? 1 : (s == 0 ? 2 : 0);
}

}  // namespace oisat_flip
