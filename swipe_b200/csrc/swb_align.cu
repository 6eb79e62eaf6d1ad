// swb_align.cu -- host-side traceback for the hits that get an alignment printed.
//
// Takes over from the reference (torognes/swipe) align() (align.cc:469-519), which hits_align
// calls for the best `-b` hits only (hits.cc:546-623): locate the local alignment's start with a
// reverse pass from its end cell, then recover the path of the enclosed GLOBAL alignment in linear
// space (Myers & Miller 1988, CABIOS 4:11-17; Huang, Hardison & Miller 1990, CABIOS 6:373-381) and
// report it as run-length operations "M<n>" (aligned pair), "I<n>" (subject symbols against a gap)
// and "D<n>" (query symbols against a gap).
//
// This is control-plane work on a handful of sequences (the score-only scan over the database is
// the GPU's job; the end cell of each hit comes from swb_search_end).  It is written independently
// from the published algorithm; what it shares with the reference are the tie-breaking rules that
// decide WHICH optimal alignment is reported, because a drop-in has to print the same one:
//   - the end cell is the first strict maximum in query-major order (align.cc:71-104),
//   - the start cell is the first cell of the reverse sweep whose score reaches the alignment's
//     (align.cc:117-157),
//   - the midpoint of a divide step prefers, in subject order, the first strictly better "pass
//     through a match column" join and then the LAST at-least-as-good "inside a query gap" join
//     (align.cc:418-444),
//   - a single query symbol against N subject symbols prefers the gap-first layout on equal boundary
//     costs and otherwise the first best substitution position (align.cc:260-326).
#include "../../include/swipe_b200.h"

#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{

typedef long long i64;

struct Scorer
{
  const int64_t *m;          // [(subject << 5) + query]
  i64 open, ext;             // gap of length k costs open + k * ext
  const uint8_t *a, *b;      // a = query, b = subject
  inline i64 sub(i64 i, i64 j) const { return m[((i64)(b[j] & 31) << 5) + (a[i] & 31)]; }
};

// Run-length op string builder ("M12D2I3").
struct Ops
{
  std::string text;
  char cur = 0;
  i64 run = 0;
  void flush()
  {
    if (run > 0)
    {
      char buf[32];
      snprintf(buf, sizeof buf, "%c%lld", cur, run);
      text += buf;
    }
    run = 0;
  }
  void add(char op, i64 n)
  {
    if (n <= 0) return;
    if (op != cur) { flush(); cur = op; }
    run += n;
  }
};

// One half of a divide step: global affine-gap DP of `rows` query symbols against the n subject
// symbols of the window, walking away from the window's corner at (qa, ba) in direction dir (+1:
// forward from the top-left corner, -1: backward from the bottom-right one).  edge = what opening a
// query-side gap costs at this corner (0 when the caller already has one open there).  On return
// H[j] = best score of the rows against the first j subject symbols (in walking order) and
// G[j] = the same ending inside a gap in the subject (a "D" run), with G[0] = H[0].
void half_sweep(const Scorer &S, i64 qa, i64 ba, int dir, i64 rows, i64 n, i64 edge,
                std::vector<i64> &H, std::vector<i64> &G)
{
  const i64 q = S.open, r = S.ext;
  H[0] = 0;
  i64 t = -q;
  for (i64 j = 1; j <= n; j++)
  {
    t -= r;
    H[j] = t;
    G[j] = t - q;
  }
  t = -edge;
  for (i64 i = 1; i <= rows; i++)
  {
    i64 diag = H[0];
    t -= r;
    i64 h = t;
    H[0] = t;
    i64 f = t - q;
    const i64 qi = qa + dir * (i - 1);
    for (i64 j = 1; j <= n; j++)
    {
      const i64 fo = h - q;
      f = (f > fo ? f : fo) - r;
      const i64 go = H[j] - q;
      G[j] = (G[j] > go ? G[j] : go) - r;
      h = diag + S.sub(qi, ba + dir * (j - 1));
      if (f > h) h = f;
      if (G[j] > h) h = G[j];
      diag = H[j];
      H[j] = h;
    }
  }
  G[0] = H[0];
}

struct Tracer
{
  const Scorer &S;
  Ops &out;
  // the four DP rows of a divide step; a level is done with them before it recurses, so one set
  // sized for the outermost window serves every level
  std::vector<i64> H, G, X, Y;
  Tracer(const Scorer &s, Ops &o, i64 n) : S(s), out(o), H((size_t)n + 1), G((size_t)n + 1),
                                           X((size_t)n + 1), Y((size_t)n + 1) {}

  // global alignment of query[qa, qa+M) with subject[ba, ba+N); tb / te = cost of opening a
  // query-side gap at the left / right end (0 if the neighbouring piece ends in one)
  void solve(i64 qa, i64 ba, i64 M, i64 N, i64 tb, i64 te)
  {
    const i64 q = S.open, r = S.ext;
    if (N == 0)
    {
      out.add('D', M);
      return;
    }
    if (M == 0)
    {
      out.add('I', N);
      return;
    }
    if (M == 1)
    {
      // one query symbol: either it is deleted before/after the N inserted symbols, or it is
      // matched to one of them
      i64 best, at;
      if (tb <= te) { best = -tb - (1 + N) * r - q; at = -1; }
      else { best = -q - (1 + N) * r - te; at = N; }
      for (i64 j = 0; j < N; j++)
      {
        i64 sc = S.sub(qa, ba + j) - r * (N - 1);
        if (j > 0) sc -= q;
        if (j < N - 1) sc -= q;
        if (sc > best) { best = sc; at = j; }
      }
      if (at == -1) { out.add('D', 1); out.add('I', N); }
      else if (at == N) { out.add('I', N); out.add('D', 1); }
      else { out.add('I', at); out.add('M', 1); out.add('I', N - 1 - at); }
      return;
    }
    const i64 top = M / 2;
    half_sweep(S, qa, ba, +1, top, N, tb, H, G);
    half_sweep(S, qa + M - 1, ba + N - 1, -1, M - top, N, te, X, Y);
    i64 best = LLONG_MIN, cut = -1;
    bool in_gap = false;
    for (i64 j = 0; j <= N; j++)
    {
      const i64 sc = H[j] + X[N - j];
      if (sc > best) { best = sc; cut = j; in_gap = false; }
    }
    for (i64 j = 0; j <= N; j++)
    {
      const i64 sc = G[j] + Y[N - j] + q;
      if (sc >= best) { best = sc; cut = j; in_gap = true; }
    }
    if (!in_gap)
    {
      solve(qa, ba, top, cut, tb, q);
      solve(qa + top, ba + cut, M - top, N - cut, q, te);
    }
    else
    {
      solve(qa, ba, top - 1, cut, tb, 0);
      out.add('D', 2);
      solve(qa + top + 1, ba + cut, M - top - 1, N - cut, 0, te);
    }
  }
};

// Forward local pass: score and end cell (first strict maximum, query-major).
void find_end(const Scorer &S, i64 M, i64 N, i64 *score, i64 *qe, i64 *de)
{
  const i64 q = S.open, r = S.ext;
  std::vector<i64> H((size_t)N, 0), G((size_t)N, -q);
  i64 best = 0;
  for (i64 i = 0; i < M; i++)
  {
    i64 h = 0, diag = 0, f = -q;
    for (i64 j = 0; j < N; j++)
    {
      const i64 fo = h - q;
      f = (f > fo ? f : fo) - r;
      const i64 go = H[j] - q;
      G[j] = (G[j] > go ? G[j] : go) - r;
      h = diag + S.sub(i, j);
      if (h < 0) h = 0;
      if (f > h) h = f;
      if (G[j] > h) h = G[j];
      diag = H[j];
      H[j] = h;
      if (h > best) { best = h; *qe = i; *de = j; }
    }
  }
  *score = best;
}

// Reverse pass from the end cell: the first cell whose backward score reaches the alignment's.
bool find_start(const Scorer &S, i64 score, i64 qe, i64 de, i64 *qs, i64 *ds)
{
  const i64 q = S.open, r = S.ext;
  std::vector<i64> H((size_t)de + 1, -1), G((size_t)de + 1, -1);
  i64 cost = 0;
  for (i64 i = qe; i >= 0; i--)
  {
    i64 h = -1, f = -1, diag = (i == qe) ? 0 : -1;
    for (i64 j = de; j >= 0; j--)
    {
      const i64 fo = h - q;
      f = (f > fo ? f : fo) - r;
      const i64 go = H[j] - q;
      G[j] = (G[j] > go ? G[j] : go) - r;
      h = diag + S.sub(i, j);
      if (f > h) h = f;
      if (G[j] > h) h = G[j];
      diag = H[j];
      H[j] = h;
      if (h > cost)
      {
        cost = h;
        *qs = i;
        *ds = j;
        if (cost >= score) return true;
      }
    }
  }
  return false;
}

}  // namespace

extern "C" int swb_align(const uint8_t *query, int64_t qlen, const uint8_t *subject, int64_t dlen,
                         const int64_t *matrix, int64_t gap_open, int64_t gap_extend,
                         int64_t *q_start, int64_t *d_start, int64_t *q_end, int64_t *d_end,
                         int64_t *score, char *ops, int64_t ops_cap, int64_t *ops_len)
{
  if (!query || !subject || !matrix || qlen < 0 || dlen < 0 || !q_start || !d_start || !q_end ||
      !d_end || !score || ops_cap < 0 || (ops_cap > 0 && !ops))
    return SWB_ERR_ARG;
  if (gap_open < 0 || gap_extend < 0) return SWB_ERR_ARG;
  Scorer S{matrix, (i64)gap_open, (i64)gap_extend, query, subject};
  i64 sc = *score, qe = *q_end, de = *d_end, qs = 0, ds = 0;
  if (sc != 0)
  {
    if (qe < 0 || qe >= qlen || de < 0 || de >= dlen) return SWB_ERR_ARG;   // a hint must name a cell
  }
  else
  {
    qe = de = 0;
    find_end(S, qlen, dlen, &sc, &qe, &de);
  }
  if (qlen == 0 || dlen == 0 || !find_start(S, sc, qe, de, &qs, &ds))
    return SWB_ERR_INTERNAL;                 // the reference: fatal("Internal error in align function.")
  Ops out;
  Tracer tr(S, out, de - ds + 1);
  tr.solve(qs, ds, qe - qs + 1, de - ds + 1, S.open, S.open);
  out.flush();
  *q_start = qs; *d_start = ds; *q_end = qe; *d_end = de; *score = sc;
  if (ops_len) *ops_len = (int64_t)out.text.size();
  if ((int64_t)out.text.size() + 1 > ops_cap) return SWB_ERR_RANGE;
  memcpy(ops, out.text.c_str(), out.text.size() + 1);
  return SWB_OK;
}
