// swb_align.cu -- host-side traceback for the few hits that get an alignment printed.
//
// Role taken over from the reference (torognes/swipe): align() (align.cc:469-519), which hits_align
// calls for the best `-b` hits only (hits.cc:546-623).  Control-plane work on a handful of sequence
// pairs; the scan over the database is the GPU's job and the end cell normally arrives as a hint
// from swb_search_end.
//
// Method (the published one: Myers & Miller 1988, CABIOS 4:11-17, in the local-alignment framing of
// Huang, Hardison & Miller 1990, CABIOS 6:373-381): find the end cell of the best local alignment,
// find its start cell with a sweep that runs backwards from the end cell, then recover the path of
// the enclosed global alignment in linear space by divide and conquer.
//
// Structure (ours): ONE lattice sweep (Lattice::sweep) serves all of it.  A Frame describes a window
// of the two sequences as seen from one of its corners (so "forward" and "backward" are the same
// loop), a Rim says what lies left of the window's first column on every row (nothing, for the
// local and the anchored sweep; a gap that started at the corner, for the global one), and a visitor
// sees every cell (to track a maximum, or to stop at the first cell that reaches a score).  The
// divide and conquer runs off an explicit work list, not recursion, and writes run-length operations
// "M<n>" (aligned pair), "I<n>" (subject symbols against a gap), "D<n>" (query symbols against a gap).
//
// What it must share with the reference are the RESULTS, ties included, because a drop-in has to
// print the same alignment among the co-optimal ones.  The rules that decide them, as observed from
// the reference's output (tests/test_align.py compares with it live and through golden vectors):
//   - the end cell is the first strict maximum in query-major order (align.cc:71-104);
//   - the start cell is the first cell of the backward sweep, again query-major from the end cell,
//     whose score strictly exceeds every earlier one and reaches the alignment's; cells outside the
//     paths through the end cell start from -1, not from minus infinity (align.cc:111-154);
//   - a divide step joins the two halves at the first subject position with the strictly best
//     "both halves end on a column" sum, unless a join inside a query-side gap is at least as good,
//     in which case the LAST such position wins (align.cc:418-444);
//   - one query symbol against N subject symbols: the gap-first layout wins on equal boundary costs,
//     otherwise the first subject position with the best substitution (align.cc:260-326).
#include "../../include/swipe_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{

typedef long long i64;

struct Model
{
  const int64_t *matrix;     // [(subject symbol << 5) + query symbol]
  i64 open, ext;             // a gap of k positions costs open + k * ext
  i64 pair(unsigned qsym, unsigned dsym) const { return matrix[((i64)(dsym & 31) << 5) + (qsym & 31)]; }
};

// A window of query x subject seen from one corner: view position (i, j), 0-based, is query symbol
// q0 + step * i against subject symbol d0 + step * j.  step = +1 looks from the top-left corner,
// step = -1 from the bottom-right one.
struct Frame
{
  const uint8_t *query, *subject;
  i64 q0, d0;
  int step;
  i64 rows, cols;
  unsigned qsym(i64 i) const { return query[q0 + step * i]; }
  unsigned dsym(i64 j) const { return subject[d0 + step * j]; }
};

// What a row finds to the left of the window's first column.
struct Rim
{
  i64 h, f, diag;            // H and F of the virtual column -1, and H of that column one row up
};

enum Kind
{
  LOCAL,                     // H floored at 0, free start anywhere
  ANCHORED,                  // must start in view cell (0, 0); everything else starts from -1
  GLOBAL                     // must start at the corner; the rim is a query-side gap hanging off it
};

struct Lattice
{
  std::vector<i64> H, G;     // per view column: best score, best score ending in a subject-side gap

  explicit Lattice(i64 cols) : H((size_t)cols + 1), G((size_t)cols + 1) {}

  // Sweeps the rows of the frame in order.  visit(i, j, h) sees every cell's H and returns true to
  // stop the sweep there.  edge (GLOBAL): cost of opening the rim gap, 0 when the neighbouring piece
  // of the alignment already ends in one.  Returns true when a visitor stopped it.
  template <Kind K, class Visit> bool sweep(const Model &m, const Frame &fr, i64 edge, Visit visit)
  {
    const i64 q = m.open, r = m.ext;
    for (i64 j = 0; j < fr.cols; j++)
    {
      const i64 reach = -q - (j + 1) * r;              // GLOBAL: the corner row is one subject-side gap
      H[(size_t)j] = K == LOCAL ? 0 : (K == ANCHORED ? -1 : reach);
      G[(size_t)j] = K == LOCAL ? -q : (K == ANCHORED ? -1 : reach - q);
    }
    i64 above = 0;                                     // GLOBAL: the rim's H one row up
    for (i64 i = 0; i < fr.rows; i++)
    {
      Rim rim;
      if (K == LOCAL) rim = Rim{0, -q, 0};
      else if (K == ANCHORED) rim = Rim{-1, -1, i == 0 ? 0 : -1};
      else
      {
        const i64 here = -edge - (i + 1) * r;
        rim = Rim{here, here - q, above};
        above = here;
      }
      i64 h = rim.h, f = rim.f, diag = rim.diag;
      const unsigned qs = fr.qsym(i);
      for (i64 j = 0; j < fr.cols; j++)
      {
        i64 &hj = H[(size_t)j], &gj = G[(size_t)j];
        const i64 f_open = h - q, g_open = hj - q;
        f = (f > f_open ? f : f_open) - r;
        gj = (gj > g_open ? gj : g_open) - r;
        h = diag + m.pair(qs, fr.dsym(j));
        if (K == LOCAL && h < 0) h = 0;
        if (f > h) h = f;
        if (gj > h) h = gj;
        diag = hj;
        hj = h;
        if (visit(i, j, h)) return true;
      }
      rim_h = rim.h;
    }
    return false;
  }
  i64 rim_h = 0;             // H of the rim on the last row swept (GLOBAL: the "no subject symbol used" score)
};

// Run-length operation string ("M12D2I3").
class OpWriter
{
  std::string text_;
  char op_ = 0;
  i64 run_ = 0;

 public:
  void put(char op, i64 n)
  {
    if (n <= 0) return;
    if (op != op_) { close(); op_ = op; }
    run_ += n;
  }
  void close()
  {
    if (run_ > 0)
    {
      char buf[32];
      snprintf(buf, sizeof buf, "%c%lld", op_, run_);
      text_ += buf;
    }
    run_ = 0;
  }
  const std::string &text() const { return text_; }
};

// One entry of the divide-and-conquer work list: either a piece still to be aligned globally --
// query[qa, qa + M) against subject[da, da + N), where lead / trail is what opening a query-side gap
// costs at its left / right end -- or operations that are already known.
struct Piece
{
  i64 qa, da, M, N, lead, trail;
  char op;                   // != 0: just write `count` of this operation
  i64 count;
};

class PathFinder
{
  const Model &m_;
  const uint8_t *query_, *subject_;
  Lattice fwd_, bwd_;
  OpWriter &out_;
  std::vector<Piece> todo_;  // a stack: pieces are pushed right to left so that they pop left to right

  void single_query_symbol(const Piece &p)
  {
    const i64 q = m_.open, r = m_.ext;
    // candidates, in the order that decides ties: the symbol deleted at the cheaper end (the left one
    // when both cost the same), then matched to subject position 0, 1, ... (first best wins)
    const bool left = p.lead <= p.trail;
    i64 best = -(left ? p.lead : p.trail) - q - (1 + p.N) * r;
    i64 where = left ? -1 : p.N;
    for (i64 j = 0; j < p.N; j++)
    {
      i64 s = m_.pair(query_[p.qa], subject_[p.da + j]) - (p.N - 1) * r;
      s -= (j > 0 ? q : 0) + (j < p.N - 1 ? q : 0);
      if (s > best) { best = s; where = j; }
    }
    if (where < 0) { out_.put('D', 1); out_.put('I', p.N); }
    else if (where == p.N) { out_.put('I', p.N); out_.put('D', 1); }
    else { out_.put('I', where); out_.put('M', 1); out_.put('I', p.N - 1 - where); }
  }

  void divide(const Piece &p)
  {
    const i64 q = m_.open;
    const i64 upper = p.M / 2, lower = p.M - upper;
    auto never = [](i64, i64, i64) { return false; };
    const Frame top{query_, subject_, p.qa, p.da, +1, upper, p.N};
    const Frame bottom{query_, subject_, p.qa + p.M - 1, p.da + p.N - 1, -1, lower, p.N};
    fwd_.sweep<GLOBAL>(m_, top, p.lead, never);
    bwd_.sweep<GLOBAL>(m_, bottom, p.trail, never);
    // score of the upper half against the first k subject symbols (k = 0: the rim), and likewise
    // for the lower half against the last N - k; the G variants end / begin inside a query-side gap
    auto upH = [&](i64 k) { return k == 0 ? fwd_.rim_h : fwd_.H[(size_t)k - 1]; };
    auto upG = [&](i64 k) { return k == 0 ? fwd_.rim_h : fwd_.G[(size_t)k - 1]; };
    auto loH = [&](i64 k) { return k == p.N ? bwd_.rim_h : bwd_.H[(size_t)(p.N - k) - 1]; };
    auto loG = [&](i64 k) { return k == p.N ? bwd_.rim_h : bwd_.G[(size_t)(p.N - k) - 1]; };
    i64 cut = 0, best = upH(0) + loH(0);
    bool through_gap = false;
    for (i64 k = 1; k <= p.N; k++)
      if (upH(k) + loH(k) > best) { best = upH(k) + loH(k); cut = k; }
    for (i64 k = 0; k <= p.N; k++)
      if (upG(k) + loG(k) + q >= best) { best = upG(k) + loG(k) + q; cut = k; through_gap = true; }
    if (through_gap)
    {
      // the rows around the cut are both inside one query-side gap: take them out of the halves
      todo_.push_back(Piece{p.qa + upper + 1, p.da + cut, lower - 1, p.N - cut, 0, p.trail, 0, 0});
      todo_.push_back(Piece{0, 0, 0, 0, 0, 0, 'D', 2});
      todo_.push_back(Piece{p.qa, p.da, upper - 1, cut, p.lead, 0, 0, 0});
    }
    else
    {
      todo_.push_back(Piece{p.qa + upper, p.da + cut, lower, p.N - cut, q, p.trail, 0, 0});
      todo_.push_back(Piece{p.qa, p.da, upper, cut, p.lead, q, 0, 0});
    }
  }

 public:
  PathFinder(const Model &m, const uint8_t *query, const uint8_t *subject, i64 cols, OpWriter &out)
      : m_(m), query_(query), subject_(subject), fwd_(cols), bwd_(cols), out_(out) {}

  void run(i64 qa, i64 da, i64 M, i64 N)
  {
    todo_.push_back(Piece{qa, da, M, N, m_.open, m_.open, 0, 0});
    while (!todo_.empty())
    {
      const Piece p = todo_.back();
      todo_.pop_back();
      if (p.op) out_.put(p.op, p.count);
      else if (p.N == 0) out_.put('D', p.M);
      else if (p.M == 0) out_.put('I', p.N);
      else if (p.M == 1) single_query_symbol(p);
      else divide(p);
    }
  }
};

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
  const Model model{matrix, (i64)gap_open, (i64)gap_extend};
  i64 total = *score, qe = *q_end, de = *d_end;
  if (total != 0)
  {
    if (qe < 0 || qe >= qlen || de < 0 || de >= dlen) return SWB_ERR_ARG;   // a hint must name a cell
  }
  else
  {
    // no hint: the end cell is the first cell, query-major, holding the overall maximum
    qe = de = 0;
    Lattice all(dlen);
    const Frame whole{query, subject, 0, 0, +1, qlen, dlen};
    all.sweep<LOCAL>(model, whole, 0, [&](i64 i, i64 j, i64 h) {
      if (h > total) { total = h; qe = i; de = j; }
      return false;
    });
  }
  if (qlen == 0 || dlen == 0) return SWB_ERR_INTERNAL;
  // the start cell: backwards from the end cell until a new running best reaches the score
  i64 qs = 0, ds = 0, running = 0;
  {
    Lattice back(de + 1);
    const Frame tail{query, subject, qe, de, -1, qe + 1, de + 1};
    const bool found = back.sweep<ANCHORED>(model, tail, 0, [&](i64 i, i64 j, i64 h) {
      if (h <= running) return false;
      running = h;
      qs = qe - i;
      ds = de - j;
      return running >= total;
    });
    if (!found) return SWB_ERR_INTERNAL;     // the reference: fatal("Internal error in align function.")
  }
  OpWriter out;
  PathFinder path(model, query, subject, de - ds + 1, out);
  path.run(qs, ds, qe - qs + 1, de - ds + 1);
  out.close();
  *q_start = qs; *d_start = ds; *q_end = qe; *d_end = de; *score = total;
  if (ops_len) *ops_len = (int64_t)out.text().size();
  if ((int64_t)out.text().size() + 1 > ops_cap) return SWB_ERR_RANGE;
  memcpy(ops, out.text().c_str(), out.text().size() + 1);
  return SWB_OK;
}
