// swb_scoring.cu -- host side of the scoring system: substitution tables and Karlin-Altschul
// statistics.
//
// Takes over from the reference (torognes/swipe):
//   score_matrix_read / _read_string / _read_file ... matrices.cc:345-591 -> swb_matrix_*
//   stats_getparams / _nt / stats_getprefs ........... stats.cc:44-325     -> swb_stats_params*, _default_gaps
//   BlastComputeLengthAdjustment (NCBI BLAST) ........ blastkar_partial.c:656-748 -> length_adjustment
//   hits_init's search-space and threshold set-up .... hits.cc:283-511     -> swb_stats_init
// The numeric tables live in swb_tables.inc (generated, data only).
#include "../../include/swipe_b200.h"
#include "swb_tables.inc"

#include <strings.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace
{

// NCBIstdaa letter -> code; '*' = 25, '-' = 0 (query.cc:55-75).  -1 = not a symbol.
int aa_code(int c)
{
  static const char sym[] = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ";
  if (c >= 'a' && c <= 'z') c -= 32;
  if (c == '-' || c == '*' || (c >= 'A' && c <= 'Z'))
    for (int i = 0; sym[i]; i++)
      if (sym[i] == c) return i;
  return -1;
}

// Fixed point of  ell = alpha/lambda * (ln K + ln((m - ell)(n - N ell))) + beta  by the bracketed
// iteration NCBI BLAST uses (at most 20 rounds), so that the same integer comes out.
long long length_adjustment(double K, double logK, double a_d_l, double beta, double m, double n,
                            double N)
{
  const double c = n * m - (m > n ? m : n) / K;
  if (c < 0) return 0;
  const double mb = m * N + n;
  double hi = 2 * c / (mb + sqrt(mb * mb - 4 * N * c));
  double lo = 0, next = 0;
  bool converged = false;
  for (int it = 1; it <= 20; it++)
  {
    const double ell = next;
    const double bar = a_d_l * (logK + log((m - ell) * (n - N * ell))) + beta;
    if (bar >= ell)
    {
      lo = ell;
      if (bar - lo <= 1.0) { converged = true; break; }
      if (lo >= hi) break;
    }
    else
      hi = ell;
    if (lo <= bar && bar <= hi) next = bar;
    else next = it == 1 ? hi : (lo + hi) / 2;
  }
  long long adj = (long long)lo;
  if (converged)
  {
    const double up = ceil(lo);
    if (up <= hi && a_d_l * (logK + log((m - up) * (n - N * up))) + beta >= up) adj = (long long)up;
  }
  return adj;
}

// the "sound" alphabet of -p 5 (query.cc:31-49): A-Z = 1..26, a-e = 27..31
int sound_code(int c)
{
  if (c >= 'A' && c <= 'Z') return c - 'A' + 1;
  if (c >= 'a' && c <= 'e') return c - 'a' + 27;
  return -1;
}

// The text format of matrix files (matrices.cc:437-517): '#' comments; a line starting with a
// blank lists the column symbols; every other line is "<row symbol> <score> <score> ...".  The
// first symbol of a data line is the first index, i.e. entry [(row << 5) + column].
int matrix_parse(const char *text, int64_t *m, int (*code_of)(int))
{
  if (!text || !m) return SWB_ERR_ARG;
  for (int i = 0; i < 1024; i++) m[i] = -1;
  int order[4096], nsym = 0;
  const char *s = text;
  while (*s)
  {
    const char *e = strchr(s, '\n');
    std::string line = e ? std::string(s, e - s) : std::string(s);
    s = e ? e + 1 : s + line.size();
    if (line.empty()) continue;
    const char c = line[0];
    if (c == '#' || c == '\n') continue;
    if (c == ' ' || c == '\t')
    {
      for (size_t k = 1; k < line.size(); k++)
        if (!strchr(" \t\n", line[k]) && nsym < 4096) order[nsym++] = code_of((unsigned char)line[k]);
      continue;
    }
    const int a = code_of((unsigned char)c);
    const char *p = line.c_str() + 1;
    for (int i = 0; i < nsym; i++)
    {
      long sc = 0;
      int used = 0;
      if (sscanf(p, "%ld%n", &sc, &used) < 1) return SWB_ERR_ARG;     // "Problem parsing score matrix file."
      const int b = order[i];
      if (a >= 0 && b >= 0 && a < 32 && b < 32) m[(a << 5) + b] = sc;
      p += used;
    }
  }
  return SWB_OK;
}


const SwbKaMatrix *ka_matrix(const char *name)
{
  for (const SwbKaMatrix &k : swb_ka_protein)
    if (strcasecmp(k.name, name) == 0) return &k;
  return nullptr;
}

}  // namespace

extern "C" {

int swb_matrix_builtin(const char *name, int64_t *m)
{
  if (!name || !m) return SWB_ERR_ARG;
  for (const SwbBuiltinMatrix &b : swb_builtin_matrices)
    if (strcasecmp(b.name, name) == 0)
    {
      for (int i = 0; i < 1024; i++) m[i] = b.m[i];
      return SWB_OK;
    }
  return SWB_ERR_ARG;
}

int swb_matrix_parse(const char *text, int64_t *m) { return matrix_parse(text, m, aa_code); }

int swb_matrix_read(const char *name_or_path, int64_t *m)
{
  if (!name_or_path || !m) return SWB_ERR_ARG;
  if (swb_matrix_builtin(name_or_path, m) == SWB_OK) return SWB_OK;
  FILE *f = fopen(name_or_path, "r");
  if (!f) return SWB_ERR_IO;
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  return swb_matrix_parse(text.c_str(), m);
}

// -p 5 ("sound", matrices.cc:284-315, :366, :443): matrix files are read with the sound alphabet; the
// built-in IDENTITY_5_1 scores 5 for equal symbols 1..31 and -1 otherwise.
int swb_matrix_read_sound(const char *name_or_path, int64_t *m)
{
  if (!name_or_path || !m) return SWB_ERR_ARG;
  if (strcasecmp(name_or_path, "identity_5_1") == 0)
  {
    for (int i = 0; i < 1024; i++) m[i] = -1;
    for (int a = 1; a < 32; a++) m[(a << 5) + a] = 5;
    return SWB_OK;
  }
  FILE *f = fopen(name_or_path, "r");
  if (!f) return SWB_ERR_IO;
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  return matrix_parse(text.c_str(), m, sound_code);
}

int swb_matrix_nucleotide(int64_t match, int64_t mismatch, int64_t *m)
{
  if (!m) return SWB_ERR_ARG;
  for (int i = 0; i < 1024; i++) m[i] = -1;
  for (int a = 1; a < 16; a++)
    for (int b = 1; b < 16; b++) m[(a << 5) + b] = a == b ? match : mismatch;   // matrices.cc:533-538
  return SWB_OK;
}

int swb_matrix_limits(const int64_t *m, int64_t *lo, int64_t *hi, int64_t *limit7, int64_t *limit16)
{
  if (!m) return SWB_ERR_ARG;
  int64_t l = 100, h = -100;                          // matrices.cc:560-571
  for (int i = 0; i < 1024; i++)
  {
    if (m[i] < l) l = m[i];
    if (m[i] > h) h = m[i];
  }
  if (lo) *lo = l;
  if (hi) *hi = h;
  if (limit7) *limit7 = 128 - h;
  if (limit16) *limit16 = 65536 - h;
  return SWB_OK;
}

// 1 = parameters found (params: lambda, K, H, alpha, beta), 0 = none for this scoring system
int swb_stats_params(const char *matrix, int64_t gap_open, int64_t gap_extend, double *params)
{
  if (!matrix || !params) return 0;
  const SwbKaMatrix *k = ka_matrix(matrix);
  if (!k) return 0;
  for (int i = 0; i < k->n; i++)
  {
    const SwbKaRow &r = k->rows[i];
    if (fabs(r.open - (double)gap_open) < 0.1 && fabs(r.extend - (double)gap_extend) < 0.1)
    {
      params[0] = r.lambda; params[1] = r.K; params[2] = r.H; params[3] = r.alpha; params[4] = r.beta;
      return 1;
    }
  }
  return 0;
}

int swb_stats_params_nt(int64_t match, int64_t mismatch, int64_t gap_open, int64_t gap_extend,
                        double *params)
{
  if (!params) return 0;
  for (const SwbKaNt &t : swb_ka_nt)
  {
    if (t.reward != match || t.penalty != mismatch) continue;
    if (gap_open >= t.gomax && gap_extend >= t.gemax) gap_open = gap_extend = 0;   // stats.cc:152-156
    for (int i = 0; i < t.n; i++)
    {
      const SwbKaRow &r = t.rows[i];
      if (fabs(r.open - (double)gap_open) < 0.1 && fabs(r.extend - (double)gap_extend) < 0.1)
      {
        params[0] = r.lambda; params[1] = r.K; params[2] = r.H; params[3] = r.alpha; params[4] = r.beta;
        return 1;
      }
    }
    return 0;
  }
  return 0;
}

// the matrix's preferred gap penalties (used when -G / -E are not given, swipe.cc:1098-1115)
int swb_stats_default_gaps(const char *matrix, int64_t *gap_open, int64_t *gap_extend)
{
  if (!matrix || !gap_open || !gap_extend) return 0;
  const SwbKaMatrix *k = ka_matrix(matrix);
  if (!k) return 0;
  for (int i = 0; i < k->n; i++)
    if (k->rows[i].preferred)
    {
      *gap_open = (int64_t)k->rows[i].open;
      *gap_extend = (int64_t)k->rows[i].extend;
      return 1;
    }
  return 0;
}

int64_t swb_stats_length_adjustment(double K, double alpha_d_lambda, double beta, int64_t qlen,
                                    int64_t dblen, int64_t nseq)
{
  // the reference passes query length and sequence count through 32-bit ints (Int4)
  return length_adjustment(K, log(K), alpha_d_lambda, beta, (double)(int)qlen, (double)dblen,
                           (double)(int)nseq);
}

// hits_init (hits.cc:283-511): effective search space and the raw-score window of the hit list.
//   symtype 0..4 as the reference's -p; qlen = nucleotide length for symtype 0/2/4, else residues;
//   symcount / seqcount = the (masked) database totals; effdbsize = -z or 0.
int swb_stats_init(int symtype, const char *matrix, int64_t match, int64_t mismatch, int64_t gap_open,
                   int64_t gap_extend, int64_t qlen, int64_t symcount, int64_t seqcount,
                   int64_t effdbsize, int64_t minscore, int64_t maxscore, double expect,
                   double minexpect, swb_stats *st)
{
  if (!st || symtype < 0 || symtype > 4) return SWB_ERR_ARG;
  memset(st, 0, sizeof *st);
  double p[5];
  int found;
  if (symtype == 0) found = swb_stats_params_nt(match, mismatch, gap_open, gap_extend, p);
  else if (symtype == 4) found = swb_stats_params(matrix, 32767, 32767, p);
  else found = swb_stats_params(matrix, gap_open, gap_extend, p);
  st->score_threshold = minscore;
  st->upper_threshold = maxscore;
  if (!found) return SWB_OK;
  st->available = 1;
  st->lambda = p[0]; st->K = p[1]; st->H = p[2]; st->alpha = p[3]; st->beta = p[4];
  st->logK = log(st->K);
  int64_t q = qlen;
  if (symtype == 2 || symtype == 4) q = qlen / 3;
  int64_t dlen = effdbsize > 0 ? effdbsize : ((symtype == 3 || symtype == 4) ? symcount / 3 : symcount);
  st->length_adjustment = length_adjustment(st->K, st->logK, st->alpha / st->lambda, st->beta,
                                            (double)(int)q, (double)dlen, (double)(int)seqcount);
  st->m = q - st->length_adjustment;
  st->n = effdbsize > 0 ? effdbsize : dlen - seqcount * st->length_adjustment;
  st->Kmn = st->K * (double)st->m * (double)st->n;
  const int64_t lo = (int64_t)ceil(-log(expect / st->Kmn) / st->lambda);
  if (lo > minscore) st->score_threshold = lo;
  if (minexpect > 0.0)
  {
    const int64_t hi = (int64_t)floor(-log(minexpect / st->Kmn) / st->lambda);
    if (hi < maxscore) st->upper_threshold = hi;
  }
  return SWB_OK;
}

double swb_stats_evalue(const swb_stats *st, int64_t score)
{
  return st && st->available ? st->Kmn * exp(-st->lambda * (double)score) : 0.0;
}

double swb_stats_bits(const swb_stats *st, int64_t score)
{
  if (!st || !st->available) return 0.0;
  return st->lambda / log(2.0) * (double)score - st->logK / log(2.0);
}

}  // extern "C"
