// ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
//
// Glue (written for this repo) that lets tests and bench.py's reference arm drive the
// UNMODIFIED reference kernels -- search7 / search7_ssse3 (search7.cc), search16
// (search16.cc), search16s (search16s.cc), fullsw (search63.cc), score_matrix_init
// (matrices.cc) and align (align.cc) -- compiled from /root/reference by oracle/Makefile
// into oracle/_ref/.  Nothing from the reference is copied here: this file only supplies
// the handful of symbols those objects import (db_getsequence, db_mapsequences, fatal,
// xmalloc, xrealloc, out, symtype, matrixname, matchscore, mismatchscore) and a driver
// that plays the role of search_chunk (swipe.cc:1365-1596) over in-memory subjects.
//
// Built only where /root/reference exists; the resulting .so travels to the GPU box.
#include "swipe.h"
#include <stdint.h>
#include <vector>
#include <atomic>
#include <thread>

// ---- symbols the reference objects import ---------------------------------------------
FILE *out = stdout;
long symtype = 1;
const char *matrixname = "BLOSUM62";
long matchscore = 1;
long mismatchscore = -3;

void fatal(const char *message)
{
  fprintf(stderr, "ref_harness fatal: %s\n", message);
  exit(1);
}

void fatal(const char *format, const char *message)
{
  fprintf(stderr, format, message);
  fprintf(stderr, "\n");
  exit(1);
}

void *xmalloc(size_t size)
{
  void *p = NULL;
  if (posix_memalign(&p, 16, size ? size : 16) != 0 || !p) fatal("out of memory");
  return p;
}

void *xrealloc(void *ptr, size_t size)
{
  void *p = realloc(ptr, size);
  if (!p) fatal("out of memory");
  return p;
}

// In-memory stand-in for the reference's database thread handle.
struct db_thread_s
{
  const unsigned char *residues;
  const int64_t *offsets;
};

// length counts one trailing separator, as the reference's reader reports it
// (database.cc:1246-1248; kernels subtract it again, search7.cc:916).
void db_getsequence(struct db_thread_s *t, long seqno, long, long,
                    char **address, long *length, long *ntlen, int)
{
  *address = (char *)(t->residues + t->offsets[seqno]);
  *length = (long)(t->offsets[seqno + 1] - t->offsets[seqno]) + 1;
  *ntlen = 0;
}

void db_mapsequences(struct db_thread_s *, long, long) {}

// ---- exported C entry points ---------------------------------------------------------
extern "C" {

// Runs the reference's own table construction.  matrix: built-in name or a file path;
// symtype 0 selects the nucleotide table from (match, mismatch).
void ref_matrix_init(const char *matrix, long sym, long match, long mismatch)
{
  if (score_matrix_63) score_matrix_free();
  symtype = sym;
  matrixname = matrix ? strdup(matrix) : "BLOSUM62";
  matchscore = match;
  mismatchscore = mismatch;
  score_matrix_init();
}

void ref_matrix_get(int64_t *m63, int64_t *limit7, int64_t *limit16)
{
  for (int i = 0; i < 1024; i++) m63[i] = score_matrix_63[i];
  *limit7 = SCORELIMIT_7;
  *limit16 = SCORELIMIT_16;
}

long ref_fullsw(const unsigned char *d, long dlen, const unsigned char *q, long qlen,
                long gapopenextend, long gapextend)
{
  long *he = (long *)xmalloc(sizeof(long) * 2 * (qlen > 0 ? qlen : 1));
  long s = fullsw((char *)d, (char *)d + dlen, (char *)q, (char *)q + qlen, he,
                  score_matrix_63, (BYTE)gapopenextend, (BYTE)gapextend);
  free(he);
  return s;
}

struct ref_worker_ctx
{
  const unsigned char *residues;
  const int64_t *offsets;
  long nseq;
  const unsigned char *q;
  long qlen;
  long goe, ge;
  long chunk;
  int ssse3;
  int64_t *scores;
  unsigned char *width;
  std::atomic<long> next;
  std::atomic<long> c7, c16, c63;
};

// One worker: the cascade of swipe.cc:1416-1594 over chunks of consecutive subjects.
static void ref_worker(ref_worker_ctx *cx)
{
  db_thread_s dbt = {cx->residues, cx->offsets};
  long qlen = cx->qlen;
  BYTE *dprofile = (BYTE *)xmalloc(4 * 16 * 32);
  BYTE *hearray = (BYTE *)xmalloc((qlen > 0 ? qlen : 1) * 32);
  BYTE **qtable = (BYTE **)xmalloc(sizeof(BYTE *) * (qlen > 0 ? qlen : 1));
  for (long i = 0; i < qlen; i++) qtable[i] = dprofile + 64 * cx->q[i];
  std::vector<long> in(cx->chunk), outl(cx->chunk), sc(cx->chunk), bp(cx->chunk);

  for (;;)
  {
    long first = cx->next.fetch_add(cx->chunk);
    if (first >= cx->nseq) break;
    long n = cx->nseq - first < cx->chunk ? cx->nseq - first : cx->chunk;
    for (long i = 0; i < n; i++) in[i] = (first + i) << 3;

    cx->c7 += n;
    if (cx->ssse3)
      search7_ssse3(qtable, (BYTE)cx->goe, (BYTE)cx->ge, (BYTE *)score_matrix_7t, dprofile,
                    hearray, &dbt, n, in.data(), sc.data(), qlen);
    else
      search7(qtable, (BYTE)cx->goe, (BYTE)cx->ge, (BYTE *)score_matrix_7, dprofile, hearray,
              &dbt, n, in.data(), sc.data(), qlen);
    long m = 0;
    for (long i = 0; i < n; i++)
    {
      long seqno = in[i] >> 3;
      if (sc[i] < SCORELIMIT_7)
      {
        cx->scores[seqno] = sc[i];
        if (cx->width) cx->width[seqno] = 7;
      }
      else
        outl[m++] = in[i];
    }
    if (m == 0) continue;

    cx->c16 += m;
    in.swap(outl);
    n = m;
    search16((WORD **)qtable, (WORD)cx->goe, (WORD)cx->ge, (WORD *)score_matrix_16,
             (WORD *)dprofile, (WORD *)hearray, &dbt, n, in.data(), sc.data(), bp.data(),
             (int)qlen);
    m = 0;
    for (long i = 0; i < n; i++)
    {
      long seqno = in[i] >> 3;
      if (sc[i] < SCORELIMIT_16)
      {
        cx->scores[seqno] = sc[i];
        if (cx->width) cx->width[seqno] = 16;
      }
      else
        outl[m++] = in[i];
    }
    if (m == 0) continue;

    cx->c63 += m;
    long *he = (long *)xmalloc(sizeof(long) * 2 * (qlen > 0 ? qlen : 1));
    for (long i = 0; i < m; i++)
    {
      long seqno = outl[i] >> 3;
      const unsigned char *d = cx->residues + cx->offsets[seqno];
      long dlen = (long)(cx->offsets[seqno + 1] - cx->offsets[seqno]);
      cx->scores[seqno] = fullsw((char *)d, (char *)d + dlen, (char *)cx->q,
                                 (char *)cx->q + qlen, he, score_matrix_63, (BYTE)cx->goe,
                                 (BYTE)cx->ge);
      if (cx->width) cx->width[seqno] = 63;
    }
    free(he);
  }
  free(dprofile);
  free(hearray);
  free(qtable);
}

// Scores every subject with the reference cascade.  counts = {compute7, compute16, compute63}
// (swipe.cc:1425, :1494, :1552).  ssse3 != 0 selects search7_ssse3 as the CLI does on SSSE3 hosts.
void ref_scan(const unsigned char *residues, const int64_t *offsets, long nseq,
              const unsigned char *q, long qlen, long gapopen, long gapextend,
              int threads, long chunk, int ssse3,
              int64_t *scores, unsigned char *width, long counts[3])
{
  ref_worker_ctx cx;
  cx.residues = residues; cx.offsets = offsets; cx.nseq = nseq;
  cx.q = q; cx.qlen = qlen;
  cx.goe = gapopen + gapextend; cx.ge = gapextend;
  cx.chunk = chunk > 0 ? chunk : 1024;
  cx.ssse3 = ssse3;
  cx.scores = scores; cx.width = width;
  cx.next = 0; cx.c7 = 0; cx.c16 = 0; cx.c63 = 0;
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(ref_worker, &cx);
  ref_worker(&cx);
  for (auto &t : pool) t.join();
  if (counts) { counts[0] = cx.c7; counts[1] = cx.c16; counts[2] = cx.c63; }
}

// search16 alone on a list of subjects (scores saturate at 65535; bestpos is the
// block-granular end column hint of search16.cc:411-414, :464).
void ref_search16(const unsigned char *residues, const int64_t *offsets, long nseq,
                  const unsigned char *q, long qlen, long gapopen, long gapextend,
                  int64_t *scores, int64_t *bestpos)
{
  db_thread_s dbt = {residues, offsets};
  BYTE *dprofile = (BYTE *)xmalloc(4 * 16 * 32);
  BYTE *hearray = (BYTE *)xmalloc((qlen > 0 ? qlen : 1) * 32);
  BYTE **qtable = (BYTE **)xmalloc(sizeof(BYTE *) * (qlen > 0 ? qlen : 1));
  for (long i = 0; i < qlen; i++) qtable[i] = dprofile + 64 * q[i];
  std::vector<long> in(nseq), sc(nseq), bp(nseq);
  for (long i = 0; i < nseq; i++) in[i] = i << 3;
  search16((WORD **)qtable, (WORD)(gapopen + gapextend), (WORD)gapextend,
           (WORD *)score_matrix_16, (WORD *)dprofile, (WORD *)hearray, &dbt, nseq, in.data(),
           sc.data(), bp.data(), (int)qlen);
  for (long i = 0; i < nseq; i++) { scores[i] = sc[i]; bestpos[i] = bp[i]; }
  free(dprofile); free(hearray); free(qtable);
}

// search16s on a list of subjects: score plus exact alignment end (bestq, bestpos),
// called the way align_chunk does (swipe.cc:381-393).
void ref_search16s(const unsigned char *residues, const int64_t *offsets, long nseq,
                   const unsigned char *q, long qlen, long gapopen, long gapextend,
                   int64_t *scores, int64_t *bestpos, int64_t *bestq)
{
  db_thread_s dbt = {residues, offsets};
  BYTE *dprofile = (BYTE *)xmalloc(4 * 16 * 32);
  BYTE *hearray = (BYTE *)xmalloc((qlen > 0 ? qlen : 1) * 32);
  BYTE **qtable = (BYTE **)xmalloc(sizeof(BYTE *) * (qlen > 0 ? qlen : 1));
  // the alignment phase builds 8 channels x 1 column profiles: 16 bytes per query symbol
  // (swipe.cc:225-255), not the 64 of the search phase (swipe.cc:1202-1232)
  for (long i = 0; i < qlen; i++) qtable[i] = dprofile + 16 * q[i];
  std::vector<long> in(nseq), sc(nseq), bp(nseq), bq(nseq);
  for (long i = 0; i < nseq; i++) in[i] = i << 3;
  db_thread_s *dbta[8];                       // one handle per SIMD channel (swipe.cc:386)
  for (int c = 0; c < 8; c++) dbta[c] = &dbt;
  search16s((WORD **)qtable, (WORD)(gapopen + gapextend), (WORD)gapextend,
            (WORD *)score_matrix_16, (WORD *)dprofile, (WORD *)hearray, dbta, nseq, in.data(),
            sc.data(), bp.data(), bq.data(), (int)qlen);
  for (long i = 0; i < nseq; i++) { scores[i] = sc[i]; bestpos[i] = bp[i]; bestq[i] = bq[i]; }
  free(dprofile); free(hearray); free(qtable);
}

// The reference's aligner (align.cc:469-519) as hits_align calls it (hits.cc:587-623): in/out
// hints (*score != 0: alignment end already known), out: begin/end coordinates and the run-length
// op string.  Uses the matrix set by ref_matrix_init.  Returns the op string length.
long ref_align(const unsigned char *q, long qlen, const unsigned char *d, long dlen,
               long gapopen, long gapextend, int64_t *coords /* qs, ds, qe, de */,
               int64_t *score, char *ops, long ops_cap)
{
  long qs = 0, ds = 0, qe = coords[2], de = coords[3], s = *score;
  char *alignment = NULL;
  align((char *)q, (char *)d, qlen, dlen, score_matrix_63, gapopen, gapextend, &qs, &ds, &qe, &de,
        &alignment, &s);
  coords[0] = qs; coords[1] = ds; coords[2] = qe; coords[3] = de;
  *score = s;
  long n = (long)strlen(alignment);
  if (n < ops_cap) memcpy(ops, alignment, n + 1);
  free(alignment);
  return n;
}

// Statistics (stats.cc, NCBI's blastkar tables) and the query-side codon table (query.cc:366-444),
// for pinning swb_stats_* and swb_translate_table.
long ref_stats_params(const char *matrix, long go, long ge, double *p)
{
  return stats_getparams(matrix, go, ge, p, p + 1, p + 2, p + 3, p + 4);
}

long ref_stats_params_nt(long match, long mismatch, long go, long ge, double *p)
{
  return stats_getparams_nt(match, mismatch, go, ge, p, p + 1, p + 2, p + 3, p + 4);
}

long ref_stats_prefs(const char *matrix, long *go, long *ge) { return stats_getprefs(matrix, go, ge); }

extern "C++" int BlastComputeLengthAdjustment(double K, double logK, double alpha_d_lambda, double beta,
                                              int query_length, long db_length, int db_num_seqs,
                                              int *length_adjustment);

long ref_length_adjustment(double K, double logK, double a_d_l, double beta, long qlen, long dblen, long nseq)
{
  int adj = 0;
  BlastComputeLengthAdjustment(K, logK, a_d_l, beta, (int)qlen, dblen, (int)nseq, &adj);
  return adj;
}

extern char q_translate[];
void ref_translate_table(long gencode, unsigned char *out)
{
  translate_init(gencode, gencode);
  memcpy(out, q_translate, 4096);
}

}  // extern "C"
