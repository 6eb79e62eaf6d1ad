/*
 * swipe_b200.h -- C ABI of the B200 (sm_100a) score-only Smith-Waterman scan.
 *
 * This is the drop-in boundary for the search path of torognes/swipe.  The reference has no
 * plugin interface; its seam is four free functions that search_chunk() calls
 * (reference swipe.h:200-258, call sites swipe.cc:1432-1455, :1499-1510, :1578-1585):
 *
 *     search7 / search7_ssse3   (swipe.h:200-222)   7-bit pass over a list of subjects
 *     search16                  (swipe.h:224-235)   16-bit pass over the survivors
 *     fullsw                    (swipe.h:251-258)   63-bit pass over what is left
 *     search16s                 (swipe.h:237-249)   16-bit pass that also reports the end cell
 *
 * A GPU cannot be fed at the reference's per-chunk granularity (a few hundred subjects,
 * swipe.cc:479-481), so the boundary sits one level up: "search_chunk over a whole database
 * shard" with the width cascade hidden inside.  The observable output of that cascade is the
 * exact 63-bit score of every subject (anything at or above a width's limit is discarded and
 * recomputed, swipe.cc:1464, :1519), and that is what these entry points return.
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes, no exceptions cross the boundary;
 *   - every function returns SWB_OK (0) or a negative swb_status; swb_strerror() names it.  The
 *     reference reports every error through fatal() -> exit(1) (swipe.cc:158-170); the host shim
 *     keeps that behaviour by calling fatal(swb_strerror(rc)) on a non-zero status;
 *   - the caller owns every buffer it passes; the library owns what swb_db_open returns;
 *   - host pointers unless a parameter says "device";
 *   - one host thread per handle at a time; different handles (one per GPU) may be driven
 *     concurrently from different threads, like the reference's per-thread search_data.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute call fails with
 *     SWB_ERR_NO_DEVICE / SWB_ERR_CUDA.
 */
#ifndef SWIPE_B200_H
#define SWIPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_ABI_VERSION 2
#define SWB_MATRIX_DIM 32 /* reference swipe.h:66-68: score tables are 32 x 32 */

typedef enum swb_status
{
  SWB_OK = 0,
  SWB_ERR_ARG = -1,       /* null pointer, negative size, symbol code > 31, bad offsets */
  SWB_ERR_NO_DEVICE = -2, /* no CUDA device / not an sm_100 part */
  SWB_ERR_CUDA = -3,      /* a CUDA runtime call failed; swb_last_cuda_error() has the text */
  SWB_ERR_NOMEM = -4,
  SWB_ERR_RANGE = -5,     /* scoring parameters outside what the kernels represent exactly */
  SWB_ERR_INTERNAL = -6,
  SWB_ERR_IO = -7         /* a database file is missing, truncated or corrupt; swb_blastdb_error() */
} swb_status;

/* Scoring parameters, exactly the values the reference hands its kernels.
 *   matrix          : 32 x 32 scores indexed [(db_symbol << 5) + query_symbol], the layout of
 *                     the reference's score_matrix_63 (matrices.cc:531-538, search63.cc:52-58).
 *   gap_open_extend : penalty of the first gap position = open + extend (swipe.cc:1126; this is
 *                     what search7/search16/fullsw receive as "gap_open_penalty").
 *   gap_extend      : penalty of each further gap position.                                   */
typedef struct swb_scoring
{
  const int64_t *matrix;
  int64_t gap_open_extend;
  int64_t gap_extend;
} swb_scoring;

/* Per-search bookkeeping, mirroring the reference's counters compute7/compute16/compute63
 * (swipe.cc:111-119, bumped at :1425, :1494, :1552) plus what the GPU path adds.             */
typedef struct swb_counters
{
  int64_t subjects;       /* subjects scored by this call                                      */
  int64_t cells;          /* sum(subject length) * qlen: numerator of the reference's GCUPS    */
  int64_t ref_width7;     /* subjects the reference would have kept from its 7-bit pass        */
  int64_t ref_width16;    /* ... from its 16-bit pass                                          */
  int64_t ref_width63;    /* ... from fullsw                                                   */
  int64_t gpu_narrow;     /* subjects finished by the first packed kernel                         */
  int64_t gpu_requeued;   /* subjects whose lane left the first kernel's exact range, re-queued  */
  int64_t gpu_middle;     /* ... of those finished by the packed int16 build (the 16-bit tier);  */
                          /* the remaining gpu_requeued - gpu_middle went to the wide kernel     */
  int64_t kernel_launches;/* CUDA kernels launched by this call                                */
  double scan_ms;         /* device time of the scan kernels (CUDA events on the handle stream)*/
  double requeue_ms;      /* device time of the wide re-queue kernels                          */
  /* which build of the scan kernel ran the first tier (0 when it was skipped): geometry 1 = a warp
     holds four pipeline stages of eight streams, 2 = a warp is one stage of 32 streams; G stages of
     R query rows each, `scan_passes` passes over the residue stream (ABI version 2)                */
  int64_t scan_geometry, scan_G, scan_R, scan_passes;
} swb_counters;

typedef struct swb_db swb_db; /* opaque: one database shard resident on one GPU */

/* ---- library / device ------------------------------------------------------------------- */
int swb_abi_version(void);
const char *swb_strerror(int status);
const char *swb_last_cuda_error(void);              /* text of the last CUDA failure (thread local) */
int swb_device_count(int *count);                    /* SWB_ERR_NO_DEVICE when there is none  */

/* Pinned host memory for callers that want full-speed uploads (optional). */
int swb_host_alloc(void **ptr, int64_t bytes);
int swb_host_free(void *ptr);
/* Device buffers of closed handles are kept for reuse (cudaMalloc/cudaFree synchronise the device);
 * swb_trim() returns them to the driver. */
int swb_trim(void);
/* Budget of that cache in bytes PER DEVICE (default 8 GiB, or SWB_CACHE_MB from the environment);
 * a buffer that would push a device's cached total beyond it is freed at once.  0 = no caching. */
int swb_set_cache_limit(int64_t bytes_per_device);

/* ---- database shard ----------------------------------------------------------------------
 * Replaces db_mapsequences + the db_getsequence pulls the reference kernels make while they
 * run (search7.cc:899-917, database.cc:1082-1131, :1237-1401): the shard is uploaded once and
 * re-laid-out on the device (length-sorted, two subjects per 32-bit lane pair, 4 columns per
 * block) so that the scan kernels stream it with coalesced loads.
 *
 *   residues : symbol codes 0..31 (protein: NCBIstdaa as stored in a .psq, query.cc:178;
 *              nucleotide: the reference's 4-bit one-hot codes, database.cc:915-921)
 *   offsets  : nseq+1 byte offsets into residues; subject i occupies
 *              residues[offsets[i] .. offsets[i+1] - trailing)
 *   trailing : bytes to drop at the end of every subject: 1 when offsets index a raw .psq
 *              (NUL separators, database.cc:1246-1248, search7.cc:916), 0 for packed input
 *   stream   : a cudaStream_t (as void*) all work of this handle is issued on, or NULL for a
 *              stream the handle creates itself
 */
int swb_db_open(int device, const uint8_t *residues, const int64_t *offsets, int64_t nseq,
                int trailing, void *stream, swb_db **db);
/* Same, but returns as soon as the copies and re-layout kernels are enqueued: the shard is cut
 * into ~256 MB chunks whose upload, re-layout and (once swb_search is called) scan overlap.  The
 * host buffers must stay valid and unchanged until swb_db_wait() or the first search on the
 * handle has returned.                                                                         */
int swb_db_open_async(int device, const uint8_t *residues, const int64_t *offsets, int64_t nseq,
                      int trailing, void *stream, swb_db **db);
int swb_db_wait(swb_db *db);
/* Closing releases the handle; its device buffers go to the reuse cache described at swb_trim()
 * (up to the per-device budget), so device memory is not necessarily returned to the driver until
 * swb_trim() is called.                                                                          */
int swb_db_close(swb_db *db);
int swb_db_info(const swb_db *db, int64_t *nseq, int64_t *total_residues, int64_t *longest);

/* ---- BLAST database files ------------------------------------------------------------------
 * Host-side reader of the version-4 BLAST databases the reference searches (db_open,
 * database.cc:406-608, :775-925): basename.pin/.psq/.phr (protein) or .nin/.nsq/.nhr
 * (nucleotide), optionally behind a .pal/.nal alias file listing several volumes (DBLIST) and a
 * membership mask (OIDLIST + MEMB_BIT).  Global sequence numbers run through the volumes in order
 * (seqno_volume, database.cc:637-660).  Files are memory-mapped; nothing is decoded until asked.
 */
typedef struct swb_blastdb swb_blastdb;
int swb_blastdb_open(const char *basename, int nucleotide, swb_blastdb **out);
int swb_blastdb_close(swb_blastdb *b);
const char *swb_blastdb_error(void);   /* text of the last open failure on this thread */
int swb_blastdb_info(const swb_blastdb *b, int64_t *nseq, int64_t *symbols, int64_t *longest,
                     int *volumes);
/* membership bit and the masked totals of an alias database (database.cc:1046-1065) */
int swb_blastdb_masked_info(const swb_blastdb *b, int64_t *memb_bit, int64_t *nseq, int64_t *symbols);
const char *swb_blastdb_title(const swb_blastdb *b);
const char *swb_blastdb_date(const swb_blastdb *b);
/* length in residues / nucleotides (database.cc:1246-1261), or a negative status */
int64_t swb_blastdb_seqlen(const swb_blastdb *b, int64_t seqno);
/* one sequence as symbol codes, as db_getsequence returns it (database.cc:1237-1353): protein
 * bytes as stored; nucleotides as 4-bit codes with ambiguity runs applied, reverse-complemented
 * when strand != 0.  *len receives the length even when cap is too small (SWB_ERR_RANGE).     */
int swb_blastdb_sequence(const swb_blastdb *b, int64_t seqno, int strand, uint8_t *buf,
                         int64_t cap, int64_t *len);
/* the raw ASN.1 defline bytes (db_getheader, database.cc:1403-1413); points into the mapping   */
int swb_blastdb_header(const swb_blastdb *b, int64_t seqno, const uint8_t **data, int64_t *len);
/* 1 when the sequence passes the alias file's membership mask (db_check_msk, database.cc:687-706) */
int swb_blastdb_included(const swb_blastdb *b, int64_t seqno);

/* Uploads sequences [first, first + count) (count < 0: to the end) as one shard.  Protein volumes
 * are copied as they lie in the .psq; nucleotide volumes are copied packed (4 bases per byte +
 * ambiguity tables) and unpacked ON THE DEVICE to the 4-bit codes db_getsequence produces.  Subject
 * i of the shard is global sequence first + i.  async != 0 behaves like swb_db_open_async (the
 * swb_blastdb must stay open until swb_db_wait or the first search returns).                    */
int swb_db_open_blast(int device, const swb_blastdb *b, int64_t first, int64_t count, int async,
                      void *stream, swb_db **db);

/* Same for a nucleotide database searched as six-frame translated protein (-p 3 / -p 4): sequence
 * first + s becomes subjects 6 s + 3 strand + frame (search_chunk's order, swipe.cc:1377-1385),
 * unpacked and translated on the device (db_translate, database.cc:1182-1218) with the codon table
 * of swb_translate_table.                                                                        */
int swb_db_open_blast_translated(int device, const swb_blastdb *b, int64_t first, int64_t count,
                                 const uint8_t *codon_table, int async, void *stream, swb_db **db);

/* ---- the hot path ------------------------------------------------------------------------
 * swb_search: what search_chunk's cascade (swipe.cc:1416-1594) yields for every subject of the
 * shard: scores[i] = exact affine-gap local alignment score of the query against subject i, in
 * the order the subjects were given to swb_db_open.  query holds symbol codes 0..31
 * (query.cc:317-325).  counters may be NULL.
 */
int swb_search(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
               int64_t *scores, swb_counters *counters);

/* swb_search_list: the same for a list of subjects given the way the reference's kernels take
 * it -- codes (seqno << 3) | (strand << 2) | frame (swipe.cc:1373-1391); strand and frame must
 * be 0 here (protein / forward-strand database symbols).  scores[k] belongs to seqnos[k], the
 * contract of search7/search16 (search7.cc:894-895).
 */
int swb_search_list(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                    const int64_t *seqnos, int64_t n, int64_t *scores, swb_counters *counters);

/* swb_search_hits: swb_search followed by the sink's admission rule ON THE DEVICE, so that only
 * the hits the reference would have kept cross the bus.  hits_enter (hits.cc:163-222) never stores
 * a score below scorethreshold or above upperscorethreshold (:180-184) and, once `keep` hits are
 * held, raises the threshold to the last kept score (:218-219); the final list is therefore the
 * best `keep` admissible subjects ordered by score descending, then sequence number descending
 * (:188-191).  This call returns exactly that list for the shard: out_seqno[k] = seqno_base +
 * subject number, out_score[k], k < *nhits <= keep; *totalhits = subjects with score >= min_score,
 * *obvious = subjects with score > upper_score (hits.cc:174-178).  Device work: a histogram of the
 * admissible scores, the bin of the keep-th score, a compaction of everything at or above it, a
 * radix sort of those candidates; keep * 16 bytes come back instead of 8 bytes per subject.
 * The per-shard lists of several GPUs are combined with swb_hits_merge.                          */
int swb_search_hits(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                    int64_t seqno_base, int64_t keep, int64_t min_score, int64_t upper_score,
                    int64_t *out_seqno, int64_t *out_score, int64_t *nhits, int64_t *totalhits,
                    int64_t *obvious, swb_counters *counters);

/* swb_search_batch / swb_search_hits_batch: several queries against the shard, the reference's query
 * loop (swipe.cc:2561) folded into as few scans as possible.  Queries that fit side by side on the 16
 * pipeline stages of the scan kernel (sum of ceil(qlen / 25) <= 16, i.e. up to 400 query rows) share ONE
 * pass over the residue stream and ONE substitution-table build per 4-column block; longer ones are
 * searched one after the other.  Results are exactly those of nqueries separate swb_search /
 * swb_search_hits calls (tests/test_gpu_batch.py): scores[k][i] / the hit list of query k.  counters,
 * when given, is an array of nqueries entries; the shared scan's time is split evenly over its queries. */
int swb_search_batch(swb_db *db, int nqueries, const uint8_t *const *queries, const int64_t *qlens,
                     const swb_scoring *scoring, int64_t *const *scores, swb_counters *counters);
int swb_search_hits_batch(swb_db *db, int nqueries, const uint8_t *const *queries, const int64_t *qlens,
                          const swb_scoring *scoring, int64_t seqno_base, int64_t keep, int64_t min_score,
                          int64_t upper_score, int64_t *const *out_seqno, int64_t *const *out_score,
                          int64_t *nhits, int64_t *totalhits, int64_t *obvious, swb_counters *counters);

/* Subject filter of the device sink ("subject_filter all | bitmap | coded list"): bit (k & 7) of
 * bitmap[k >> 3] says whether subject k may enter the hit list -- what db_check_inclusion decides per sequence
 * for alias masks and taxid lists (swipe.cc:1373-1376, database.cc:687-733).  Excluded subjects are still
 * scored (the scan streams the whole shard) but neither counted in totalhits / obvious nor admitted by
 * swb_search_hits / swb_search_hits_batch; swb_search (dense scores) ignores the filter.  The bitmap is
 * copied; NULL removes the filter.  A coded list is what swb_search_list takes.                          */
int swb_db_set_filter(swb_db *db, const uint8_t *bitmap);

/* Merges hit lists that are each in the sink's order (what swb_search_hits returns) into the best
 * `keep` overall -- the master's merge of the reference's MPI build (swipe.cc:1957-1974) and the
 * host-side step of a multi-GPU search.  Returns the number of hits written or a negative status. */
int64_t swb_hits_merge(int nlists, const int64_t *const *seqnos, const int64_t *const *scores,
                       const int64_t *n, int64_t keep, int64_t *out_seqno, int64_t *out_score);

/* swb_search_end: search16s's contract (swipe.h:237-249, search16s.cc:390-405, called from
 * align_chunk swipe.cc:381-393): exact score plus the alignment end -- bestpos = first subject
 * column (0-based) in which the maximum is reached, bestq = smallest query row reaching it in
 * that column; both -1 when the score is 0.  Scores are exact (never saturated at 65535).
 * A set strand bit (seqnos[k] & 4) scores the reverse complement of a nucleotide subject, which is
 * how align_chunk asks for minus-strand hits (swipe.cc:355-361, database.cc:1327-1353).
 */
int swb_search_end(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                   const int64_t *seqnos, int64_t n, int64_t *scores, int64_t *bestpos,
                   int64_t *bestq);

/* ---- scoring system and statistics (host) ---------------------------------------------------
 * Substitution tables as score_matrix_init builds them (matrices.cc:520-591): 32 x 32,
 * [(subject << 5) + query], -1 where undefined.
 */
int swb_matrix_builtin(const char *name, int64_t *matrix);        /* BLOSUM45/50/62/80/90, PAM30/70/250, identity_5_1 */
int swb_matrix_parse(const char *text, int64_t *matrix);          /* matrix file text (matrices.cc:437-517) */
int swb_matrix_read(const char *name_or_path, int64_t *matrix);   /* built-in name, else a file */
int swb_matrix_read_sound(const char *name_or_path, int64_t *matrix);   /* -p 5: IDENTITY_5_1 or a file in the sound alphabet */
int swb_matrix_nucleotide(int64_t match, int64_t mismatch, int64_t *matrix);   /* matrices.cc:533-538 */
int swb_matrix_limits(const int64_t *matrix, int64_t *lo, int64_t *hi, int64_t *limit7, int64_t *limit16);

/* Karlin-Altschul parameters (NCBI BLAST's tables; stats.cc:44-325): params = lambda, K, H, alpha,
 * beta.  Return 1 when the scoring system is tabulated, 0 otherwise.                            */
int swb_stats_params(const char *matrix, int64_t gap_open, int64_t gap_extend, double *params);
int swb_stats_params_nt(int64_t match, int64_t mismatch, int64_t gap_open, int64_t gap_extend, double *params);
int swb_stats_default_gaps(const char *matrix, int64_t *gap_open, int64_t *gap_extend);
int64_t swb_stats_length_adjustment(double K, double alpha_d_lambda, double beta, int64_t qlen,
                                    int64_t dblen, int64_t nseq);

/* The search space and raw-score window hits_init derives (hits.cc:283-511). */
typedef struct swb_stats
{
  int available;                 /* 0: no parameters for this scoring system, scores only          */
  double lambda, K, H, alpha, beta, logK, Kmn;
  int64_t length_adjustment, m, n;
  int64_t score_threshold;       /* hits below are dropped: max(minscore, ceil(-ln(expect/Kmn)/lambda)) */
  int64_t upper_threshold;       /* hits above are "obvious" and dropped (-u / -k)                  */
} swb_stats;
int swb_stats_init(int symtype, const char *matrix, int64_t match, int64_t mismatch, int64_t gap_open,
                   int64_t gap_extend, int64_t qlen, int64_t symcount, int64_t seqcount,
                   int64_t effdbsize, int64_t minscore, int64_t maxscore, double expect,
                   double minexpect, swb_stats *out);
double swb_stats_evalue(const swb_stats *st, int64_t score);     /* Kmn * exp(-lambda * score) */
double swb_stats_bits(const swb_stats *st, int64_t score);       /* (lambda * score - ln K) / ln 2 */

/* ---- query text, translation, deflines (host) ------------------------------------------------ */
/* One FASTA record -> symbol codes (query.cc:244-355); returns the bytes of text consumed.
 * nucleotide: 0 = amino acids, 1 = nucleotides, 2 = the sound alphabet of -p 5.                  */
int64_t swb_query_parse(const char *text, int64_t text_len, int nucleotide, uint8_t *seq,
                        int64_t seq_cap, int64_t *seq_len, char *descr, int64_t descr_cap);
int swb_revcomp(const uint8_t *seq, int64_t len, uint8_t *out);   /* 4-bit nt codes (query.cc:357-364) */
/* 4096-entry codon table of NCBI genetic code 1..23 over 4-bit nucleotide codes (query.cc:366-436) */
int swb_translate_table(int gencode, uint8_t *table);
const char *swb_gencode_name(int gencode);
int64_t swb_translate(const uint8_t *nt, int64_t len, int strand, int frame, const uint8_t *table,
                      uint8_t *out);                               /* query.cc:450-506 */
/* The deflines of a .phr/.nhr record as text, one per line (asnparse.cc:753-887); returns their
 * count, *needed = bytes required (SWB_ERR_RANGE when cap is too small).  memb / taxids filter the
 * deflines like the reference's -x list and alias MEMB_BIT do (database.cc:718-733, :1457-1481).  */
int64_t swb_defline_text(const uint8_t *data, int64_t len, int show_gis, int show_taxid, int64_t memb,
                         const uint8_t *taxids, int64_t taxid_bytes, char *buf, int64_t cap,
                         int64_t *needed);

/* ---- alignment of a hit (host) -------------------------------------------------------------
 * swb_align: the reference's align() (align.cc:469-519) as hits_align calls it for the best -b hits
 * (hits.cc:587-623).  Host code: a reverse pass from the end cell finds the start, then a
 * linear-space divide-and-conquer recovers the path.  Ties are broken as the reference breaks
 * them, so the same alignment is reported.
 *   matrix, query, subject : as for the scan ([(subject_symbol << 5) + query_symbol])
 *   gap_open, gap_extend   : -G / -E as given on the command line (NOT open+extend)
 *   *score, *q_end, *d_end : in/out.  *score != 0 on entry = the end cell is already known (the
 *                            hint from swb_search_end / search16s, hits.cc:589-600); 0 = find it
 *                            with a forward pass (first strict maximum in query-major order)
 *   q_start, d_start       : out, first aligned query / subject position (0-based)
 *   ops                    : out, NUL-terminated run-length string "M<n>I<n>D<n>..." (M aligned
 *                            pair, I subject symbols against a gap, D query symbols against a gap);
 *                            *ops_len receives its length, SWB_ERR_RANGE if ops_cap is too small
 */
int swb_align(const uint8_t *query, int64_t qlen, const uint8_t *subject, int64_t dlen,
              const int64_t *matrix, int64_t gap_open, int64_t gap_extend, int64_t *q_start,
              int64_t *d_start, int64_t *q_end, int64_t *d_end, int64_t *score, char *ops,
              int64_t ops_cap, int64_t *ops_len);

/* ---- the sink ----------------------------------------------------------------------------
 * swb_topk_merge: the insertion rule of hits_enter (hits.cc:163-222) applied to the scores of
 * one or more shards: reject score < min_score or > upper_score, order by score descending then
 * seqno descending, keep at most `keep` entries; totalhits / obvious as hits.cc:174-178.
 * shard k contributes n[k] subjects whose global sequence numbers are seqno_base[k] + i.
 * Returns the number of hits kept (>= 0) or a negative status.
 */
int64_t swb_topk_merge(int nshards, const int64_t *const *scores, const int64_t *n,
                       const int64_t *seqno_base, int64_t keep, int64_t min_score,
                       int64_t upper_score, int64_t *out_seqno, int64_t *out_score,
                       int64_t *totalhits, int64_t *obvious);

/* ---- tuning / introspection (not part of the reference's surface) -------------------------- */
/* Measured issue rate of the packed 16x2 DPX instructions (VIADDMNMX / VIMNMX3) the scan is bound by,
 * in warp instructions per clock per SM, and the SM clock the measurement ran at.  A DP cell pair
 * costs 3.5 of them, so rate * 64 / 3.5 cells per clock per SM is the scan's integer-issue ceiling on
 * this device -- the roofline denominator bench.py reports (calibrated by microbenchmark). */
int swb_alu_peak(int device, double *dpx_per_clk_per_sm, double *sm_clock_mhz);
/* Forces every subject through one kernel family: 0 = cascade (default), 1 = packed 16-bit
 * lanes only is not allowed (would not be exact) -> rejected; 2 = wide kernel for everything.  */
int swb_set_mode(swb_db *db, int mode);
/* Device time of the open: first byte sent .. last chunk laid out, and first layout kernel ..
 * last chunk laid out (the two overlap), in ms.  Waits for the open to finish.                 */
int swb_db_open_ms(const swb_db *db, double *upload_ms, double *layout_ms);
/* Test hook: pin the scan kernel shape (G threads per stream, R query rows per thread) and the
 * lane arithmetic (0 = DPX int16 only, 1 = DPX + fp16-pattern adds); (0, 0, -1) = automatic.   */
int swb_set_shape(swb_db *db, int G, int R, int lane_mode);
/* Test hook: which scan kernel geometry may be chosen: 1 = a warp holds four pipeline stages of eight
 * streams (swb_scan_kernel), 2 = a warp is one stage of 32 streams (swb_scan2_kernel), 0 = automatic.
 * Call before swb_set_shape when both are used.                                                  */
int swb_set_geometry(swb_db *db, int geometry);

#ifdef __cplusplus
}
#endif
#endif /* SWIPE_B200_H */
