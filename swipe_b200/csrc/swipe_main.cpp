// swipe_main.cpp -- command-line front end over the C ABI of include/swipe_b200.h: the reference's
// options, run header, hit list, statistics and report formats (plain, XML, tab-separated) with the
// database scan on one or more B200s.
//
// Takes over from the reference (torognes/swipe):
//   args_init / args_show ................. swipe.cc:665-782, :827-1157
//   work / main ........................... swipe.cc:2436-2611
//   search_chunk's strand / frame loops ... swipe.cc:1365-1596 (the scoring itself: swb_search)
//   hits_enter / hits_sort ................ hits.cc:78-222
//   align_chunk / hits_align .............. swipe.cc:339-414, hits.cc:546-623 (swb_search_end + swb_align)
//   hits_show_plain / _xml / _tsv ......... hits.cc:647-1176, :1660-1945
//   show_deflines ......................... asnparse.cc:889-971
// Not carried over: the MPI master / slave.  `-a` (threads in the
// reference) selects how many GPUs share the database, one host thread each.
#include "../../include/swipe_b200.h"

#include <getopt.h>
#include <strings.h>
#include <sys/times.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define SWB_CLI_VERSION "0.1 (B200)"

namespace
{

FILE *out = stdout;

[[noreturn]] void fatal(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n");
  exit(1);
}

void check(int rc, const char *what)
{
  if (rc == SWB_OK) return;
  const char *detail = rc == SWB_ERR_IO ? swb_blastdb_error() : swb_last_cuda_error();
  fatal("%s: %s%s%s", what, swb_strerror(rc), detail && *detail ? " - " : "", detail ? detail : "");
}

struct Options
{
  long gapopen = 0, gapextend = 0;
  std::string matrixname, queryname = "-", databasename;
  long minscore = 1, maxscore = LONG_MAX, maxmatches = 250, alignments = 100, threads = 1, view = 0;
  long symtype = 1, show_gis = 0, show_taxid = 0;
  double expect = 10.0, minexpect = 0.0;
  long matchscore = 1, mismatchscore = -3, querystrands = 3, query_gencode = 1, db_gencode = 1;
  long effdbsize = 0, dump = 0;
  const char *outfile = nullptr, *taxidfile = nullptr;
};

const char SYM_AA[] = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ####";
const char SYM_NT[] = "-acmgrsvtwyhkdbn################";
const char SYM_SOUND[] = "-ABCDEFGHIJKLMNOPQRSTUVWXYZabcde";

void usage(const char *prog)
{
  fprintf(out, "Usage: %s [OPTIONS]\n", prog);
  fprintf(out, "  -h, --help                 show help\n");
  fprintf(out, "  -d, --db=FILE              sequence database base name (required)\n");
  fprintf(out, "  -i, --query=FILE           query sequence filename (stdin)\n");
  fprintf(out, "  -M, --matrix=NAME/FILE     score matrix name or filename (BLOSUM62)\n");
  fprintf(out, "  -q, --penalty=NUM          penalty for nucleotide mismatch (-3)\n");
  fprintf(out, "  -r, --reward=NUM           reward for nucleotide match (1)\n");
  fprintf(out, "  -G, --gapopen=NUM          gap open penalty (11)\n");
  fprintf(out, "  -E, --gapextend=NUM        gap extension penalty (1)\n");
  fprintf(out, "  -v, --num_descriptions=NUM sequence descriptions to show (250)\n");
  fprintf(out, "  -b, --num_alignments=NUM   sequence alignments to show (100)\n");
  fprintf(out, "  -e, --evalue=REAL          maximum expect value of sequences to show (10.0)\n");
  fprintf(out, "  -k, --minevalue=REAL       minimum expect value of sequences to show (0.0)\n");
  fprintf(out, "  -c, --min_score=NUM        minimum score of sequences to show (1)\n");
  fprintf(out, "  -u, --max_score=NUM        maximum score of sequences to show (inf.)\n");
  fprintf(out, "  -a, --num_threads=NUM      number of GPUs to use, one host thread each (1)\n");
  fprintf(out, "  -m, --outfmt=NUM           output format [0,7-9,99=plain,xml,tsv,tsv+,paralign xml] (0)\n");
  fprintf(out, "  -I, --show_gis             show gi numbers in results (no)\n");
  fprintf(out, "  -p, --symtype=NAME/NUM     symbol type/translation [0-4] (1)\n");
  fprintf(out, "  -S, --strand=NAME/NUM      query strands to search [1-3] (3)\n");
  fprintf(out, "  -Q, --query_gencode=NUM    query genetic code [1-23] (1)\n");
  fprintf(out, "  -D, --db_gencode=NUM       database genetic code [1-23] (1)\n");
  fprintf(out, "  -x, --taxidlist=FILE       taxid list filename (none)\n");
  fprintf(out, "  -N, --dump=NUM             dump database [0-2=no,yes,split headers] (0)\n");
  fprintf(out, "  -H, --show_taxid           show taxid etc in results (no)\n");
  fprintf(out, "  -o, --out=FILE             output file (stdout)\n");
  fprintf(out, "  -z, --dbsize=NUM           set effective database size (0)\n");
}

Options parse_args(int argc, char **argv)
{
  Options o;
  static struct option longopts[] = {
      {"db", required_argument, nullptr, 'd'}, {"query", required_argument, nullptr, 'i'},
      {"matrix", required_argument, nullptr, 'M'}, {"penalty", required_argument, nullptr, 'q'},
      {"reward", required_argument, nullptr, 'r'}, {"gapopen", required_argument, nullptr, 'G'},
      {"gapextend", required_argument, nullptr, 'E'}, {"strand", required_argument, nullptr, 'S'},
      {"num_descriptions", required_argument, nullptr, 'v'}, {"num_alignments", required_argument, nullptr, 'b'},
      {"min_score", required_argument, nullptr, 'c'}, {"max_score", required_argument, nullptr, 'u'},
      {"evalue", required_argument, nullptr, 'e'}, {"minevalue", required_argument, nullptr, 'k'},
      {"num_threads", required_argument, nullptr, 'a'}, {"outfmt", required_argument, nullptr, 'm'},
      {"symtype", required_argument, nullptr, 'p'}, {"taxid", required_argument, nullptr, 'x'},
      {"comp_based_stats", required_argument, nullptr, 'C'}, {"query_gencode", required_argument, nullptr, 'Q'},
      {"db_gencode", required_argument, nullptr, 'D'}, {"filter", required_argument, nullptr, 'F'},
      {"subalignments", required_argument, nullptr, 'K'}, {"dump", required_argument, nullptr, 'N'},
      {"out", required_argument, nullptr, 'o'}, {"dbsize", required_argument, nullptr, 'z'},
      {"show_gis", no_argument, nullptr, 'I'}, {"show_taxid", no_argument, nullptr, 'H'},
      {"help", no_argument, nullptr, 'h'}, {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "d:i:M:q:r:G:E:S:v:b:c:u:e:k:a:m:p:x:C:Q:D:F:K:N:o:z:IHh", longopts, nullptr)) != -1)
  {
    switch (c)
    {
      case 'a': o.threads = atol(optarg); break;
      case 'b': o.alignments = atol(optarg); break;
      case 'c': o.minscore = atol(optarg); break;
      case 'C':
        if (strcasecmp(optarg, "F") != 0 && strcmp(optarg, "0") != 0)
          fatal("Composition-based score adjustments not supported.");
        break;
      case 'd': o.databasename = optarg; break;
      case 'D': o.db_gencode = atol(optarg); break;
      case 'e': o.expect = atof(optarg); break;
      case 'E': o.gapextend = atol(optarg); break;
      case 'F':
        if (strlen(optarg) != 0 && strcasecmp(optarg, "F") != 0) fatal("Query sequence filtering not supported.");
        break;
      case 'G': o.gapopen = atol(optarg); break;
      case 'h':
        fprintf(out, "SWIPE-B200 %s\n\nScore-only Smith-Waterman database search on NVIDIA B200, "
                     "command-line compatible with SWIPE\n(T. Rognes (2011) BMC Bioinformatics, 12:221).\n\n",
                SWB_CLI_VERSION);
        usage(argv[0]);
        exit(1);
      case 'H': o.show_taxid = 1; break;
      case 'i': o.queryname = optarg; break;
      case 'I': o.show_gis = 1; break;
      case 'k': o.minexpect = atof(optarg); break;
      case 'K': break;
      case 'm': o.view = atol(optarg); break;
      case 'M': o.matrixname = optarg; break;
      case 'N': o.dump = atol(optarg); break;
      case 'o': o.outfile = optarg; break;
      case 'p':
        if (!strcmp(optarg, "blastn")) o.symtype = 0;
        else if (!strcmp(optarg, "blastp")) o.symtype = 1;
        else if (!strcmp(optarg, "blastx")) o.symtype = 2;
        else if (!strcmp(optarg, "tblastn")) o.symtype = 3;
        else if (!strcmp(optarg, "tblastx")) o.symtype = 4;
        else if (!strcmp(optarg, "sound")) o.symtype = 5;
        else o.symtype = atol(optarg);
        break;
      case 'q': o.mismatchscore = atol(optarg); break;
      case 'Q': o.query_gencode = atol(optarg); break;
      case 'r': o.matchscore = atol(optarg); break;
      case 'S':
        if (!strcmp(optarg, "plus")) o.querystrands = 1;
        else if (!strcmp(optarg, "minus")) o.querystrands = 2;
        else if (!strcmp(optarg, "both")) o.querystrands = 3;
        else o.querystrands = atol(optarg);
        break;
      case 'u': o.maxscore = atol(optarg); break;
      case 'v': o.maxmatches = atol(optarg); break;
      case 'x': o.taxidfile = optarg; break;
      case 'z': o.effdbsize = atol(optarg); break;
      default:
        usage(argv[0]);
        exit(1);
    }
  }
  if (o.outfile)
  {
    FILE *f = fopen(o.outfile, "w");
    if (!f) fatal("Unable to open output file for writing.");
    out = f;
  }
  // defaults that depend on the scoring system (swipe.cc:1089-1126)
  if (o.symtype == 0)
  {
    if (o.gapopen == 0) o.gapopen = 5;
    if (o.gapextend == 0) o.gapextend = 2;
  }
  else if (o.symtype < 5)
  {
    if (o.matrixname.empty()) o.matrixname = "BLOSUM62";
    int64_t go = 0, ge = 0;
    if (swb_stats_default_gaps(o.matrixname.c_str(), &go, &ge))
    {
      if (o.gapopen == 0) o.gapopen = go;
      if (o.gapextend == 0) o.gapextend = ge;
    }
    else if (o.gapopen == 0 && o.gapextend == 0)
      fatal("Unknown score matrix. Gap penalties must be specified (-G and -E).");
  }
  else if (o.symtype == 5)
  {
    if (o.matrixname.empty()) o.matrixname = "IDENTITY_5_1";
    if (o.gapopen == 0) o.gapopen = 15;
    if (o.gapextend == 0) o.gapextend = 5;
  }
  if (o.effdbsize < 0) fatal("Illegal effective db size specified");
  if (o.threads < 1 || o.threads > 256) fatal("Illegal number of threads specified");
  if (o.databasename.empty()) fatal("No database specified.");
  if (!(o.view == 0 || o.view == 7 || o.view == 8 || o.view == 9 || o.view == 99)) fatal("Illegal view type.");
  if (o.gapopen < 0 || o.gapextend < 0 || o.gapopen + o.gapextend < 1) fatal("Illegal gap penalties.");
  if (o.symtype < 0 || o.symtype > 5) fatal("Illegal symbol type.");
  if (o.querystrands < 1 || o.querystrands > 3) fatal("Illegal query strands specified.");
  if (o.querystrands == 2 && (o.symtype == 1 || o.symtype == 3 || o.symtype == 4))
    fatal("Illegal strand specified for protein query.");
  if (o.query_gencode < 1 || o.query_gencode > 23 || !swb_gencode_name((int)o.query_gencode))
    fatal("Illegal query genetic code specified.");
  if (o.db_gencode < 1 || o.db_gencode > 23 || !swb_gencode_name((int)o.db_gencode))
    fatal("Illegal database genetic code specified.");
  if (o.dump < 0 || o.dump > 2) fatal("Illegal dump mode.");
  return o;
}

// ---- the query and its variants -----------------------------------------------------------------
struct Query
{
  std::string description;
  std::vector<uint8_t> nt[2];          // forward / reverse complement (symtype 0, 2, 4)
  std::vector<uint8_t> aa[6];          // [3 * strand + frame]; aa[0] = the query for symtype 1, 3
};

struct Hit
{
  int64_t seqno = 0, score = 0;
  int qstrand = 0, qframe = 0, dstrand = 0, dframe = 0;
  int64_t bestq = -1, align_hint = -1;
  std::string header;                   // raw ASN.1 defline bytes
  std::vector<uint8_t> dseq;
  int64_t dlen = 0, dlennt = 0;
  int64_t aqs = 0, aqe = 0, ads = 0, ade = 0, score_align = 0;
  std::string ops;
};

// (score desc, seqno desc, then the order the reference's loops enter equal hits: hits.cc:188-191)
bool hit_before(const Hit &a, const Hit &b)
{
  if (a.score != b.score) return a.score > b.score;
  if (a.seqno != b.seqno) return a.seqno > b.seqno;
  if (a.qstrand != b.qstrand) return a.qstrand < b.qstrand;
  if (a.qframe != b.qframe) return a.qframe < b.qframe;
  if (a.dstrand != b.dstrand) return a.dstrand < b.dstrand;
  return a.dframe < b.dframe;
}

struct Shard
{
  int device = 0;
  int64_t first = 0, count = 0;         // source sequences of the BLAST database
  swb_db *db = nullptr;
  bool sink_filter = false;             // mask / -x list handed to the device sink (swb_db_set_filter)
  int64_t nincluded = 0;                // sequences of the shard that pass it
};

struct Run
{
  Options o;
  swb_blastdb *bdb = nullptr;
  int64_t nseq = 0, symcount = 0, longest = 0;
  int64_t memb_bit = 0, masked_nseq = 0, masked_symcount = 0;   // alias-file mask (database.cc:1046-1065)
  std::vector<uint8_t> taxids;          // -x: bitmap of the taxids to keep (database.cc:735-772)
  bool db_nt = false, translated_db = false;
  int64_t matrix[1024];
  uint8_t qtable[4096], dtable[4096];
  swb_stats st;
  std::vector<Shard> shards;
  std::vector<Hit> hits;
  int64_t keephits = 0;
  int64_t totalhits = 0, obvious = 0, compute7 = 0, queryno = 0;   // hits.cc:174-178, swipe.cc:111; totalhits is never reset
  std::string started, completed;
  double elapsed = 0, speed = 0;
  std::mutex mu;
};

const std::vector<uint8_t> &query_variant(const Run &R, const Query &q, int qstrand, int qframe)
{
  if (R.o.symtype == 0) return q.nt[qstrand];
  return q.aa[3 * qstrand + qframe];
}

void merge_hits(Run &R, std::vector<Hit> &local, int64_t tot, int64_t obv, int64_t computed)
{
  std::lock_guard<std::mutex> g(R.mu);
  R.totalhits += tot;
  R.obvious += obv;
  R.compute7 += computed;
  R.hits.insert(R.hits.end(), local.begin(), local.end());
  if ((int64_t)R.hits.size() > R.keephits)
  {
    std::sort(R.hits.begin(), R.hits.end(), hit_before);
    R.hits.resize((size_t)R.keephits);
  }
}

// db_check_inclusion (database.cc:1465-1481): the membership bit of a masked database, then the
// -x list: a sequence stays when at least one of its deflines passes both filters
bool included(const Run &R, int64_t seqno)
{
  if (!swb_blastdb_included(R.bdb, seqno)) return false;
  if (R.taxids.empty()) return true;
  const uint8_t *hp = nullptr;
  int64_t hl = 0;
  if (swb_blastdb_header(R.bdb, seqno, &hp, &hl) != SWB_OK) return false;
  char tmp[1];
  int64_t need = 0;
  const int64_t n = swb_defline_text(hp, hl, 0, 0, R.memb_bit, R.taxids.data(), (int64_t)R.taxids.size(), tmp, 0, &need);
  return n > 0 || (n == SWB_ERR_RANGE && need > 1);
}

// one GPU's share of search_chunk: every query strand / frame against every subject of the shard
void search_shard(Run &R, const Query &q, Shard &S)
{
  const Options &o = R.o;
  const int unit = R.translated_db ? 6 : 1;
  const int64_t nsub = S.count * unit;
  std::vector<int64_t> scores((size_t)std::max<int64_t>(nsub, 1));
  swb_scoring sc = {R.matrix, o.gapopen + o.gapextend, o.gapextend};
  const int qs1 = (o.symtype == 0 || o.symtype == 2 || o.symtype == 4) ? (o.querystrands == 2 ? 1 : 0) : 0;
  const int qs2 = (o.symtype == 0 || o.symtype == 2 || o.symtype == 4) ? (o.querystrands == 1 ? 0 : 1) : 0;
  const int qf2 = (o.symtype == 2 || o.symtype == 4) ? 2 : 0;
  std::vector<Hit> local;
  int64_t tot = 0, obv = 0, computed = 0;
  for (int qstrand = qs1; qstrand <= qs2; qstrand++)
    for (int qframe = 0; qframe <= qf2; qframe++)
    {
      const std::vector<uint8_t> &qv = query_variant(R, q, qstrand, qframe);
      const bool filtered = R.memb_bit != 0 || !R.taxids.empty();
      if ((!filtered || S.sink_filter) && unit == 1 && R.keephits > 0)
      {
        // subject = sequence, and every subject takes part or the device knows which do: the sink runs
        // on the device (swb_search_hits) and only the hits hits_enter would have kept come back
        std::vector<int64_t> hs((size_t)R.keephits), hv((size_t)R.keephits);
        int64_t nh = 0, t = 0, ob = 0;
        check(swb_search_hits(S.db, qv.data(), (int64_t)qv.size(), &sc, S.first, R.keephits, R.st.score_threshold,
                              R.st.upper_threshold, hs.data(), hv.data(), &nh, &t, &ob, nullptr), "search");
        computed += filtered ? (o.view == 99 ? S.nincluded : 0) : nsub;        // as the dense path below counts
        tot += t;
        obv += ob;
        for (int64_t k = 0; k < nh; k++)
        {
          Hit h;
          h.seqno = hs[(size_t)k];
          h.score = hv[(size_t)k];
          if (o.symtype == 0 && qstrand) { h.qstrand = 0; h.dstrand = 1; }     // swipe.cc:1470-1471
          else { h.qstrand = qstrand; h.qframe = qframe; h.dstrand = 0; h.dframe = 0; }
          local.push_back(h);
        }
        continue;
      }
      check(swb_search(S.db, qv.data(), (int64_t)qv.size(), &sc, scores.data(), nullptr), "search");
      int64_t threshold = R.st.score_threshold;
      if (!filtered) computed += nsub;
      for (int64_t j = 0; j < nsub; j++)
      {
        const int64_t s = scores[(size_t)j];
        if (filtered && o.view == 99 && included(R, S.first + j / unit)) computed++;   // only -m 99 reports it
        if (s < R.st.score_threshold) continue;                // below the initial threshold: not even counted
        const int64_t seqno = S.first + j / unit;
        if (!included(R, seqno)) continue;
        tot++;                                                 // hits_enter: score >= init_threshold (hits.cc:177)
        if (s > R.st.upper_threshold) { obv++; continue; }     // hits.cc:174
        if (s < threshold) continue;
        Hit h;
        h.seqno = seqno;
        h.score = s;
        if (o.symtype == 0 && qstrand) { h.qstrand = 0; h.dstrand = 1; }     // swipe.cc:1470-1471
        else
        {
          h.qstrand = qstrand; h.qframe = qframe;
          h.dstrand = unit == 6 ? (int)(j % 6) / 3 : 0;
          h.dframe = unit == 6 ? (int)(j % 3) : 0;
        }
        local.push_back(h);
        if ((int64_t)local.size() >= 4 * R.keephits + 1024)
        {
          std::sort(local.begin(), local.end(), hit_before);
          local.resize((size_t)R.keephits);
          threshold = std::max(threshold, local.back().score);   // hits_enter: the list is full (hits.cc:218-219)
        }
      }
    }
  std::sort(local.begin(), local.end(), hit_before);
  if ((int64_t)local.size() > R.keephits) local.resize((size_t)R.keephits);
  merge_hits(R, local, tot, obv, computed);
}

// the subject as the aligner sees it (hits_align, hits.cc:562-571): strand / frame applied
void fetch_subject(Run &R, Hit &h)
{
  const int64_t n = swb_blastdb_seqlen(R.bdb, h.seqno);
  std::vector<uint8_t> raw((size_t)std::max<int64_t>(n, 1));
  int64_t got = 0;
  const int strand = (R.o.symtype == 0) ? h.dstrand : 0;
  check(swb_blastdb_sequence(R.bdb, h.seqno, strand, raw.data(), n, &got), "reading a database sequence");
  if (R.translated_db)
  {
    h.dlennt = got;
    std::vector<uint8_t> prot((size_t)std::max<int64_t>(got / 3 + 1, 1));
    const int64_t plen = swb_translate(raw.data(), got, h.dstrand, h.dframe, R.dtable, prot.data());
    prot.resize((size_t)plen);
    h.dseq.swap(prot);
  }
  else
  {
    raw.resize((size_t)got);
    h.dseq.swap(raw);
  }
  h.dlen = (int64_t)h.dseq.size();
}

// align_chunk + hits_align: end cells from the GPU, traceback on the host
void align_hits(Run &R, const Query &q)
{
  const Options &o = R.o;
  const int64_t nalign = std::min<int64_t>((int64_t)R.hits.size(), o.alignments);
  int64_t lo = 0, hi = 0, l7 = 0, limit16 = 0;
  swb_matrix_limits(R.matrix, &lo, &hi, &l7, &limit16);
  swb_scoring sc = {R.matrix, o.gapopen + o.gapextend, o.gapextend};
  const int unit = R.translated_db ? 6 : 1;
  for (const Shard &S : R.shards)
    for (int qv = 0; qv < 6; qv++)
    {
      std::vector<int64_t> list, idx;
      for (int64_t i = 0; i < nalign; i++)
      {
        const Hit &h = R.hits[(size_t)i];
        if (3 * h.qstrand + h.qframe != qv || h.seqno < S.first || h.seqno >= S.first + S.count) continue;
        int64_t local = (h.seqno - S.first) * unit;
        int64_t code;
        if (unit == 6) code = (local + 3 * h.dstrand + h.dframe) << 3;
        else code = (local << 3) | ((int64_t)h.dstrand << 2);
        list.push_back(code);
        idx.push_back(i);
      }
      if (list.empty()) continue;
      const std::vector<uint8_t> &qq = query_variant(R, q, qv / 3, qv % 3);
      std::vector<int64_t> s(list.size()), bp(list.size()), bq(list.size());
      check(swb_search_end(S.db, qq.data(), (int64_t)qq.size(), &sc, list.data(), (int64_t)list.size(),
                           s.data(), bp.data(), bq.data()), "alignment end search");
      for (size_t k = 0; k < list.size(); k++)
        if (s[k] < limit16)                                   // swipe.cc:401-402
        {
          R.hits[(size_t)idx[k]].bestq = bq[k];
          R.hits[(size_t)idx[k]].align_hint = bp[k];
        }
    }
  for (size_t i = 0; i < R.hits.size(); i++)
  {
    Hit &h = R.hits[i];
    const uint8_t *hp = nullptr;
    int64_t hl = 0;
    check(swb_blastdb_header(R.bdb, h.seqno, &hp, &hl), "reading a database header");
    h.header.assign((const char *)hp, (size_t)hl);
    if ((int64_t)i >= o.alignments) continue;
    fetch_subject(R, h);
    const std::vector<uint8_t> &qq = (o.symtype == 0) ? q.nt[0] : q.aa[3 * h.qstrand + h.qframe];
    if (h.bestq > 0 && h.align_hint != 0)                     // hits.cc:589-600
    {
      h.score_align = h.score; h.aqe = h.bestq; h.ade = h.align_hint;
    }
    else
    {
      h.score_align = 0; h.aqe = 0; h.ade = 0;
    }
    std::vector<char> ops(16 * (qq.size() + h.dseq.size()) + 64);
    int64_t n = 0;
    const int rc = swb_align(qq.data(), (int64_t)qq.size(), h.dseq.data(), h.dlen, R.matrix, o.gapopen,
                             o.gapextend, &h.aqs, &h.ads, &h.aqe, &h.ade, &h.score_align, ops.data(),
                             (int64_t)ops.size(), &n);
    if (rc != SWB_OK) fatal("Internal error in align function.");
    h.ops.assign(ops.data(), (size_t)n);
  }
}

// ---- report -----------------------------------------------------------------------------------------
// show_deflines (asnparse.cc:889-971)
void show_header(const Run &R, const Hit &h, long show_gis, long indent, size_t maxlen, long linelen,
                 long maxdeflines, bool show_descr)
{
  int64_t need = 0;
  std::vector<char> buf(h.header.size() * 4 + 4096);
  const uint8_t *tx = R.taxids.empty() ? nullptr : R.taxids.data();
  int64_t n = swb_defline_text((const uint8_t *)h.header.data(), (int64_t)h.header.size(), (int)show_gis,
                               (int)R.o.show_taxid, R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(),
                               (int64_t)buf.size(), &need);
  if (n == SWB_ERR_RANGE)
  {
    buf.resize((size_t)need + 1);
    n = swb_defline_text((const uint8_t *)h.header.data(), (int64_t)h.header.size(), (int)show_gis,
                         (int)R.o.show_taxid, R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(),
                         (int64_t)buf.size(), &need);
  }
  if (n < 0) fatal("Error parsing binary ASN.1 in database sequence definition.");
  std::vector<std::string> lines;
  {
    std::string all(buf.data());
    size_t p = 0;
    for (int64_t k = 0; k < n; k++)
    {
      size_t e = all.find('\n', p);
      if (e == std::string::npos) e = all.size();
      lines.push_back(all.substr(p, e - p));
      p = e + 1;
    }
  }
  for (size_t x = 0; x < lines.size() && (long)x < maxdeflines; x++)
  {
    std::string d = lines[x];
    size_t show = d.size();
    if (maxlen && show > maxlen) show = maxlen;
    if (show < d.size() && show >= 3) d.replace(show - 3, 3, "...");
    size_t pos = 0;
    long line = 0;
    while (pos < show)
    {
      long col = 0;
      if (maxdeflines > 1)
      {
        if (line)
          while (col < 1 + indent) { putc(' ', out); col++; }
        else { putc(x ? ' ' : '>', out); col++; }
      }
      while (pos < show && col < linelen)
      {
        const char c = d[pos];
        if (!show_descr && c == ' ') pos = show;
        else { putc(c, out); pos++; col++; }
      }
      if (linelen < LONG_MAX)
        while (col < linelen) { putc(' ', out); col++; }
      if (maxdeflines > 1) putc('\n', out);
      line++;
    }
  }
}

struct AlignView
{
  long identities = 0, positives = 0, indels = 0, aligned = 0, gaps = 0;
  long q_first = 0, q_last = 0, d_first = 0, d_last = 0;
  int poswidth = 1;
  std::string qline, aline, dline, opline;       // one character per alignment column
};

// count_align / whole_align (hits.cc:815-1176)
AlignView view_alignment(const Run &R, const Query &q, const Hit &h, bool xml_marks)
{
  const Options &o = R.o;
  AlignView v;
  const char *sym = o.symtype == 0 ? SYM_NT : (o.symtype == 5 ? SYM_SOUND : SYM_AA);
  const std::vector<uint8_t> &qs = o.symtype == 0 ? q.nt[h.qstrand] : q.aa[3 * h.qstrand + h.qframe];
  int64_t qp = h.aqs, dp = h.ads;
  const char *p = h.ops.c_str();
  while (*p)
  {
    const char op = *p++;
    long len = 0;
    int used = 0;
    sscanf(p, "%ld%n", &len, &used);
    p += used;
    v.aligned += len;
    v.opline.append((size_t)len, op);
    for (long j = 0; j < len; j++)
    {
      if (op == 'D')
      {
        v.qline += sym[qs[(size_t)qp++]]; v.aline += ' '; v.dline += '-';
      }
      else if (op == 'I')
      {
        v.qline += '-'; v.aline += ' '; v.dline += sym[h.dseq[(size_t)dp++]];
      }
      else
      {
        const int a = qs[(size_t)qp++], b = h.dseq[(size_t)dp++];
        v.qline += sym[a];
        v.dline += sym[b];
        if (a == b)
        {
          v.identities++; v.positives++;
          v.aline += xml_marks || o.symtype == 0 ? '|' : sym[a];
        }
        else if (R.matrix[32 * a + b] > 0)
        {
          v.positives++;
          v.aline += o.symtype == 0 && !xml_marks ? ' ' : '+';
        }
        else
          v.aline += ' ';
      }
    }
    if (op != 'M') { v.gaps++; v.indels += len; }
  }
  long qf = h.aqs, ql = h.aqe, df = h.ads, dl = h.ade;
  const long qlen = (long)qs.size(), qlen_nt = (long)q.nt[0].size();
  if (o.symtype == 0)
  {
    if (h.qstrand) { qf = qlen - 1 - qf; ql = qlen - 1 - ql; }
    if (h.dstrand) { df = h.dlen - 1 - df; dl = h.dlen - 1 - dl; }
  }
  if (o.symtype == 2 || o.symtype == 4)
  {
    if (h.qstrand) { qf = qlen_nt - 1 - 3 * qf - h.qframe; ql = qlen_nt - 1 - 3 * ql - h.qframe - 2; }
    else { qf = 3 * qf + h.qframe; ql = 3 * ql + h.qframe + 2; }
  }
  if (o.symtype == 3 || o.symtype == 4)
  {
    if (h.dstrand) { df = h.dlennt - 1 - 3 * df - h.dframe; dl = h.dlennt - 1 - 3 * dl - h.dframe - 2; }
    else { df = 3 * df + h.dframe; dl = 3 * dl + h.dframe + 2; }
  }
  v.q_first = qf + 1; v.q_last = ql + 1; v.d_first = df + 1; v.d_last = dl + 1;
  long maxpos = std::max(std::max(v.q_first, v.q_last), std::max(v.d_first, v.d_last));
  while (maxpos > 9) { maxpos /= 10; v.poswidth++; }
  return v;
}

void show_expect(double e)
{
  char temp[32];
  if (e < 1e-180) fprintf(out, "0.0  ");
  else if (e < 9.5e-100) { snprintf(temp, sizeof temp, "%-6.0e", e); fputs(temp + 1, out); }
  else if (e < 0.00095) fprintf(out, "%-5.0e", e);
  else if (e < 0.0995) fprintf(out, "%-5.3f", e);
  else if (e < 0.95) fprintf(out, "%-5.2f", e);
  else if (e < 9.5) fprintf(out, "%-5.1f", e);
  else fprintf(out, "%5.0f", e);
}

// show_align / putalignop (hits.cc:647-813): 60 columns per block
void show_alignment_blocks(const Run &R, const Query &q, const Hit &h, const AlignView &v)
{
  const Options &o = R.o;
  const long qlen_nt = (long)q.nt[0].size();
  long qpos = h.aqs, dpos = h.ads;
  const size_t total = v.qline.size();
  for (size_t at = 0; at < total; at += 60)
  {
    const size_t n = std::min<size_t>(60, total - at);
    const long qstart = qpos, dstart = dpos;
    for (size_t k = 0; k < n; k++)
    {
      if (v.opline[at + k] != 'I') qpos++;
      if (v.opline[at + k] != 'D') dpos++;
    }
    long q1 = qstart + 1, q2 = qpos, d1 = dstart + 1, d2 = dpos;
    if (o.symtype == 0 && h.dstrand) { d1 = h.dlen - d1 + 1; d2 = h.dlen - d2 + 1; }
    if (o.symtype == 2 || o.symtype == 4)
    {
      if (h.qstrand) { q1 = qlen_nt - 3 * qstart - h.qframe; q2 = qlen_nt - 3 * qpos - h.qframe + 1; }
      else { q1 = 3 * qstart + h.qframe + 1; q2 = 3 * qpos + h.qframe; }
    }
    if (o.symtype == 3 || o.symtype == 4)
    {
      if (h.dstrand) { d1 = h.dlennt - 3 * dstart - h.dframe; d2 = h.dlennt - 3 * dpos - h.dframe + 1; }
      else { d1 = 3 * dstart + h.dframe + 1; d2 = 3 * dpos + h.dframe; }
    }
    fprintf(out, "\n");
    fprintf(out, "Query: %*ld %s %ld\n", v.poswidth, q1, v.qline.substr(at, n).c_str(), q2);
    fprintf(out, "       %*s %s\n", v.poswidth, "", v.aline.substr(at, n).c_str());
    fprintf(out, "Sbjct: %*ld %s %ld\n", v.poswidth, d1, v.dline.substr(at, n).c_str(), d2);
  }
}

void show_description_word(const std::string &d)
{
  for (char c : d)
  {
    if (c == ' ') break;
    putc(c, out);
  }
}

void report_plain(Run &R, const Query &q, long showalignments, long showhits)
{
  const Options &o = R.o;
  if (R.hits.empty())
  {
    fprintf(out, "\nNo hits.\n");
    return;
  }
  if (R.st.available)
  {
    fprintf(out, "                                                                 Score    E\n");
    fprintf(out, "Sequences producing significant alignments:                      (bits) Value\n\n");
  }
  else
    fprintf(out, "Sequences producing significant alignments:                         Score\n\n");
  for (long i = 0; i < showhits; i++)
  {
    const Hit &h = R.hits[(size_t)i];
    long headerlen = 67;
    if (o.symtype == 0) headerlen = 65;
    else if (o.symtype == 2 || o.symtype == 3) headerlen = 64;
    else if (o.symtype == 4) headerlen = 61;
    show_header(R, h, o.show_gis, 0, (size_t)headerlen, headerlen, 1, true);
    if (o.symtype == 0) fprintf(out, " %c", h.dstrand ? '-' : '+');
    else if (o.symtype == 2) fprintf(out, " %c%d", h.qstrand ? '-' : '+', h.qframe + 1);
    else if (o.symtype == 3) fprintf(out, " %c%d", h.dstrand ? '-' : '+', h.dframe + 1);
    else if (o.symtype == 4)
      fprintf(out, " %c%d/%c%d", h.qstrand ? '-' : '+', h.qframe + 1, h.dstrand ? '-' : '+', h.dframe + 1);
    if (R.st.available)
    {
      const long bits = (long)floor(swb_stats_bits(&R.st, h.score) + 0.5);
      fprintf(out, " %5ld", bits);
      fprintf(out, "   ");
      show_expect(swb_stats_evalue(&R.st, h.score));
    }
    else
      fprintf(out, " %5ld", (long)h.score);
    putc('\n', out);
  }
  for (long i = 0; i < showalignments; i++)
  {
    const Hit &h = R.hits[(size_t)i];
    fprintf(out, "\n");
    show_header(R, h, o.show_gis, 10, 0, 79, LONG_MAX, true);
    fprintf(out, "          Length = %ld\n", (long)((o.symtype == 3 || o.symtype == 4) ? h.dlennt : h.dlen));
    fprintf(out, "\n");
    if (R.st.available)
    {
      fprintf(out, " Score = %.1lf bits (%ld), Expect = ", swb_stats_bits(&R.st, h.score), (long)h.score);
      show_expect(swb_stats_evalue(&R.st, h.score));
    }
    else
      fprintf(out, " Score = %ld", (long)h.score);
    putc('\n', out);
    const AlignView v = view_alignment(R, q, h, false);
    fprintf(out, " Identities = %ld/%ld (%ld%%)", v.identities, v.aligned, v.identities * 100 / v.aligned);
    if (o.symtype > 0)
      fprintf(out, ", Positives = %ld/%ld (%ld%%)", v.positives, v.aligned, v.positives * 100 / v.aligned);
    if (v.indels) fprintf(out, ", Gaps = %ld/%ld (%ld%%)", v.indels, v.aligned, v.indels * 100 / v.aligned);
    fprintf(out, "\n");
    if (o.symtype == 0) fprintf(out, " Strand = %s\n", h.dstrand ? "Plus / Minus" : "Plus / Plus");
    else if (o.symtype == 2) fprintf(out, " Frame = %c%d\n", h.qstrand ? '-' : '+', h.qframe + 1);
    else if (o.symtype == 3) fprintf(out, " Frame = %c%d\n", h.dstrand ? '-' : '+', h.dframe + 1);
    else if (o.symtype == 4)
      fprintf(out, " Frame = %c%d / %c%d\n", h.qstrand ? '-' : '+', h.qframe + 1, h.dstrand ? '-' : '+', h.dframe + 1);
    show_alignment_blocks(R, q, h, v);
    fprintf(out, "\n");
  }
}

void report_xml(Run &R, const Query &q, long showalignments, long showhits)
{
  fprintf(out, "<result>\n  <general>\n    <hitcount>%d</hitcount>\n  </general>\n  <hits>\n", (int)R.hits.size());
  for (long i = 0; i < showhits; i++)
  {
    const Hit &h = R.hits[(size_t)i];
    fprintf(out, "    <hit>\n      <hitno>%ld</hitno>\n      <track>%ld</track>\n      <query>", i + 1, (long)h.seqno);
    show_description_word(q.description);
    fprintf(out, "</query>\n      <name>");
    show_header(R, h, R.o.show_gis, 0, 0, LONG_MAX, 1, true);
    fprintf(out, "</name>\n      <len>%ld</len>\n      <score>%ld</score>\n", (long)h.dlen, (long)h.score);
    if (i < showalignments)
    {
      const AlignView v = view_alignment(R, q, h, true);
      fprintf(out, "      <alignment>%s</alignment>\n", h.ops.c_str());
      fprintf(out, "      <qpos>%ld,%ld</qpos>\n      <dpos>%ld,%ld</dpos>\n", v.q_first, v.q_last, v.d_first, v.d_last);
      fprintf(out, "      <qseq>%s</qseq>\n      <aseq>%s</aseq>\n      <dseq>%s</dseq>\n", v.qline.c_str(),
              v.aline.c_str(), v.dline.c_str());
    }
    fprintf(out, "    </hit>\n");
  }
  fprintf(out, "  </hits>\n</result>\n");
}

// hits_show_xml_paralign (hits.cc:1215-1648)
std::vector<std::string> deflines_of(const Run &R, const Hit &h)
{
  std::vector<char> buf(h.header.size() * 4 + 4096);
  int64_t need = 0;
  const uint8_t *tx = R.taxids.empty() ? nullptr : R.taxids.data();
  int64_t n = swb_defline_text((const uint8_t *)h.header.data(), (int64_t)h.header.size(), 1, (int)R.o.show_taxid,
                               R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(), (int64_t)buf.size(), &need);
  if (n == SWB_ERR_RANGE)
  {
    buf.resize((size_t)need + 1);
    n = swb_defline_text((const uint8_t *)h.header.data(), (int64_t)h.header.size(), 1, (int)R.o.show_taxid,
                         R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(), (int64_t)buf.size(), &need);
  }
  if (n < 0) fatal("Error parsing binary ASN.1 in database sequence definition.");
  std::vector<std::string> lines;
  const std::string all(buf.data());
  size_t p = 0;
  for (int64_t k = 0; k < n; k++)
  {
    size_t e = all.find('\n', p);
    if (e == std::string::npos) e = all.size();
    lines.push_back(all.substr(p, e - p));
    p = e + 1;
  }
  return lines;
}

// "gi|N|" prefix, link = text up to the first blank, rest = the title (hits_defline_split, hits.cc:1256-1287)
void split_defline(const std::string &d, long *gi, std::string *link, std::string *rest)
{
  const char *p = d.c_str();
  int len = 0;
  *gi = 0;
  if (sscanf(p, "gi|%ld%n", gi, &len) > 0) p += len;
  if (*p == '|') p++;
  const char *r = strchr(p, ' ');
  link->clear();
  if (r) { link->assign(p, r - p); *rest = r + 1; }
  else *rest = p;
}

std::string anchor_of(const Run &R, const Hit &h)
{
  char a[200];
  const long qn = (long)R.queryno, s = (long)h.seqno;
  switch (R.o.symtype)
  {
    case 0: snprintf(a, sizeof a, "%ld_%ld__%c__+", qn, s, h.qstrand ? '-' : '+'); break;
    case 2: snprintf(a, sizeof a, "%ld_%ld_%d_%c__", qn, s, h.qframe + 1, h.qstrand ? '-' : '+'); break;
    case 3: snprintf(a, sizeof a, "%ld_%ld___%d_%c", qn, s, h.dframe + 1, h.dstrand ? '-' : '+'); break;
    case 4: snprintf(a, sizeof a, "%ld_%ld_%d_%c_%d_%c", qn, s, h.qframe + 1, h.qstrand ? '-' : '+', h.dframe + 1, h.dstrand ? '-' : '+'); break;
    default: snprintf(a, sizeof a, "%ld_%ld____", qn, s); break;
  }
  return a;
}

void report_paralign(Run &R, const Query &q, long showalignments, long showhits)
{
  const Options &o = R.o;
  const bool aa_query = o.symtype == 1 || o.symtype == 3;
  const std::vector<uint8_t> &qs = aa_query ? q.aa[0] : q.nt[0];
  const char *qsym = (o.symtype == 1 || o.symtype == 3) ? SYM_AA : (o.symtype == 5 ? SYM_SOUND : SYM_NT);
  const bool nt_db = o.symtype == 0 || o.symtype == 3 || o.symtype == 4;
  const char *ncbidb = nt_db ? "Nucleotide" : "Protein", *ncbiopt = nt_db ? "GenBank" : "GenPept";
  fprintf(out, "\t<paralignOutput>\n");
  fprintf(out, "\t\t<queryInformation>\n");
  fprintf(out, "\t\t\t<queryFilename>./%s</queryFilename>\n", o.queryname.c_str());
  fprintf(out, "\t\t\t<querySequencetype>%s</querySequencetype>\n", aa_query ? "Amino Acid" : "Nucleotide");
  fprintf(out, "\t\t\t<queryDescription>%s</queryDescription>\n", q.description.c_str());
  fprintf(out, "\t\t\t<queryLength>%ld</queryLength>\n", (long)qs.size());
  fprintf(out, "\t\t\t<querySequence>");
  for (uint8_t c : qs) putc(qsym[c & 31], out);
  fprintf(out, "</querySequence>\n\t\t</queryInformation>\n");
  fprintf(out, "\t\t<databaseInformation>\n");
  fprintf(out, "\t\t\t<databaseFilename>%s</databaseFilename>\n", o.databasename.c_str());
  fprintf(out, "\t\t\t<databaseSequencetype>%s</databaseSequencetype>\n", nt_db ? "Nucleotide" : "Amino Acid");
  fprintf(out, "\t\t\t<databaseDescription>%s</databaseDescription>\n", swb_blastdb_title(R.bdb));
  fprintf(out, "\t\t\t<databaseVersion>%ld</databaseVersion>\n", 4L);
  fprintf(out, "\t\t\t<databaseDate>%s</databaseDate>\n", swb_blastdb_date(R.bdb));
  fprintf(out, "\t\t\t<residueCount>%ld</residueCount>\n", (long)R.masked_symcount);
  fprintf(out, "\t\t\t<sequenceCount>%ld</sequenceCount>\n", (long)R.masked_nseq);
  fprintf(out, "\t\t\t<longestSequenceLength>%ld</longestSequenceLength>\n", (long)R.longest);
  fprintf(out, "\t\t</databaseInformation>\n");
  static const char *const strands[] = {"", "Plus", "Minus", "Both"};
  fprintf(out, "\t\t<options>\n\t\t\t<algorithm>Smith-Waterman</algorithm>\n");
  if (o.symtype == 0 || o.symtype == 2 || o.symtype == 4)
    fprintf(out, "\t\t\t<queryStrands>%s</queryStrands>\n", strands[o.querystrands]);
  if (o.symtype == 0) fprintf(out, "\t\t\t<scoreMatrix>NT</scoreMatrix>\n");
  else fprintf(out, "\t\t\t<scoreMatrix>%s</scoreMatrix>\n", o.matrixname.c_str());
  fprintf(out, "\t\t\t<gapPenalties>\n");
  fprintf(out, "\t\t\t\t<gapPenaltyOpen>%ld</gapPenaltyOpen>\n", o.gapopen);
  fprintf(out, "\t\t\t\t<gapPenaltyExtension>%ld</gapPenaltyExtension>\n", o.gapextend);
  for (const char *kind : {"ungapped", "gapped"})
  {
    fprintf(out, "\t\t\t\t<%s>\n", kind);
    fprintf(out, "\t\t\t\t\t<%sLambda>%.4g</%sLambda>\n", kind, R.st.lambda, kind);
    fprintf(out, "\t\t\t\t\t<%sKappa>%.4g</%sKappa>\n", kind, R.st.K, kind);
    fprintf(out, "\t\t\t\t\t<%sEta>%.4g</%sEta>\n", kind, R.st.H, kind);
    fprintf(out, "\t\t\t\t</%s>\n", kind);
  }
  fprintf(out, "\t\t\t</gapPenalties>\n");
  fprintf(out, "\t\t\t<expectRange>\n\t\t\t\t<expectRangeFrom>%.2g</expectRangeFrom>\n", o.minexpect);
  fprintf(out, "\t\t\t\t<expectRangeTo>%.2g</expectRangeTo>\n\t\t\t</expectRange>\n", o.expect);
  fprintf(out, "\t\t\t<displayLimits>\n\t\t\t\t<hitLimit>%ld</hitLimit>\n", o.maxmatches);
  fprintf(out, "\t\t\t\t<alignmentLimit>%ld</alignmentLimit>\n", o.alignments);
  fprintf(out, "\t\t\t\t<subalignmentLimit>%ld</subalignmentLimit>\n\t\t\t</displayLimits>\n", 1L);
  fprintf(out, "\t\t\t<threads>%ld</threads>\n\t\t</options>\n", o.threads);
  fprintf(out, "\t\t\t<searchInformation>\n");
  fprintf(out, "\t\t\t\t<searchStarted>%s</searchStarted>\n", R.started.c_str());
  fprintf(out, "\t\t\t\t<searchCompleted>%s</searchCompleted>\n", R.completed.c_str());
  fprintf(out, "\t\t\t\t<searchElapsedTime>%.2fs</searchElapsedTime>\n", R.elapsed);
  fprintf(out, "\t\t\t\t<searchSpeed>%.3f GCUPS</searchSpeed>\n", R.speed / 1e9);
  fprintf(out, "\t\t\t\t<searchSWAlignments>\n\t\t\t\t\t<SWAbsolute>%ld</SWAbsolute>\n", (long)R.compute7);
  fprintf(out, "\t\t\t\t\t<SWPercent>100</SWPercent>\n\t\t\t\t</searchSWAlignments>\n\t\t\t</searchInformation>\n");
  fprintf(out, "\t\t<resultInformation>\n\t\t\t<resultHits>\n");
  fprintf(out, "\t\t\t\t<totalCount>%ld</totalCount>\n", (long)R.totalhits);
  fprintf(out, "\t\t\t\t<obviousCount>%ld</obviousCount>\n", (long)R.obvious);
  fprintf(out, "\t\t\t\t<shownCount>%ld</shownCount>\n\t\t\t</resultHits>\n", showhits);
  fprintf(out, "\t\t\t<alignmentCount>%ld</alignmentCount>\n\t\t</resultInformation>\n", showalignments);
  auto links = [&](const char *ver, const char *tabs, long gi, const std::string &link) {
    if (gi)
    {
      fprintf(out, "%s<%sVersionLink>\n", tabs, ver);
      fprintf(out, "%s\t<%sVersionLinkDestination>http://www.ncbi.nlm.nih.gov/entrez/query.fcgi?cmd=Retrieve&amp;db=%s&amp;list_uids=%ld&amp;dopt=%s</%sVersionLinkDestination>\n", tabs, ver, ncbidb, gi, ncbiopt, ver);
      fprintf(out, "%s\t<%sVersionLinkText>gi|%ld</%sVersionLinkText>\n", tabs, ver, gi, ver);
      fprintf(out, "%s</%sVersionLink>\n", tabs, ver);
    }
    fprintf(out, "%s<%sVersionLink>\n", tabs, ver);
    fprintf(out, "%s\t<%sVersionLinkDestination>http://www.ncbi.nlm.nih.gov/entrez/query.fcgi?cmd=Search&amp;db=%s&amp;term=%s&amp;doptcmdl=%s</%sVersionLinkDestination>\n", tabs, ver, ncbidb, link.c_str(), ncbiopt, ver);
    fprintf(out, "%s\t<%sVersionLinkText>%s</%sVersionLinkText>\n", tabs, ver, link.c_str(), ver);
    fprintf(out, "%s</%sVersionLink>\n", tabs, ver);
  };
  fprintf(out, "\t\t<shortVersionHits>\n");
  for (long i = 0; i < showhits; i++)
  {
    const Hit &h = R.hits[(size_t)i];
    const std::vector<std::string> dl = deflines_of(R, h);
    long gi = 0;
    std::string link, title;
    split_defline(dl.empty() ? std::string() : dl[0], &gi, &link, &title);
    fprintf(out, "\t\t\t<shortVersionHit>\n\t\t\t\t<shortVersionAnchor>%s</shortVersionAnchor>\n", anchor_of(R, h).c_str());
    links("short", "\t\t\t\t", gi, link);
    fprintf(out, "\t\t\t\t<shortVersionName>%.35s</shortVersionName>\n", title.c_str());
    if (o.symtype == 0) fprintf(out, "\t\t\t\t<shortVersionStrand>%c</shortVersionStrand>\n", h.qstrand ? '-' : '+');
    else if (o.symtype == 2) fprintf(out, "\t\t\t\t<shortVersionFrame>%c%d</shortVersionFrame>\n", h.qstrand ? '-' : '+', h.qframe + 1);
    else if (o.symtype == 3) fprintf(out, "\t\t\t\t<shortVersionFrame>%c%d</shortVersionFrame>\n", h.dstrand ? '-' : '+', h.dframe + 1);
    else if (o.symtype == 4)
      fprintf(out, "\t\t\t\t<shortVersionFrame>%c%d/%c%d</shortVersionFrame>\n", h.qstrand ? '-' : '+', h.qframe + 1, h.dstrand ? '-' : '+', h.dframe + 1);
    fprintf(out, "\t\t\t\t<shortVersionScore>%ld</shortVersionScore>\n", (long)h.score);
    fprintf(out, "\t\t\t\t<shortVersionEValue>%.2g</shortVersionEValue>\n\t\t\t</shortVersionHit>\n", swb_stats_evalue(&R.st, h.score));
  }
  fprintf(out, "\t\t</shortVersionHits>\n");
  if (showalignments)
  {
    fprintf(out, "\t\t<longVersionHits>\n");
    for (long i = 0; i < showalignments; i++)
    {
      const Hit &h = R.hits[(size_t)i];
      fprintf(out, "\t\t\t<longVersionHit>\n\t\t\t\t<longVersionAnchor>%s</longVersionAnchor>\n", anchor_of(R, h).c_str());
      fprintf(out, "\t\t\t\t<linkContainer>\n");
      for (const std::string &d : deflines_of(R, h))
      {
        long gi = 0;
        std::string link, title;
        split_defline(d, &gi, &link, &title);
        links("long", "\t\t\t\t\t", gi, link);
        fprintf(out, "\t\t\t\t\t<longVersionName>%s</longVersionName>\n", title.c_str());
      }
      fprintf(out, "\t\t\t\t</linkContainer>\n");
      if (o.symtype == 0) fprintf(out, "\t\t\t\t<databaseSequenceLength>%ld nt</databaseSequenceLength>\n", (long)h.dlen);
      else if (o.symtype == 3 || o.symtype == 4) fprintf(out, "\t\t\t\t<databaseSequenceLength>%ld nt</databaseSequenceLength>\n", (long)h.dlennt);
      else fprintf(out, "\t\t\t\t<databaseSequenceLength>%ld aa</databaseSequenceLength>\n", (long)h.dlen);
      if (o.symtype == 0)
        fprintf(out, "\t\t\t\t<alignmentMatchLocation>%s</alignmentMatchLocation>\n",
                h.qstrand ? "Matches on complementary strands." : "Matches on same strands.");
      else if (o.symtype >= 2 && o.symtype <= 4)
      {
        fprintf(out, "\t\t\t\t<longVersionFrames>\n");
        if (o.symtype == 2 || o.symtype == 4)
          fprintf(out, "\t\t\t\t\t<longVersionQueryFrame>\n\t\t\t\t\t\t<queryStrand>%c</queryStrand>\n\t\t\t\t\t\t<queryFrame>%d</queryFrame>\n\t\t\t\t\t</longVersionQueryFrame>\n",
                  h.qstrand ? '-' : '+', h.qframe + 1);
        if (o.symtype == 3 || o.symtype == 4)
          fprintf(out, "\t\t\t\t\t<longVersionDatabaseFrame>\n\t\t\t\t\t\t<databaseStrand>%c</databaseStrand>\n\t\t\t\t\t\t<databaseFrame>%d</databaseFrame>\n\t\t\t\t\t</longVersionDatabaseFrame>\n",
                  h.dstrand ? '-' : '+', h.dframe + 1);
        fprintf(out, "\t\t\t\t</longVersionFrames>\n");
      }
      const AlignView v = view_alignment(R, q, h, true);
      fprintf(out, "\t\t\t\t<alignment>\n\t\t\t\t\t<subalignment>\n");
      fprintf(out, "\t\t\t\t\t\t<longVersionScore>%ld</longVersionScore>\n", (long)h.score);
      fprintf(out, "\t\t\t\t\t\t<longVersionEValue>%.2g</longVersionEValue>\n", swb_stats_evalue(&R.st, h.score));
      fprintf(out, "\t\t\t\t\t\t<identical>\n\t\t\t\t\t\t\t<identicalNominator>%ld</identicalNominator>\n\t\t\t\t\t\t\t<identicalDenominator>%ld</identicalDenominator>\n\t\t\t\t\t\t\t<identicalPercentage>%.1f</identicalPercentage>\n\t\t\t\t\t\t</identical>\n",
              v.identities, v.aligned, 100.0 * v.identities / v.aligned);
      if (o.symtype != 0)
        fprintf(out, "\t\t\t\t\t\t<positive>\n\t\t\t\t\t\t\t<positiveNominator>%ld</positiveNominator>\n\t\t\t\t\t\t\t<positiveDenominator>%ld</positiveDenominator>\n\t\t\t\t\t\t\t<positivePercentage>%.1f</positivePercentage>\n\t\t\t\t\t\t</positive>\n",
                v.positives, v.aligned, 100.0 * v.positives / v.aligned);
      fprintf(out, "\t\t\t\t\t\t<indels>\n\t\t\t\t\t\t\t<indelsNominator>%ld</indelsNominator>\n\t\t\t\t\t\t\t<indelsDenominator>%ld</indelsDenominator>\n\t\t\t\t\t\t\t<indelsPercentage>%.1f</indelsPercentage>\n\t\t\t\t\t\t</indels>\n",
              v.indels, v.aligned, 100.0 * v.indels / v.aligned);
      fprintf(out, "\t\t\t\t\t\t<gaps>%ld</gaps>\n", v.gaps);
      fprintf(out, "\t\t\t\t\t\t<alignmentQuery>\n\t\t\t\t\t\t\t<alignmentQueryStart>%ld</alignmentQueryStart>\n\t\t\t\t\t\t\t<alignmentQueryLine>%s</alignmentQueryLine>\n\t\t\t\t\t\t\t<alignmentQueryEnd>%ld</alignmentQueryEnd>\n\t\t\t\t\t\t</alignmentQuery>\n",
              v.q_first, v.qline.c_str(), v.q_last);
      fprintf(out, "\t\t\t\t\t\t<alignmentLine>%s</alignmentLine>\n", v.aline.c_str());
      fprintf(out, "\t\t\t\t\t\t<alignmentDatabase>\n\t\t\t\t\t\t\t<alignmentDatabaseStart>%ld</alignmentDatabaseStart>\n\t\t\t\t\t\t\t<alignmentDatabaseLine>%s</alignmentDatabaseLine>\n\t\t\t\t\t\t\t<alignmentDatabaseEnd>%ld</alignmentDatabaseEnd>\n\t\t\t\t\t\t</alignmentDatabase>\n",
              v.d_first, v.dline.c_str(), v.d_last);
      fprintf(out, "\t\t\t\t\t</subalignment>\n\t\t\t\t</alignment>\n\t\t\t</longVersionHit>\n");
    }
    fprintf(out, "\t\t</longVersionHits>\n");
  }
  fprintf(out, "\t</paralignOutput>\n");
}

void report_tsv(Run &R, const Query &q, long showalignments, bool comments)
{
  if (comments)
  {
    fprintf(out, "# SWIPE-B200 %s - score-only Smith-Waterman scan on NVIDIA B200, SWIPE-compatible output "
                 "(T. Rognes (2011) BMC Bioinformatics, 12:221).\n", SWB_CLI_VERSION);
    fprintf(out, "# Query: %s\n", q.description.c_str());
    fprintf(out, "# Database: %s\n", R.o.databasename.c_str());
    if (R.st.available)
      fprintf(out, "# Fields: Query id, Subject id, %% identity, alignment length, mismatches, gap openings, q. start, q. end, s. start, s. end, e-value, bit score\n");
    else
      fprintf(out, "# Fields: Query id, Subject id, %% identity, alignment length, mismatches, gap openings, q. start, q. end, s. start, s. end, score\n");
  }
  for (long i = 0; i < showalignments; i++)
  {
    const Hit &h = R.hits[(size_t)i];
    show_description_word(q.description);
    putc('\t', out);
    show_header(R, h, 1, 0, 0, LONG_MAX, 1, false);
    const AlignView v = view_alignment(R, q, h, false);
    fprintf(out, "\t%.2f\t%ld\t%ld\t%ld\t%ld\t%ld\t%ld\t%ld", 100.0 * v.identities / v.aligned, v.aligned,
            v.aligned - v.identities - v.indels, v.gaps, v.q_first, v.q_last, v.d_first, v.d_last);
    if (R.st.available)
      fprintf(out, "\t%.2g\t%.1f", swb_stats_evalue(&R.st, h.score), swb_stats_bits(&R.st, h.score));
    else
      fprintf(out, "\t%ld", (long)h.score);
    fprintf(out, "\n");
  }
}

void show_run_header(const Run &R, const Query &q)
{
  const Options &o = R.o;
  static const char *const symtypes[] = {"Nucleotide", "Amino acid", "Translated query", "Translated database",
                                         "Both translated", "Sound"};
  fprintf(out, "Database file:     %s\n", o.databasename.c_str());
  fprintf(out, "Database title:    %s\n", swb_blastdb_title(R.bdb));
  fprintf(out, "Database time:     %s\n", swb_blastdb_date(R.bdb));
  fprintf(out, "Database size:     %ld residues in %ld sequences\n", (long)R.masked_symcount, (long)R.masked_nseq);
  fprintf(out, "Longest db seq:    %ld residues\n", (long)R.longest);
  if (o.effdbsize > 0) fprintf(out, "Effecive db size:  %ld\n", o.effdbsize);
  fprintf(out, "Query file name:   %s\n", o.queryname.c_str());
  const long qlen = (o.symtype == 0 || o.symtype == 2 || o.symtype == 4) ? (long)q.nt[0].size() : (long)q.aa[0].size();
  fprintf(out, "Query length:      %ld residues\n", qlen);
  for (size_t i = 0; i < q.description.size(); i += 60)
    fprintf(out, "%s%-60.60s\n", i ? "                   " : "Query description: ", q.description.c_str() + i);
  if (o.symtype == 0)
  {
    static const char *const strands[] = {"", "Plus", "Minus", "Plus and minus"};
    fprintf(out, "Query strands:     %s\n", strands[o.querystrands]);
    fprintf(out, "Score matrix:      %ld/%ld\n", o.matchscore, o.mismatchscore);
  }
  else
    fprintf(out, "Score matrix:      %s\n", o.matrixname.c_str());
  fprintf(out, "Gap penalty:       %ld+%ldk\n", o.gapopen, o.gapextend);
  fprintf(out, "Max expect shown:  %-g\n", o.expect);
  fprintf(out, "Min score shown:   %ld\n", o.minscore);
  fprintf(out, "Max matches shown: %ld\n", o.maxmatches);
  fprintf(out, "Alignments shown:  %ld\n", o.alignments);
  fprintf(out, "Show gi's:         %ld\n", o.show_gis);
  fprintf(out, "Show taxid's:      %ld\n", o.show_taxid);
  fprintf(out, "Threads:           %ld\n", o.threads);
  fprintf(out, "Symbol type:       %s\n", symtypes[o.symtype]);
  if (o.symtype == 2 || o.symtype == 4)
    fprintf(out, "Query genetic code:%s (%ld)\n", swb_gencode_name((int)o.query_gencode), o.query_gencode);
  if (o.symtype == 3 || o.symtype == 4)
    fprintf(out, "DB genetic code:   %s (%ld)\n", swb_gencode_name((int)o.db_gencode), o.db_gencode);
  if (o.taxidfile) fprintf(out, "Taxid filename:    %s\n", o.taxidfile);
  fprintf(out, "\n");
}

void build_query(Run &R, Query &q, std::vector<uint8_t> &seq)
{
  const Options &o = R.o;
  for (auto &v : q.nt) v.clear();
  for (auto &v : q.aa) v.clear();
  if (o.symtype == 0 || o.symtype == 2 || o.symtype == 4)
  {
    q.nt[0] = seq;
    if (o.querystrands & 2)
    {
      q.nt[1].resize(seq.size());
      swb_revcomp(seq.data(), (int64_t)seq.size(), q.nt[1].data());
    }
    if (o.symtype != 0)
      for (int s = 0; s < 2; s++)
        if ((s + 1) & o.querystrands)
          for (int f = 0; f < 3; f++)
          {
            q.aa[3 * s + f].resize(seq.size() / 3 + 1);
            const int64_t n = swb_translate(seq.data(), (int64_t)seq.size(), s, f, R.qtable, q.aa[3 * s + f].data());
            q.aa[3 * s + f].resize((size_t)std::max<int64_t>(n, 0));
          }
  }
  else
    q.aa[0] = seq;
}

void work(Run &R, Query &q)
{
  const Options &o = R.o;
  if (o.view == 0) show_run_header(R, q);
  // hits_init (hits.cc:283-511)
  R.keephits = std::max(o.maxmatches, o.alignments);
  int64_t maxhits = R.masked_nseq;
  if (o.symtype == 0) maxhits *= o.querystrands == 3 ? 2 : 1;
  else if (o.symtype == 2) maxhits *= o.querystrands == 3 ? 6 : 3;
  else if (o.symtype == 3) maxhits *= 6;
  else if (o.symtype == 4) maxhits *= o.querystrands == 3 ? 36 : 18;
  R.keephits = std::min(R.keephits, maxhits);
  const int64_t qlen = (o.symtype == 0 || o.symtype == 2 || o.symtype == 4) ? (int64_t)q.nt[0].size() : (int64_t)q.aa[0].size();
  if (o.symtype == 5)
  {
    memset(&R.st, 0, sizeof R.st);                    // no statistics for the sound alphabet (hits.cc:404)
    R.st.score_threshold = o.minscore;
    R.st.upper_threshold = o.maxscore;
  }
  else
    check(swb_stats_init((int)o.symtype, o.matrixname.c_str(), o.matchscore, o.mismatchscore, o.gapopen,
                         o.gapextend, qlen, R.masked_symcount, R.masked_nseq, o.effdbsize, o.minscore, o.maxscore,
                         o.expect, o.minexpect, &R.st), "statistics");
  if (!R.st.available && o.view == 0)
    fprintf(out, "Statistical parameters are not available for the scoring system specified.\nBit scores and E-values will not be computed.\n\n");
  R.hits.clear();
  R.obvious = 0;
  R.compute7 = 0;
  if (o.view == 0)
  {
    fprintf(out, "Searching...");
    fflush(out);
  }
  struct tms t1, t2;
  const time_t w1 = time(nullptr);
  const clock_t c1 = times(&t1);
  {
    std::vector<std::thread> pool;
    for (Shard &S : R.shards) pool.emplace_back([&R, &q, &S]() { search_shard(R, q, S); });
    for (std::thread &t : pool) t.join();
  }
  std::sort(R.hits.begin(), R.hits.end(), hit_before);
  if ((int64_t)R.hits.size() > R.keephits) R.hits.resize((size_t)R.keephits);
  const clock_t c2 = times(&t2);
  const time_t w2 = time(nullptr);
  {
    char b1[40], b2[40];
    struct tm tmv;
    gmtime_r(&w1, &tmv); strftime(b1, sizeof b1, "%a, %e %b %Y %T UTC", &tmv);
    gmtime_r(&w2, &tmv); strftime(b2, sizeof b2, "%a, %e %b %Y %T UTC", &tmv);
    const double elapsed = (double)(c2 - c1) / (double)sysconf(_SC_CLK_TCK);
    double speed = (double)R.masked_symcount;
    if (o.symtype == 0) speed *= (double)q.nt[0].size() * (o.querystrands == 3 ? 2 : 1);
    else if (o.symtype == 1) speed *= (double)q.aa[0].size();
    else if (o.symtype == 2) speed *= (double)q.nt[0].size() * (o.querystrands == 3 ? 2 : 1);
    else if (o.symtype == 3) speed *= 2.0 * (double)q.aa[0].size();
    else speed *= 2.0 * (double)q.nt[0].size() * (o.querystrands == 3 ? 2 : 1);
    R.started = b1; R.completed = b2; R.elapsed = elapsed; R.speed = speed / elapsed;
    if (o.view == 0)
    {
      fprintf(out, "...............................................done\n\n");
      fprintf(out, "Search started:    %s\n", b1);
      fprintf(out, "Search completed:  %s\n", b2);
      fprintf(out, "Elapsed:           %.2fs\n", elapsed);
      fprintf(out, "Speed:             %.3f GCUPS\n", speed / elapsed / 1e9);
      fprintf(out, "\n");
    }
  }
  align_hits(R, q);
  const long showhits = (long)std::min<int64_t>((int64_t)R.hits.size(), o.maxmatches);
  const long showalignments = (long)std::min<int64_t>((int64_t)R.hits.size(), o.alignments);
  if (o.view == 0) report_plain(R, q, showalignments, showhits);
  else if (o.view == 7) report_xml(R, q, showalignments, showhits);
  else if (o.view == 99) report_paralign(R, q, showalignments, showhits);
  else report_tsv(R, q, showalignments, o.view == 9);
  R.queryno++;
}

}  // namespace

int main(int argc, char **argv)
{
  Run R;
  R.o = parse_args(argc, argv);
  const Options &o = R.o;
  R.db_nt = o.symtype == 0 || o.symtype == 3 || o.symtype == 4;
  R.translated_db = o.symtype == 3 || o.symtype == 4;
  check(swb_blastdb_open(o.databasename.c_str(), R.db_nt ? 1 : 0, &R.bdb), "database");
  int vols = 0;
  swb_blastdb_info(R.bdb, &R.nseq, &R.symcount, &R.longest, &vols);
  swb_blastdb_masked_info(R.bdb, &R.memb_bit, &R.masked_nseq, &R.masked_symcount);
  if (o.taxidfile)
  {
    FILE *f = fopen(o.taxidfile, "r");
    if (!f) fatal("Unable to open taxid file %s.", o.taxidfile);
    R.taxids.assign(64 * 1024, 0);
    unsigned long t;
    while (fscanf(f, "%lu\n", &t) > 0)
    {
      if (t / 8 >= R.taxids.size()) R.taxids.resize(t / 8 + 1, 0);
      R.taxids[t / 8] |= (uint8_t)(1u << (t & 7));
    }
    fclose(f);
  }
  if (o.dump)
  {
    // db_show_fasta (database.cc:1483-1537) for every sequence; needs no GPU
    const char *sym = R.db_nt ? "-ACMGRSVTWYHKDBN################" : (o.symtype == 5 ? SYM_SOUND : SYM_AA);
    std::vector<uint8_t> seq;
    std::vector<char> buf(1 << 16);
    for (int64_t s = 0; s < R.nseq; s++)
    {
      const uint8_t *hp = nullptr;
      int64_t hl = 0, need = 0;
      check(swb_blastdb_header(R.bdb, s, &hp, &hl), "reading a database header");
      const uint8_t *tx = R.taxids.empty() ? nullptr : R.taxids.data();
      int64_t n = swb_defline_text(hp, hl, 1, (int)o.show_taxid, R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(),
                                   (int64_t)buf.size(), &need);
      if (n == SWB_ERR_RANGE)
      {
        buf.resize((size_t)need + 1);
        n = swb_defline_text(hp, hl, 1, (int)o.show_taxid, R.memb_bit, tx, (int64_t)R.taxids.size(), buf.data(),
                             (int64_t)buf.size(), &need);
      }
      if (n < 0) fatal("Error parsing binary ASN.1 in database sequence definition.");
      if (n == 0) continue;
      const int64_t len = swb_blastdb_seqlen(R.bdb, s);
      seq.resize((size_t)std::max<int64_t>(len, 1));
      int64_t got = 0;
      check(swb_blastdb_sequence(R.bdb, s, 0, seq.data(), len, &got), "reading a database sequence");
      auto print_seq = [&]() {
        for (int64_t i = 0; i < got; i += 80)
        {
          for (int64_t k = i; k < std::min<int64_t>(got, i + 80); k++) putc(sym[seq[(size_t)k] & 31], out);
          fprintf(out, "\n");
        }
      };
      const char *p = buf.data();
      for (int64_t i = 0; i < n; i++)
      {
        const char *e = strchr(p, '\n');
        std::string line = e ? std::string(p, e - p) : std::string(p);
        p = e ? e + 1 : p + line.size();
        if (o.dump == 2)
        {
          fprintf(out, ">%s\n", line.c_str());
          print_seq();
        }
        else
        {
          if (i) fprintf(out, " ");
          fprintf(out, ">%s", line.c_str());
          if (i == n - 1)
          {
            fprintf(out, "\n");
            print_seq();
          }
        }
      }
    }
    swb_blastdb_close(R.bdb);
    if (o.outfile) fclose(out);
    return 0;
  }

  if (o.symtype == 0) swb_matrix_nucleotide(o.matchscore, o.mismatchscore, R.matrix);
  else
  {
    const int rc = o.symtype == 5 ? swb_matrix_read_sound(o.matrixname.c_str(), R.matrix)
                                  : swb_matrix_read(o.matrixname.c_str(), R.matrix);
    if (rc == SWB_ERR_IO) fatal("Cannot open score matrix file.");
    if (rc != SWB_OK) fatal("Problem parsing score matrix file.");
  }
  swb_translate_table((int)o.query_gencode, R.qtable);
  swb_translate_table((int)o.db_gencode, R.dtable);

  // the database: one shard per GPU, cut by sequence number into equal residue shares
  int ndev = 0;
  check(swb_device_count(&ndev), "CUDA");
  const int ngpu = (int)std::min<long>(o.threads, ndev);
  {
    std::vector<int64_t> cum((size_t)R.nseq + 1, 0);
    for (int64_t s = 0; s < R.nseq; s++) cum[(size_t)s + 1] = cum[(size_t)s] + swb_blastdb_seqlen(R.bdb, s);
    int64_t first = 0;
    for (int g = 0; g < ngpu; g++)
    {
      const int64_t target = cum[(size_t)R.nseq] * (g + 1) / ngpu;
      int64_t last = g + 1 == ngpu ? R.nseq
                                   : (int64_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
      last = std::max(first, std::min(last, R.nseq));
      Shard S;
      S.device = g; S.first = first; S.count = last - first;
      R.shards.push_back(S);
      first = last;
    }
  }
  {
    std::vector<std::thread> pool;
    std::vector<int> rcs(R.shards.size(), 0);
    for (size_t g = 0; g < R.shards.size(); g++)
      pool.emplace_back([&R, &rcs, g]() {
        Shard &S = R.shards[g];
        rcs[g] = R.translated_db
                     ? swb_db_open_blast_translated(S.device, R.bdb, S.first, S.count, R.dtable, 0, nullptr, &S.db)
                     : swb_db_open_blast(S.device, R.bdb, S.first, S.count, 0, nullptr, &S.db);
        if (rcs[g] == SWB_OK && !R.translated_db && (R.memb_bit != 0 || !R.taxids.empty()))
        {
          // db_check_inclusion once per sequence (swipe.cc:1373-1376), as one bit each for the device sink
          std::vector<uint8_t> bits((size_t)(S.count + 7) / 8 + 1, 0);
          for (int64_t j = 0; j < S.count; j++)
            if (included(R, S.first + j))
            {
              bits[(size_t)(j >> 3)] |= (uint8_t)(1u << (j & 7));
              S.nincluded++;
            }
          rcs[g] = swb_db_set_filter(S.db, bits.data());
          S.sink_filter = rcs[g] == SWB_OK;
        }
      });
    for (std::thread &t : pool) t.join();
    for (int rc : rcs) check(rc, "uploading the database");
  }

  // queries
  std::string text;
  {
    FILE *f = o.queryname == "-" ? stdin : fopen(o.queryname.c_str(), "r");
    if (!f) fatal("Cannot open query file.");
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    if (f != stdin) fclose(f);
  }
  if (o.view == 0)
    fprintf(out, "SWIPE-B200 %s\n\nScore-only Smith-Waterman database search on NVIDIA B200, command-line compatible with\n"
                 "SWIPE: T. Rognes (2011) BMC Bioinformatics, 12:221.\n\n", SWB_CLI_VERSION);
  else if (o.view == 7)
    fprintf(out, "<?xml version=\"1.0\"?>\n");
  else if (o.view == 99)
  {
    fprintf(out, "<?xml version=\"1.0\"?>\n");
    fprintf(out, "<ParalignXML xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\" xsi:noNamespaceSchemaLocation=\"http://www.paralign.org/ParalignXML.xsd\">\n");
    fprintf(out, "\t<programInformation>\n\t\t<programName>swipe</programName>\n");
    fprintf(out, "\t\t<programVersion>SWIPE-B200 %s</programVersion>\n", SWB_CLI_VERSION);
    fprintf(out, "\t\t<programDescription>Smith-Waterman database searches with inter-sequence SIMD parallelisation</programDescription>\n");
    fprintf(out, "\t\t<articleReferences>\n\t\t\t<reference>T. Rognes (2011) Faster Smith-Waterman database searches with inter-sequence SIMD parallelisation, BMC Bioinformatics, 12:221.</reference>\n\t\t</articleReferences>\n");
    fprintf(out, "\t\t<license>SWIPE is available under the GNU Affero General Public License, version 3</license>\n");
    fprintf(out, "\t</programInformation>\n");
  }
  const bool nt_query = o.symtype == 0 || o.symtype == 2 || o.symtype == 4;
  size_t at = 0;
  while (at < text.size())
  {
    std::vector<uint8_t> seq(text.size() - at + 1);
    std::vector<char> descr(text.size() - at + 2);
    int64_t n = 0;
    const int64_t used = swb_query_parse(text.data() + at, (int64_t)(text.size() - at),
                                         nt_query ? 1 : (o.symtype == 5 ? 2 : 0), seq.data(),
                                         (int64_t)seq.size(), &n, descr.data(), (int64_t)descr.size());
    if (used <= 0) break;
    at += (size_t)used;
    seq.resize((size_t)n);
    Query q;
    q.description = descr.data();
    build_query(R, q, seq);
    work(R, q);
  }
  if (o.view == 99) fprintf(out, "</ParalignXML>\n");
  for (Shard &S : R.shards) swb_db_close(S.db);
  swb_blastdb_close(R.bdb);
  if (o.outfile) fclose(out);
  return 0;
}
