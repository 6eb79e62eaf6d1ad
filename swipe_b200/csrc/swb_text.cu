// swb_text.cu -- host-side text handling around the scan: query FASTA parsing, symbol maps,
// reverse complement, codon translation tables and the ASN.1 sequence deflines of BLAST databases.
//
// Takes over from the reference (torognes/swipe):
//   map_ncbi_aa / map_ncbi_nt16 / ntcompl / sym_* ...... query.cc:49-128, :174-178
//   query_read ........................................... query.cc:244-355   -> swb_query_parse
//   translate_createtable / translate ................... query.cc:366-506   -> swb_translate_table, swb_translate
//   parse_blast_def_line_set & friends .................. asnparse.cc:94-1014 -> swb_defline_text
#include "../../include/swipe_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{

const char SYM_AA[] = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ";
const char SYM_NT16[] = "-ACMGRSVTWYHKDBN";

int aa_of(int c)
{
  if (c >= 'a' && c <= 'z') c -= 32;
  if (c == 0) return -1;
  const char *p = strchr(SYM_AA, c);
  return p ? (int)(p - SYM_AA) : -1;
}

int nt_of(int c)
{
  if (c >= 'a' && c <= 'z') c -= 32;
  if (c == 'U') c = 'T';
  if (c == '-' || c == 0) return -1;                 // '-' is not a nucleotide query symbol
  const char *p = strchr(SYM_NT16, c);
  return p ? (int)(p - SYM_NT16) : -1;
}

// the "sound" alphabet of -p 5 (query.cc:31-49, :179): A-Z = 1..26, a-e = 27..31, case sensitive
int sound_of(int c)
{
  if (c >= 'A' && c <= 'Z') return c - 'A' + 1;
  if (c >= 'a' && c <= 'e') return c - 'a' + 27;
  return -1;
}

inline int nt_complement(int c)                      // bit-reversed 4-bit code (query.cc:112)
{
  return ((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3);
}

// NCBI genetic codes 1..23 in TCAG order (the published translation tables; gaps in the numbering
// are codes NCBI never assigned)
const char *const GENCODE[23] = {
    "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG",
    "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    nullptr,
    nullptr,
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG",
    "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    nullptr,
    nullptr,
    nullptr,
    nullptr,
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"};

// ---- BER walking ------------------------------------------------------------------------------
// Deflines are BER with indefinite lengths on constructed values (terminated by 00 00) and short
// definite lengths on primitives; long-form definite lengths are accepted too.
struct Ber
{
  const uint8_t *p, *end;
  bool ok = true;
  bool at_end() const { return p >= end; }
  bool eoc() const { return p + 2 <= end && p[0] == 0 && p[1] == 0; }
  int peek() const { return p < end ? *p : -1; }
  // reads a tag + length; returns the tag, sets len (-1 = indefinite)
  int open(long long *len)
  {
    if (p + 2 > end) { ok = false; return -1; }
    const int tag = *p++;
    int l = *p++;
    if (l == 0x80) *len = -1;
    else if (l < 0x80) *len = l;
    else
    {
      int n = l & 0x7f;
      long long v = 0;
      if (n > 8 || p + n > end) { ok = false; return -1; }
      while (n--) v = (v << 8) | *p++;
      *len = v;
    }
    return tag;
  }
  void close_indef()
  {
    if (eoc()) p += 2;
    else ok = false;
  }
};

unsigned long long ber_uint(Ber &b)
{
  long long len;
  if (b.open(&len) != 0x02 || len < 0 || b.p + len > b.end) { b.ok = false; return 0; }
  unsigned long long v = 0;
  for (long long i = 0; i < len; i++) v = (v << 8) | *b.p++;
  return v;
}

std::string ber_string(Ber &b)
{
  long long len;
  const int tag = b.open(&len);
  if (tag != 0x1a || len < 0 || b.p + len > b.end) { b.ok = false; return std::string(); }
  std::string s((const char *)b.p, (size_t)len);
  b.p += len;
  return s;
}

// skips one complete value of any shape
void ber_skip(Ber &b)
{
  long long len;
  const int tag = b.open(&len);
  if (!b.ok) return;
  (void)tag;
  if (len >= 0)
  {
    if (b.p + len > b.end) { b.ok = false; return; }
    b.p += len;
    return;
  }
  while (b.ok && !b.at_end() && !b.eoc()) ber_skip(b);
  b.close_indef();
}

// runs body() on the content of a constructed value with the given tag
template <typename F> bool ber_within(Ber &b, int tag, F body)
{
  long long len;
  if (b.peek() != tag) return false;
  b.open(&len);
  if (len >= 0)
  {
    Ber in{b.p, b.p + len};
    if (in.end > b.end) { b.ok = false; return true; }
    body(in);
    b.ok = b.ok && in.ok;
    b.p = in.end;
  }
  else
  {
    Ber in{b.p, b.end};
    body(in);
    b.ok = b.ok && in.ok;
    b.p = in.p;
    b.close_indef();
  }
  return true;
}

bool more(const Ber &b) { return b.ok && !b.at_end() && !b.eoc(); }

struct TextId { std::string name, accession, release; unsigned long long version = 0; };

// Object-id ::= CHOICE { id [0] INTEGER, str [1] VisibleString } (asnparse.cc:253-273)
void object_id(Ber &b, std::string &str, unsigned long long &num)
{
  str.clear();
  num = 0;
  if (b.peek() == 0xa0) ber_within(b, 0xa0, [&](Ber &v) { num = ber_uint(v); });
  else if (b.peek() == 0xa1) ber_within(b, 0xa1, [&](Ber &v) { str = ber_string(v); });
  else if (b.peek() == 0x02) num = ber_uint(b);
  else if (b.peek() == 0x1a) str = ber_string(b);
  else if (more(b)) ber_skip(b);
}

// one Seq-id (already inside the SEQUENCE OF): returns its printed form, "" when suppressed
std::string seq_id(Ber &b, bool show_gis)
{
  static const char *const db[] = {"lcl", "bbs", "bbm", "gim", "gb", "emb", "pir", "sp", "pat", "ref",
                                   "gnl", "gi", "dbj", "prf", "pdb", "tpg", "tpe", "tpd", "gpp", "nat"};
  const int tag = b.peek();
  std::string out;
  if (tag < 0xa0 || tag > 0xb3) { ber_skip(b); return out; }
  const std::string dbn = db[tag - 0xa0];
  char num[64];
  ber_within(b, tag, [&](Ber &in) {
    switch (tag)
    {
      case 0xa0:                                         // local: Object-id
      {
        std::string s; unsigned long long n;
        object_id(in, s, n);
        if (!s.empty()) out = dbn + "|" + s;
        else { snprintf(num, sizeof num, "%llu", n); out = dbn + "|" + num; }
        break;
      }
      case 0xa1: case 0xa2:                              // gibbsq / gibbmt: INTEGER
        snprintf(num, sizeof num, "%llu", ber_uint(in));
        out = dbn + "|" + num;
        break;
      case 0xa3:                                         // giim: Giimport-id, first field = id
        ber_within(in, 0x30, [&](Ber &g) {
          if (g.peek() == 0xa0) ber_within(g, 0xa0, [&](Ber &v) { snprintf(num, sizeof num, "%llu", ber_uint(v)); });
          while (more(g)) ber_skip(g);
        });
        out = dbn + "|" + num;
        break;
      case 0xaa:                                         // general: Dbtag { db, tag Object-id }
        ber_within(in, 0x30, [&](Ber &g) {
          std::string dbname, s; unsigned long long n = 0;
          if (g.peek() == 0xa0) ber_within(g, 0xa0, [&](Ber &v) { dbname = ber_string(v); });
          if (g.peek() == 0xa1) ber_within(g, 0xa1, [&](Ber &v) { object_id(v, s, n); });
          while (more(g)) ber_skip(g);
          if (!s.empty()) out = dbn + "|" + dbname + "|" + s;
          else { snprintf(num, sizeof num, "%llu", n); out = dbn + "|" + dbname + "|" + num; }
        });
        break;
      case 0xab:                                         // gi: shown only on request (-I)
        snprintf(num, sizeof num, "%llu", ber_uint(in));
        if (show_gis) out = dbn + "|" + num;
        break;
      case 0xae:                                         // pdb: { mol, chain, rel }
        ber_within(in, 0x30, [&](Ber &g) {
          std::string mol; unsigned long long chain = 32;
          if (g.peek() == 0xa0) ber_within(g, 0xa0, [&](Ber &v) { mol = ber_string(v); });
          if (g.peek() == 0xa1) ber_within(g, 0xa1, [&](Ber &v) { chain = ber_uint(v); });
          while (more(g)) ber_skip(g);
          std::string ch;
          if (chain > 95) { ch.push_back((char)(chain - 32)); ch.push_back((char)(chain - 32)); }
          else ch.push_back((char)chain);
          out = dbn + "|" + mol + "|" + ch;
        });
        break;
      case 0xa8:                                         // patent: { seqid, cit { country, id } }
        ber_within(in, 0x30, [&](Ber &g) {
          unsigned long long seq = 0; std::string country, id; bool granted = true;
          if (g.peek() == 0xa0) ber_within(g, 0xa0, [&](Ber &v) { seq = ber_uint(v); });
          if (g.peek() == 0xa1)
            ber_within(g, 0xa1, [&](Ber &v) {
              ber_within(v, 0x30, [&](Ber &c) {
                if (c.peek() == 0xa0) ber_within(c, 0xa0, [&](Ber &w) { country = ber_string(w); });
                if (c.peek() == 0xa1)
                  ber_within(c, 0xa1, [&](Ber &w) {
                    const int t = w.peek();                 // number [0] (granted) or app-number [1]
                    granted = t == 0xa0;
                    ber_within(w, t, [&](Ber &x) { id = ber_string(x); });
                  });
                while (more(c)) ber_skip(c);
              });
            });
          while (more(g)) ber_skip(g);
          snprintf(num, sizeof num, "%llu", seq);
          out = std::string(granted ? "pat" : "pgp") + "|" + country + "|" + id + "|" + num;
        });
        break;
      default:                                           // the Textseq-id family
        ber_within(in, 0x30, [&](Ber &g) {
          TextId t;
          if (g.peek() == 0xa0) ber_within(g, 0xa0, [&](Ber &v) { t.name = ber_string(v); });
          if (g.peek() == 0xa1) ber_within(g, 0xa1, [&](Ber &v) { t.accession = ber_string(v); });
          if (g.peek() == 0xa2) ber_within(g, 0xa2, [&](Ber &v) { t.release = ber_string(v); });
          if (g.peek() == 0xa3) ber_within(g, 0xa3, [&](Ber &v) { t.version = ber_uint(v); });
          while (more(g)) ber_skip(g);
          std::string d = dbn;
          if (d == "sp" && t.release == "unreviewed") d = "tr";
          if (t.version) { snprintf(num, sizeof num, ".%llu", t.version); out = d + "|" + t.accession + num + "|" + t.name; }
          else out = d + "|" + t.accession + "|" + t.name;
        });
        break;
    }
    while (more(in)) ber_skip(in);
  });
  return out;
}

}  // namespace

extern "C" {

// Query text -> symbol codes (nucleotide: 0 = amino acids, 1 = nucleotides, 2 = the sound alphabet).  FASTA: an optional '>' description line, then sequence lines up to
// the next '>' or the end.  Characters outside the alphabet are dropped (query.cc:317-325).
// Returns the number of bytes of `text` consumed (the next record starts there), 0 at the end.
int64_t swb_query_parse(const char *text, int64_t text_len, int nucleotide, uint8_t *seq,
                        int64_t seq_cap, int64_t *seq_len, char *descr, int64_t descr_cap)
{
  if (!text || text_len < 0 || !seq_len || seq_cap < 0 || (seq_cap > 0 && !seq)) return SWB_ERR_ARG;
  *seq_len = 0;
  if (descr && descr_cap > 0) descr[0] = 0;
  if (text_len == 0) return 0;
  int64_t i = 0;
  if (text[0] == '>')
  {
    int64_t e = 1;
    while (e < text_len && text[e] != '\n') e++;
    if (descr && descr_cap > 0)
    {
      const int64_t n = std::min<int64_t>(e - 1, descr_cap - 1);
      memcpy(descr, text + 1, (size_t)n);
      descr[n] = 0;
    }
    i = e < text_len ? e + 1 : e;
  }
  int64_t n = 0;
  bool line_start = true;
  for (; i < text_len; i++)
  {
    const unsigned char c = (unsigned char)text[i];
    if (line_start && c == '>') break;
    line_start = c == '\n';
    const int m = nucleotide == 1 ? nt_of(c) : (nucleotide == 2 ? sound_of(c) : aa_of(c));
    if (m >= 0)
    {
      if (n < seq_cap) seq[n] = (uint8_t)m;
      n++;
    }
  }
  *seq_len = n;
  if (n > seq_cap) return SWB_ERR_RANGE;
  return i;
}

int swb_revcomp(const uint8_t *seq, int64_t len, uint8_t *out)
{
  if (len < 0 || (len > 0 && (!seq || !out))) return SWB_ERR_ARG;
  for (int64_t i = 0; i < len; i++) out[i] = (uint8_t)nt_complement(seq[len - 1 - i] & 15);
  return SWB_OK;
}

// table[256 a + 16 b + c] = amino-acid code of the codon of 4-bit nucleotide codes (a, b, c) under
// NCBI genetic code `gencode`: the common translation of every compatible plain codon, B / Z when
// they only differ as D/N or E/Q, X otherwise (query.cc:366-436).
int swb_translate_table(int gencode, uint8_t *table)
{
  if (gencode < 1 || gencode > 23 || !GENCODE[gencode - 1] || !table) return SWB_ERR_ARG;
  static const int tcag[4] = {2, 1, 3, 0};             // bit 0..3 = A C G T -> position in TCAG order
  const char *code = GENCODE[gencode - 1];
  for (int a = 0; a < 16; a++)
    for (int b = 0; b < 16; b++)
      for (int c = 0; c < 16; c++)
      {
        char aa = '-';
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++)
            {
              if (!((a >> i) & 1) || !((b >> j) & 1) || !((c >> k) & 1)) continue;
              const char x = code[tcag[i] * 16 + tcag[j] * 4 + tcag[k]];
              if (aa == '-' || aa == x) aa = x;
              else if (aa == 'B' && (x == 'D' || x == 'N')) {}
              else if ((aa == 'D' && (x == 'B' || x == 'N')) || (aa == 'N' && (x == 'B' || x == 'D'))) aa = 'B';
              else if (aa == 'Z' && (x == 'Q' || x == 'E')) {}
              else if ((aa == 'E' && (x == 'Z' || x == 'Q')) || (aa == 'Q' && (x == 'Z' || x == 'E'))) aa = 'Z';
              else aa = 'X';
            }
        if (aa == '-') aa = 'X';
        table[256 * a + 16 * b + c] = (uint8_t)aa_of(aa);
      }
  return SWB_OK;
}

const char *swb_gencode_name(int gencode)
{
  static const char *const names[23] = {
      "Standard Code", "Vertebrate Mitochondrial Code", "Yeast Mitochondrial Code",
      "Mold, Protozoan, and Coelenterate Mitochondrial Code and Mycoplasma/Spiroplasma Code",
      "Invertebrate Mitochondrial Code", "Ciliate, Dasycladacean and Hexamita Nuclear Code", nullptr, nullptr,
      "Echinoderm and Flatworm Mitochondrial Code", "Euplotid Nuclear Code",
      "Bacterial, Archaeal and Plant Plastid Code", "Alternative Yeast Nuclear Code",
      "Ascidian Mitochondrial Code", "Alternative Flatworm Mitochondrial Code", "Blepharisma Nuclear Code",
      "Chlorophycean Mitochondrial Code", nullptr, nullptr, nullptr, nullptr, "Trematode Mitochondrial Code",
      "Scenedesmus obliquus Mitochondrial Code", "Thraustochytrium Mitochondrial Code"};
  return gencode >= 1 && gencode <= 23 ? names[gencode - 1] : nullptr;
}

// One reading frame (query.cc:450-506): strand 0 reads forward from `frame`, strand 1 reads the
// reverse complement from the far end.  out must hold (len - frame) / 3 codes; returns that count.
int64_t swb_translate(const uint8_t *nt, int64_t len, int strand, int frame, const uint8_t *table,
                      uint8_t *out)
{
  if (len < 0 || frame < 0 || frame > 2 || !table || (len > 0 && !nt)) return SWB_ERR_ARG;
  const int64_t plen = len - frame >= 0 ? (len - frame) / 3 : 0;
  if (plen > 0 && !out) return SWB_ERR_ARG;
  if (!strand)
  {
    int64_t pos = frame;
    for (int64_t k = 0; k < plen; k++, pos += 3)
      out[k] = table[((nt[pos] & 15) << 8) | ((nt[pos + 1] & 15) << 4) | (nt[pos + 2] & 15)];
  }
  else
  {
    int64_t pos = len - 1 - frame;
    for (int64_t k = 0; k < plen; k++, pos -= 3)
      out[k] = table[(nt_complement(nt[pos] & 15) << 8) | (nt_complement(nt[pos - 1] & 15) << 4) |
                     nt_complement(nt[pos - 2] & 15)];
  }
  return plen;
}

// The deflines of one sequence header as text, one per line ('\n' separated), each
// "<seqids joined by |>[|taxid|N][|link|N][|memb|N] <title>" (asnparse.cc:753-887).  memb != 0
// keeps only deflines whose membership bits include it.  Returns the number of deflines kept.
// taxids != NULL: a bitmap (bit t & 7 of byte t / 8, db_check_taxid database.cc:718-733); deflines
// whose taxid is not in it are dropped, like deflines failing the membership test.
int64_t swb_defline_text(const uint8_t *data, int64_t len, int show_gis, int show_taxid, int64_t memb,
                         const uint8_t *taxids, int64_t taxid_bytes, char *buf, int64_t cap,
                         int64_t *needed)
{
  if (!data || len < 0 || cap < 0 || (cap > 0 && !buf)) return SWB_ERR_ARG;
  Ber top{data, data + len};
  std::string all;
  int64_t count = 0;
  const bool was = ber_within(top, 0x30, [&](Ber &set) {
    while (more(set))
    {
      if (set.peek() != 0x30) { set.ok = false; break; }
      ber_within(set, 0x30, [&](Ber &d) {
        std::string title = "unnamed protein product", ids;
        unsigned long long taxid = 0, membership = 0, links = 0;
        if (d.peek() == 0xa0) ber_within(d, 0xa0, [&](Ber &v) { title = ber_string(v); });
        if (d.peek() == 0xa1)
          ber_within(d, 0xa1, [&](Ber &v) {
            ber_within(v, 0x30, [&](Ber &s) {
              while (more(s))
              {
                const std::string id = seq_id(s, show_gis != 0);
                // the reference joins with '|' even when an id is suppressed only if ids is non-empty
                if (!ids.empty()) ids += "|";
                ids += id;
              }
            });
          });
        if (d.peek() == 0xa2) ber_within(d, 0xa2, [&](Ber &v) { taxid = ber_uint(v); });
        if (d.peek() == 0xa3)
          ber_within(d, 0xa3, [&](Ber &v) { ber_within(v, 0x30, [&](Ber &s) { while (more(s)) membership = ber_uint(s); }); });
        if (d.peek() == 0xa4)
          ber_within(d, 0xa4, [&](Ber &v) { ber_within(v, 0x30, [&](Ber &s) { while (more(s)) links = ber_uint(s); }); });
        while (more(d)) ber_skip(d);
        if (((long long)membership & memb) != memb) return;
        if (taxids && !((long long)(taxid >> 3) < taxid_bytes && ((taxids[taxid >> 3] >> (taxid & 7)) & 1))) return;
        std::string line = ids;
        char num[64];
        if (show_taxid)
        {
          if (taxid) { snprintf(num, sizeof num, "|taxid|%llu", taxid); line += num; }
          if (links) { snprintf(num, sizeof num, "|link|%llu", links); line += num; }
          if (membership) { snprintf(num, sizeof num, "|memb|%llu", membership); line += num; }
        }
        if (!line.empty() && !title.empty()) line += " ";
        line += title;
        if (count) all += "\n";
        all += line;
        count++;
      });
    }
  });
  if (!was || !top.ok) return SWB_ERR_IO;      // "Error parsing binary ASN.1 in database sequence definition."
  if (needed) *needed = (int64_t)all.size() + 1;
  if ((int64_t)all.size() + 1 > cap) return SWB_ERR_RANGE;
  memcpy(buf, all.c_str(), all.size() + 1);
  return count;
}

}  // extern "C"
