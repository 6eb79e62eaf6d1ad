// swb_blastdb.cu -- host-side reader of BLAST version-4 databases, the residue source of the scan.
//
// Takes over from the reference (torognes/swipe):
//   db_open / db_read_alias / db_open_xin ...... database.cc:406-608, :775-925
//   seqno_volume ............................... database.cc:637-660
//   db_check_msk ............................... database.cc:687-706
//   db_getsequence (host decode, for alignments) database.cc:1237-1401
// The scan itself never pulls sequences through this interface: swb_db_open_blast (swb_api.cu)
// uploads whole volumes' byte ranges and decodes nucleotide data on the device.
#include "../../include/swipe_b200.h"
#include "swb_blastdb.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace
{

thread_local std::string g_blast_error;

bool map_file(const std::string &name, const uint8_t **adr, size_t *len, std::string &err)
{
  const int fd = open(name.c_str(), O_RDONLY);
  if (fd < 0)
  {
    err = "Unable to open file " + name + ".";
    return false;
  }
  struct stat st;
  if (fstat(fd, &st) != 0)
  {
    close(fd);
    err = "Unable to stat file " + name + ".";
    return false;
  }
  *len = (size_t)st.st_size;
  *adr = nullptr;
  if (*len > 0)
  {
    void *p = mmap(nullptr, *len, PROT_READ, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED)
    {
      close(fd);
      err = "Unable to map file " + name + " in memory. It may be empty or too large.";
      return false;
    }
    *adr = (const uint8_t *)p;
  }
  close(fd);
  return true;
}

void unmap(const uint8_t *adr, size_t len)
{
  if (adr && len) munmap((void *)adr, len);
}

// one volume: index header then three (two for protein) BE u32 tables (database.cc:567-603)
bool open_volume(bool nt, const std::string &base, SwbVolume &v, std::string &err)
{
  v.base = base;
  const char *xi = nt ? ".nin" : ".pin", *xh = nt ? ".nhr" : ".phr", *xs = nt ? ".nsq" : ".psq";
  if (!map_file(base + xi, &v.idx, &v.len_idx, err)) return false;
  if (!map_file(base + xh, &v.hdr, &v.len_hdr, err)) return false;
  if (!map_file(base + xs, &v.seq, &v.len_seq, err)) return false;
  const uint8_t *p = v.idx, *e = v.idx + v.len_idx;
  auto need = [&](size_t n) { return (size_t)(e - p) >= n; };
  if (!need(12)) { err = "Truncated index file " + base + xi + "."; return false; }
  const uint32_t version = SwbVolume::be32(p); p += 4;
  if (version != 4) { err = "Illegal database version (must be 4)."; return false; }
  const uint32_t type = SwbVolume::be32(p); p += 4;
  if ((type == 1) == nt) { err = "Database " + base + " holds the other kind of sequences."; return false; }
  const uint32_t tlen = SwbVolume::be32(p); p += 4;
  if (!need((size_t)tlen + 4)) { err = "Truncated index file " + base + xi + "."; return false; }
  v.title.assign((const char *)p, tlen); p += tlen;
  const uint32_t dlen = SwbVolume::be32(p); p += 4;
  if (!need((size_t)dlen)) { err = "Truncated index file " + base + xi + "."; return false; }
  v.date.assign((const char *)p, dlen); p += dlen;
  while ((p - v.idx) & 3) p++;                       // database.cc:587-592
  if (!need(16)) { err = "Truncated index file " + base + xi + "."; return false; }
  v.nseq = SwbVolume::be32(p); p += 4;
  uint64_t sym = 0;
  memcpy(&sym, p, 8); p += 8;                        // little-endian, unlike the rest (database.cc:595)
  v.symcount = (long long)sym;
  v.longest = SwbVolume::be32(p); p += 4;
  const size_t tab = 4 * (size_t)(v.nseq + 1);
  if (!need(tab * (nt ? 3 : 2))) { err = "Truncated index file " + base + xi + "."; return false; }
  v.tab_hdr = p;
  v.tab_seq = p + tab;
  v.tab_amb = nt ? p + 2 * tab : nullptr;
  // the tables must stay inside the sequence file and be monotone; checked once here so that the
  // upload and the device decode can trust them
  long long prev = 0;
  for (long long s = 0; s <= v.nseq; s++)
  {
    const long long o = v.seq_off(s);
    if (o < prev || (size_t)o > v.len_seq) { err = "Corrupt sequence offsets in " + base + xi + "."; return false; }
    if (nt && s < v.nseq)
    {
      const long long a = v.amb_off(s);
      if (a < o || a > v.seq_off(s + 1)) { err = "Corrupt ambiguity offsets in " + base + xi + "."; return false; }
    }
    prev = o;
  }
  return true;
}

struct Alias
{
  bool present = false;
  std::string title;
  std::vector<std::string> dblist, oidlist;
  long long length = 0, nseq = 0, maxoid = 0, memb_bit = 0;
};

// names separated by blanks or double quotes, exactly as getnames splits them (database.cc:290-327)
std::vector<std::string> names_of(const char *line)
{
  static const char ws[] = " \t\r\n\"";
  std::vector<std::string> out;
  const char *p = line;
  for (;;)
  {
    p += strspn(p, ws);
    const size_t n = strcspn(p, ws);
    if (n == 0) break;
    out.emplace_back(p, n);
    p += n;
  }
  return out;
}

// .pal / .nal alias files (database.cc:406-490)
bool read_alias(bool nt, const std::string &base, Alias &a, std::string &err)
{
  FILE *f = fopen((base + (nt ? ".nal" : ".pal")).c_str(), "r");
  if (!f) return true;
  a.present = true;
  char line[10000];
  bool ok = true;
  while (fgets(line, sizeof line, f))
  {
    if (!strncmp(line, "TITLE ", 6))
    {
      const char *s = line + 6;
      s += strspn(s, " \t");
      a.title.assign(s, strcspn(s, "\r\n"));
    }
    else if (!strncmp(line, "DBLIST", 6)) a.dblist = names_of(line + 6);
    else if (!strncmp(line, "OIDLIST", 7)) a.oidlist = names_of(line + 7);
    else if (!strncmp(line, "GILIST", 6))
    {
      err = "GILIST in database alias files not implemented.";
      ok = false;
      break;
    }
    else if (!strncmp(line, "LENGTH ", 7)) a.length = atol(line + 7);
    else if (!strncmp(line, "NSEQ ", 5)) a.nseq = atol(line + 5);
    else if (!strncmp(line, "MAXOID ", 7)) a.maxoid = atol(line + 7);
    else if (!strncmp(line, "MEMB_BIT ", 9)) a.memb_bit = atol(line + 9);
  }
  fclose(f);
  if (a.title.empty()) a.title = base;
  return ok;
}

std::string dir_of(const std::string &base)
{
  const size_t k = base.rfind('/');
  return k == std::string::npos ? std::string() : base.substr(0, k + 1);
}

}  // namespace

extern "C" {

const char *swb_blastdb_error(void) { return g_blast_error.c_str(); }

int swb_blastdb_open(const char *basename, int nucleotide, swb_blastdb **out)
{
  if (!out) return SWB_ERR_ARG;
  *out = nullptr;
  if (!basename) return SWB_ERR_ARG;
  swb_blastdb *b = new (std::nothrow) swb_blastdb;
  if (!b) return SWB_ERR_NOMEM;
  b->nucleotide = nucleotide != 0;
  const bool nt = b->nucleotide;
  const std::string base = basename, path = dir_of(base);
  std::string err;
  auto fail = [&]() {
    g_blast_error = err;
    swb_blastdb_close(b);
    return SWB_ERR_IO;
  };
  auto add_volume = [&](const std::string &vbase, const Alias *mask_from, const std::string &mskfile) {
    b->vols.emplace_back();
    SwbVolume &v = b->vols.back();
    if (!open_volume(nt, vbase, v, err)) return false;
    if (mask_from)
    {
      v.masked_maxoid = mask_from->maxoid;
      v.masked_nseq = mask_from->nseq;
      v.masked_length = mask_from->length;
      if (!map_file(path + mskfile, &v.msk, &v.len_msk, err)) return false;
    }
    return true;
  };
  Alias top;
  if (!read_alias(nt, base, top, err)) return fail();
  if (top.present)
  {
    // one level of nesting, as the reference handles (database.cc:790-880)
    b->title = top.title;
    b->memb_bit = top.memb_bit;
    for (size_t i = 0; i < top.dblist.size(); i++)
    {
      const std::string base2 = path + top.dblist[i];
      Alias sub;
      if (!read_alias(nt, base2, sub, err)) return fail();
      if (sub.present)
      {
        if (b->memb_bit && (sub.oidlist.size() != 1 || sub.dblist.size() != 1))
        {
          err = "Illegal alias file (2).";
          return fail();
        }
        for (size_t j = 0; j < sub.dblist.size(); j++)
          if (!add_volume(path + sub.dblist[j], b->memb_bit ? &sub : nullptr,
                          b->memb_bit ? sub.oidlist[j] : std::string()))
            return fail();
      }
      else
      {
        if (top.oidlist.empty()) b->memb_bit = 0;
        if (b->memb_bit && (top.oidlist.size() != 1 || top.dblist.size() != 1))
        {
          err = "Illegal alias file (1).";
          return fail();
        }
        if (!add_volume(base2, b->memb_bit ? &top : nullptr, b->memb_bit ? top.oidlist[i] : std::string()))
          return fail();
      }
    }
    if (b->vols.empty())
    {
      err = "Alias file " + base + " lists no databases.";
      return fail();
    }
  }
  else
  {
    if (!add_volume(base, nullptr, std::string())) return fail();
    b->title = b->vols[0].title;
  }
  long long first = 0;
  for (SwbVolume &v : b->vols)
  {
    v.first = first;
    first += v.nseq;
    b->symcount += v.symcount;
    b->masked_nseq += v.masked_nseq;
    b->masked_symcount += v.masked_length;
    if (v.longest > b->longest) b->longest = v.longest;
  }
  b->nseq = first;
  b->date = b->vols[0].date;
  *out = b;
  return SWB_OK;
}

int swb_blastdb_close(swb_blastdb *b)
{
  if (!b) return SWB_OK;
  for (SwbVolume &v : b->vols)
  {
    unmap(v.idx, v.len_idx);
    unmap(v.hdr, v.len_hdr);
    unmap(v.seq, v.len_seq);
    unmap(v.msk, v.len_msk);
  }
  delete b;
  return SWB_OK;
}

int swb_blastdb_info(const swb_blastdb *b, int64_t *nseq, int64_t *symbols, int64_t *longest,
                     int *volumes)
{
  if (!b) return SWB_ERR_ARG;
  if (nseq) *nseq = b->nseq;
  if (symbols) *symbols = b->symcount;
  if (longest) *longest = b->longest;
  if (volumes) *volumes = (int)b->vols.size();
  return SWB_OK;
}

// totals hits_init uses for a masked database (db_getseqcount_masked / db_getsymcount_masked,
// database.cc:1046-1065): the alias file's NSEQ / LENGTH; equal to the plain totals when not masked
int swb_blastdb_masked_info(const swb_blastdb *b, int64_t *memb_bit, int64_t *nseq, int64_t *symbols)
{
  if (!b) return SWB_ERR_ARG;
  if (memb_bit) *memb_bit = b->memb_bit;
  if (nseq) *nseq = b->memb_bit ? b->masked_nseq : b->nseq;
  if (symbols) *symbols = b->memb_bit ? b->masked_symcount : b->symcount;
  return SWB_OK;
}

const char *swb_blastdb_title(const swb_blastdb *b) { return b ? b->title.c_str() : ""; }
const char *swb_blastdb_date(const swb_blastdb *b) { return b ? b->date.c_str() : ""; }

int64_t swb_blastdb_seqlen(const swb_blastdb *b, int64_t seqno)
{
  if (!b || seqno < 0 || seqno >= b->nseq) return SWB_ERR_ARG;
  long long s = 0;
  const SwbVolume *v = b->volume_of(seqno, &s);
  if (b->nucleotide) return swb_nt_length(*v, s);
  return v->seq_off(s + 1) - v->seq_off(s) - 1;      // NUL separated (database.cc:1246-1248)
}

// Host-side fetch of one sequence as symbol codes, what db_getsequence hands the reference's
// aligner: protein bytes as stored; nucleotides unpacked to 4-bit codes with the ambiguity runs
// patched in, reverse-complemented when strand != 0 (database.cc:1257-1353).
int swb_blastdb_sequence(const swb_blastdb *b, int64_t seqno, int strand, uint8_t *buf,
                         int64_t cap, int64_t *len)
{
  if (!b || seqno < 0 || seqno >= b->nseq || cap < 0 || (cap > 0 && !buf)) return SWB_ERR_ARG;
  long long s = 0;
  const SwbVolume *v = b->volume_of(seqno, &s);
  const long long o1 = v->seq_off(s), o2 = v->seq_off(s + 1);
  if (!b->nucleotide)
  {
    const long long n = o2 - o1 - 1;
    if (len) *len = n;
    if (n > cap) return SWB_ERR_RANGE;
    memcpy(buf, v->seq + o1, (size_t)std::max<long long>(n, 0));
    return SWB_OK;
  }
  const long long n = swb_nt_length(*v, s);
  if (len) *len = n;
  if (n > cap) return SWB_ERR_RANGE;
  const uint8_t *src = v->seq + o1;
  for (long long i = 0; i < n; i++) buf[i] = (uint8_t)(1u << ((src[i >> 2] >> ((3 - (i & 3)) << 1)) & 3));
  const long long a0 = v->amb_off(s), abytes = o2 - a0;
  if (abytes >= 4)
  {
    const uint8_t *p = v->seq + a0;
    const uint32_t head = SwbVolume::be32(p);
    p += 4;
    if (head >> 31)
      for (long long k = 0; k < (abytes - 4) / 8; k++, p += 8)
      {
        const uint64_t e = ((uint64_t)SwbVolume::be32(p) << 32) | SwbVolume::be32(p + 4);
        const uint64_t code = e >> 60, run = ((e >> 48) & 0xfff) + 1, off = e & 0x0000ffffffffffffULL;
        for (uint64_t r = 0; r < run && (long long)(off + r) < n; r++) buf[off + r] = (uint8_t)code;
      }
    else
      for (long long k = 0; k < (abytes - 4) / 4; k++, p += 4)
      {
        const uint32_t e = SwbVolume::be32(p);
        const uint32_t code = e >> 28, run = ((e >> 24) & 0xf) + 1, off = e & 0x00ffffff;
        for (uint32_t r = 0; r < run && (long long)(off + r) < n; r++) buf[off + r] = (uint8_t)code;
      }
  }
  if (strand)
  {
    // complement = bit-reversed 4-bit code (query.cc:112), order reversed
    for (long long i = 0, j = n - 1; i <= j; i++, j--)
    {
      const uint8_t x = buf[i], y = buf[j];
      buf[i] = (uint8_t)(((y & 1) << 3) | ((y & 2) << 1) | ((y & 4) >> 1) | ((y & 8) >> 3));
      buf[j] = (uint8_t)(((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3));
    }
  }
  return SWB_OK;
}

// The raw ASN.1 defline bytes of a sequence (.phr / .nhr; db_getheader, database.cc:1403-1413).
int swb_blastdb_header(const swb_blastdb *b, int64_t seqno, const uint8_t **data, int64_t *len)
{
  if (!b || seqno < 0 || seqno >= b->nseq || !data || !len) return SWB_ERR_ARG;
  long long s = 0;
  const SwbVolume *v = b->volume_of(seqno, &s);
  const long long h1 = v->hdr_off(s), h2 = v->hdr_off(s + 1);
  if (h2 < h1 || (size_t)h2 > v->len_hdr) return SWB_ERR_IO;
  *data = v->hdr + h1;
  *len = h2 - h1;
  return SWB_OK;
}

// membership bit of a masked database (.msk named by OIDLIST; db_check_msk, database.cc:687-706)
int swb_blastdb_included(const swb_blastdb *b, int64_t seqno)
{
  if (!b || seqno < 0 || seqno >= b->nseq) return 0;
  if (!b->memb_bit) return 1;
  long long s = 0;
  const SwbVolume *v = b->volume_of(seqno, &s);
  if (!v->msk || s > v->masked_maxoid) return 0;
  const long long byteno = s >> 3;
  if ((size_t)(4 + byteno) >= v->len_msk) return 0;
  return (v->msk[4 + byteno] >> (7 - (s & 7))) & 1;
}

}  // extern "C"
