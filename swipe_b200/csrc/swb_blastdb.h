// swb_blastdb.h -- internal: the parsed form of a BLAST version-4 database (one or more volumes)
// shared between the reader (swb_blastdb.cu) and the shard upload (swb_api.cu).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

struct SwbVolume
{
  std::string base;
  const uint8_t *idx = nullptr, *seq = nullptr, *hdr = nullptr, *msk = nullptr;   // mmaps
  size_t len_idx = 0, len_seq = 0, len_hdr = 0, len_msk = 0;
  long long nseq = 0, symcount = 0, longest = 0;
  std::string title, date;
  const uint8_t *tab_hdr = nullptr, *tab_seq = nullptr, *tab_amb = nullptr;        // BE u32 [nseq+1]
  long long masked_maxoid = 0, masked_nseq = 0, masked_length = 0;   // from the alias file (NSEQ / LENGTH)
  long long first = 0;                  // global number of the volume's first sequence

  static inline uint32_t be32(const uint8_t *p)
  {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
  long long seq_off(long long s) const { return be32(tab_seq + 4 * s); }
  long long amb_off(long long s) const { return be32(tab_amb + 4 * s); }
  long long hdr_off(long long s) const { return be32(tab_hdr + 4 * s); }
};

struct swb_blastdb
{
  bool nucleotide = false;
  long long memb_bit = 0;              // MEMB_BIT of the alias file; 0 = not masked
  long long masked_nseq = 0, masked_symcount = 0;
  std::vector<SwbVolume> vols;
  long long nseq = 0, symcount = 0, longest = 0;
  std::string title, date, error;

  // volume holding global sequence number s (database.cc:637-660)
  const SwbVolume *volume_of(long long s, long long *local) const
  {
    for (const SwbVolume &v : vols)
      if (s < v.first + v.nseq)
      {
        *local = s - v.first;
        return &v;
      }
    return nullptr;
  }
};

// nucleotide length of local sequence s of volume v (database.cc:1257-1261)
inline long long swb_nt_length(const SwbVolume &v, long long s)
{
  const long long o1 = v.seq_off(s), o3 = v.amb_off(s);
  const long long packed = o3 - o1;
  if (packed <= 0) return 0;
  return 4 * (packed - 1) + (v.seq[o3 - 1] & 3);
}
