// swb_api.cu -- the C ABI of include/swipe_b200.h: database shard upload and device-side
// re-layout, the scoring-table setup, the width cascade and the top-K sink.
//
// Host-side roles taken over from the reference (torognes/swipe):
//   search_chunk's cascade ............ swipe.cc:1416-1594   -> swb_search / swb_search_list
//   db_mapsequences / db_getsequence .. database.cc:1082-1131, :1237-1401 -> swb_db_open
//   score limits ...................... matrices.cc:560-577  -> Tables::prepare
//   hits_enter ........................ hits.cc:163-222      -> swb_topk_merge
#include "../../include/swipe_b200.h"
#include "sw_kernels.cuh"
#include "swb_blastdb.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace
{

thread_local std::string g_cuda_error;

#define SWB_CUDA(call)                                                                        \
  do                                                                                          \
  {                                                                                           \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
    {                                                                                         \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,       \
               __LINE__, #call);                                                              \
      g_cuda_error = buf_;                                                                    \
      (void)cudaGetLastError();                                                               \
      return e_ == cudaErrorMemoryAllocation ? SWB_ERR_NOMEM : SWB_ERR_CUDA;                  \
    }                                                                                         \
  } while (0)

#define SWB_TRY(expr)                 \
  do                                  \
  {                                   \
    int rc_ = (expr);                 \
    if (rc_ != SWB_OK) return rc_;    \
  } while (0)

// Device allocations are recycled through a small per-process cache: opening and closing a shard
// per search (the end-to-end path) would otherwise spend more time in cudaMalloc / cudaFree --
// which also synchronise the device -- than in the scan.  Sizes are rounded up to 1/8 octave.
struct DevCache
{
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void *> free_blocks;
  size_t cached[64] = {0};        // bytes held per device
  long long limit = -1;           // per-device budget; -1 = not set yet (SWB_CACHE_MB or 8 GiB)
  size_t budget()
  {
    if (limit < 0)
    {
      const char *env = getenv("SWB_CACHE_MB");
      limit = env ? std::max<long long>(0, atoll(env)) << 20 : (long long)8 << 30;
    }
    return (size_t)limit;
  }
  static size_t round_up(size_t bytes)
  {
    size_t b = bytes < 512 ? 512 : bytes;
    size_t p = 512;
    while (p < b) p <<= 1;
    const size_t step = p >> 4 ? p >> 4 : 1;        // p/2 < b <= p: granularity p/16
    return (b + step - 1) / step * step;
  }
  void *take(int dev, size_t rounded)
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = free_blocks.find(std::make_pair(dev, rounded));
    if (it == free_blocks.end()) return nullptr;
    void *p = it->second;
    free_blocks.erase(it);
    cached[dev & 63] -= rounded;
    return p;
  }
  void give(int dev, size_t rounded, void *p)
  {
    {
      std::lock_guard<std::mutex> g(mu);
      if (cached[dev & 63] + rounded <= budget())
      {
        free_blocks.emplace(std::make_pair(dev, rounded), p);
        cached[dev & 63] += rounded;
        return;
      }
    }
    cudaFree(p);
  }
  void trim()
  {
    std::lock_guard<std::mutex> g(mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : free_blocks)
    {
      cudaSetDevice(kv.first.first);
      cudaFree(kv.second);
    }
    cudaSetDevice(cur);
    free_blocks.clear();
    for (size_t &c : cached) c = 0;
  }
};
DevCache g_cache;

template <typename T> struct DevBuf
{
  T *p = nullptr;
  size_t cap = 0;        // elements
  size_t bytes = 0;      // rounded allocation size
  int dev = 0;
  int reserve(size_t n)
  {
    if (n <= cap && p) return SWB_OK;
    release();
    const size_t want = DevCache::round_up((n ? n : 1) * sizeof(T));
    int d = 0;
    SWB_CUDA(cudaGetDevice(&d));
    void *q = g_cache.take(d, want);
    if (!q)
    {
      cudaError_t e = cudaMalloc(&q, want);
      if (e == cudaErrorMemoryAllocation)
      {
        (void)cudaGetLastError();
        g_cache.trim();                         // give cached blocks back and retry once
        e = cudaMalloc(&q, want);
      }
      SWB_CUDA(e);
    }
    p = (T *)q;
    bytes = want;
    cap = want / sizeof(T);
    dev = d;
    return SWB_OK;
  }
  void release()
  {
    if (p) g_cache.give(dev, bytes, p);
    p = nullptr;
    cap = 0;
    bytes = 0;
  }
};

// The device-resident form of a set of subjects: length-sorted (descending), paired, cut
// into 4-column blocks.  Built for the whole shard at open time and for ad-hoc lists.
struct Layout
{
  long long first = 0;    // first subject of the chunk (0 for list layouts)
  long long n = 0;        // subjects
  long long npairs = 0;
  long long cap_blocks = 0;
  long long res_bytes = 0;                          // bound of the residues it covers
  cudaEvent_t ev_ready = nullptr;                   // layout complete (recorded on the layout stream)
  DevBuf<u32> keys_in, keys_out, idx_in, idx_out;   // idx_out[k] = list position of k-th longest
  DevBuf<long long> nblk, pairblk;                  // [npairs+1]
  DevBuf<uint2> blocks;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<u32> pair_scores;
  DevBuf<int> stream_pair;
  int stream_pair_n = -1;                           // streams the partition was computed for
  void release()
  {
    keys_in.release(); keys_out.release(); idx_in.release(); idx_out.release();
    nblk.release(); pairblk.release(); blocks.release(); cub_tmp.release();
    pair_scores.release(); stream_pair.release();
    if (ev_ready) cudaEventDestroy(ev_ready);
    ev_ready = nullptr;
  }
};

struct Shape { int G, R; };

// A batched scan (swb_search_batch) has already run the first tier for this query: the packed lane
// maxima of chunk k are at pair_scores[k] and search_impl starts from there.
struct Prescan
{
  std::vector<const u32 *> pair_scores;
};

// swb_search_hits: the sink's admission rule applied on the device (see swb_hist_kernel)
struct HitsReq
{
  long long seqno_base = 0, keep = 0, min_score = 0, upper = 0;
  long long *out_seqno = nullptr, *out_score = nullptr;
  long long nhits = 0, totalhits = 0, obvious = 0;
};

}  // namespace

struct swb_db
{
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long nseq = 0, total_res = 0, longest = 0;
  int trailing = 0;
  int mode = 0;
  DevBuf<unsigned char> residues;
  DevBuf<long long> offsets;
  DevBuf<unsigned char> packed;        // .nsq bytes as uploaded (nucleotide databases only)
  DevBuf<long long> pk_start;          // [nseq+1] start of every subject's record in packed
  DevBuf<u32> pk_len;                  // [nseq] bytes of packed bases in the record (rest: ambiguity table)
  DevBuf<unsigned char> nt_residues;   // decoded nucleotides of a translated shard (source of the 6 frames)
  DevBuf<long long> nt_offsets;        // [nsrc+1]
  DevBuf<unsigned char> ttable;        // [4096] codon table
  std::vector<Layout *> chunks;   // the whole shard, cut into upload/layout/scan pipeline chunks
  Layout tmp;      // ad-hoc list layouts
  Layout tmp2;     // the re-queue list of the 16-bit pass
  DevBuf<long long> requeue2, codes;
  cudaStream_t copy_stream = nullptr, layout_stream = nullptr;
  cudaStream_t stream2 = nullptr;      // second compute stream: overlapping launches while the shard arrives
  cudaEvent_t ev_setup = nullptr;
  cudaEvent_t ev_uploaded = nullptr;   // every byte of the shard is on the device
  cudaEvent_t ev_open[3] = {nullptr, nullptr, nullptr};
  bool opened_sync = false;
  // per-search device scratch
  DevBuf<short> m16;
  DevBuf<unsigned short> qrow_off;
  DevBuf<long long> matrix;
  DevBuf<unsigned char> query;
  DevBuf<long long> scores, bestpos, bestq, requeue, list;
  DevBuf<unsigned long long> counters;       // [0] requeue count, [1..3] width counts
  DevBuf<unsigned char> he;
  DevBuf<uint4> bndH, bndF;
  DevBuf<int> slot_flags;         // multi-pass scratch regions in use (ScanParams::slot_flags)
  DevBuf<unsigned char> filter;   // swb_db_set_filter: one bit per subject, honoured by the device sink
  std::vector<unsigned char> h_filter;
  bool has_filter = false;
  DevBuf<ScanSeg> segs;           // chunk table of a merged (whole-shard) scan launch
  DevBuf<unsigned> hist;          // [SWB_HIST_BINS] admissible scores, [SWB_HIST_BINS] = the cut bin
  DevBuf<unsigned long long> cand, cand_sorted;   // score << 32 | subject of the candidates
  DevBuf<unsigned char> sort_tmp;
  std::vector<unsigned long long> h_cand;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_group[2] = {nullptr, nullptr};   // completion of the last two scan launches
  cudaEvent_t ev_batch[2] = {nullptr, nullptr};   // around a batched scan launch
  double upload_ms = 0, layout_ms = 0;
  int force_G = 0, force_R = 0, force_mode = -1;   // test hooks (swb_set_shape)
  int force_geom = 0;                              // 0 = automatic, 1 / 2 = scan kernel geometry (swb_set_geometry)
};

namespace
{

int build_layout(swb_db *db, Layout &L, const long long *d_list, long long first, long long n,
                 long long res_bound, cudaStream_t st)
{
  L.first = first;
  L.n = n;
  L.npairs = (n + 1) / 2;
  L.stream_pair_n = -1;
  if (n == 0) return SWB_OK;
  if (n > 0x7fffffffLL) return SWB_ERR_ARG;
  SWB_TRY(L.keys_in.reserve(n)); SWB_TRY(L.keys_out.reserve(n + 1));
  SWB_TRY(L.idx_in.reserve(n)); SWB_TRY(L.idx_out.reserve(n + 1));
  SWB_TRY(L.nblk.reserve(L.npairs + 1)); SWB_TRY(L.pairblk.reserve(L.npairs + 1));
  SWB_TRY(L.pair_scores.reserve(L.npairs));
  const int T = 256;
  const long long *offs = db->offsets.p + first;       // subject k of the chunk = first + k
  swb_len_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(offs, db->trailing, d_list, n,
                                                            L.keys_in.p, L.idx_in.p);
  SWB_CUDA(cudaGetLastError());
  size_t tmp1 = 0, tmp2 = 0;
  int bits = 1;
  while (bits < 32 && (db->longest >> bits) != 0) bits++;
  SWB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp1, L.keys_in.p, L.keys_out.p,
                                                     L.idx_in.p, L.idx_out.p, (int)n, 0, bits, st));
  SWB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, L.nblk.p, L.pairblk.p,
                                         (int)(L.npairs + 1), st));
  SWB_TRY(L.cub_tmp.reserve(std::max(tmp1, tmp2)));
  size_t tb = L.cub_tmp.cap;
  SWB_CUDA(cub::DeviceRadixSort::SortPairsDescending(L.cub_tmp.p, tb, L.keys_in.p, L.keys_out.p,
                                                     L.idx_in.p, L.idx_out.p, (int)n, 0, bits, st));
  // an odd subject count leaves the last pair's second lane empty
  SWB_CUDA(cudaMemsetAsync(L.keys_out.p + n, 0, sizeof(u32), st));
  SWB_CUDA(cudaMemsetAsync(L.nblk.p + L.npairs, 0, sizeof(long long), st));
  swb_pairblk_kernel<<<(unsigned)((L.npairs + T - 1) / T), T, 0, st>>>(L.keys_out.p, n, L.npairs,
                                                                      L.nblk.p);
  SWB_CUDA(cudaGetLastError());
  tb = L.cub_tmp.cap;
  SWB_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp.p, tb, L.nblk.p, L.pairblk.p,
                                         (int)(L.npairs + 1), st));
  if (d_list)
  {
    // a caller's list may name a subject any number of times, so the residues of the shard do not
    // bound it: read the exact block total back (the list paths synchronise anyway)
    long long total_blocks = 0;
    SWB_CUDA(cudaMemcpyAsync(&total_blocks, L.pairblk.p + L.npairs, sizeof total_blocks,
                             cudaMemcpyDeviceToHost, st));
    SWB_CUDA(cudaStreamSynchronize(st));
    L.res_bytes = 4 * total_blocks;
    L.cap_blocks = total_blocks + 1;
  }
  else
  {
    // upper bound of the block total without a round trip: sum(max len) <= total residues
    const long long bound_res = std::min<long long>(res_bound, n * db->longest);
    L.res_bytes = bound_res;
    L.cap_blocks = (bound_res + 3 * L.npairs) / 4 + 1;
  }
  SWB_TRY(L.blocks.reserve((size_t)L.cap_blocks));
  const long long threads = L.npairs * 32;
  swb_fill_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, st>>>(
      db->residues.p, offs, db->trailing, d_list, L.keys_out.p, L.idx_out.p, n, L.npairs,
      L.pairblk.p, L.blocks.p);
  SWB_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---- scoring tables ------------------------------------------------------------------------
struct Tables
{
  int nq = 0;                  // distinct query symbols = table rows
  int rowof[32];               // query symbol -> table row
  long long hi = 0, lo = 0;    // over the whole 32x32 table (matrices.cc:560-571)
  long long limit7 = 0, limit16 = 0;
  bool narrow_ok = false;      // the packed kernels can represent this scoring system
  bool hybrid_ok = false;
  std::vector<short> m16;      // [33][34] + padding = SWB_M16_BYTES, the image staged in shared memory
  std::vector<unsigned short> qrow;
};

inline short enc16(long long v, int mode)
{
  if (mode != SWB_MODE_INT16 && v < 0) return (short)(0x8000u | (unsigned)(-v));
  return (short)v;
}

// pad_after: table rows between the last symbol row and the padding row (geometry 1 keeps a row of
// START / END flags there, geometry 2 does not)
int prepare_tables(Tables &t, const unsigned char *query, long long qlen, const swb_scoring *sc,
                   int mode, int rows_padded, int pad_after = 1)
{
  for (int i = 0; i < 32; i++) t.rowof[i] = -1;
  t.nq = 0;
  for (long long i = 0; i < qlen; i++)
  {
    if (query[i] > 31) return SWB_ERR_ARG;
    if (t.rowof[query[i]] < 0) t.rowof[query[i]] = t.nq++;
  }
  t.hi = -100; t.lo = 100;
  for (int i = 0; i < 1024; i++)
  {
    t.hi = std::max(t.hi, (long long)sc->matrix[i]);
    t.lo = std::min(t.lo, (long long)sc->matrix[i]);
  }
  t.limit7 = 128 - t.hi;
  t.limit16 = 65536 - t.hi;
  const long long q = sc->gap_open_extend, r = sc->gap_extend;
  t.narrow_ok = t.nq <= 32 && t.hi <= 1024 && t.lo >= -1024 && q >= 0 && q <= 8192 && r >= 0 &&
                r <= 8192;
  t.hybrid_ok = t.narrow_ok && t.hi <= 256 && t.lo >= -1023 && q <= 1023 && r <= 1023;
  t.m16.assign(SWB_M16_BYTES / sizeof(short), 0);
  for (int d = 0; d < SWB_MROWS; d++)
    for (int s = 0; s < 32; s++)
    {
      long long v = SWB_PAD_SCORE;
      if (d < 32)
        for (int qs = 0; qs < 32; qs++)
          if (t.rowof[qs] == s) v = sc->matrix[(d << 5) + qs];
      t.m16[d * SWB_MS_STRIDE + s] = t.narrow_ok ? enc16(v, mode) : (short)0;
    }
  t.qrow.assign((size_t)rows_padded, (unsigned short)((t.nq + pad_after) * 16));
  for (long long i = 0; i < qlen && i < rows_padded; i++)
    t.qrow[(size_t)i] = (unsigned short)(t.rowof[query[i]] * 16);
  return SWB_OK;
}

// ---- kernel shapes -----------------------------------------------------------------------------
typedef void (*scan_fn)(const ScanParams);
struct ShapeEntry { int G, R, mode; scan_fn fn, fn_mp; u32 kq, kr; int geom; };   // single-pass / multi-pass builds

// hybrid-mode encodings of the two default scoring systems' penalties, compiled in as immediates:
// BLOSUM62 11+1k (open+extend 12, extend 1) and nucleotide 5+2k (7, 2)
#define SWB_KQ(q) (0x8000u | (q)) | ((0x8000u | (q)) << 16)
#define SWB_KR(r) ((0x10000u - (r)) | ((0x10000u - (r)) << 16))

#define SWB_SHAPE_OF(KERNEL, GEOM, G, R)                                                         \
  {G, R, SWB_MODE_INT16, KERNEL<G, R, SWB_MODE_INT16, false>,                                    \
   KERNEL<G, R, SWB_MODE_INT16, true>, 0, 0, GEOM},                                              \
  {G, R, SWB_MODE_HYBRID, KERNEL<G, R, SWB_MODE_HYBRID, false>,                                  \
   KERNEL<G, R, SWB_MODE_HYBRID, true>, 0, 0, GEOM},                                             \
  {G, R, SWB_MODE_HYBRID, KERNEL<G, R, SWB_MODE_HYBRID, false, SWB_KQ(12), SWB_KR(1)>,           \
   KERNEL<G, R, SWB_MODE_HYBRID, true, SWB_KQ(12), SWB_KR(1)>, SWB_KQ(12), SWB_KR(1), GEOM},     \
  {G, R, SWB_MODE_HYBRID, KERNEL<G, R, SWB_MODE_HYBRID, false, SWB_KQ(7), SWB_KR(2)>,            \
   KERNEL<G, R, SWB_MODE_HYBRID, true, SWB_KQ(7), SWB_KR(2)>, SWB_KQ(7), SWB_KR(2), GEOM}
#define SWB_SHAPE(G, R) SWB_SHAPE_OF(swb_scan_kernel, 1, G, R)
#define SWB_SHAPE2(G, R) SWB_SHAPE_OF(swb_scan2_kernel, 2, G, R)

const ShapeEntry g_shapes[] = {
    SWB_SHAPE(4, 25),  // one warp per CTA, no inter-warp hand-off at all: measured SLOWER (4.3 TCUPS at 100 aa: the lone warp builds every table itself); kept for the comparison
    SWB_SHAPE(8, 8),   SWB_SHAPE(8, 13),  SWB_SHAPE(8, 16),  SWB_SHAPE(16, 12), SWB_SHAPE(16, 16),
    SWB_SHAPE(16, 20), SWB_SHAPE(16, 24), SWB_SHAPE(32, 12), SWB_SHAPE(32, 16), SWB_SHAPE(32, 20),
    SWB_SHAPE(32, 24), SWB_SHAPE(32, 28), SWB_SHAPE(32, 32),
    // geometry 2 (one warp = one stage of 32 streams)
    SWB_SHAPE2(16, 20), SWB_SHAPE2(16, 21), SWB_SHAPE2(16, 24), SWB_SHAPE2(16, 25), SWB_SHAPE2(4, 25),
};
const int g_nshapes = (int)(sizeof(g_shapes) / sizeof(g_shapes[0]));

// rows covered per pass = G*R; cost ~ padded rows * (1 + per-step overhead / R)
// Measured throughput of every compiled shape on padded rows (TCUPS of the hybrid build with the
// penalties compiled in, 2 M-subject shard, tools/tune_shapes.py -> profiles/r1_tune_shapes.txt):
// single pass / multi-pass.  Shapes without a multi-pass measurement use 0.93 x single pass.
struct ShapeEff { int geom, G, R; double sp, mp; };
const ShapeEff g_eff[] = {
    // geometry 1: r1 table scaled by what the round-2 row loop / boundary prefetch measured (x 1.025 single
    // pass, x 1.037 multi-pass), with the shapes re-measured in round 2 entered as measured
    // (profiles/r2_tune_shapes.txt)
    {1, 4, 25, 4.31, 3.5},  {1, 8, 8, 5.12, 0},     {1, 8, 13, 5.70, 0},    {1, 8, 16, 6.15, 0},    {1, 16, 12, 6.58, 0},   {1, 16, 16, 7.22, 0},
    {1, 16, 20, 7.27, 0},  {1, 16, 24, 7.61, 6.41}, {1, 32, 12, 6.52, 5.92}, {1, 32, 16, 6.81, 6.68}, {1, 32, 20, 6.97, 6.99},
    {1, 32, 24, 7.69, 6.69}, {1, 32, 28, 7.57, 6.65}, {1, 32, 32, 6.36, 6.62},
    // geometry 2 (measured): only ahead for short queries, where its four-stage CTA amortises the table build
    {2, 16, 20, 6.5, 5.9}, {2, 16, 21, 6.6, 6.08}, {2, 16, 24, 6.92, 5.3}, {2, 16, 25, 6.9, 5.3}, {2, 4, 25, 5.82, 5.0},
};

// (kq, kr): the penalties in the mode's packed encoding; a build with exactly these compiled in is
// preferred over the generic one of the same shape (except where it measured slower).  The shape
// that scores the query's rows fastest wins: efficiency x qlen / (passes x G x R).
inline int shape_streams(const ShapeEntry &s) { return s.geom == 2 ? SWB2_STREAMS : SWB_STREAMS; }
inline int shape_threads(const ShapeEntry &s) { return s.geom == 2 ? swb_scan2_threads(s.G) : swb_scan_threads(s.G); }
inline size_t shape_smem(const ShapeEntry &s, int nq)
{
  return s.geom == 2 ? swb_scan2_smem(s.G, nq) : swb_scan_smem(s.G, nq);
}
const size_t SWB_SMEM_LIMIT = 227 * 1024;      // dynamic shared memory one CTA may ask for on sm_100

const ShapeEntry *choose_shape(const swb_db *db, long long qlen, int mode, u32 kq, u32 kr, int *npass, int nq)
{
  const ShapeEntry *best = nullptr;
  double best_rate = 0;
  const bool allow_spec = getenv("SWB_NO_SPEC") == nullptr;
  int geom = db->force_geom;                   // 0: both geometries compete on measured efficiency
  if (geom == 0)
    if (const char *env = getenv("SWB_GEOM")) geom = atoi(env);
  for (int i = 0; i < g_nshapes; i++)
  {
    const ShapeEntry &s = g_shapes[i];
    if (s.mode != mode) continue;
    const bool spec = (s.kq | s.kr) != 0;
    if (spec && !(allow_spec && s.kq == kq && s.kr == kr)) continue;
    if (db->force_G && (s.G != db->force_G || s.R != db->force_R)) continue;
    if (geom && s.geom != geom) continue;
    if (shape_smem(s, nq) + 1024 > SWB_SMEM_LIMIT) continue;      // (+ the kernel's static shared memory)
    const long long rows = (long long)s.G * s.R;
    const long long np = std::max<long long>(1, (qlen + rows - 1) / rows);
    double eff = 5.0;
    for (const ShapeEff &e : g_eff)
      if (e.G == s.G && e.R == s.R && e.geom == s.geom) eff = np > 1 ? (e.mp > 0 ? e.mp : 0.93 * e.sp) : e.sp;
    if (!spec) eff *= 0.955;                                   // generic builds read the penalties from registers
    if (mode == SWB_MODE_INT16) eff *= 0.84;
    const double rate = eff * (double)std::max<long long>(qlen, 1) / (double)(np * rows);
    if (!best || rate > best_rate) { best = &s; best_rate = rate; *npass = (int)np; }
  }
  return best;
}

int run_wide(swb_db *db, const long long *d_list, const long long *d_sel, long long nsel,
             const unsigned char *d_query, long long qlen, const swb_scoring *sc, bool end,
             bool use64, int *launches)
{
  if (nsel == 0) return SWB_OK;
  cudaStream_t st = db->stream;
  // (a long query over very many subjects would want more pass-boundary scratch than is reasonable: that
  // unusual combination stays on the one-thread-per-subject kernel below, which works in batches)
  const bool end_scratch_ok = qlen <= 1024 ||
                              (double)nsel * 2.0 * (double)std::max<long long>(db->longest, 1) * sizeof(int) <= 4e9;
  if (end && !use64 && end_scratch_ok && getenv("SWB_NO_END_KERNEL") == nullptr)
  {
    // the alignment phase's few subjects: one warp each (swb_end_kernel), state in registers; queries
    // beyond 1024 rows go in passes with the pass boundary (2 ints per subject column) in scratch
    WideParams W;
    memset(&W, 0, sizeof W);
    if (qlen > 1024)
    {
      SWB_TRY(db->he.reserve((size_t)nsel * 2 * (size_t)std::max<long long>(db->longest, 1) * sizeof(int)));
      W.he = db->he.p;
      W.stride = std::max<long long>(db->longest, 1);
    }
    W.residues = db->residues.p; W.offsets = db->offsets.p; W.trailing = db->trailing;
    W.list = d_list; W.sel = d_sel; W.nsel = nsel;
    W.query = d_query; W.qlen = (int)qlen; W.matrix = db->matrix.p;
    W.q = sc->gap_open_extend; W.r = sc->gap_extend;
    W.scores = db->scores.p; W.bestpos = db->bestpos.p; W.bestq = db->bestq.p;
    const unsigned grid = (unsigned)((nsel * 32 + 127) / 128);
    if (qlen <= 256) swb_end_kernel<8><<<grid, 128, 0, st>>>(W);
    else if (qlen <= 512) swb_end_kernel<16><<<grid, 128, 0, st>>>(W);
    else swb_end_kernel<32><<<grid, 128, 0, st>>>(W);
    SWB_CUDA(cudaGetLastError());
    (*launches)++;
    return SWB_OK;
  }
  const long long batch = 1 << 16;
  const size_t cell = use64 ? 8 : 4;
  SWB_TRY(db->he.reserve((size_t)(2 * std::max<long long>(qlen, 1)) *
                         (size_t)std::min(batch, nsel) * cell));
  for (long long first = 0; first < nsel; first += batch)
  {
    const long long m = std::min(batch, nsel - first);
    WideParams W;
    W.residues = db->residues.p; W.offsets = db->offsets.p; W.trailing = db->trailing;
    W.list = d_list; W.sel = d_sel ? d_sel + first : nullptr; W.nsel = m;
    W.query = d_query; W.qlen = (int)qlen; W.matrix = db->matrix.p;
    W.q = sc->gap_open_extend; W.r = sc->gap_extend;
    W.he = db->he.p; W.stride = m;
    W.scores = db->scores.p; W.bestpos = db->bestpos.p; W.bestq = db->bestq.p;
    if (!d_sel)
    {
      // dense positions first .. first+m-1: shift the output views instead of building a list
      W.scores += first; W.bestpos = W.bestpos ? W.bestpos + first : nullptr;
      W.bestq = W.bestq ? W.bestq + first : nullptr;
      W.list = d_list ? d_list + first : nullptr;
      if (!d_list) { W.offsets += first; }
    }
    const unsigned grid = (unsigned)((m + 127) / 128);
    if (use64)
    {
      if (end) swb_wide_kernel<long long, true><<<grid, 128, 0, st>>>(W);
      else swb_wide_kernel<long long, false><<<grid, 128, 0, st>>>(W);
    }
    else
    {
      if (end) swb_wide_kernel<int, true><<<grid, 128, 0, st>>>(W);
      else swb_wide_kernel<int, false><<<grid, 128, 0, st>>>(W);
    }
    SWB_CUDA(cudaGetLastError());
    (*launches)++;
  }
  return SWB_OK;
}

int check_args(const swb_db *db, const unsigned char *query, long long qlen, const swb_scoring *sc)
{
  if (!db || !sc || !sc->matrix || qlen < 0 || (qlen > 0 && !query)) return SWB_ERR_ARG;
  if (qlen > 0x3fffffff) return SWB_ERR_ARG;
  if (sc->gap_open_extend < 0 || sc->gap_extend < 0) return SWB_ERR_ARG;
  return SWB_OK;
}

// The cascade over n subjects (the whole shard when h_list == NULL).
int search_impl(swb_db *db, const unsigned char *query, long long qlen, const swb_scoring *sc,
                const long long *h_list, long long n, long long *scores, long long *bestpos,
                long long *bestq, swb_counters *ctr, HitsReq *hits = nullptr, const Prescan *pre = nullptr)
{
  SWB_TRY(check_args(db, query, qlen, sc));
  if (n < 0 || (n > 0 && !scores && !hits)) return SWB_ERR_ARG;
  SWB_CUDA(cudaSetDevice(db->device));
  cudaStream_t st = db->stream;
  const bool want_end = bestpos != nullptr;
  swb_counters c;
  memset(&c, 0, sizeof c);
  c.subjects = n;
  int launches = 0;

  if (h_list)
    for (long long k = 0; k < n; k++)
    {
      // frames never reach the kernels (translated subjects are separate entries of the shard);
      // the strand bit is honoured by the end-cell search only
      if ((h_list[k] & 3) != 0 || ((h_list[k] & 4) != 0 && !want_end)) return SWB_ERR_ARG;
      const long long s = h_list[k] >> 3;
      if (s < 0 || s >= db->nseq) return SWB_ERR_ARG;
    }

  // scoring tables; the score range decides whether the wide kernel needs 64-bit cells
  int mode = SWB_MODE_HYBRID;
  if (db->force_mode >= 0) mode = db->force_mode;
  Tables tb;
  int npass = 1;
  {
    Tables probe;
    SWB_TRY(prepare_tables(probe, query, qlen, sc, mode, 0));
    if (mode != SWB_MODE_INT16 && !probe.hybrid_ok) mode = SWB_MODE_INT16;
  }
  u32 kq = 0, kr = 0;
  if (mode == SWB_MODE_HYBRID && sc->gap_open_extend <= 1023 && sc->gap_extend >= 1 && sc->gap_extend <= 1023)
  {
    const u32 a = (u32)(unsigned short)enc16(-sc->gap_open_extend, mode), b = (u32)(unsigned short)(short)(-sc->gap_extend);
    kq = a | (a << 16);
    kr = b | (b << 16);
  }
  int nq_probe = 0;
  {
    Tables probe;
    SWB_TRY(prepare_tables(probe, query, qlen, sc, mode, 0));
    nq_probe = probe.nq;
  }
  const ShapeEntry *shape = choose_shape(db, qlen, mode, kq, kr, &npass, nq_probe);
  if (!shape) return SWB_ERR_INTERNAL;
  const long long rows_padded = (long long)npass * shape->G * shape->R;
  SWB_TRY(prepare_tables(tb, query, qlen, sc, mode, (int)rows_padded, shape->geom == 2 ? 0 : 1));
  const long long maxcell = std::max<long long>(tb.hi, 0) * std::min<long long>(qlen, db->longest);
  const bool use64 = maxcell + sc->gap_open_extend + 65536 > 0x7fffffffLL ||
                     sc->gap_open_extend > 0x3fffffff || sc->gap_extend > 0x3fffffff ||
                     tb.lo < -0x3fffffff;

  SWB_TRY(db->scores.reserve((size_t)std::max<long long>(n, 1)));
  SWB_TRY(db->counters.reserve(8));
  SWB_TRY(db->matrix.reserve(1024));
  SWB_TRY(db->query.reserve((size_t)std::max<long long>(qlen, 1)));
  SWB_CUDA(cudaMemcpyAsync(db->matrix.p, sc->matrix, 1024 * sizeof(long long),
                           cudaMemcpyHostToDevice, st));
  if (qlen > 0)
    SWB_CUDA(cudaMemcpyAsync(db->query.p, query, (size_t)qlen, cudaMemcpyHostToDevice, st));
  SWB_CUDA(cudaMemsetAsync(db->counters.p, 0, 8 * sizeof(unsigned long long), st));
  if (want_end)
  {
    SWB_TRY(db->bestpos.reserve((size_t)std::max<long long>(n, 1)));
    SWB_TRY(db->bestq.reserve((size_t)std::max<long long>(n, 1)));
  }
  const long long *d_list = nullptr;
  if (h_list && n > 0)
  {
    SWB_TRY(db->list.reserve((size_t)n));
    SWB_CUDA(cudaMemcpyAsync(db->list.p, h_list, (size_t)n * sizeof(long long),
                             cudaMemcpyHostToDevice, st));
    d_list = db->list.p;
  }

  long long nrequeue = 0;
  const bool narrow = !want_end && db->mode == 0 && tb.narrow_ok && qlen > 0 && n > 0;
  if (n == 0 || qlen == 0)
  {
    // nothing to scan, but the open's copies may still be reading the caller's buffers: the
    // contract lets them go once a search has returned
    SWB_CUDA(cudaStreamWaitEvent(st, db->ev_uploaded, 0));
    if (n > 0) SWB_CUDA(cudaMemsetAsync(db->scores.p, 0, (size_t)n * sizeof(long long), st));
    else SWB_CUDA(cudaStreamSynchronize(st));
  }
  else if (narrow)
  {
    std::vector<Layout *> layouts;
    if (d_list)
    {
      SWB_CUDA(cudaStreamWaitEvent(st, db->ev_uploaded, 0));
      SWB_TRY(build_layout(db, db->tmp, d_list, 0, n, db->total_res, st));
      launches += 5;
      layouts.push_back(&db->tmp);
    }
    else
      layouts = db->chunks;
    const int limit = (mode != SWB_MODE_INT16 ? 2047 : 32767) - (int)std::max<long long>(tb.hi, 0);
    int stagger = 1;
    if (const char *env = getenv("SWB_STAGGER")) stagger = atoi(env) != 0;
    SWB_TRY(db->requeue.reserve((size_t)n));
    if (pre)
    {
      // first tier done by a batched launch: unpack its lane maxima, queue what left the exact range
      SWB_CUDA(cudaEventRecord(db->ev[0], st));
      size_t k = 0;
      for (Layout *L : layouts)
      {
        if (L->n == 0) continue;
        if (k >= pre->pair_scores.size()) return SWB_ERR_INTERNAL;
        swb_finish_kernel<<<(unsigned)((L->n + 255) / 256), 256, 0, st>>>(
            pre->pair_scores[k++], L->idx_out.p, L->n, L->first, limit, nullptr, db->scores.p, db->requeue.p,
            db->counters.p);
        SWB_CUDA(cudaGetLastError());
        launches++;
      }
    }
    else
    {
    // launch geometry: one CTA = 8 streams x G stages; as many CTAs per SM as shared memory
    // and registers allow
    const int threads = shape_threads(*shape);
    const size_t smem = shape_smem(*shape, tb.nq);
    const int cta_streams = shape_streams(*shape);
    const scan_fn fn = npass > 1 ? shape->fn_mp : shape->fn;
    SWB_CUDA(cudaFuncSetAttribute((const void *)fn,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, threads,
                                                           smem));
    if (occ < 1) return SWB_ERR_INTERNAL;
    if (const char *env = getenv("SWB_CTAS_PER_SM"))     // tuning hook: run below full occupancy
      occ = std::max(1, std::min(occ, atoi(env)));
    // More CTAs than resident slots: the hardware hands a fresh CTA to whichever SM finishes one,
    // which evens out the warp scheduler's unfairness between co-resident CTAs (one persistent
    // wave left SMs running 1-2 warps per scheduler towards the end of every launch).
    int oversub = 2;
    if (const char *env = getenv("SWB_OVERSUB")) oversub = std::max(1, atoi(env));
    long long min_blocks = 0;
    for (Layout *L : layouts) min_blocks = std::max(min_blocks, L->cap_blocks);
    while (oversub > 1 && min_blocks / ((long long)db->sm_count * occ * oversub * cta_streams) < 40 * shape->G)
      oversub--;                                   // keep streams much longer than the pipeline fill
    SWB_TRY(db->m16.reserve(SWB_M16_BYTES / sizeof(short)));
    SWB_TRY(db->qrow_off.reserve((size_t)rows_padded));
    SWB_CUDA(cudaMemcpyAsync(db->m16.p, tb.m16.data(), tb.m16.size() * sizeof(short),
                             cudaMemcpyHostToDevice, st));
    SWB_CUDA(cudaMemcpyAsync(db->qrow_off.p, tb.qrow.data(), tb.qrow.size() * sizeof(unsigned short),
                             cudaMemcpyHostToDevice, st));
    // Chunks are scanned in GROUPS of consecutive chunks, one launch per group (blockIdx.y = chunk
    // of the group), so that the SMs run from one chunk into the next and the ragged end of a launch
    // is paid once per group.  Resident shard: one group.  Shard still arriving (asynchronous open):
    // a group is whatever has been laid out by the time the GPU is about to need more work -- at most
    // two launches are kept in flight -- so upload, re-layout and scan overlap and the launches grow
    // as the upload gets ahead of the scan.
    std::vector<Layout *> work;
    for (Layout *L : layouts)
      if (L->n > 0) work.push_back(L);
    bool merge = work.size() <= 65535;                               // gridDim.y limit
    if (const char *env = getenv("SWB_MERGE")) merge = merge && atoi(env) != 0;
    bool resident = true;
    for (Layout *L : work)
      if (L->ev_ready && cudaEventQuery(L->ev_ready) != cudaSuccess) resident = false;
    (void)cudaGetLastError();
    if (!resident && getenv("SWB_WAIT_RESIDENT") != nullptr)      // tuning hook: no overlap of upload and scan
    {
      SWB_TRY(swb_db_wait(db));
      resident = true;
    }
    if (merge && resident && !getenv("SWB_OVERSUB")) oversub = work.size() >= 4 ? 1 : oversub;
    // a shard that is still arriving is scanned chunk group by chunk group anyway: one CTA per resident slot
    // and chunk keeps the streams of its small first chunks as long as possible (measured: 15.2 -> 14.5 ms end
    // to end for an eighth of the 5 M database, 97.3 -> 96.4 ms for all of it)
    if (!resident && !getenv("SWB_OVERSUB")) oversub = 1;
    const int grid = db->sm_count * occ * oversub;
    const int nstreams = grid * cta_streams;
    long long sum_blocks = 0, all_blocks = 0;
    std::vector<ScanSeg> segs(work.size());
    for (size_t k = 0; k < work.size(); k++)
    {
      Layout *L = work[k];
      if (L->stream_pair_n != nstreams)
      {
        SWB_TRY(L->stream_pair.reserve((size_t)nstreams + 1));
        L->stream_pair_n = -nstreams - 1;                            // reserved, partition still to run
      }
      ScanSeg &S = segs[k];
      S.blocks = L->blocks.p; S.pairblk = L->pairblk.p; S.stream_pair = L->stream_pair.p;
      S.pair_scores = L->pair_scores.p;
      S.bnd_base = all_blocks;
      all_blocks += L->cap_blocks;
      sum_blocks = std::max(sum_blocks, L->cap_blocks);
    }
    // Multi-pass scratch (ScanParams::bndH): one entry per block of the shard while that fits the budget
    // (SWB_BND_BUDGET_MB, default 64 GiB: the faster layout, 4 x the shard's own size), else one region per
    // resident CTA, each holding the longest stream any chunk can give it: an equal share of the chunk's
    // blocks plus one pair's worth of slack.
    long long bnd_budget = 64ll << 30;
    if (const char *env = getenv("SWB_BND_BUDGET_MB")) bnd_budget = std::max<long long>(0, atoll(env)) << 20;
    int nslots = 0;
    long long bnd_stream = 0, bnd_cta = 0;
    if (npass > 1)
    {
      if (2 * all_blocks * (long long)sizeof(uint4) > bnd_budget)
      {
        nslots = db->sm_count * occ;
        bnd_stream = sum_blocks / nstreams + (db->longest + 3) / 4 + 2;
        bnd_cta = bnd_stream * cta_streams;
        SWB_TRY(db->slot_flags.reserve((size_t)nslots));
        SWB_CUDA(cudaMemsetAsync(db->slot_flags.p, 0, (size_t)nslots * sizeof(int), st));
      }
      const size_t entries = nslots ? (size_t)nslots * (size_t)bnd_cta : (size_t)all_blocks;
      SWB_TRY(db->bndH.reserve(entries));
      SWB_TRY(db->bndF.reserve(entries));
    }
    SWB_TRY(db->segs.reserve(std::max<size_t>(segs.size(), 1)));
    if (!segs.empty())
      SWB_CUDA(cudaMemcpyAsync(db->segs.p, segs.data(), segs.size() * sizeof(ScanSeg),
                               cudaMemcpyHostToDevice, st));
    const long long q = sc->gap_open_extend, r = sc->gap_extend;
    const unsigned nq16 = (unsigned)(unsigned short)enc16(-q, mode);
    const unsigned nr16 = (unsigned)(unsigned short)(short)(-r);
    const unsigned pad16 = (unsigned)(unsigned short)enc16(SWB_PAD_SCORE, mode);
    ScanParams P;
    memset(&P, 0, sizeof P);
    P.m16 = db->m16.p; P.qrow_off = db->qrow_off.p;
    P.bndH = db->bndH.p; P.bndF = db->bndF.p;
    P.slot_flags = db->slot_flags.p; P.nslots = nslots; P.bnd_cta = bnd_cta; P.bnd_stream = bnd_stream;
    P.nq = tb.nq; P.npass = npass;
    P.negq = nq16 | (nq16 << 16); P.negr = nr16 | (nr16 << 16); P.padword = pad16 | (pad16 << 16);
    P.stagger = stagger;
    P.start_mask = 1u; P.end_mask = 1u << (shape->G - 1);
    SWB_CUDA(cudaEventRecord(db->ev[0], st));
    // While the shard is arriving, consecutive groups alternate between two streams: the CTAs of the
    // next launch fill the SMs as those of the previous one drain, so the launch boundary costs nothing.
    const bool two_streams = !resident && db->stream2 && getenv("SWB_ONE_STREAM") == nullptr;
    if (two_streams)
    {
      SWB_CUDA(cudaEventRecord(db->ev_setup, st));               // tables, chunk table, counters are set up on st
      SWB_CUDA(cudaStreamWaitEvent(db->stream2, db->ev_setup, 0));
    }
    size_t next = 0;
    int group_no = 0;
    while (next < work.size())
    {
      size_t end = next + 1;
      cudaStream_t gs = (two_streams && (group_no & 1)) ? db->stream2 : st;
      if (!resident)
      {
        if (group_no >= 2) SWB_CUDA(cudaEventSynchronize(db->ev_group[group_no & 1]));   // group_no - 2 is done
        if (work[next]->ev_ready) SWB_CUDA(cudaEventSynchronize(work[next]->ev_ready));
      }
      if (merge)
        while (end < work.size() &&
               (resident || !work[end]->ev_ready || cudaEventQuery(work[end]->ev_ready) == cudaSuccess))
          end++;
      (void)cudaGetLastError();
      for (size_t k = next; k < end; k++)
      {
        Layout *L = work[k];
        if (L->ev_ready) SWB_CUDA(cudaStreamWaitEvent(gs, L->ev_ready, 0));
        if (L->stream_pair_n != nstreams)
        {
          swb_partition_kernel<<<(nstreams + 1 + 255) / 256, 256, 0, gs>>>(L->pairblk.p, L->npairs,
                                                                           nstreams, L->stream_pair.p);
          SWB_CUDA(cudaGetLastError());
          launches++;
          L->stream_pair_n = nstreams;
        }
        SWB_CUDA(cudaMemsetAsync(L->pair_scores.p, 0, (size_t)L->npairs * sizeof(u32), gs));
      }
      P.seg = segs[next];
      P.segs = db->segs.p + next;
      fn<<<dim3((unsigned)grid, (unsigned)(end - next)), threads, smem, gs>>>(P);
      SWB_CUDA(cudaGetLastError());
      launches++;
      for (size_t k = next; k < end; k++)
      {
        Layout *L = work[k];
        swb_finish_kernel<<<(unsigned)((L->n + 255) / 256), 256, 0, gs>>>(
            L->pair_scores.p, L->idx_out.p, L->n, L->first, limit, nullptr, db->scores.p, db->requeue.p,
            db->counters.p);
        SWB_CUDA(cudaGetLastError());
        launches++;
      }
      if (!resident) SWB_CUDA(cudaEventRecord(db->ev_group[group_no & 1], gs));
      group_no++;
      next = end;
    }
    if (two_streams && group_no >= 2)
      SWB_CUDA(cudaStreamWaitEvent(st, db->ev_group[1], 0));     // the last launch on stream2 (odd groups)
    }
    SWB_CUDA(cudaEventRecord(db->ev[1], st));
    unsigned long long h_nreq = 0;
    SWB_CUDA(cudaMemcpyAsync(&h_nreq, db->counters.p, sizeof h_nreq, cudaMemcpyDeviceToHost, st));
    SWB_CUDA(cudaStreamSynchronize(st));
    nrequeue = (long long)h_nreq;
    SWB_CUDA(cudaEventRecord(db->ev[2], st));
    const long long *wide_sel = db->requeue.p;
    long long nwide = nrequeue;
    if (nrequeue > 0 && mode == SWB_MODE_HYBRID && getenv("SWB_NO_MIDDLE") == nullptr)
    {
      // Middle tier of the cascade (the reference's 16-bit pass, swipe.cc:1483-1540): lanes that left
      // the 11-bit range of the fp16-pattern build are scanned again by the pure int16 build (limit
      // 32767 - hi) over a layout of just those subjects; only what overflows that goes to the wide
      // kernel.
      Tables t16;
      int np16 = 1;
      const ShapeEntry *sh16 = choose_shape(db, qlen, SWB_MODE_INT16, 0, 0, &np16, nq_probe);
      if (!sh16) return SWB_ERR_INTERNAL;
      const long long rows16 = (long long)np16 * sh16->G * sh16->R;
      SWB_TRY(prepare_tables(t16, query, qlen, sc, SWB_MODE_INT16, (int)rows16, sh16->geom == 2 ? 0 : 1));
      SWB_TRY(db->codes.reserve((size_t)nrequeue));
      SWB_TRY(db->requeue2.reserve((size_t)nrequeue));
      SWB_TRY(db->qrow_off.reserve((size_t)rows16));
      swb_requeue_codes_kernel<<<(unsigned)((nrequeue + 255) / 256), 256, 0, st>>>(db->requeue.p, nrequeue,
                                                                                  d_list, db->codes.p);
      SWB_CUDA(cudaGetLastError());
      SWB_CUDA(cudaStreamWaitEvent(st, db->ev_uploaded, 0));
      SWB_TRY(build_layout(db, db->tmp2, db->codes.p, 0, nrequeue, db->total_res, st));
      launches += 6;
      SWB_CUDA(cudaMemcpyAsync(db->m16.p, t16.m16.data(), t16.m16.size() * sizeof(short),
                               cudaMemcpyHostToDevice, st));
      SWB_CUDA(cudaMemcpyAsync(db->qrow_off.p, t16.qrow.data(), t16.qrow.size() * sizeof(unsigned short),
                               cudaMemcpyHostToDevice, st));
      SWB_CUDA(cudaMemsetAsync(db->counters.p, 0, sizeof(unsigned long long), st));
      const int threads16 = shape_threads(*sh16);
      const size_t smem16 = shape_smem(*sh16, t16.nq);
      const int streams16 = shape_streams(*sh16);
      const scan_fn fn16 = np16 > 1 ? sh16->fn_mp : sh16->fn;
      SWB_CUDA(cudaFuncSetAttribute((const void *)fn16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
      int occ16 = 0;
      SWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ16, (const void *)fn16, threads16, smem16));
      if (occ16 < 1) return SWB_ERR_INTERNAL;
      // few subjects: no more streams than pairs (a stream walks its pairs one after the other)
      Layout &L2 = db->tmp2;
      int grid16 = db->sm_count * occ16;
      grid16 = (int)std::max<long long>(1, std::min<long long>(grid16, (L2.npairs + streams16 - 1) / streams16));
      const int nstreams16 = grid16 * streams16;
      SWB_TRY(L2.stream_pair.reserve((size_t)nstreams16 + 1));
      swb_partition_kernel<<<(nstreams16 + 1 + 255) / 256, 256, 0, st>>>(L2.pairblk.p, L2.npairs, nstreams16,
                                                                         L2.stream_pair.p);
      SWB_CUDA(cudaGetLastError());
      L2.stream_pair_n = nstreams16;
      SWB_CUDA(cudaMemsetAsync(L2.pair_scores.p, 0, (size_t)L2.npairs * sizeof(u32), st));
      const int nslots16 = db->sm_count * occ16;
      const long long bnd_stream16 = L2.cap_blocks / nstreams16 + (db->longest + 3) / 4 + 2;
      if (np16 > 1)
      {
        SWB_TRY(db->bndH.reserve((size_t)nslots16 * (size_t)bnd_stream16 * (size_t)streams16));
        SWB_TRY(db->bndF.reserve((size_t)nslots16 * (size_t)bnd_stream16 * (size_t)streams16));
        SWB_TRY(db->slot_flags.reserve((size_t)nslots16));
        SWB_CUDA(cudaMemsetAsync(db->slot_flags.p, 0, (size_t)nslots16 * sizeof(int), st));
      }
      ScanParams P2;
      memset(&P2, 0, sizeof P2);
      P2.seg.blocks = L2.blocks.p; P2.seg.pairblk = L2.pairblk.p; P2.seg.stream_pair = L2.stream_pair.p;
      P2.seg.pair_scores = L2.pair_scores.p;
      P2.m16 = db->m16.p; P2.qrow_off = db->qrow_off.p; P2.bndH = db->bndH.p; P2.bndF = db->bndF.p;
      P2.slot_flags = db->slot_flags.p; P2.nslots = nslots16; P2.bnd_stream = bnd_stream16;
      P2.bnd_cta = bnd_stream16 * streams16;
      P2.nq = t16.nq; P2.npass = np16;
      const unsigned q16 = (unsigned)(unsigned short)enc16(-sc->gap_open_extend, SWB_MODE_INT16);
      const unsigned r16 = (unsigned)(unsigned short)(short)(-sc->gap_extend);
      const unsigned p16 = (unsigned)(unsigned short)enc16(SWB_PAD_SCORE, SWB_MODE_INT16);
      P2.negq = q16 | (q16 << 16); P2.negr = r16 | (r16 << 16); P2.padword = p16 | (p16 << 16);
      P2.stagger = stagger;
      P2.start_mask = 1u; P2.end_mask = 1u << (sh16->G - 1);
      fn16<<<grid16, threads16, smem16, st>>>(P2);
      SWB_CUDA(cudaGetLastError());
      const int limit16 = 32767 - (int)std::max<long long>(tb.hi, 0);
      swb_finish_kernel<<<(unsigned)((nrequeue + 255) / 256), 256, 0, st>>>(
          L2.pair_scores.p, L2.idx_out.p, nrequeue, 0, limit16, db->requeue.p, db->scores.p, db->requeue2.p,
          db->counters.p);
      SWB_CUDA(cudaGetLastError());
      launches += 4;
      unsigned long long h_n2 = 0;
      SWB_CUDA(cudaMemcpyAsync(&h_n2, db->counters.p, sizeof h_n2, cudaMemcpyDeviceToHost, st));
      SWB_CUDA(cudaStreamSynchronize(st));
      nwide = (long long)h_n2;
      wide_sel = db->requeue2.p;
      c.gpu_middle = nrequeue - nwide;
    }
    SWB_TRY(run_wide(db, d_list, wide_sel, nwide, db->query.p, qlen, sc, false, use64, &launches));
    SWB_CUDA(cudaEventRecord(db->ev[3], st));
  }
  else
  {
    SWB_CUDA(cudaStreamWaitEvent(st, db->ev_uploaded, 0));
    SWB_CUDA(cudaEventRecord(db->ev[0], st));
    SWB_CUDA(cudaEventRecord(db->ev[1], st));
    SWB_CUDA(cudaEventRecord(db->ev[2], st));
    SWB_TRY(run_wide(db, d_list, nullptr, n, db->query.p, qlen, sc, want_end, use64, &launches));
    SWB_CUDA(cudaEventRecord(db->ev[3], st));
    nrequeue = n;
  }

  if (hits && n > 0 && !use64)
  {
    // The sink on the device: only the best `keep` (subject, score) pairs cross the bus.
    SWB_TRY(db->hist.reserve(SWB_HIST_BINS + 1));
    SWB_TRY(db->cand.reserve((size_t)n));
    SWB_CUDA(cudaMemsetAsync(db->hist.p, 0, (SWB_HIST_BINS + 1) * sizeof(unsigned), st));
    const unsigned hgrid = (unsigned)std::min<long long>((n + 256 * 16 - 1) / (256 * 16), (long long)db->sm_count * 8);
    const unsigned char *filter = db->has_filter ? db->filter.p : nullptr;
    swb_hist_kernel<<<std::max(hgrid, 1u), 256, 0, st>>>(db->scores.p, n, tb.limit7, tb.limit16, hits->min_score,
                                                         hits->upper, filter, db->counters.p + 1, db->hist.p);
    SWB_CUDA(cudaGetLastError());
    swb_cut_kernel<<<1, 1024, 0, st>>>(db->hist.p, hits->keep, db->hist.p + SWB_HIST_BINS);
    SWB_CUDA(cudaGetLastError());
    launches += 2;
    unsigned long long h_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (hits->keep > 0)
    {
      swb_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
          db->scores.p, n, hits->min_score, hits->upper, filter, db->hist.p + SWB_HIST_BINS, db->cand.p,
          db->counters.p + 6);
      SWB_CUDA(cudaGetLastError());
      launches++;
    }
    SWB_CUDA(cudaMemcpyAsync(h_cnt, db->counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
    SWB_CUDA(cudaStreamSynchronize(st));
    c.ref_width7 = (long long)h_cnt[1];
    c.ref_width16 = (long long)h_cnt[2];
    c.ref_width63 = (long long)h_cnt[3];
    hits->totalhits = (long long)h_cnt[4];
    hits->obvious = (long long)h_cnt[5];
    const long long ncand = (long long)h_cnt[6];
    const long long take = std::min(ncand, hits->keep);
    if (take > 0)
    {
      const unsigned long long *sorted = db->cand.p;
      if (ncand > 1)
      {
        SWB_TRY(db->cand_sorted.reserve((size_t)ncand));
        size_t tmp = 0;
        SWB_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp, db->cand.p, db->cand_sorted.p, (int)ncand, 0, 64, st));
        SWB_TRY(db->sort_tmp.reserve(tmp));
        tmp = db->sort_tmp.cap;
        SWB_CUDA(cub::DeviceRadixSort::SortKeysDescending(db->sort_tmp.p, tmp, db->cand.p, db->cand_sorted.p, (int)ncand, 0, 64, st));
        launches += 4;
        sorted = db->cand_sorted.p;
      }
      db->h_cand.resize((size_t)take);
      SWB_CUDA(cudaMemcpyAsync(db->h_cand.data(), sorted, (size_t)take * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, st));
      SWB_CUDA(cudaStreamSynchronize(st));
      for (long long k = 0; k < take; k++)
      {
        hits->out_seqno[k] = hits->seqno_base + (long long)(db->h_cand[(size_t)k] & 0xffffffffull);
        hits->out_score[k] = (long long)(db->h_cand[(size_t)k] >> 32);
      }
    }
    hits->nhits = take;
    if (qlen > 0)
    {
      float ms = 0;
      SWB_CUDA(cudaEventElapsedTime(&ms, db->ev[0], db->ev[1]));
      c.scan_ms = ms;
      SWB_CUDA(cudaEventElapsedTime(&ms, db->ev[2], db->ev[3]));
      c.requeue_ms = ms;
    }
  }
  else if (n > 0)
  {
    std::vector<long long> dense;
    if (hits && !scores)
    {
      // scores that may not fit the 32-bit half of a candidate key: dense read-back, host sink
      dense.resize((size_t)n);
      scores = dense.data();
    }
    swb_widthcount_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        db->scores.p, n, tb.limit7, tb.limit16, db->counters.p + 1);
    SWB_CUDA(cudaGetLastError());
    launches++;
    unsigned long long h_cnt[4] = {0, 0, 0, 0};
    SWB_CUDA(cudaMemcpyAsync(scores, db->scores.p, (size_t)n * sizeof(long long),
                             cudaMemcpyDeviceToHost, st));
    if (want_end)
    {
      SWB_CUDA(cudaMemcpyAsync(bestpos, db->bestpos.p, (size_t)n * sizeof(long long),
                               cudaMemcpyDeviceToHost, st));
      SWB_CUDA(cudaMemcpyAsync(bestq, db->bestq.p, (size_t)n * sizeof(long long),
                               cudaMemcpyDeviceToHost, st));
    }
    SWB_CUDA(cudaMemcpyAsync(h_cnt, db->counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
    SWB_CUDA(cudaStreamSynchronize(st));
    c.ref_width7 = (long long)h_cnt[1];
    c.ref_width16 = (long long)h_cnt[2];
    c.ref_width63 = (long long)h_cnt[3];
    if (qlen > 0)
    {
      float ms = 0;
      SWB_CUDA(cudaEventElapsedTime(&ms, db->ev[0], db->ev[1]));
      c.scan_ms = ms;
      SWB_CUDA(cudaEventElapsedTime(&ms, db->ev[2], db->ev[3]));
      c.requeue_ms = ms;
    }
    if (hits)
    {
      std::vector<long long> kept;             // the host sink of the wide-cell path under a subject filter
      if (db->has_filter)
      {
        kept.assign(scores, scores + n);
        for (long long k = 0; k < n; k++)
          if (!((db->h_filter[(size_t)(k >> 3)] >> (k & 7)) & 1)) kept[(size_t)k] = LLONG_MIN;
      }
      const int64_t *arr[1] = {(const int64_t *)(db->has_filter ? kept.data() : scores)};
      const int64_t cnt[1] = {n}, base[1] = {hits->seqno_base};
      int64_t tot = 0, obv = 0;
      const int64_t k = swb_topk_merge(1, arr, cnt, base, hits->keep, hits->min_score, hits->upper,
                                       (int64_t *)hits->out_seqno, (int64_t *)hits->out_score, &tot, &obv);
      if (k < 0) return (int)k;
      hits->nhits = k; hits->totalhits = tot; hits->obvious = obv;
    }
  }
  if (narrow && !pre)
  {
    c.scan_geometry = shape->geom; c.scan_G = shape->G; c.scan_R = shape->R; c.scan_passes = npass;
  }
  c.gpu_requeued = nrequeue;
  c.gpu_narrow = n - nrequeue;
  c.kernel_launches = launches;
  // cells: the reference's GCUPS numerator, symbols * qlen (swipe.cc:1744-1775)
  if (!h_list) c.cells = db->total_res * qlen;
  if (ctr) *ctr = c;
  return SWB_OK;
}

// ---- several queries per scan (the reference loops over its queries one search at a
// time, swipe.cc:2561) --------------------------------------------------------------------------------
// Queries are laid one after the other along the pipeline of the geometry-2 kernel, each on whole stages
// (ScanParams::start_mask / end_mask): one pass over the residue stream and ONE table build per block
// serve all of them, which is what a short query cannot amortise on its own.  The first tier runs
// batched; unpacking, the re-queue tiers and the sink then run per query through search_impl (Prescan).
const ShapeEntry *find_shape(int geom, int G, int R, int mode, u32 kq, u32 kr)
{
  const ShapeEntry *generic = nullptr;
  const bool allow_spec = getenv("SWB_NO_SPEC") == nullptr;
  for (int i = 0; i < g_nshapes; i++)
  {
    const ShapeEntry &s = g_shapes[i];
    if (s.geom != geom || s.G != G || s.R != R || s.mode != mode) continue;
    if ((s.kq | s.kr) == 0) generic = &s;
    else if (allow_spec && s.kq == kq && s.kr == kr) return &s;
  }
  return generic;
}

int batch_impl(swb_db *db, int nqueries, const uint8_t *const *queries, const int64_t *qlens,
               const swb_scoring *sc, int64_t *const *scores, HitsReq *hits, swb_counters *counters)
{
  if (!db || nqueries < 0 || (nqueries > 0 && (!queries || !qlens))) return SWB_ERR_ARG;
  if (!hits && nqueries > 0 && !scores) return SWB_ERR_ARG;
  for (int k = 0; k < nqueries; k++)
    SWB_TRY(check_args(db, queries[k], qlens[k], sc));
  SWB_CUDA(cudaSetDevice(db->device));
  SWB_TRY(swb_db_wait(db));
  cudaStream_t st = db->stream;
  auto single = [&](int k, const Prescan *pre) {
    return search_impl(db, queries[k], qlens[k], sc, nullptr, db->nseq, scores ? (long long *)scores[k] : nullptr,
                       nullptr, nullptr, counters ? counters + k : nullptr, hits ? hits + k : nullptr, pre);
  };
  // can the packed kernels hold this scoring system at all, and in which lane arithmetic?
  int mode = SWB_MODE_HYBRID;
  if (db->force_mode >= 0) mode = db->force_mode;
  bool packed = db->mode == 0 && db->nseq > 0 && db->force_geom != 1 && getenv("SWB_NO_BATCH") == nullptr;
  u32 kq = 0, kr = 0;
  if (packed)
  {
    Tables probe;
    const unsigned char none = 0;
    SWB_TRY(prepare_tables(probe, &none, 0, sc, mode, 0));
    if (mode != SWB_MODE_INT16 && !probe.hybrid_ok) mode = SWB_MODE_INT16;
    packed = probe.narrow_ok;
    if (mode == SWB_MODE_HYBRID && sc->gap_open_extend <= 1023 && sc->gap_extend >= 1 && sc->gap_extend <= 1023)
    {
      const u32 a = (u32)(unsigned short)enc16(-sc->gap_open_extend, mode), b = (u32)(unsigned short)(short)(-sc->gap_extend);
      kq = a | (a << 16);
      kr = b | (b << 16);
    }
  }
  const int G = 16;
  const int cand_R[] = {25, 24, 21, 20};
  int k0 = 0;
  while (k0 < nqueries)
  {
    // the next group: as many of the following queries as fit the 16 stages, with the rows-per-stage
    // that wastes the fewest pipeline rows
    int best_m = 1, best_R = 0;
    double best_fill = 0;
    for (int R : cand_R)
    {
      if (!packed || !find_shape(2, G, R, mode, kq, kr)) continue;
      int stages = 0, m = 0;
      long long rows = 0;
      for (int k = k0; k < nqueries; k++)
      {
        const long long need = (qlens[k] + R - 1) / R;
        if (qlens[k] <= 0 || stages + need > G) break;
        stages += (int)need;
        rows += qlens[k];
        m++;
      }
      const double fill = (double)rows / (double)(G * R);
      if (m >= 2 && fill > best_fill) { best_fill = fill; best_m = m; best_R = R; }
    }
    if (best_m < 2)
    {
      SWB_TRY(single(k0, nullptr));
      k0++;
      continue;
    }
    const int m = best_m, R = best_R;
    const ShapeEntry *shape = find_shape(2, G, R, mode, kq, kr);
    // tables over the union of the group's symbols; query rows in pipeline order, padded per query
    std::vector<unsigned char> cat;
    for (int k = k0; k < k0 + m; k++) cat.insert(cat.end(), queries[k], queries[k] + qlens[k]);
    Tables tb;
    SWB_TRY(prepare_tables(tb, cat.data(), (long long)cat.size(), sc, mode, 0, 0));
    const size_t smem = shape_smem(*shape, tb.nq);
    if (smem + 1024 > SWB_SMEM_LIMIT)
    {
      SWB_TRY(single(k0, nullptr));              // too many distinct symbols for the one-CTA ring
      k0++;
      continue;
    }
    ScanParams P;
    memset(&P, 0, sizeof P);
    std::vector<unsigned short> qrow((size_t)G * R, (unsigned short)(tb.nq * 16));
    int stage = 0;
    for (int k = k0; k < k0 + m; k++)
    {
      const int need = (int)((qlens[k] + R - 1) / R);
      P.start_mask |= 1u << stage;
      P.end_mask |= 1u << (stage + need - 1);
      for (int g = stage; g < stage + need; g++) P.stage_query[g] = (unsigned char)(k - k0);
      for (long long i = 0; i < qlens[k]; i++)
        qrow[(size_t)stage * R + (size_t)i] = (unsigned short)(tb.rowof[queries[k][i]] * 16);
      stage += need;
    }
    if (stage < G) P.start_mask |= 1u << stage;    // idle tail stages: cut off from the last query
    const int threads = shape_threads(*shape);
    const scan_fn fn = shape->fn;
    SWB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, threads, smem));
    if (occ < 1) return SWB_ERR_INTERNAL;
    const int grid = db->sm_count * occ;
    const int nstreams = grid * shape_streams(*shape);
    std::vector<Layout *> work;
    for (Layout *L : db->chunks)
      if (L->n > 0) work.push_back(L);
    if (work.size() > 65535) return SWB_ERR_RANGE;
    std::vector<ScanSeg> segs(work.size());
    for (size_t c = 0; c < work.size(); c++)
    {
      Layout *L = work[c];
      if (L->stream_pair_n != nstreams)
      {
        SWB_TRY(L->stream_pair.reserve((size_t)nstreams + 1));
        swb_partition_kernel<<<(nstreams + 1 + 255) / 256, 256, 0, st>>>(L->pairblk.p, L->npairs, nstreams,
                                                                         L->stream_pair.p);
        SWB_CUDA(cudaGetLastError());
        L->stream_pair_n = nstreams;
      }
      SWB_TRY(L->pair_scores.reserve((size_t)L->npairs * (size_t)m));
      SWB_CUDA(cudaMemsetAsync(L->pair_scores.p, 0, (size_t)L->npairs * (size_t)m * sizeof(u32), st));
      ScanSeg &S = segs[c];
      S.blocks = L->blocks.p; S.pairblk = L->pairblk.p; S.stream_pair = L->stream_pair.p;
      S.pair_scores = L->pair_scores.p; S.score_stride = L->npairs;
    }
    SWB_TRY(db->m16.reserve(SWB_M16_BYTES / sizeof(short)));
    SWB_TRY(db->qrow_off.reserve(qrow.size()));
    SWB_TRY(db->segs.reserve(std::max<size_t>(segs.size(), 1)));
    SWB_CUDA(cudaMemcpyAsync(db->m16.p, tb.m16.data(), tb.m16.size() * sizeof(short), cudaMemcpyHostToDevice, st));
    SWB_CUDA(cudaMemcpyAsync(db->qrow_off.p, qrow.data(), qrow.size() * sizeof(unsigned short),
                             cudaMemcpyHostToDevice, st));
    SWB_CUDA(cudaMemcpyAsync(db->segs.p, segs.data(), segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    const unsigned nq16 = (unsigned)(unsigned short)enc16(-sc->gap_open_extend, mode);
    const unsigned nr16 = (unsigned)(unsigned short)(short)(-sc->gap_extend);
    const unsigned pad16 = (unsigned)(unsigned short)enc16(SWB_PAD_SCORE, mode);
    P.m16 = db->m16.p; P.qrow_off = db->qrow_off.p;
    P.nq = tb.nq; P.npass = 1;
    P.negq = nq16 | (nq16 << 16); P.negr = nr16 | (nr16 << 16); P.padword = pad16 | (pad16 << 16);
    P.stagger = 1;
    if (const char *env = getenv("SWB_STAGGER")) P.stagger = atoi(env) != 0;
    P.seg = segs[0];
    P.segs = db->segs.p;
    SWB_CUDA(cudaEventRecord(db->ev_batch[0], st));
    fn<<<dim3((unsigned)grid, (unsigned)work.size()), threads, smem, st>>>(P);
    SWB_CUDA(cudaGetLastError());
    SWB_CUDA(cudaEventRecord(db->ev_batch[1], st));
    for (int k = k0; k < k0 + m; k++)
    {
      Prescan pre;
      for (Layout *L : work) pre.pair_scores.push_back(L->pair_scores.p + (size_t)(k - k0) * (size_t)L->npairs);
      SWB_TRY(single(k, &pre));
    }
    float ms = 0;
    SWB_CUDA(cudaEventElapsedTime(&ms, db->ev_batch[0], db->ev_batch[1]));
    if (counters)
      for (int k = k0; k < k0 + m; k++)
      {
        counters[k].scan_ms += ms / m;             // the shared scan, split evenly over the group
        counters[k].scan_geometry = 2; counters[k].scan_G = G; counters[k].scan_R = R; counters[k].scan_passes = 1;
        counters[k].kernel_launches += k == k0 ? 1 : 0;
      }
    k0 += m;
  }
  return SWB_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int swb_search_batch(swb_db *db, int nqueries, const uint8_t *const *queries, const int64_t *qlens,
                     const swb_scoring *scoring, int64_t *const *scores, swb_counters *counters)
{
  if (nqueries > 0 && !scores) return SWB_ERR_ARG;
  return batch_impl(db, nqueries, queries, qlens, scoring, scores, nullptr, counters);
}

int swb_search_hits_batch(swb_db *db, int nqueries, const uint8_t *const *queries, const int64_t *qlens,
                          const swb_scoring *scoring, int64_t seqno_base, int64_t keep, int64_t min_score,
                          int64_t upper_score, int64_t *const *out_seqno, int64_t *const *out_score,
                          int64_t *nhits, int64_t *totalhits, int64_t *obvious, swb_counters *counters)
{
  if (nqueries < 0 || keep < 0 || (nqueries > 0 && (!nhits || (keep > 0 && (!out_seqno || !out_score)))))
    return SWB_ERR_ARG;
  std::vector<HitsReq> H((size_t)std::max(nqueries, 0));
  for (int k = 0; k < nqueries; k++)
  {
    H[(size_t)k].seqno_base = seqno_base; H[(size_t)k].keep = keep;
    H[(size_t)k].min_score = min_score; H[(size_t)k].upper = upper_score;
    H[(size_t)k].out_seqno = keep > 0 ? (long long *)out_seqno[k] : nullptr;
    H[(size_t)k].out_score = keep > 0 ? (long long *)out_score[k] : nullptr;
  }
  const int rc = batch_impl(db, nqueries, queries, qlens, scoring, nullptr, H.data(), counters);
  if (rc != SWB_OK) return rc;
  for (int k = 0; k < nqueries; k++)
  {
    nhits[k] = H[(size_t)k].nhits;
    if (totalhits) totalhits[k] = H[(size_t)k].totalhits;
    if (obvious) obvious[k] = H[(size_t)k].obvious;
  }
  return SWB_OK;
}

int swb_abi_version(void) { return SWB_ABI_VERSION; }

const char *swb_strerror(int status)
{
  switch (status)
  {
    case SWB_OK: return "ok";
    case SWB_ERR_ARG: return "invalid argument";
    case SWB_ERR_NO_DEVICE: return "no usable CUDA device (an sm_100 GPU is required; there is no CPU path)";
    case SWB_ERR_CUDA: return "CUDA runtime error";
    case SWB_ERR_NOMEM: return "out of device or host memory";
    case SWB_ERR_RANGE: return "scoring parameters out of range";
    case SWB_ERR_INTERNAL: return "internal error";
    case SWB_ERR_IO: return "database file missing, truncated or corrupt";
    default: return "unknown status";
  }
}

const char *swb_last_cuda_error(void) { return g_cuda_error.c_str(); }

int swb_device_count(int *count)
{
  if (!count) return SWB_ERR_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
  {
    g_cuda_error = cudaGetErrorString(e);
    (void)cudaGetLastError();
    *count = 0;
    return SWB_ERR_NO_DEVICE;
  }
  *count = n;
  return SWB_OK;
}

int swb_host_alloc(void **ptr, int64_t bytes)
{
  if (!ptr || bytes < 0) return SWB_ERR_ARG;
  SWB_CUDA(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)));
  return SWB_OK;
}

int swb_trim(void)
{
  g_cache.trim();
  return SWB_OK;
}

int swb_host_free(void *ptr)
{
  if (ptr) SWB_CUDA(cudaFreeHost(ptr));
  return SWB_OK;
}

// Uploads from PAGEABLE host memory (a memory-mapped database file, a plain malloc) go through a few
// pinned staging buffers filled by a handful of host threads: cudaMemcpyAsync on pageable memory
// stages through the driver at a few GB/s, which would make opening a shard cost several scans.
// Used by the synchronous opens only (the asynchronous open must return at once and documents that
// its buffers should be pinned).
struct Stager
{
  static const int NBUF = 4;
  static const size_t BUF = (size_t)32 << 20;
  std::mutex mu;
  unsigned char *buf[NBUF] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done[NBUF] = {nullptr, nullptr, nullptr, nullptr};
  bool used[NBUF] = {false, false, false, false};
  int next = 0;
  bool ready = false, broken = false;

  bool init()
  {
    if (ready || broken) return ready;
    for (int i = 0; i < NBUF; i++)
      if (cudaHostAlloc((void **)&buf[i], BUF, cudaHostAllocPortable) != cudaSuccess ||
          cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess)
      {
        (void)cudaGetLastError();
        broken = true;
        return false;
      }
    ready = true;
    return true;
  }

  static void parallel_copy(unsigned char *dst, const unsigned char *src, size_t n)
  {
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw ? hw : 1, 6), n >> 22));
    if (nt == 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nt; t++)
      pool.emplace_back([=]() { memcpy(dst + n * t / nt, src + n * t / nt, n * (t + 1) / nt - n * t / nt); });
    for (std::thread &t : pool) t.join();
  }

  // dst (device) <- src (pageable host), enqueued on st; returns a CUDA error code
  cudaError_t copy(unsigned char *dst, const unsigned char *src, size_t n, cudaStream_t st)
  {
    std::lock_guard<std::mutex> g(mu);
    if (!init()) return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
    for (size_t off = 0; off < n; off += BUF)
    {
      const size_t len = std::min(BUF, n - off);
      const int k = next;
      next = (next + 1) % NBUF;
      if (used[k])
      {
        cudaError_t e = cudaEventSynchronize(done[k]);      // the copy that last used this buffer
        if (e != cudaSuccess) return e;
      }
      parallel_copy(buf[k], src + off, len);
      cudaError_t e = cudaMemcpyAsync(dst + off, buf[k], len, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      e = cudaEventRecord(done[k], st);
      if (e != cudaSuccess) return e;
      used[k] = true;
    }
    return cudaSuccess;
  }

  // before another device's stream reuses the buffers
  void drain()
  {
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < NBUF; i++)
      if (used[i]) { cudaEventSynchronize(done[i]); used[i] = false; }
  }
};
Stager g_stagers[64];               // one per device: events and streams must share a device
inline Stager &stager_of(int device) { return g_stagers[device & 63]; }

bool is_pageable(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    (void)cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// Where the subjects of a shard come from.  `offsets` are byte offsets of the (decoded) residues
// in the device residue buffer; `extents` name the host memory holding contiguous runs of subjects:
// raw symbol bytes (copied as they are) or .nsq records (copied to the packed buffer and decoded on
// the device by swb_nt_decode_kernel).
struct Extent
{
  long long s0, s1;            // subjects [s0, s1)
  const uint8_t *src;          // host address of the first byte of subject s0 (its record, if nsq)
};
struct OpenSrc
{
  long long nseq = 0;
  int trailing = 0;
  const long long *offsets = nullptr;     // [nseq+1], offsets[0] == 0
  std::vector<long long> own_offsets;
  long long total = 0, longest = 0;
  bool nsq = false;
  std::vector<long long> pk_start;        // [nsrc+1] record offsets in the packed buffer (nsq)
  std::vector<u32> pk_len;                // [nsrc]
  std::vector<Extent> extents;            // in source sequences
  // translated shards: every nucleotide source sequence s yields the six protein subjects
  // 6 s + 3 strand + frame; offsets / nseq describe those, nt_offsets the decoded nucleotides
  bool translate = false;
  std::vector<long long> nt_offsets;      // [nsrc+1]
  const uint8_t *ttable = nullptr;        // [4096]
};

// SWB_TRACE_OPEN=1: host-side timeline of an open on stderr (tuning aid)
struct OpenTrace
{
  bool on = getenv("SWB_TRACE_OPEN") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char *what)
  {
    if (!on) return;
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[swb open] %8.1f us  %s\n", us, what);
  }
};

static int open_impl(int device, OpenSrc &S, void *stream, swb_db **out, bool wait)
{
  OpenTrace trace;
  trace.mark("open_impl");
  const long long nseq = S.nseq;
  const long long *offsets = S.offsets;
  int ndev = 0;
  SWB_TRY(swb_device_count(&ndev));
  if (device < 0 || device >= ndev) return SWB_ERR_NO_DEVICE;
  SWB_CUDA(cudaSetDevice(device));
  int cc_major = 0, cc_minor = 0, sm_count = 0;       // (cudaGetDeviceProperties costs milliseconds)
  SWB_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  SWB_CUDA(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  SWB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major != 10)
  {
    g_cuda_error = std::string("device is sm_") + std::to_string(cc_major * 10 + cc_minor) +
                   ", the kernels are built for sm_100a only";
    return SWB_ERR_NO_DEVICE;
  }
  swb_db *db = new (std::nothrow) swb_db;
  if (!db) return SWB_ERR_NOMEM;
  db->device = device;
  db->sm_count = sm_count;
  db->nseq = nseq; db->total_res = S.total; db->longest = S.longest; db->trailing = S.trailing;

  // Pipeline chunks: contiguous subject ranges of about SWB_CHUNK_BYTES of residues each, so the
  // upload of chunk c+1 overlaps the re-layout (and, once a search is issued, the scan) of chunk c.
  long long chunk_bytes = 512LL << 20;    // (measured: 87.9 ms with 512 MB chunks, 89.6 with 256 MB, 90.5 with 1 GB for the 5 M-subject scan)
  if (const char *env = getenv("SWB_CHUNK_BYTES"))      // test hook: force many small chunks
    chunk_bytes = std::max<long long>(1, atoll(env));
  const int unit = S.translate ? 6 : 1;             // subjects per source sequence
  const long long nsrc = nseq / unit;
  const long long *cut_off = S.translate ? S.nt_offsets.data() : offsets;
  // a small shard (one GPU's share of a database spread over eight) is still cut into about four chunks:
  // the scan then runs as several waves of CTAs whose ragged ends overlap (measured on a 0.6 M-subject
  // shard: 11.95 -> 11.68 ms)
  // A shard that is still arriving is cut into about eight chunks (32 to 256 MB): the scan of the LAST chunk
  // cannot start before the upload ends, so it must be a small part of the whole -- with the 256 MB chunks of
  // a big shard, one GPU's eighth of the database ended in a chunk holding more than half of it.
  if (getenv("SWB_CHUNK_BYTES") == nullptr && nsrc > 0)
  {
    if (wait && cut_off[nsrc] < 2 * chunk_bytes) chunk_bytes = std::max<long long>(32LL << 20, cut_off[nsrc] / 4 + 1);
    if (!wait) chunk_bytes = std::min<long long>(chunk_bytes, std::max<long long>(32LL << 20, cut_off[nsrc] / 8 + 1));
  }
  std::vector<long long> cut;                       // chunk c covers source sequences [cut[c], cut[c+1])
  cut.push_back(0);
  // asynchronous open: the first chunks are small so that the scan can start while most of the shard
  // is still on the wire (32, 64, 128 MB, then full size)
  const bool ramp = !wait && getenv("SWB_CHUNK_BYTES") == nullptr;   // only an asynchronous open overlaps
  while (cut.back() < nsrc)
  {
    const long long lo = cut.back();
    long long budget = chunk_bytes;
    if (ramp && cut.size() <= 3) budget = std::min<long long>(chunk_bytes, (32LL << 20) << (cut.size() - 1));
    const long long *e = std::upper_bound(cut_off + lo + 1, cut_off + nsrc + 1, cut_off[lo] + budget);
    long long hi = (long long)(e - cut_off) - 1;     // last boundary within the byte budget
    if (hi <= lo) hi = lo + 1;
    if (nsrc - hi < (hi - lo) / 4) hi = nsrc;        // do not leave a sliver behind
    cut.push_back(hi);
  }

  auto body = [&]() -> int {
    if (stream) db->stream = (cudaStream_t)stream;
    else
    {
      SWB_CUDA(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
      db->own_stream = true;
    }
    SWB_CUDA(cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking));
    SWB_CUDA(cudaStreamCreateWithFlags(&db->layout_stream, cudaStreamNonBlocking));
    SWB_CUDA(cudaStreamCreateWithFlags(&db->stream2, cudaStreamNonBlocking));
    SWB_CUDA(cudaEventCreateWithFlags(&db->ev_setup, cudaEventDisableTiming));
    for (int i = 0; i < 4; i++) SWB_CUDA(cudaEventCreate(&db->ev[i]));
    for (int i = 0; i < 2; i++) SWB_CUDA(cudaEventCreateWithFlags(&db->ev_group[i], cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) SWB_CUDA(cudaEventCreate(&db->ev_open[i]));
    for (int i = 0; i < 2; i++) SWB_CUDA(cudaEventCreate(&db->ev_batch[i]));
    SWB_CUDA(cudaEventCreateWithFlags(&db->ev_uploaded, cudaEventDisableTiming));
    trace.mark("streams and events created");
    SWB_TRY(db->residues.reserve((size_t)offsets[nseq] + 16));
    SWB_TRY(db->offsets.reserve((size_t)nseq + 1));
    SWB_CUDA(cudaEventRecord(db->ev_open[0], db->copy_stream));
    SWB_CUDA(cudaMemcpyAsync(db->offsets.p, offsets, ((size_t)nseq + 1) * sizeof(long long),
                             cudaMemcpyHostToDevice, db->copy_stream));
    // host-side coordinates of the bytes that are copied: residue offsets, or .nsq record offsets
    const long long *coord = offsets;
    unsigned char *dst_base = db->residues.p;
    if (S.nsq)
    {
      coord = S.pk_start.data();
      SWB_TRY(db->packed.reserve((size_t)coord[nsrc] + 16));
      SWB_TRY(db->pk_start.reserve((size_t)nsrc + 1));
      SWB_TRY(db->pk_len.reserve((size_t)std::max<long long>(nsrc, 1)));
      SWB_CUDA(cudaMemcpyAsync(db->pk_start.p, coord, ((size_t)nsrc + 1) * sizeof(long long),
                               cudaMemcpyHostToDevice, db->copy_stream));
      SWB_CUDA(cudaMemcpyAsync(db->pk_len.p, S.pk_len.data(), (size_t)nsrc * sizeof(u32),
                               cudaMemcpyHostToDevice, db->copy_stream));
      dst_base = db->packed.p;
    }
    if (S.translate)
    {
      SWB_TRY(db->nt_residues.reserve((size_t)S.nt_offsets[(size_t)nsrc] + 16));
      SWB_TRY(db->nt_offsets.reserve((size_t)nsrc + 1));
      SWB_TRY(db->ttable.reserve(4096));
      SWB_CUDA(cudaMemcpyAsync(db->nt_offsets.p, S.nt_offsets.data(), ((size_t)nsrc + 1) * sizeof(long long),
                               cudaMemcpyHostToDevice, db->copy_stream));
      SWB_CUDA(cudaMemcpyAsync(db->ttable.p, S.ttable, 4096, cudaMemcpyHostToDevice, db->copy_stream));
    }
    // pageable vectors owned by the caller's frame must be consumed before it returns
    if (!S.own_offsets.empty() || S.nsq) SWB_CUDA(cudaStreamSynchronize(db->copy_stream));
    size_t ext = 0;
    for (size_t c = 0; c + 1 < cut.size(); c++)
    {
      Layout *L = new (std::nothrow) Layout;
      if (!L) return SWB_ERR_NOMEM;
      db->chunks.push_back(L);
      const long long lo = cut[c], hi = cut[c + 1];
      while (ext < S.extents.size() && S.extents[ext].s1 <= lo) ext++;
      for (size_t x = ext; x < S.extents.size() && S.extents[x].s0 < hi; x++)
      {
        const Extent &E = S.extents[x];
        const long long a = std::max(lo, E.s0), b = std::min(hi, E.s1);
        const long long bytes = coord[b] - coord[a];
        if (bytes > 0)
        {
          const uint8_t *src = E.src + (coord[a] - coord[E.s0]);
          if (wait && bytes >= (8 << 20) && getenv("SWB_NO_STAGING") == nullptr && is_pageable(src))
            SWB_CUDA(stager_of(db->device).copy(dst_base + coord[a], src, (size_t)bytes, db->copy_stream));
          else
            SWB_CUDA(cudaMemcpyAsync(dst_base + coord[a], src, (size_t)bytes, cudaMemcpyHostToDevice,
                                     db->copy_stream));
        }
      }
      SWB_CUDA(cudaEventCreateWithFlags(&L->ev_ready, cudaEventDisableTiming));
      cudaEvent_t up;
      SWB_CUDA(cudaEventCreateWithFlags(&up, cudaEventDisableTiming));
      SWB_CUDA(cudaEventRecord(up, db->copy_stream));
      SWB_CUDA(cudaStreamWaitEvent(db->layout_stream, up, 0));
      SWB_CUDA(cudaEventDestroy(up));             // released once the wait has consumed it
      if (c == 0) SWB_CUDA(cudaEventRecord(db->ev_open[1], db->layout_stream));
      if (S.nsq && hi > lo)
      {
        const long long warps = hi - lo;
        const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
        swb_nt_decode_kernel<<<grid, 256, 0, db->layout_stream>>>(
            db->packed.p, db->pk_start.p, db->pk_len.p, S.translate ? db->nt_offsets.p : db->offsets.p,
            S.translate ? db->nt_residues.p : db->residues.p, lo, hi - lo);
        SWB_CUDA(cudaGetLastError());
        if (S.translate)
        {
          swb_translate_kernel<<<grid, 256, 0, db->layout_stream>>>(
              db->nt_residues.p, db->nt_offsets.p, db->ttable.p, db->offsets.p, db->residues.p, lo, hi - lo);
          SWB_CUDA(cudaGetLastError());
        }
      }
      SWB_TRY(build_layout(db, *L, nullptr, lo * unit, (hi - lo) * unit,
                           offsets[hi * unit] - offsets[lo * unit], db->layout_stream));
      SWB_CUDA(cudaEventRecord(L->ev_ready, db->layout_stream));
    }
    // "uploaded" = the residue buffer is complete (for nsq: decoded), which the wide kernel needs
    if (S.nsq)
    {
      cudaEvent_t done;
      SWB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      SWB_CUDA(cudaEventRecord(done, db->layout_stream));
      SWB_CUDA(cudaStreamWaitEvent(db->copy_stream, done, 0));
      SWB_CUDA(cudaEventDestroy(done));
    }
    trace.mark("all chunks enqueued");
    SWB_CUDA(cudaEventRecord(db->ev_uploaded, db->copy_stream));
    SWB_CUDA(cudaEventRecord(db->ev_open[2], db->layout_stream));
    if (db->chunks.empty()) SWB_CUDA(cudaEventRecord(db->ev_open[1], db->layout_stream));
    if (wait)
    {
      SWB_TRY(swb_db_wait(db));
      stager_of(db->device).drain();
    }
    return SWB_OK;
  };
  const int rc = body();
  if (rc != SWB_OK)
  {
    swb_db_close(db);
    return rc;
  }
  *out = db;
  return SWB_OK;
}

// raw symbol bytes + offsets, as swb_db_open documents
static int open_raw(int device, const uint8_t *residues, const int64_t *offsets, int64_t nseq,
                    int trailing, void *stream, swb_db **out, bool wait)
{
  if (!out) return SWB_ERR_ARG;
  *out = nullptr;
  if (nseq < 0 || !offsets || trailing < 0 || trailing > 1 || nseq > 0x7ffffff0LL) return SWB_ERR_ARG;
  const long long span = offsets[nseq] - offsets[0];
  if (span < 0 || (span > 0 && !residues)) return SWB_ERR_ARG;
  OpenSrc S;
  S.nseq = nseq;
  S.trailing = trailing;
  {
    // one pass over the offsets (range check, total, longest), cut over a few host threads when the
    // shard is large: at 5 M subjects a serial loop costs more host time than enqueueing the upload
    struct Acc { long long total = 0, longest = 0, shortest = 0; };
    auto pass = [&](Acc &a, long long lo, long long hi) {
      long long tot = 0, mx = 0, mn = 0;
      for (long long i = lo; i < hi; i++)
      {
        const long long len = offsets[i + 1] - offsets[i] - trailing;
        tot += len;
        mx = len > mx ? len : mx;
        mn = len < mn ? len : mn;
      }
      a.total = tot; a.longest = mx; a.shortest = mn;
    };
    unsigned hw = std::thread::hardware_concurrency();
    const long long nt = std::max<long long>(1, std::min<long long>(std::min<long long>(hw ? hw : 1, 8), nseq >> 19));
    std::vector<Acc> acc((size_t)nt);
    if (nt == 1) pass(acc[0], 0, nseq);
    else
    {
      std::vector<std::thread> pool;
      for (long long t = 0; t < nt; t++)
        pool.emplace_back([&, t]() { pass(acc[(size_t)t], nseq * t / nt, nseq * (t + 1) / nt); });
      for (std::thread &t : pool) t.join();
    }
    for (const Acc &a : acc)
    {
      if (a.shortest < 0 || a.longest > 0x7fffffffLL) return SWB_ERR_ARG;
      S.total += a.total;
      S.longest = std::max(S.longest, a.longest);
    }
  }
  S.offsets = (const long long *)offsets;
  if (offsets[0] != 0)                             // rebase so that residues[0] is the first byte uploaded
  {
    S.own_offsets.resize((size_t)nseq + 1);
    for (long long i = 0; i <= nseq; i++) S.own_offsets[(size_t)i] = offsets[i] - offsets[0];
    S.offsets = S.own_offsets.data();
  }
  if (nseq > 0) S.extents.push_back(Extent{0, nseq, residues + offsets[0]});
  return open_impl(device, S, stream, out, wait);
}

int swb_db_open(int device, const uint8_t *residues, const int64_t *offsets, int64_t nseq,
                int trailing, void *stream, swb_db **out)
{
  return open_raw(device, residues, offsets, nseq, trailing, stream, out, true);
}

int swb_db_open_async(int device, const uint8_t *residues, const int64_t *offsets, int64_t nseq,
                      int trailing, void *stream, swb_db **out)
{
  return open_raw(device, residues, offsets, nseq, trailing, stream, out, false);
}

// Upload subjects [first, first + count) of a BLAST database (all of it with count < 0): protein
// volumes are copied as they lie in the .psq (NUL separated -> trailing = 1); nucleotide volumes are
// copied packed (4 bases per byte + ambiguity tables) and unpacked on the device to the 4-bit codes
// db_getsequence produces (database.cc:1257-1323).
int swb_db_open_blast(int device, const swb_blastdb *b, int64_t first, int64_t count, int async,
                      void *stream, swb_db **out)
{
  if (!out) return SWB_ERR_ARG;
  *out = nullptr;
  if (!b || first < 0 || first > b->nseq) return SWB_ERR_ARG;
  if (count < 0 || first + count > b->nseq) count = b->nseq - first;
  if (count > 0x7ffffff0LL) return SWB_ERR_ARG;
  OpenSrc S;
  S.nseq = count;
  S.nsq = b->nucleotide;
  S.trailing = b->nucleotide ? 0 : 1;
  S.own_offsets.resize((size_t)count + 1);
  if (S.nsq)
  {
    S.pk_start.resize((size_t)count + 1);
    S.pk_len.resize((size_t)std::max<int64_t>(count, 1));
  }
  long long k = 0, off = 0, pk = 0;
  for (const SwbVolume &v : b->vols)
  {
    const long long a = std::max<long long>(first, v.first) - v.first;
    const long long e = std::min<long long>(first + count, v.first + v.nseq) - v.first;
    if (e <= a) continue;
    S.extents.push_back(Extent{k, k + (e - a), v.seq + v.seq_off(a)});
    for (long long s = a; s < e; s++, k++)
    {
      const long long o1 = v.seq_off(s), o2 = v.seq_off(s + 1);
      long long len;
      if (S.nsq)
      {
        len = swb_nt_length(v, s);
        S.pk_start[(size_t)k] = pk;
        S.pk_len[(size_t)k] = (u32)(v.amb_off(s) - o1);
        pk += o2 - o1;
        S.own_offsets[(size_t)k] = off;
        off += len;
      }
      else
      {
        len = o2 - o1 - 1;
        if (len < 0) return SWB_ERR_ARG;
        S.own_offsets[(size_t)k] = off;
        off += o2 - o1;
      }
      S.total += len;
      S.longest = std::max(S.longest, len);
    }
  }
  S.own_offsets[(size_t)count] = off;
  if (S.nsq) S.pk_start[(size_t)count] = pk;
  S.offsets = S.own_offsets.data();
  return open_impl(device, S, stream, out, async == 0);
}

// A nucleotide database as six-frame translated protein subjects (the reference's -p 3 / -p 4,
// db_translate database.cc:1182-1218): sequence s of the shard becomes subjects 6 s + 3 strand +
// frame (the order search_chunk lists them in, swipe.cc:1377-1385), translated ON THE DEVICE with
// the 4096-entry codon table of swb_translate_table.
int swb_db_open_blast_translated(int device, const swb_blastdb *b, int64_t first, int64_t count,
                                 const uint8_t *codon_table, int async, void *stream, swb_db **out)
{
  if (!out) return SWB_ERR_ARG;
  *out = nullptr;
  if (!b || !b->nucleotide || !codon_table || first < 0 || first > b->nseq) return SWB_ERR_ARG;
  if (count < 0 || first + count > b->nseq) count = b->nseq - first;
  if (count > 0x7ffffff0LL / 6) return SWB_ERR_ARG;
  OpenSrc S;
  S.nseq = 6 * count;
  S.nsq = true;
  S.translate = true;
  S.trailing = 0;
  S.ttable = codon_table;
  S.own_offsets.resize((size_t)(6 * count) + 1);
  S.nt_offsets.resize((size_t)count + 1);
  S.pk_start.resize((size_t)count + 1);
  S.pk_len.resize((size_t)std::max<int64_t>(count, 1));
  long long k = 0, off = 0, ntoff = 0, pk = 0;
  for (const SwbVolume &v : b->vols)
  {
    const long long a = std::max<long long>(first, v.first) - v.first;
    const long long e = std::min<long long>(first + count, v.first + v.nseq) - v.first;
    if (e <= a) continue;
    S.extents.push_back(Extent{k, k + (e - a), v.seq + v.seq_off(a)});
    for (long long s = a; s < e; s++, k++)
    {
      const long long o1 = v.seq_off(s), o2 = v.seq_off(s + 1);
      const long long len = swb_nt_length(v, s);
      S.pk_start[(size_t)k] = pk;
      S.pk_len[(size_t)k] = (u32)(v.amb_off(s) - o1);
      pk += o2 - o1;
      S.nt_offsets[(size_t)k] = ntoff;
      ntoff += len;
      for (int f = 0; f < 6; f++)
      {
        const long long plen = len - (f % 3) >= 0 ? (len - (f % 3)) / 3 : 0;
        S.own_offsets[(size_t)(6 * k + f)] = off;
        off += plen;
        S.total += plen;
        S.longest = std::max(S.longest, plen);
      }
    }
  }
  S.own_offsets[(size_t)(6 * count)] = off;
  S.nt_offsets[(size_t)count] = ntoff;
  S.pk_start[(size_t)count] = pk;
  S.offsets = S.own_offsets.data();
  return open_impl(device, S, stream, out, async == 0);
}

int swb_db_wait(swb_db *db)
{
  if (!db) return SWB_ERR_ARG;
  SWB_CUDA(cudaSetDevice(db->device));
  SWB_CUDA(cudaStreamSynchronize(db->copy_stream));
  SWB_CUDA(cudaStreamSynchronize(db->layout_stream));
  if (!db->opened_sync)
  {
    float ms = 0;
    SWB_CUDA(cudaEventSynchronize(db->ev_open[2]));
    SWB_CUDA(cudaEventElapsedTime(&ms, db->ev_open[0], db->ev_open[2]));
    db->upload_ms = ms;                      // first byte sent .. last layout done (pipelined)
    SWB_CUDA(cudaEventElapsedTime(&ms, db->ev_open[1], db->ev_open[2]));
    db->layout_ms = ms;
    db->opened_sync = true;
  }
  return SWB_OK;
}

int swb_db_close(swb_db *db)
{
  if (!db) return SWB_OK;
  cudaSetDevice(db->device);
  if (db->copy_stream) cudaStreamSynchronize(db->copy_stream);
  if (db->layout_stream) cudaStreamSynchronize(db->layout_stream);
  if (db->stream2) cudaStreamSynchronize(db->stream2);
  if (db->stream) cudaStreamSynchronize(db->stream);
  db->residues.release(); db->offsets.release(); db->tmp.release(); db->tmp2.release();
  db->requeue2.release(); db->codes.release();
  db->packed.release(); db->pk_start.release(); db->pk_len.release();
  db->nt_residues.release(); db->nt_offsets.release(); db->ttable.release();
  for (Layout *L : db->chunks) { L->release(); delete L; }
  db->chunks.clear();
  db->m16.release(); db->qrow_off.release(); db->matrix.release(); db->query.release();
  db->scores.release(); db->bestpos.release(); db->bestq.release(); db->requeue.release();
  db->list.release(); db->counters.release(); db->he.release(); db->bndH.release();
  db->bndF.release(); db->segs.release();
  db->slot_flags.release(); db->filter.release();
  db->hist.release(); db->cand.release(); db->cand_sorted.release(); db->sort_tmp.release();
  for (int i = 0; i < 4; i++)
    if (db->ev[i]) cudaEventDestroy(db->ev[i]);
  for (int i = 0; i < 3; i++)
    if (db->ev_open[i]) cudaEventDestroy(db->ev_open[i]);
  for (int i = 0; i < 2; i++)
    if (db->ev_group[i]) cudaEventDestroy(db->ev_group[i]);
  for (int i = 0; i < 2; i++)
    if (db->ev_batch[i]) cudaEventDestroy(db->ev_batch[i]);
  if (db->ev_uploaded) cudaEventDestroy(db->ev_uploaded);
  if (db->copy_stream) cudaStreamDestroy(db->copy_stream);
  if (db->layout_stream) cudaStreamDestroy(db->layout_stream);
  if (db->stream2) { cudaStreamSynchronize(db->stream2); cudaStreamDestroy(db->stream2); }
  if (db->ev_setup) cudaEventDestroy(db->ev_setup);
  if (db->own_stream && db->stream) cudaStreamDestroy(db->stream);
  delete db;
  (void)cudaGetLastError();
  return SWB_OK;
}

int swb_db_info(const swb_db *db, int64_t *nseq, int64_t *total_residues, int64_t *longest)
{
  if (!db) return SWB_ERR_ARG;
  if (nseq) *nseq = db->nseq;
  if (total_residues) *total_residues = db->total_res;
  if (longest) *longest = db->longest;
  return SWB_OK;
}

int swb_search(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
               int64_t *scores, swb_counters *counters)
{
  if (!db) return SWB_ERR_ARG;
  return search_impl(db, query, qlen, scoring, nullptr, db->nseq, (long long *)scores, nullptr,
                     nullptr, counters);
}

int swb_search_list(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                    const int64_t *seqnos, int64_t n, int64_t *scores, swb_counters *counters)
{
  if (!db || (n > 0 && !seqnos)) return SWB_ERR_ARG;
  return search_impl(db, query, qlen, scoring, (const long long *)seqnos, n, (long long *)scores,
                     nullptr, nullptr, counters);
}

int swb_search_hits(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                    int64_t seqno_base, int64_t keep, int64_t min_score, int64_t upper_score,
                    int64_t *out_seqno, int64_t *out_score, int64_t *nhits, int64_t *totalhits,
                    int64_t *obvious, swb_counters *counters)
{
  if (!db || keep < 0 || (keep > 0 && (!out_seqno || !out_score)) || !nhits) return SWB_ERR_ARG;
  HitsReq H;
  H.seqno_base = seqno_base; H.keep = keep; H.min_score = min_score; H.upper = upper_score;
  H.out_seqno = (long long *)out_seqno; H.out_score = (long long *)out_score;
  const int rc = search_impl(db, query, qlen, scoring, nullptr, db->nseq, nullptr, nullptr, nullptr,
                             counters, &H);
  if (rc != SWB_OK) return rc;
  *nhits = H.nhits;
  if (totalhits) *totalhits = H.totalhits;
  if (obvious) *obvious = H.obvious;
  return SWB_OK;
}

int swb_db_set_filter(swb_db *db, const uint8_t *bitmap)
{
  if (!db) return SWB_ERR_ARG;
  SWB_CUDA(cudaSetDevice(db->device));
  SWB_CUDA(cudaStreamSynchronize(db->stream));              // no search of this handle is using the old one
  if (!bitmap)
  {
    db->has_filter = false;
    return SWB_OK;
  }
  const size_t bytes = (size_t)((db->nseq + 7) / 8);
  db->h_filter.assign(bitmap, bitmap + bytes);
  SWB_TRY(db->filter.reserve(std::max<size_t>(bytes, 1)));
  if (bytes)
    SWB_CUDA(cudaMemcpyAsync(db->filter.p, db->h_filter.data(), bytes, cudaMemcpyHostToDevice, db->stream));
  SWB_CUDA(cudaStreamSynchronize(db->stream));
  db->has_filter = true;
  return SWB_OK;
}

int64_t swb_hits_merge(int nlists, const int64_t *const *seqnos, const int64_t *const *scores,
                       const int64_t *n, int64_t keep, int64_t *out_seqno, int64_t *out_score)
{
  if (nlists < 0 || keep < 0 || (nlists > 0 && (!seqnos || !scores || !n))) return SWB_ERR_ARG;
  if (keep > 0 && (!out_seqno || !out_score)) return SWB_ERR_ARG;
  // every list is already in the sink's order (score descending, then sequence number descending,
  // hits.cc:188-191): a k-way merge by repeated selection of the best head; the lists are short
  std::vector<int64_t> head((size_t)nlists, 0);
  for (int s = 0; s < nlists; s++)
    if (n[s] < 0 || (n[s] > 0 && (!seqnos[s] || !scores[s]))) return SWB_ERR_ARG;
  int64_t out = 0;
  while (out < keep)
  {
    int best = -1;
    for (int s = 0; s < nlists; s++)
    {
      if (head[(size_t)s] >= n[s]) continue;
      if (best < 0) { best = s; continue; }
      const int64_t a = scores[s][head[(size_t)s]], b = scores[best][head[(size_t)best]];
      if (a > b || (a == b && seqnos[s][head[(size_t)s]] > seqnos[best][head[(size_t)best]])) best = s;
    }
    if (best < 0) break;
    out_seqno[out] = seqnos[best][head[(size_t)best]];
    out_score[out] = scores[best][head[(size_t)best]];
    head[(size_t)best]++;
    out++;
  }
  return out;
}

int swb_set_cache_limit(int64_t bytes_per_device)
{
  if (bytes_per_device < 0) return SWB_ERR_ARG;
  {
    std::lock_guard<std::mutex> g(g_cache.mu);
    g_cache.limit = bytes_per_device;
  }
  return SWB_OK;
}

int swb_search_end(swb_db *db, const uint8_t *query, int64_t qlen, const swb_scoring *scoring,
                   const int64_t *seqnos, int64_t n, int64_t *scores, int64_t *bestpos,
                   int64_t *bestq)
{
  if (!db || (n > 0 && (!seqnos || !bestpos || !bestq))) return SWB_ERR_ARG;
  long long dummy = 0;
  return search_impl(db, query, qlen, scoring, (const long long *)seqnos, n, (long long *)scores,
                     n > 0 ? (long long *)bestpos : &dummy, n > 0 ? (long long *)bestq : &dummy,
                     nullptr);
}

int64_t swb_topk_merge(int nshards, const int64_t *const *scores, const int64_t *n,
                       const int64_t *seqno_base, int64_t keep, int64_t min_score,
                       int64_t upper_score, int64_t *out_seqno, int64_t *out_score,
                       int64_t *totalhits, int64_t *obvious)
{
  if (nshards < 0 || keep < 0 || (nshards > 0 && (!scores || !n || !seqno_base))) return SWB_ERR_ARG;
  if (keep > 0 && (!out_seqno || !out_score)) return SWB_ERR_ARG;
  for (int s = 0; s < nshards; s++)
    if (n[s] < 0 || (n[s] > 0 && !scores[s])) return SWB_ERR_ARG;
  // hits_enter keeps (score desc, seqno desc) and, once full, only admits scores >= the last
  // kept one; the final list therefore is the top `keep` of the admissible hits in that order,
  // whatever the arrival order -- which is what makes the multi-GPU merge (and the split of one
  // shard over host threads below) deterministic.
  struct Hit { int64_t score, seqno; };
  auto worse = [](const Hit &a, const Hit &b) {
    return a.score != b.score ? a.score > b.score : a.seqno > b.seqno;
  };
  struct Part { std::vector<Hit> heap; int64_t tot = 0, obv = 0; };
  auto scan = [&](Part &P, const int64_t *sc, int64_t lo, int64_t hi, int64_t base) {
    std::vector<Hit> heap;                           // min-heap on (score, seqno) of the best so far
    heap.reserve((size_t)std::min<int64_t>(keep, hi - lo) + 1);
    int64_t tot = 0, obv = 0;
    int64_t cut = min_score;                         // scores below cannot enter (raised once the heap is full)
    // highest sequence number first: among equal scores the higher number wins (hits.cc:188-191), so
    // in this order a tie never displaces an entry and the heap only changes on a better score
    for (int64_t i = hi - 1; i >= lo; i--)
    {
      const int64_t v = sc[i];
      tot += v >= min_score;
      obv += v > upper_score;
      if (v < cut || v > upper_score || keep == 0) continue;
      const Hit h = {v, base + i};
      if ((int64_t)heap.size() < keep)
      {
        heap.push_back(h);
        std::push_heap(heap.begin(), heap.end(), worse);
      }
      else if (worse(h, heap.front()))
      {
        std::pop_heap(heap.begin(), heap.end(), worse);
        heap.back() = h;
        std::push_heap(heap.begin(), heap.end(), worse);
      }
      if ((int64_t)heap.size() == keep) cut = std::max(cut, heap.front().score);
    }
    P.heap.swap(heap);
    P.tot = tot;
    P.obv = obv;
  };
  // large shards are cut over a few host threads; every piece keeps its own top `keep`
  struct Piece { int shard; int64_t lo, hi; };
  std::vector<Piece> pieces;
  const int64_t grain = 1 << 18;
  unsigned hw = std::thread::hardware_concurrency();
  const int64_t maxpar = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, 8));
  for (int s = 0; s < nshards; s++)
  {
    const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(maxpar, n[s] / grain));
    for (int64_t k = 0; k < parts; k++) pieces.push_back(Piece{s, n[s] * k / parts, n[s] * (k + 1) / parts});
  }
  std::vector<Part> parts(pieces.size());
  if (pieces.size() <= 1)
  {
    for (size_t k = 0; k < pieces.size(); k++)
      scan(parts[k], scores[pieces[k].shard], pieces[k].lo, pieces[k].hi, seqno_base[pieces[k].shard]);
  }
  else
  {
    std::vector<std::thread> pool;
    std::atomic<size_t> next(0);
    const size_t nthreads = std::min<size_t>(pieces.size(), (size_t)maxpar);
    for (size_t t = 0; t < nthreads; t++)
      pool.emplace_back([&]() {
        for (size_t k = next++; k < pieces.size(); k = next++)
          scan(parts[k], scores[pieces[k].shard], pieces[k].lo, pieces[k].hi, seqno_base[pieces[k].shard]);
      });
    for (std::thread &t : pool) t.join();
  }
  std::vector<Hit> all;
  int64_t tot = 0, obv = 0;
  for (Part &P : parts)
  {
    all.insert(all.end(), P.heap.begin(), P.heap.end());
    tot += P.tot;
    obv += P.obv;
  }
  std::sort(all.begin(), all.end(), worse);
  if ((int64_t)all.size() > keep) all.resize((size_t)keep);
  for (size_t k = 0; k < all.size(); k++)
  {
    out_seqno[k] = all[k].seqno;
    out_score[k] = all[k].score;
  }
  if (totalhits) *totalhits = tot;
  if (obvious) *obvious = obv;
  return (int64_t)all.size();
}

int swb_set_mode(swb_db *db, int mode)
{
  if (!db || (mode != 0 && mode != 2)) return SWB_ERR_ARG;
  db->mode = mode;
  return SWB_OK;
}

int swb_db_open_ms(const swb_db *db, double *upload_ms, double *layout_ms)
{
  if (!db) return SWB_ERR_ARG;
  SWB_TRY(swb_db_wait(const_cast<swb_db *>(db)));
  if (upload_ms) *upload_ms = db->upload_ms;
  if (layout_ms) *layout_ms = db->layout_ms;
  return SWB_OK;
}

/* test hook: pin the scan kernel shape (G threads per stream, R rows per thread) and lane
   arithmetic (0 = int16 DPX only, 1 = DPX + fp16-pattern adds); zeros / -1 restore the default. */
int swb_set_geometry(swb_db *db, int geometry)
{
  if (!db || geometry < 0 || geometry > 2) return SWB_ERR_ARG;
  db->force_geom = geometry;
  return SWB_OK;
}

int swb_set_shape(swb_db *db, int G, int R, int lane_mode)
{
  if (!db) return SWB_ERR_ARG;
  if (G != 0)
  {
    bool found = false;
    for (int i = 0; i < g_nshapes; i++)
      found = found || (g_shapes[i].G == G && g_shapes[i].R == R &&
                        (db->force_geom == 0 || g_shapes[i].geom == db->force_geom));
    if (!found) return SWB_ERR_ARG;
  }
  if (lane_mode < -1 || lane_mode > 1) return SWB_ERR_ARG;
  db->force_G = G; db->force_R = R; db->force_mode = lane_mode;
  return SWB_OK;
}

}  // extern "C"
