// sw_kernels.cuh -- sm_100a kernels of the score-only affine-gap Smith-Waterman scan.
//
// Reference behaviour being reproduced (torognes/swipe): the recurrence of fullsw
// (search63.cc:28-89) evaluated for every database subject, which is what the reference's
// 7-bit -> 16-bit -> 63-bit cascade (search7.cc:755-958, search16.cc:320-546,
// swipe.cc:1416-1594) delivers.  Nothing here is derived from the reference's SSE code; the
// design is a systolic, register-resident scan built for the B200 SM:
//
//   * two subjects share every 32-bit register as 16-bit lanes (DPX s16x2 instructions:
//     VIADDMNMX / VIMNMX3 do add+max and 3-way max in one issue);
//   * a group of G threads owns one *stream* of subject pairs; thread g keeps the H and E
//     values of query rows [g*R, g*R+R) in registers for the whole scan (no spill of the DP
//     column), and the group works as a pipeline: at step t thread g processes the 4-column
//     block t-g of the stream, handing the bottom H/F of its strip to thread g+1 by shuffle;
//   * subjects follow each other in the stream without draining the pipeline (a START flag on
//     a block resets the strip), so the fill/drain cost is paid once per stream, not per subject;
//   * substitution scores come from a per-block table built once per 4 columns in shared memory
//     (row = query symbol, 4 words = the 4 columns, each word = the two lanes' scores) and
//     shared by all G threads, so the inner loop does one conflict-free LDS.128 per 4 cell
//     pairs and no byte shuffling;
//   * queries longer than G*R rows are scanned in passes; the bottom row of a pass is kept in
//     global memory (32 B per block) and fed to thread 0 in the next pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned int u32;
typedef unsigned long long u64;

#define SWB_PAD_CODE 32      // internal subject symbol for padding columns (scores SWB_PAD_SCORE)
#define SWB_PAD_SCORE (-1)
#define SWB_MROWS 33         // 32 symbol codes + the pad code
#define SWB_SMEM_HEADER 2304  // the staged [33][34] s16 score matrix, rounded up to 128 B
#define SWB_M16_BYTES 2256    // its image in global memory: 33 * 34 halfwords padded to a multiple of 16 B
#define SWB_FLAG_START 1u
#define SWB_FLAG_END 2u
#ifndef SWB_BUILD_BATCH
#define SWB_BUILD_BATCH 1    // table rows whose loads are in flight together (geometry 1); > 1 is an experiment
#endif

enum { SWB_MODE_INT16 = 0, SWB_MODE_HYBRID = 1 };

// One re-laid-out chunk of the shard.  A launch covers one chunk (ScanParams::seg) or, once the
// whole shard is resident, all of them at once: blockIdx.y picks the chunk from ScanParams::segs,
// so the SMs run from one chunk into the next without a per-launch tail.
struct ScanSeg
{
  const uint2 *blocks;        // [total_blocks] 8 residues each: bytes 0-3 lane A cols 0-3, 4-7 lane B
  const long long *pairblk;   // [npairs+1] exclusive prefix of blocks per pair
  const int *stream_pair;     // [nstreams+1] first pair of every stream
  u32 *pair_scores;           // [npairs] packed lane maxima (a batch: [queries][score_stride])
  long long score_stride;     // batched queries: distance between two queries' pair_scores
  long long bnd_base;         // shard-sized multi-pass scratch: first entry of this chunk in bndH / bndF
};

struct ScanParams
{
  ScanSeg seg;
  const ScanSeg *segs;        // [gridDim.y] device array, or NULL: use seg
  const short *m16;           // [33][34] (SWB_M16_BYTES) score of (subject code, table row) in the mode's
                              // encoding, laid out as it is staged in shared memory
  const unsigned short *qrow_off; // [npass*G*R] 16 * (table row of every query row)
  // Multi-pass scans: the bottom row of a pass waits here for the next pass, 2 x 16 B per block.  Two
  // layouts: (nslots == 0) one entry per block of the shard, indexed like the blocks -- the faster one
  // (measured: 263 vs 278 ms for 1000 aa x 5 M subjects), 32 B per 8 residue bytes; (nslots > 0) sized for
  // the RESIDENT CTAs only, for shards whose full scratch would not fit: a CTA claims one of `nslots`
  // regions of bnd_cta entries (bnd_stream per stream) when it starts and gives it back when it ends.
  uint4 *bndH;                // bottom H of a pass (only when npass > 1)
  uint4 *bndF;
  int *slot_flags;            // [nslots] 0 = free
  int nslots;
  long long bnd_cta, bnd_stream;
  int nq;                     // table rows in use (distinct query symbols)
  int npass;
  u32 negq;                   // both lanes: -(gap open + extend) in the mode's encoding
  u32 negr;                   // both lanes: -(gap extend), two's complement
  u32 padword;                // both lanes: score of a padding query row
  int stagger;                // geometry 2: odd stages build the next block's tables after their tile
  // Batched queries (geometry 2, single pass): several queries lie one after the other in the pipeline,
  // each on whole stages.  Bit g of start_mask: stage g holds the first rows of a query (its input is
  // zeros, not the previous stage's bottom row); bit g of end_mask: stage g holds a query's last rows and
  // stores its scores; stage_query[g]: which query that is.  A single query: 1, 1 << (G - 1), zeros.
  u32 start_mask, end_mask;
  unsigned char stage_query[32];
};

__device__ __forceinline__ u32 swb_hadd2(u32 a, u32 b)
{
  u32 r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// relu(a + b) per fp16 lane: fma.rn.relu(a, 1.0, b).  Exact on the integer bit patterns like swb_hadd2.
__device__ __forceinline__ u32 swb_hadd2_relu(u32 a, u32 b)
{
  u32 r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3c003c00u), "r"(b));
  return r;
}
// fp16x2 maximum: on bit patterns of non-negative integers (and sign-magnitude negatives) this is
// the integer maximum; a 2-input op that issues at twice the rate of the 3-input DPX forms.
__device__ __forceinline__ u32 swb_hmax2(u32 a, u32 b)
{
  u32 r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// Shared-memory access by 32-bit shared-window address (no generic-pointer arithmetic, no
// alignment masks in the instruction stream).  The "memory" clobber keeps them ordered against
// __syncthreads at the compiler level; ptxas schedules them like any other LDS/STS.
__device__ __forceinline__ uint4 swb_lds128(u32 a)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ u32 swb_lds32(u32 a)
{
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ u32 swb_lds16(u32 a)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void swb_sts128(u32 a, uint4 v)
{
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void swb_sts32(u32 a, u32 v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void swb_sts16(u32 a, unsigned short v)
{
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}

// database blocks are read-only for the kernel's lifetime: non-coherent global loads (the chunk
// table is read through a generic pointer, so the compiler cannot infer the state space itself)
__device__ __forceinline__ uint2 swb_ldg_blk(const uint2 *p)
{
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(__cvta_generic_to_global(p)));
  return v;
}

// lo | hi << 16 as an integer multiply-add: runs on the FMA pipe, which has room, instead of a PRMT
// on the ALU pipe, which is the bottleneck of the scan.
__device__ __forceinline__ u32 swb_pack16(u32 lo, u32 hi)
{
  u32 r;
  asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
  return r;
}

// Multi-pass scratch: claim / release a region (see ScanParams::slot_flags).  One thread per CTA calls these.
__device__ __forceinline__ int swb_claim_slot(const ScanParams &P)
{
  int s = (int)((blockIdx.x + blockIdx.y * gridDim.x) % (unsigned)P.nslots);
  for (long long probes = 0;; probes++)
  {
    if (atomicCAS(P.slot_flags + s, 0, 1) == 0) return s;
    s = s + 1 == P.nslots ? 0 : s + 1;
    if (probes > (1ll << 26)) __trap();            // more CTAs resident than regions: never hang the GPU
  }
}
__device__ __forceinline__ void swb_release_slot(const ScanParams &P, int s)
{
  __threadfence();
  atomicExch(P.slot_flags + s, 0);
}

// One DP cell for both lanes.  hd = H(i-1,j-1), s = score word, e = E(i,j), f = F(i,j).
// Produces h = H(i,j) and advances e -> E(i,j+1), f -> F(i+1,j); smax accumulates max H.
template <int MODE>
__device__ __forceinline__ void swb_cell(u32 hd, u32 s, u32 &e, u32 &f, u32 &h, u32 &smax,
                                         const u32 negq, const u32 negr)
{
  if (MODE == SWB_MODE_INT16)
  {
    u32 t = __viaddmax_s16x2(hd, s, e);        // max(hd + s, e)
    h = __vimax_s16x2_relu(t, f);              // max(.., f, 0)
    smax = __vmaxs2(smax, h);
    u32 hq = __vadd2(h, negq);                 // h - (open + extend)
    e = __viaddmax_s16x2(e, negr, hq);         // max(e - extend, h - open - extend)
    f = __viaddmax_s16x2(f, negr, hq);
  }
  else
  {
    // Values 0..2047 are their own fp16 bit patterns (value n * 2^-24), so an fp16x2 add is an
    // exact integer add there and runs on the FMA pipe beside the DPX ops on the ALU pipe.
    // Negative results come out sign-magnitude, i.e. as large negative s16 values, and are
    // removed by the relu of the following max.
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    smax = __vmaxs2(smax, h);
    u32 hq = swb_hadd2(h, negq);
    e = __viaddmax_s16x2_relu(e, negr, hq);
    f = __viaddmax_s16x2_relu(f, negr, hq);
  }
}

// The same cell in two halves, for loops that want the first operation of a cell issued early (it only
// needs the diagonal value, the score and -- in the int16 build -- E, all known before the cell's turn):
// a = swb_cell_pre(hd, s, e), then swb_cell_post(a, e, f, h, smax).
template <int MODE> __device__ __forceinline__ u32 swb_cell_pre(u32 hd, u32 s, u32 e)
{
  return MODE == SWB_MODE_INT16 ? __viaddmax_s16x2(hd, s, e) : swb_hadd2(hd, s);
}
template <int MODE>
__device__ __forceinline__ void swb_cell_post(u32 a, u32 &e, u32 &f, u32 &h, u32 &smax, const u32 negq, const u32 negr)
{
  if (MODE == SWB_MODE_INT16)
  {
    h = __vimax_s16x2_relu(a, f);
    smax = __vmaxs2(smax, h);
    const u32 hq = __vadd2(h, negq);
    e = __viaddmax_s16x2(e, negr, hq);
    f = __viaddmax_s16x2(f, negr, hq);
  }
  else
  {
    h = __vimax3_s16x2_relu(a, e, f);
    smax = __vmaxs2(smax, h);
    const u32 hq = swb_hadd2(h, negq);
    e = __viaddmax_s16x2_relu(e, negr, hq);
    f = __viaddmax_s16x2_relu(f, negr, hq);
  }
}

// Shared-memory geometry of the scan kernel (host and device agree through these).
#define SWB_STREAMS 8                        // streams per CTA = threads per quarter-warp
#define SWB_MS_STRIDE 34                     // halfwords per subject code in the staged score matrix
#define SWB_XFER_BYTES 48                    // mailbox entry per stream: H[4], F[4], running maximum (+pad)
__host__ __device__ inline int swb_scan_threads(int G) { return SWB_STREAMS * G; }
__host__ __device__ inline size_t swb_scan_smem(int G, int nq)
{
  const size_t warps = (size_t)(SWB_STREAMS * G) / 32;
  return SWB_SMEM_HEADER + (size_t)(G + 1) * (nq + 2) * 128 +
         (1 + 2 * warps) * SWB_STREAMS * SWB_XFER_BYTES;
}

// Thread geometry: a CTA runs 8 streams through G pipeline stages, thread = (stage g, stream k)
// with k = tid & 7 and g = tid >> 3.  The 8 threads of a quarter-warp are therefore at the SAME
// stage, i.e. on the same query rows, of 8 different streams; with the 8 streams' tables
// interleaved (16 B per stream in every 128-B table row) their LDS.128 requests hit one
// 128-B line: no bank conflicts, whatever the query symbols are.  Stage g hands its bottom row to
// stage g+1 by a shuffle over 8 lanes inside a warp and through a double-buffered shared-memory
// mailbox between warps (stage 0 reads an all-zero mailbox); one __syncthreads per step orders
// tables, mailboxes and ring reuse (the ring has G+1 slots so that the slot being rebuilt was
// last read before the barrier).
//
// Table build: after the barrier of step t the CTA builds the tables of block t+1 (read from
// step t+1 on).  Thread (g, k) fills column g & 3 of rows g>>2, g>>2 + G/4, ... of stream k's
// table: one residue pair to decode, two LDS.U16 + PRMT + STS.32 per word, and the 32 words a
// warp stores at a time fill exactly one 128-B table row.
//
// The DP tile itself runs unconditionally (threads outside their stream's block range chew on
// stale tables; only their side effects are predicated off), which keeps the whole step one
// straight-line region for the instruction scheduler.  MP = the query needs more than one pass
// (compiled out otherwise).  All shared-memory traffic goes through word/quad-word indices of
// one typed array, so the compiler addresses it with 32-bit shared offsets.
// KQ / KR != 0: gap penalties compiled in as immediates (both lanes, the mode's encoding) -- the DPX and
// fp16 ops then read one register less each, which is worth ~4 % of the inner loop
// (profiles/r1_ubench_tile_v2.txt, "immediate penalties"); 0 / 0 = read them from ScanParams.
#ifndef SWB_MIN_CTAS
#define SWB_MIN_CTAS(G) (64 / (G))       // CTAs per SM the register allocation is held to (experiments override)
#endif
template <int G, int R, int MODE, bool MP, u32 KQ = 0, u32 KR = 0>
__global__ void __launch_bounds__(SWB_STREAMS * G, SWB_MIN_CTAS(G)) swb_scan_kernel(const ScanParams P)
{
  extern __shared__ uint4 smem4[];
  const u32 sbase = (u32)__cvta_generic_to_shared(smem4);     // staged score matrix at sbase
  constexpr int NSLOT = G + 1;
  constexpr int NWARP = SWB_STREAMS * G / 32;
  constexpr int RG = G / 4;                           // row groups of the table build
  constexpr int NBJ = (32 + RG - 1) / RG;             // rows one thread may have to build
  const int tid = threadIdx.x;
  const int k = tid & 7;
  const int g = tid >> 3;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nq = P.nq;
  const u32 slot_bytes = (u32)(nq + 2) * 128u;
  const u32 ring = sbase + SWB_SMEM_HEADER;                   // shared-window byte addresses
  const u32 ring_bytes = (u32)NSLOT * slot_bytes;
  const u32 zbox = ring + ring_bytes;                         // 8 all-zero mailbox entries
  const u32 xfer = zbox + SWB_STREAMS * SWB_XFER_BYTES;       // [2][NWARP][8] entries of 48 B

  // The score profile is staged once per CTA by the bulk-copy engine (TMA, cp.async.bulk): one thread
  // arms an mbarrier with the byte count and issues the copy; everybody waits on the barrier's phase
  // after the rest of the set-up below.
  __shared__ __align__(8) unsigned long long tma_bar;
  const u32 bar = (u32)__cvta_generic_to_shared(&tma_bar);
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)SWB_M16_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sbase), "l"(__cvta_generic_to_global(P.m16)), "r"((u32)SWB_M16_BYTES), "r"(bar) : "memory");
  }
  // header rows start out flag-free; the pad row of every slot is written once and never rebuilt
  for (int i = tid; i < NSLOT * SWB_STREAMS; i += blockDim.x)
  {
    const u32 base = ring + (u32)(i >> 3) * slot_bytes + (u32)(i & 7) * 16u;
    swb_sts128(base + (u32)nq * 128u, make_uint4(0, 0, 0, 0));
    swb_sts128(base + (u32)(nq + 1) * 128u, make_uint4(P.padword, P.padword, P.padword, P.padword));
  }
  for (int i = tid; i < SWB_STREAMS * SWB_XFER_BYTES / 4; i += blockDim.x) swb_sts32(zbox + 4u * i, 0);
  __syncthreads();                                            // also publishes the mbarrier's initialisation
  {
    u32 done = 0;
    for (int spins = 0; !done; spins++)
    {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
      if (spins > (1 << 24)) __trap();                        // a copy that never lands must not hang the GPU
    }
  }

  ScanSeg S = P.seg;
  if (P.segs) S = P.segs[blockIdx.y];
  const int stream = blockIdx.x * SWB_STREAMS + k;
  const int p0 = S.stream_pair[stream];
  const int p1 = S.stream_pair[stream + 1];
  const long long b0 = S.pairblk[p0];
  const int nblk = (int)(S.pairblk[p1] - b0);
  const uint2 *blk = S.blocks + b0;
  __shared__ int bnd_slot;
  if (MP && P.nslots > 0)
  {
    if (tid == 0) bnd_slot = swb_claim_slot(P);
    __syncthreads();
  }
  const long long bnd0 = !MP ? 0 : (P.nslots > 0 ? (long long)bnd_slot * P.bnd_cta + (long long)k * P.bnd_stream
                                                 : S.bnd_base + b0);
  const int nblk_max = __reduce_max_sync(0xffffffffu, nblk);   // every warp holds all 8 streams
  const int nsteps = nblk_max > 0 ? nblk_max + G - 1 : 0;
  const u32 negq = (KQ | KR) ? KQ : P.negq, negr = (KQ | KR) ? KR : P.negr;
  const bool first_quarter = (lane >> 3) == 0;
  const bool last_quarter = (lane >> 3) == 3 && g < G - 1;
  const int brow = g >> 2;                             // first table row this thread builds
  const u32 bshift = 8u * (u32)(g & 3);                // ... in table column g & 3
  const u32 bdst = ring + (u32)k * 16u + (u32)(g & 3) * 4u + (u32)brow * 128u;
  const u32 bsrc = sbase + 2u * (u32)brow;
  const u32 hdr = ring + (u32)nq * 128u + (u32)k * 16u;       // this stream's flag word in slot 0
  // mailbox addresses; the double buffer is walked by XOR-ing with a ^ b of its two halves
  const u32 xhalf = (u32)NWARP * SWB_STREAMS * SWB_XFER_BYTES;
  const u32 xin0 = warp == 0 ? zbox + (u32)k * SWB_XFER_BYTES
                             : xfer + (u32)((warp - 1) * SWB_STREAMS + k) * SWB_XFER_BYTES;
  const u32 xout0 = xfer + (u32)(warp * SWB_STREAMS + k) * SWB_XFER_BYTES;
  const u32 xin_toggle = warp == 0 ? 0u : (xin0 ^ (xin0 + xhalf));
  const u32 xout_toggle = xout0 ^ (xout0 + xhalf);

  // builds column g & 3 of stream k's table for block words (x, y) at byte offset slot_off
  u32 bmask = 0;                                     // bit j: this thread builds row brow + j * RG
#pragma unroll
  for (int j = 0; j < NBJ; j++)
    if (brow + j * RG < nq) bmask |= 1u << j;
  // (measured: keeping the build behind its branch beats predicating it through one R2P of the mask --
  // 88.6 vs 96.8 ms for the 5 M-subject scan -- the branch-free form costs four more registers)
  auto build = [&](const uint2 blkw, const u32 slot_off) {
    const u32 da = bsrc + ((blkw.x >> bshift) & 63u) * (2u * SWB_MS_STRIDE);
    const u32 db = bsrc + ((blkw.y >> bshift) & 63u) * (2u * SWB_MS_STRIDE);
    const u32 dst = bdst + slot_off;
    // (SWB_BUILD_BATCH > 1, experiment: the loads of several rows issued back to back so that their
    // latencies overlap -- a fifth of the kernel's stall samples sit in these few instructions -- at the
    // price of registers the tile can hardly spare)
#if SWB_BUILD_BATCH == 1
#pragma unroll
    for (int j = 0; j < NBJ; j++)
      if (bmask & (1u << j))
        swb_sts32(dst + j * RG * 128, swb_pack16(swb_lds16(da + j * RG * 2), swb_lds16(db + j * RG * 2)));
#else
#pragma unroll
    for (int j = 0; j < NBJ; j += SWB_BUILD_BATCH)
    {
      u32 lo[SWB_BUILD_BATCH], hi[SWB_BUILD_BATCH];
#pragma unroll
      for (int u = 0; u < SWB_BUILD_BATCH; u++)
        if (j + u < NBJ && (bmask & (1u << (j + u))))
        {
          lo[u] = swb_lds16(da + (j + u) * RG * 2);
          hi[u] = swb_lds16(db + (j + u) * RG * 2);
        }
#pragma unroll
      for (int u = 0; u < SWB_BUILD_BATCH; u++)
        if (j + u < NBJ && (bmask & (1u << (j + u)))) swb_sts32(dst + (j + u) * RG * 128, swb_pack16(lo[u], hi[u]));
    }
#endif
    if (g == 0) swb_sts32(hdr + slot_off, (blkw.x >> 6) & 3u);
  };

  const int npass = MP ? P.npass : 1;
  for (int pass = 0; pass < npass; pass++)
  {
    u32 rq[R];
#pragma unroll
    for (int i = 0; i < R; i++)
      rq[i] = ring + (u32)P.qrow_off[(pass * G + g) * R + i] * 8u + (u32)k * 16u;

    u32 H[R], E[R];
#pragma unroll
    for (int i = 0; i < R; i++) { H[i] = 0; E[i] = 0; }
    u32 smax = 0, dtop = 0;
    u32 ih0 = 0, ih1 = 0, ih2 = 0, ih3 = 0, if0 = 0, if1 = 0, if2 = 0, if3 = 0, is = 0;
    int pair_out = p0;
    const bool feed = MP && (pass > 0) && (g == 0);    // stage 0 reads the previous pass's bottom row
    const bool spill = MP && (pass + 1 < npass) && (g == G - 1);

    if (MP && pass > 0) __syncthreads();               // the previous pass is done with ring and mailboxes
    uint4 pfh = make_uint4(0, 0, 0, 0), pff = make_uint4(0, 0, 0, 0);
    if (MP && feed && nblk > 0) { pfh = P.bndH[bnd0]; pff = P.bndF[bnd0]; }
    uint2 nxt = make_uint2(0, 0);                      // block t + 1
    if (nblk > 0) build(swb_ldg_blk(blk), 0);
    if (nblk > 1) nxt = swb_ldg_blk(blk + 1);
    const uint2 *pnext = blk + 2;
    u32 woff = slot_bytes;                             // ((t + 1) % NSLOT) * slot_bytes
    u32 roff = (u32)((NSLOT - g) % NSLOT) * slot_bytes;  // ((t - g) mod NSLOT) * slot_bytes
    u32 xin = xin0 ^ xin_toggle;                       // the half written at step t - 1
    u32 xout = xout0;
    int b = -g;

    for (int t = 0; t < nsteps; t++)
    {
      __syncthreads();
      // ---- tables of block t + 1 (first read after the next barrier) ------------------------------
      const uint2 cur = nxt;
      if (t + 2 < nblk) nxt = swb_ldg_blk(pnext);
      pnext++;
      if (t + 1 < nblk) build(cur, woff);

      // ---- stage g works on block b = t - g ----------------------------------------------------------
      const bool active = b >= 0 && b < nblk;
      const u32 flags = swb_lds32(hdr + roff);
      if (first_quarter)
      {
        const uint4 vh = swb_lds128(xin), vf = swb_lds128(xin + 16);
        ih0 = vh.x; ih1 = vh.y; ih2 = vh.z; ih3 = vh.w;
        if0 = vf.x; if1 = vf.y; if2 = vf.z; if3 = vf.w;
        is = swb_lds32(xin + 32);
      }
      if (MP && feed)
      {
        // the previous pass's bottom row of this block was fetched one step ago; the fetch for the next
        // block is issued now, a whole tile ahead of its use, so its latency never stalls the pipeline
        if (active)
        {
          ih0 = pfh.x; ih1 = pfh.y; ih2 = pfh.z; ih3 = pfh.w;
          if0 = pff.x; if1 = pff.y; if2 = pff.z; if3 = pff.w;
        }
        if (b + 1 < nblk) { pfh = P.bndH[bnd0 + b + 1]; pff = P.bndF[bnd0 + b + 1]; }
      }
      if (flags & SWB_FLAG_START)
      {
#pragma unroll
        for (int i = 0; i < R; i++) { H[i] = 0; E[i] = 0; }
        smax = 0;
        dtop = 0;
      }
      smax = __vmaxs2(smax, is);
      u32 hup0 = ih0, hup1 = ih1, hup2 = ih2, hup3 = ih3;
      u32 f0 = if0, f1 = if1, f2 = if2, f3 = if3;
      // H[i] = H(row i, last column of the previous block) is the diagonal input of row i + 1; the first
      // operation of that row's first cell is issued one row early (with its score load), so H[i]'s
      // register is free again when row i's own last column is written back into it -- no register
      // rotation at the loop's back edge.
      uint4 sc = swb_lds128(rq[0] + roff);
      u32 a0 = swb_cell_pre<MODE>(dtop, sc.x, E[0]);
      dtop = ih3;
#pragma unroll
      for (int i = 0; i < R; i++)
      {
        uint4 scn = sc;
        u32 an = 0;
        if (i + 1 < R)
        {
          scn = swb_lds128(rq[i + 1] + roff);
          an = swb_cell_pre<MODE>(H[i], scn.x, E[i + 1]);
        }
        u32 e = E[i], h, a;
        swb_cell_post<MODE>(a0, e, f0, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup0, sc.y, e); hup0 = h;
        swb_cell_post<MODE>(a, e, f1, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup1, sc.z, e); hup1 = h;
        swb_cell_post<MODE>(a, e, f2, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup2, sc.w, e); hup2 = h;
        swb_cell_post<MODE>(a, e, f3, h, smax, negq, negr);
        hup3 = h;
        H[i] = h;
        E[i] = e;
        sc = scn;
        a0 = an;
      }
      if (MP && spill && active)
      {
        P.bndH[bnd0 + b] = make_uint4(hup0, hup1, hup2, hup3);
        P.bndF[bnd0 + b] = make_uint4(f0, f1, f2, f3);
      }
      if (g == G - 1 && active && (flags & SWB_FLAG_END))
      {
        u32 v = smax;
        if (MP && pass > 0) v = __vmaxs2(v, S.pair_scores[pair_out]);
        S.pair_scores[pair_out] = v;
        pair_out++;
      }

      // ---- hand the strip's bottom row to the next stage --------------------------------------------
      if (last_quarter)
      {
        swb_sts128(xout, make_uint4(hup0, hup1, hup2, hup3));
        swb_sts128(xout + 16, make_uint4(f0, f1, f2, f3));
        swb_sts32(xout + 32, smax);
      }
      ih0 = __shfl_up_sync(0xffffffffu, hup0, 8);
      ih1 = __shfl_up_sync(0xffffffffu, hup1, 8);
      ih2 = __shfl_up_sync(0xffffffffu, hup2, 8);
      ih3 = __shfl_up_sync(0xffffffffu, hup3, 8);
      if0 = __shfl_up_sync(0xffffffffu, f0, 8);
      if1 = __shfl_up_sync(0xffffffffu, f1, 8);
      if2 = __shfl_up_sync(0xffffffffu, f2, 8);
      if3 = __shfl_up_sync(0xffffffffu, f3, 8);
      is = __shfl_up_sync(0xffffffffu, smax, 8);
      woff = woff + slot_bytes == ring_bytes ? 0u : woff + slot_bytes;
      roff = roff + slot_bytes == ring_bytes ? 0u : roff + slot_bytes;
      xin ^= xin_toggle;
      xout ^= xout_toggle;
      b++;
    }
  }
  if (MP && P.nslots > 0)
  {
    __syncthreads();
    if (tid == 0) swb_release_slot(P, bnd_slot);
  }
}

// ---- scan kernel, second geometry: one warp = one pipeline stage of 32 streams -----------------------
// Same systolic scan, same tables, same block stream; what changes is who sits where.  In
// swb_scan_kernel a warp holds four consecutive stages of eight streams, so its four quarter-warps
// read four different ring slots and every LDS.128 of the inner loop needs a per-thread address add.
// Here a CTA is G warps x 32 streams and warp g IS stage g: ring slot, mailbox and query-row offsets
// are warp-uniform, the ring offset rides in a uniform register of the LDS ([R + UR]) and costs no
// instruction, every hand-off goes through a shared-memory mailbox (three LDS + three STS per step
// instead of nine shuffles), and the START / END flags of a block travel down the pipeline in the two
// spare sign bits of the running-maximum word instead of a table row of their own.  One CTA of 32 G
// threads per SM (G = 16: 512 threads at 128 registers, the whole register file).
// Shared memory: header | ring [G+1][nq+1][32 streams][16 B] | mailboxes [G][2][H 512 B | F 512 B | S 128 B];
// mailbox g is the INPUT of stage g (mailbox 0: zeros, or the previous pass's bottom row, which
// cp.async brings in one step ahead so that the load latency never stalls the pipeline).
#define SWB2_STREAMS 32
#define SWB2_XFER 1152                       // one mailbox, one parity
#define SWB2_FLAG_START 0x00008000u          // in the running-maximum word of a mailbox
#define SWB2_FLAG_END 0x80000000u
__host__ __device__ inline int swb_scan2_threads(int G) { return SWB2_STREAMS * G; }
__host__ __device__ inline size_t swb_scan2_smem(int G, int nq)
{
  return SWB_SMEM_HEADER + (size_t)(G + 1) * (nq + 1) * 512 + (size_t)G * 2 * SWB2_XFER;
}

__device__ __forceinline__ void swb_cp_async16(u32 dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}

template <int G, int R, int MODE, bool MP, u32 KQ = 0, u32 KR = 0>
__global__ void __launch_bounds__(SWB2_STREAMS * G, 16 / G) swb_scan2_kernel(const ScanParams P)
{
  extern __shared__ uint4 smem4[];
  const u32 sbase = (u32)__cvta_generic_to_shared(smem4);     // staged score matrix at sbase
  constexpr int NSLOT = G + 1;
  constexpr int RG = G >= 4 ? G / 4 : 1;              // row groups of the table build
  constexpr int NBJ = (32 + RG - 1) / RG;             // rows one thread may have to build
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = __shfl_sync(0xffffffffu, tid >> 5, 0);        // stage = warp, known to be warp-uniform
  const int nq = P.nq;
  const u32 slot_bytes = (u32)(nq + 1) * 512u;
  const u32 ring = sbase + SWB_SMEM_HEADER;
  const u32 ring_bytes = (u32)NSLOT * slot_bytes;
  const u32 xfer = ring + ring_bytes;

  __shared__ __align__(8) unsigned long long tma_bar;
  const u32 bar = (u32)__cvta_generic_to_shared(&tma_bar);
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)SWB_M16_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sbase), "l"(__cvta_generic_to_global(P.m16)), "r"((u32)SWB_M16_BYTES), "r"(bar) : "memory");
  }
  // the pad row of every slot is written once and never rebuilt; mailbox 0 starts out all zero
  for (int i = tid; i < NSLOT * SWB2_STREAMS; i += blockDim.x)
    swb_sts128(ring + (u32)(i >> 5) * slot_bytes + (u32)nq * 512u + (u32)(i & 31) * 16u,
               make_uint4(P.padword, P.padword, P.padword, P.padword));
  for (int i = tid; i < 2 * SWB2_XFER / 4; i += blockDim.x) swb_sts32(xfer + 4u * i, 0);
  __syncthreads();
  {
    u32 done = 0;
    for (int spins = 0; !done; spins++)
    {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
      if (spins > (1 << 24)) __trap();
    }
  }

  ScanSeg S = P.seg;
  if (P.segs) S = P.segs[blockIdx.y];
  const int stream = blockIdx.x * SWB2_STREAMS + lane;
  const int p0 = S.stream_pair[stream];
  const int p1 = S.stream_pair[stream + 1];
  const long long b0 = S.pairblk[p0];
  const int nblk = (int)(S.pairblk[p1] - b0);
  const uint2 *blk = S.blocks + b0;
  __shared__ int bnd_slot;
  if (MP && P.nslots > 0)
  {
    if (tid == 0) bnd_slot = swb_claim_slot(P);
    __syncthreads();
  }
  const long long bnd0 = !MP ? 0 : (P.nslots > 0 ? (long long)bnd_slot * P.bnd_cta + (long long)lane * P.bnd_stream
                                                 : S.bnd_base + b0);
  const int nblk_max = __reduce_max_sync(0xffffffffu, nblk);   // every warp holds all 32 streams
  const int nsteps = nblk_max > 0 ? nblk_max + G - 1 : 0;
  const u32 negq = (KQ | KR) ? KQ : P.negq, negr = (KQ | KR) ? KR : P.negr;
  const bool qstart = (P.start_mask >> g) & 1u, qend = (P.end_mask >> g) & 1u;
  u32 *qscores = S.pair_scores + (long long)P.stage_query[g] * S.score_stride;
  // table build: this thread fills ONE column of its stream's table, rows brow, brow + RG, ...; the
  // column rotates with the lane's octet so that the 32 STS.32 of a warp fall into 32 different banks
  const int bcol = ((lane >> 3) + g) & 3;
  const int brow = G >= 4 ? g >> 2 : 0;
  const u32 bshift = 8u * (u32)bcol;
  const u32 bdst = ring + (u32)lane * 16u + (u32)bcol * 4u + (u32)brow * 512u;
  const u32 bsrc = sbase + 2u * (u32)brow;
  const u32 lane16 = (u32)lane * 16u;
  const u32 xin0 = xfer + (u32)g * (2u * SWB2_XFER) + lane16;          // input mailbox, parity 0
  const u32 xout0 = xin0 + 2u * SWB2_XFER;                             // = input mailbox of stage g + 1
  const u32 xs0 = xin0 - lane16 + 1024u + (u32)lane * 4u;              // running maximum + flags word of the input

  u32 bmask = 0;                                     // bit j: this thread builds row brow + j * RG
#pragma unroll
  for (int j = 0; j < NBJ; j++)
    if (brow + j * RG < nq) bmask |= 1u << j;
  auto build = [&](const uint2 blkw, const u32 slot_off, const bool on) {
    const u32 da = bsrc + ((blkw.x >> bshift) & 63u) * (2u * SWB_MS_STRIDE);
    const u32 db = bsrc + ((blkw.y >> bshift) & 63u) * (2u * SWB_MS_STRIDE);
    const u32 dst = bdst + slot_off;
    const u32 bm = on ? bmask : 0u;
#pragma unroll
    for (int j = 0; j < NBJ; j++)
      if (bm & (1u << j))
        swb_sts32(dst + j * RG * 512, swb_pack16(swb_lds16(da + j * RG * 2), swb_lds16(db + j * RG * 2)));
  };

  const int npass = MP ? P.npass : 1;
  for (int pass = 0; pass < npass; pass++)
  {
    u32 rq[R];
#pragma unroll
    for (int i = 0; i < R; i++)
      rq[i] = ring + (u32)P.qrow_off[(pass * G + g) * R + i] * 32u + lane16;

    u32 H[R], E[R];
#pragma unroll
    for (int i = 0; i < R; i++) { H[i] = 0; E[i] = 0; }
    u32 smax = 0, dtop = 0;
    int pair_out = p0;
    const bool feed = MP && (pass > 0) && (g == 0);    // stage 0 reads the previous pass's bottom row
    const bool spill = MP && (pass + 1 < npass) && (g == G - 1);

    if (MP && pass > 0) __syncthreads();               // the previous pass is done with ring and mailboxes
    uint2 now = make_uint2(0, 0), nxt = make_uint2(0, 0);     // blocks t and t + 1
    if (nblk > 0) { now = swb_ldg_blk(blk); build(now, 0, true); }
    if (nblk > 1) nxt = swb_ldg_blk(blk + 1);
    if (MP && feed)
    {
      if (nblk > 0)
      {
        swb_cp_async16(xin0, P.bndH + bnd0);            // block 0's top row -> parity 0
        swb_cp_async16(xin0 + 512u, P.bndF + bnd0);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint2 *pnext = blk + 2;
    u32 woff = slot_bytes;                             // ((t + 1) % NSLOT) * slot_bytes
    u32 roff = (u32)((NSLOT - g) % NSLOT) * slot_bytes;  // ((t - g) mod NSLOT) * slot_bytes
    u32 par = 0;                                       // (t & 1) * SWB2_XFER
    int b = -g;

    for (int t = 0; t < nsteps; t++)
    {
      if (MP && feed) asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      // ---- tables of block t + 1 (first read after the next barrier) ------------------------------
      const uint2 cur = nxt;
      if (t + 2 < nblk) nxt = swb_ldg_blk(pnext);
      pnext++;
      // Even stages build before their tile, odd stages after it: the barrier puts all G warps of the SM
      // at the same point of the step, and this keeps half of them in the (ALU-free) build phase while
      // the other half is in the DP tile.
      const bool build_late = P.stagger && (g & 1);
      if (!build_late) build(cur, woff, t + 1 < nblk);
      if (MP && feed)
      {
        if (t + 1 < nblk)                              // top row of block t + 1 -> the other parity
        {
          swb_cp_async16(xin0 + (par ^ SWB2_XFER), P.bndH + bnd0 + t + 1);
          swb_cp_async16(xin0 + (par ^ SWB2_XFER) + 512u, P.bndF + bnd0 + t + 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }

      // ---- stage g works on block b = t - g ----------------------------------------------------------
      const bool active = b >= 0 && b < nblk;
      uint4 vh = make_uint4(0, 0, 0, 0), vf = make_uint4(0, 0, 0, 0);
      u32 is;
      if (g == 0 || !qstart) { vh = swb_lds128(xin0 + par); vf = swb_lds128(xin0 + par + 512u); }
      if (g == 0) is = ((now.x >> 6) & 1u) * SWB2_FLAG_START + ((now.x >> 7) & 1u) * SWB2_FLAG_END;
      else is = swb_lds32(xs0 + par);
      if (qstart) is &= SWB2_FLAG_START | SWB2_FLAG_END;     // a query's first stage inherits no maximum
      if (is & SWB2_FLAG_START)
      {
#pragma unroll
        for (int i = 0; i < R; i++) { H[i] = 0; E[i] = 0; }
        smax = 0;
        dtop = 0;
      }
      const u32 flagbits = is & (SWB2_FLAG_START | SWB2_FLAG_END);
      smax = __vmaxs2(smax, is & ~(SWB2_FLAG_START | SWB2_FLAG_END));
      u32 hup0 = vh.x, hup1 = vh.y, hup2 = vh.z, hup3 = vh.w;
      u32 f0 = vf.x, f1 = vf.y, f2 = vf.z, f3 = vf.w;
      // H[i] holds H(row i, last column of the previous block) = the diagonal input of row i + 1.  It is
      // consumed by the first operation of row i + 1's first cell, issued one row EARLY (software
      // pipelining of the score load and that operation), so that the register is free again when row
      // i's own last column is written back into it.
      uint4 sc = swb_lds128(rq[0] + roff);
      u32 a0 = swb_cell_pre<MODE>(dtop, sc.x, E[0]);
      dtop = vh.w;
#pragma unroll
      for (int i = 0; i < R; i++)
      {
        uint4 scn = sc;
        u32 an = 0;
        if (i + 1 < R)
        {
          scn = swb_lds128(rq[i + 1] + roff);
          an = swb_cell_pre<MODE>(H[i], scn.x, E[i + 1]);
        }
        u32 e = E[i], h, a;
        swb_cell_post<MODE>(a0, e, f0, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup0, sc.y, e); hup0 = h;
        swb_cell_post<MODE>(a, e, f1, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup1, sc.z, e); hup1 = h;
        swb_cell_post<MODE>(a, e, f2, h, smax, negq, negr);
        a = swb_cell_pre<MODE>(hup2, sc.w, e); hup2 = h;
        swb_cell_post<MODE>(a, e, f3, h, smax, negq, negr);
        hup3 = h;
        H[i] = h;
        E[i] = e;
        sc = scn;
        a0 = an;
      }
      if (qend)
      {
        if (MP && spill && active)
        {
          P.bndH[bnd0 + b] = make_uint4(hup0, hup1, hup2, hup3);
          P.bndF[bnd0 + b] = make_uint4(f0, f1, f2, f3);
        }
        if (active && (flagbits & SWB2_FLAG_END))
        {
          u32 v = smax;
          if (MP && pass > 0) v = __vmaxs2(v, qscores[pair_out]);
          qscores[pair_out] = v;
          pair_out++;
        }
      }
      if (g < G - 1)
      {
        // ---- hand the strip's bottom row to the next stage ------------------------------------------
        // (written to the parity the next stage reads at step t + 1; it is reading the other one now)
        const u32 wpar = par ^ SWB2_XFER;
        swb_sts128(xout0 + wpar, make_uint4(hup0, hup1, hup2, hup3));
        swb_sts128(xout0 + wpar + 512u, make_uint4(f0, f1, f2, f3));
        swb_sts32(xs0 + 2u * SWB2_XFER + wpar, smax | flagbits);
      }
      if (build_late) build(cur, woff, t + 1 < nblk);
      now = cur;
      woff = woff + slot_bytes == ring_bytes ? 0u : woff + slot_bytes;
      roff = roff + slot_bytes == ring_bytes ? 0u : roff + slot_bytes;
      par ^= SWB2_XFER;
      b++;
    }
  }
  if (MP)
  {
    asm volatile("cp.async.wait_group 0;" ::: "memory");       // (stage 0's last prefetch, if any)
    __syncthreads();
    if (tid == 0 && P.nslots > 0) swb_release_slot(P, bnd_slot);
  }
}

// ---- wide kernel: one thread per subject, 32- or 64-bit cells ------------------------------
// Used for subjects whose 16-bit lane reached the overflow limit, for scoring systems the
// packed kernel cannot represent, and (with END = true) for search16s's alignment-end contract
// (search16s.cc:390-405): first column reaching the final maximum, smallest row within it.
// H/E of the query rows live in a global scratch laid out [row][thread] so a warp's accesses
// coalesce.
struct WideParams
{
  const unsigned char *residues;   // raw database symbols as uploaded
  const long long *offsets;        // [nseq+1]
  int trailing;
  const long long *list;           // caller's list of subject numbers, or NULL (position = subject)
  const long long *sel;            // positions of that list to score, or NULL (all of them)
  long long nsel;
  const unsigned char *query;
  int qlen;
  const long long *matrix;         // [32][32] (subject << 5) + query
  long long q, r;
  void *he;                        // [2*qlen][stride] of T
  long long stride;
  long long *scores;               // indexed by list position
  long long *bestpos, *bestq;      // END only, indexed like scores
};

template <typename T, bool END>
__global__ void __launch_bounds__(128) swb_wide_kernel(const WideParams P)
{
  __shared__ T Msh[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) Msh[i] = (T)P.matrix[i];
  __syncthreads();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= P.nsel) return;
  const long long item = P.sel ? P.sel[tid] : tid;
  const long long subj = P.list ? (P.list[item] >> 3) : item;
  // strand bit of the coded subject number (swipe.cc:1373-1391): score against the reverse
  // complement of a nucleotide subject, as db_getsequence serves it (database.cc:1327-1353)
  const bool rc = P.list ? ((P.list[item] >> 2) & 1) != 0 : false;
  const long long o0 = P.offsets[subj];
  const long long dlen = P.offsets[subj + 1] - o0 - P.trailing;
  const unsigned char *d = P.residues + o0;
  T *he = (T *)P.he + tid;
  const long long st = P.stride;
  const int qlen = P.qlen;
  const T q = (T)P.q, r = (T)P.r;
  for (int i = 0; i < 2 * qlen; i++) he[i * st] = 0;
  T best = 0;
  long long bq = -1, bd = -1;
  for (long long j = 0; j < dlen; j++)
  {
    unsigned sym = d[rc ? dlen - 1 - j : j] & 31;
    if (rc) sym = ((sym & 1) << 3) | ((sym & 2) << 1) | ((sym & 4) >> 1) | ((sym & 8) >> 3);
    const T *row = Msh + (sym << 5);
    T f = 0, h = 0;
    for (int i = 0; i < qlen; i++)
    {
      const T hn = he[(2 * i) * st];
      T e = he[(2 * i + 1) * st];
      h += row[P.query[i]];
      h = h > e ? h : e;
      h = h > f ? h : f;
      h = h > 0 ? h : 0;
      if (h > best) { best = h; if (END) { bq = i; bd = j; } }
      he[(2 * i) * st] = h;
      const T hq = h - q;
      e -= r; f -= r;
      e = e > hq ? e : hq;
      f = f > hq ? f : hq;
      he[(2 * i + 1) * st] = e;
      h = hn;
    }
  }
  P.scores[item] = (long long)best;
  if (END) { P.bestpos[item] = bd; P.bestq[item] = bq; }
}

// ---- end-cell kernel: one WARP per subject -----------------------------------------------------------
// search16s's contract (search16s.cc:390-405, called for the hits that get an alignment, swipe.cc:381-393):
// exact score, the first subject column in which it is reached and the smallest query row reaching it in
// that column.  The 32 lanes of a warp own consecutive strips of S query rows (H and E of the strip in
// registers) and sweep the subject as a wavefront: at step t lane l works on column t - l and hands the
// bottom H / F of its strip to lane l + 1 by shuffle.  Every lane remembers its own first best cell; the
// lanes' candidates are then reduced by (value, then column, then row).  32-bit cells; 64-bit scoring
// systems stay on swb_wide_kernel.
template <int S>
__global__ void __launch_bounds__(128) swb_end_kernel(const WideParams P)
{
  __shared__ int Msh[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) Msh[i] = (int)P.matrix[i];
  __syncthreads();
  const long long item0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item0 >= P.nsel) return;                       // whole warps leave together
  const int lane = threadIdx.x & 31;
  const long long item = P.sel ? P.sel[item0] : item0;
  const long long subj = P.list ? (P.list[item] >> 3) : item;
  const bool rc = P.list ? ((P.list[item] >> 2) & 1) != 0 : false;
  const long long o0 = P.offsets[subj];
  const int dlen = (int)(P.offsets[subj + 1] - o0 - P.trailing);
  const unsigned char *d = P.residues + o0;
  const int qlen = P.qlen;
  const int q = (int)P.q, r = (int)P.r;
  // queries of more than 32 S rows: passes of 32 S rows; the bottom row of a pass (H and F per subject
  // column) waits in global scratch for lane 0 of the next pass.  Lane 31 writes column j 31 steps after
  // lane 0 has read it, so one pair of arrays serves every pass.
  const int npass = (qlen + 32 * S - 1) / (32 * S);
  int *bndH = (int *)P.he + item0 * 2 * P.stride, *bndF = bndH + P.stride;
  int best = 0, bcol = -1, brow = -1;
  for (int pass = 0; pass < npass; pass++)
  {
    const int row0 = (pass * 32 + lane) * S;
    int qsym[S], H[S], E[S];
#pragma unroll
    for (int i = 0; i < S; i++)
    {
      qsym[i] = row0 + i < qlen ? (int)P.query[row0 + i] : 0;
      H[i] = 0;
      E[i] = 0;
    }
    int hup_prev = 0;                                // H of the row above the strip, previous column
    int hbot = 0, fbot = 0;                          // bottom of this strip, the column just done
    __syncwarp();
    for (int t = 0; t < dlen + 31; t++)
    {
      // the strip above finished column t - lane one step ago
      int hup = __shfl_up_sync(0xffffffffu, hbot, 1);
      int fup = __shfl_up_sync(0xffffffffu, fbot, 1);
      const int j = t - lane;
      if (lane == 0)
      {
        hup = 0; fup = 0;
        if (pass > 0 && j < dlen) { hup = bndH[j]; fup = bndF[j]; }
      }
      if (j >= 0 && j < dlen)
      {
        unsigned sym = d[rc ? dlen - 1 - j : j] & 31;
        if (rc) sym = ((sym & 1) << 3) | ((sym & 2) << 1) | ((sym & 4) >> 1) | ((sym & 8) >> 3);
        const int *mrow = Msh + (sym << 5);
        int diag = hup_prev, f = fup;
#pragma unroll
        for (int i = 0; i < S; i++)
        {
          int h = diag + mrow[qsym[i]];
          h = max(max(h, E[i]), max(f, 0));
          // first column reaching the maximum, smallest row in it: later passes hold larger rows, so on
          // equal values only a smaller column takes over
          if ((h > best || (h == best && h > 0 && j < bcol)) && row0 + i < qlen) { best = h; bcol = j; brow = row0 + i; }
          const int hq = h - q;
          E[i] = max(E[i] - r, hq);
          f = max(f - r, hq);
          diag = H[i];
          H[i] = h;
        }
        hbot = H[S - 1];
        fbot = f;
        hup_prev = hup;
        if (lane == 31 && pass + 1 < npass) { bndH[j] = hbot; bndF[j] = fbot; }
      }
    }
  }
  // the warp's best cell: highest value, then first column, then smallest row
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
  {
    const int ob = __shfl_down_sync(0xffffffffu, best, off);
    const int oc = __shfl_down_sync(0xffffffffu, bcol, off);
    const int orow = __shfl_down_sync(0xffffffffu, brow, off);
    const bool better = ob > best || (ob == best && ob > 0 && (oc < bcol || (oc == bcol && orow < brow)));
    if (better) { best = ob; bcol = oc; brow = orow; }
  }
  if (lane == 0)
  {
    P.scores[item] = best;
    P.bestpos[item] = bcol;
    P.bestq[item] = brow;
  }
}

// ---- nucleotide ingest ---------------------------------------------------------------------------
// One warp per subject: unpack the .nsq record (4 bases per byte, most significant pair first; the
// last byte carries the remaining len % 4 bases) to the 4-bit one-hot codes A=1 C=2 G=4 T=8 and
// patch in the ambiguity runs that follow the packed bases -- what db_getsequence does per pull
// on the host (database.cc:1257-1323), done once per shard here.  Table entries are big-endian:
// 32-bit {code:4, run-1:4, offset:24}, or 64-bit {code:4, run-1:12, offset:48} when bit 31 of
// the leading count word is set.
__device__ __forceinline__ u32 swb_be32(const unsigned char *p)
{
  return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3];
}

__global__ void swb_nt_decode_kernel(const unsigned char *packed, const long long *pk_start,
                                     const u32 *pk_len, const long long *offsets,
                                     unsigned char *residues, long long first, long long n)
{
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long s = first + w;
  const unsigned char *src = packed + pk_start[s];
  unsigned char *dst = residues + offsets[s];
  const long long len = offsets[s + 1] - offsets[s];
  const long long nbytes = (len + 3) >> 2;
  const bool aligned = (((unsigned long long)dst) & 3ull) == 0;
  for (long long j = lane; j < nbytes; j += 32)
  {
    const u32 b = src[j];
    const u32 word = (1u << ((b >> 6) & 3)) | ((1u << ((b >> 4) & 3)) << 8) |
                     ((1u << ((b >> 2) & 3)) << 16) | ((1u << (b & 3)) << 24);
    if (aligned && 4 * j + 4 <= len)
      *(u32 *)(dst + 4 * j) = word;
    else
      for (int c = 0; c < 4; c++)
        if (4 * j + c < len) dst[4 * j + c] = (unsigned char)(word >> (8 * c));
  }
  const long long abytes = pk_start[s + 1] - pk_start[s] - (long long)pk_len[s];
  if (abytes < 4) return;
  __syncwarp();
  const unsigned char *amb = src + pk_len[s];
  const u32 head = swb_be32(amb);
  if (head >> 31)
  {
    const long long entries = (abytes - 4) >> 3;
    for (long long k = lane; k < entries; k += 32)
    {
      const u64 e = ((u64)swb_be32(amb + 4 + 8 * k) << 32) | swb_be32(amb + 8 + 8 * k);
      const unsigned char code = (unsigned char)(e >> 60);
      const long long run = (long long)((e >> 48) & 0xfff) + 1, off = (long long)(e & 0x0000ffffffffffffULL);
      for (long long r = 0; r < run && off + r < len; r++) dst[off + r] = code;
    }
  }
  else
  {
    const long long entries = (abytes - 4) >> 2;
    for (long long k = lane; k < entries; k += 32)
    {
      const u32 e = swb_be32(amb + 4 + 4 * k);
      const unsigned char code = (unsigned char)(e >> 28);
      const long long run = (long long)((e >> 24) & 0xf) + 1, off = (long long)(e & 0x00ffffffu);
      for (long long r = 0; r < run && off + r < len; r++) dst[off + r] = code;
    }
  }
}

// Six-frame translation of decoded nucleotide subjects (db_translate, database.cc:1182-1218, with
// the table of query.cc:366-436): one warp per nucleotide sequence s writes the protein subjects
// 6 s + 3 strand + frame.  Strand 1 reads the reverse complement from the far end.
__global__ void swb_translate_kernel(const unsigned char *nt, const long long *nt_offsets,
                                     const unsigned char *table, const long long *offsets,
                                     unsigned char *residues, long long first, long long n)
{
  __shared__ unsigned char tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = table[i];
  __syncthreads();
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long s = first + w;
  const unsigned char *src = nt + nt_offsets[s];
  const long long len = nt_offsets[s + 1] - nt_offsets[s];
  for (int k = 0; k < 6; k++)
  {
    const int strand = k / 3, frame = k % 3;
    const long long plen = len - frame >= 0 ? (len - frame) / 3 : 0;
    unsigned char *dst = residues + offsets[6 * s + k];
    for (long long p = lane; p < plen; p += 32)
    {
      u32 a, b, c;
      if (!strand)
      {
        const long long pos = frame + 3 * p;
        a = src[pos] & 15; b = src[pos + 1] & 15; c = src[pos + 2] & 15;
      }
      else
      {
        const long long pos = len - 1 - frame - 3 * p;
        a = src[pos] & 15; b = src[pos - 1] & 15; c = src[pos - 2] & 15;
        a = ((a & 1) << 3) | ((a & 2) << 1) | ((a & 4) >> 1) | ((a & 8) >> 3);
        b = ((b & 1) << 3) | ((b & 2) << 1) | ((b & 4) >> 1) | ((b & 8) >> 3);
        c = ((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3);
      }
      dst[p] = tab[(a << 8) | (b << 4) | c];
    }
  }
}

// ---- layout kernels ------------------------------------------------------------------------------
// keys for the length sort: subject k of the list -> its length
__global__ void swb_len_kernel(const long long *offsets, int trailing, const long long *list,
                               long long n, u32 *keys, u32 *vals)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long s = list ? (list[k] >> 3) : k;
  long long len = offsets[s + 1] - offsets[s] - trailing;
  if (len < 0) len = 0;
  keys[k] = (u32)len;
  vals[k] = (u32)k;
}

// blocks per pair from the sorted lengths (descending): pair p = sorted entries 2p, 2p+1
__global__ void swb_pairblk_kernel(const u32 *sorted_len, long long n, long long npairs,
                                   long long *nblk)
{
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const u32 la = sorted_len[2 * p];           // descending: the first of the pair is the longer
  nblk[p] = (la + 3) / 4;
}

// one warp per pair: interleave the two subjects into 8-byte blocks
__global__ void swb_fill_kernel(const unsigned char *residues, const long long *offsets,
                                int trailing, const long long *list, const u32 *sorted_len,
                                const u32 *sorted_idx, long long n, long long npairs,
                                const long long *pairblk, uint2 *blocks)
{
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= npairs) return;
  const long long nb = pairblk[p + 1] - pairblk[p];
  if (nb == 0) return;
  const u32 ka = sorted_idx[2 * p];
  const long long sa = list ? (list[ka] >> 3) : (long long)ka;
  const unsigned char *da = residues + offsets[sa];
  const u32 la = sorted_len[2 * p];
  const unsigned char *db = da;
  u32 lb = 0;
  if (2 * p + 1 < n)
  {
    const u32 kb = sorted_idx[2 * p + 1];
    const long long sb = list ? (list[kb] >> 3) : (long long)kb;
    db = residues + offsets[sb];
    lb = sorted_len[2 * p + 1];
  }
  uint2 *out = blocks + pairblk[p];
  for (long long b = lane; b < nb; b += 32)
  {
    u32 x = 0, y = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
      const long long j = b * 4 + c;
      const u32 ra = j < la ? (u32)(da[j] & 31) : (u32)SWB_PAD_CODE;
      const u32 rb = j < lb ? (u32)(db[j] & 31) : (u32)SWB_PAD_CODE;
      x |= ra << (8 * c);
      y |= rb << (8 * c);
    }
    u32 flags = 0;
    if (b == 0) flags |= SWB_FLAG_START;
    if (b == nb - 1) flags |= SWB_FLAG_END;
    x |= flags << 6;
    out[b] = make_uint2(x, y);
  }
}

// stream s covers pairs [stream_pair[s], stream_pair[s+1]): equal shares of the block total
__global__ void swb_partition_kernel(const long long *pairblk, long long npairs, int nstreams,
                                     int *stream_pair)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nstreams) return;
  if (s == nstreams) { stream_pair[s] = (int)npairs; return; }
  const long long total = pairblk[npairs];
  const long long target = (long long)(((__int128)total * s) / nstreams);
  long long lo = 0, hi = npairs;                 // first p with pairblk[p] >= target
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if (pairblk[mid] >= target) hi = mid; else lo = mid + 1;
  }
  stream_pair[s] = (int)lo;
}

// unpack lane maxima into per-subject scores; lanes at or above the limit are queued for the next
// wider pass instead (the reference's re-queue, swipe.cc:1459-1479, :1514-1537).  remap != NULL:
// the layout was built over a re-queue list, entry k of it stands for position remap[k].
__global__ void swb_finish_kernel(const u32 *pair_scores, const u32 *sorted_idx, long long n,
                                  long long first, int limit, const long long *remap, long long *scores,
                                  long long *requeue, unsigned long long *nrequeue)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const u32 w = pair_scores[k >> 1];
  const int v = (int)((k & 1) ? (w >> 16) : (w & 0xffffu));
  long long pos = first + sorted_idx[k];          // position in the caller's list / shard
  if (remap) pos = remap[pos];
  if (v >= limit)
  {
    const unsigned long long slot = atomicAdd(nrequeue, 1ull);
    requeue[slot] = pos;
  }
  else
    scores[pos] = v;
}

// coded subject numbers (seqno << 3) of the positions queued for the next pass
__global__ void swb_requeue_codes_kernel(const long long *requeue, long long n, const long long *list,
                                         long long *codes)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long pos = requeue[k];
  codes[k] = list ? (list[pos] & ~7ll) : (pos << 3);
}

// ---- the sink on the device: hits_enter's admission rule (hits.cc:163-222) over the final scores ---
// The reference never materialises scores outside [scorethreshold, upperscorethreshold] and, once
// its list is full, tightens the threshold to the K-th score (hits.cc:180-184, :218-219).  Here:
// (1) one pass builds a histogram of the admissible scores (plus the width bookkeeping and the
// totalhits / obvious counts), (2) one block finds the K-th score's bin, (3) one pass appends
// (score << 32 | subject) of everything at or above it to a candidate list, which is then sorted
// descending -- score first, subject number second, the reference's order -- and cut to K.
#define SWB_HIST_BINS 4096     // scores >= 4095 share the last bin (the cut is then conservative)

// counts: [0..2] subjects the reference's 7 / 16 / 63-bit pass would have kept, [3] totalhits
// (score >= min_score), [4] obvious (score > upper)
// filter: one bit per subject (bit k & 7 of byte k >> 3), a clear bit keeps the subject out of the sink
// (the reference's db_check_inclusion, swipe.cc:1373-1376); NULL = every subject takes part
__device__ __forceinline__ bool swb_included(const unsigned char *filter, long long k)
{
  return !filter || ((filter[k >> 3] >> (k & 7)) & 1);
}

__global__ void __launch_bounds__(256) swb_hist_kernel(const long long *scores, long long n,
                                                        long long limit7, long long limit16,
                                                        long long min_score, long long upper,
                                                        const unsigned char *filter,
                                                        unsigned long long *counts, unsigned *hist)
{
  __shared__ unsigned sh[SWB_HIST_BINS];
  __shared__ unsigned long long cnt[5];
  for (int i = threadIdx.x; i < SWB_HIST_BINS; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned w7 = 0, w16 = 0, w63 = 0, tot = 0, obv = 0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
  {
    const long long v = scores[k];
    if (v < limit7) w7++; else if (v < limit16) w16++; else w63++;
    if (!swb_included(filter, k)) continue;
    tot += v >= min_score;
    obv += v > upper;
    if (v >= min_score && v <= upper)
      atomicAdd(&sh[v < 0 ? 0 : (v > SWB_HIST_BINS - 1 ? SWB_HIST_BINS - 1 : (int)v)], 1u);
  }
  w7 = __reduce_add_sync(0xffffffffu, w7);
  w16 = __reduce_add_sync(0xffffffffu, w16);
  w63 = __reduce_add_sync(0xffffffffu, w63);
  tot = __reduce_add_sync(0xffffffffu, tot);
  obv = __reduce_add_sync(0xffffffffu, obv);
  if ((threadIdx.x & 31) == 0)
  {
    if (w7) atomicAdd(&cnt[0], (unsigned long long)w7);
    if (w16) atomicAdd(&cnt[1], (unsigned long long)w16);
    if (w63) atomicAdd(&cnt[2], (unsigned long long)w63);
    if (tot) atomicAdd(&cnt[3], (unsigned long long)tot);
    if (obv) atomicAdd(&cnt[4], (unsigned long long)obv);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SWB_HIST_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
  if (threadIdx.x < 5 && cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], cnt[threadIdx.x]);
}

// one block of 1024 threads: cut[0] = the highest bin c with count(bin >= c) >= keep (0 when fewer
// than keep scores are admissible): everything in bins >= c is a candidate
__global__ void __launch_bounds__(1024) swb_cut_kernel(const unsigned *hist, long long keep, unsigned *cut)
{
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  constexpr int PER = SWB_HIST_BINS / 1024;
  unsigned long long mine = 0;                    // thread t owns bins top-down: BINS-1 - PER*t - j
  for (int j = 0; j < PER; j++) mine += hist[SWB_HIST_BINS - 1 - (PER * t + j)];
  part[t] = mine;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1)              // inclusive scan (Hillis-Steele; 10 rounds)
  {
    const unsigned long long add = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  const unsigned long long before = part[t] - mine;
  if (t == 0) cut[0] = 0;
  __syncthreads();
  if (keep > 0 && before < (unsigned long long)keep && part[t] >= (unsigned long long)keep)
  {
    unsigned long long acc = before;
    for (int j = 0; j < PER; j++)
    {
      const int bin = SWB_HIST_BINS - 1 - (PER * t + j);
      acc += hist[bin];
      if (acc >= (unsigned long long)keep) { cut[0] = (unsigned)bin; break; }
    }
  }
}

// cand[slot] = score << 32 | subject for every admissible score whose bin is >= cut[0]
__global__ void __launch_bounds__(256) swb_compact_kernel(const long long *scores, long long n,
                                                           long long min_score, long long upper,
                                                           const unsigned char *filter,
                                                           const unsigned *cut, unsigned long long *cand,
                                                           unsigned long long *ncand)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = (long long)cut[0];
  bool take = false;
  long long v = 0;
  if (k < n)
  {
    v = scores[k];
    const long long bin = v > SWB_HIST_BINS - 1 ? SWB_HIST_BINS - 1 : v;
    take = v >= min_score && v <= upper && bin >= c && swb_included(filter, k);
  }
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(ncand, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (take) cand[base + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)v << 32) | (unsigned long long)k;
}

// reference-width bookkeeping: which of the reference's passes would have kept each score
__global__ void swb_widthcount_kernel(const long long *scores, long long n, long long limit7,
                                      long long limit16, unsigned long long *counts)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int w = -1;
  if (k < n)
  {
    const long long v = scores[k];
    w = v < limit7 ? 0 : (v < limit16 ? 1 : 2);
  }
  for (int c = 0; c < 3; c++)
  {
    const unsigned m = __ballot_sync(0xffffffffu, w == c);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counts[c], (unsigned long long)__popc(m));
  }
}
