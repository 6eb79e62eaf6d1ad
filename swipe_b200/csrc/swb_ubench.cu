// swb_ubench.cu -- swb_alu_peak: measures, on the device it runs on, the issue rate of the packed
// 16x2 DPX instructions (VIADDMNMX / VIMNMX3) that bound the scan kernel.  bench.py turns the
// figure into the roofline denominator of the run it belongs to (SURVEY 8d asks for a
// microbenchmark-calibrated integer peak: there is no driver-measured one).
#include "../../include/swipe_b200.h"
#include <cuda_runtime.h>

namespace
{
typedef unsigned int u32;

// 8 independent dependency chains per thread, 1024 threads per CTA, one CTA per SM: enough
// independent work that the figure is pipe throughput, not latency.
__global__ void __launch_bounds__(1024, 1) swb_dpx_rate_kernel(u32 *out, const u32 *in, int iters,
                                                                long long *cycles)
{
  u32 x[8], y[8];
  const u32 c0 = in[0], c1 = in[1];
#pragma unroll
  for (int k = 0; k < 8; k++) { x[k] = in[2 + k] + threadIdx.x; y[k] = in[10 + k] ^ threadIdx.x; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        x[k] = __viaddmax_s16x2_relu(x[k], c0, y[k]);      // the E / F update of a cell
        y[k] = __vimax3_s16x2_relu(y[k], x[k], c1);        // the H update
      }
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) acc ^= x[k] ^ y[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
}  // namespace

extern "C" int swb_alu_peak(int device, double *dpx_per_clk_per_sm, double *sm_clock_mhz)
{
  if (!dpx_per_clk_per_sm) return SWB_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
  {
    (void)cudaGetLastError();
    return SWB_ERR_NO_DEVICE;
  }
  int sms = 0;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return SWB_ERR_CUDA;
  const int iters = 4096;
  u32 *out = nullptr, *in = nullptr;
  long long *cyc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = SWB_ERR_CUDA;
  u32 h_in[18];
  for (int i = 0; i < 18; i++) h_in[i] = 0x00010001u * (u32)(i + 1);
  h_in[0] = 0xffffffffu;                                          // -1 in both lanes
  do
  {
    if (cudaMalloc(&out, (size_t)sms * 1024 * sizeof(u32)) != cudaSuccess) break;
    if (cudaMalloc(&in, sizeof h_in) != cudaSuccess) break;
    if (cudaMalloc(&cyc, (size_t)sms * sizeof(long long)) != cudaSuccess) break;
    if (cudaMemcpy(in, h_in, sizeof h_in, cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
    swb_dpx_rate_kernel<<<sms, 1024>>>(out, in, 64, cyc);         // warm-up: clocks, instruction cache
    cudaEventRecord(e0);
    swb_dpx_rate_kernel<<<sms, 1024>>>(out, in, iters, cyc);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h_cyc[256];
    const int n = sms < 256 ? sms : 256;
    if (cudaMemcpy(h_cyc, cyc, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) break;
    long long worst = 1;
    for (int i = 0; i < n; i++) worst = h_cyc[i] > worst ? h_cyc[i] : worst;
    const double warp_instr = (double)iters * 4 * 8 * 2 * (1024 / 32);   // per SM
    *dpx_per_clk_per_sm = warp_instr / (double)worst;
    if (sm_clock_mhz) *sm_clock_mhz = ms > 0 ? (double)worst / (ms * 1e3) : 0.0;
    rc = SWB_OK;
  } while (0);
  (void)cudaGetLastError();
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(out); cudaFree(in); cudaFree(cyc);
  return rc;
}
