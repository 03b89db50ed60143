// dft64.cu -- kernel and launcher of the fp64 engine's DFT stage (rate_kernels_f64.cuh). Its own translation unit so
// that the kernel can be rebuilt and profiled without recompiling engine.cu. With -DB200RATE_EMU (tests/emu only) the
// launch is a serial loop over the same program.
#include "engine.hpp"
#include "rate_kernels_f64.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifndef B200RATE_EMU
#include <cuda_runtime.h>
#endif

namespace b200rate {

#ifndef B200RATE_EMU
extern __shared__ __align__(16) unsigned char rr_smem_raw[];

// DFT stage of the fp64 engine (rate_kernels_f64.cuh): a CTA is `groups` independent groups of `gthreads` threads,
// each a persistent worker with its own buffer behind the shared pass-twiddle rows.
template <bool UPL>
__global__ void __launch_bounds__(kD64MaxThreads, 1) dft64_kernel(const __grid_constant__ Dft64Params dp, long long nwork)
{
  CD *twt = reinterpret_cast<CD *>(rr_smem_raw);
  for (int i = threadIdx.x; i < dp.ntw; i += blockDim.x) twt[i] = dp.tw[i];
  const int gi = (int)threadIdx.x / dp.gthreads;
  const Grp g{(int)threadIdx.x - gi * dp.gthreads, dp.gthreads, 1 + gi};
  CD *buf = twt + dp.ntw + (size_t)gi * dp.group_slots;
  __shared__ D64Item items[kD64MaxGroups][2];
  __syncthreads();
  int w = (int)blockIdx.x * dp.groups + gi;               // 32-bit work counters (the host checks the range)
  const int stride = (int)gridDim.x * dp.groups, nw = (int)nwork;
  if (w < nw && g.tid == 0) items[gi][0] = d64_make_item(dp, w);
  for (int n = 0; w < nw; w += stride, n ^= 1) {
    const int next = w + stride < nw ? w + stride : -1;
    dft64_program<UPL>(dp, g, twt, items[gi], n, next, buf);
  }
}

namespace {
struct D64Launch { int max_dyn = -1; std::map<std::pair<int, size_t>, int> resident; };
std::mutex g_d64_mu;
std::map<std::pair<int, const void *>, D64Launch> g_d64_launch;     // (device, kernel)

int d64_fail(cudaError_t e, const char *what)
{
  set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? RR_ENOMEM : RR_INTERNAL;
}
#define D64_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return d64_fail(e_, #expr); } while (0)

// Resident CTAs per device for (threads, smem); the opt-in shared-memory limit is raised once per (device, kernel) to
// the device maximum and never lowered (handles on different host threads share the kernels).
template <class Kernel> int d64_geometry(Kernel kernel, int threads, size_t smem, long long *resident)
{
  int dev = 0;
  D64_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_d64_mu);
  D64Launch &l = g_d64_launch[std::make_pair(dev, reinterpret_cast<const void *>(kernel))];
  if (l.max_dyn < 0) {
    cudaFuncAttributes fa;
    D64_TRY(cudaFuncGetAttributes(&fa, kernel));
    int optin = 0;
    D64_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int max_dyn = optin - static_cast<int>(fa.sharedSizeBytes);
    D64_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    l.max_dyn = max_dyn;
  }
  if (static_cast<long long>(smem) > l.max_dyn) { set_last_error("dft64_kernel needs more shared memory than an SM has"); return RR_INTERNAL; }
  auto it = l.resident.find(std::make_pair(threads, smem));
  if (it == l.resident.end()) {
    int occ = 0, sms = 0;
    D64_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    D64_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (occ < 1) { set_last_error("dft64_kernel does not fit on an SM"); return RR_INTERNAL; }
    it = l.resident.emplace(std::make_pair(threads, smem), occ * sms).first;
  }
  *resident = it->second;
  return RR_OK;
}
}  // namespace
#endif  // !B200RATE_EMU

// nwork = blocks x lanes
int launch_dft64(const Dft64Params &dp, long long nwork, void *stream)
{
  if (nwork <= 0) return RR_OK;
#ifdef B200RATE_EMU
  (void)stream;
  std::vector<CD> mem(static_cast<size_t>(dp.group_slots) + 1);
  const Grp g{0, 1, 0};
  for (long long w = 0; w < nwork; ++w) {
    D64Item items[2];
    items[0] = d64_make_item(dp, w);
    if (dp.mode == D64_UP2 && dp.up_bits > 1) dft64_program<true>(dp, g, dp.tw, items, 0, -1, mem.data());
    else dft64_program<false>(dp, g, dp.tw, items, 0, -1, mem.data());
  }
  return RR_OK;
#else
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(CD) * (static_cast<size_t>(dp.ntw) + static_cast<size_t>(dp.groups) * dp.group_slots);
  const int threads = dp.groups * dp.gthreads;
  auto go = [&](auto kernel) -> int {
    long long resident = 0;
    const int rc = d64_geometry(kernel, threads, smem, &resident);
    if (rc != RR_OK) return rc;
    const long long ctas = (nwork + dp.groups - 1) / dp.groups;
    kernel<<<static_cast<unsigned>(std::min(ctas, resident)), threads, smem, s>>>(dp, nwork);
    D64_TRY(cudaGetLastError());
    return RR_OK;
  };
  if (dp.mode == D64_UP2 && dp.up_bits > 1) return go(dft64_kernel<true>);
  return go(dft64_kernel<false>);
#endif
}

}  // namespace b200rate
