// Microbenchmark: how fast can 1-D bulk-async copies (cp.async.bulk, global -> shared) stream a large array on B200?
// Each block loops over chunks of CH bytes with NBUF staging buffers (NBUF copies in flight per block); nothing is computed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NBUF>
__global__ void k_stream(const double* __restrict__ src, int64_t nchunks, int chunk_bytes, int split, double* out, const double* __restrict__ xsrc = nullptr, int nx = 0, int64_t xstride = 0)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[NBUF];
  const int t = threadIdx.x;
  constexpr int NBUFS_TOTAL = NBUF;
  if (t == 0) {
    for (int b = 0; b < NBUF; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t chunk, int b) {
    const char* g = (const char*)src + chunk * (int64_t)chunk_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(chunk_bytes + nx * 1056) : "memory");
    for (int k = 0; k < nx; ++k)      // nx small segment copies (the x segments of the SpMV), 1056 B each, from xsrc
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)NBUFS_TOTAL * chunk_bytes + (b * 9 + k) * 1056)),
                   "l"(xsrc + chunk * 128 + (k % 3) * 264 + (k / 3) * xstride), "r"(1056), "r"(smem_u32(&bar[b]))
                   : "memory");
    const int piece = chunk_bytes / split;
    for (int k = 0; k < split; ++k)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)b * chunk_bytes + k * piece)),
                   "l"(g + k * piece), "r"(piece), "r"(smem_u32(&bar[b]))
                   : "memory");
  };
  double acc = 0.0;
  int64_t chunk = blockIdx.x;
  if (t == 0) for (int b = 0; b < NBUF; ++b) if (chunk + (int64_t)b * gridDim.x < nchunks) issue(chunk + (int64_t)b * gridDim.x, b);
  uint32_t phasebits = 0;
  int it = 0;
  for (; chunk < nchunks; chunk += gridDim.x, ++it) {
    const int b = it % NBUF;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"((phasebits >> b) & 1u) : "memory");
    phasebits ^= 1u << b;
    acc += ((const double*)(smem + (size_t)b * chunk_bytes))[t];
    __syncthreads();
    if (t == 0 && chunk + (int64_t)NBUF * gridDim.x < nchunks) issue(chunk + (int64_t)NBUF * gridDim.x, b);
  }
  if (acc == 1.2345) out[0] = acc;
}

__global__ void k_ldg(const double2* __restrict__ src, int64_t n2, double* out)
{
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(src + i));
    acc += v.x + v.y;
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int NBUF>
static void run(const double* d, size_t bytes, int chunk_bytes, int blocks_per_sm, int threads, int split, double* out, const double* xs = nullptr, int nx = 0)
{
  const int64_t nchunks = bytes / chunk_bytes;
  const size_t smem = (size_t)NBUF * chunk_bytes + (size_t)NBUF * 9 * 1056;
  cudaFuncSetAttribute(k_stream<NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_stream<NBUF><<<148 * blocks_per_sm, threads, smem>>>(d, nchunks, chunk_bytes, split, out, xs, nx, 66048);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("bulk nx=%d chunk=%6d B  nbuf=%d  blocks/SM=%d  threads=%3d  split=%d  smem/SM=%4zu KB : %7.1f GB/s  %s\n", nx, chunk_bytes, NBUF, blocks_per_sm, threads, split,
         smem * blocks_per_sm / 1024, nchunks * (double)chunk_bytes / best * 1e-6, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
  const size_t bytes = (size_t)3680 << 20;
  double *d, *out;
  cudaMalloc(&d, bytes); cudaMalloc(&out, 8);
  cudaMemset(d, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int g : {8, 16}) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      k_ldg<<<148 * g, 256>>>((const double2*)d, bytes / 16, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("ldg.v2 read-only, %d blocks/SM x 256 : %7.1f GB/s\n", g, bytes / best * 1e-6);
  }
  double* xs; cudaMalloc(&xs, (size_t)160 << 20); cudaMemset(xs, 0, (size_t)160 << 20);
  run<1>(d, bytes, 27648, 5, 128, 1, out);
  run<1>(d, bytes, 27648, 5, 128, 1, out, xs, 9);
  run<1>(d, bytes, 27648, 5, 128, 1, out, xs, 3);
  run<2>(d, bytes, 27648, 3, 128, 1, out, xs, 9);
  run<1>(d, bytes, 27648, 7, 128, 1, out);
  run<2>(d, bytes, 27648, 3, 128, 1, out);
  run<2>(d, bytes, 27648, 4, 128, 1, out);
  run<3>(d, bytes, 27648, 2, 128, 1, out);
  run<1>(d, bytes, 27648, 7, 128, 4, out);
  run<2>(d, bytes, 27648, 4, 128, 4, out);
  run<2>(d, bytes, 13824, 8, 64, 1, out);
  run<4>(d, bytes, 13824, 4, 64, 1, out);
  run<2>(d, bytes, 6912, 16, 32, 1, out);
  run<4>(d, bytes, 6912, 8, 128, 1, out);
  run<8>(d, bytes, 6912, 4, 128, 1, out);
  run<4>(d, bytes, 55296, 1, 128, 1, out);
  run<2>(d, bytes, 55296, 2, 128, 1, out);
  run<2>(d, bytes, 110592, 1, 128, 1, out);
  run<4>(d, bytes, 27648, 2, 128, 1, out);
  run<8>(d, bytes, 27648, 1, 128, 1, out);
  return 0;
}
