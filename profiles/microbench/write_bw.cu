// write_bw.cu — calibrates the write-only HBM ceiling on B200 for the store patterns the assembly kernel can use.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_store8(double* out, int64_t n, double v)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_store16(double2* out, int64_t n2, double v)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) out[i] = make_double2(v, v);
}
__global__ void k_store32(double4* out, int64_t n4, double v)
{
  double4 w = make_double4(v, v, v, v);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) out[i] = w;
}
// one warp writes a contiguous span of 864 doubles (32 rows x 27), non-persistent grid like the assembly kernel
__global__ void k_span(double* out, int64_t nwarps, const double* S)
{
  __shared__ double sS[32];
  if (threadIdx.x < 27) sS[threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= nwarps) return;
  double* o = out + w * 864;
  int mm = lane % 27;
#pragma unroll
  for (int it = 0; it < 27; ++it) { o[it * 32 + lane] = sS[mm]; mm += 5; if (mm >= 27) mm -= 27; }
}
__global__ void k_span_cs(double* out, int64_t nwarps, const double* S)
{
  __shared__ double sS[32];
  if (threadIdx.x < 27) sS[threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= nwarps) return;
  double* o = out + w * 864;
  int mm = lane % 27;
#pragma unroll
  for (int it = 0; it < 27; ++it) { __stcs(o + it * 32 + lane, sS[mm]); mm += 5; if (mm >= 27) mm -= 27; }
}

int main()
{
  const int64_t n = 460078858 / 864 * 864;   // ~ nnz of the M2-256 matrix
  double* d; double* S;
  CK(cudaMalloc(&d, n * 8)); CK(cudaMalloc(&S, 32 * 8)); CK(cudaMemset(S, 0, 256));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  auto report = [&](const char* name) { cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms /= 5; printf("%-28s %.3f ms  %.0f GB/s\n", name, ms, n * 8.0 / ms / 1e6); };
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) cudaMemsetAsync(d, 0, n * 8); cudaEventRecord(e1); report("cudaMemset");
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_store8<<<g, 256>>>(d, n, 1.0); cudaEventRecord(e1);
      char b[64]; snprintf(b, 64, "store8 grid=%d", g); report(b);
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_store16<<<g, 256>>>((double2*)d, n / 2, 1.0); cudaEventRecord(e1);
      snprintf(b, 64, "store16 grid=%d", g); report(b);
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_store32<<<g, 256>>>((double4*)d, n / 4, 1.0); cudaEventRecord(e1);
      snprintf(b, 64, "store32 grid=%d", g); report(b);
    }
    int64_t nw = n / 864;
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span<<<(unsigned)((nw * 32 + 255) / 256), 256>>>(d, nw, S); cudaEventRecord(e1); report("span 27x32 (st)");
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span_cs<<<(unsigned)((nw * 32 + 255) / 256), 256>>>(d, nw, S); cudaEventRecord(e1); report("span 27x32 (st.cs)");
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span<<<(unsigned)((nw * 32 + 127) / 128), 128>>>(d, nw, S); cudaEventRecord(e1); report("span 27x32 block128");
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
