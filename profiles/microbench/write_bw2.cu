// write_bw2.cu — which store pattern reaches the write-only HBM ceiling when the rows of the value array are 27
// doubles (216 B) long and therefore start at arbitrary 8-byte-aligned addresses?  Variants of the Jacobian fill:
//   A  warp-span, 128 B aligned spans            (ceiling, write_bw.cu "span 27x32")
//   B  warp-span, spans shifted by 8 B           (every warp store straddles sectors)
//   C  block-per-run of 255 rows, 128 threads, block-strided 8 B stores, run starts unaligned (the current kernel)
//   D  as C but the loop index is aligned to 256 B: every warp store is one full 256 B segment, edges masked
//   E  as D with 16 B stores (double2), aligned to 512 B per warp store
//   F  persistent grid (148 x 8 blocks) over runs, aligned 16 B stores
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

#define NS 27
__global__ void k_span(double* out, int64_t nwarps, const double* S)
{
  __shared__ double sS[32];
  if (threadIdx.x < NS) sS[threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= nwarps) return;
  double* o = out + w * 864;
  int mm = lane % NS;
#pragma unroll
  for (int it = 0; it < NS; ++it) { o[it * 32 + lane] = sS[mm]; mm += 5; if (mm >= NS) mm -= NS; }
}

// runs: run r covers doubles [start[r], start[r+1]) of out
__global__ void __launch_bounds__(128) k_run_unaligned(double* out, const int64_t* __restrict__ start, const double* S)
{
  __shared__ double sS[32];
  if (threadIdx.x < NS) sS[threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  const int64_t a = start[blockIdx.x];
  const int n = (int)(start[blockIdx.x + 1] - a);
  double* o = out + a;
  int mm = threadIdx.x % NS;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 128) { o[i] = sS[mm]; mm += 128 % NS; if (mm >= NS) mm -= NS; }
}

__global__ void __launch_bounds__(128) k_run_aligned8(double* out, const int64_t* __restrict__ start, const double* S)
{
  __shared__ double sS[32];
  if (threadIdx.x < NS) sS[threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  const int64_t a = start[blockIdx.x], b = start[blockIdx.x + 1];
  const int64_t a0 = a & ~(int64_t)31;                 // 256 B aligned
  int64_t i = a0 + threadIdx.x;
  int mm = (int)((i - a + 128 * NS) % NS);             // (i - a) mod NS, i may be < a
#pragma unroll 4
  for (; i < b; i += 128) { if (i >= a) out[i] = sS[mm]; mm += 128 % NS; if (mm >= NS) mm -= NS; }
}

__global__ void __launch_bounds__(128) k_run_aligned16(double* out, const int64_t* __restrict__ start, const double* S)
{
  __shared__ double sS[64];
  if (threadIdx.x < 2 * NS) sS[threadIdx.x] = S[threadIdx.x % NS];
  __syncthreads();
  const int64_t a = start[blockIdx.x], b = start[blockIdx.x + 1];
  const int64_t a0 = a & ~(int64_t)63;                 // 512 B aligned
  int64_t i = a0 + 2 * threadIdx.x;
  int mm = (int)((i - a + 256 * NS) % NS);
#pragma unroll 4
  for (; i < b; i += 256) {
    if (i >= a && i + 1 < b) *reinterpret_cast<double2*>(out + i) = make_double2(sS[mm], sS[mm + 1]);
    else { if (i >= a) out[i] = sS[mm]; if (i + 1 >= a && i + 1 < b) out[i + 1] = sS[mm + 1]; }
    mm += 256 % NS; if (mm >= NS) mm -= NS;
  }
}

// persistent: block b takes runs b, b+grid, ...
__global__ void __launch_bounds__(256) k_run_persistent16(double* out, const int64_t* __restrict__ start, int nruns, const double* S)
{
  __shared__ double sS[64];
  if (threadIdx.x < 2 * NS) sS[threadIdx.x] = S[threadIdx.x % NS];
  __syncthreads();
  for (int r = blockIdx.x; r < nruns; r += gridDim.x) {
    const int64_t a = start[r], b = start[r + 1];
    const int64_t a0 = a & ~(int64_t)63;
    int64_t i = a0 + 2 * threadIdx.x;
    int mm = (int)((i - a + 512 * NS) % NS);
#pragma unroll 4
    for (; i < b; i += 512) {
      if (i >= a && i + 1 < b) *reinterpret_cast<double2*>(out + i) = make_double2(sS[mm], sS[mm + 1]);
      else { if (i >= a) out[i] = sS[mm]; if (i + 1 >= a && i + 1 < b) out[i + 1] = sS[mm + 1]; }
      mm += 512 % NS; if (mm >= NS) mm -= NS;
    }
  }
}

// flat: ignore run structure entirely; thread t of the grid writes doubles 2t, 2t+1 of [lo, hi) with value S[(i - phase) % NS]
// (what a kernel could do if a whole child block were one run): upper bound for 16 B stores with a modulo per element
__global__ void __launch_bounds__(256) k_flat16(double* out, int64_t n, const double* S)
{
  __shared__ double sS[64];
  if (threadIdx.x < 2 * NS) sS[threadIdx.x] = S[threadIdx.x % NS];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * 512;
  int64_t i = blockIdx.x * (int64_t)512 + 2 * threadIdx.x;
  int mm = (int)(i % NS);
  const int step = (int)(stride % NS);
  for (; i + 1 < n; i += stride) {
    *reinterpret_cast<double2*>(out + i) = make_double2(sS[mm], sS[mm + 1]);
    mm += step; if (mm >= NS) mm -= NS;
  }
}

int main()
{
  const int64_t rows = 17040642;
  const int64_t n = rows * NS;
  double* d; double* S;
  CK(cudaMalloc(&d, (n + 64) * 8)); CK(cudaMalloc(&S, 32 * 8)); CK(cudaMemset(S, 0, 256));
  // runs of 255 rows separated by 2 irregular rows of 18 entries (an x-line of the 256^3 mesh)
  const int nruns = (int)(rows / 257);
  int64_t* h = new int64_t[nruns + 1];
  int64_t p = 5;
  for (int r = 0; r < nruns; ++r) { h[r] = p; p += 255 * NS + 36; }
  h[nruns] = p;
  // the run kernels write [h[r], h[r+1]) including the 36 entries of the irregular rows: same bytes, contiguous coverage
  int64_t* ds; CK(cudaMalloc(&ds, (nruns + 1) * 8)); CK(cudaMemcpy(ds, h, (nruns + 1) * 8, cudaMemcpyHostToDevice));
  const int64_t nrun_bytes = (p - 5) * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  auto report = [&](const char* name, int64_t bytes) { cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms /= 5; printf("%-44s %.3f ms  %.0f GB/s\n", name, ms, bytes / ms / 1e6); };
  for (int rep = 0; rep < 2; ++rep) {
    int64_t nw = n / 864;
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) cudaMemsetAsync(d, 0, n * 8); cudaEventRecord(e1); report("cudaMemset", n * 8);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span<<<(unsigned)((nw * 32 + 127) / 128), 128>>>(d, nw, S); cudaEventRecord(e1); report("A span 27x32 aligned", nw * 864 * 8);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span<<<(unsigned)((nw * 32 + 127) / 128), 128>>>(d + 1, nw, S); cudaEventRecord(e1); report("B span 27x32 shifted 8 B", nw * 864 * 8);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_span<<<(unsigned)((nw * 32 + 127) / 128), 128>>>(d + 4, nw, S); cudaEventRecord(e1); report("B' span 27x32 shifted 32 B", nw * 864 * 8);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_run_unaligned<<<nruns, 128>>>(d, ds, S); cudaEventRecord(e1); report("C run/block unaligned 8 B stores", nrun_bytes);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_run_aligned8<<<nruns, 128>>>(d, ds, S); cudaEventRecord(e1); report("D run/block 256 B aligned 8 B stores", nrun_bytes);
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_run_aligned16<<<nruns, 128>>>(d, ds, S); cudaEventRecord(e1); report("E run/block 512 B aligned 16 B stores", nrun_bytes);
    for (int g : {148 * 2, 148 * 4, 148 * 8}) {
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_run_persistent16<<<g, 256>>>(d, ds, nruns, S); cudaEventRecord(e1);
      char bb[64]; snprintf(bb, 64, "F persistent grid=%d 16 B stores", g); report(bb, nrun_bytes);
    }
    for (int g : {148 * 4, 148 * 8, 148 * 16}) {
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_flat16<<<g, 256>>>(d, n, S); cudaEventRecord(e1);
      char bb[64]; snprintf(bb, 64, "G flat grid=%d 16 B stores + modulo", g); report(bb, n * 8);
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
