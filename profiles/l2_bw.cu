// Measured L2 -> SM read bandwidth of this GPU (the practical ceiling bm25_select_kernel's row / postings reads run
// against).  Every CTA streams the same L2-resident buffer with 16-byte read-only loads, four in flight per thread,
// exactly the access shape of the BM25 kernel's pass A.  Build + run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/l2_bw.bin profiles/l2_bw.cu && profiles/l2_bw.bin
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 4) read_kernel(const uint4* __restrict__ buf, size_t n16, int passes, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    // rotate the start so that CTAs do not read the same line at the same time
    size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x + size_t(p) * 7919u * 256u) % n16;
    for (size_t k = 0; k + 3 * stride < n16; k += 4 * stride) {
      size_t a = i + k, b = a + stride, c = b + stride, d = c + stride;
      a -= a >= n16 ? n16 : 0; b -= b >= n16 ? n16 : 0; c -= c >= n16 ? n16 : 0; d -= d >= n16 ? n16 : 0;
      const uint4 x = __ldg(buf + a), y = __ldg(buf + b), z = __ldg(buf + c), w = __ldg(buf + d);
      acc ^= x.x ^ x.y ^ x.z ^ x.w ^ y.x ^ y.y ^ y.z ^ y.w ^ z.x ^ z.y ^ z.z ^ z.w ^ w.x ^ w.y ^ w.z ^ w.w;
    }
  }
  if (acc == 0x12345678u) out[0] = acc;   // keeps the loads alive
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int grid = 4 * prop.multiProcessorCount;
  for (size_t mb : {8, 16, 32, 48, 64, 96}) {
    const size_t bytes = mb << 20, n16 = bytes / 16;
    uint4* buf;
    unsigned* out;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&out, 4);
    cudaMemset(buf, 1, bytes);
    const int passes = int((size_t(4) << 30) / bytes);   // ~4 GB per launch
    read_kernel<<<grid, 256>>>(buf, n16, 2, out);          // warm L2
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0);
      read_kernel<<<grid, 256>>>(buf, n16, passes, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    // bytes actually requested: passes * floor(n16 / (4 * stride)) * 4 * stride * 16
    const size_t stride = size_t(grid) * 256;
    const double moved = double(passes) * double(n16 / (4 * stride)) * 4.0 * double(stride) * 16.0;
    printf("{\"buffer_MiB\": %zu, \"l2_read_GBps\": %.1f, \"ms\": %.3f, \"err\": \"%s\"}\n", mb, moved / (best * 1e-3) / 1e9, best,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(buf);
    cudaFree(out);
  }
  return 0;
}
