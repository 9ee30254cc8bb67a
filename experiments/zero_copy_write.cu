// Microbenchmark behind DESIGN.md section 3 ("direct output"): how fast can SM stores fill page-locked host memory over
// PCIe, against the copy engine, for the sizes of a 752x480 / 1280x720 / 3840x2160 cloud?  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zc_write zero_copy_write.cu && ./zc_write
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// linear: thread i writes element i (a warp = 512 contiguous bytes, consecutive warps consecutive chunks)
__global__ void fill_linear(float4 *dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcs(dst + i, make_float4((float)i, 1.f, 2.f, 1.f));
}
// strips: a warp owns a 32-point column block and walks down `rows` rows of a `row_pts`-point-wide image (the store
// pattern of the fused callback kernel: 512 B per row, row_pts * 16 B apart)
__global__ void fill_strips(float4 *dst, int row_pts, int n_rows, int strip_rows) {
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n_cb = (row_pts + 31) / 32, n_strip = (n_rows + strip_rows - 1) / strip_rows;
  for (int unit = blockIdx.x * wpc + wic; unit < n_cb * n_strip; unit += gridDim.x * wpc) {
    const int strip = unit / n_cb, cb = unit - strip * n_cb;
    const int x = cb * 32 + lane;
    for (int y = strip * strip_rows; y < min((strip + 1) * strip_rows, n_rows); ++y)
      if (x < row_pts) __stcs(dst + (size_t)y * row_pts + x, make_float4((float)x, (float)y, 2.f, 1.f));
  }
}
// v8: 32-byte stores (two points per lane): a warp = 1 KB contiguous
__global__ void fill_linear32(float4 *dst, size_t n) {
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n; i += (size_t)gridDim.x * blockDim.x * 2) {
    const float4 a = make_float4((float)i, 1.f, 2.f, 1.f);
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + i), "f"(a.x), "f"(a.y), "f"(a.z),
                 "f"(a.w), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                 : "memory");
  }
}

int main() {
  const struct { int w, h; } frames[] = {{752, 480}, {1280, 720}, {3840, 2160}};
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (auto f : frames) {
    const int row_pts = f.w - 80, n_rows = f.h - 80;
    const size_t n = (size_t)row_pts * n_rows, bytes = n * 16;
    float4 *h, *hd, *d;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    CK(cudaMalloc(&d, bytes));
    auto time_it = [&](auto fn, const char *name) {
      float best = 1e30f, sum = 0;
      const int it = 30;
      for (int i = 0; i < it + 5; ++i) {
        CK(cudaEventRecord(e0, s));
        fn();
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 5) { best = ms < best ? ms : best; sum += ms; }
      }
      printf("%4dx%-4d %-34s best %8.1f us (%5.1f GB/s)  mean %8.1f us (%5.1f GB/s)\n", f.w, f.h, name, best * 1e3,
             bytes / best / 1e6, sum / it * 1e3, bytes / (sum / it) / 1e6);
    };
    time_it([&] { CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s)); }, "copy engine D2H");
    for (int grid : {148, 148 * 4, 148 * 16})
      time_it([&] { fill_linear<<<grid, 256, 0, s>>>(hd, n); }, grid == 148 ? "SM stores, linear, 148 CTAs" : grid == 592 ? "SM stores, linear, 592 CTAs" : "SM stores, linear, 2368 CTAs");
    time_it([&] { fill_linear32<<<148 * 4, 256, 0, s>>>(hd, n); }, "SM stores, linear, 32 B per lane");
    for (int strip : {2, 8, 64})
      time_it([&] { fill_strips<<<148 * 6, 128, 0, s>>>(hd, row_pts, n_rows, strip); },
              strip == 2 ? "SM stores, column strips of 2 rows" : strip == 8 ? "SM stores, column strips of 8 rows" : "SM stores, column strips of 64 rows");
    time_it([&] { fill_linear<<<148 * 4, 256, 0, s>>>(d, n); }, "(device memory, linear)");
    CK(cudaFree(d));
    CK(cudaFreeHost(h));
  }
  return 0;
}
