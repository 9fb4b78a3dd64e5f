// ubench_discard.cu -- does `discard.global.L2` keep dead activation lines out of HBM on B200?
//
// The pass kernel hands activations from layer to layer through L2; once the next layer has read a row block those lines
// are dead, but they are dirty, and with six lanes' workspaces (177 MB) cycling through a 126 MB L2 they are written
// back to HBM when they are evicted.  `discard.global.L2 [a], 128` drops a line without the write-back.
//
//   ./ubench_discard [mode] [chunks] [reps]
//     mode 0: write chunk after chunk (32 MB each, `chunks` distinct chunks, default 8 = 256 MB > L2)
//     mode 1: the same, every chunk discarded right after it has been written
//   Timed with CUDA events; under `ncu --cache-control none --metrics dram__bytes_write.sum,dram__bytes_read.sum` the
//   per-launch DRAM bytes show whether the write-backs happen.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ubench_discard tools/ubench_discard.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void write_chunk(uint4* p, size_t n16, unsigned v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(v, v + 1, v + 2, static_cast<unsigned>(i));
}
__global__ void read_chunk(const uint4* p, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = p[i];
    acc += x.x ^ x.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void discard_chunk(char* p, size_t bytes) {
  for (size_t off = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 128; off < bytes; off += (size_t)gridDim.x * blockDim.x * 128)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + off) : "memory");
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int chunks = argc > 2 ? atoi(argv[2]) : 8;
  const int reps = argc > 3 ? atoi(argv[3]) : 4;
  const size_t chunk_bytes = 32u << 20;
  char* buf; unsigned* sink;
  CK(cudaMalloc(&buf, chunk_bytes * chunks));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 0, chunk_bytes * chunks));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r)
      for (int c = 0; c < chunks; ++c) {
        char* p = buf + chunk_bytes * c;
        write_chunk<<<148 * 8, 256>>>(reinterpret_cast<uint4*>(p), chunk_bytes / 16, r * 131 + c);
        read_chunk<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(p), chunk_bytes / 16, sink);   // the "next layer"
        if (mode == 1) discard_chunk<<<148 * 4, 256>>>(p, chunk_bytes);
      }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass == 1)
      printf("mode %d (%s): %d reps x %d chunks x 32 MB written + read in %.3f ms = %.1f GB/s of writes\n", mode,
             mode ? "write, read, discard" : "write, read", reps, chunks, ms, reps * (double)chunks * chunk_bytes / ms * 1e-6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
