// wr_probe.cu -- what does this box sustain for a WRITE-ONLY stream issued the way the env kernels issue theirs?
// (timing aid, not product code)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/wr_probe tools/wr_probe.cu
//
// Every CTA owns one or two shared-memory buffers holding a constant pattern and streams "pieces" of `piece` bytes to
// global memory: piece p = blockIdx.x + k * gridDim.x (the env kernels' assignment: at any moment the CTAs write one
// contiguous window of gridDim.x pieces).  Variants: TMA bulk stores (one thread issues, `depth` pieces in flight per CTA)
// or 16-byte st.global by all threads; piece size; CTAs per SM.  Prints GB/s per variant.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)__cvta_generic_to_global(g)),
               "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
}

template <int DEPTH>
__global__ void __launch_bounds__(256) k_tma(uint8_t* out, long long n_pieces, int piece, int spin) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < piece * DEPTH / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01020304u * (i & 63);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int buf = 0;
  for (long long p = blockIdx.x; p < n_pieces; p += gridDim.x) {
    if (threadIdx.x == 0) {
      if (DEPTH == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      uint8_t* g = out + p * (long long)piece;
      const uint8_t* s = smem + buf * piece;
      for (int off = 0; off < piece; off += 16384) bulk_store(g + off, s + off, (uint32_t)min(16384, piece - off));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (spin) {  // stand-in for the work between two pieces (expansion / painting): all threads, `spin` dependent FMAs
      float x = (float)threadIdx.x;
      for (int i = 0; i < spin; ++i) x = x * 1.0001f + 0.5f;
      if (x == 12345.f) out[0] = 1;
      __syncthreads();
    }
    buf = (buf + 1) % DEPTH;
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) k_stg(uint8_t* out, long long n_pieces, int piece) {
  const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
  for (long long p = blockIdx.x; p < n_pieces; p += gridDim.x) {
    uint4* g = reinterpret_cast<uint4*>(out + p * (long long)piece);
    for (int i = threadIdx.x; i < piece / 16; i += blockDim.x) __stcs(g + i, v);
  }
}

int main(int argc, char** argv) {
  const size_t total = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)131072 * 42336;
  uint8_t* d;
  if (cudaMalloc(&d, total + (1 << 20)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(d, 0, total);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const int pieces[] = {42336, 21168, 10080, 4032, 2016 * 16};
  const int ctas[] = {2, 5, 10};
  printf("total %.1f MB, %d SMs\n", total / 1e6, n_sm);
  auto run = [&](const char* tag, int piece, int per_sm, int depth, int spin, int stg) {
    const long long n = (long long)(total / piece);
    const int grid = n_sm * per_sm;
    const size_t smem = (size_t)piece * depth;
    if (!stg && smem * per_sm > 220 * 1024) return;
    if (!stg) {
      if (depth == 1) cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      else cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (stg) k_stg<<<grid, 256>>>(d, n, piece);
      else if (depth == 1) k_tma<1><<<grid, 256, smem>>>(d, n, piece, spin);
      else k_tma<2><<<grid, 256, smem>>>(d, n, piece, spin);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%-4s piece %6d  %2d CTAs/SM  depth %d  spin %5d : %8.1f us  %7.0f GB/s %s\n", tag, piece, per_sm, depth, spin, best * 1e3,
           (double)n * piece / (best * 1e-3) / 1e9, err ? cudaGetErrorString(err) : "");
  };
  for (int piece : pieces)
    for (int c : ctas) {
      run("tma", piece, c, 1, 0, 0);
      run("tma", piece, c, 2, 0, 0);
    }
  for (int c : {5, 10}) { run("tma", 42336, c > 5 ? 5 : c, 1, 2000, 0); run("tma", 21168, c, 2, 1000, 0); }
  for (int piece : {42336, 4032}) for (int c : {5, 8}) run("stg", piece, c, 1, 0, 1);
  // a window that advances linearly at fine grain: pieces of 4 KB with many CTAs = the fill kernel's pattern
  run("stg", 4096, 8, 1, 0, 1);
  run("tma", 4096, 10, 2, 0, 0);
  return 0;
}
