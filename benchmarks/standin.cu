// Bench-only stand-in for one layer of attention ("a dummy attention-like kernel, HBM-bound,
// ~T_layer, runs on the main stream").  It streams `n_vec` 16-byte vectors through the SMs and -- like an engine's
// KV-write epilogue would -- releases the layer's ready flag from inside the kernel (last block), so the producer
// pays no extra launch or stream operation per layer.  Not part of the product libraries.
#include <cuda_runtime_api.h>
#include <cstdint>

__global__ void __launch_bounds__(256) standin_layer_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_vec,
                                                            uint32_t* ready_flag, uint32_t value, unsigned int* counter, int fma_iters)
{
  // fma_iters > 0 adds a dependent FMA chain per vector: arithmetic intensity knob, so the stand-in can model a
  // layer that is compute-bound (prefill attention / MLP) rather than one that saturates HBM on its own
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    uint4 v = in[i];
    float a = __uint_as_float(v.x), b = __uint_as_float(v.y);
    for (int k = 0; k < fma_iters; ++k) {
      a = fmaf(a, 1.0001f, b);
      b = fmaf(b, 0.9999f, a);
    }
    v.x = __float_as_uint(a) + 1;
    v.y = __float_as_uint(b);
    out[i] = v;
  }
  if (ready_flag == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(counter, 1u);
    if (old == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ready_flag), "r"(value) : "memory");
    }
  }
}

extern "C" cudaError_t standin_attention_layer(const void* in, void* out, size_t n_vec, uint32_t* ready_flag, uint32_t value,
                                               unsigned int* counter, int blocks, int fma_iters, cudaStream_t stream)
{
  standin_layer_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), n_vec, ready_flag,
                                                   value, counter, fma_iters);
  return cudaGetLastError();
}

// Bench-only: what the v1 block manager's D2D path and UCX cuda_ipc do -- one cudaMemcpyAsync per (block, layer, outer)
// chunk (lib/llm/src/block_manager/block/transfer/cuda.rs:299-391).  Host arrays of addresses.
extern "C" cudaError_t memcpy_per_chunk(void* const* src, void* const* dst, size_t n, size_t bytes, cudaStream_t stream)
{
  for (size_t i = 0; i < n; ++i) {
    cudaError_t e = cudaMemcpyAsync(dst[i], src[i], bytes, cudaMemcpyDefault, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
