// tq_common.h — shared helpers for the tedq_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/tedq_b200.h"

namespace tq {

// ---- error plumbing (no exceptions across the C ABI) -----------------------
void set_error(const char* fmt, ...);

#define TQ_CUDA_OK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::tq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TQ_E_CUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define TQ_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::tq::set_error(__VA_ARGS__);  \
      return code;                   \
    }                                \
  } while (0)

// ---- device ownership: a plan's tables live on the device that was current when it was created ---------------
inline int current_device() {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess) {
    (void)cudaGetLastError();
    d = -1;
  }
  return d;
}
#define TQ_REQUIRE_DEVICE(plan, what)                                                                          \
  do {                                                                                                         \
    const int _cur = ::tq::current_device();                                                                   \
    TQ_REQUIRE((plan)->device < 0 || _cur == (plan)->device, TQ_E_INVALID,                                     \
               "%s: the plan was created on CUDA device %d but device %d is current (create one plan per "    \
               "device, inside that device's context)", what, (plan)->device, _cur);                          \
  } while (0)

// ---- complex value type -----------------------------------------------------
template <typename R>
struct __align__(2 * sizeof(R)) cx {
  R x, y;
};

template <typename R>
__host__ __device__ __forceinline__ cx<R> mk(R x, R y) {
  cx<R> r;
  r.x = x;
  r.y = y;
  return r;
}
template <typename R>
__host__ __device__ __forceinline__ cx<R> cmul(cx<R> a, cx<R> b) {
  return mk<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + a*b
template <typename R>
__host__ __device__ __forceinline__ cx<R> cfma(cx<R> a, cx<R> b, cx<R> acc) {
  acc.x += a.x * b.x;
  acc.x -= a.y * b.y;
  acc.y += a.x * b.y;
  acc.y += a.y * b.x;
  return acc;
}
// acc + conj(a)*b
template <typename R>
__host__ __device__ __forceinline__ cx<R> cfma_conj(cx<R> a, cx<R> b, cx<R> acc) {
  acc.x += a.x * b.x;
  acc.x += a.y * b.y;
  acc.y += a.x * b.y;
  acc.y -= a.y * b.x;
  return acc;
}
template <typename R>
__host__ __device__ __forceinline__ cx<R> conj_(cx<R> a) {
  return mk<R>(a.x, -a.y);
}
// Re(conj(a)*b)
template <typename R>
__host__ __device__ __forceinline__ R re_conj_mul(cx<R> a, cx<R> b) {
  return a.x * b.x + a.y * b.y;
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ __forceinline__ uint32_t insert_zero_bit(uint32_t v, int p) {
  return ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u));
}

}  // namespace tq
