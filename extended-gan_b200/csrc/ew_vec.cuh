// 16-byte vector access along the channel axis of NHWC tensors for the element-wise / reduction kernels around the convs.
#pragma once
#include "common.cuh"

namespace cgat {

// element access for T in {float, bf16}: V consecutive channels (16 bytes when V * sizeof(T) == 16)
template <typename T, int V>
__device__ __forceinline__ void na_load(const T* p, float (&v)[V]) {
  if constexpr (V == 1) {
    v[0] = DT<T>::to_f(p[0]);
  } else if constexpr (sizeof(T) == 4) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
// the same access split in two: the raw 16-byte (or scalar) load first, the conversion later -- lets a loop keep several
// loads in flight per thread at 4 registers each
template <typename T, int V> struct NaRaw { uint4 q; };
template <typename T> struct NaRaw<T, 1> { T q; };
template <typename T, int V>
__device__ __forceinline__ NaRaw<T, V> na_load_raw(const T* p) {
  NaRaw<T, V> r;
  if constexpr (V == 1) r.q = p[0];
  else r.q = *reinterpret_cast<const uint4*>(p);
  return r;
}
template <typename T, int V>
__device__ __forceinline__ void na_unpack(const NaRaw<T, V>& r, float (&v)[V]) {
  if constexpr (V == 1) {
    v[0] = DT<T>::to_f(r.q);
  } else if constexpr (sizeof(T) == 4) {
    v[0] = __uint_as_float(r.q.x); v[1] = __uint_as_float(r.q.y); v[2] = __uint_as_float(r.q.z); v[3] = __uint_as_float(r.q.w);
  } else {
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
template <typename T, int V>
__device__ __forceinline__ void na_store(T* p, const float (&v)[V]) {
  if constexpr (V == 1) {
    p[0] = DT<T>::from_f(v[0]);
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// vector width the channel count and the pointers allow: 16 bytes, else scalar
inline int ew_vec(int dtype, int c, const void* a, const void* b = nullptr, const void* o = nullptr) {
  const int v = dtype == CGAT_F32 ? 4 : 8;
  if (c % v) return 1;
  if (!aligned16(a) || (b && !aligned16(b)) || (o && !aligned16(o))) return 1;
  return v;
}

}  // namespace cgat
