// Shared device/host helpers for the PSRS engine (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ols_b200.h"

#ifndef __CUDA_ARCH__
#define OLS_HOST_ONLY 1
#endif

// ------------------------------------------------------------------------------------------ errors
void ols_set_error(const char* fmt, ...);
void ols_note_launch(int n);  // process-wide count of the library's own kernel launches (ols_kernel_launches)

#define OLS_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ols_set_error(__VA_ARGS__);         \
      return (code);                      \
    }                                     \
  } while (0)

#define OLS_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ols_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return OLS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define OLS_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    ols_note_launch(1);                                                                   \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ols_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return OLS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// SMs of the current device, queried once per device at run time (B200: 148 = 2 dies x 74; that is also the fallback when
// no device answers); grids are sized in multiples of this
int ols_sm_count(void);
#define OLS_SM_COUNT ols_sm_count()
constexpr int OLS_MAX_ACTIONS = 16;

static inline int ols_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------ records
// meta word: bits 0..7 action, bit 8 terminal, bit 9 "tame": every logged probability of the record lies in
// [2^-40, 2^40] (set at packing time; lets the accept pre-filter skip its per-pop range checks)
__host__ __device__ inline uint32_t ols_pack_meta(int action, int terminal, int tame = 0) {
  return (uint32_t)(action & 0xff) | ((terminal ? 1u : 0u) << 8) | ((tame ? 1u : 0u) << 9);
}
__host__ __device__ inline int ols_meta_tame(uint32_t m) { return (int)((m >> 9) & 1); }
__host__ __device__ inline int ols_meta_action(uint32_t m) { return (int)(m & 0xff); }
__host__ __device__ inline int ols_meta_terminal(uint32_t m) { return (int)((m >> 8) & 1); }

// ---- hot records (ols_queues_t.hot): what the accept rule needs of a <= 4-action record, in 8 bytes.
// Four IEEE half images of the logged probabilities (all >= 0, so the sign bits are free): bit 15 of halves 0 and 1
// carries the action, bit 15 of half 2 the "hot-tame" flag (every probability in [2^-13, 1]: a normal half, relative
// rounding error <= 2^-11).  The pop loop decides from these with a 2^-9 band (threshold estimate within 1 + 2^-10 + 2^-19
// of the exact one) and falls back to the full 32-byte record inside the band or when the flag is clear.
#ifdef __CUDACC__
__device__ __forceinline__ uint2 ols_hot_pack(float4 p, int action, int n_actions) {
  const float v[4] = {p.x, p.y, p.z, p.w};
  unsigned c[4];
  bool tame = ((unsigned)action < 4u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool in_range = (v[j] >= 1.220703125e-4f) && (v[j] <= 1.0f);  // false for NaN
    tame = tame && (j >= n_actions || in_range);
    c[j] = (unsigned)__half_as_ushort(__float2half_rn(in_range ? v[j] : 1.0f)) & 0x7fffu;
  }
  uint2 h;
  h.x = c[0] | (c[1] << 16) | (((unsigned)action & 1u) << 15) | ((((unsigned)action >> 1) & 1u) << 31);
  h.y = c[2] | (c[3] << 16) | ((tame ? 1u : 0u) << 15);
  return h;
}
// TAIL hot record (OLS_HOT_TAIL): what a pop needs of a record whose queue holds one distribution — the action for the
// accept rule, and reward / next state / terminal flag if it is accepted
__device__ __forceinline__ uint2 ols_hot_tail_pack(int action, float reward, int z_next, int terminal) {
  return make_uint2(__float_as_uint(reward),
                    ((uint32_t)z_next & 0x1fffffffu) | (((uint32_t)action & 3u) << 29) | ((terminal ? 1u : 0u) << 31));
}
__device__ __forceinline__ void ols_hot_unpack(uint2 h, float (&p)[4], int& action, bool& tame) {
  action = (int)(((h.x >> 15) & 1u) | ((h.x >> 31) << 1));
  tame = ((h.y >> 15) & 1u) != 0u;
  const unsigned lo = h.x & 0x7fff7fffu, hi = h.y & 0x7fff7fffu;
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
}
#endif

// ------------------------------------------------------------------------------------------ Philox4x32-10
// Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Counter-based: the value
// for (key, counter) is a pure function, so lane k of a lane group can evaluate draw (base + k).
struct ols_u4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline ols_u4 ols_philox4x32_10(ols_u4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)M0 * ctr.x;
    uint64_t p1 = (uint64_t)M1 * ctr.z;
    ols_u4 n;
    n.x = (uint32_t)(p1 >> 32) ^ ctr.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ ctr.w ^ k1;
    n.w = (uint32_t)p0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}

// ---- the simulator's uniform streams (protocol v2; restated in oracle/ols_oracle.c and tests/helpers.py) ---------------
// Draw d of stream `which` (0 = rejection uniforms, 1 = arg-max tie-breaks) of replica `rid` under key `seed` is a 53-bit
// uniform assembled from two Philox4x32-7 words (7 rounds: the Crush-resistant minimum of Salmon et al.; counter-based, so
// any lane can evaluate any draw):
//     blk = d >> 2,  w = d & 3
//     hi  = Philox4x32-7(ctr = {blk_lo, blk_hi, rid_lo, rid_hi ^ (which << 28)},       key = seed)[w]          32 bits
//     lo  = Philox4x32-7(ctr = {blk_lo, blk_hi, rid_lo, rid_hi ^ ((which + 2) << 28)}, key = seed)[w] >> 11    21 bits
//     u   = ((hi << 21) | lo) * 2^-53
// One block of the first stream carries the high words of FOUR consecutive draws, and those 32 bits are all the accept
// pre-filter needs (|float(hi) * 2^-32 - u| <= 2^-32 + 2^-24 u); the low word is generated only when a pop falls inside
// the filter's band or hi < 2^20 — 3e-4 of the pops — so a pop costs a quarter of a 7-round block instead of half of a
// 10-round one (protocol v1), with the exact float64 decision preserved.
// The round keys depend on the seed only: the host computes them once per launch and passes them as a kernel parameter,
// so the device reads them as constant-bank operands.
struct ols_philox_keys {
  uint32_t k0[10], k1[10];
};
inline ols_philox_keys ols_philox_make_keys(uint64_t seed) {
  ols_philox_keys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    K.k0[i] = a;
    K.k1[i] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return K;
}
constexpr int OLS_PHILOX_ROUNDS = 7;
__device__ __forceinline__ ols_u4 ols_philox4x32_7_keyed(ols_u4 ctr, const ols_philox_keys& K) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < OLS_PHILOX_ROUNDS; ++i) {
    const uint64_t p0 = (uint64_t)M0 * ctr.x;  // one IMAD.WIDE each
    const uint64_t p1 = (uint64_t)M1 * ctr.z;
    ols_u4 n;
    n.x = (uint32_t)(p1 >> 32) ^ ctr.y ^ K.k0[i];
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ ctr.w ^ K.k1[i];
    n.w = (uint32_t)p0;
    ctr = n;
  }
  return ctr;
}
// the block of stream `which` that holds draws 4 * blk .. 4 * blk + 3
__device__ __forceinline__ ols_u4 ols_stream_block(const ols_philox_keys& K, uint64_t rid, uint32_t which, uint64_t blk) {
  ols_u4 c;
  c.x = (uint32_t)blk;
  c.y = (uint32_t)(blk >> 32);
  c.z = (uint32_t)rid;
  c.w = (uint32_t)(rid >> 32) ^ (which << 28);
  return ols_philox4x32_7_keyed(c, K);
}
__device__ __forceinline__ uint32_t ols_u4_word(const ols_u4& o, int w) { return w == 0 ? o.x : w == 1 ? o.y : w == 2 ? o.z : o.w; }

// 53-bit uniform in [0,1) from 64 random bits — the same mapping NumPy's Generator.random() uses.
__host__ __device__ inline double ols_u64_to_unit(uint64_t v) { return (double)(v >> 11) * (1.0 / 9007199254740992.0); }
__host__ __device__ inline double ols_u53_to_unit(uint64_t v53) { return (double)v53 * (1.0 / 9007199254740992.0); }

// the exact draw (cold paths: inside the filter's band, the scalar step): both words.  Out of line, and keyed by the SEED
// (round keys re-derived on the fly): taking the address of the kernel's key parameter would move it from the constant bank
// to local memory for the hot path as well.
static __device__ __noinline__ uint64_t ols_stream_u53_seed(uint64_t seed, uint64_t rid, uint32_t which, uint64_t d) {
  uint32_t w[2];
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const uint64_t blk = d >> 2;
    uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = (uint32_t)rid, c3 = (uint32_t)(rid >> 32) ^ ((which + 2u * s) << 28);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int i = 0; i < OLS_PHILOX_ROUNDS; ++i) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      c1 = (uint32_t)p1;
      c3 = (uint32_t)p0;
      c0 = n0;
      c2 = n2;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const int wi = (int)(d & 3);
    w[s] = wi == 0 ? c0 : wi == 1 ? c1 : wi == 2 ? c2 : c3;
  }
  return ((uint64_t)w[0] << 21) | (uint64_t)(w[1] >> 11);
}
__device__ __forceinline__ uint64_t ols_keys_seed(const ols_philox_keys& K) { return (uint64_t)K.k0[0] | ((uint64_t)K.k1[0] << 32); }
__device__ __forceinline__ uint64_t ols_stream_u53(const ols_philox_keys& K, uint64_t rid, uint32_t which, uint64_t d) {
  return ols_stream_u53_seed(ols_keys_seed(K), rid, which, d);
}
__device__ __forceinline__ double ols_stream_uniform(const ols_philox_keys& K, uint64_t rid, uint32_t which, uint64_t d) {
  return ols_u53_to_unit(ols_stream_u53(K, rid, which, d));
}
// float image of a draw from its high word: |error| <= 2^-32 + 2^-24 u, i.e. < 2^-19 relative once u >= 2^-12; the rare
// smaller draws (hi < 2^20: 2.4e-4) are converted from all 53 bits instead.  tiny: u < 2^-60, i.e. u == 0 on the 53-bit
// grid — not resolved by a float32 pre-filter.
__device__ __forceinline__ float ols_stream_float(const ols_philox_keys& K, uint64_t rid, uint32_t which, uint64_t d, uint32_t hi,
                                                  bool& tiny) {
  float f = __uint2float_rn(hi) * 2.3283064365386963e-10f;
  if (hi < (1u << 20)) f = __ull2float_rn(ols_stream_u53(K, rid, which, d)) * 1.1102230246251565e-16f;
  tiny = !(f >= 8.6736174e-19f);
  return f;
}

// Philox4x32-10, two doubles per block: the SYNTHETIC-LOG generator's own stream (ols_synth.cu: row i of the benchmark log is a
// pure function of (seed, i)); the simulator's streams are the 7-round protocol above
__host__ __device__ inline double ols_philox_uniform(uint64_t seed, uint64_t stream_id, uint32_t which, uint64_t d) {
  ols_u4 c;
  uint64_t blk = d >> 1;
  c.x = (uint32_t)blk;
  c.y = (uint32_t)(blk >> 32);
  c.z = (uint32_t)stream_id;
  c.w = (uint32_t)(stream_id >> 32) ^ (which << 28);
  ols_u4 o = ols_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t v = (d & 1) ? (((uint64_t)o.w << 32) | o.z) : (((uint64_t)o.y << 32) | o.x);
  return ols_u64_to_unit(v);
}

// Ring offsets (replica decorrelation, not in the reference): 32-bit murmur3 finaliser hashing.
//   h_rep   = fmix32(seed_lo ^ fmix32(seed_hi ^ fmix32(rid_lo ^ fmix32(rid_hi))))        once per replica
//   queue b : start = mulhi32(fmix32(h_rep ^ ((b + 1) * 0x9E3779B1)), len_b)             uniform in [0, len_b)
//   init    : start = mulhi32(fmix32(h_rep ^ 0x85EBCA6B), n_init)
// (restated in oracle/ols_oracle.c and tests/helpers.py)
__host__ __device__ inline uint32_t ols_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__host__ __device__ inline uint32_t ols_ring_replica_hash(uint64_t ring_seed, uint64_t replica) {
  uint32_t h = ols_fmix32((uint32_t)(replica >> 32));
  h = ols_fmix32((uint32_t)replica ^ h);
  h = ols_fmix32((uint32_t)(ring_seed >> 32) ^ h);
  return ols_fmix32((uint32_t)ring_seed ^ h);
}
__host__ __device__ inline uint32_t ols_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__host__ __device__ inline uint32_t ols_ring_start(uint32_t h_rep, uint32_t bucket, uint32_t len) {
  return ols_mulhi32(ols_fmix32(h_rep ^ ((bucket + 1u) * 0x9E3779B1u)), len);
}
__host__ __device__ inline uint32_t ols_ring_init_start(uint32_t h_rep, uint32_t n_init) {
  return ols_mulhi32(ols_fmix32(h_rep ^ 0x85EBCA6Bu), n_init);
}

// ------------------------------------------------------------------------------------------ PCG64 (NumPy)
// PCG XSL-RR 128/64 exactly as numpy.random.PCG64: state/inc are 128-bit, kept as (hi, lo).
struct ols_u128 {
  uint64_t hi, lo;
};
#define OLS_PCG_MULT_HI 0x2360ED051FC65DA4ull
#define OLS_PCG_MULT_LO 0x4385DF649FCCF645ull

__host__ __device__ inline ols_u128 ols_mul128(ols_u128 a, ols_u128 b) {
  ols_u128 r;
  r.lo = a.lo * b.lo;
#ifdef __CUDA_ARCH__
  r.hi = __umul64hi(a.lo, b.lo) + a.hi * b.lo + a.lo * b.hi;
#else
  r.hi = (uint64_t)(((unsigned __int128)a.lo * b.lo) >> 64) + a.hi * b.lo + a.lo * b.hi;
#endif
  return r;
}
__host__ __device__ inline ols_u128 ols_add128(ols_u128 a, ols_u128 b) {
  ols_u128 r;
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1 : 0);
  return r;
}
__host__ __device__ inline ols_u128 ols_pcg_step(ols_u128 s, ols_u128 inc) {
  ols_u128 m = {OLS_PCG_MULT_HI, OLS_PCG_MULT_LO};
  return ols_add128(ols_mul128(s, m), inc);
}
__host__ __device__ inline uint64_t ols_pcg_output(ols_u128 s) {
  uint64_t x = s.hi ^ s.lo;
  unsigned rot = (unsigned)(s.hi >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}

// SeedSequence(seed).generate_state(4, uint64) for a 64-bit seed, then the PCG64 seeding recipe
// (numpy/random/bit_generator.pyx, _pcg64.pyx; restated in oracle/np_random.py and checked there).
__host__ __device__ inline void ols_pcg64_seed(uint64_t seed, ols_u128* state, ols_u128* inc) {
  const uint32_t INIT_A = 0x43B0D7E5u, MULT_A = 0x931E8875u, INIT_B = 0x8B51F9DDu, MULT_B = 0x58F38DEDu;
  const uint32_t MIX_L = 0xCA01F9DDu, MIX_R = 0x4973F715u;
  uint32_t ent[4] = {(uint32_t)seed, (uint32_t)(seed >> 32), 0u, 0u};
  uint32_t pool[4];
  uint32_t hc = INIT_A;
  auto hashmix = [&](uint32_t v) {
    v ^= hc;
    hc *= MULT_A;
    v *= hc;
    v ^= v >> 16;
    return v;
  };
  auto mix = [&](uint32_t x, uint32_t y) {
    uint32_t r = MIX_L * x - MIX_R * y;
    r ^= r >> 16;
    return r;
  };
  for (int i = 0; i < 4; ++i) pool[i] = hashmix(ent[i]);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = mix(pool[d], hashmix(pool[s]));
  uint32_t hb = INIT_B;
  uint32_t w[8];
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= MULT_B;
    v *= hb;
    v ^= v >> 16;
    w[i] = v;
  }
  uint64_t w0 = w[0] | ((uint64_t)w[1] << 32), w1 = w[2] | ((uint64_t)w[3] << 32);
  uint64_t w2 = w[4] | ((uint64_t)w[5] << 32), w3 = w[6] | ((uint64_t)w[7] << 32);
  ols_u128 initstate = {w0, w1}, initseq = {w2, w3};
  inc->hi = (initseq.hi << 1) | (initseq.lo >> 63);
  inc->lo = (initseq.lo << 1) | 1ull;
  ols_u128 s = {0, 0};
  s = ols_pcg_step(s, *inc);
  s = ols_add128(s, initstate);
  s = ols_pcg_step(s, *inc);
  *state = s;
}
