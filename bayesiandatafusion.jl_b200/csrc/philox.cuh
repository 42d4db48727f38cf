// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the standard-normal stream the engine uses when no
// noise is injected. The stream is a pure function of (seed, sweep, stream id, global row, latent index), so the
// draws do not depend on the number of GPUs, the work-item order or the chunking of heavy rows.
#pragma once
#include <cstdint>

namespace bdf {

struct u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline u32x4 philox4x32_10(u32x4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)M0 * ctr.x;
    const uint64_t p1 = (uint64_t)M1 * ctr.z;
    u32x4 n;
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

// Stream ids: the sampler kind in the high byte, the entity / relation index (< 2^24) below it, so the streams of different
// samplers can never overlap however many entities or relations a model has.
enum PhiloxKind : uint32_t { PHILOX_ROW = 0, PHILOX_NW = 1, PHILOX_BETA = 2, PHILOX_LAMBDA_BETA = 3, PHILOX_ALPHA = 4, PHILOX_RELFEAT = 5 };
__host__ __device__ constexpr uint32_t philox_stream(PhiloxKind kind, uint32_t index) { return ((uint32_t)kind << 24) | (index & 0xFFFFFFu); }

// two independent uniforms in (0,1) with 52 random bits each
__host__ __device__ inline void philox_uniform2(uint64_t seed, uint64_t sweep, uint32_t stream, uint64_t row, uint32_t idx,
                                                double& u1, double& u2) {
  u32x4 c;
  c.x = (uint32_t)row;
  c.y = (uint32_t)(row >> 32) ^ (idx << 8);
  c.z = (uint32_t)sweep;
  c.w = (uint32_t)(sweep >> 32) ^ stream;  // the whole 32-bit stream id (see philox_stream): sweep counters stay far below 2^32
  const u32x4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
  u1 = ((double)(a >> 12) + 0.5) * (1.0 / 4503599627370496.0);
  u2 = ((double)(b >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

// standard normal number `j` of (stream, row): Box–Muller on pair j/2
__host__ __device__ inline double philox_normal(uint64_t seed, uint64_t sweep, uint32_t stream, uint64_t row, int j) {
  double u1, u2;
  philox_uniform2(seed, sweep, stream, row, (uint32_t)(j >> 1), u1, u2);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  return (j & 1) ? r * s : r * c;
}

// both normals of pair j/2 at once (z0 = number 2·(j/2), z1 = number 2·(j/2)+1): the same values philox_normal returns one at a time
__host__ __device__ inline void philox_normal_pair(uint64_t seed, uint64_t sweep, uint32_t stream, uint64_t row, int pair, double& z0, double& z1) {
  double u1, u2;
  philox_uniform2(seed, sweep, stream, row, (uint32_t)pair, u1, u2);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

}  // namespace bdf
