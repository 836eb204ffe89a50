/* ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-C restatement of oracle/allreduce_oracle.py (same
 * citations: torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-33 scale + allreduce,
 * :57-92 cast; SURVEY.md 7.3-4 / 8(c) parity definition).  Built by oracle/build_oracle.py into
 * oracle/_build/liballreduce_oracle.so; tests cross-check it bit-for-bit against the numpy version
 * and use it where numpy is too slow.  Never linked into the product.
 *
 *   wire_r = cast_wire(f32(in_r) * pre)          pre  = scale (PRE) | 1 (POST)
 *   acc    = f32(wire_0) + f32(wire_1) + ...     fp32, rank order
 *   out    = cast_out(f32(cast_wire(acc * post)))
 * dtype codes match include/tok8s.h: 0 = f32, 1 = bf16 (uint16 bits), 2 = f16.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static float bf16_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static uint16_t f32_to_bf16(float f) { /* round to nearest even; NaN -> 0x7FFF */
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7FFF;
  uint32_t lsb = (u >> 16) & 1u;
  return (uint16_t)((u + 0x7FFFu + lsb) >> 16);
}

static float f16_to_f32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else { /* subnormal */
      int e = -1;
      do { e++; man <<= 1; } while (!(man & 0x400u));
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7F800000u | (man << 13);
  } else {
    u = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static uint16_t f32_to_f16(float f) { /* round to nearest even, IEEE binary16 */
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t abs = u & 0x7fffffffu;
  if (abs > 0x7f800000u) return (uint16_t)(sign | 0x7E00u);  /* NaN */
  if (abs >= 0x47800000u) return (uint16_t)(sign | 0x7C00u); /* >= 65536 (and inf) -> inf */
  if (abs < 0x33000000u) return (uint16_t)sign;              /* < 2^-25 -> 0 */
  const int32_t exp = (int32_t)(abs >> 23) - 127;
  const uint32_t man = (abs & 0x7FFFFFu) | 0x800000u;        /* 24-bit significand */
  uint32_t shift = 13, base = 0;
  if (exp < -14)
    shift = 13u + (uint32_t)(-14 - exp);                     /* subnormal result: shift further */
  else
    base = (uint32_t)(exp + 15) << 10;
  uint32_t q = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u), mid = 1u << (shift - 1);
  if (rem > mid || (rem == mid && (q & 1u))) q++;
  /* normal: drop the implicit bit; a mantissa carry (q == 0x800) correctly bumps the exponent */
  uint32_t r = (exp < -14) ? q : base + q - 0x400u;
  if (r >= 0x7C00u) r = 0x7C00u;
  return (uint16_t)(sign | r);
}

static float load(const void* p, size_t i, int dt) {
  switch (dt) {
    case 0: return ((const float*)p)[i];
    case 1: return bf16_to_f32(((const uint16_t*)p)[i]);
    default: return f16_to_f32(((const uint16_t*)p)[i]);
  }
}

static float round_to(float x, int dt) { /* value after a cast to dt and back to f32 */
  switch (dt) {
    case 0: return x;
    case 1: return bf16_to_f32(f32_to_bf16(x));
    default: return f16_to_f32(f32_to_f16(x));
  }
}

static void store(void* p, size_t i, int dt, float x) {
  switch (dt) {
    case 0: ((float*)p)[i] = x; break;
    case 1: ((uint16_t*)p)[i] = f32_to_bf16(x); break;
    default: ((uint16_t*)p)[i] = f32_to_f16(x); break;
  }
}

int oracle_allreduce(const void* const* ins, int world, size_t count, int in_dt, int wire_dt,
                     int out_dt, float scale, int post, void* out) {
  if (world < 1 || !ins || !out) return -1;
  const float pre = post ? 1.0f : scale, pst = post ? scale : 1.0f;
  for (size_t i = 0; i < count; ++i) {
    volatile float acc = round_to(load(ins[0], i, in_dt) * pre, wire_dt);
    for (int r = 1; r < world; ++r) {
      volatile float w = round_to(load(ins[r], i, in_dt) * pre, wire_dt);
      acc = acc + w;
    }
    volatile float scaled = acc * pst;
    store(out, i, out_dt, round_to(scaled, wire_dt));
  }
  return 0;
}

/* exposed for the unit tests of the conversions */
uint16_t oracle_f32_to_bf16(float f) { return f32_to_bf16(f); }
uint16_t oracle_f32_to_f16(float f) { return f32_to_f16(f); }
float oracle_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
