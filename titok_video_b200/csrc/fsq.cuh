// Finite scalar quantization arithmetic, shared by the standalone FSQ kernels and the fused encoder head.
//
// Mirrors model/quantizer/fsq.py op by op in fp32 (the reference disables autocast and calls z.float()):
//   bound            fsq.py:78-83    tanh(z + shift) * half_l - offset
//   round_ste        fsq.py:48-51    round-half-even
//   quantize         fsq.py:85-90    / half_width
//   codes_to_indices fsq.py:105-109  sum((codes*half_width + half_width) * basis) -> int32
// The per-dimension constants (half_l, offset, shift, half_width, basis) are computed by the host with
// torch exactly as the reference computes them, so they are bit-identical; mul/add are kept as separate
// roundings (no FMA contraction) because the reference executes them as separate fp32 kernels.
#pragma once

#include "common.cuh"

namespace ttk {

constexpr int FSQ_MAX_D = 8;

struct FsqConsts {
  int D;
  float half_l[FSQ_MAX_D];
  float offset[FSQ_MAX_D];
  float shift[FSQ_MAX_D];
  float half_width[FSQ_MAX_D];
  float rcp_half_width[FSQ_MAX_D];  // RN(1 / half_width), filled by the launchers
  float basis[FSQ_MAX_D];
  int levels[FSQ_MAX_D];
  int ibasis[FSQ_MAX_D];
};

// q / hw, correctly rounded, without the IEEE-divide slow path: one Newton correction on q * RN(1/hw) (Markstein).
// Verified exhaustively against fp32 division for every integer |q| <= hw + 2, hw = 1..16 (tests/test_host_logic.py).
__device__ __forceinline__ float fsq_div_hw(float q, float hw, float rcp_hw) {
  const float r0 = __fmul_rn(q, rcp_hw);
  const float e = __fmaf_rn(-hw, r0, q);
  return __fmaf_rn(e, rcp_hw, r0);
}

// returns the fp32 code; accumulates the (exact, integer-valued) fp32 index term
__device__ __forceinline__ float fsq_quantize_dim(float z, const FsqConsts& c, int d, float& idx_acc) {
  const float b = __fsub_rn(__fmul_rn(tanhf(__fadd_rn(z, c.shift[d])), c.half_l[d]), c.offset[d]);
  // round-half-even of |b| < 2^22 through the 1.5 * 2^23 magic constant (rintf is a conversion-pipe instruction)
  const float q = __fsub_rn(__fadd_rn(b, 12582912.0f), 12582912.0f);
  const float code = fsq_div_hw(q, c.half_width[d], c.rcp_half_width[d]);
  const float lvl = __fadd_rn(__fmul_rn(code, c.half_width[d]), c.half_width[d]);
  idx_acc = __fadd_rn(idx_acc, __fmul_rn(lvl, c.basis[d]));
  return code;
}

}  // namespace ttk
