// incept_tc.cuh — interface of the two fused inception-block kernels (tcgen05 / TMA, sm_100a).
//
// One inception block (reference layers.py:87-139) = kernel A + kernel B:
//   A  x -> { b1 = relu(conv1x1(maxpool3(x))), b2 = relu(conv1x1(x)) } -> Y[:, 0:96]
//          { b3a | b4a | b5a = relu(conv1x1(x)), stem = conv1x1(x) }   -> planes P3, P4, P5 [M][32], PS [M][48]
//      the 1x3 max pool is computed in shared memory from the x tile (in place, after the p1 MMAs consumed it)
//   B  T -> b3 = relu(conv1x3(b3a)), b4 = relu(conv1x5(b4a)), b5 = relu(stem + conv1x1(relu(conv1x3(b5a)))) -> Y[:, 96:240]
//      the intermediate [128 x 64] tile of branch 5 never leaves the SM (TMEM -> registers -> smem -> MMA)
// HBM traffic per row: A 480 B in + 480 B out, B 288 B in + 288 B out (16-bit), against 4.3 KB for the seven
// separate launches they replace.
#pragma once
#include "ops.cuh"

namespace dsb {

// Split-operand mode (DSB_PREC_FP16X2): every activation matrix has a lo twin (v = hi + lo, both fp16); the *l maps address
// the twins.  The weights stay single fp16: [xh | xl] · [W ; W] re-uses each weight block for both parts.
struct alignas(64) InceptAMaps {
  CUtensorMap x, wp1, wb1, ys, p3s, p4s, p5s, pss;      // *s: [32 x 32] store boxes over Y and the planes
  CUtensorMap xl, ysl, p3sl, p4sl, p5sl, pssl;          // lo twins (split mode only)
};
struct InceptAParams {
  int M = 0, Lp = 0, L = 0;
  int nkb = 4;                 // 64-wide K blocks of the input channels (256 or 240 -> 4)
  void* Y = nullptr; int ldY = 0;
  void* P3 = nullptr; void* P4 = nullptr; void* P5 = nullptr;   // [M][32] planes: b3a, b4a, b5a
  void* PS = nullptr;                                           // [M][48] plane: residual stem
  int ablate = 0;              // timing experiments only (DSB_ABLATE_A): 1 skip pool math, 2 skip epilogue stores
  const float* bias = nullptr; // [240]: b1 (48) then p1 (b2 48 | b3a 32 | b4a 32 | b5a 32 | stem 48)
  float biasv[240];            // the same values by value: read through the constant bank (LDC), not shared memory
  long long lo_y = 0, lo_p3 = 0, lo_p4 = 0, lo_p5 = 0, lo_ps = 0;   // element offsets hi -> lo twin (split mode)
};
struct InceptALaunch { InceptAMaps maps; InceptAParams p; bool fp16 = false; bool split = false; };

struct alignas(64) InceptBMaps {
  CUtensorMap t3, t4, t5, w3, w4, w5b, w5c, ys;         // ys: Y as [32 x 32] store boxes
  CUtensorMap t3l, t4l, t5l, ysl;                       // lo twins (split mode only)
};
struct InceptBParams {
  int M = 0, Lp = 0, L = 0;
  const void* PS = nullptr;    // [M][48] residual stem plane
  void* Y = nullptr; int ldY = 0;
  int ablate = 0;              // timing experiments only (DSB_ABLATE_B): 1 skip staging math, 2 skip epilogue stores
  const float* bias = nullptr; // [208]: b3 48 | b4 48 | b5c 48 | b5b 64
  float biasv[208];            // by value: constant-bank reads keep the bias off the shared-memory data pipe
  long long lo_ps = 0;         // element offset PS -> its lo twin (split mode)
};
struct InceptBLaunch { InceptBMaps maps; InceptBParams p; bool fp16 = false; bool split = false; };

constexpr int INCEPT_A_TILE_ROWS = 126;   // 128-row tiles overlap by one row on each side (pool halo)

// ablate: timing experiments only (dsb_plan_options): 1 skip pool / staging math, 2 skip epilogue stores, 4 (kernel A) skip the whole epilogue,
// 8 (kernel A) no b1 MMAs, 16 (kernel A) no lo-part p1 MMAs
void launch_incept_a(const InceptALaunch& L, int M, int sm_count, int ablate, cudaStream_t s);
void launch_incept_b(const InceptBLaunch& L, int M, int sm_count, int ablate, cudaStream_t s);

}  // namespace dsb
