// front_tc.cuh — interface of the fused CNN front end (tcgen05 / TMA, sm_100a):
//   signals f32[n][S] -> conv1 1x7 s2 + BN + ReLU -> maxpool 1x3 s2   (reference layers.py:142-163, CUDA cores, fp32)
//                     -> conv2 1x1 64->128 + BN + ReLU                 (tcgen05, K = 64)
//                     -> conv3 1x3 128->256 + BN + ReLU                (tcgen05, three row-shifted taps, K = 384)
//                     -> X0 [n*Lp][256] 16-bit (input of the first inception block)
// The three separate launches wrote and re-read 1.1 KB per row through HBM; here P1 and C2 never leave the SM.
#pragma once
#include "ops.cuh"

namespace dsb {

struct alignas(64) FrontMaps {
  CUtensorMap w2;     // conv2 weights [128][64], box 128 x 64
  CUtensorMap w3;     // conv3 weights [256][384] (K = tap*128 + channel), box 128 x 32
  CUtensorMap w1;     // front2: conv1 (hi | hi | lo) parts [192][64], box 192 x 64
};

struct FrontParams {
  const float* signals = nullptr;   // [n][S]
  int S = 0;
  const float* w1 = nullptr;        // [64][7] folded conv1
  const float* b1 = nullptr;        // [64]
  const float* b2 = nullptr;        // [128]
  const float* b3 = nullptr;        // [256]
  int L1 = 0, L = 0, Lp = 0;        // conv1 output length, pooled length (90), padded rows per sample (92)
  int conv_pl = 0, pool_pl = 0;     // SAME-padding left pads
  void* X0 = nullptr; int ldX = 0;  // [n*Lp][ldX]
  long long lo_x0 = 0;              // element offset X0 -> its lo twin (split mode)
  int M = 0;                        // n * Lp
  float b1v[64]; float b2v[128]; float b3v[256];   // front2: biases by value (constant bank)
};

struct FrontLaunch { FrontMaps maps; FrontParams p; bool fp16 = false; bool split = false; };

constexpr int FRONT_TILE_ROWS = 126;   // 128-row tiles overlap by one row on each side (conv3 halo)

void launch_front2(const FrontLaunch& L, const float* signals, int n, int sm_count, cudaStream_t s);   // conv1 on the tensor cores too (front2_tc.cu)

}  // namespace dsb
