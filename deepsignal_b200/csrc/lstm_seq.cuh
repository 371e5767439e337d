// lstm_seq.cuh — interface of the time-persistent BiLSTM kernel (sm_100a only).
#pragma once
#include "ops.cuh"

namespace dsb {

constexpr int LSTM_SEQ_LAYERS = 3;

struct alignas(64) LstmSeqMaps {
  CUtensorMap f;                          // layer-0 feature rows F [KL*cap][FW], box 128 x 64
  CUtensorMap h[2][LSTM_SEQ_LAYERS];      // H[dir][layer] [(KL+1)*cap][256], box 128 x 64 (slot t+1 = output of step t)
  CUtensorMap w2[2][LSTM_SEQ_LAYERS];     // gate weights [1024][Ktot] (columns 4*u+g), 128-row box: half a 256-row block per CTA of a pair
};

struct LstmSeqParams {
  const float* bias[2][LSTM_SEQ_LAYERS];  // [1024] each
  void* H[2][LSTM_SEQ_LAYERS];
  float* c[2];                            // per-direction cell state scratch [cap][256] fp32, tile-interleaved (re-used by every layer)
  long long cap = 0;                      // rows per H slot / F position
  int n = 0;                              // valid samples
  int KL = 0;                             // time steps
};

struct LstmSeqLaunch {
  LstmSeqMaps maps;
  LstmSeqParams p;
  bool fp16 = false;
  double flops_per_sample = 0;
};

// ablate: timing experiments only (dsb_plan_options): 1 no cell-state traffic, 2 no h store
void launch_lstm_seq2(const LstmSeqLaunch& L, int n, int sm_count, int ablate, cudaStream_t s);   // CTA-pair kernel (lstm_seq2.cu)

}  // namespace dsb
