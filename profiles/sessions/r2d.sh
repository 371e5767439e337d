#!/usr/bin/env bash
# GPU session r2d: LSTM / CNN SM partitioning (dsb_plan_options.lstm_sms) in the split and single modes
set -u
O=gpurun_out
: > $O/r2d_lstm_sms.jsonl
for prec in fp16x2 fp16; do
  for k in 0 12 16 20 24 32 40; do
    timeout 120 python profiles/prof_forward.py --precision $prec --chunks 4 --opt lstm_sms=$k >> $O/r2d_lstm_sms.jsonl 2>&1
  done
done
cut -c1-200 $O/r2d_lstm_sms.jsonl
