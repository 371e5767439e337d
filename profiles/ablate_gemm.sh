run() { ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 208 -c 104 --csv python bench.py --steps 1 --warmup 3 --samples-per-step 32768 --max-chunk 32768 --e2e-samples 16384 --no-cpu-baseline 2>/dev/null | grep '^"[0-9]' | grep gemm_tc | python -c "
import sys,csv
a=b=0
for r in csv.reader(sys.stdin):
    t=float(r[-1].replace(',',''))/1e3
    if '<64, 1' in r[4]: a+=t
    else: b+=t
print('LSTM gemm total %.0f us  conv gemm total %.0f us'%(a,b))"; }
echo baseline; run
echo ABLATE_G=2 "(epilogue skips c/h/output global traffic)"; DSB_ABLATE_G=2 run
