run() { ncu --metrics gpu__time_duration.sum --clock-control none -k regex:incept -s 44 -c 22 --csv python bench.py --steps 1 --warmup 3 --samples-per-step 32768 --max-chunk 32768 --e2e-samples 16384 --no-cpu-baseline 2>/dev/null | grep '^"[0-9]' | grep incept | python -c "
import sys,csv
a=b=0
for r in csv.reader(sys.stdin):
    t=float(r[-1].replace(',',''))/1e3
    if 'incept_a' in r[4]: a+=t
    else: b+=t
print('A total %.0f us  B total %.0f us'%(a,b))"; }
echo baseline; run
echo ABLATE_A=1; DSB_ABLATE_A=1 run
echo ABLATE_A=2; DSB_ABLATE_A=2 run
echo ABLATE_A=3; DSB_ABLATE_A=3 run
echo ABLATE_B=2; DSB_ABLATE_B=2 run
