#!/bin/bash
# where the shared-memory bank conflicts of jp_glm_tc_kernel come from: source-level ncu capture of one launch at cfg3
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jp_glm_tc_kernel --launch-skip 4 -c 1 \
  -o /tmp/ncu/r4p_tc -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r4p_ncu.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/ncu/r4p_tc.ncu-rep --page source --csv > /tmp/ncu/r4p_tc_source.csv 2>/dev/null
python - <<'PY'
import csv
csv.field_size_limit(10**9)
rows=list(csv.reader(open("/tmp/ncu/r4p_tc_source.csv")))
h=rows[1]; ix={c:i for i,c in enumerate(h)}
out=[]
for k,r in enumerate(rows[2:]):
    def g(c):
        try: return int(r[ix[c]])
        except: return 0
    w=g('L1 Wavefronts Shared'); e=g('L1 Wavefronts Shared Excessive'); cf=g('L1 Conflicts Shared N-Way'); idl=g('L1 Wavefronts Shared Ideal')
    if w or e or cf: out.append((w,e,idl,cf,k,r[ix['Source']][:90], g('Instructions Executed')))
out.sort(reverse=True)
with open("gpurun_out/r4p_tc_shared_by_instruction.txt","w") as f:
    f.write("wavefronts excessive ideal nway line instr_executed source\n")
    for o in out[:60]: f.write("%d %d %d %d %d %d %s\n"%(o[0],o[1],o[2],o[3],o[4],o[6],o[5]))
print(open("gpurun_out/r4p_tc_shared_by_instruction.txt").read()[:6000])
PY
ncu -i /tmp/ncu/r4p_tc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
for a,b in zip(h,v):
    if 'shared' in a or 'bank' in a: print(a,b)
" > gpurun_out/r4p_tc_shared_metrics.txt
