#!/bin/bash
# round 2, GPU call 45: quartet kernel at B = 2^20 (CUDA events, launch durations, ncu --set full); operand build with
# the 8-CTAs-per-SM grid on the query side
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 120 python profiles/r02c_quartet_probe.py
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:"quartet_fwd|scale_by" -s 4 -c 2 -o gpurun_out/r45_quartet python profiles/r02c_quartet_probe.py > gpurun_out/r45_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r45_quartet.ncu-rep --page raw --csv > gpurun_out/r45_quartet_raw.csv 2>/dev/null
ncu -i gpurun_out/r45_quartet.ncu-rep --page source --csv -k regex:quartet_fwd > gpurun_out/r45_quartet_source.csv 2>/dev/null
ncu -i gpurun_out/r45_quartet.ncu-rep --page details > gpurun_out/r45_quartet_details.txt 2>/dev/null
rm -f gpurun_out/r45_quartet.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r45_quartet_raw.csv')))
hdr=rows[0]
def g(r,k):
    return r[hdr.index(k)] if k in hdr else '-'
for r in rows[2:]:
    print(g(r,'Kernel Name')[:44], 'us', g(r,'gpu__time_duration.sum'), 'dramR', g(r,'dram__bytes_read.sum'), 'dramW', g(r,'dram__bytes_write.sum'), 'issue', g(r,'sm__inst_issued.avg.pct_of_peak_sustained_active'), 'occ', g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'), 'regs', g(r,'launch__registers_per_thread'), 'inst', g(r,'smsp__inst_executed.sum'))
PY
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"build_operands" --csv --log-file gpurun_out/r45_launches.csv python profiles/run_screen.py 1000000 > gpurun_out/r45_ncu2.log 2>&1
grep "build_" gpurun_out/r45_launches.csv | tail -2 | awk -F, '{print $5, $NF}' | cut -c1-100
# the wide-row first tier (C5-shaped, four waves against 500k x 2000) under the three cell-plan variants: did the
# TF32 assignment change the tiles swept / time of the 2M x 2000 configuration?
for a in 2 0; do
echo "== VVB_CELLS_ASSIGN=$a"
VVB_CELLS_ASSIGN=$a NQ_BLOCKS=592 timeout -s KILL 300 python profiles/r02_prof_wide.py 2>&1 | tail -1 | cut -c1-420
done
