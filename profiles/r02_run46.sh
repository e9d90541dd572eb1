#!/bin/bash
# round 2, GPU call 46: quartet kernel with 32 quartets per warp -- parity (fixtures under both kernels, large batches
# against the oracle), time at B = 2^20 both ways, ncu of the new kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_step.py -x -q -m gpu -k "quartet or fit or step or mds" > gpurun_out/r46_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r46_tests.log)"
for b in 0 1; do
echo "== VVB_QUARTET_BIG=$b"; VVB_QUARTET_BIG=$b timeout -s KILL 120 python profiles/r02c_quartet_probe.py
done
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:"quartet_fwd" -s 4 -c 1 -o gpurun_out/r46_quartet python profiles/r02c_quartet_probe.py > gpurun_out/r46_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r46_quartet.ncu-rep --page raw --csv > gpurun_out/r46_quartet_raw.csv 2>/dev/null
rm -f gpurun_out/r46_quartet.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r46_quartet_raw.csv')))
hdr=rows[0]
def g(r,k):
    return r[hdr.index(k)] if k in hdr else '-'
for r in rows[2:]:
    print(g(r,'Kernel Name')[:44], 'us', g(r,'gpu__time_duration.sum'), 'dramR', g(r,'dram__bytes_read.sum'), 'dramW', g(r,'dram__bytes_write.sum'), 'dram%', g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), 'issue', g(r,'sm__inst_issued.avg.pct_of_peak_sustained_active'), 'occ', g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'), 'regs', g(r,'launch__registers_per_thread'), 'inst', g(r,'smsp__inst_executed.sum'))
PY
