#!/bin/bash
# round 2, GPU call 49: ncu --set full of the fused training-step kernel (batch 512: 16 CTAs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:"vae_step_kernel" -s 20 -c 1 -o gpurun_out/r49_step python profiles/r02_fit_probe.py > gpurun_out/r49_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r49_step.ncu-rep --page raw --csv > gpurun_out/r49_step_raw.csv 2>/dev/null
ncu -i gpurun_out/r49_step.ncu-rep --page source --csv > gpurun_out/r49_step_source.csv 2>/dev/null
ncu -i gpurun_out/r49_step.ncu-rep --page details > gpurun_out/r49_step_details.txt 2>/dev/null
rm -f gpurun_out/r49_step.ncu-rep
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Warp Cycles Per Issued|One or More Eligible|Registers Per|Achieved Occupancy|Shared Memory Configuration|Dynamic Shared" gpurun_out/r49_step_details.txt | head -12
