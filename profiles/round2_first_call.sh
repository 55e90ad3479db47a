#!/bin/bash
# First GPU call of round 2 (one B200, ~3 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash profiles/round2_first_call.sh'
# 1. the kernel variants written blind at the end of round 1 meet a GPU (tests gated on DMI_EXPERIMENTAL=1)
# 2. per-mode timings against the kernels they replace + step A/B per fused_panel bit (profiles/panel_tc_modes_probe.py)
# 3. ncu --set full of the default tcgen05 dpre pass and its merged-column-sum variant (why 43.6 us and not ~25 us?)
# 4. launch list of one bench step with the round-1 default schedule
# Every step has its own timeout and writes under gpurun_out/; a failing step does not stop the others.
set -u
mkdir -p gpurun_out
export DMI_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_panel_gpu.py -q > gpurun_out/r2_experimental_tests.log 2>&1; echo "experimental tests rc=$?"
tail -15 gpurun_out/r2_experimental_tests.log
timeout 90 python profiles/panel_tc_modes_probe.py > gpurun_out/r2_panel_modes.log 2>&1; echo "modes probe rc=$?"
tail -45 gpurun_out/r2_panel_modes.log
timeout 150 ncu --set full --import-source on --clock-control none -k regex:panel_tc_kernel -c 2 -f -o gpurun_out/r2_panel_tc \
    python profiles/panel_tc_ncu_probe.py > gpurun_out/r2_panel_tc_ncu.log 2>&1; echo "ncu panel_tc rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_step_ncu.log 2>&1; echo "launch list rc=$?"
