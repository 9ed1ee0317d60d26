#!/bin/bash
# tools/ncu_capture.sh TAG: on the GPU box -- the bench command plainly first, then its launch list
# (gpu__time_duration, no clock control) and one `--set full` capture of K1 / K2b / K3 / K4 of one 128-frame launch chain.
# Outputs: gpurun_out/TAG_plain.log, TAG_launches.csv, TAG_full.ncu-rep (summarise here with tools/ncu_summary.py / ncu_phases.py)
tag=${1:-cap}
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --no-verify --steps 1 --warmup 3 --images 256 --sub-batch 128 --exact-sub-batch --depth 1"
mkdir -p gpurun_out
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
K='regex:k1_|k2_|k2b_|k3_|k4_|k5_|k_zero'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k1_transform_p420|k2b_tables|k3_pack_tiles|k4_stuff' -s 4 -c 4 -f -o gpurun_out/${tag}_full $B > gpurun_out/${tag}_ncu_f.log 2>&1
tail -2 gpurun_out/${tag}_ncu_f.log; ls -la gpurun_out/ | grep ${tag}
