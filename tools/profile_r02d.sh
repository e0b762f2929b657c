#!/bin/bash
# Final captures of round 2 (work-proportional CTA classes, fused matching cost).  Run under gpurun on ONE B200.
set -u
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
$BENCH > $O/r02d_plain_bench2.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_agg_flow -s 6 -c 6 --csv \
    --log-file $O/r02d_agg_flow_fused_c4_batch8_dram.csv $BENCH > $O/r02d_ncu_dram.log 2>&1
echo "dram rc=$?"
python tools/probe_pair.py c4 -1 > $O/r02d_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_agg_flow -s 4 -c 2 -f -o $O/r02d_agg_flow_fused_c4_single python tools/probe_pair.py c4 -1 > $O/r02d_ncu_c4.log 2>&1
echo "c4 full rc=$?"
$BENCH > $O/r02d_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02d_launches_bench_c4_batch8.csv $BENCH > $O/r02d_ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la $O | grep r02d_
