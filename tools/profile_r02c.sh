#!/bin/bash
# DRAM bytes of the fused batched aggregation launches + a full capture of the fused kernel on one C4 pair.
set -u
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
python tools/probe_pair.py c4 -1 > $O/r02b_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_agg_flow -s 4 -c 2 -f -o $O/r02b_agg_flow_fused_c4_single python tools/probe_pair.py c4 -1 > $O/r02b_ncu_c4.log 2>&1
echo "c4 full rc=$?"
$BENCH > $O/r02b_plain_bench2.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_agg_flow -s 6 -c 6 --csv \
    --log-file $O/r02b_agg_flow_fused_c4_batch8_dram.csv $BENCH > $O/r02b_ncu_dram.log 2>&1
echo "dram rc=$?"
ls -la $O | grep r02b_
