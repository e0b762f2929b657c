#!/bin/bash
# Round-2 profiling captures (run under gpurun on ONE B200; outputs into gpurun_out/, summaries are copied into profiles/).
# Every ncu command is preceded by a plain run of the same command line that has to exit 0 (B200_PROFILING.md).
set -u
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
# 1. launch list of the bench command (cold-cache, serialised: shares, not absolutes)
$BENCH > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_bench_c4_batch8.csv $BENCH > $O/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
# 2. DRAM bytes of the batched aggregation launches of the same command (one pass: no replay of the 100 GB working set)
$BENCH > $O/r02_plain_bench2.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_agg_flow -s 6 -c 6 --csv \
    --log-file $O/r02_agg_flow_c4_batch8_dram.csv $BENCH > $O/r02_ncu_dram.log 2>&1
echo "dram rc=$?"
# 3. full capture of the dataflow kernel on one C2 pair (both views in one launch) and of the cluster walk on the FLIR pair
python tools/probe_pair.py c2 -1 > $O/r02_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_agg_flow -s 2 -c 1 -f -o $O/r02_agg_flow_c2_single python tools/probe_pair.py c2 -1 > $O/r02_ncu_c2.log 2>&1
echo "c2 full rc=$?"
python tools/probe_pair.py flir 0 > $O/r02_plain_flir.log 2>&1 &&
# (the cluster launch is the first k_agg_flow launch of every run_dense: launches 0, 3, 6, ... of the filter)
ncu --set full --clock-control none --import-source on -k regex:k_agg_flow -s 6 -c 1 -f -o $O/r02_agg_flow_cluster_flir python tools/probe_pair.py flir 0 > $O/r02_ncu_flir.log 2>&1
echo "flir cluster full rc=$?"
ls -la $O | grep r02_
