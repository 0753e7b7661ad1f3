#!/bin/sh
# Round-2 evidence run (one B200): every output goes to gpurun_out/, from where
# tools/refresh_profiles_r2.py makes the tracked summaries under profiles/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'sh tools/profile_r2.sh'
set -x
O=gpurun_out
Q="--lines 20000000 --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs"
# 1. the bench as the driver runs it
timeout -s KILL 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err || exit 1
# 2. all styles, both kernels, and the geometry knobs
timeout -s KILL 300 python tools/hot_sweep.py --lines 20000000 "HOT_DEBUG=1" "" HOT=0 HOT_ASCII=-1 HOT_DEFER=-1 \
    HOT_PRODUCERS=2 HOT_PRODUCERS=4 HOT_WARPS=24 HOT_BATCH=128 > $O/r2_sweep_instant.txt 2>&1
for sty in last full; do
  timeout -s KILL 200 python tools/hot_sweep.py --lines 20000000 --style $sty "" HOT=0 HOT_FORCE=1 > $O/r2_sweep_$sty.txt 2>&1
done
timeout -s KILL 200 python tools/hot_sweep.py --lines 20000000 --program single_auto "" HOT=0 > $O/r2_sweep_single.txt 2>&1
# 3. the same short command first without ncu (must exit 0), then the launch list and one full capture
timeout -s KILL 300 python bench.py $Q > $O/r2_short.json 2> $O/r2_short.err || exit 1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py $Q > $O/r2_ncu_list.log 2>&1
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:batch_hot --launch-skip 3 -c 1 \
    -o $O/r2_hot python bench.py $Q > $O/r2_ncu_full.log 2>&1
# 4. the long stream: launch list of one 4 GB call per style
timeout -s KILL 300 python tools/profile_stream.py > $O/r2_stream.txt 2>&1
echo "--- RED_B200_STREAM_HOT=0 (the walk on the full table, chunk engine)" >> $O/r2_stream.txt
RED_B200_STREAM_HOT=0 timeout -s KILL 300 python tools/profile_stream.py 2>&1 | grep union64 >> $O/r2_stream.txt
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $O/r2_stream_launches.csv python tools/profile_stream.py --once > $O/r2_ncu_stream.log 2>&1
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:stream_hot -c 1 \
    -o $O/r2_stream_hot python tools/profile_stream.py --once > $O/r2_ncu_stream_full.log 2>&1
ls -la $O
