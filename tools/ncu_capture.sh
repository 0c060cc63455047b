#!/bin/bash
# usage: tools/ncu_capture.sh NAME KERNEL_REGEX SKIP COUNT <run_config args...>
# Runs the command plainly first, then under `ncu --set full`, and exports the raw (and source) pages as CSV on the box:
# the .ncu-rep files are too large for gpurun_out's 64 MiB limit and are removed.
NAME=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
python tools/run_config.py "$@" > gpurun_out/plain_$NAME.log 2>&1 || { echo "plain run of $NAME failed"; tail -5 gpurun_out/plain_$NAME.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -o /tmp/$NAME python tools/run_config.py "$@" > gpurun_out/ncu_$NAME.log 2>&1
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv > gpurun_out/${NAME}_source.csv 2>/dev/null
ls -la /tmp/$NAME.ncu-rep gpurun_out/${NAME}_raw.csv gpurun_out/${NAME}_source.csv
