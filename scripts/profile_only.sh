#!/bin/bash
# ncu full capture of one kernel only (fast): $1 tag, $2 kernel regex, rest: command
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o gpurun_out/prof_$TAG "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
