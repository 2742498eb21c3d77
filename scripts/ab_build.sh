#!/bin/bash
# Development helper: build an A/B variant of the library next to the product one (fast build: a handful of window
# sizes) -- scripts/ab_build.sh <tag> "<extra nvcc flags>"  ->  bio_b200/lib/ab/libb200sketch_<tag>.so
# Select it with B200SK_LIB_PATH=... (bio_b200/_cabi.py).
set -e
TAG=$1; shift
cd "$(dirname "$0")/../bio_b200/csrc"
mkdir -p ../lib/ab
make -s OBJDIR=../lib/ab/obj_$TAG OUT=../lib/ab/libb200sketch_$TAG.so PTXLOG=../lib/ab/ptxas_$TAG.log EXTRA="-DB200SK_FAST_BUILD $*"
ls -la ../lib/ab/libb200sketch_$TAG.so
