#!/bin/bash
# Freeze what a GPU probe job needs into variants/<tag>/ so that edits made while the job waits in gpurun's queue cannot
# reach it (gpurun snapshots the tree when the job STARTS).  Usage: tools/freeze_job.sh <tag>; the job then runs
#   cd variants/<tag> && DSAC_SKIP_BUILD=1 python tools/...
set -e
T=variants/$1
rm -rf $T; mkdir -p $T/dsac_b200 $T/tools $T/oracle $T/tests
cp dsac_b200/*.py dsac_b200/*.so $T/dsac_b200/
cp tools/*.py $T/tools/
cp oracle/*.py oracle/*.so $T/oracle/ 2>/dev/null || true
cp -r tests/*.py tests/golden $T/tests/ 2>/dev/null || true
cp bench.py MEASURED_PEAKS.json $T/ 2>/dev/null || true
echo frozen $T
