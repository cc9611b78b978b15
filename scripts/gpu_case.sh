#!/bin/bash
# one golden CLI case under several engine settings: rows and the rows missing against the expected file
CASE=${CASE:-edx_quick_forage}
cd tests/golden/cli/$CASE
ARGS=$(python -c "import json; print(' '.join('/tmp/out.b6' if a=='OUT' else a for a in json.load(open('case.json'))['args']))")
sort expected.b6 > /tmp/want.b6
for env in "BURST_B200_SEED_IMPL=1" "BURST_B200_SEED_IMPL=0" "BURST_B200_SEED_IMPL=1 BURST_B200_SEED_NCH=4" "BURST_B200_SEED_IMPL=1 BURST_B200_SEED_FB=2" "BURST_B200_SEED_FILTER=0"; do
  env $env ../../../../burst_b200/host/burst-b200 $ARGS --noprogress > /tmp/run.log 2>&1; rc=$?
  sort /tmp/out.b6 > /tmp/got.b6
  echo "[$env] rc=$rc rows=$(wc -l < /tmp/out.b6) missing=$(comm -13 /tmp/got.b6 /tmp/want.b6 | wc -l) extra=$(comm -23 /tmp/got.b6 /tmp/want.b6 | wc -l)"
  comm -13 /tmp/got.b6 /tmp/want.b6 | head -4
done
