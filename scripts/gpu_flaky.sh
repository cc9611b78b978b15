#!/bin/bash
# repeat a golden CLI case under different debug knobs and count wrong outputs
cd tests/golden/cli/${CASE:-fasta_best}
ARGS=$(python -c "import json; print(' '.join('/tmp/out.b6' if a=='OUT' else a for a in json.load(open('case.json'))['args']))")
sort expected.b6 > /tmp/want.b6
for MODE in ${MODES:-"BURST_B200_DBG=0" "BURST_B200_DBG=1" "BURST_B200_DBG=2" "BURST_B200_DBG=3" "BURST_B200_SEED_CHUNK=1"}; do
  bad=0
  for i in $(seq 1 ${REPS:-12}); do
    env $MODE ../../../../burst_b200/host/burst-b200 $ARGS --noprogress > /tmp/run.log 2>&1 || { bad=$((bad+1)); tail -2 /tmp/run.log; continue; }
    sort /tmp/out.b6 | cmp -s - /tmp/want.b6 || bad=$((bad+1))
  done
  echo "$MODE: $bad bad of ${REPS:-12}"
done
