#!/bin/bash
# memcheck + racecheck of the drop-in binary on one golden CLI case (CASE) and of a python test selection (PYTEST_K)
CASE=${CASE:-fasta_allpaths_whitespace}
cd tests/golden/cli/$CASE
ARGS=$(python -c "import json; print(' '.join('/tmp/out.b6' if a=='OUT' else a for a in json.load(open('case.json'))['args']))")
for i in 1 2 3; do ../../../../burst_b200/host/burst-b200 $ARGS --noprogress > /tmp/run.log 2>&1; echo "rc=$? rows=$(wc -l < /tmp/out.b6)"; done
timeout 600 compute-sanitizer --tool memcheck ../../../../burst_b200/host/burst-b200 $ARGS --noprogress 2>&1 | grep -v "^$" | head -40
timeout 600 compute-sanitizer --tool racecheck ../../../../burst_b200/host/burst-b200 $ARGS --noprogress 2>&1 | grep -v "^$" | head -30
