#!/bin/bash
# Multi-GPU forms of the drop-in binary on the golden accelerator cases (needs >= 2 GPUs: gpurun --gpus 2):
#   --gpus N               query batches dealt to N GPUs, database replicated, per-read minima merged on the host
#   --gpus N --shard-refs  database cut into N clump ranges, ncclAllReduce(MIN) on the per-read minima between extend and select
# every output must equal the file the reference binary wrote (sorted).
N=${NGPU:-2}
fail=0
for CASE in acx_best acx_allpaths_fr acx_capitalist_tax_iupac acx_forage_mixed_lengths acx_y_wildcard; do
  d=tests/golden/cli/$CASE
  gunzip -c $d/db.acx.gz > /tmp/db.acx
  ARGS=$(python -c "import json; print(' '.join({'OUT':'/tmp/out.b6','db.acx':'/tmp/db.acx'}.get(a,a) for a in json.load(open('$d/case.json'))['args']))")
  sort $d/expected.b6 > /tmp/want.b6
  for mode in "--gpus $N" "--gpus $N --shard-refs -sa" "--gpus $N -t 8"; do
    # (FORAGE rows depend on the bunch size through the reference's first-seen-wins duplicate suppression, SURVEY.md App. A: -t changes them in the reference too)
    if [ "$CASE" = acx_forage_mixed_lengths ] && [ "$mode" = "--gpus $N -t 8" ]; then continue; fi
    (cd $d && ../../../../burst_b200/host/burst-b200 $ARGS --noprogress $mode > /tmp/run.log 2>&1); rc=$?
    sort /tmp/out.b6 > /tmp/got.b6
    if [ $rc -eq 0 ] && cmp -s /tmp/got.b6 /tmp/want.b6; then echo "ok   $CASE [$mode] rows=$(wc -l < /tmp/out.b6)"; else echo "FAIL $CASE [$mode] rc=$rc rows=$(wc -l < /tmp/out.b6) want=$(wc -l < /tmp/want.b6)"; tail -3 /tmp/run.log; fail=1; fi
  done
done
grep -h "shard\|Accel\]" /tmp/run.log | tail -4
exit $fail
