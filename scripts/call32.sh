#!/bin/bash
mkdir -p gpurun_out
BURST_B200_DEBUG=1 timeout 1500 python scripts/manuscript_fixture.py --every 3 --skip-reference > gpurun_out/manuscript_debug.json 2> gpurun_out/manuscript_debug.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/manuscript_debug.json'))
for l in d['ours_stderr_tail']: print(l[:400])
print(d['ours_stdout_tail'][-3:])
PY
