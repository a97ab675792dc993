#!/bin/bash
# e2e of config 2 over sub-batch sizes of the pipelined submit
for n in 200000 262144 330000 400000 520000; do for f in 65536 98304 160000; do
  echo -n "PIPE_SEQS $n FIRST $f: "; PF_PIPELINE_SEQS=$n PF_PIPELINE_FIRST=$f python tools/e2e_pipe.py 4000 5 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print(' '.join('%.1f+%.1f=%.1f(subs %d)' % (x['submit_ms'], x['collect_ms'], x['submit_ms']+x['collect_ms'], x['subs']) for x in r))"
done; done
