#!/bin/bash
# tools/quick.sh <test> <real> [size]  -> one line: GLUP/s and roofline fraction (no suite, no e2e, no cpu), SM clock under load
python bench.py --test $1 --real $2 --size ${3:-512x256x256} --suite none --no-e2e --no-cpu --steps ${STEPS:-10} --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d.get('clocks') or {}; print(d['config']['workload'], round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'us/sweep', round(d['roofline']['avg_launch_us'],1), 'sm_mhz', c.get('sm_mhz'), c.get('reasons'))"
