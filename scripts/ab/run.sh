#!/bin/bash
# A/B of library builds on one box: alternate the variants, three rounds
for r in 1 2 3; do for v in "$@"; do
  echo -n "$v: "; SMB200_LIB=$PWD/scripts/ab/$v python bench.py --no-cpu-baseline --steps 20000 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
