#!/bin/bash
# cfg3 (RACER + LSTM) device time per step for library variants
for v in "$@"; do echo -n "$v: "; SMB200_LIB=$PWD/scripts/ab/$v PROF_STEPS=400 python scripts/cfg3_probe.py 2>&1 | grep "^cfg3:" | cut -c1-120; done
