#!/bin/bash
# Sustained (power-capped) kernel-only throughput of the CROP kernel vs rows per work unit / CTAs per SM:
# bash tools/power_sweep.sh <config> <frames per step> "<rows> <ctas>" ...   (bench.py reads D2PC_ROWS_PER_UNIT / D2PC_CTAS_PER_SM)
cfg=$1; frames=$2; shift 2
for spec in "$@"; do set -- $spec; D2PC_ROWS_PER_UNIT=$1 D2PC_CTAS_PER_SM=$2 python bench.py --config $cfg --frames $frames --no-cpu --no-ceiling 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('rows $1 ctas $2 config',d['config']['config'],'frames/step',d['config']['units_per_step_per_gpu'],round(d['value']),round(d['roofline']['frac'],4),round(d['roofline']['launch_us'],1),d['clocks']['sm_mhz'],d['clocks']['reasons'])"; done
