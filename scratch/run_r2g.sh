#!/bin/bash
for c in up4; do echo "== $c dual"; timeout 120 python scratch/trace_stream.py $c 2>&1 | tail -30; echo "== $c single"; TBI_TC_NO_DUAL=1 timeout 120 python scratch/trace_stream.py $c 2>&1 | tail -24; done
