#!/bin/bash
cd /root/repo
for c in "$@"; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
