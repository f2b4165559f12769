#!/bin/bash
O=gpurun_out/${1:-r2q}; mkdir -p $O
HPV_STAMP_OUT=$O HPV_LIB=$PWD/tools/variants/libhpv_stamps.so timeout 300 python tools/fwd_stamps.py > $O/fwd_stamps.txt 2>&1; grep -v "^     \|wait for\|exit\|CTA start" $O/fwd_stamps.txt
tools/gpu_r2m.sh $1
