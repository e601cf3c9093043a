#!/bin/bash
# A/B the kernel variants built by tools/build_variant.py on the GPU box: two bench runs each (kernel ms, step ms).
for v in "$@"; do
  for k in 1 2; do
    IQ2A_LIB=_ab/libiq2a_$v.so timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-others > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"
  done
done
