#!/bin/bash
# usage (on the GPU box): tools/slots_sweep.sh "2 3"  -- per-kernel ms vs the input-ring depth of the conv_tc kernels
for v in $1; do echo "AASIST_TC_SLOTS=$v"; AASIST_TC_SLOTS=$v tools/variants_bench.sh product --steps 10; done
