#!/bin/bash
# Time outside the frame loop of the persistent OPNet kernels: the same launches at T = 2 and T = 4 (launch + workspace memset +
# weight prologue + epilogue), against T = 300.
for t in 2 4 300; do
echo "== T=$t backward"; TT=$t timeout 100 python tools/split_bwd_debug.py 2>&1 | head -2
echo "== T=$t forward"; TT=$t timeout 100 python tools/split_fwd_debug.py 2>&1 | grep -i "single\|producer\|split" | head -3
done
