#!/bin/bash
# Root-cause demonstration: the fused backward with its landing slots deliberately filled with "ready-looking" words
# (what another kernel's leftovers can be) must fail the [11-37] parity case the way round 1 saw it; the shipped library passes.
mkdir -p gpurun_out
SEQ='tests/test_gpu_kernels.py::test_opnet_fused_forward_matches_separate_kernels'
echo "== poisoned landing slots (libopnet_b200_poison.so)" > gpurun_out/r02_03_rootcause.log
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_poison.so timeout 300 python -m pytest "$SEQ" -m gpu -q --tb=line -p no:cacheprovider -k "11-37 or 8-300 or 32-64" >> gpurun_out/r02_03_rootcause.log 2>&1
echo "== shipped library" >> gpurun_out/r02_03_rootcause.log
timeout 300 python -m pytest "$SEQ" -m gpu -q --tb=line -p no:cacheprovider -k "11-37 or 8-300 or 32-64" >> gpurun_out/r02_03_rootcause.log 2>&1
tail -30 gpurun_out/r02_03_rootcause.log
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_03_suite.log 2>&1; tail -5 gpurun_out/r02_03_suite.log
