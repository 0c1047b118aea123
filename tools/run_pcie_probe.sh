#!/bin/bash
# run on an N-GPU box: topology + the concurrent host<->device copy ceilings (results under gpurun_out/)
mkdir -p gpurun_out
bash tools/probe_topology.sh > gpurun_out/topology.txt 2>&1
B=tools/micro/pcie_multi.bin
N=$(nvidia-smi -L | wc -l)
ALL=$(seq -s, 0 $((N-1)))
SETS="0"
[ $N -ge 2 ] && SETS="$SETS 1 0,1"
[ $N -ge 4 ] && SETS="$SETS 0,2 0,3 2,3 0,1,2,3"
[ $N -ge 8 ] && SETS="$SETS 4 0,4 4,5 4,5,6,7 0,1,4,5 0,2,4,6 $ALL"
{
timeout 120 $B -k flat $SETS
timeout 120 $B -k flat -p 4 $SETS
timeout 120 $B -k flat -n -2 $SETS
timeout 60 $B -k flat -n 0 $ALL
timeout 60 $B -k flat -n 1 $ALL
timeout 120 $B -k 2d $SETS
timeout 120 $B -k d2h 0 $ALL
timeout 120 $B -k wc 0 $ALL
} > gpurun_out/pcie_multi.jsonl 2>gpurun_out/pcie_multi.err
cat gpurun_out/topology.txt | head -60
cat gpurun_out/pcie_multi.jsonl
tail -5 gpurun_out/pcie_multi.err
