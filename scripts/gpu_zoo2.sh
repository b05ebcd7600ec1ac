#!/bin/bash
mkdir -p gpurun_out
python scripts/kernel_zoo2.py > gpurun_out/zoo2_plain.log 2>&1 || { tail -20 gpurun_out/zoo2_plain.log; exit 1; }
N=$(tail -1 gpurun_out/zoo2_plain.log | sed 's/launches [0-9]*\.\.\([0-9]*\):.*/\1/'); N=$((N+1))
ncu --set full --clock-control none --import-source on -s $N -c $N -o gpurun_out/prof_zoo2 python scripts/kernel_zoo2.py > gpurun_out/zoo2_ncu.log 2>&1
cat gpurun_out/zoo2_plain.log; ls -la gpurun_out/prof_zoo2.ncu-rep
