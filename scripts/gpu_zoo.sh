#!/bin/bash
mkdir -p gpurun_out
python scripts/kernel_zoo.py > gpurun_out/zoo_plain.log 2>&1 || { tail -20 gpurun_out/zoo_plain.log; exit 1; }
SKIP=$(grep -c . /dev/null); 
# the first pass (warm-up) launches the same number of kernels as the measured pass: skip them
N=$(tail -1 gpurun_out/zoo_plain.log | sed 's/launches [0-9]*\.\.\([0-9]*\):.*/\1/'); N=$((N+1))
ncu --set full --clock-control none --import-source on -s $N -c $N -o gpurun_out/prof_zoo python scripts/kernel_zoo.py > gpurun_out/zoo_ncu.log 2>&1
cat gpurun_out/zoo_plain.log; ls -la gpurun_out/prof_zoo.ncu-rep
