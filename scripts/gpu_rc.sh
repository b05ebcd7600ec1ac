#!/bin/bash
# run a pytest selection and report the REAL exit code + stderr tail (catches crashes at interpreter exit)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "$1" > gpurun_out/rc_out.log 2> gpurun_out/rc_err.log; echo "exit code $?"
tail -3 gpurun_out/rc_out.log; head -30 gpurun_out/rc_err.log
