#!/bin/bash
# round-2 validation: full GPU suite, smoke, 1-GPU bench line (with extras) + reference arm, launch list + full ncu capture of the C2 kernel
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log; echo "tests took $(( $(date +%s) - t0 )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
t1=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; echo "bench took $(( $(date +%s) - t1 )) s"
t1=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; echo "ref took $(( $(date +%s) - t1 )) s"
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_launches.csv $CMD > gpurun_out/c2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 1 -o gpurun_out/r02_prof_c2 $CMD > gpurun_out/c2_ncu_full.log 2>&1
tail -1 gpurun_out/c2_ncu_full.log
head -c 1500 gpurun_out/bench_n1.log
# launch lists of the other configs as committed (eager calls: every generated / library kernel of one gradient)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c3_launches.csv python scripts/profile_config.py c3 4 > gpurun_out/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c5_launches.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c4_launches.csv python scripts/profile_config.py c4 2 > gpurun_out/ncu_c4.log 2>&1
for c in c3 c4 c5; do python scripts/summarize_launches.py gpurun_out/r02_${c}_launches.csv 6; done
