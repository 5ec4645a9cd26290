#!/bin/bash
# Final evidence of a round, in one GPU call (each capture only after the same command has exited 0 without ncu):
#   ncu --set full of the dominant kernel of configs 2 / 4, launch lists of configs 2 / 5.  Outputs under gpurun_out/.
set -u
O=gpurun_out
B2="python bench.py --steps 1 --warmup 1 --no-postorder --no-extra --no-cpu --no-e2e"
$B2 > $O/cap_c2_plain.json 2> $O/cap_c2_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k regex:do_linear_multi -c 1 -o $O/r02_final_c2_multi $B2 > $O/cap_c2_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches_config2.csv $B2 > $O/cap_c2_launches.log 2>&1
B4="python bench.py --workload config4 --pairs 100000 --steps 2 --warmup 3 --no-cpu --no-e2e"
$B4 > $O/cap_c4_plain.json 2> $O/cap_c4_plain.err
ncu --set full --import-source on --clock-control none -k regex:do_explicit_kernel -c 1 -o $O/r02_final_c4_explicit $B4 > $O/cap_c4_ncu.log 2>&1
B5="python bench.py --workload config5 --po-trees 256 --po-steps 1 --no-cpu --no-fresh-trees"
$B5 > $O/cap_c5_plain.json 2> $O/cap_c5_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_final_launches_config5_256trees.csv $B5 > $O/cap_c5_launches.log 2>&1
echo captured
