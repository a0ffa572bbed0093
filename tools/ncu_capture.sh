#!/bin/bash
# ncu evidence for profiles/: the launch list of one config-2 step and --set full captures of the top kernels
# (B200_PROFILING.md recipe; run under gpurun on one GPU).  tools/prof_step.py runs three identical steps; the
# captures take a launch of the third.
#   bash tools/ncu_capture.sh [kernel-regex ...]      default: every kernel worth a look
python tools/prof_step.py c2 > gpurun_out/prof_plain.log 2>&1 || exit 1
# the launches of the third step: prof_step prints "<n> launches in 3 steps"
L=$(( $(grep -o '[0-9]* launches in 3 steps' gpurun_out/prof_plain.log | cut -d' ' -f1) / 3 ))
ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/r02_launches_c2.csv python tools/prof_step.py c2 > /dev/null 2>&1
[ "$1" = "launches" ] && exit 0
capture() {  # regex, launches to skip, tag
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$1" -s "$2" -c 1 -o "gpurun_out/r02_ncu_$3" -f python tools/prof_step.py c2 > "gpurun_out/ncu_$3.log" 2>&1
}
capture 'onesweep_kernel<unsigned int, \(int\)256, \(int\)5, \(int\)8' 10 onesweep_u32
capture 'mems::segment_flag_kernel' 2 segment_flag
capture 'mems::run_hits_kernel' 2 run_hits
capture 'mems::hit_describe_kernel' 2 hit_describe
capture 'mems::extract_kernel' 2 extract
capture 'mems::long_walk_right_kernel' 2 long_walk_right
capture 'void mems::walk_right_kernel' 2 walk_right
capture 'void mems::walk_left_kernel' 2 walk_left
ls -la gpurun_out/*.ncu-rep
