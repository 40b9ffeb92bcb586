#!/bin/bash
# compute-sanitizer over the hand-written kernels; logs under gpurun_out/sanitizer/ (summaries are copied to profiles/).
# usage: bash tools/run_sanitizer.sh [tools...]   (default: memcheck racecheck synccheck)
mkdir -p gpurun_out/sanitizer
TOOLS=${@:-"memcheck racecheck synccheck"}
for tool in $TOOLS; do
  for target in fast "precise:BESO_PREC_LAYOUT=stacked" "precise:BESO_PREC_LAYOUT=p128" wide_fast wide_precise train "fast:BESO_FAST_CG=2" "fast:BESO_FAST_MC=2"; do
    name=${target%%:*}; envs=${target#*:}; [ "$envs" = "$target" ] && envs=""
    tag=${name}${envs:+_${envs//=/}}
    log=gpurun_out/sanitizer/${tool}_${tag}.log
    env $envs timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $name > $log 2>&1
    echo "$tool $tag: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
