#!/bin/bash
# One parameterised evidence script (run on the GPU box through gpurun; outputs land in gpurun_out/, copy what is to be
# judged into profiles/).  Replaces the per-run one-off scripts of round 1.
#   tools/gpu_evidence.sh tests [TAG]          pytest -m gpu + smoke
#   tools/gpu_evidence.sh bench [TAG] [ARGS]   python bench.py ARGS
#   tools/gpu_evidence.sh launches [TAG] CMD   ncu launch list (gpu__time_duration.sum) of CMD
#   tools/gpu_evidence.sh ncufull [TAG] KERNEL_REGEX CMD   one ncu --set full capture of the first matching launch
#   tools/gpu_evidence.sh sanitize [TAG]       compute-sanitizer memcheck + racecheck + synccheck of tools/sanitize_smoke.py
set -u
mode=${1:-tests}; tag=${2:-r02}; shift 2 || true
out=gpurun_out; mkdir -p $out
case $mode in
  tests)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $out/${tag}_pytest_gpu.log; cat $out/${tag}_pytest_gpu.log
    python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/${tag}_smoke.log ;;
  bench)
    python bench.py "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.err; head -c 1500 $out/${tag}_bench.json ;;
  launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv "$@" > $out/${tag}_under_ncu.log 2>&1
    python tools/launch_summary.py $out/${tag}_launches.csv | tee $out/${tag}_launch_summary.txt ;;
  ncufull)
    regex=$1; shift
    ncu --set full --clock-control none --import-source on -k "regex:$regex" -c 1 -o $out/${tag}_ncu_full -f "$@" > $out/${tag}_ncufull.log 2>&1
    ncu -i $out/${tag}_ncu_full.ncu-rep --page details --csv > $out/${tag}_ncu_full_details.csv 2>/dev/null
    ncu -i $out/${tag}_ncu_full.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw.csv 2>/dev/null
    tail -5 $out/${tag}_ncufull.log ;;
  sanitize)
    for tool in memcheck racecheck synccheck; do
      for part in flat quant misc train; do
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $part > $out/${tag}_sanitize_${tool}_${part}.log 2>&1
        echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_sanitize_${tool}_${part}.log | tail -1)"
      done
    done | tee $out/${tag}_sanitize_summary.txt ;;
esac
