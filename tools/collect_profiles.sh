#!/usr/bin/env bash
# Evidence of one round for profiles/: bench lines of every BASELINE config, the ncu launch list of the
# bench command, one `ncu --set full` capture of the streaming kernels. Usage: collect_profiles.sh TAG
tag="${1:-rX}"
out=gpurun_out
mkdir -p $out
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --steps 100 --warmup 5 > $out/${tag}_config$c.json 2> $out/${tag}_config$c.err
  tail -c 300 $out/${tag}_config$c.err
done
timeout 600 python bench.py --steps 300 --warmup 5 > $out/${tag}_config2.json 2> $out/${tag}_config2.err
# launch list (per-launch durations are cold-cache and serialised: the SHARE of a kernel is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-overlap > $out/${tag}_ncu_launches.log 2>&1
# full capture of the streaming kernels and the tails of config 2 (eager launches, one capture each)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"match_lse_fast|detect_bound|classify_kernel|mine_kernel|bwd_patch|detect_refine|detect_nms" --launch-skip 28 --launch-count 14 \
  -f -o $out/prof_${tag}_step python bench.py --steps 2 --warmup 3 --no-graph --no-overlap > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log
