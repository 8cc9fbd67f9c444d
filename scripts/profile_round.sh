#!/bin/bash
# Round profile on ONE GPU (run under gpurun): launch list + one --set full capture of the three sweep kernels
# of the bench workload and of the dissipative sweeps.  Only CSV / text summaries are kept (the .ncu-rep files are
# deleted: gpurun copies back at most 64 MiB).  Outputs under gpurun_out/; copy the summaries to profiles/.
tag=${1:-r01b}
out=gpurun_out
# (1) launch list of the timed region's kernels (cold-cache, serialised: compare SHARES only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_bench_$tag.log 2>&1
# (2) full set for the x, y, z+epilogue sweeps of the first timed stage (27 sweep launches of warm-up skipped)
ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 27 -c 3 -o /tmp/prof_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
python profiles/ncu_summary.py $out/prof_${tag}_raw.csv > $out/ncu_full_${tag}_summary.txt
for k in sweep_rows sweep_march; do
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null > /tmp/src_$k.csv
  python scripts/ncu_source_mix.py 4194304 < /tmp/src_$k.csv > $out/mix_${tag}_$k.txt
  python scripts/ncu_source_stalls.py 60 < /tmp/src_$k.csv > $out/stalls_${tag}_$k.txt
done
# (3) dissipative sweeps (x, z) at 512^3
ncu --set full --clock-control none -k regex:visc_ -s 6 -c 1 -o /tmp/prof_${tag}_viscx -f \
    python scripts/visc_sweeps.py > $out/ncu_visc_$tag.log 2>&1
ncu --set full --clock-control none -k regex:visc_ -s 20 -c 1 -o /tmp/prof_${tag}_viscz -f \
    python scripts/visc_sweeps.py >> $out/ncu_visc_$tag.log 2>&1
for k in viscx viscz; do
  ncu -i /tmp/prof_${tag}_$k.ncu-rep --page raw --csv > $out/prof_${tag}_${k}_raw.csv 2>/dev/null
  python profiles/ncu_summary.py $out/prof_${tag}_${k}_raw.csv >> $out/ncu_full_${tag}_summary.txt
done
cat $out/ncu_full_${tag}_summary.txt
