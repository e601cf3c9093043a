#!/bin/bash
# Round evidence on ONE GPU: full GPU test suite, the default bench line, the reference arm, the ncu launch list and
# full captures (each only after the plain command has exited 0).  Everything lands in gpurun_out/ev_*; the summaries
# that are judged are copied into profiles/ by hand (profiles/README.md lists them).
set -u
O=gpurun_out
B12="python bench.py --seconds 12 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/ev_pytest_gpu.log 2>&1; tail -2 $O/ev_pytest_gpu.log
timeout 600 python bench.py > $O/ev_bench_1gpu.json 2> $O/ev_bench_1gpu.err || tail -5 $O/ev_bench_1gpu.err
timeout 600 python bench.py --impl reference > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err || tail -5 $O/ev_bench_reference.err
if timeout 300 $B12 > $O/ev_b12.json 2> $O/ev_b12.err; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(channelize|tail|head|state|pre|scan|seq|fix|fir_fft64r|fir_repair|fir_rows|mix|forward5|mac5)" -c 400 --csv --log-file $O/ev_launches.csv $B12 > $O/ev_ncu_l.log 2>&1
  timeout 500 $NCU -k regex:^k_channelize5 -o $O/ev_k_channelize5 $B12 > $O/ev_ncu_c5.log 2>&1
  timeout 500 $NCU -k regex:^k_tail_fused -o $O/ev_k_tail_fused $B12 > $O/ev_ncu_tail.log 2>&1
fi
C3="python tools/bench_cfg3.py --seconds 4 --child"
if timeout 300 python tools/bench_cfg3.py --seconds 10 > $O/ev_cfg3.json 2> $O/ev_cfg3.err; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(channelize|tail|head|state|pre|scan|seq|fix|fir_fft64r|fir_repair|fir_rows|mix|forward5|mac5)" -c 400 --csv --log-file $O/ev_cfg3_launches.csv python tools/bench_cfg3.py --seconds 10 --child > $O/ev_cfg3_ncu_l.log 2>&1
  for k in k_fir_fft64r k_fir_repair k_seq_dc k_seq_agc k_mix_exact; do
    timeout 400 $NCU -k regex:^$k -o $O/ev_$k $C3 > $O/ev_ncu_$k.log 2>&1
  done
fi
if timeout 300 python tools/bench_other.py cfg5_c256 > $O/ev_c256.json 2> $O/ev_c256.err; then
  for k in k_forward5 k_mac5; do
    timeout 400 $NCU -k regex:^$k -o $O/ev_$k python tools/bench_other.py cfg5_c256 --reps 1 > $O/ev_ncu_$k.log 2>&1
  done
fi
# reports -> csv pages; only the channel bank's report itself is kept (64 MiB limit on what comes back)
for r in $O/ev_k_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv > ${b}_src.csv 2>/dev/null
  case $r in *k_channelize5*) ;; *) rm -f $r ;; esac
done
timeout 300 python tools/bench_other.py cfg1 > $O/ev_cfg1.json 2> $O/ev_cfg1.err
[ -x tools/tma_rows_ubench ] && timeout 120 tools/tma_rows_ubench > $O/ev_tma_rows_ubench.txt 2>&1
[ -x tools/tma_issue_probe ] && timeout 120 tools/tma_issue_probe > $O/ev_tma_issue_probe.txt 2>&1
timeout 600 python tools/bench_pipeline_file.py > $O/ev_pipeline_file.json 2> $O/ev_pipeline_file.err
ls -la $O | grep ev_ | wc -l
