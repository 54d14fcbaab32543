O=gpurun_out
python bench.py > $O/r02_bench_default.json 2> $O/bench_err.log; tail -2 $O/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/b_ncu.log 2>&1
cap() {  # name workload batch kernel-id
  ncu --set full --clock-control none --import-source on -k regex:jq_traj_kernel -c 1 -f -o $O/tmp_$1 python tools/ncu_target.py $2 $3 > $O/n_$1.log 2>&1
  python tools/ncu_extract.py $O/tmp_$1.ncu-rep $O/r02_ncu_full_$1.csv $4 $5 > $O/r02_ncu_entry_$1.json 2>> $O/n_$1.log
  ncu -i $O/tmp_$1.ncu-rep --page source --csv > $O/tmp_src.csv 2>/dev/null && python tools/ncu_mix.py $O/tmp_src.csv > $O/r02_ncu_mix_$1.txt
  rm -f $O/tmp_$1.ncu-rep $O/tmp_src.csv
}
cap tile_cnot2 cnot2 16384 16384 4
cap tile_cnot3 cnot3 2368 2368 4
cap fiber_risk_neutral risk_neutral 4096 36864 3
cap fiber_cnot1 cnot1 32768 32768 3
cap latency_cnot2_single cnot2 1 1 5
python tools/bench_all.py r02 > $O/bench_all.log 2>&1
python tools/generic_bench.py > $O/r02_generic.md 2>&1
python tools/single_eval_latency.py > $O/r02_latency.md 2>&1
du -sh $O; echo done
