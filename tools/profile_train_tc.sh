# ncu capture of the fused TF32 layer-step kernel of the training forward pass (one launch of each layer's instance)
O=gpurun_out
CMD="python bench.py --workload train64tf32 --no-cpu-baseline --steps 1 --warmup 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gru_fwd_tc_kernel -s 70 -c 1 -f -o $O/prof_train_tc $CMD > $O/r02d_ncu_train_tc.log 2>&1
python tools/ncu_summary.py $O/prof_train_tc.ncu-rep > $O/r02d_train_tc_ncu_summary.txt 2>&1
ncu -i $O/prof_train_tc.ncu-rep --page source --csv > $O/src_tc.csv 2>/dev/null && python tools/ncu_hot.py $O/src_tc.csv 40 > $O/r02d_train_tc_hot.txt 2>&1
rm -f $O/src_tc.csv $O/prof_train_tc.ncu-rep
cat $O/r02d_train_tc_ncu_summary.txt
head -70 $O/r02d_train_tc_hot.txt
