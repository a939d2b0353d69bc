# Round-2 evidence run (one GPU): every command first runs plain (exit 0), then under ncu.  Outputs in gpurun_out/;
# tools/summarize_profiles.py r2 turns them into profiles/r2_*.
set -x
timeout 300 python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/bench_nograph_r2.json 2> gpurun_out/bench_nograph_r2.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/ncu_launches_r2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'layer_|wgrad|adamw|sqnorm|pack_images' -s 40 -c 12 -o gpurun_out/prof_train_r2 -f python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/ncu_train_r2.log 2>&1
timeout 200 python tools/prof_field.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_field -s 2 -c 1 -o gpurun_out/prof_field_r2 -f python tools/prof_field.py > gpurun_out/ncu_field_r2.log 2>&1
timeout 200 python tools/prof_block1.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layer_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_block1_r2 -f python tools/prof_block1.py > gpurun_out/ncu_block1_r2.log 2>&1
timeout 200 python tools/prof_predict.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 2 -c 1 -o gpurun_out/prof_fused_r2 -f python tools/prof_predict.py > gpurun_out/ncu_fused_r2.log 2>&1
timeout 300 python bench.py --config 4 --no-graph --steps 6 --warmup 3 > gpurun_out/c4_nograph_r2.json || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_spatial -s 8 -c 2 -o gpurun_out/prof_sparse_r2 -f python bench.py --config 4 --no-graph --steps 6 --warmup 3 > gpurun_out/ncu_sparse_r2.log 2>&1
ls -la gpurun_out/*_r2.ncu-rep
