set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'layer_|wgrad|adamw|sqnorm|pack_images' -s 40 -c 14 -o gpurun_out/prof_train_r1_final python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/ncu_train.log 2>&1
timeout 200 python tools/prof_predict.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 2 -c 1 -o gpurun_out/prof_fused_r1_final python tools/prof_predict.py > gpurun_out/ncu_fused.log 2>&1
PRED_LAYERED=1 timeout 200 python tools/prof_predict.py || exit 1
PRED_LAYERED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:layer_fwd -s 6 -c 3 -o gpurun_out/prof_layered_r1_final python tools/prof_predict.py > gpurun_out/ncu_layered.log 2>&1
ls -la gpurun_out/*.ncu-rep
