mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
ICL_DISABLE_BIGW=0 $NCU -k regex:bigw_gemm_k -s 2 -c 1 -o gpurun_out/r02p_bigw_fwd_rows16 python tools/skinny_bench.py fwd 16 > /dev/null 2>&1
$NCU -k regex:bigw_gemm_k -s 2 -c 1 -o gpurun_out/r02p_bigw_fwd_rows128 python tools/skinny_bench.py fwd 128 > /dev/null 2>&1
$NCU -k regex:bigw_gemm_k -s 2 -c 1 -o gpurun_out/r02p_bigw_dgrad_rows16 python tools/skinny_bench.py dgrad 16 > /dev/null 2>&1
$NCU -k regex:sgd_factored_umma_k -s 2 -c 1 -o gpurun_out/r02p_sgd_factored_R32 python tools/skinny_bench.py sgd 16 32 > /dev/null 2>&1
$NCU -k regex:sgd_factored_umma_k -s 2 -c 1 -o gpurun_out/r02p_sgd_factored_R1024 python tools/skinny_bench.py sgd 16 1024 > /dev/null 2>&1
$NCU -k regex:class_stats_row -s 4 -c 4 -o gpurun_out/r02p_class_stats_row_k16 python tools/loss_bench.py 16 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-profile --no-cpu-baseline --no-also > gpurun_out/r02p_bench_under_ncu.log 2>&1
ls -la gpurun_out/r02p_*
