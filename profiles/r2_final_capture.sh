set -x
mkdir -p gpurun_out
# (1) launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2f_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r2f_bench_under_ncu.log | cut -c1-200
# (2) full-set capture of one step's own kernels (scatter on), plain and parameter variants
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'norm_2pwl|density_kernel|ct_gradients|ct_eval|bilinear_back|zero_stalled|dbn_2pwl|bin_kernel|head_sum|head_cut|rank_inverse|cost_kernel|order_kernel|realize_kernel|resolve_kernel|final_kernel' -c 40 -o gpurun_out/r2f_step python profiles/run_step.py 1 1000 1 0.3 > gpurun_out/r2f_step.log 2>&1; tail -2 gpurun_out/r2f_step.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'realize_quad' -c 2 -o gpurun_out/r2f_quad python profiles/run_step.py 1 1000 10 0.3 1 > gpurun_out/r2f_quad.log 2>&1; tail -2 gpurun_out/r2f_quad.log
ls -la gpurun_out/*.ncu-rep
# (gpurun merges at most 64 MiB back per call: the two --set full captures were taken in separate calls)
