set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/r02_bench_large_64x20s.json 2> gpurun_out/r02_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
timeout 300 python tests/parity_report.py > gpurun_out/r02_parity_report.txt 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload base_32x15s --no-cpu-baseline > gpurun_out/r02_bench_base_32x15s.json 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload large_64x30s --no-cpu-baseline --no-incremental > gpurun_out/r02_bench_large_64x30s.json 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload stream_large_b1 --steps 10 > gpurun_out/r02_stream_large_b1_cluster.json 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload stream_large_b1 --steps 10 --step-impl 1 --no-cpu-baseline > gpurun_out/r02_stream_large_b1.json 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload stream_large_b16 --steps 10 > gpurun_out/r02_stream_large_b16.json 2>> gpurun_out/r02_bench.err
timeout 200 python tools/cluster_trace.py 20 > gpurun_out/r02_stream_cluster_trace.txt 2>> gpurun_out/r02_bench.err
timeout 200 python tools/stream_kernel_probe.py 1 > gpurun_out/r02_stream_b1_kernels.txt 2>> gpurun_out/r02_bench.err
timeout 200 python tools/stream_kernel_probe.py 16 > gpurun_out/r02_stream_b16_kernels.txt 2>> gpurun_out/r02_bench.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-incremental > gpurun_out/ncu_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/traffic.csv large_64x20s > gpurun_out/r02_traffic_large_64x20s.json 2>> gpurun_out/r02_bench.err
tail -3 gpurun_out/r02_bench.err
