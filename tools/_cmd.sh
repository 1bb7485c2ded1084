timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01f_pytest.log 2>&1; tail -3 gpurun_out/r01f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err; cat gpurun_out/r01f_bench_n1.json | cut -c1-300
timeout 300 python tools/graph_trace.py gpurun_out/r01f_graph_trace.txt > gpurun_out/r01f_trace.log 2>&1; tail -3 gpurun_out/r01f_trace.log
