python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/h2d_probe_multi.py > gpurun_out/r2_h2d_probe_n8.jsonl 2> gpurun_out/r2_h2d_probe_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 tools/h2d_probe_multi.py > gpurun_out/r2_h2d_probe_n4.jsonl 2> gpurun_out/r2_h2d_probe_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -2 gpurun_out/r2_bench_n8.err
nvidia-smi topo -m > gpurun_out/r2_topo_n8.log 2>&1; nproc >> gpurun_out/r2_topo_n8.log; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/r2_topo_n8.log
