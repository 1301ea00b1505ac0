python -m pytest tests -m gpu -x -q -k "sharded_two_gpus or repeated_batches or batched_handoff" > gpurun_out/r2_pytest_n2.log 2>&1
tail -3 gpurun_out/r2_pytest_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/h2d_probe_multi.py > gpurun_out/r2_h2d_probe_n2.jsonl 2> gpurun_out/r2_h2d_probe_n2.err
python tools/h2d_probe_multi.py > gpurun_out/r2_h2d_probe_n1.jsonl 2> gpurun_out/r2_h2d_probe_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.err
nvidia-smi topo -m > gpurun_out/r2_topo_n2.log 2>&1; nproc >> gpurun_out/r2_topo_n2.log
