python -m pytest tests -m gpu -x -q -k "reference or nms or gauss or config or paf_to_pose or random_shapes or batched_handoff or coco" 2>&1 | tail -2
python tools/time_configs.py 2>&1 | grep "reference" | cut -c1-250
