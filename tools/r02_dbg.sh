for d in 0 1 2 4 5; do for sh in "10 256 64 64" "10 128 128 128" "1 256 64 64"; do echo -n "debug=$d: "; FLAIR_CONV_DEBUG=$d python tests/gpu_probes/conv_graph.py $sh; done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
