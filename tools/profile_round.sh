#!/bin/bash
# Run on the GPU box (via gpurun): plain bench first, then the ncu launch list and one full capture of the top kernels.
# Outputs go to gpurun_out/; tools/summarize_profiles.py turns them into profiles/*.md here.
TAG=${1:-r02}
set -x
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-large-graph --no-sharded --no-cuda-graph"
python bench.py $ARGS > gpurun_out/${TAG}_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c2.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_launch_c2.log 2>&1
python bench.py $ARGS > gpurun_out/${TAG}_plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"knn_gram|knn_rerank|cg_small|cg_cluster|cg_resident|row_gather|graph_weights" -s 18 -c 6 \
    -o gpurun_out/${TAG}_c2_top python bench.py $ARGS > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
if [ -n "$SKIP_C4" ]; then ls -la gpurun_out/${TAG}_*; exit 0; fi  # the CG kernels did not change: keep the previous capture
ARGS4="--steps 2 --warmup 3 --no-cpu-baseline --workload c4 --no-sharded --no-cuda-graph"
python bench.py $ARGS4 > gpurun_out/${TAG}_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"cg_resident" -s 6 -c 2 \
    -o gpurun_out/${TAG}_c4_cg python bench.py $ARGS4 > gpurun_out/${TAG}_ncu_full_c4.log 2>&1
ls -la gpurun_out/${TAG}_*
