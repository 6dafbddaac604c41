#!/usr/bin/env bash
# ncu --set full captures of the eval and mesh workloads' kernels, summarised ON the box (the reports themselves exceed
# what gpurun copies back): gpurun_out/<tag>_{eval,mesh}_ncu_full.md + ncu_traffic_{eval,mesh}.json
O=gpurun_out; mkdir -p $O; T=${1:-inf}
CMD_E="python bench.py --workload eval --steps 1 --warmup 1"
timeout 900 ncu --set full --clock-control none -k regex:"k_sdf_tc2|k_knn_points|k_knn_slots|k_sampler_iter|k_color_fwd_tc2|k_head_fwd_tc2" -s 40 -c 20 -f -o /tmp/${T}_eval $CMD_E > $O/${T}_eval_ncu.log 2>&1; echo "eval rc=$?"
python tools/summarise_ncu_full.py /tmp/${T}_eval.ncu-rep "$CMD_E" --traffic-json $O/ncu_traffic_eval.json > $O/${T}_eval_ncu_full.md; wc -l $O/${T}_eval_ncu_full.md
CMD_M="python bench.py --workload mesh --steps 1 --warmup 1"
timeout 900 ncu --set full --clock-control none -k regex:"k_sdf_tc2|k_knn_points|k_grid_points_mask|k_scatter_f32" -s 10 -c 12 -f -o /tmp/${T}_mesh $CMD_M > $O/${T}_mesh_ncu.log 2>&1; echo "mesh rc=$?"
python tools/summarise_ncu_full.py /tmp/${T}_mesh.ncu-rep "$CMD_M" --traffic-json $O/ncu_traffic_mesh.json > $O/${T}_mesh_ncu_full.md; wc -l $O/${T}_mesh_ncu_full.md
