#!/usr/bin/env bash
# Round-2 profiling session (run on the GPU box through gpurun; every program runs once WITHOUT ncu first).
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
python benchmarks/kernels.py k2 > $O/r2_kernels_k2.jsonl 2> $O/r2_kernels_k2.err
# launch list of the bench's timed region (batch 32, 2 steps): per-launch times are cold-cache and serialised -> compare SHARES
python bench.py --batch 32 --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-extras --no-parity-gate > $O/r2_bench_b32_plain.json 2> $O/r2_bench_b32_plain.err && \
$NCU --metrics gpu__time_duration.sum --profile-from-start off -c 4000 --csv --log-file $O/r2_launches_bench_b32.csv \
    python bench.py --batch 32 --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-extras --no-parity-gate > $O/r2_ncu_bench.log 2>&1
cap() {  # name, kernel regex, count, one_kernel arguments...
  local name=$1 rx=$2 cnt=$3; shift 3
  python benchmarks/one_kernel.py "$@" > $O/plain_$name.log 2>&1 && \
  $NCU --set full --import-source on -k "regex:$rx" -c $cnt -f -o $O/$name python benchmarks/one_kernel.py "$@" > $O/ncu_$name.log 2>&1
}
FSD_N=64 cap r2_k1_upscale2x_nhwc_b64 k1_upscale2x 1 k1_nhwc 3
cap r2_k1_sixteenths_c1 k1_sixteenths 1 k1_c1 3
cap r2_k2a_gate_decode "k2_(gate|pose_decode)" 2 k2 2
cap r2_k3_greedynmm_n1024 "k3_merge_kernel" 1 k3 2
cap r2_k3_cluster_n9900 "k3_merge_cluster" 1 k3_big 2
FSD_FRAMES=8 cap r2_k4_crop_stitch_f8 "k4_(crop|stitch)" 2 k4 2
ls -la $O/*.ncu-rep
