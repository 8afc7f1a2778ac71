#!/usr/bin/env bash
# compute-sanitizer over one small parametrisation of every kernel's GPU parity test.
#   memcheck  — out-of-bounds / misaligned global + shared accesses (incl. the TMA boxes and bulk stores of K1 / K4)
#   racecheck — shared-memory hazards: K1's double-buffered TMA tile marching, K3's named-barrier phases and double-buffered
#               match bits, K4's TMA staging, K7's cp.async tiles (cluster-scope / global-memory races are outside its model)
#   initcheck — reads of device memory nobody wrote (candidate rows, workspaces, concat buffers)
# Usage (on the GPU box):  bash benchmarks/sanitize.sh [outdir]      -> <outdir>/sanitize_<tool>.log + sanitize_summary.txt
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
TESTS=(
  "tests/test_k1_gather_letterbox_gpu.py::test_slices_match_oracle"                    # -k below keeps the two small geometries
  "tests/test_k1_gather_letterbox_gpu.py::test_sixteenths_path_equals_general_kernel_and_oracle"
  "tests/test_k2_pose_decode_gpu.py::test_decode_matches_oracle"
  "tests/test_k2_pose_decode_gpu.py::test_stage1_nms_and_finalize_match_oracle"
  "tests/test_k3_merge_gpu.py::test_hand_made_edge_cases"
  "tests/test_k3_merge_gpu.py::test_stage2_matches_sahi_oracle"
  "tests/test_k3_merge_gpu.py::test_small_segments_in_a_large_capacity_launch"
  "tests/test_k3_merge_gpu.py::test_null_parent_with_pre_cap"
  "tests/test_k4_esrgan_gpu.py::test_batched_crop_equals_per_frame_and_tma_equals_load_store"
  "tests/test_k4_esrgan_gpu.py::test_pre_pad_reflect"
  "tests/test_backbone_kernels_gpu.py::test_bias_act_into_concat_slot_with_residual"
  "tests/test_backbone_kernels_gpu.py::test_stem_conv_matches_torch"
  "tests/test_backbone_kernels_gpu.py::test_pointwise_conv_matches_torch"
  "tests/test_backbone_kernels_gpu.py::test_sppf_pool_equals_cascaded_maxpool"
  "tests/test_widerface_eval.py::test_device_evaluator_reproduces_reference_evaluator_golden"
)
# one parametrisation each: small shapes, both dtypes where the kernel is templated on it
KEXPR="(case4 or case5 or not case) and (not 2500) and (not 5000) and (GREEDYNMM-IOS-0.5-box_lex or not test_stage2_matches) and (shape2 or not test_stem) and (kn1 or not test_pointwise) and (hw2 or not test_sppf)"
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck initcheck; do
  log="$OUT/sanitize_${tool}.log"
  extra=""
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout 1500 "$SAN" --tool "$tool" $extra --error-exitcode 77 --print-limit 30 --launch-timeout 0 \
      python -m pytest "${TESTS[@]}" -q -m gpu -k "$KEXPR" -p no:cacheprovider > "$log" 2>&1
  rc=$?
  {
    echo "== $tool: exit code $rc (77 = sanitizer reported errors; 124 = timeout)"
    grep -E "passed|failed|error" "$log" | tail -3
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized" "$log" | sort | uniq -c | sort -rn | head -12
  } >> "$OUT/sanitize_summary.txt"
done
cat "$OUT/sanitize_summary.txt"
