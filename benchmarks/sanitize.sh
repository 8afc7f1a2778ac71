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
TESTS=(  # one small parametrisation of each kernel (both dtypes / layouts where the kernel is templated on them)
  "tests/test_k1_gather_letterbox_gpu.py::test_slices_match_oracle[dtype0-case4]"     # general TMA kernel, letterbox border, fp16
  "tests/test_k1_gather_letterbox_gpu.py::test_slices_match_oracle[dtype1-case4]"     # ... fp32
  "tests/test_k1_gather_letterbox_gpu.py::test_slices_match_oracle[dtype0-case5]"     # copy: sixteenths path
  "tests/test_k1_gather_letterbox_gpu.py::test_sixteenths_path_equals_general_kernel_and_oracle[True-case4]"
  "tests/test_k1_gather_letterbox_gpu.py::test_exact_2x_fast_path_equals_general_kernel_and_oracle[True-size0]"
  "tests/test_k1_gather_letterbox_gpu.py::test_full_image_pass_matches_oracle[shape6]" # down-scale tap pattern Q = 10
  "tests/test_k2_pose_decode_gpu.py::test_decode_matches_oracle[0.01-True-dtype1]"
  "tests/test_k2_pose_decode_gpu.py::test_decode_matches_oracle[0.5-False-dtype0]"
  "tests/test_k2_pose_decode_gpu.py::test_stage1_nms_and_finalize_match_oracle[shape1]"
  "tests/test_k3_merge_gpu.py::test_hand_made_edge_cases"
  "tests/test_k3_merge_gpu.py::test_stage2_matches_sahi_oracle[box_lex-0.5-IOS-GREEDYNMM]"
  "tests/test_k3_merge_gpu.py::test_stage2_matches_sahi_oracle[index-0.5-IOU-NMM]"
  "tests/test_k3_merge_gpu.py::test_stage2_matches_sahi_oracle[box_lex-0.5-IOU-NMS]"
  "tests/test_k3_merge_gpu.py::test_small_segments_in_a_large_capacity_launch"        # both tiers + the cluster kernel
  "tests/test_k3_merge_gpu.py::test_null_parent_with_pre_cap"
  "tests/test_k4_esrgan_gpu.py::test_batched_crop_equals_per_frame_and_tma_equals_load_store"
  "tests/test_k4_esrgan_gpu.py::test_pre_pad_reflect"
  "tests/test_k4_esrgan_gpu.py::test_round_half_even_and_clamp"
  "tests/test_backbone_kernels_gpu.py::test_bias_act_into_concat_slot_with_residual[silu]"
  "tests/test_backbone_kernels_gpu.py::test_stem_conv_matches_torch[shape2]"
  "tests/test_backbone_kernels_gpu.py::test_pointwise_conv_matches_torch[silu-kn1]"
  "tests/test_backbone_kernels_gpu.py::test_sppf_pool_equals_cascaded_maxpool[hw2]"
  "tests/test_widerface_eval.py::test_device_evaluator_reproduces_reference_evaluator_golden"
)
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck initcheck; do
  log="$OUT/sanitize_${tool}.log"
  extra=""
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout 1500 "$SAN" --tool "$tool" $extra --error-exitcode 77 --print-limit 30 --launch-timeout 0 \
      python -m pytest "${TESTS[@]}" -q -m gpu -p no:cacheprovider > "$log" 2>&1
  rc=$?
  {
    echo "== $tool: exit code $rc (77 = sanitizer reported errors; 124 = timeout)"
    grep -E "passed|failed|error" "$log" | tail -3
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized" "$log" | sort | uniq -c | sort -rn | head -12
  } >> "$OUT/sanitize_summary.txt"
done
cat "$OUT/sanitize_summary.txt"
